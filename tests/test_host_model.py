"""The product's __host__ __device__ model headers (quad_model.cuh, hessian_local.cuh, pid.cuh), compiled
with g++ (tests/host_check), against the oracle: float path, hyper-dual first/second derivatives, PID."""
import ctypes

import numpy as np

from oracle import oracle_np as o
from oracle.oracle_np import Jet

FP = ctypes.POINTER(ctypes.c_float)


def P(a):
    return a.ctypes.data_as(FP)


def _envp(p):
    return np.array([p.m, p.g, p.max_thrust, p.dt, p.alpha_bodyrate, p.action_scale, p.pos_limit, *p.max_omega,
                     p.max_steps_in_episode], dtype=np.float32)


def _mk(z, fd, pt, vt):
    zz = np.zeros((10, 3))
    return o.QuadState(pos=z[0:3], quat=z[3:7], vel=z[7:10], omega=z[10:13], f_disturb=[float(v) for v in fd], time=5,
                       pos_tar=[float(v) for v in pt], vel_tar=[float(v) for v in vt], pos_traj=zz, vel_traj=zz)


def test_model_headers_match_oracle(host_check_lib):
    lib = host_check_lib
    p = o.EnvParams()
    envp = _envp(p)
    rng = np.random.default_rng(1)
    iu = np.triu_indices(17)
    for trial in range(25):
        x = np.zeros(13, np.float32)
        x[0:3] = rng.normal(0, .5, 3)
        q = rng.normal(0, 1, 4) + np.array([0, 0, 0, 3.])
        x[3:7] = q / np.linalg.norm(q) * (1 + 0.01 * rng.normal())
        x[7:10] = rng.normal(0, 1, 3)
        x[10:13] = rng.normal(0, 2, 3)
        u = rng.uniform(-1.3, 1.3, 4).astype(np.float32)
        if trial % 3 == 0:
            u[1], u[0] = 1.0, -1.0  # clip ties
        fd = rng.normal(0, .1, 3).astype(np.float32)
        pt = rng.normal(0, .5, 3).astype(np.float32)
        vt = rng.normal(0, .5, 3).astype(np.float32)
        xn = np.zeros(13, np.float32)
        r, d = ctypes.c_float(), ctypes.c_int()
        lib.hc_step(P(envp), P(x), P(u), P(fd), P(pt), P(vt), 5, P(xn), ctypes.byref(r), ctypes.byref(d))
        s = _mk([float(v) for v in x], fd, pt, vt)
        rr = o.tracking_penyaw_reward(s)
        nx = o.step_env(s, [float(v) for v in u], p)
        xo = np.array(list(nx.pos) + list(nx.quat) + list(nx.vel) + list(nx.omega))
        assert np.abs(xn - xo).max() < 2e-6 and abs(r.value - rr) < 2e-6 and d.value == 0
        eye = np.eye(17)
        z = [Jet(float(v), eye[i].copy()) for i, v in enumerate(list(x) + list(u))]
        sj = _mk(z, fd, pt, vt)
        c = -1.0 * o.tracking_penyaw_reward(sj)
        nxj = o.step_env(sj, z[13:17], p)
        F = list(nxj.pos) + list(nxj.quat) + list(nxj.vel) + list(nxj.omega) + [c]
        Go = np.stack([f.g if f.g is not None else np.zeros(17) for f in F])
        To = np.stack([(f.h if f.h is not None else np.zeros((17, 17)))[iu] for f in F])
        G = np.zeros((14, 17), np.float32)
        T = np.zeros((14, 153), np.float32)
        lib.hc_local(P(envp), P(x), P(u), P(fd), P(pt), P(vt), P(G), P(T))
        assert (np.abs(G - Go) / (1 + np.abs(Go))).max() < 2e-6
        assert (np.abs(T - To) / (1 + np.abs(To))).max() < 5e-6
        act = np.zeros(4, np.float32)
        at = rng.normal(0, .1, 3).astype(np.float32)
        lib.hc_pid(P(envp), P(x), P(pt), P(vt), P(at), P(act))
        ao = o.pid_action(s, p, acc_tar=at.astype(np.float64))
        assert np.abs(act - ao).max() / (1 + np.abs(ao).max()) < 5e-6


def test_termination_flags(host_check_lib):
    lib = host_check_lib
    p = o.EnvParams()
    envp = _envp(p)
    x = np.zeros(13, np.float32)
    x[6] = 1
    u = np.zeros(4, np.float32)
    z3 = np.zeros(3, np.float32)
    xn = np.zeros(13, np.float32)
    r, d = ctypes.c_float(), ctypes.c_int()
    for time, pos1, exp in ((299, 0.0, 0), (300, 0.0, 1), (0, 3.0, 0), (0, 3.0001, 1), (0, -3.5, 1)):
        x[1] = pos1
        lib.hc_step(P(envp), P(x), P(u), P(z3), P(z3), P(z3), time, P(xn), ctypes.byref(r), ctypes.byref(d))
        assert d.value == exp


def test_result_artefacts_match_the_reference_formats(tmp_path):
    """eval_err_pos_{name}.pkl and state_seq_{name}.pkl (envs/quadrotor.py:581-591, 655-666): host-side only."""
    import os
    import pickle

    from covo_mpc_b200 import harness as hz  # loads libcovo_b200.so (no GPU needed until a handle is created)

    p1 = hz.save_eval_results(np.array([0.05, 0.06]), "covo_online", str(tmp_path))
    assert os.path.basename(p1) == "eval_err_pos_covo_online.pkl"
    assert np.allclose(pickle.load(open(p1, "rb")), [0.05, 0.06])
    seq = [{"pos": np.zeros(3), "quat": np.array([0, 0, 0, 1.0]), "pos_tar": np.zeros(3), "f_disturb": np.zeros(3), "pos_traj": np.zeros((320, 3))}]
    p2 = hz.save_state_seq(seq, "", str(tmp_path))
    assert os.path.basename(p2) == "state_seq_.pkl"  # scripts/vis.py:71 loads exactly this name by default
    back = pickle.load(open(p2, "rb"))
    assert isinstance(back, list) and set(["pos", "quat", "pos_tar", "f_disturb", "pos_traj"]) <= set(back[0])
