"""The product's __host__ __device__ model headers (quad_model.cuh, hessian_local.cuh, pid.cuh), compiled
with g++ (tests/host_check), against the oracle: float path, hyper-dual first/second derivatives, PID."""
import ctypes

import numpy as np
import pytest

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


def test_plugin_call_marshalling_fast_path():
    """The per-step plugin call reuses staging buffers and cached pointers (no per-call allocations): with a recording stand-in
    for the library, what covo_step receives through Handle.step_state / Handle.step(eps=None) must equal what the generic
    path (fresh arrays per call) hands over, for float32 and float64 / list inputs alike, and the action must come back as a
    fresh array every time."""
    import ctypes as C

    import covo_mpc_b200 as cm
    from covo_mpc_b200 import _lib

    calls = []

    class RecordingLib:
        def covo_step(self, h, sp, tp, ep, op):
            s = np.ctypeslib.as_array(sp, shape=(24,)).copy()
            t = int(tp[0])
            calls.append((s, t, ep is None))
            for k in range(4):
                op[k] = 0.25 * (k + 1) + t
            return 0

        def __getattr__(self, name):
            return lambda *a: 0

    h = object.__new__(_lib.Handle)
    h.lib, h._h, h.E, h.H, h.n, h.n_local = RecordingLib(), C.c_void_p(1), 1, 8, 32, 64
    env = cm.Quad3D("tracking_zigzag")
    _, info, st = env.reset(np.random.default_rng(3))
    ns = info["noisy_state"].replace(time=17)
    a1 = h.step_state(ns)
    a2 = h.step(ns.to_state24(), [ns.time])[0]
    a3 = h.step(ns.to_state24().astype(np.float64).tolist(), np.array([ns.time], np.int64))[0]
    a4 = h.step(ns.to_state24(), [ns.time], eps=np.zeros((1, 64, 32), np.float32))[0]
    ref = ns.to_state24()
    assert len(calls) == 4
    for s, t, no_eps in calls:
        assert np.array_equal(s, ref) and t == 17
    assert [c[2] for c in calls] == [True, True, True, False]
    for a in (a1, a2, a3, a4):
        assert a.dtype == np.float32 and np.allclose(a, [17.25, 17.5, 17.75, 18.0])
    assert a1 is not a2 and not np.shares_memory(a1, a2)  # callers may keep actions: never a view of the staging buffer
    with pytest.raises(ValueError):
        h.step(np.zeros(23, np.float32), [0])
    with pytest.raises(ValueError):
        h.step(np.zeros(24, np.float32), [0, 1])
    # the controller's functional params update without the dataclass constructor keeps every field
    ctl, cp = cm.get_controller(env, "covo-online", "N64_H8_lam0.01")
    ctl._handle, ctl._cfg.traj_len = h, ns.pos_traj.shape[0]
    h.cfg = _lib.CovoConfig()
    h.cfg.traj_len = ns.pos_traj.shape[0]
    action, cp2, info2 = ctl(None, st, env.default_params, None, cp, {"noisy_state": ns})
    assert np.allclose(action, [17.25, 17.5, 17.75, 18.0]) and info2 is None
    assert type(cp2) is type(cp) and cp2 is not cp
    assert (cp2.gamma_mean, cp2.gamma_sigma, cp2.discount, cp2.sample_sigma) == (cp.gamma_mean, cp.gamma_sigma, cp.discount, cp.sample_sigma)
    assert cp2._gen == ctl._generation and np.array_equal(np.asarray(cp.a_mean), np.tile(np.array([cp.a_mean[0][0], 0, 0, 0], np.float32), (8, 1)))
    cp3 = cp2.replace(a_mean=np.zeros((8, 4), np.float32))  # the reference-style replace still works on the returned object
    assert cp3.a_mean.shape == (8, 4) and cp3.sample_sigma == cp.sample_sigma


def test_fast_path_pointers_are_accepted_by_the_real_library():
    """The cached ctypes pointers go through the real library's argtypes: with a NULL handle covo_step must come back with
    its own 'null argument' status (no GPU involved), not a ctypes conversion error."""
    import covo_mpc_b200 as cm
    from covo_mpc_b200 import _lib

    h = object.__new__(_lib.Handle)
    h.lib, h._h, h.E, h.H, h.n, h.n_local = _lib.load(), None, 1, 8, 32, 64
    env = cm.Quad3D("hovering")
    _, info, st = env.reset(np.random.default_rng(0))
    with pytest.raises(ValueError, match="null argument"):
        h.step_state(info["noisy_state"])
    with pytest.raises(ValueError, match="null argument"):
        h.step(st.to_state24(), [0])
    h._h = None  # keep __del__ from touching the library


def _standin_handle(cm, _lib, calls, H=8, N=64):
    """A Handle whose library is a recording stand-in (no GPU): enough surface for the controllers' host logic."""
    import ctypes as C

    class Lib:
        def covo_step(self, h, sp, tp, ep, op):
            calls.append(("step",))
            for k in range(4):
                op[k] = 0.5
            return 0

        def covo_set_mean(self, h, p):
            calls.append(("set_mean", np.ctypeslib.as_array(p, shape=(4 * H,)).copy()))
            return 0

        def covo_get_mean(self, h, p):
            calls.append(("get_mean",))
            np.ctypeslib.as_array(p, shape=(4 * H,))[:] = 7.0
            return 0

        def covo_set_env_params(self, h, m, g, mt, dt, al, sc, mo, ms):
            calls.append(("set_env_params", float(m), float(mt), int(ms), [float(mo[k]) for k in range(3)]))
            return 0

        def __getattr__(self, name):
            def f(*a):
                calls.append((name.replace("covo_", ""),))
                return 0
            return f

    h = object.__new__(_lib.Handle)
    h.lib, h._h, h.E, h.H, h.n, h.n_local = Lib(), C.c_void_p(1), 1, H, 4 * H, N
    h.cfg = _lib.CovoConfig()
    return h


def test_offline_reset_with_current_params_keeps_the_resident_mean():
    """render_env's pattern (envs/quadrotor.py:637-639): after `done`, controller.reset is called with the CURRENT control_params.
    The reference keeps a_mean across that reset (covo.py:101-104); here the returned object must stay usable -- no 'stale
    control_params' -- and must not trigger a re-upload of the mean."""
    import covo_mpc_b200 as cm
    from covo_mpc_b200 import _lib

    calls = []
    env = cm.Quad3D("tracking_zigzag")
    _, info, st = env.reset(np.random.default_rng(5))
    ns = info["noisy_state"]
    ctl, cp = cm.get_controller(env, "covo-offline", "N64_H8_lam0.01")
    h = _standin_handle(cm, _lib, calls)
    ctl._handle, ctl._cfg.traj_len = h, ns.pos_traj.shape[0]
    h.cfg.traj_len = ns.pos_traj.shape[0]
    cp = ctl.reset(st, env.default_params, cp, None)
    _, cp, _ = ctl(None, st, env.default_params, None, cp, {"noisy_state": ns})
    assert [c[0] for c in calls].count("set_mean") == 1  # the initial (host) mean, uploaded once
    _, cp, _ = ctl(None, st, env.default_params, None, cp, {"noisy_state": ns})
    cp_r = ctl.reset(st, env.default_params, cp, None)  # reset with the params of the previous step
    assert cp_r._gen == cp._gen == ctl._generation
    n_set = [c[0] for c in calls].count("set_mean")
    action, cp2, _ = ctl(None, st, env.default_params, None, cp_r, {"noisy_state": ns})  # raised RuntimeError before the fix
    assert [c[0] for c in calls].count("set_mean") == n_set and np.allclose(action, 0.5)
    assert np.allclose(np.asarray(cp2.a_mean), 7.0)  # materialises from the (stand-in) device
    # a reset with the INITIAL params re-uploads the hover mean, as the reference's eval_env does (quadrotor.py:548-550)
    cp_i = ctl.reset(st, env.default_params, ctl.init_control_params, None)
    ctl(None, st, env.default_params, None, cp_i, {"noisy_state": ns})
    assert [c[0] for c in calls].count("set_mean") == n_set + 1


def test_call_time_env_params_reach_the_handle():
    """The reference plans with the env_params of the call (covo.py:187-283): a modified mass must be forwarded, once per object."""
    import covo_mpc_b200 as cm
    from covo_mpc_b200 import _lib

    calls = []
    env = cm.Quad3D("tracking_zigzag")
    _, info, st = env.reset(np.random.default_rng(6))
    ns = info["noisy_state"]
    ctl, cp = cm.get_controller(env, "mppi", "N64_H8_lam0.01")
    h = _standin_handle(cm, _lib, calls)
    ctl._handle, ctl._cfg.traj_len = h, ns.pos_traj.shape[0]
    h.cfg.traj_len = ns.pos_traj.shape[0]
    _, cp, _ = ctl(None, st, env.default_params, None, cp, {"noisy_state": ns})
    heavy = env.default_params.replace(m=0.04)
    _, cp, _ = ctl(None, st, heavy, None, cp, {"noisy_state": ns})
    _, cp, _ = ctl(None, st, heavy, None, cp, {"noisy_state": ns})
    sets = [c for c in calls if c[0] == "set_env_params"]
    assert len(sets) == 2 and abs(sets[0][1] - 0.027) < 1e-9 and abs(sets[1][1] - 0.04) < 1e-9 and sets[1][3] == 300
    assert cm.get_controller(env, "covo_offline_online")[0].mode == "online"  # "online" is tested first (quadrotor.py:731-737)
