"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the float64-Hessian oracle).
CPU: the oracle still reproduces them.  GPU: the CUDA path reproduces them through the C-ABI."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle_np as o

HERE = os.path.dirname(os.path.abspath(__file__))
# reference_*.npz are the vectors produced by the reference's own source; tests/test_reference_golden.py reads those
FILES = sorted(f for f in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(f).startswith("reference_"))


def _state(g):
    s = g["state24"]
    return o.make_state(s[0:3], s[3:7], s[7:10], s[10:13], s[13:16], int(g["time"]), g["pos_traj"], g["vel_traj"], s[16:19],
                        s[19:22], dtype=np.float32)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    p = o.EnvParams()
    ns = _state(g)
    H = g["a_mean"].shape[0]
    if "mppi" in os.path.basename(path):
        u, new_mean, _, _, dbg = o.mppi_call(ns, g["a_mean"], g["a_cov"], g["eps"].reshape(-1, H, 4), p, lam=0.01, return_debug=True)
    else:
        u, new_mean, a_cov, _, dbg = o.covo_call(ns, g["a_mean"], g["eps"], p, lam=0.01, return_debug=True)
        assert np.abs(dbg["R"] - g["R"]).max() <= 1e-9 * max(1, np.abs(g["R"]).max())
        assert np.abs(a_cov - g["a_cov"]).max() < 1e-6
    assert np.abs(dbg["cost"] - g["cost"]).max() < 1e-5
    assert np.abs(new_mean - g["a_mean_new"]).max() < 1e-6 and np.abs(u - g["action"]).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_path_reproduces_golden(path):
    from covo_mpc_b200 import _lib

    g = np.load(path)
    H = g["a_mean"].shape[0]
    N = g["eps"].shape[0]
    mppi = "mppi" in os.path.basename(path)
    cfg = _lib.default_config()
    cfg.mode = _lib.MODE_MPPI if mppi else _lib.MODE_COVO_ONLINE
    cfg.n_samples, cfg.horizon, cfg.traj_len = N, H, g["pos_traj"].shape[0]
    h = _lib.Handle(cfg)
    h.set_reference(g["pos_traj"][None], g["vel_traj"][None])
    h.set_mean(g["a_mean"][None])
    if mppi:
        h.set_cov(g["a_cov"][None])
    else:
        R = h.hessian(g["state24"], [int(g["time"])], g["a_mean"][None], shift=True)[0]
        assert np.abs(R - g["R"]).max() < 2e-5 * max(1, np.abs(g["R"]).max())
    act = h.step(g["state24"], [int(g["time"])], g["eps"][None])[0]
    assert np.abs(act - g["action"]).max() < 2e-4
    assert np.abs(h.get_mean()[0] - g["a_mean_new"]).max() < 2e-4
    if not mppi:
        cov = h.get_cov()[0]
        assert np.linalg.norm(cov - g["a_cov"]) / np.linalg.norm(g["a_cov"]) < 2e-5


# ---- vectors produced by executing the reference's own source (tests/golden/make_reference_golden.py, sections 6-8) -------
REF_CALLS = ["reference_call_covo_online_lam0.01", "reference_call_covo_online_lam1.0", "reference_call_mppi"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", REF_CALLS)
def test_cuda_path_reproduces_reference_call(name):
    """One whole controller call through the C-ABI vs the same call executed from /root/reference (NumPy shim, logged draws)."""
    from covo_mpc_b200 import _lib

    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    mppi = "mppi" in name
    N, H = int(g["N"]), int(g["H"])
    cfg = _lib.default_config()
    cfg.mode = _lib.MODE_MPPI if mppi else _lib.MODE_COVO_ONLINE
    cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.lam = N, H, g["pos_traj"].shape[0], float(g["lam"])
    h = _lib.Handle(cfg)
    h.set_reference(g["pos_traj"][None], g["vel_traj"][None])
    h.set_mean(g["a_mean"][None])
    if mppi:
        h.set_cov(g["a_cov"][None])
    else:
        R = h.hessian(g["state24"], [int(g["time"])], g["a_mean"][None], shift=True)[0]
        assert np.abs(R - g["R"]).max() < 2e-5 * np.abs(g["R"]).max()
    act = h.step(g["state24"], [int(g["time"])], g["eps"].reshape(1, N, 4 * H))[0]
    tol = 2e-4 if float(g["lam"]) < 0.1 else 5e-5
    assert np.abs(act - g["action"]).max() < tol
    assert np.abs(h.get_mean()[0] - g["a_mean_new"]).max() < tol
    if not mppi:
        cov = h.get_cov()[0]
        assert np.linalg.norm(cov - g["a_cov"]) / np.linalg.norm(g["a_cov"]) < 5e-5


@pytest.mark.gpu
def test_cuda_offline_schedule_reproduces_reference():
    """covo_reset_offline vs reset_a_cov_offline executed from /root/reference (first 3 table entries)."""
    from covo_mpc_b200 import _lib

    g = np.load(os.path.join(HERE, "golden", "reference_covo_offline_schedule.npz"))
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len = _lib.MODE_COVO_OFFLINE, 64, int(g["H"]), g["pos_traj"].shape[0]
    h = _lib.Handle(cfg)
    h.set_reference(g["pos_traj"][None], g["vel_traj"][None])
    h.reset_offline(g["state24"], [int(g["time"])], 3)
    tab = h.get_cov_offline(3)
    for k in range(3):
        assert np.linalg.norm(tab[k] - g["a_cov_offline"][k]) / np.linalg.norm(g["a_cov_offline"][k]) < 1e-4, k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["covo-online", "mppi"])
def test_same_key_in_same_action_out(name):
    """The drop-in claim end to end: the reference's controller call, executed from its own source with rng_act = a PRNGKey,
    vs the product's controller object handed the SAME key -- sampling happens in the kernel (Threefry twin of jax.random)."""
    import covo_mpc_b200 as cm

    g = np.load(os.path.join(HERE, "golden", f"reference_call_{name.replace('-', '_')}_keyed.npz"))
    N, H = int(g["N"]), int(g["H"])
    env = cm.Quad3D("tracking_zigzag")
    ctl, cp = cm.get_controller(env, name, f"N{N}_H{H}_lam{float(g['lam'])}")
    s = g["state24"]
    z3 = np.zeros(3, np.float32)
    st = cm.EnvState3D(pos=s[0:3], quat=s[3:7], vel=s[7:10], omega=s[10:13], f_disturb=s[13:16], pos_tar=s[16:19], vel_tar=s[19:22],
                       acc_tar=z3, pos_traj=g["pos_traj"], vel_traj=g["vel_traj"], acc_traj=np.zeros_like(g["pos_traj"]), time=int(g["time"]))
    action, cp2, _ = ctl(None, st, env.default_params, g["rng_act"], cp.replace(a_mean=g["a_mean"]), {"noisy_state": st})
    assert np.abs(np.asarray(action) - g["action"]).max() < 2e-4
    assert np.abs(np.asarray(cp2.a_mean) - g["a_mean_new"]).max() < 2e-4
    if name != "mppi":
        assert np.linalg.norm(np.asarray(cp2.a_cov) - g["a_cov"]) / np.linalg.norm(g["a_cov"]) < 5e-5


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["gaussian", "gamma_sigma"])
def test_mppi_beyond_the_defaults_same_key_in_same_action_out(tag):
    """MPPIController.__call__ executed from the reference's source with a PRNGKey (tests/golden/make_reference_golden.py, section 10b)
    vs the product handed the same key:
      * gaussian (the reference's default disturb_type): the rollouts run a STOCHASTIC step_env with one step_key for every sample and
        horizon step (mppi.py:74) -- the controller derives that force from the key and hands it to the rollout kernel;
      * gamma_sigma = 0.3 passed at CALL time: the covariance update of mppi.py:119-125."""
    import covo_mpc_b200 as cm

    g = np.load(os.path.join(HERE, "golden", f"reference_call_mppi_keyed_{tag}.npz"))
    N, H = int(g["N"]), int(g["H"])
    env = cm.Quad3D("tracking_zigzag", disturb_type="gaussian" if tag == "gaussian" else "none")
    ctl, cp = cm.get_controller(env, "mppi", f"N{N}_H{H}_lam{float(g['lam'])}")
    s = g["state24"]
    z3 = np.zeros(3, np.float32)
    st = cm.EnvState3D(pos=s[0:3], quat=s[3:7], vel=s[7:10], omega=s[10:13], f_disturb=s[13:16], pos_tar=s[16:19], vel_tar=s[19:22],
                       acc_tar=z3, pos_traj=g["pos_traj"], vel_traj=g["vel_traj"], acc_traj=np.zeros_like(g["pos_traj"]), time=int(g["time"]))
    cp_in = cp.replace(a_mean=g["a_mean"], gamma_sigma=float(g["gamma_sigma"]))
    action, cp2, _ = ctl(None, st, env.default_params, g["rng_act"], cp_in, {"noisy_state": st})
    assert np.abs(np.asarray(action) - g["action"]).max() < 2e-4
    assert np.abs(np.asarray(cp2.a_mean) - g["a_mean_new"]).max() < 2e-4
    assert np.abs(np.asarray(cp2.a_cov).reshape(H, 4, 4) - g["a_cov"]).max() < 2e-4
    if tag == "gamma_sigma":
        assert np.abs(g["a_cov"] - g["a_cov_in"]).max() > 1e-3  # the reference did update the covariance
    ctl.close()
