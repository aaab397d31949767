"""Reference-executed fixtures at the BASELINE sizes (tests/golden/reference_fullsize_*.npz, written by
tests/golden/make_reference_golden_fullsize.py from the UNMODIFIED reference source under the NumPy shim): optimize_sigma at n = 200,
one MPPI call and one CoVO-offline call at N = 8192, H = 50.  The N x 4H draws are regenerated from the recorded seed (the generator
asserts that rule against its logged draws).  CPU tests hold the oracle to them, GPU tests the kernels (through the C-ABI)."""
import os

import numpy as np
import pytest

from oracle import oracle_np as o

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    path = os.path.join(HERE, "golden", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated yet (tests/golden/make_reference_golden_fullsize.py)")
    return np.load(path)


def _state(g):
    s = g["state24"]
    return o.make_state(s[0:3], s[3:7], s[7:10], s[10:13], s[13:16], int(g["time"]), g["pos_traj"], g["vel_traj"], s[16:19], s[19:22],
                        dtype=np.float32)


def _eps(g, mppi):
    N, H = int(g["N"]), int(g["H"])
    rng = np.random.default_rng(int(g["seed"]))
    if mppi:
        return rng.standard_normal((N * H, 4)).astype(np.float32).reshape(N, H, 4)
    return rng.standard_normal((N, 4 * H)).astype(np.float32)


# ---- CPU: the oracle against the reference's own execution -----------------------------------------------------------------------
def test_oracle_optimize_sigma_n200():
    g = _load("reference_fullsize_optimize_sigma_n200.npz")
    S = o.optimize_sigma(g["R"], 0.5, dtype=np.float32)
    assert np.linalg.norm(S - g["Sigma"]) / np.linalg.norm(g["Sigma"]) < 5e-6  # same float32 LAPACK eigh, same formula
    S64 = o.optimize_sigma(g["R"].astype(np.float64), 0.5, np.float64)
    # the reference's float32 result against exact arithmetic: the yardstick for the kernels below
    print("reference float32 optimize_sigma vs float64:", np.linalg.norm(g["Sigma"] - S64) / np.linalg.norm(S64))


def test_oracle_mppi_call_fullsize():
    g = _load("reference_fullsize_mppi_N8192_H50.npz")
    H = int(g["H"])
    a_cov = np.tile(np.eye(4, dtype=np.float32) * 0.25, (H, 1, 1))
    u, mean, _, info = o.mppi_call(_state(g), g["a_mean"], a_cov, _eps(g, True), o.EnvParams(), lam=float(g["lam"]))
    assert np.abs(u - g["action"]).max() < 2e-5 and np.abs(mean - g["a_mean_new"]).max() < 2e-5
    assert np.abs(info["pos_mean"] - g["pos_mean"]).max() < 1e-4 and np.abs(info["pos_std"] - g["pos_std"]).max() < 1e-4


def _softmax_tolerance(cost, lam, base=2e-5):
    """The update is an arg-min softened by lam = 0.01: d w / d cost = w (1 - w) / lam.  Two float32 evaluations of the same rollouts
    differ by a few ulp of the cost (~1e-6 at cost ~12), which moves the weights of the two best samples by w (1 - w) ulp / lam and the
    mean by that times their distance (<= 2 per component after the clip).  Allow 4 ulp."""
    c = np.sort(cost.astype(np.float64))
    w = np.exp(-(c - c[0]) / lam)
    w /= w.sum()
    return base + 4.0 * float(np.spacing(np.float32(c[0]))) * float(w[0] * (1.0 - w[0])) / lam * 2.0


def test_oracle_covo_offline_call_fullsize():
    """(this fixture is a near-tie: the two best of the 8192 samples are 1.04 lambda apart, weights 0.74 / 0.26)"""
    g = _load("reference_fullsize_covo_offline_N8192_H50.npz")
    u, mean, _, info, dbg = o.covo_call(_state(g), g["a_mean"], _eps(g, False), o.EnvParams(), lam=float(g["lam"]), a_cov=g["a_cov"],
                                        return_debug=True)
    tol = _softmax_tolerance(dbg["cost"], float(g["lam"]))
    assert np.abs(u - g["action"]).max() < tol and np.abs(mean - g["a_mean_new"]).max() < tol
    assert np.abs(info["pos_mean"] - g["pos_mean"]).max() < 1e-4


# ---- GPU: the kernels against the reference's own execution ----------------------------------------------------------------------
def _handle(mode, g, **kw):
    from covo_mpc_b200 import _lib

    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.lam = mode, int(g["N"]) if "N" in g else 64, int(g["H"]), 320, 0.01
    for k, v in kw.items():
        setattr(cfg, k, v)
    return _lib.Handle(cfg)


@pytest.mark.gpu
def test_gpu_optimize_sigma_n200():
    from covo_mpc_b200 import _lib

    g = _load("reference_fullsize_optimize_sigma_n200.npz")
    h = _handle(_lib.MODE_COVO_ONLINE, g)
    S = h.optimize_sigma(g["R"][None])[0]
    S64 = o.optimize_sigma(g["R"].astype(np.float64), 0.5, np.float64)
    ref_err = np.linalg.norm(g["Sigma"] - S64) / np.linalg.norm(S64)  # the reference's own float32 rounding
    assert np.linalg.norm(S - S64) / np.linalg.norm(S64) < max(1e-5, 2 * ref_err)
    assert np.linalg.norm(S - g["Sigma"]) / np.linalg.norm(g["Sigma"]) < max(2e-5, 3 * ref_err)
    h.close()


@pytest.mark.gpu
def test_gpu_mppi_call_fullsize():
    from covo_mpc_b200 import _lib

    g = _load("reference_fullsize_mppi_N8192_H50.npz")
    N, H = int(g["N"]), int(g["H"])
    h = _handle(_lib.MODE_MPPI, g)
    h.set_reference(g["pos_traj"][None], g["vel_traj"][None])
    h.set_mean(g["a_mean"][None])
    act = h.step(g["state24"], [int(g["time"])], _eps(g, True).reshape(1, N, 4 * H))[0]
    assert np.abs(act - g["action"]).max() < 2e-5
    assert np.abs(h.get_mean()[0].reshape(H, 4) - g["a_mean_new"]).max() < 2e-5
    h.close()


@pytest.mark.gpu
def test_gpu_covo_offline_call_fullsize():
    from covo_mpc_b200 import _lib

    g = _load("reference_fullsize_covo_offline_N8192_H50.npz")
    N, H = int(g["N"]), int(g["H"])
    h = _handle(_lib.MODE_COVO_OFFLINE, g)
    h.set_reference(g["pos_traj"][None], g["vel_traj"][None])
    h.set_cov_offline(g["a_cov"][None])  # one-entry table: every time index clamps to it (the reference call looked up this entry)
    h.set_mean(g["a_mean"][None])
    eps = _eps(g, False)
    act = h.step(g["state24"], [int(g["time"])], eps[None])[0]
    dbg = o.covo_call(_state(g), g["a_mean"], eps, o.EnvParams(), lam=float(g["lam"]), a_cov=g["a_cov"], return_debug=True)[4]
    tol = _softmax_tolerance(dbg["cost"], float(g["lam"]))  # near-tie fixture, see the CPU test
    assert np.abs(act - g["action"]).max() < tol
    assert np.abs(h.get_mean()[0].reshape(H, 4) - g["a_mean_new"]).max() < tol
    h.close()
