"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol the header
declares; the host-only entry point works; creating a handle without a GPU fails loudly (no CPU path)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    import ctypes

    hdr = open(os.path.join(ROOT, "include", "covo_b200.h")).read()
    declared = set(re.findall(r"\b(covo_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"covo_config", "covo_handle"}
    assert len(declared) >= 30
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/covo_b200.h but not exported"
    from covo_mpc_b200 import _lib

    assert set(_lib.EXPORTED) == declared


def test_zolotarev_nodes_match_scipy(built_lib):
    from scipy.special import ellipj, ellipk

    from covo_mpc_b200 import _lib

    for m, M, N in ((0.01, 2.56, 16), (0.01 * (1 - 1e-7), 0.01 * 4 ** 8, 16), (1.0, 1e6, 12)):
        t, w = _lib.zolotarev_nodes(m, M, N)
        k2 = 1 - m / M
        K = ellipk(k2)
        u = (np.arange(1, N + 1) - 0.5) * K / N
        sn, cn, dn, _ = ellipj(u, k2)
        assert np.allclose(t, m * (sn / cn) ** 2, rtol=1e-9)
        assert np.allclose(w, (2 * K * np.sqrt(m) / (np.pi * N)) * dn / cn ** 2, rtol=1e-9)
        # the rational function approximates x^(-1/2) on [m, M]
        x = np.geomspace(m, M, 400)
        approx = (w[None] / (x[:, None] + t[None])).sum(1)
        assert np.abs(approx * np.sqrt(x) - 1).max() < (1e-8 if N == 16 else 1e-5)


def test_config_defaults_are_the_reference_defaults(built_lib):
    from covo_mpc_b200 import _lib

    c = _lib.default_config()
    assert (c.n_samples, c.horizon, c.n_env) == (8192, 32, 1)  # envs/quadrotor.py:673-676
    assert abs(c.lam - 0.01) < 1e-7 and abs(c.sample_sigma - 0.5) < 1e-7
    assert abs(c.m - 0.027) < 1e-7 and abs(c.max_thrust - 0.8) < 1e-7 and c.max_steps_in_episode == 300
    assert [round(x, 5) for x in c.max_omega] == [10.0, 10.0, 3.0]


def test_no_cpu_fallback(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from covo_mpc_b200 import _lib

    with pytest.raises(_lib.CovoCudaError, match="no CPU path"):
        _lib.Handle(_lib.default_config())


def test_get_controller_dispatch_and_errors(built_lib):
    import covo_mpc_b200 as cm

    env = cm.Quad3D("tracking_zigzag")
    ctl, cp = cm.get_controller(env, "covo_offline", "N64_H8_lam0.05")
    assert isinstance(ctl, cm.CoVOController) and ctl.mode == "offline" and (ctl.N, ctl.H, ctl.lam) == (64, 8, 0.05)
    assert cp.a_mean.shape == (8, 4) and abs(cp.a_mean[0, 0] - (-0.337825)) < 1e-6 and cp.a_cov.shape == (32, 32)
    ctl, cp = cm.get_controller(env, "covo", "")
    assert ctl.mode == "online" and (ctl.N, ctl.H) == (8192, 32)
    ctl, cp = cm.get_controller(env, "mppi", "N128_H32_lam0.01")
    assert isinstance(ctl, cm.MPPIController) and cp.a_cov.shape == (32, 4, 4) and cp.a_cov[3, 2, 2] == 0.25
    with pytest.raises(NotImplementedError):
        cm.get_controller(env, "lqr")
    with pytest.raises(NotImplementedError):
        cm.Quad3D("jumping")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "covo_mpc_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle/oracle_np.py", "").replace("the oracle", ""), fn
