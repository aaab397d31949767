"""csrc/sigma_dense.cu (the tridiagonalisation-free optimize_sigma, default for single environments) executed on the CPU stand-in for the CUDA execution model
(tests/emu/cuda_runtime.h: every CUDA thread a cooperative fiber; barriers, named barriers, shuffles and ballots block until
the peers arrive; shared memory poisoned; a barrier that can never complete aborts) against the float64 eigen-decomposition
of the oracle.  Checks the kernels' LOGIC -- indexing, hand-overs through shared memory, barrier placement -- with the launch
geometry and shared-memory layout of launch_sigma_dense; it says nothing about timing or the GPU memory model."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_np as o
from tests.util import scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
kFirstCheck = 16  # csrc/sigma_dense.cu: kLanczosFirstCheck
kDensePoles = 13  # csrc/sigma.cuh


def _emu_lib():
    so = os.path.join(EMU, "libemu_sigma_dense.so")
    srcs = [os.path.join(EMU, "run_sigma_dense.cpp"), os.path.join(EMU, "cuda_runtime.h"),
            os.path.join(ROOT, "covo_mpc_b200", "csrc", "sigma_dense.cu"), os.path.join(ROOT, "covo_mpc_b200", "csrc", "common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-DCOVO_CPU_EMU", "-I" + EMU, "-I" + os.path.join(ROOT, "include"), "-shared", "-fPIC",
                               "-o", so, srcs[0]])
    return C.CDLL(so)


def test_dense_sigma_at_the_headline_size(variant=3):
    """n = 200 (H = 50): the production kernels (8-CTA cluster Lanczos with its checker warp, float64 blocked Gauss-Jordan on an 8-CTA cluster
    per pole) run with all CTAs of a cluster interleaved."""
    emu = _emu_lib()
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=3, H=50, warm_steps=6)
    R = o.get_hessian(ns, o.shift_mean(a_mean), p, dtype=np.float64).astype(np.float32)
    S_ref = o.optimize_sigma(R.astype(np.float64), 0.5, np.float64)
    cov = np.full((200, 200), np.nan, np.float32)
    scal, status, tab = np.zeros(4), np.zeros(1, np.int32), _zolo_table()
    rc = emu.emu_sigma_dense(200, C.c_float(0.5), R.ctypes.data_as(C.POINTER(C.c_float)), tab.ctypes.data_as(C.POINTER(C.c_double)),
                             cov.ctypes.data_as(C.POINTER(C.c_float)), scal.ctypes.data_as(C.POINTER(C.c_double)),
                             status.ctypes.data_as(C.POINTER(C.c_int)), variant)
    assert rc == 0 and status[0] == 0 and np.isfinite(cov).all()
    lam = np.linalg.eigvalsh(0.5 * (R + R.T).astype(np.float64))
    assert abs(scal[0] - lam[0]) < 2e-8 and kFirstCheck <= scal[3] <= 64  # adaptive Lanczos: converged, and says how many steps it took
    assert np.linalg.norm(cov - S_ref) / np.linalg.norm(S_ref) < 1e-6


def _zolo_table():
    from covo_mpc_b200 import _lib

    lib = _lib.load()  # the ladder of sigma.cu, through the C-ABI's host-side helper (no GPU involved)
    lib.covo_zolotarev_nodes.argtypes = [C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    m = 1e-2 * (1 - 1e-7)
    tab = np.zeros((10, 2, kDensePoles))
    for i in range(10):
        sh, w = np.zeros(kDensePoles), np.zeros(kDensePoles)
        assert lib.covo_zolotarev_nodes(m, m * 4.0 ** (4 + i), kDensePoles, sh.ctypes.data_as(C.POINTER(C.c_double)), w.ctypes.data_as(C.POINTER(C.c_double))) == 0
        tab[i, 0], tab[i, 1] = sh, w
    return tab


@pytest.mark.parametrize("variant", [3, 3 | 16])  # 8 (default) and 16 pivots per elimination step (n = 200 with 16: run once by hand, 2 min)
@pytest.mark.parametrize("H", [2, 8, 9])  # n = 8 (Krylov space exhausted before the first checkpoint), 32, and the ragged 36 (n_pad = 40)
def test_dense_sigma_kernels_on_the_cpu_execution_model(H, variant):
    emu = _emu_lib()
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=3, H=H, warm_steps=6)
    R = o.get_hessian(ns, o.shift_mean(a_mean), p, dtype=np.float64).astype(np.float32)
    n = 4 * H
    S_ref = o.optimize_sigma(R.astype(np.float64), 0.5, np.float64)
    lam = np.linalg.eigvalsh(0.5 * (R + R.T).astype(np.float64))
    cov = np.full((n, n), np.nan, np.float32)
    scal, status, tab = np.zeros(4), np.zeros(1, np.int32), _zolo_table()
    rc = emu.emu_sigma_dense(n, C.c_float(0.5), R.ctypes.data_as(C.POINTER(C.c_float)), tab.ctypes.data_as(C.POINTER(C.c_double)),
                             cov.ctypes.data_as(C.POINTER(C.c_float)), scal.ctypes.data_as(C.POINTER(C.c_double)),
                             status.ctypes.data_as(C.POINTER(C.c_int)), variant)
    assert rc == 0 and status[0] == 0
    width = lam[-1] - lam[0]
    assert abs(scal[0] - lam[0]) < 1e-8 * max(1.0, width)  # Lanczos + multisection, fp64; needed: << 2e-7 absolute (1e-2 offset x 2e-5)
    assert -1e-9 * width <= scal[1] - lam[-1] < 0.5 * width   # lambda_max only selects the approximation interval: a coarse UPPER bound
    assert abs(scal[2] - np.log(lam - lam[0] + 1e-2).sum()) < 1e-4          # log det from the fp32 factorisation
    assert np.isfinite(cov).all() and np.array_equal(cov, cov.T)
    assert np.linalg.norm(cov - S_ref) / np.linalg.norm(S_ref) < 2e-6


def test_adaptive_lanczos_on_a_hessian_that_defeats_24_steps():
    """The last matrix of tests/golden/hessians/hard_hessians_n200.npz (gap 6e-2, width 1640): after 24 steps the smallest Ritz value is
    still 6e-2 above lambda_min (A would be indefinite); the checker warp keeps the recurrence going until the residual says so."""
    emu = _emu_lib()
    R = np.ascontiguousarray(np.load(os.path.join(ROOT, "tests", "golden", "hessians", "hard_hessians_n200.npz"))["R"][-1])
    lam = np.linalg.eigvalsh(0.5 * (R + R.T).astype(np.float64))
    S_ref = o.optimize_sigma(R.astype(np.float64), 0.5, np.float64)
    cov = np.full((200, 200), np.nan, np.float32)
    scal, status, tab = np.zeros(4), np.zeros(1, np.int32), _zolo_table()
    rc = emu.emu_sigma_dense(200, C.c_float(0.5), R.ctypes.data_as(C.POINTER(C.c_float)), tab.ctypes.data_as(C.POINTER(C.c_double)),
                             cov.ctypes.data_as(C.POINTER(C.c_float)), scal.ctypes.data_as(C.POINTER(C.c_double)),
                             status.ctypes.data_as(C.POINTER(C.c_int)), 3)
    assert rc == 0 and status[0] == 0
    assert scal[3] > 24 and abs(scal[0] - lam[0]) < 2e-8
    assert np.linalg.norm(cov - S_ref) / np.linalg.norm(S_ref) < 1e-6  # float64 inverses: what is left is the 13-pole approximation (2e-7)
