"""Round-2 feasibility study (CPU, NumPy/SciPy): optimize_sigma without an eigen-decomposition -- lambda_min by Lanczos (fp64
arithmetic on the fp32 matrix), A^(-1/2) by Zolotarev partial fractions with dense fp32 Cholesky solves, log det A from one more
factorisation -- against the float64 eigen-decomposition, across tasks / horizons / states.  The CUDA counterpart is
csrc/sigma_dense.cu (the opt-in fast path, COVO_SIGMA=dense).  Prints per scenario: |lambda_min error| and the relative Frobenius error of
Sigma for (Lanczos steps, poles) = (16, 8), (24, 8), (32, 10).

    python tools/study_dense_sigma.py
"""
import ctypes as C, sys, time
import numpy as np, scipy.linalg as sl
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import oracle_np as o
from tests.util import scenario
from covo_mpc_b200 import _lib
lib = _lib.load()
lib.covo_zolotarev_nodes.argtypes = [C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
def zolo(m, M, k):
    sh = np.zeros(k); w = np.zeros(k)
    assert lib.covo_zolotarev_nodes(m, M, k, sh.ctypes.data_as(C.POINTER(C.c_double)), w.ctypes.data_as(C.POINTER(C.c_double))) == 0
    return sh, w
def lanczos(A32, k):
    """matrix entries fp32, all arithmetic fp64 (what a kernel with fp64 accumulators would do)"""
    A = A32.astype(np.float64); n = A.shape[0]
    v = np.cos(0.37 * np.arange(n) + 0.1) + 0.01 * np.arange(n) / n; v /= np.linalg.norm(v)   # deterministic start vector
    vp = np.zeros(n); beta = 0.0; al = []; be = []
    for _ in range(k):
        w = A @ v - beta * vp
        a = w @ v; w -= a * v
        beta = np.linalg.norm(w)
        al.append(a); be.append(beta)
        if beta < 1e-12: break
        vp, v = v, w / beta
    m = len(al)
    T = np.diag(al) + np.diag(be[:m-1], 1) + np.diag(be[:m-1], -1)
    ev = np.linalg.eigvalsh(T)
    return ev[0], ev[-1]
def sigma_dense(R32, lmin, lmax, poles):
    n = R32.shape[0]
    Rs = R32.astype(np.float64) - (lmin - 1e-2) * np.eye(n)
    width = lmax - lmin
    m, M = 1e-2 * 0.9, (width + 1e-2) * 1.1      # safety margins on both ends
    sh, w = zolo(m, M, poles)
    Lc = np.linalg.cholesky(Rs.astype(np.float32).astype(np.float64))
    logdet = 2 * np.log(np.diag(Lc)).sum()
    acc = np.zeros((n, n))
    for s_, w_ in zip(sh, w):
        A = (Rs + s_ * np.eye(n)).astype(np.float32)
        c, low = sl.cho_factor(A, lower=True)
        acc += w_ * sl.cho_solve((c, low), np.eye(n, dtype=np.float32)).astype(np.float64)
    log_const = (n * np.log(0.5) * 2 * 2 + logdet) / n
    return np.exp(0.5 * log_const) * acc
worst = 0
for task in ("tracking_zigzag", "tracking", "hovering"):
    for H in (8, 20, 32, 50):
        for seed, warm in ((1, 0), (2, 5), (3, 40)):
            p, ns, a_mean, rng = scenario(task, seed=seed, H=H, warm_steps=warm)
            R = o.get_hessian(ns, o.shift_mean(a_mean), p, dtype=np.float64).astype(np.float32)
            R = ((R + R.T) / 2).astype(np.float32)
            lam = np.linalg.eigvalsh(R.astype(np.float64))
            S_ref = o.optimize_sigma(R.astype(np.float64), 0.5, np.float64)
            out = []
            for k, poles in ((16, 8), (24, 8), (32, 10)):
                l0, l1 = lanczos(R, min(k, R.shape[0]))
                S = sigma_dense(R, l0, l1, poles)
                out.append((abs(l0 - lam[0]), np.linalg.norm(S - S_ref) / np.linalg.norm(S_ref)))
            worst = max(worst, out[1][1])
            print(f"{task:16s} H={H:2d} warm={warm:2d} gap={lam[1]-lam[0]:.2e} width={lam[-1]-lam[0]:.1f} | " +
                  "  ".join(f"dl={a:.1e} err={b:.1e}" for a, b in out))
print("worst (k=24, 8 poles):", worst)
