import sys, numpy as np, scipy.linalg as sl
sys.path.insert(0, '.')
from oracle import oracle_np as o
from tests.test_emu_sigma_dense import _zolo_table
tab = _zolo_table()
Rall = np.load('tests/golden/hard_hessians_n200.npz')['R']
f32 = np.float32
def gj(A, order=None):
    A = A.astype(f32).copy(); n = A.shape[0]
    order = range(n) if order is None else order
    for k in order:
        pv = A[k, k]; r = A[k, :].copy(); c = A[:, k].copy()
        g = (r / pv).astype(f32)
        A = (A - np.outer(c, g)).astype(f32)
        A[k, :] = g; A[:, k] = (c / pv).astype(f32); A[k, k] = -f32(1) / pv   # sweep operator convention: full sweep gives -inverse... check sign below
    return A
def blocked_gj(A, nb=8):
    """block sweep like the kernel: A_IJ -= A_IK P^-1 A_KJ etc."""
    A = A.astype(f32).copy(); n = A.shape[0]
    for k0 in range(0, n, nb):
        K = slice(k0, min(k0 + nb, n))
        P = A[K, K].astype(f32); Pi = np.linalg.inv(P.astype(np.float64)).astype(f32)
        G = (Pi @ A[K, :]).astype(f32); C = A[:, K].copy()
        A = (A - C @ G).astype(f32)
        A[K, :] = G; A[:, K] = -(C @ Pi).astype(f32); A[K, K] = Pi
    return A
for R in Rall:
    n = 200
    Rs = (0.5 * (R + R.T)).astype(f32).astype(np.float64)
    lam = np.linalg.eigvalsh(Rs); W = lam[-1] - lam[0]
    S_ref = o.optimize_sigma(Rs, 0.5, np.float64)
    Mb = 1.02 * W + 1e-2; Mi = 1e-2 * (1 - 1e-7) * 256.0; lad = 0
    while lad < 9 and Mi < Mb: Mi *= 4; lad += 1
    sh, w = tab[lad]
    A = Rs - (lam[0] - 1e-2) * np.eye(n)
    c = np.exp(0.5 * (4 * n * np.log(0.5) + np.log(lam - lam[0] + 1e-2).sum()) / n)
    res = {}
    def run(name, inv):
        acc = np.zeros((n, n))
        for t, wt in zip(sh, w):
            M = (A + t * np.eye(n)).astype(f32)   # fp32 matrix incl. shift rounding
            X = inv(M).astype(np.float64); acc += wt * 0.5 * (X + X.T)
        res[name] = np.linalg.norm(c * acc - S_ref) / np.linalg.norm(S_ref)
    run("exact inv of fp32 shifted matrix", lambda M: np.linalg.inv(M.astype(np.float64)))
    run("f32 LU", lambda M: np.linalg.inv(M))
    run("f32 chol", lambda M: sl.cho_solve(sl.cho_factor(M, lower=True), np.eye(n, dtype=f32)))
    run("f32 blocked GJ natural", lambda M: blocked_gj(M))
    d = np.argsort(-np.diag(A))
    def gjp(M):
        X = blocked_gj(M[np.ix_(d, d)]); out = np.empty_like(X); out[np.ix_(d, d)] = X; return out
    run("f32 blocked GJ diag-sorted", gjp)
    dr = d[::-1].copy()
    def gjr(M):
        X = blocked_gj(M[np.ix_(dr, dr)]); out = np.empty_like(X); out[np.ix_(dr, dr)] = X; return out
    run("f32 blocked GJ reverse-sorted", gjr)
    rv = np.arange(n)[::-1].copy()
    def gjrev(M):
        X = blocked_gj(M[np.ix_(rv, rv)]); out = np.empty_like(X); out[np.ix_(rv, rv)] = X; return out
    run("GJ reversed natural", gjrev)
    for k in ("exact inv of fp32 shifted matrix", "f32 LU", "f32 blocked GJ diag-sorted"): res.pop(k)
    print(f"W {W:.0f}: " + "  ".join(f"{k}: {v:.1e}" for k, v in res.items()))
