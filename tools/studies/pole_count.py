"""How many Zolotarev poles does the dense optimize_sigma path need once the pole inverses are exact (float64), and which poles would have
to be float64 in a mixed-precision variant?  Run from the repo root: python tools/studies/pole_count.py.  Result on the four hard closed-loop
Hessians (tests/golden/hessians/hard_hessians_n200.npz): Sigma error vs the float64 eigen-decomposition 8 poles 1e-4, 10: 8e-6, 12: 7e-7,
13: 2e-7, 14: 5e-8, 16: 4e-9; with 16 poles and float32 LAPACK inverses for the poles t_j >= 1 (6-7 poles in float64): 1e-8 .. 7e-8."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
from oracle import oracle_np as o
from covo_mpc_b200 import _lib
lib = _lib.load()
lib.covo_zolotarev_nodes.argtypes = [C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
def nodes(M, npol):
    m = 1e-2 * (1 - 1e-7)
    sh, w = np.zeros(npol), np.zeros(npol)
    assert lib.covo_zolotarev_nodes(m, M, npol, sh.ctypes.data_as(C.POINTER(C.c_double)), w.ctypes.data_as(C.POINTER(C.c_double))) == 0
    return sh, w
Rs_all = np.load("tests/golden/hessians/hard_hessians_n200.npz")["R"]
print(Rs_all.shape)
n = 200
for idx in range(len(Rs_all)):
    R = Rs_all[idx]
    Rs = (0.5 * (R + R.T)).astype(np.float32).astype(np.float64)
    lam = np.linalg.eigvalsh(Rs)
    S_ref = o.optimize_sigma(Rs, 0.5, np.float64)
    W = lam[-1] - lam[0]
    Mb = 1.02 * W + 1e-2; Mi = 1e-2 * (1 - 1e-7) * 256.0; lad = 0
    while lad < 9 and Mi < Mb: Mi *= 4; lad += 1
    A = Rs - (lam[0] - 1e-2) * np.eye(n)
    logdet = np.log(lam - lam[0] + 1e-2).sum()
    c = np.exp(0.5 * (4 * n * np.log(0.5) + logdet) / n)
    line = f"[{idx}] W {W:.0f} lad {lad}:"
    for npol in (8, 10, 11, 12, 13, 14, 16):
        sh, w = nodes(Mi, npol)
        acc = np.zeros((n, n))
        for t, wt in zip(sh, w):
            acc += wt * np.linalg.inv(A + t * np.eye(n))
        line += f"  m{npol} {np.linalg.norm(c * acc - S_ref) / np.linalg.norm(S_ref):.1e}"
    print(line)
    sh, w = nodes(Mi, 16)
    for thr in (1.0, 10.0, 100.0):
        acc = np.zeros((n, n)); n64 = 0
        for t, wt in zip(sh, w):
            M = A + t * np.eye(n)
            if t < thr: X = np.linalg.inv(M); n64 += 1
            else: X = np.linalg.inv(M.astype(np.float32)).astype(np.float64)
            acc += wt * 0.5 * (X + X.T)
        print(f"      hybrid thr {thr}: {n64} f64 poles, err {np.linalg.norm(c * acc - S_ref) / np.linalg.norm(S_ref):.1e}   poles {np.array2string(np.sort(sh), precision=2)}")
