import sys, numpy as np
sys.path.insert(0, '.')
from oracle import oracle_np as o
from tests.test_emu_sigma_dense import _zolo_table
tab = _zolo_table()
Rall = np.load('tests/golden/hard_hessians_n200.npz')['R']
f32 = np.float32
def inv8_f32(P):
    P = P.astype(f32).copy(); m = P.shape[0]
    for s in range(m):
        p = P[s, s]; rinv = f32(1) / p
        r = P[s, :].copy(); c = P[:, s].copy()
        P = (P - np.outer(c, r * rinv)).astype(f32)
        P[s, :] = r * rinv; P[:, s] = -(c * rinv); P[s, s] = rinv
    return P
def blocked_gj(A, nb=8, pinv="exact", use_sym=False, fused=False):
    A = A.astype(f32).copy(); n = A.shape[0]; swept = np.zeros(n, bool)
    for k0 in range(0, n, nb):
        K = slice(k0, min(k0 + nb, n))
        P = A[K, K]
        Pi = np.linalg.inv(P.astype(np.float64)).astype(f32) if pinv == "exact" else inv8_f32(P)
        raw = A[K, :].copy()
        G = (Pi @ raw).astype(f32)
        if use_sym:
            sig = np.where(swept, -1.0, 1.0).astype(f32)
            C = (raw.T * sig[:, None]).astype(f32)     # A_iK = sigma(i) raw[s][i]
        else:
            C = A[:, K].copy()
        if fused:
            acc = A.astype(np.float64)
            for s in range(G.shape[0]):
                acc = (acc - np.outer(C[:, s].astype(np.float64), G[s].astype(np.float64))).astype(f32).astype(np.float64)  # fma: one rounding per term
            A = acc.astype(f32)
        else:
            A = (A - C @ G).astype(f32)
        A[K, :] = G
        A[:, K] = -(C @ Pi).astype(f32) if not use_sym else -(G.T * np.where(swept, -1.0, 1.0)[:, None]).astype(f32)
        A[K, K] = Pi
        swept[K] = True
    return A
for R in Rall[[1, 3]]:
    n = 200
    Rs = (0.5 * (R + R.T)).astype(f32).astype(np.float64)
    lam = np.linalg.eigvalsh(Rs); W = lam[-1] - lam[0]
    Mb = 1.02 * W + 1e-2; Mi = 1e-2 * (1 - 1e-7) * 256.0; lad = 0
    while lad < 9 and Mi < Mb: Mi *= 4; lad += 1
    sh, w = tab[lad]
    A = Rs - (lam[0] - 1e-2) * np.eye(n)
    rv = np.arange(n)[::-1].copy()
    M = (A + sh[0] * np.eye(n)).astype(f32)[np.ix_(rv, rv)]
    ex = np.linalg.inv(M.astype(np.float64))
    for kw in (dict(), dict(pinv="f32"), dict(use_sym=True), dict(fused=True), dict(pinv="f32", use_sym=True, fused=True)):
        X = blocked_gj(M, **kw).astype(np.float64); X = 0.5 * (X + X.T) if not kw.get("use_sym") else np.tril(X) + np.tril(X, -1).T
        print(f"W {W:.0f} pole0 reversed {kw}: rel err {np.linalg.norm(X - ex) / np.linalg.norm(ex):.2e}")
