import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
from oracle import oracle_c, oracle_np as o
from tools import tracking_protocol as tp
from tests.test_emu_sigma_dense import _zolo_table
import scipy.linalg as sl
N, H, LAM = 512, 50, 0.01
steps = int(sys.argv[1]); ep = 0
tab = _zolo_table()
p = o.EnvParams()
s = o.reset_env(tp.TASK, p, np.random.default_rng(tp.episode_seeds(ep)[0]), dtype=np.float32, zero_disturb=False)
noise, eps_rng = tp.episode_noise(ep, steps), tp.episode_eps_rng(ep)
mean = o.hover_mean(H, p)
n = 4 * H
def gj_inverse_f32(A):
    """in-place Gauss-Jordan sweep without pivoting, float32 arithmetic"""
    A = A.astype(np.float32).copy(); n = A.shape[0]
    for k in range(n):
        pv = A[k, k]; r = A[k, :].copy(); c = A[:, k].copy()
        A -= np.outer(c, r / pv).astype(np.float32)
        A[k, :] = r / pv; A[:, k] = -c / pv; A[k, k] = np.float32(1) / pv
    # sign convention: sweeping all indices yields -A^-1?? fix by comparing
    return A
for i in range(steps):
    eps = eps_rng.standard_normal((8192, 200)).astype(np.float32)[:N]
    ns = o.noisy_state(s, p, tp.SeqRng(noise[i, :13]))
    a_mean = o.shift_mean(mean.astype(np.float32))
    R = oracle_c.hessian(ns, a_mean, p)
    Rs = (0.5 * (R + R.T)).astype(np.float32).astype(np.float64)
    lam = np.linalg.eigvalsh(Rs)
    S_ref = o.optimize_sigma(Rs, 0.5, np.float64)
    W = lam[-1] - lam[0]
    Mb = 1.02 * W + 1e-2; Mi = 1e-2 * (1 - 1e-7) * 256.0; lad = 0
    while lad < 9 and Mi < Mb: Mi *= 4; lad += 1
    sh, w = tab[lad]
    A = Rs - (lam[0] - 1e-2) * np.eye(n)
    logdet = np.log(lam - lam[0] + 1e-2).sum()
    c = np.exp(0.5 * (4 * n * np.log(0.5) + logdet) / n)
    out = {}
    for name, inv in (("f64", lambda M: np.linalg.inv(M)),
                      ("f32 lapack", lambda M: np.linalg.inv(M.astype(np.float32)).astype(np.float64)),
                      ("f32 chol", lambda M: sl.cho_solve(sl.cho_factor(M.astype(np.float32), lower=True), np.eye(n, dtype=np.float32)).astype(np.float64))):
        acc = np.zeros((n, n))
        for t, wt in zip(sh, w):
            X = inv(A + t * np.eye(n)); acc += wt * 0.5 * (X + X.T)
        out[name] = np.linalg.norm(c * acc - S_ref) / np.linalg.norm(S_ref)
    print(f"step {i} W {W:.0f} ladder {lad} smallest pole {sh.min():.2e}: " + "  ".join(f"{k} {v:.2e}" for k, v in out.items()), flush=True)
    cov32 = o.optimize_sigma(R, 0.5, dtype=np.float32)
    L = np.linalg.cholesky(cov32.astype(np.float64)).astype(np.float32)
    a_s = o.sample_actions(a_mean, L, eps)
    cost = oracle_c.rollout_costs(ns, a_s, p)
    mean, _ = o.softmax_update(a_mean, a_s, cost, LAM)
    s, _, _, _ = o.env_step(s, mean[0], p, tp.SeqRng(noise[i + 1, 13:16]), "none")
