"""CPU study: Lanczos-24 (kernel start vector) on the Hessians of the episode-0 closed loop (oracle loop)."""
import sys, numpy as np
sys.path.insert(0, '.')
from oracle import oracle_c, oracle_np as o
from tools import tracking_protocol as tp
N, H, LAM = 1024, 50, 0.01
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ep = int(sys.argv[2]) if len(sys.argv) > 2 else 0
def start(n):
    j = np.arange(n, dtype=np.uint64)
    hsh = (((j + 1) * 2654435761) & 0xffffffff) >> 8 & 0xffff
    x = 1.0 + hsh.astype(np.float32).astype(np.float64) / 65536.0
    return x / np.linalg.norm(x)
def lanczos(A, k, v):
    n = A.shape[0]; vp = np.zeros(n); beta = 0.0; al = []; be = []
    for _ in range(k):
        u = A @ v - beta * vp
        a = u @ v
        b2 = u @ u - a * a
        w = u - a * v
        if b2 < 1e-3 * (u @ u): b2 = w @ w
        beta = np.sqrt(b2)
        al.append(a); be.append(beta)
        vp, v = v, w / beta
    m = len(al)
    T = np.diag(al) + np.diag(be[:m-1], 1) + np.diag(be[:m-1], -1)
    ev, V = np.linalg.eigh(T)
    res = be[-1] * abs(V[-1, 0])
    return ev[0], ev[-1], res
p = o.EnvParams()
s = o.reset_env(tp.TASK, p, np.random.default_rng(tp.episode_seeds(ep)[0]), dtype=np.float32, zero_disturb=False)
noise, eps_rng = tp.episode_noise(ep, steps), tp.episode_eps_rng(ep)
mean = o.hover_mean(H, p)
for i in range(steps):
    eps = eps_rng.standard_normal((8192, 4 * H)).astype(np.float32)[:N]
    ns = o.noisy_state(s, p, tp.SeqRng(noise[i, :13]))
    a_mean = o.shift_mean(mean.astype(np.float32))
    R = oracle_c.hessian(ns, a_mean, p)
    Rs = (0.5 * (R + R.T)).astype(np.float32).astype(np.float64)
    lam = np.linalg.eigvalsh(Rs)
    l0, l1, res = lanczos(Rs, 24, start(Rs.shape[0]))
    cov = o.optimize_sigma(R, 0.5, dtype=np.float32)
    L = np.linalg.cholesky(cov.astype(np.float64)).astype(np.float32)
    a_s = o.sample_actions(a_mean, L, eps)
    cost = oracle_c.rollout_costs(ns, a_s, p)
    new_mean, _ = o.softmax_update(a_mean, a_s, cost, LAM)
    mean = new_mean
    print(f"step {i:3d} lam0 {lam[0]:.5f} lam1 {lam[1]:.5f} gap {lam[1]-lam[0]:.3e} lmax {lam[-1]:.1f} | lanczos dl {l0-lam[0]:.3e} res2 {res*res:.2e} lmaxerr {lam[-1]-l1:.2e}", flush=True)
    s, _, _, _ = o.env_step(s, new_mean[0], p, tp.SeqRng(noise[i + 1, 13:16]), "none")
