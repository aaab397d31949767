"""CPU study: how many Lanczos steps (no reorthogonalisation, fp64, kernel start vector) until |ritz_min - lam_min| < tol, along closed loops."""
import sys, numpy as np
sys.path.insert(0, '.')
from oracle import oracle_c, oracle_np as o
from tools import tracking_protocol as tp
N, H, LAM = 512, 50, 0.01
steps = int(sys.argv[1]); ep = int(sys.argv[2])
def start(n):
    j = np.arange(n, dtype=np.uint64)
    hsh = (((j + 1) * 2654435761) & 0xffffffff) >> 8 & 0xffff
    x = 1.0 + hsh.astype(np.float32).astype(np.float64) / 65536.0
    return x / np.linalg.norm(x)
def lanczos_trace(A, kmax, v, lam0):
    n = A.shape[0]; vp = np.zeros(n); beta = 0.0; al = []; be = []
    out = {}
    for k in range(1, kmax + 1):
        u = A @ v - beta * vp
        a = u @ v
        w = u - a * v
        beta = np.linalg.norm(w)
        al.append(a); be.append(beta)
        vp, v = v, w / beta
        if k >= 16 and k % 8 == 0:
            T = np.diag(al) + np.diag(be[:k-1], 1) + np.diag(be[:k-1], -1)
            ev, V = np.linalg.eigh(T)
            out[k] = (ev[0] - lam0, (be[-1] * V[-1, 0]) ** 2)
    return out
p = o.EnvParams()
s = o.reset_env(tp.TASK, p, np.random.default_rng(tp.episode_seeds(ep)[0]), dtype=np.float32, zero_disturb=False)
noise, eps_rng = tp.episode_noise(ep, steps), tp.episode_eps_rng(ep)
mean = o.hover_mean(H, p)
need = []
for i in range(steps):
    eps = eps_rng.standard_normal((8192, 4 * H)).astype(np.float32)[:N]
    ns = o.noisy_state(s, p, tp.SeqRng(noise[i, :13]))
    a_mean = o.shift_mean(mean.astype(np.float32))
    R = oracle_c.hessian(ns, a_mean, p)
    Rs = (0.5 * (R + R.T)).astype(np.float32).astype(np.float64)
    lam = np.linalg.eigvalsh(Rs)
    tr = lanczos_trace(Rs, 128, start(Rs.shape[0]), lam[0])
    kk = next((k for k in sorted(tr) if abs(tr[k][0]) < 2e-7), 999)
    kres = next((k for k in sorted(tr) if tr[k][1] < 9e-10), 999)
    need.append((kk, kres))
    if kk > 24 or kres > 32:
        print(f"step {i:3d} gap {lam[1]-lam[0]:.2e} width {lam[-1]-lam[0]:.0f} k_true {kk} k_res {kres} " + " ".join(f"{k}:{tr[k][0]:.0e}/{tr[k][1]:.0e}" for k in (24, 32, 48, 64, 96, 128)), flush=True)
    cov = o.optimize_sigma(R, 0.5, dtype=np.float32)
    L = np.linalg.cholesky(cov.astype(np.float64)).astype(np.float32)
    a_s = o.sample_actions(a_mean, L, eps)
    cost = oracle_c.rollout_costs(ns, a_s, p)
    mean, _ = o.softmax_update(a_mean, a_s, cost, LAM)
    s, _, _, _ = o.env_step(s, mean[0], p, tp.SeqRng(noise[i + 1, 13:16]), "none")
need = np.array(need)
for k in (16, 24, 32, 40, 48, 64, 96, 128, 999):
    print(f"k<={k}: true-converged {np.mean(need[:,0] <= k):.3f}  residual-converged {np.mean(need[:,1] <= k):.3f}")
