"""CPU study: Lanczos warm-started from the previous step's lowest eigenvector (time-shifted) vs cold hash start."""
import sys, numpy as np
sys.path.insert(0, '.')
from oracle import oracle_c, oracle_np as o
from tools import tracking_protocol as tp
N, H, LAM = 512, 50, 0.01
steps = int(sys.argv[1]); ep = int(sys.argv[2]); mix = float(sys.argv[3])
def start(n):
    j = np.arange(n, dtype=np.uint64)
    hsh = (((j + 1) * 2654435761) & 0xffffffff) >> 8 & 0xffff
    x = 1.0 + hsh.astype(np.float32).astype(np.float64) / 65536.0
    return x / np.linalg.norm(x)
def k_needed(A, kmax, v, lam0, tol=2e-7):
    n = A.shape[0]; vp = np.zeros(n); beta = 0.0; al = []; be = []
    for k in range(1, kmax + 1):
        u = A @ v - beta * vp
        a = u @ v
        w = u - a * v
        beta = np.linalg.norm(w)
        al.append(a); be.append(beta)
        vp, v = v, w / beta
        if k >= 4 and k % 4 == 0:
            T = np.diag(al) + np.diag(be[:k-1], 1) + np.diag(be[:k-1], -1)
            ev, V = np.linalg.eigh(T)
            if (be[-1] * V[-1, 0]) ** 2 < 9e-10: return k
    return 999
p = o.EnvParams()
s = o.reset_env(tp.TASK, p, np.random.default_rng(tp.episode_seeds(ep)[0]), dtype=np.float32, zero_disturb=False)
noise, eps_rng = tp.episode_noise(ep, steps), tp.episode_eps_rng(ep)
mean = o.hover_mean(H, p)
need = []; v0_prev = None
for i in range(steps):
    eps = eps_rng.standard_normal((8192, 4 * H)).astype(np.float32)[:N]
    ns = o.noisy_state(s, p, tp.SeqRng(noise[i, :13]))
    a_mean = o.shift_mean(mean.astype(np.float32))
    R = oracle_c.hessian(ns, a_mean, p)
    Rs = (0.5 * (R + R.T)).astype(np.float32).astype(np.float64)
    lam, V = np.linalg.eigh(Rs)
    n = Rs.shape[0]
    cold = k_needed(Rs, 96, start(n), lam[0])
    if v0_prev is not None:
        w = np.concatenate([v0_prev[4:], v0_prev[-4:]])
        w = w / np.linalg.norm(w)
        if w.sum() < 0: w = -w
        x = w + mix * start(n); x /= np.linalg.norm(x)
        warm = k_needed(Rs, 96, x, lam[0])
        ov = abs(w @ V[:, 0])
    else:
        warm, ov = cold, 0
    need.append((cold, warm, ov))
    v0_prev = V[:, 0]
    cov = o.optimize_sigma(R, 0.5, dtype=np.float32)
    L = np.linalg.cholesky(cov.astype(np.float64)).astype(np.float32)
    a_s = o.sample_actions(a_mean, L, eps)
    cost = oracle_c.rollout_costs(ns, a_s, p)
    mean, _ = o.softmax_update(a_mean, a_s, cost, LAM)
    s, _, _, _ = o.env_step(s, mean[0], p, tp.SeqRng(noise[i + 1, 13:16]), "none")
need = np.array(need)
print("mean k cold", need[:,0].mean(), "warm", need[:,1].mean(), "median overlap", np.median(need[:,2]))
print("hist cold", np.bincount(need[:,0].astype(int)//4)[:26])
print("hist warm", np.bincount(need[:,1].astype(int)//4)[:26])
