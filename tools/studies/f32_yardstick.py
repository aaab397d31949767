"""How far is the REFERENCE's float32 pipeline (float32 Hessian, float32 eigh) from exact arithmetic on the same inputs? (closed loop, episode 0)"""
import sys, numpy as np
sys.path.insert(0, '.')
from oracle import oracle_c, oracle_np as o
from tools import tracking_protocol as tp
N, H, LAM = 512, 50, 0.01
steps = int(sys.argv[1])
p = o.EnvParams()
s = o.reset_env(tp.TASK, p, np.random.default_rng(tp.episode_seeds(0)[0]), dtype=np.float32, zero_disturb=False)
noise, eps_rng = tp.episode_noise(0, steps), tp.episode_eps_rng(0)
mean = o.hover_mean(H, p)
errs = []
for i in range(steps):
    eps = eps_rng.standard_normal((8192, 200)).astype(np.float32)[:N]
    ns = o.noisy_state(s, p, tp.SeqRng(noise[i, :13]))
    a_mean = o.shift_mean(mean.astype(np.float32))
    R32 = oracle_c.hessian(ns, a_mean, p)
    R64 = oracle_c.hessian_f64(ns, a_mean, p)
    S64 = o.optimize_sigma(R64, 0.5, np.float64)
    S32 = o.optimize_sigma(R32, 0.5, np.float32).astype(np.float64)
    S32x = o.optimize_sigma(R32.astype(np.float64), 0.5, np.float64)   # exact Sigma of the float32 Hessian
    lam = np.linalg.eigvalsh(R64)
    nr = np.linalg.norm(S64)
    errs.append((np.linalg.norm(S32 - S64) / nr, np.linalg.norm(S32x - S64) / nr, np.abs(R32 - R64).max(), lam[-1] - lam[0]))
    print(f"step {i:3d} W {errs[-1][3]:7.0f} |R32-R64|max {errs[-1][2]:.1e}  Sigma(f32 pipeline) vs truth {errs[-1][0]:.1e}  Sigma(exact fn of f32 Hessian) vs truth {errs[-1][1]:.1e}", flush=True)
    L = np.linalg.cholesky(S32).astype(np.float32)
    a_s = o.sample_actions(a_mean, L, eps)
    cost = oracle_c.rollout_costs(ns, a_s, p)
    mean, _ = o.softmax_update(a_mean, a_s, cost, LAM)
    s, _, _, _ = o.env_step(s, mean[0], p, tp.SeqRng(noise[i + 1, 13:16]), "none")
e = np.array(errs)
print("median / max Sigma error of the float32 pipeline:", np.median(e[:, 0]), e[:, 0].max(), " of the exact function of the float32 Hessian:", np.median(e[:, 1]), e[:, 1].max())
