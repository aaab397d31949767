"""CPU: the production dense kernels (on the execution-model stand-in) along a closed loop: lambda_min error, steps taken, Sigma error."""
import sys, ctypes as C, numpy as np, time
sys.path.insert(0, '.')
from oracle import oracle_c, oracle_np as o
from tools import tracking_protocol as tp
from tests.test_emu_sigma_dense import _emu_lib, _zolo_table
N, H, LAM = 512, int(sys.argv[3]) if len(sys.argv) > 3 else 50, 0.01
steps = int(sys.argv[1]); ep = int(sys.argv[2])
emu = _emu_lib(); tab = _zolo_table(); VAR = int(sys.argv[4]) if len(sys.argv) > 4 else 3
p = o.EnvParams()
s = o.reset_env(tp.TASK, p, np.random.default_rng(tp.episode_seeds(ep)[0]), dtype=np.float32, zero_disturb=False)
noise, eps_rng = tp.episode_noise(ep, steps), tp.episode_eps_rng(ep)
mean = o.hover_mean(H, p)
n = 4 * H
worst = 0; ks = []
for i in range(steps):
    eps = eps_rng.standard_normal((8192, 4 * 50)).astype(np.float32)[:N, :n]
    ns = o.noisy_state(s, p, tp.SeqRng(noise[i, :13]))
    a_mean = o.shift_mean(mean.astype(np.float32))
    R = oracle_c.hessian(ns, a_mean, p)
    Rs = (0.5 * (R + R.T)).astype(np.float32).astype(np.float64)
    lam = np.linalg.eigvalsh(Rs)
    S_ref = o.optimize_sigma(Rs, 0.5, np.float64)
    cov = np.full((n, n), np.nan, np.float32); scal = np.zeros(4); status = np.zeros(1, np.int32)
    t0 = time.time()
    rc = emu.emu_sigma_dense(n, C.c_float(0.5), R.ctypes.data_as(C.POINTER(C.c_float)), tab.ctypes.data_as(C.POINTER(C.c_double)),
                             cov.ctypes.data_as(C.POINTER(C.c_float)), scal.ctypes.data_as(C.POINTER(C.c_double)),
                             status.ctypes.data_as(C.POINTER(C.c_int)), VAR)
    err = np.linalg.norm(cov - S_ref) / np.linalg.norm(S_ref) if VAR == 3 else 0.0
    worst = max(worst, err); ks.append(scal[3])
    if status[0]: np.save(f'/tmp/bad_R_{i}.npy', R)
    print(f"step {i:3d} k {int(scal[3]):2d} status {status[0]} dl {scal[0]-lam[0]:+.2e} gap {lam[1]-lam[0]:.2e} W {lam[-1]-lam[0]:.0f} Sigma err {err:.2e}  ({time.time()-t0:.1f}s)", flush=True)
    cov32 = o.optimize_sigma(R, 0.5, dtype=np.float32)
    if VAR == 3: print("      float32 LAPACK oracle err", np.linalg.norm(cov32 - S_ref) / np.linalg.norm(S_ref), "asym", np.abs(R-R.T).max())
    L = np.linalg.cholesky(cov32.astype(np.float64)).astype(np.float32)
    a_s = o.sample_actions(a_mean, L, eps)
    cost = oracle_c.rollout_costs(ns, a_s, p)
    mean, _ = o.softmax_update(a_mean, a_s, cost, LAM)
    s, _, _, _ = o.env_step(s, mean[0], p, tp.SeqRng(noise[i + 1, 13:16]), "none")
print("worst Sigma err", worst, "mean k", np.mean(ks), "hist", np.bincount(np.array(ks, int))[16:])
