"""Hessians of a CoVO-online closed loop on which a FIXED 24-step Lanczos iteration fails (the smallest Ritz value is still 1e-4 .. 6e-2
above lambda_min, so A = R - lambda_min + 1e-2 would be indefinite or Sigma off by per cents): regression inputs for the adaptive
Lanczos kernel of csrc/sigma_dense.cu.  Oracle closed loop (N = 1024, H = 50, episode 0 of tools/tracking_protocol.py), steps
10, 54, 55, 57 -> tests/golden/hessians/hard_hessians_n200.npz.  TEST INFRASTRUCTURE.

    python tests/golden/make_hard_hessians.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_c, oracle_np as o  # noqa: E402
from tools import tracking_protocol as tp  # noqa: E402

N, H, LAM, KEEP = 1024, 50, 0.01, (10, 54, 55, 57)


def main():
    p = o.EnvParams()
    s = o.reset_env(tp.TASK, p, np.random.default_rng(tp.episode_seeds(0)[0]), dtype=np.float32, zero_disturb=False)
    noise, eps_rng = tp.episode_noise(0, 60), tp.episode_eps_rng(0)
    mean = o.hover_mean(H, p)
    out = []
    for i in range(max(KEEP) + 1):
        eps = eps_rng.standard_normal((8192, 4 * H)).astype(np.float32)[:N]
        ns = o.noisy_state(s, p, tp.SeqRng(noise[i, :13]))
        a_mean = o.shift_mean(mean.astype(np.float32))
        R = oracle_c.hessian(ns, a_mean, p)
        if i in KEEP:
            out.append(R.astype(np.float32))
        cov = o.optimize_sigma(R, 0.5, dtype=np.float32)
        L = np.linalg.cholesky(cov.astype(np.float64)).astype(np.float32)
        cost = oracle_c.rollout_costs(ns, o.sample_actions(a_mean, L, eps), p)
        mean, _ = o.softmax_update(a_mean, o.sample_actions(a_mean, L, eps), cost, LAM)
        s, _, _, _ = o.env_step(s, mean[0], p, tp.SeqRng(noise[i + 1, 13:16]), "none")
    np.savez_compressed(os.path.join(HERE, "hessians", "hard_hessians_n200.npz"), R=np.stack(out), steps=np.array(KEEP))
    for R in out:
        lam = np.linalg.eigvalsh(0.5 * (R + R.T).astype(np.float64))
        print("lambda_min %.5f gap %.3e width %.1f" % (lam[0], lam[1] - lam[0], lam[-1] - lam[0]))


if __name__ == "__main__":
    main()
