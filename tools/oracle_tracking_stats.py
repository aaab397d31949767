"""Oracle arm of the powered closed-loop tracking comparison: 40 episodes x 300 steps of the reference algorithm (oracle/ port,
C/OpenMP heavy loops) at the headline size, protocol in tools/tracking_protocol.py.  CPU only (~20 min on 8 cores for covo-online);
writes tests/golden/tracking/oracle_tracking_{controller}_N{N}_H{H}.npz with per-episode, per-step err_pos and actions.

    python tools/oracle_tracking_stats.py [--controller covo-online|mppi] [--episodes 40] [--N 8192] [--H 50]

TEST INFRASTRUCTURE: the fixture it writes is what bench.py's `tracking_cost` and tests/test_tracking_gpu.py compare the device with."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_c, oracle_np as o  # noqa: E402
from tools import tracking_protocol as tp  # noqa: E402


def mppi_step(ns, mean_prev, a_cov_blk, eps, p, lam):
    """MPPIController.__call__ (controllers/mppi.py:28-134) with gamma_sigma = 0 (the covariance stays 0.25 I, so its shift is the
    identity): oracle_np.mppi_call with the rollout loop in C."""
    a_mean = o.shift_mean(mean_prev.astype(np.float32))
    Lblk = np.linalg.cholesky(a_cov_blk.astype(np.float64)).astype(np.float32)
    a_s = o.sample_actions_blockdiag(a_mean, Lblk, eps)
    cost = oracle_c.rollout_costs(ns, a_s, p)
    new_mean, _ = o.softmax_update(a_mean, a_s, cost, lam)
    return new_mean[0].copy(), new_mean


def run_episode(k, controller, N, H, lam, n_steps):
    p = o.EnvParams()
    traj_seed, _, _ = tp.episode_seeds(k)
    s = o.reset_env(tp.TASK, p, np.random.default_rng(traj_seed), dtype=np.float32, zero_disturb=False)
    noise = tp.episode_noise(k, n_steps)
    eps_rng = tp.episode_eps_rng(k)
    mean = o.hover_mean(H, p)
    a_cov_blk = np.tile(np.eye(4, dtype=np.float32) * 0.25, (H, 1, 1))
    errs, acts = np.zeros(n_steps, np.float32), np.zeros((n_steps, 4), np.float32)
    ns = o.noisy_state(s, p, tp.SeqRng(noise[0, :13]))
    for i in range(n_steps):
        if controller == "mppi":
            eps = eps_rng.standard_normal((N, H, 4)).astype(np.float32)
            u, mean = mppi_step(ns, mean, a_cov_blk, eps, p, lam)
        else:
            eps = eps_rng.standard_normal((N, 4 * H)).astype(np.float32)
            u, mean = oracle_c.covo_step(ns, mean, eps, p, lam)
        acts[i] = u
        s, _, _, e = o.env_step(s, u, p, tp.SeqRng(noise[i + 1, 13:16]), "none")
        errs[i] = e
        ns = o.noisy_state(s, p, tp.SeqRng(noise[i + 1, :13]))
    return errs, acts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--controller", default="covo-online")
    ap.add_argument("--episodes", type=int, default=tp.N_EPISODES)
    ap.add_argument("--steps", type=int, default=tp.EP_STEPS)
    ap.add_argument("--N", type=int, default=8192)
    ap.add_argument("--H", type=int, default=50)
    ap.add_argument("--lam", type=float, default=0.01)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    out = a.out or os.path.join(ROOT, "tests", "golden", "tracking", f"oracle_tracking_{a.controller}_N{a.N}_H{a.H}.npz")
    errs = np.zeros((a.episodes, a.steps), np.float32)
    acts = np.zeros((a.episodes, a.steps, 4), np.float32)
    t0 = time.time()
    for k in range(a.episodes):
        errs[k], acts[k] = run_episode(k, a.controller, a.N, a.H, a.lam, a.steps)
        print(f"episode {k}: mean err_pos {errs[k].mean():.5f} (first 100: {errs[k, :100].mean():.5f})  [{time.time() - t0:.0f} s]", flush=True)
        np.savez_compressed(out, err_pos=errs[:k + 1], actions=acts[:k + 1], N=a.N, H=a.H, lam=a.lam, controller=a.controller,
                            threads=oracle_c.num_threads())
    m = errs.mean(axis=1)
    print(f"{a.controller}: mean {m.mean():.5f} +- {m.std(ddof=1) / np.sqrt(len(m)):.5f} (s.e.), episode std {m.std(ddof=1):.5f}")


if __name__ == "__main__":
    main()
