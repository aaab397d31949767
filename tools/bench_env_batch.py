"""BASELINE config 5 on ONE GPU: E environments batched behind one handle (CoVO-online, N=1024, H=50); with G GPUs
each rank runs the same thing on its own E environments (no collective), so the 4096-environment configuration is
8 x this at E = 512.  Prints one JSON line.  `python tools/bench_env_batch.py [E] [steps]` on a GPU box."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as o  # noqa: E402
from covo_mpc_b200 import _lib  # noqa: E402
import covo_mpc_b200.env as envmod  # noqa: E402


def main():
    import torch

    E = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    N, H = 1024, 50
    p = o.EnvParams()
    env = envmod.Quad3D("tracking_zigzag")
    trajs = [env.reset(np.random.default_rng(100 + (e % 16)))[2] for e in range(min(E, 16))]
    pos = np.stack([trajs[e % 16].pos_traj for e in range(E)])
    vel = np.stack([trajs[e % 16].vel_traj for e in range(E)])
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.n_env, cfg.seed = _lib.MODE_COVO_ONLINE, N, H, pos.shape[1], E, 3
    h = _lib.Handle(cfg)
    h.set_reference(pos, vel)
    h.set_mean(np.tile(o.hover_mean(H, p)[None], (E, 1, 1)))
    s0 = np.zeros((E, 24), np.float32)
    s0[:, 6] = 1.0
    s0[:, 16:19] = pos[:, 0]
    s0[:, 19:22] = vel[:, 0]
    h.env_reset(s0, np.zeros(E, np.int32))
    h.closed_loop(3, noise_seed=1)  # warm-up
    torch.cuda.synchronize()
    h.set_profiling(True)
    t0 = time.perf_counter()
    _, _, err = h.closed_loop(K, noise_seed=2)
    dt = time.perf_counter() - t0
    km = h.kernel_ms()
    out = {"metric": "env_mpc_steps_per_sec", "value": E * K / dt, "unit": "env-steps/s", "n_gpus": 1, "steps": K,
           "ms_per_batched_step": 1e3 * dt / K, "config": {"workload": f"covo-online tracking_zigzag, {E} envs x N={N} H={H} per GPU, "
                                                           "device-resident closed loop (controller + env step)"},
           "kernel_ms_last_step": {k: float(v) for k, v in zip(["hessian", "tridiag", "trifunc", "sandwich", "cholesky", "rollout"], km)},
           "mean_err_pos": float(err.mean()), "status_ok": bool((h.status() == 0).all())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
