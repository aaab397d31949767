"""Spread of the closed-loop tracking error over noise seeds, device loop (development tool, GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as o  # noqa: E402
from covo_mpc_b200 import _lib  # noqa: E402
import covo_mpc_b200.env as envmod  # noqa: E402


def main():
    N, H = 8192, 50
    p = o.EnvParams()
    env = envmod.Quad3D("tracking_zigzag")
    _, _, st = env.reset(np.random.default_rng(100))
    for mode, name in ((_lib.MODE_COVO_ONLINE, "covo-online"), (_lib.MODE_MPPI, "mppi")):
        for seed in range(1, 7):
            cfg = _lib.default_config()
            cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.seed = mode, N, H, st.pos_traj.shape[0], seed
            h = _lib.Handle(cfg)
            h.set_reference(st.pos_traj[None], st.vel_traj[None])
            h.set_mean(o.hover_mean(H, p)[None])
            s0 = np.zeros(24, np.float32)
            s0[6] = 1.0
            s0[16:19] = st.pos_traj[0]
            s0[19:22] = st.vel_traj[0]
            h.env_reset(s0[None], [0])
            _, _, err = h.closed_loop(100, noise_seed=1000 + seed)
            print(name, seed, float(err.mean()), float(err[:50].mean()), float(err[50:].mean()))
            h.close()


if __name__ == "__main__":
    main()
