"""Four CoVO-online MPC steps at the headline size (N=8192, H=50) and nothing else: the target of the ncu captures
(`-s 27 -c 9` = the 9 kernels of the fourth step).  GPU box only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as o  # noqa: E402
from covo_mpc_b200 import _lib  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    p = o.EnvParams()
    rng = np.random.default_rng(100)
    s = o.reset_env("tracking_zigzag", p, rng, dtype=np.float32, zero_disturb=True)
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len = _lib.MODE_COVO_ONLINE, N, H, s.pos_traj.shape[0]
    h = _lib.Handle(cfg)
    h.set_reference(s.pos_traj[None], s.vel_traj[None])
    for i in range(4):
        ns = o.noisy_state(s, p, rng)
        a = h.step(o.state_to_vec24(ns), [ns.time])[0]
        s, _, _, _ = o.env_step(s, a, p, rng, "none")
    print("ok", a)


if __name__ == "__main__":
    main()
