"""Accuracy of the covariance kernels along a closed loop (development tool, GPU box): for the Hessians met in the
first steps of the headline episode, device optimize_sigma / tridiagonal spectrum vs the float64 oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as o  # noqa: E402
from covo_mpc_b200 import _lib  # noqa: E402


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    n = 4 * H
    p = o.EnvParams()
    rng = np.random.default_rng(11)
    s = o.reset_env("tracking_zigzag", p, rng, dtype=np.float32, zero_disturb=True)
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len = _lib.MODE_COVO_ONLINE, 1024, H, s.pos_traj.shape[0]
    h = _lib.Handle(cfg)
    h.set_reference(s.pos_traj[None], s.vel_traj[None])
    mean = o.hover_mean(H, p)
    for i in range(steps):
        ns = o.noisy_state(s, p, np.random.default_rng(1000 + i))
        eps = np.random.default_rng(2000 + i).standard_normal((1024, n)).astype(np.float32)
        R = h.hessian(o.state_to_vec24(ns), [ns.time], o.shift_mean(mean)[None] if i else mean[None])[0]
        S = h.optimize_sigma(R[None])[0]
        d, e, sc = h.debug_tridiag()
        So = o.optimize_sigma(R.astype(np.float64), 0.5, np.float64)
        T = np.diag(d) + np.diag(e[:-1], 1) + np.diag(e[:-1], -1)
        w = np.linalg.eigvalsh(((R + R.T) / 2).astype(np.float64))
        wt = np.linalg.eigvalsh(T)
        print(f"step {i}: |S-So|/|So| = {np.linalg.norm(S - So) / np.linalg.norm(So):.3e}  max|eig(T)-eig(R)| = "
              f"{np.abs(wt - w).max():.3e}  lam_min dev {sc[0]:.8f} ref {w[0]:.8f}  finite {np.isfinite(S).all()} status {h.status()[0]}")
        a, mean, _, _ = o.covo_call(ns, mean, eps, p, lam=0.01)
        s, _, _, _ = o.env_step(s, a, p, rng, "none")


if __name__ == "__main__":
    main()
