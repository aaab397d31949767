"""In-kernel phase breakdown of one CoVO-online MPC step (clock64 stamps written by block 0 of every kernel)
next to the per-kernel CUDA-event times.  Development tool: `python tools/phase_profile.py [N] [H]` on a GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from covo_mpc_b200 import _lib  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    import torch

    bench.N_SAMPLES, bench.HORIZON = N, H
    env, states, times, traj = bench.record_states(12, 100, "covo-online", device=0)
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.device = _lib.MODE_COVO_ONLINE, N, H, int(traj[0].shape[0]), 0
    h = _lib.Handle(cfg)
    h.set_reference(traj[0][None], traj[1][None])
    for i in range(4):
        h.step(states[i], times[i:i + 1])
    h.set_profiling(True)
    acc = np.zeros(6)
    for i in range(4, 10):
        h.step(states[i], times[i:i + 1])
        acc += h.kernel_ms()
    h.set_profiling(False)
    names = ["hessian (3 kernels)", "E1 tridiag (cluster)", "E2 trifunc", "E3 sandwich", "cholesky", "rollout"]
    print("kernel ms (CUDA events, mean of 6):", {k: round(float(v) / 6, 4) for k, v in zip(names, acc)})
    h.phase_clocks(True)
    h.step(states[10], times[10:11])
    torch.cuda.synchronize()
    c = h.phase_clocks(True, read=True)
    mhz = 1965.0

    def d(a, b):
        return (c[b] - c[a]) / mhz

    print("clock64 deltas in us @1965 MHz")
    print("hess_assemble: ", [round(d(i, i + 1), 2) for i in range(0, 5)], " local end stamp", c[6] - c[0])
    print("E1 (register-resident): load/sym", round(d(8, 15), 2), "loop", round(d(15, 9), 2), "store+WY", round(d(9, 16), 2))
    print("E1 loop sections (warp 0 of CTA 0, us): scalar chain + rank-2 + publish", round(c[40] / mhz, 2), "matvec+transpose+send",
          round(c[42] / mhz, 2), "mbarrier wait", round(c[43] / mhz, 2))
    print("E2: load+gersh", round(d(17, 10), 2), "multisect", round(d(10, 11), 2), "pivots", round(d(11, 12), 2),
          "generators", round(d(12, 13), 2), "F rows", round(d(13, 14), 2))
    print("cholesky: load", round(d(23, 24), 2), "first panel", round(d(24, 26), 2), "panels", round(d(26, 25), 2),
          "store", round(d(25, 27), 2))
    print("rollout: ", [round(d(i, i + 1), 2) for i in range(32, 39)])


if __name__ == "__main__":
    main()
