"""In-kernel phase breakdown (clock64 stamps of CTA 0) of the tridiagonalisation-free optimize_sigma kernels (csrc/sigma_dense.cu), next
to the per-kernel CUDA-event times.  Development tool: `COVO_SIGMA=dense python tools/dense_profile.py` on a GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("COVO_SIGMA", "dense")

import bench  # noqa: E402
from covo_mpc_b200 import _lib  # noqa: E402


def main():
    import torch

    env, states, times, traj = bench.record_states(12, 100, "covo-online", device=0)
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.device = _lib.MODE_COVO_ONLINE, 8192, 50, int(traj[0].shape[0]), 0
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
    print(os.environ["COVO_SIGMA"], "kernel us:", [round(float(v) / 6 * 1e3, 1) for v in acc])
    h.phase_clocks(True)
    h.step(states[10], times[10:11])
    torch.cuda.synchronize()
    c = h.phase_clocks(True, read=True)
    mhz = 1965.0
    d = lambda a, b: round((c[b] - c[a]) / mhz, 2)
    print("lanczos: load", d(48, 49), "iterations", d(49, 50), "multisection", d(50, 51), "us")
    print("lanczos exchange (send + mbarrier wait, all iterations, CTA 0 thread 0):", round(c[52] / mhz, 2), "us")
    us = lambda k: round(c[k] / mhz, 2)
    print("lanczos recurrence (CTA 0 thread 0, sums over", int(c[45]), "steps, us): matvec + reductions", us(40), "sends + |v|^2", us(41), "exchange wait", us(42),
          "alpha / beta", us(43), "next vector + verdict + barrier", us(44))
    print("gjb (CTA 0, sums over the 25 steps, us): solver warp 0: wait for the pivot block", us(54), "8x8 inversion", us(55), "wait for the rows + P^-1 barrier", us(60),
          "multipliers", us(56), "step barrier", us(57), "| update warp 0: at step barrier", us(58), "open -> step done", us(59))
    print("gjb look-ahead on the owning CTA (its own clock, sums over 24 blocks, us): pivot columns", us(62), "pivot block to the copy unit", us(63), "rest of the row", us(61))
    print("gjb bulk rank-8 update alone (update warp 0, sum over 25 steps, us):", us(53))
    print("raw stamps 48..63:", [int(x) for x in c[48:64]])


if __name__ == "__main__":
    main()
