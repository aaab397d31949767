"""Shared scenario builders for the tests (oracle side)."""
import numpy as np

from oracle import oracle_np as o


def scenario(task="tracking_zigzag", seed=0, H=50, time=0, warm_steps=0, zero_disturb=True, dtype=np.float32):
    """A noisy state on a reference trajectory after `warm_steps` of hover-thrust flight, plus a nominal mean."""
    p = o.EnvParams()
    rng = np.random.default_rng(seed)
    s = o.reset_env(task, p, rng, dtype=np.float64, zero_disturb=zero_disturb)
    hover = o.hover_mean(1, p, np.float64)[0]
    for _ in range(warm_steps):
        a = hover + rng.normal(0, 0.2, 4)
        s, _, _, _ = o.env_step(s, a, p, rng, "none")
    if time:
        s.time = int(time)
        ti = min(time, s.pos_traj.shape[0] - 1)
        s.pos_tar = [np.float64(x) for x in s.pos_traj[ti]]
        s.vel_tar = [np.float64(x) for x in s.vel_traj[ti]]
    ns = o.noisy_state(s, p, rng)
    ns32 = o.make_state(ns.pos, ns.quat, ns.vel, ns.omega, ns.f_disturb, ns.time, ns.pos_traj, ns.vel_traj, ns.pos_tar,
                        ns.vel_tar, dtype=dtype)
    a_mean = o.hover_mean(H, p, np.float32) + rng.normal(0, 0.1, (H, 4)).astype(np.float32)
    return p, ns32, a_mean, rng


def pack_lt(L, n_pad=None):
    """Python restatement of the packed k-major factor layout (csrc/common.cuh: lt_col_offset)."""
    n = L.shape[0]
    n_pad = n_pad or (n + 7) // 8 * 8
    out = []
    for k in range(n):
        rs = k & ~7
        col = np.zeros(n_pad - rs, np.float32)
        col[k - rs:n - rs] = L[k:, k]
        out.append(col)
    return np.concatenate(out)
