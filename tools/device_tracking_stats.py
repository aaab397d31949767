"""Device arms of the powered closed-loop tracking comparison (protocol: tools/tracking_protocol.py; oracle arm: the fixture written
by tools/oracle_tracking_stats.py).  GPU box only.

  arm "production":  40 episodes x 300 steps as ONE batched handle (E = 40, per-environment reference trajectory), the whole closed
                     loop on the device (covo_closed_loop), the SAME observation-noise stream per episode as the oracle arm, the
                     device's own Philox sample field.  Statistical comparison: mean +- s.e. over episodes and the z-score of the
                     difference of means, plus the paired per-episode differences (same trajectory, same noise).
  arm "identical":   the first K episodes one controller call at a time with the oracle arm's eps (numpy default_rng(9000 + k)) and
                     the host environment of the oracle: per-step action distance to the oracle's recorded actions, first diverging
                     step, and the paired err_pos difference.

    python tools/device_tracking_stats.py [--controller covo-online] [--identical 8] > gpurun_out/r2_tracking_stats.json"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import tracking_protocol as tp  # noqa: E402


def fixture_path(controller, N, H):
    return os.path.join(ROOT, "tests", "golden", "tracking", f"oracle_tracking_{controller}_N{N}_H{H}.npz")


def initial_states(n_ep):
    from oracle import oracle_np as o

    p = o.EnvParams()
    return [o.reset_env(tp.TASK, p, np.random.default_rng(tp.episode_seeds(k)[0]), dtype=np.float32, zero_disturb=False) for k in range(n_ep)]


def production_arm(controller, N, H, lam, n_ep, n_steps, seed=2024, device=0):
    """Returns err_pos [n_ep][n_steps] of the batched device closed loop."""
    from covo_mpc_b200 import _lib
    from oracle import oracle_np as o

    s0 = initial_states(n_ep)
    cfg = _lib.default_config()
    cfg.mode = {"covo-online": _lib.MODE_COVO_ONLINE, "mppi": _lib.MODE_MPPI}[controller]
    cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.n_env, cfg.lam, cfg.seed, cfg.device = N, H, s0[0].pos_traj.shape[0], n_ep, lam, seed, device
    h = _lib.Handle(cfg)
    h.set_reference(np.stack([s.pos_traj for s in s0]), np.stack([s.vel_traj for s in s0]))
    h.env_reset(np.stack([o.state_to_vec24(s) for s in s0]), [0] * n_ep)
    noise = np.stack([tp.episode_noise(k, n_steps) for k in range(n_ep)], axis=1)  # [n_steps + 1][E][16]
    _, _, err = h.closed_loop(n_steps, noise=noise)
    status = h.status()
    h.close()
    return err.T.copy(), status


def identical_arm(controller, N, H, lam, k, n_steps, oracle_actions, device=0):
    from covo_mpc_b200 import _lib
    from oracle import oracle_np as o

    p = o.EnvParams()
    s = initial_states(k + 1)[k]
    cfg = _lib.default_config()
    cfg.mode = {"covo-online": _lib.MODE_COVO_ONLINE, "mppi": _lib.MODE_MPPI}[controller]
    cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.lam, cfg.device = N, H, s.pos_traj.shape[0], lam, device
    h = _lib.Handle(cfg)
    h.set_reference(s.pos_traj[None], s.vel_traj[None])
    noise, eps_rng = tp.episode_noise(k, n_steps), tp.episode_eps_rng(k)
    errs, dist = np.zeros(n_steps, np.float32), np.zeros(n_steps, np.float32)
    ns = o.noisy_state(s, p, tp.SeqRng(noise[0, :13]))
    for i in range(n_steps):
        eps = eps_rng.standard_normal((N, 4 * H)).astype(np.float32)
        u = h.step(o.state_to_vec24(ns), [ns.time], eps[None])[0]
        dist[i] = np.abs(u - oracle_actions[i]).max()
        s, _, _, e = o.env_step(s, u, p, tp.SeqRng(noise[i + 1, 13:16]), "none")
        errs[i] = e
        ns = o.noisy_state(s, p, tp.SeqRng(noise[i + 1, :13]))
    h.close()
    return errs, dist


def summarise(dev_err, ora_err):
    """Means over episodes of the per-episode mean err_pos (the reference's metric, envs/quadrotor.py:573-579)."""
    n = min(len(dev_err), len(ora_err))
    d, r = dev_err[:n].mean(axis=1), ora_err[:n].mean(axis=1)
    se = lambda x: float(x.std(ddof=1) / np.sqrt(len(x))) if len(x) > 1 else float("nan")
    diff = d - r
    return {"episodes": int(n), "steps": int(dev_err.shape[1]),
            "device_mean": float(d.mean()), "device_se": se(d), "oracle_mean": float(r.mean()), "oracle_se": se(r),
            "rel_delta": float((d.mean() - r.mean()) / r.mean()),
            "z_unpaired": float((d.mean() - r.mean()) / np.sqrt(se(d) ** 2 + se(r) ** 2)) if n > 1 else None,
            "paired_mean_diff": float(diff.mean()), "paired_se": se(diff), "z_paired": float(diff.mean() / se(diff)) if n > 1 else None,
            "first100": {"device_mean": float(dev_err[:n, :100].mean()), "oracle_mean": float(ora_err[:n, :100].mean())}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--controller", default="covo-online")
    ap.add_argument("--N", type=int, default=8192)
    ap.add_argument("--H", type=int, default=50)
    ap.add_argument("--identical", type=int, default=8)
    a = ap.parse_args()
    fx = np.load(fixture_path(a.controller, a.N, a.H))
    ora_err, ora_act = fx["err_pos"], fx["actions"]
    n_ep, n_steps = ora_err.shape
    dev_err, status = production_arm(a.controller, a.N, a.H, float(fx["lam"]), n_ep, n_steps)
    out = {"controller": a.controller, "N": a.N, "H": a.H, "protocol": "tools/tracking_protocol.py",
           "production": summarise(dev_err, ora_err), "status_nonzero": int(np.count_nonzero(status))}
    ident = []
    for k in range(min(a.identical, n_ep)):
        errs, dist = identical_arm(a.controller, a.N, a.H, float(fx["lam"]), k, n_steps, ora_act[k])
        far = np.nonzero(dist >= 5e-4)[0]
        ident.append({"episode": k, "first_diverging_step": int(far[0]) if len(far) else None, "device_mean_err": float(errs.mean()),
                      "oracle_mean_err": float(ora_err[k].mean()), "prefix_rel_delta": None if not len(far) or far[0] == 0 else
                      float(abs(errs[:far[0]].sum() - ora_err[k, :far[0]].sum()) / ora_err[k, :far[0]].sum())})
    if ident:
        d = np.array([x["device_mean_err"] for x in ident]) - np.array([x["oracle_mean_err"] for x in ident])
        out["identical_eps"] = {"episodes": ident, "paired_mean_diff": float(d.mean()),
                                "paired_se": float(d.std(ddof=1) / np.sqrt(len(d))) if len(d) > 1 else None}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
