"""Development tool: Lanczos step counts and per-step device time of the dense optimize_sigma path along (a) the bench's replayed states,
(b) the plugin call loop (is the handle still on the dense path at the end?), (c) the device-resident closed loop."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import covo_mpc_b200 as cm  # noqa: E402
from covo_mpc_b200 import _lib  # noqa: E402


def main():
    import torch

    K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    states_h, times_h, trajs, _ = bench.synthetic_states(290, 100)
    traj = trajs[0]
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.device = _lib.MODE_COVO_ONLINE, 8192, 50, int(traj[0].shape[0]), 0
    cfg.lam, cfg.seed = bench.LAM, 100
    h = _lib.Handle(cfg)
    h.set_reference(traj[0][None], traj[1][None])
    dev = torch.device("cuda:0")
    states, times = torch.from_numpy(states_h).to(dev), torch.from_numpy(times_h).to(dev)
    actions = torch.zeros((290, 4), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    steps, ms = [], []
    for i in range(K):
        j = i % 290
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        h.step_device(states.data_ptr() + 96 * j, times.data_ptr() + 4 * j, 0, actions.data_ptr() + 16 * j, stream)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
        steps.append(h.debug_tridiag()[2][3])
    steps, ms = np.array(steps), np.array(ms)
    print(f"(a) replayed states, warm L2: path {h.sigma_path()} status {h.status()}; Lanczos steps min {steps.min():.0f} mean {steps.mean():.1f} max {steps.max():.0f}; "
          f"step ms mean {ms[5:].mean():.4f} p99 {np.percentile(ms[5:], 99):.4f}")
    print("    steps histogram:", dict(zip(*np.unique(steps, return_counts=True))))
    env = cm.Quad3D(bench.TASK)
    ctl, cp = cm.get_controller(env, "covo-online", f"N8192_H50_lam{bench.LAM}", device=0, seed=100)
    f32 = np.float32
    st0 = cm.EnvState3D(pos=np.zeros(3, f32), vel=np.zeros(3, f32), quat=np.array([0, 0, 0, 1], f32), omega=np.zeros(3, f32),
                        pos_traj=traj[0], vel_traj=traj[1], acc_traj=np.zeros_like(traj[0]), pos_tar=np.zeros(3, f32),
                        vel_tar=np.zeros(3, f32), acc_tar=np.zeros(3, f32), time=0, f_disturb=np.zeros(3, f32))
    hs = [st0.replace(pos=s[0:3], quat=s[3:7], vel=s[7:10], omega=s[10:13], f_disturb=s[13:16], pos_tar=s[16:19], vel_tar=s[19:22], time=int(t))
          for s, t in zip(states_h, times_h)]
    for i in range(5):
        _, cp, _ = ctl(None, hs[i], env.default_params, None, cp, {"noisy_state": hs[i]})
    tcall = []
    for i in range(K):
        t0 = time.perf_counter()
        _, cp, _ = ctl(None, hs[(5 + i) % 290], env.default_params, None, cp, {"noisy_state": hs[(5 + i) % 290]})
        tcall.append(time.perf_counter() - t0)
    tcall = np.array(tcall) * 1e3
    hh = ctl._handle if hasattr(ctl, "_handle") else None
    print(f"(b) plugin call: mean {tcall.mean():.4f} ms, first 20 {tcall[:20].mean():.4f}, last 20 {tcall[-20:].mean():.4f}; sigma path at the end:",
          hh.sigma_path() if hh is not None else "?")
    ctl.close()
    h.set_mean(bench._hover())
    h.env_reset(states_h[0][None], times_h[:1])
    h.closed_loop(5, noise_seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h.closed_loop(K, noise_seed=2)
    print(f"(c) closed loop: {(time.perf_counter() - t0) / K * 1e3:.4f} ms/step; path {h.sigma_path()} status {h.status()} last Lanczos steps {h.debug_tridiag()[2][3]:.0f}")


if __name__ == "__main__":
    main()
