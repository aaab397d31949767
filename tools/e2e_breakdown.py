"""Where the plugin call's time goes (development tool, GPU box): device time of the step, the bare C-ABI call covo_step with host
buffers, the ctypes wrapper Handle.step_state, and the full Controller.__call__."""
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

    K = 300
    states_h, times_h, trajs, _ = bench.synthetic_states(290, 100)
    traj = trajs[0]
    env = cm.Quad3D(bench.TASK)
    ctl, cp = cm.get_controller(env, "covo-online", f"N{bench.N_SAMPLES}_H{bench.HORIZON}_lam{bench.LAM}")
    f32 = np.float32
    st0 = cm.EnvState3D(pos=np.zeros(3, f32), vel=np.zeros(3, f32), quat=np.array([0, 0, 0, 1], f32), omega=np.zeros(3, f32),
                        pos_traj=traj[0], vel_traj=traj[1], acc_traj=np.zeros_like(traj[0]), pos_tar=np.zeros(3, f32),
                        vel_tar=np.zeros(3, f32), acc_tar=np.zeros(3, f32), time=0, f_disturb=np.zeros(3, f32))
    hs = []
    for i in range(290):
        s = states_h[i]
        hs.append(st0.replace(pos=s[0:3], quat=s[3:7], vel=s[7:10], omega=s[10:13], f_disturb=s[13:16], pos_tar=s[16:19], vel_tar=s[19:22],
                              time=int(times_h[i])))
    for i in range(5):
        _, cp, _ = ctl(None, hs[i], env.default_params, None, cp, {"noisy_state": hs[i]})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        _, cp, _ = ctl(None, hs[i % 290], env.default_params, None, cp, {"noisy_state": hs[i % 290]})
    t_ctl = (time.perf_counter() - t0) / K
    h = ctl._handle
    t0 = time.perf_counter()
    for i in range(K):
        h.step_state(hs[i % 290])
    t_wrap = (time.perf_counter() - t0) / K
    lib = h.lib
    s = np.ascontiguousarray(states_h[3]); t = np.ascontiguousarray(times_h[3:4]); out = np.zeros(4, f32)
    sp, tp, op = _lib.fptr(s), _lib.iptr(t), _lib.fptr(out)
    t0 = time.perf_counter()
    for i in range(K):
        lib.covo_step(h._h, sp, tp, None, op)
    t_c = (time.perf_counter() - t0) / K
    # device time of the same step (events around step_device, no flush)
    sd = torch.from_numpy(states_h).cuda(); td = torch.from_numpy(times_h).cuda(); ad = torch.zeros((290, 4), device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for i in range(5):
        h.step_device(sd.data_ptr() + 96 * i, td.data_ptr() + 4 * i, 0, ad.data_ptr() + 16 * i, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        j = i % 290
        h.step_device(sd.data_ptr() + 96 * j, td.data_ptr() + 4 * j, 0, ad.data_ptr() + 16 * j, stream)
    e1.record()
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / K * 1e-3
    print(f"device (back-to-back graph launches, warm L2): {t_dev * 1e6:.1f} us | covo_step (C-ABI, host buffers): {t_c * 1e6:.1f} us | "
          f"Handle.step_state: {t_wrap * 1e6:.1f} us | Controller.__call__: {t_ctl * 1e6:.1f} us")


if __name__ == "__main__":
    main()
