#!/usr/bin/env python
"""bench.py -- MPC control steps/sec of the CoVO-MPC hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

Workload (config.workload): CoVO-online, tracking_zigzag, N=8192 samples, H=50, u_dim=4 -- BASELINE.json's
headline single-GPU configuration.  One "step" = one controller call (shift -> exact Hessian -> optimize_sigma
-> Cholesky -> sample -> N x H rollout -> softmax update) for one environment, driven over the noisy states of
a closed-loop episode recorded beforehand (synthetic zigzag reference trajectory, observation noise as in
envs/quadrotor.py:323-351).

  value  : states / times resident in HBM, covo_step_device on the launch stream, CUDA events per step,
           L2 flushed (256 MiB memset) between steps and excluded from the timing.
  e2e    : the same steps through the reference-facing plugin call controller(obs, state, params, rng,
           control_params, info) with HOST numpy state: H2D (pinned) + kernels + D2H inside the timed region.
  N > 1  : one process per GPU (torch.distributed / NCCL for barrier + max-over-ranks), one independent environment per rank
           (no data-path collective, weak scaling).
  The same JSON line carries the BASELINE configurations the replica number does not show:
    nsample_shard : config 4 -- CoVO-offline, the N samples of ONE environment split over the ranks with one exchange of the
                    (min cost, sum w, sum w u) records per step; device-timed and wall, with the strong-scaling efficiency against
                    the same handle at world = 1 run in the same process, at N = 8192 and N = 65536;
    env_batch     : config 5 -- 512 environments per GPU x N = 1024, device-resident closed loop, per-kernel times;
    tracking_cost : 40 x 300-step episodes, device vs the oracle fixture (mean +- s.e., z-scores);
    cpu_baseline  : the oracle port on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

_REAL_STDOUT = sys.stdout  # replaced by main(): the JSON line goes to the real stdout, everything else to stderr

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES, HORIZON, LAM, TASK = 8192, 50, 0.01, "tracking_zigzag"
METRIC, UNIT = "mpc_control_steps_per_sec", "steps/s"


# ---------------------------------------------------------------------------------------------------
def synthetic_states(n_states: int, seed: int):
    """The inputs BOTH arms time: noisy states along closed-loop episodes of the reference's own PID policy (gains of
    envs/quadrotor.py:692-699) on zigzag reference trajectories, observation noise as envs/quadrotor.py:323-351, disturb none.
    Generated on the host without any kernel under test (oracle/ restatement of the environment, untimed) so that `--impl b200`
    and `--impl reference` see bit-identical states (same inputs / seeds).  Returns (states [n][24], times [n],
    (pos_traj, vel_traj) of each state's episode, episode index per state)."""
    from oracle import oracle_np as o

    p = o.EnvParams()
    rng = np.random.default_rng(seed)
    states, times, trajs, ep_of = [], [], [], []
    while len(states) < n_states:
        s = o.reset_env(TASK, p, rng, dtype=np.float32, zero_disturb=True)
        trajs.append((np.ascontiguousarray(s.pos_traj, np.float32), np.ascontiguousarray(s.vel_traj, np.float32)))
        for _ in range(min(290, n_states - len(states))):
            ns = o.noisy_state(s, p, rng)
            states.append(o.state_to_vec24(ns))
            times.append(ns.time)
            ep_of.append(len(trajs) - 1)
            s, _, _, _ = o.env_step(s, o.pid_action(s, p, Kp=10.0, Kd=5.0, Ki=0.0, Kp_att=10.0), p, rng, "none")
    return np.stack(states).astype(np.float32), np.array(times, dtype=np.int32), trajs, ep_of


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def n_samples(self) -> int:
        return len(self.lines) if self.proc is not None else 1 << 30

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(n, H, N, parity=False):
    """SURVEY 8(d) / DESIGN.md: mandatory HBM bytes per MPC step, per kernel."""
    nn = 4 * n * n
    rollout = nn // 2 + 4 * 4 * n + 4 * (H + 1) * 6 + 96 + 2 * 4 * n + (4 * N * n if parity else 0)
    return {
        "lanczos": nn + nn,                          # R in; (R + R^T)/2 out for the pole kernels
        "pole_inverses": 14 * nn + 13 * nn // 2,     # every pole cluster (and the log det one) reads the matrix; 13 weighted inverses (triangles) out
        "combine": 13 * nn // 2 + nn,                # 13 triangles in; Sigma out
        # state, mean, ref; per-step derivative records written + read; [A|B], S, D hand-over written + read; R
        "hessian": 96 + 4 * n + 4 * H * 6 + 2 * 4 * H * (14 * 153 + 14 * 17) + 2 * 4 * H * 328 + nn,
        "tridiag": nn + nn + 2 * 8 * n,  # R in; reflectors out; (d, e) fp64 out (Q^T: qacc kernel, side stream, +2 nn)
        "trifunc": 2 * 8 * n + nn,       # (d, e) in; F out (full symmetric)
        "sandwich": nn + nn + nn,        # Q^T, F in; Sigma out
        "cholesky": nn + nn // 2,        # Sigma in; packed factor out
        "rollout": rollout,
    }


# ---------------------------------------------------------------------------------------------------
def record_states(n_states: int, seed: int, controller_name="covo-online", N=1024, device=0):
    """Compatibility shim for the tools: (env, states, times, (pos_traj, vel_traj)) of ONE synthetic episode."""
    import covo_mpc_b200 as cm

    st, tm, trajs, _ = synthetic_states(min(n_states, 290), seed)
    idx = np.arange(n_states) % len(st)
    return cm.Quad3D(TASK), st[idx], tm[idx], trajs[0]


def _hover(E=1):
    return np.tile(np.array([(0.027 * 9.81 / 0.8) * 2 - 1, 0, 0, 0], np.float32), (E, HORIZON, 1))


def _events(n):
    import torch

    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def bench_nsample_shard(dist, world, rank, dev, states, times, traj, K, W, flush, n_samples):
    """BASELINE config 4: CoVO-offline, the N samples of ONE environment split over the ranks, one exchange of the (min cost,
    sum w, sum w u) records per MPC step (controllers/covo.py:266-275 merged across ranks).  Device-timed (CUDA events, max over
    ranks) and wall-clock, next to the same handle type at world = 1 run in the same process: strong-scaling efficiency =
    t(world 1) / (world x t(world))."""
    import torch

    from covo_mpc_b200 import _lib

    stream = torch.cuda.current_stream().cuda_stream

    def make(r, w):
        cfg = _lib.default_config()
        cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.device = _lib.MODE_COVO_OFFLINE, n_samples, HORIZON, int(traj[0].shape[0]), dev.index
        cfg.lam, cfg.seed, cfg.rank, cfg.world = LAM, 100, r, w
        h = _lib.Handle(cfg)
        h.set_reference(traj[0][None], traj[1][None])
        h.reset_offline(states[0].cpu().numpy(), [0], 300)
        return h

    actions = torch.zeros((W + K, 4), dtype=torch.float32, device=dev)

    def timed(step_fn):
        for i in range(W):
            step_fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = _events(K)
        t0 = time.perf_counter()
        for i in range(K):
            flush.zero_()
            ev[i][0].record()
            step_fn(W + i)
            ev[i][1].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_ms = float(sum(a.elapsed_time(b) for a, b in ev))
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / K, float(t[1]) / K

    h1 = make(0, 1)
    one = lambda i: h1.step_device(states.data_ptr() + 96 * i, times.data_ptr() + 4 * i, 0, actions.data_ptr() + 16 * i, stream)
    d1, w1 = timed(one)
    a1_first = actions[0].clone()  # the first step: later ones carry the mean, and an arg-min near-tie flips on a 1e-7 difference
    out = {"n_samples": n_samples, "world1_device_ms_per_step": d1, "world1_wall_ms_per_step_incl_flush": w1}
    if world > 1:
        # (a) the exchange as a collective call between two kernels: kernel -> ncclAllGather -> kernel
        hs = make(rank, world)
        pbuf, pn = hs.partial_buffer()

        class _W:
            __cuda_array_interface__ = {"shape": (pn,), "typestr": "<f4", "data": (pbuf, False), "version": 2}

        part = torch.as_tensor(_W(), device=dev)
        gathered = torch.zeros((world, pn), dtype=torch.float32, device=dev)

        def sharded(i):
            hs.step_partial_device(states.data_ptr() + 96 * i, times.data_ptr() + 4 * i, 0, stream)
            dist.all_gather_into_tensor(gathered.view(-1), part)
            hs.step_merge_device(gathered.data_ptr(), actions.data_ptr() + 16 * i, stream)

        dw, ww = timed(sharded)
        a_nccl, a_nccl_first = actions[W:W + K].clone(), actions[0].clone()
        hs.close()
        # (b) the fused exchange: the rollout kernel's finalising CTA stores the record into every rank's buffer (peer memory mapped
        # through CUDA IPC) and raises a flag; the merge kernel behind it waits for the flags.  No collective, no host involvement.
        hf = make(rank, world)
        handle, _ = hf.exchange_info()
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        for r in range(world):
            if r != rank:
                hf.exchange_attach(r, ipc_handle=handles[r])
        dist.barrier()

        def fused(i):
            hf.step_sharded_device(states.data_ptr() + 96 * i, times.data_ptr() + 4 * i, 0, actions.data_ptr() + 16 * i, stream)

        df, wf = timed(fused)
        a_fused, a_fused_first = actions[W:W + K].clone(), actions[0].clone()
        ok = bool((hf.status() == 0).all())
        dist.barrier()
        hf.close()
        out.update({"device_ms_per_step": df, "wall_ms_per_step_incl_flush": wf, "steps_per_sec": 1e3 / df,
                    "strong_scaling_efficiency": d1 / (world * df), "speedup_vs_world1": d1 / df,
                    "exchange": "fused: %d B record stored into each of the %d ranks' buffers over NVLink (CUDA IPC peer memory) by the rollout "
                                "kernel + one flag per rank; merge kernel waits on the flags (covo_step_sharded_device)" % (4 * pn, world),
                    "nvlink_bytes_per_step_per_rank": (4 * pn + 4) * (world - 1),
                    "launches_per_step": 2, "exchange_status_ok": ok,
                    "first_step_action_diff_vs_world1": float((a_fused_first - a1_first).abs().max()),
                    "fused_equals_allgather_bitwise": bool(torch.equal(a_fused, a_nccl) and torch.equal(a_fused_first, a_nccl_first)),
                    "nccl_allgather_variant": {"device_ms_per_step": dw, "wall_ms_per_step_incl_flush": ww,
                                               "exchange": "kernel -> ncclAllGather of %d B per rank -> merge kernel (3 launches + a D2D copy)" % (4 * pn)},
                    "limiter": "the rollout grid is ceil(N/world/64) CTAs: at N=8192 it is ONE wave (128 CTAs on 148 SMs) on one GPU already, so "
                               "sharding cannot shorten the per-CTA chain (GEMM || 50-step rollout ~25 us) and only adds the exchange; it pays "
                               "when N/64 >> 148 (N=65536: 1024 CTAs = 7 waves on one GPU)"})
    h1.close()
    return out


def bench_env_batch(dist, world, rank, dev, K, E=512, N=1024):
    """BASELINE config 5: E environments per GPU behind one handle (CoVO-online, N = 1024, H = 50), the whole closed loop
    (noisy state -> controller -> Quad3D.step_env) on the device; ranks are independent (no collective)."""
    import torch

    from covo_mpc_b200 import _lib

    _, _, trajs, _ = synthetic_states(16 * 290, 1000 + rank)
    pos = np.stack([trajs[e % len(trajs)][0] for e in range(E)])
    vel = np.stack([trajs[e % len(trajs)][1] for e in range(E)])
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.n_env, cfg.seed, cfg.device = _lib.MODE_COVO_ONLINE, N, HORIZON, pos.shape[1], E, 3 + rank, dev.index
    cfg.lam = LAM
    h = _lib.Handle(cfg)
    h.set_reference(pos, vel)
    h.set_mean(_hover(E))
    s0 = np.zeros((E, 24), np.float32)
    s0[:, 6] = 1.0
    s0[:, 16:19] = pos[:, 0]
    s0[:, 19:22] = vel[:, 0]
    h.env_reset(s0, np.zeros(E, np.int32))
    h.closed_loop(3, noise_seed=1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    _, _, err = h.closed_loop(K, noise_seed=2)
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    h.set_profiling(True)
    h.closed_loop(1, noise_seed=3)
    km = h.kernel_ms()
    ok = bool((h.status() == 0).all())
    h.close()
    return {"metric": "env_mpc_steps_per_sec", "value": E * world * K / float(tt.item()), "unit": "env-steps/s", "envs_per_gpu": E, "n_samples": N,
            "horizon": HORIZON, "steps": K, "ms_per_batched_step": 1e3 * float(tt.item()) / K, "scaling": "weak",
            "timing": "wall clock around covo_closed_loop (K batched steps, one D2H of the logs at the end), max over ranks",
            "kernel_ms_one_step": {k: float(v) for k, v in zip(["hessian", "sigma_stage1", "sigma_stage2", "sigma_stage3", "cholesky", "rollout"], km)},
            "mean_err_pos": float(err.mean()), "status_ok": ok}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import covo_mpc_b200 as cm
    from covo_mpc_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    n_states = min(W + K, 290)
    seed = 100 + rank  # one independent environment per rank
    mode_name = args.controller
    states_h, times_h, trajs, _ = synthetic_states(n_states, seed)
    traj = trajs[0]
    env = cm.Quad3D(TASK)
    sidx = lambda i: i % n_states

    cfg = _lib.default_config()
    cfg.mode = {"covo-online": _lib.MODE_COVO_ONLINE, "covo-offline": _lib.MODE_COVO_OFFLINE, "mppi": _lib.MODE_MPPI}[mode_name]
    cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.device = N_SAMPLES, HORIZON, int(traj[0].shape[0]), local_rank
    cfg.lam, cfg.seed = LAM, seed
    h = _lib.Handle(cfg)
    h.set_reference(traj[0][None], traj[1][None])
    if cfg.mode == _lib.MODE_COVO_OFFLINE:
        h.reset_offline(states_h[0], [0], 300)
    states = torch.from_numpy(states_h).to(dev)
    times = torch.from_numpy(times_h).to(dev)
    actions = torch.zeros((n_states, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def one_step(i):
        i = sidx(i)
        h.step_device(states.data_ptr() + 96 * i, times.data_ptr() + 4 * i, 0, actions.data_ptr() + 16 * i, stream)

    for i in range(W):
        one_step(i)
    torch.cuda.synchronize()
    # per-kernel device time (instrumented pass, not the timed one)
    kernel_ms = None
    if cfg.mode != _lib.MODE_MPPI:
        h.set_profiling(True)
        acc = np.zeros(6)
        reps = min(10, K)
        for i in range(reps):
            flush.zero_()
            one_step(W + i)
            acc += h.kernel_ms()
        kernel_ms = acc / reps
        h.set_profiling(False)
        # restart the controller state so the timed pass sees the same sequence again
        h.set_mean(_hover())
        for i in range(W):
            one_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = _events(K)
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()  # L2 flush, outside the event pair
        ev[i][0].record()
        one_step(W + i)
        ev[i][1].record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    # nvidia-smi needs a few hundred ms to deliver its first line (longer on an 8-GPU box) and K steps can be over before that: keep the
    # same load running, untimed, until two clock samples have been taken under it
    t_lim, j = time.perf_counter() + 4.0, 0
    while sampler.n_samples() < 2 and time.perf_counter() < t_lim:
        for _ in range(50):
            flush.zero_()
            one_step(W + (j % K))
            j += 1
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(step_ms.sum())
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = K * world / (total_ms_max / 1e3)
    assert torch.isfinite(actions).all(), "non-finite actions"
    status_ok = bool((h.status() == 0).all()) if cfg.mode == _lib.MODE_COVO_ONLINE else True
    sigma_path = h.sigma_path()
    slot_names = h.kernel_slot_names()

    # ---- the other optimize_sigma path (same states, same timing protocol; rank-local, not the headline) --------
    PATH_NAMES = {0: "tridiagonal (E1-E3: cluster Householder + float64 Sturm/Zolotarev on the tridiagonal + sandwich)",
                  3: "dense (D1-D3: adaptive float64 cluster Lanczos + 13 float64 cluster Gauss-Jordan pole inverses + combine)"}
    fast_sigma = None
    if cfg.mode == _lib.MODE_COVO_ONLINE:
        alt_path = 0 if sigma_path == 3 else 3
        h.set_sigma_path(alt_path)
        h.set_mean(_hover())
        for i in range(W):
            one_step(i)
        torch.cuda.synchronize()
        ev2 = _events(K)
        for i in range(K):
            flush.zero_()
            ev2[i][0].record()
            one_step(W + i)
            ev2[i][1].record()
        torch.cuda.synchronize()
        ms2 = np.array([a.elapsed_time(b) for a, b in ev2])
        ok2 = bool((h.status() == 0).all())
        h.set_profiling(True)
        acc2 = np.zeros(6)
        for i in range(min(10, K)):
            flush.zero_()
            one_step(W + i)
            acc2 += h.kernel_ms()
        names2 = h.kernel_slot_names()
        h.set_profiling(False)
        fast_sigma = {"value": K / (float(ms2.sum()) / 1e3), "unit": UNIT, "ms_per_step": float(ms2.mean()), "step_ms_p99": float(np.percentile(ms2, 99)),
                      "numeric_status_ok": ok2, "kernel_us": {k: round(float(v) / min(10, K) * 1e3, 1) for k, v in zip(names2, acc2)},
                      "sigma_path": PATH_NAMES[alt_path],
                      "note": "the optimize_sigma path that is NOT the default of this handle (covo_set_sigma_path / COVO_SIGMA=tridiag|dense); "
                              "rank-local number, not the headline"}
        h.set_sigma_path(sigma_path)
        h.set_mean(_hover())
        for i in range(W):
            one_step(i)
        torch.cuda.synchronize()

    # ---- e2e through the plugin surface with host buffers -------------------------------------------
    ctl, cp = cm.get_controller(env, mode_name, f"N{N_SAMPLES}_H{HORIZON}_lam{LAM}", device=local_rank, seed=seed)
    f32 = np.float32
    st0 = cm.EnvState3D(pos=np.zeros(3, f32), vel=np.zeros(3, f32), quat=np.array([0, 0, 0, 1], f32), omega=np.zeros(3, f32),
                        pos_traj=traj[0], vel_traj=traj[1], acc_traj=np.zeros_like(traj[0]), pos_tar=np.zeros(3, f32),
                        vel_tar=np.zeros(3, f32), acc_tar=np.zeros(3, f32), time=0, f_disturb=np.zeros(3, f32))

    def mk(i):
        s = states_h[i]
        return st0.replace(pos=s[0:3], quat=s[3:7], vel=s[7:10], omega=s[10:13], f_disturb=s[13:16], pos_tar=s[16:19],
                           vel_tar=s[19:22], time=int(times_h[i]))

    host_states = [mk(i) for i in range(n_states)]
    if mode_name == "covo-offline":
        cp = ctl.reset(host_states[0], env.default_params, cp, None)
    for i in range(W):
        _, cp, _ = ctl(None, host_states[i], env.default_params, None, cp, {"noisy_state": host_states[i]})
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        hs_i = host_states[sidx(W + i)]
        act, cp, _ = ctl(None, hs_i, env.default_params, None, cp, {"noisy_state": hs_i})
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    tt = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e = {"value": K * world / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": 24 * 4 + 4, "d2h_bytes_per_step": 16 + 4,
           "ms_per_step": 1e3 * float(tt.item()) / K,
           "path": "get_controller(...)(obs, state, env_params, rng, control_params, info) with numpy state -> covo_step (pinned staging) -> action"}
    ctl.close()

    # ---- device-resident closed loop: controller + environment step, no host round trip -----
    h.set_mean(_hover())
    h.env_reset(states_h[0][None], times_h[:1])
    h.closed_loop(W, noise_seed=seed)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, _, err_cl = h.closed_loop(K, noise_seed=seed + 1)
    tc = time.perf_counter() - t0
    closed = {"value": K * world / tc, "unit": UNIT, "ms_per_step": 1e3 * tc / K, "mean_err_pos": float(err_cl.mean()),
              "note": "covo_closed_loop: K x [noisy state -> controller -> Quad3D.step_env] on the device, one D2H of the logs at the end (wall clock)"}
    # kernels inside the replayed CUDA graph of one step (+ the counter kernel that replaces the per-step launch arguments)
    launches_per_step = {"covo-online": (8 if sigma_path == 3 else 9) + 1, "covo-offline": 1 + 1, "mppi": 1 + 1}[mode_name]
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{mode_name} {TASK} N={N_SAMPLES} H={HORIZON} u_dim=4 lam={LAM} sigma=0.5, 1 env per GPU",
                   "controller": mode_name, "n_samples": N_SAMPLES, "horizon": HORIZON, "envs_per_gpu": 1,
                   "parallelism": "env-replicas x%d (no collective)" % world,
                   "inputs": "noisy states of a PID closed loop on a zigzag reference (host-generated, shared with --impl reference)",
                   "rng": "in-kernel Philox (production mode)", "l2": "256 MiB memset between steps, excluded from the event timing",
                   "timing": "CUDA events per step on the launch stream, sum over K steps, max over ranks",
                   "sigma_path": PATH_NAMES.get(sigma_path, str(sigma_path)) if cfg.mode == _lib.MODE_COVO_ONLINE else None},
        "wall_ms_per_step_incl_flush": 1e3 * t_wall / K,
        "step_ms_p50": float(np.median(step_ms)), "step_ms_p99": float(np.percentile(step_ms, 99)),
        "gpu_launches": launches_per_step * K, "clocks": clocks, "numeric_status_ok": status_ok,
        "e2e": e2e, "closed_loop": closed,
    }
    if fast_sigma is not None:
        out["alt_sigma"] = fast_sigma
    # ---- roofline ---------------------------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    n = 4 * HORIZON
    ab = algorithmic_bytes(n, HORIZON, N_SAMPLES)
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    if kernel_ms is not None and cfg.mode == _lib.MODE_COVO_ONLINE:
        per = {k: float(v) for k, v in zip(slot_names, kernel_ms) if v > 1e-4}
        dom = max(per, key=per.get)
        rl = {}
        for k, ms in per.items():
            b = ab.get(k, ab.get("sigma"))
            ach = b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            rl[k] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic.get(k),
                     "ms": float(ms), "algorithmic_bytes": b, "share_of_step": float(ms / sum(per.values()))}
        out["roofline"] = dict(rl[dom], kernel=dom, peak_source=peak_src,
                               note="latency-bound by construction (SURVEY 8d): the whole step moves ~1.7 MB through HBM; the dominant "
                                    "kernel is a chain of dependent small-matrix steps, see DESIGN.md section 4")
        out["roofline_kernels"] = rl
        flops = N_SAMPLES * n * (n + 1) + N_SAMPLES * HORIZON * 200 + 2 * N_SAMPLES * n
        out["rollout_fp32"] = {"algorithmic_gflop": flops / 1e9, "achieved_tflops": flops / (per["rollout"] * 1e-3) / 1e12,
                               "nominal_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12}
    else:
        ms = total_ms_max / K
        ach = ab["rollout"] / (ms * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                           "traffic": traffic.get("rollout"), "kernel": "rollout", "peak_source": peak_src}
    # ---- the BASELINE configurations the replica line does not show --------------------------------------------
    if not args.no_subrecords:
        h.close()
        Ks = min(K, 100)
        if world > 1:  # config 4 is ONE environment: every rank holds rank 0's states and reference (the replica run above has one per rank)
            s0_h, t0_h, trajs0, _ = synthetic_states(n_states, 100)
            states, times, traj = torch.from_numpy(s0_h).to(dev), torch.from_numpy(t0_h).to(dev), trajs0[0]
        out["nsample_shard"] = {
            "workload": f"covo-offline {TASK} H={HORIZON}, the N samples of ONE environment split over {world} rank(s) (BASELINE config 4)",
            "N8192": bench_nsample_shard(dist, world, rank, dev, states, times, traj, Ks, W, flush, 8192),
            "N65536": bench_nsample_shard(dist, world, rank, dev, states, times, traj, Ks, W, flush, 65536)}
        out["env_batch"] = bench_env_batch(dist, world, rank, dev, min(K, 20))
        out["env_batch"]["workload"] = f"{512 * world} envs x covo-online N=1024 H={HORIZON}, env-batch sharded over {world} GPU(s), no collective (BASELINE config 5)"
    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(states_h, times_h, traj, budget_s=args.cpu_budget, mode_name=mode_name)
        if mode_name == "covo-online":
            out["tracking_cost"] = tracking_cost(local_rank)
    if rank == 0:
        print(json.dumps(out), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def cpu_baseline(states_h, times_h, traj, budget_s=20.0, mode_name="covo-online", max_steps=None, warmup=0):
    """The reference algorithm (oracle/ port) on the host CPU: full MPC steps on the same recorded states."""
    from oracle import oracle_np as o

    try:
        from oracle import oracle_c

        fast = oracle_c.available()
    except Exception:
        fast = False
    p = o.EnvParams()
    a_mean = o.hover_mean(HORIZON, p)
    rng = np.random.default_rng(0)
    n = 4 * HORIZON
    t_used, done = 0.0, 0
    i = 0
    cores = oracle_c.num_threads() if fast else 1  # OpenMP threads the port actually runs on
    a_cov = None
    while True:
        s = states_h[i % len(states_h)]
        ns = o.make_state(s[0:3], s[3:7], s[7:10], s[10:13], s[13:16], int(times_h[i % len(times_h)]), traj[0], traj[1], s[16:19],
                          s[19:22], dtype=np.float32)
        eps = rng.standard_normal((N_SAMPLES, n)).astype(np.float32)
        t0 = time.perf_counter()
        if fast:
            _, a_mean = oracle_c.covo_step(ns, a_mean, eps, p, LAM, online=(mode_name == "covo-online"), a_cov=a_cov)
        else:
            if mode_name == "covo-online" or a_cov is None:
                _, a_mean, a_cov, _ = o.covo_call(ns, a_mean, eps, p, lam=LAM, hessian_dtype=np.float32)
            else:
                _, a_mean, a_cov, _ = o.covo_call(ns, a_mean, eps, p, lam=LAM, a_cov=a_cov)
        dt = time.perf_counter() - t0
        i += 1
        if i <= warmup:
            continue
        t_used += dt
        done += 1
        if (max_steps is not None and done >= max_steps) or (max_steps is None and t_used >= budget_s) or t_used > 170.0:
            break
    return {"value": done / t_used, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} full MPC steps ({mode_name}, N={N_SAMPLES}, H={HORIZON}) of the oracle restatement "
                      f"({'C/OpenMP + LAPACK' if fast else 'NumPy float32 + LAPACK eigh'}); JAX is not installable here",
            "steps": done, "seconds": t_used}


def tracking_cost(device=0, controller="covo-online"):
    """BASELINE metric, second half ("tracking cost delta vs ref"), POWERED: the reference's protocol (envs/quadrotor.py:564-579:
    4 trajectories x 10 episodes x 300 steps, mean over episodes of the per-episode mean ||pos_tar - pos||) on both sides with the
    SAME trajectories, initial states and observation-noise streams per episode (tools/tracking_protocol.py).  Oracle arm: the
    fixture tests/golden/tracking/oracle_tracking_*.npz written by tools/oracle_tracking_stats.py (CPU, ~35 min; committed with the script).
    Device arm: all 40 episodes as one batched device-resident closed loop with the production Philox sample field.  The sample
    streams differ (numpy vs Philox), so this is a statistical comparison: means +- s.e., z-scores unpaired and paired by episode."""
    from tools import device_tracking_stats as dts

    fx_path = dts.fixture_path(controller, N_SAMPLES, HORIZON)
    if not os.path.exists(fx_path):
        return {"unavailable": f"{os.path.relpath(fx_path, ROOT)} missing: run tools/oracle_tracking_stats.py"}
    fx = np.load(fx_path)
    ora = fx["err_pos"]
    dev_err, status = dts.production_arm(controller, N_SAMPLES, HORIZON, float(fx["lam"]), ora.shape[0], ora.shape[1], device=device)
    out = dts.summarise(dev_err, ora)
    out.update({"metric": "mean ||pos_tar - pos|| over 300-step episodes (m), mean over episodes", "protocol": "tools/tracking_protocol.py",
                "oracle_fixture": os.path.relpath(fx_path, ROOT), "device_numeric_status_nonzero": int(np.count_nonzero(status)),
                "note": "same trajectories / initial states / observation noise per episode, independent sample streams; identical-eps "
                        "parity is tests/test_tracking_gpu.py"})
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    world = max(1, args.gpus)
    # all host threads this process may use: torchrun exports OMP_NUM_THREADS=1, which would silently serialise the
    # OpenMP loops of the port (libgomp reads it when the library is first loaded, i.e. below) and LAPACK/BLAS
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(ncores)
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=ncores)
    except Exception:
        pass
    # The GPU arm at --gpus G steps G independent environments (one per GPU); the host steps the SAME G environments one after the
    # other (same generator, seeds 100 .. 100 + G - 1), so value = environment-steps / s is a ratio of equal work at every G.
    per_env = []
    for r in range(world):
        st, tm, trajs, _ = synthetic_states(min(W + K, 290), 100 + r)
        per_env.append((st, tm, trajs[0]))
    steps_each = max(1, K // world)
    t_used, done, cb = 0.0, 0, None
    for r in range(world):
        st, tm, traj = per_env[r]
        cb = cpu_baseline(st, tm, traj, mode_name=args.controller, max_steps=steps_each, warmup=W if r == 0 else 0)
        t_used += cb["seconds"]
        done += cb["steps"]
    value = done / t_used
    cb.update({"value": value, "steps": done, "seconds": t_used,
               "sample": f"{done} full MPC steps ({args.controller}, N={N_SAMPLES}, H={HORIZON}) over {world} environment(s) stepped one after "
                         f"the other on the host, {cb['sample'].split(') of ', 1)[-1]}"})
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
           "warmup": W, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.controller} {TASK} N={N_SAMPLES} H={HORIZON} u_dim=4 lam={LAM} sigma=0.5, 1 env per GPU",
                      "controller": args.controller, "n_samples": N_SAMPLES, "horizon": HORIZON, "envs_per_gpu": 1,
                      "parallelism": "env-replicas x%d (no collective)" % world,
                      "inputs": "noisy states of a PID closed loop on a zigzag reference (host-generated, shared with --impl reference)",
                      "equal_work": f"the {world} environment(s) of the GPU arm, stepped sequentially on all host cores: steps/s here is "
                                    "environment-steps per second, the same unit as the GPU arm's whole-job value",
                      "note": "CPU restatement of the reference algorithm (JAX cannot be installed offline)"},
           "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), file=_REAL_STDOUT, flush=True)


def _claim_stdout():
    """The contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner does) are sent to stderr; the line
    itself goes to the real stdout through the returned file object."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--controller", default="covo-online", choices=["covo-online", "covo-offline", "mppi"])
    ap.add_argument("--no-subrecords", action="store_true", help="skip the nsample_shard / env_batch records (configs 4 and 5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 60:
            args.steps = 60  # bounded sample: each step is one full N=8192, H=50 MPC step on the CPU
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
