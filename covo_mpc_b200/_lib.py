"""ctypes binding of libcovo_b200.so (include/covo_b200.h).

The library is the product; there is no CPU implementation behind it.  Loading fails loudly if the
shared object is missing, and ``Handle`` creation fails loudly if no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcovo_b200.so")

MODE_MPPI, MODE_COVO_ONLINE, MODE_COVO_OFFLINE = 0, 1, 2
ERR_INVALID, ERR_NOT_IMPLEMENTED, ERR_CUDA, ERR_NUMERIC = 1, 2, 3, 4


class CovoConfig(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("n_samples", C.c_int), ("horizon", C.c_int), ("n_env", C.c_int),
        ("traj_len", C.c_int), ("device", C.c_int), ("rank", C.c_int), ("world", C.c_int),
        ("lam", C.c_float), ("sample_sigma", C.c_float), ("gamma_mean", C.c_float),
        ("gamma_sigma", C.c_float), ("discount", C.c_float),
        ("m", C.c_float), ("g", C.c_float), ("max_thrust", C.c_float), ("dt", C.c_float),
        ("alpha_bodyrate", C.c_float), ("action_scale", C.c_float), ("pos_limit", C.c_float),
        ("max_omega", C.c_float * 3), ("max_steps_in_episode", C.c_int),
        ("seed", C.c_ulonglong),
    ]


class CovoCudaError(RuntimeError):
    pass


class CovoNumericError(ArithmeticError):
    pass


_lib: Optional[C.CDLL] = None

_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int)
_D = C.POINTER(C.c_double)
_H = C.c_void_p

_SIGNATURES = {
    "covo_default_config": [C.POINTER(CovoConfig)],
    "covo_create": [C.POINTER(CovoConfig), C.POINTER(_H)],
    "covo_destroy": [_H],
    "covo_set_reference": [_H, _F, _F, _F],
    "covo_set_mean": [_H, _F],
    "covo_get_mean": [_H, _F],
    "covo_set_cov": [_H, _F],
    "covo_get_cov": [_H, _F],
    "covo_set_cov_offline": [_H, _F, C.c_int],
    "covo_get_cov_offline": [_H, _F, C.c_int],
    "covo_reset_offline": [_H, _F, _I, C.c_int],
    "covo_reset_offline_disturbed": [_H, _F, _I, C.c_int, _F],
    "covo_step": [_H, _F, _I, _F, _F],
    "covo_set_rollout_disturbance": [_H, _F],
    "covo_set_env_params": [_H, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _F, C.c_int],
    "covo_step_device": [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "covo_step_partial_device": [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "covo_partial_buffer": [_H, C.POINTER(C.c_void_p), _I],
    "covo_step_merge_device": [_H, C.c_void_p, C.c_void_p, C.c_void_p],
    "covo_exchange_info": [_H, C.c_void_p, C.POINTER(C.c_void_p)],
    "covo_exchange_attach": [_H, C.c_int, C.c_void_p, C.c_void_p],
    "covo_step_sharded_device": [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "covo_hessian": [_H, _F, _I, _F, C.c_int, _F],
    "covo_optimize_sigma": [_H, _F, _F],
    "covo_cholesky": [_H, _F, _F],
    "covo_rollout": [_H, _F, _I, _F, C.c_int, _F, _F, _F, _F, _F, _F],
    "covo_get_pos_stats": [_H, _F, _F],
    "covo_enable_pos_stats": [_H, C.c_int],
    "covo_debug_eps": [_H, C.c_uint, _F],
    "covo_set_jax_key": [_H, C.POINTER(C.c_uint)],
    "covo_pid_action": [_H, _F, _I, C.c_float, C.c_float, C.c_float, C.c_float, _F, _F],
    "covo_debug_tridiag": [_H, _D, _D, _D],
    "covo_zolotarev_nodes": [C.c_double, C.c_double, C.c_int, _D, _D],
    "covo_get_status": [_H, _I],
    "covo_set_profiling": [_H, C.c_int],
    "covo_get_kernel_ms": [_H, _F],
    "covo_debug_phase_clocks": [_H, C.c_int, C.POINTER(C.c_longlong)],
    "covo_rng_step": [_H, C.POINTER(C.c_uint)],
    "covo_get_sigma_path": [_H, _I],
    "covo_set_sigma_path": [_H, C.c_int],
    "covo_env_reset": [_H, _F, _I],
    "covo_env_get_state": [_H, _F, _I],
    "covo_env_step": [_H, _F, _F, C.c_ulonglong, C.c_uint, C.c_int, C.c_float, C.c_float, _F, _F, _F, _I],
    "covo_env_set_reset_pool": [_H, C.c_int, _F, _I, _F, _F, _F],
    "covo_closed_loop": [_H, C.c_int, C.c_ulonglong, C.c_int, C.c_float, C.c_float, _F, _F, _F, _F],
    "covo_local_samples": [_H, _I, _I],
}
EXPORTED = sorted(list(_SIGNATURES) + ["covo_last_error", "covo_version"])


def load() -> C.CDLL:
    """Load the shared library (built in-tree by covo_mpc_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m covo_mpc_b200.build` (nvcc, sm_100a). "
            "covo_mpc_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.covo_last_error.restype = C.c_char_p
    lib.covo_last_error.argtypes = []
    lib.covo_version.restype = C.c_char_p
    lib.covo_version.argtypes = []
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Map C status codes to the exception types the reference raises for the same conditions."""
    if rc == 0:
        return
    msg = load().covo_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_NOT_IMPLEMENTED:
        raise NotImplementedError(msg)  # envs/quadrotor.py:751, controllers/covo.py:114
    if rc == ERR_NUMERIC:
        raise CovoNumericError(msg)
    raise CovoCudaError(msg)


def f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def fptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_F)


def iptr(a: np.ndarray):
    return a.ctypes.data_as(_I)


def default_config() -> CovoConfig:
    cfg = CovoConfig()
    check(load().covo_default_config(C.byref(cfg)))
    return cfg


class Handle:
    """Owner of one ``covo_handle`` (one device, one stream)."""

    def __init__(self, cfg: CovoConfig):
        self.lib = load()
        self.cfg = cfg
        self._h = _H()
        check(self.lib.covo_create(C.byref(cfg), C.byref(self._h)))
        self.H = cfg.horizon
        self.n = 4 * cfg.horizon
        self.E = cfg.n_env
        nl, off = C.c_int(), C.c_int()
        check(self.lib.covo_local_samples(self._h, C.byref(nl), C.byref(off)))
        self.n_local, self.sample_offset = nl.value, off.value

    def close(self):
        if self._h:
            self.lib.covo_destroy(self._h)
            self._h = _H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- control params -----------------------------------------------------------------------
    def set_reference(self, pos, vel, acc=None):
        pos, vel = f32(pos), f32(vel)
        acc = None if acc is None else f32(acc)
        exp = (self.E * self.cfg.traj_len * 3,)
        if pos.size != exp[0] or vel.size != exp[0]:
            raise ValueError(f"reference trajectory must have {self.E}x{self.cfg.traj_len}x3 entries")
        check(self.lib.covo_set_reference(self._h, fptr(pos), fptr(vel), fptr(acc)))

    def set_mean(self, a_mean):
        a = f32(a_mean)
        if a.size != self.E * self.n:
            raise ValueError("a_mean has the wrong size")
        check(self.lib.covo_set_mean(self._h, fptr(a)))

    def get_mean(self) -> np.ndarray:
        out = np.empty((self.E, self.H, 4), dtype=np.float32)
        check(self.lib.covo_get_mean(self._h, fptr(out)))
        return out

    def cov_shape(self):
        return (self.E, self.H, 4, 4) if self.cfg.mode == MODE_MPPI else (self.E, self.n, self.n)

    def set_cov(self, a_cov):
        a = f32(a_cov)
        if a.size != int(np.prod(self.cov_shape())):
            raise ValueError("a_cov has the wrong size")
        check(self.lib.covo_set_cov(self._h, fptr(a)))

    def get_cov(self) -> np.ndarray:
        out = np.empty(self.cov_shape(), dtype=np.float32)
        check(self.lib.covo_get_cov(self._h, fptr(out)))
        return out

    def set_cov_offline(self, table):
        t = f32(table)
        check(self.lib.covo_set_cov_offline(self._h, fptr(t), int(t.shape[0])))

    def get_cov_offline(self, t_sched: int) -> np.ndarray:
        out = np.empty((t_sched, self.n, self.n), dtype=np.float32)
        check(self.lib.covo_get_cov_offline(self._h, fptr(out), t_sched))
        return out

    def pid_action(self, state24, time, Kp=10.0, Kd=5.0, Ki=0.0, Kp_att=10.0, integral=None) -> np.ndarray:
        """PIDController.__call__ on the device for the handle's environments (covo_pid_action)."""
        s, t = f32(state24), i32(time)
        if s.size != self.E * 24 or t.size != self.E:
            raise ValueError("state24 / time have the wrong size")
        g = None if integral is None else f32(integral)
        if g is not None and g.size != self.E * 3:
            raise ValueError("integral must be [E][3]")
        out = np.empty((self.E, 4), dtype=np.float32)
        check(self.lib.covo_pid_action(self._h, fptr(s), iptr(t), Kp, Kd, Ki, Kp_att, fptr(g), fptr(out)))
        return out

    def set_rollout_disturbance(self, fdist_seq):
        """MPPI, disturb_type gaussian (mppi.py:74): fdist_seq [E][H][3] = the force the rollouts see after step h; None: off."""
        if fdist_seq is None:
            check(self.lib.covo_set_rollout_disturbance(self._h, None))
            return
        f = f32(fdist_seq)
        if f.size != self.E * self.H * 3:
            raise ValueError("fdist_seq must be [E][H][3]")
        check(self.lib.covo_set_rollout_disturbance(self._h, fptr(f)))

    def reset_offline(self, state24, time, t_sched: int, f_disturb=None):
        """f_disturb [t_sched][3]: the gaussian disturbance force after every path step (covo_reset_offline_disturbed); None: none."""
        s, t = f32(state24), i32(time)
        if f_disturb is None:
            check(self.lib.covo_reset_offline(self._h, fptr(s), iptr(t), t_sched))
            return
        d = f32(f_disturb)
        if d.shape != (t_sched, 3):
            raise ValueError("f_disturb must be [t_sched][3]")
        check(self.lib.covo_reset_offline_disturbed(self._h, fptr(s), iptr(t), t_sched, fptr(d)))

    # -- the MPC step -----------------------------------------------------------------------------
    def set_jax_key(self, act_key):
        """The next sampling call draws what the reference would draw from ``act_key`` (covo_set_jax_key)."""
        k = np.ascontiguousarray(act_key, dtype=np.uint32).ravel()
        if k.size != 2:
            raise ValueError("a JAX key is two uint32 words")
        check(self.lib.covo_set_jax_key(self._h, k.ctypes.data_as(C.POINTER(C.c_uint))))

    def _step_buffers(self):
        """Staging arrays of the per-step call and their ctypes pointers, made once: building numpy arrays and pointer objects
        per call cost more than the H2D copy they feed (17 us -> 5 us per call, measured with a no-op library)."""
        sb = getattr(self, "_step_bufs", None)
        if sb is None:
            s = np.empty(self.E * 24, dtype=np.float32)
            t = np.empty(self.E, dtype=np.int32)
            out = np.empty((self.E, 4), dtype=np.float32)
            sb = self._step_bufs = (s, t, out, fptr(s), iptr(t), fptr(out))
        return sb

    def step(self, state24, time, eps=None) -> np.ndarray:
        if eps is None:  # production path: no per-call allocations besides the returned action
            s, t, out, sp, tp, op = self._step_buffers()
            a = np.asarray(state24)
            if a.size != s.size or np.size(time) != t.size:
                raise ValueError("state24 / time have the wrong size")
            s[:] = a.reshape(-1)
            t[:] = time
            check(self.lib.covo_step(self._h, sp, tp, None, op))
            return out.copy()
        s, t = f32(state24), i32(time)
        if s.size != self.E * 24 or t.size != self.E:
            raise ValueError("state24 / time have the wrong size")
        e = f32(eps)
        if e.size != self.E * self.n_local * self.n:
            raise ValueError("eps must be [E][N_local][4H]")
        out = np.empty((self.E, 4), dtype=np.float32)
        check(self.lib.covo_step(self._h, fptr(s), iptr(t), fptr(e), fptr(out)))
        return out

    def step_state(self, env_state) -> np.ndarray:
        """covo_step for ONE environment straight from an EnvState3D-like object (``pack_into``, ``time``): the plugin call's
        path, without the intermediate 24-float array."""
        if self.E != 1:
            raise ValueError("step_state is the single-environment path")
        s, t, out, sp, tp, op = self._step_buffers()
        env_state.pack_into(s)
        t[0] = env_state.time
        check(self.lib.covo_step(self._h, sp, tp, None, op))
        return out[0].copy()

    def set_env_params(self, m, g, max_thrust, dt, alpha_bodyrate, action_scale, max_omega, max_steps_in_episode):
        mo = f32(max_omega)
        check(self.lib.covo_set_env_params(self._h, m, g, max_thrust, dt, alpha_bodyrate, action_scale, fptr(mo), int(max_steps_in_episode)))

    def step_device(self, state24_ptr: int, time_ptr: int, eps_ptr: int, action_ptr: int, stream: int = 0):
        check(self.lib.covo_step_device(self._h, state24_ptr, time_ptr, eps_ptr or None, action_ptr, stream or None))

    def step_partial_device(self, state24_ptr: int, time_ptr: int, eps_ptr: int, stream: int = 0):
        check(self.lib.covo_step_partial_device(self._h, state24_ptr, time_ptr, eps_ptr or None, stream or None))

    def partial_buffer(self):
        p, n = C.c_void_p(), C.c_int()
        check(self.lib.covo_partial_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def step_merge_device(self, gathered_ptr: int, action_ptr: int, stream: int = 0):
        check(self.lib.covo_step_merge_device(self._h, gathered_ptr, action_ptr, stream or None))

    # -- fused exchange of the N-sharded step (no collective call: records go peer to peer from inside the rollout kernel) ------
    def exchange_info(self):
        """(64-byte CUDA IPC handle, raw device pointer) of this rank's exchange buffer."""
        hd = C.create_string_buffer(64)
        p = C.c_void_p()
        check(self.lib.covo_exchange_info(self._h, hd, C.byref(p)))
        return hd.raw, p.value

    def exchange_attach(self, peer_rank: int, ipc_handle: Optional[bytes] = None, dev_ptr: Optional[int] = None):
        """Make rank `peer_rank`'s exchange buffer known: by IPC handle (another process) or by pointer (a handle of this process)."""
        buf = C.create_string_buffer(ipc_handle, 64) if ipc_handle is not None else None
        check(self.lib.covo_exchange_attach(self._h, int(peer_rank), buf, dev_ptr))

    def step_sharded_device(self, state_ptr: int, time_ptr: int, eps_ptr: int, action_ptr: int, stream: int = 0):
        check(self.lib.covo_step_sharded_device(self._h, state_ptr, time_ptr, eps_ptr or None, action_ptr, stream or None))

    # -- operators -------------------------------------------------------------------------------
    def hessian(self, state24, time, a_mean, shift: bool = False) -> np.ndarray:
        s, t, a = f32(state24), i32(time), f32(a_mean)
        out = np.empty((self.E, self.n, self.n), dtype=np.float32)
        check(self.lib.covo_hessian(self._h, fptr(s), iptr(t), fptr(a), int(shift), fptr(out)))
        return out

    def optimize_sigma(self, R) -> np.ndarray:
        r = f32(R)
        if r.size != self.E * self.n * self.n:
            raise ValueError("R has the wrong size")
        out = np.empty((self.E, self.n, self.n), dtype=np.float32)
        check(self.lib.covo_optimize_sigma(self._h, fptr(r), fptr(out)))
        return out

    def cholesky(self, a_cov) -> np.ndarray:
        c = f32(a_cov)
        out = np.empty((self.E, self.n, self.n), dtype=np.float32)
        check(self.lib.covo_cholesky(self._h, fptr(c), fptr(out)))
        return out

    def rollout(self, state24, time, a_mean, shift=False, eps=None, fdist_seq=None, want_costs=False, want_samples=False):
        s, t, a = f32(state24), i32(time), f32(a_mean)
        e = None if eps is None else f32(eps)
        fd = None if fdist_seq is None else f32(fdist_seq)
        a_out = np.empty((self.E, self.H, 4), dtype=np.float32)
        act = np.empty((self.E, 4), dtype=np.float32)
        costs = np.empty((self.E, self.n_local), dtype=np.float32) if want_costs else None
        samples = np.empty((self.E, self.n_local, self.H, 4), dtype=np.float32) if want_samples else None
        check(self.lib.covo_rollout(self._h, fptr(s), iptr(t), fptr(a), int(shift), fptr(e), fptr(fd), fptr(a_out),
                                    fptr(act), fptr(costs), fptr(samples)))
        return a_out, act, costs, samples

    # -- device-resident environment / closed loop ----------------------------------
    def env_reset(self, state24, time):
        s, t = f32(state24), i32(time)
        if s.size != self.E * 24 or t.size != self.E:
            raise ValueError("state24 / time have the wrong size")
        check(self.lib.covo_env_reset(self._h, fptr(s), iptr(t)))

    def env_set_reset_pool(self, state24, time, pos_traj, vel_traj, a_mean_init=None):
        """Auto-reset on the device (envs/base.py:27-38): state24 [P][E][24], time [P][E], pos_traj / vel_traj [P][E][T][3] are the
        reset_env draws each environment continues from when its pre-step state is terminal; a_mean_init [H][4]: also reset the
        controller's mean.  state24 None switches it off."""
        if state24 is None:
            check(self.lib.covo_env_set_reset_pool(self._h, 0, None, None, None, None, None))
            return
        s, t, pt, vt = f32(state24), i32(time), f32(pos_traj), f32(vel_traj)
        P = s.size // (self.E * 24)
        if s.size != P * self.E * 24 or t.size != P * self.E or pt.size != vt.size or pt.size % (P * self.E * 3):
            raise ValueError("reset pool arrays have inconsistent sizes")
        m = None if a_mean_init is None else f32(a_mean_init)
        if m is not None and m.size != self.n:
            raise ValueError("a_mean_init must be [H][4]")
        check(self.lib.covo_env_set_reset_pool(self._h, P, fptr(s), iptr(t), fptr(pt), fptr(vt), fptr(m)))

    def env_state(self):
        s = np.empty((self.E, 24), dtype=np.float32)
        t = np.empty(self.E, dtype=np.int32)
        check(self.lib.covo_env_get_state(self._h, fptr(s), iptr(t)))
        return s, t

    def env_step(self, action, noise=None, noise_seed=0, noise_step=0, gaussian=False, obs_noise_scale=0.05, dyn_noise_scale=0.05):
        """One Quad3D.step_env + get_info transition on the device (action None: only the noisy copy of the state).
        Returns (noisy24 [E][24], reward [E], err_pos [E], done [E]); the last three describe the PRE-step state."""
        a = None if action is None else f32(action)
        if a is not None and a.size != self.E * 4:
            raise ValueError("action must be [E][4]")
        z = None if noise is None else f32(noise)
        if z is not None and z.size != self.E * 16:
            raise ValueError("noise must be [E][16]")
        noisy = np.empty((self.E, 24), dtype=np.float32)
        rew = np.zeros(self.E, dtype=np.float32)
        err = np.zeros(self.E, dtype=np.float32)
        done = np.zeros(self.E, dtype=np.int32)
        check(self.lib.covo_env_step(self._h, fptr(a), fptr(z), int(noise_seed), int(noise_step), int(gaussian), float(obs_noise_scale),
                                     float(dyn_noise_scale), fptr(noisy), fptr(rew), fptr(err), iptr(done)))
        return noisy, rew, err, done

    def closed_loop(self, n_steps, noise=None, noise_seed=0, gaussian=False, obs_noise_scale=0.05, dyn_noise_scale=0.05):
        """n_steps x [controller (production RNG) -> env step] without a host round trip.
        Returns (actions [n][E][4], rewards [n][E], err_pos [n][E])."""
        z = None if noise is None else f32(noise)
        if z is not None and z.size != (n_steps + 1) * self.E * 16:
            raise ValueError("noise must be [n_steps+1][E][16]")
        act = np.empty((n_steps, self.E, 4), dtype=np.float32)
        rew = np.empty((n_steps, self.E), dtype=np.float32)
        err = np.empty((n_steps, self.E), dtype=np.float32)
        check(self.lib.covo_closed_loop(self._h, int(n_steps), int(noise_seed), int(gaussian), float(obs_noise_scale), float(dyn_noise_scale),
                                        fptr(z), fptr(act), fptr(rew), fptr(err)))
        return act, rew, err

    # -- introspection ------------------------------------------------------------------------------
    def enable_pos_stats(self, on=True):
        check(self.lib.covo_enable_pos_stats(self._h, int(on)))

    def pos_stats(self):
        m = np.empty((self.E, self.H, 3), dtype=np.float32)
        s = np.empty((self.E, self.H, 3), dtype=np.float32)
        check(self.lib.covo_get_pos_stats(self._h, fptr(m), fptr(s)))
        return m, s

    def debug_eps(self, stream_id: int) -> np.ndarray:
        out = np.empty((self.n_local, self.n), dtype=np.float32)
        check(self.lib.covo_debug_eps(self._h, stream_id, fptr(out)))
        return out

    def debug_tridiag(self):
        d = np.empty(self.n, dtype=np.float64)
        e = np.empty(self.n, dtype=np.float64)
        sc = np.empty(5, dtype=np.float64)
        check(self.lib.covo_debug_tridiag(self._h, d.ctypes.data_as(_D), e.ctypes.data_as(_D), sc.ctypes.data_as(_D)))
        return d, e, sc

    def status(self) -> np.ndarray:
        out = np.empty(self.E, dtype=np.int32)
        check(self.lib.covo_get_status(self._h, iptr(out)))
        return out

    def set_profiling(self, on=True):
        check(self.lib.covo_set_profiling(self._h, int(on)))

    def kernel_ms(self) -> np.ndarray:
        out = np.empty(6, dtype=np.float32)
        check(self.lib.covo_get_kernel_ms(self._h, fptr(out)))
        return out

    def phase_clocks(self, on=True, read=False):
        out = np.zeros(64, dtype=np.int64) if read else None
        check(self.lib.covo_debug_phase_clocks(self._h, int(on), None if out is None else out.ctypes.data_as(C.POINTER(C.c_longlong))))
        return out

    def sigma_path(self) -> int:
        v = C.c_int()
        check(self.lib.covo_get_sigma_path(self._h, C.byref(v)))
        return v.value

    def set_sigma_path(self, path: int):
        check(self.lib.covo_set_sigma_path(self._h, int(path)))

    def kernel_slot_names(self):
        """Names of the six covo_get_kernel_ms slots for this handle's optimize_sigma path."""
        mid = ["lanczos", "pole_inverses", "combine"] if self.sigma_path() else ["tridiag", "trifunc", "sandwich"]
        return ["hessian", *mid, "cholesky", "rollout"]

    def rng_step(self) -> int:
        v = C.c_uint()
        check(self.lib.covo_rng_step(self._h, C.byref(v)))
        return v.value


def zolotarev_nodes(m: float, M: float, n_poles: int):
    t = np.empty(n_poles, dtype=np.float64)
    w = np.empty(n_poles, dtype=np.float64)
    check(load().covo_zolotarev_nodes(m, M, n_poles, t.ctypes.data_as(_D), w.ctypes.data_as(_D)))
    return t, w
