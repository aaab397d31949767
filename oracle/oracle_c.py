"""ctypes front end of oracle/covo_oracle.c (C/OpenMP restatement of the reference's two heavy loops).

TEST INFRASTRUCTURE / CPU BASELINE ONLY -- see the header of covo_oracle.c.  Everything that is not a heavy
loop (shift, eigh-based optimize_sigma via LAPACK, Cholesky, the sampling GEMM, the softmax update) is the
NumPy oracle, i.e. the same LAPACK/BLAS class of routines jax.numpy dispatches to on CPU."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle_np as o

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
FP = C.POINTER(C.c_float)


def _load():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libcovo_oracle.so")
        src = os.path.join(HERE, "covo_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(src) > os.path.getmtime(path):
            subprocess.check_call(["make", "-s", "-C", HERE])
        lib = C.CDLL(path)
        lib.oracle_rollout_costs.argtypes = [FP, FP, C.c_int, FP, FP, C.c_int, FP, C.c_int, C.c_int, C.c_float, FP]
        lib.oracle_hessian_fof.argtypes = [FP, FP, C.c_int, FP, FP, C.c_int, FP, C.c_int, FP]
        DP = C.POINTER(C.c_double)
        lib.oracle_hessian_fof_f64.argtypes = [FP, FP, C.c_int, FP, FP, C.c_int, FP, C.c_int, DP]
        lib.oracle_num_threads.restype = C.c_int
        _LIB = lib
    return _LIB


def available() -> bool:
    try:
        _load()
        return True
    except Exception:
        return False


def num_threads() -> int:
    return int(_load().oracle_num_threads())


def _envp(p: o.EnvParams) -> np.ndarray:
    return np.array([p.m, p.g, p.max_thrust, p.dt, p.alpha_bodyrate, p.action_scale, p.pos_limit, *p.max_omega,
                     p.max_steps_in_episode], dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(FP)


def rollout_costs(ns: o.QuadState, a_sampled: np.ndarray, p: o.EnvParams, discount: float = 1.0) -> np.ndarray:
    a = np.ascontiguousarray(a_sampled, dtype=np.float32)
    N, H, _ = a.shape
    st = o.state_to_vec24(ns)
    pt = np.ascontiguousarray(ns.pos_traj, dtype=np.float32)
    vt = np.ascontiguousarray(ns.vel_traj, dtype=np.float32)
    cost = np.empty(N, dtype=np.float32)
    envp = _envp(p)
    _load().oracle_rollout_costs(_p(envp), _p(st), int(ns.time), _p(pt), _p(vt), pt.shape[0], _p(a), N, H, discount, _p(cost))
    return cost


def hessian(ns: o.QuadState, a_mean: np.ndarray, p: o.EnvParams) -> np.ndarray:
    am = np.ascontiguousarray(a_mean, dtype=np.float32)
    H = am.shape[0]
    st = o.state_to_vec24(ns)
    pt = np.ascontiguousarray(ns.pos_traj, dtype=np.float32)
    vt = np.ascontiguousarray(ns.vel_traj, dtype=np.float32)
    R = np.zeros((4 * H, 4 * H), dtype=np.float32)
    envp = _envp(p)
    _load().oracle_hessian_fof(_p(envp), _p(st), int(ns.time), _p(pt), _p(vt), pt.shape[0], _p(am), H, _p(R))
    return R


def hessian_f64(ns: o.QuadState, a_mean: np.ndarray, p: o.EnvParams) -> np.ndarray:
    """The same forward-over-forward Hessian in float64 arithmetic on the same float32 inputs: what the float32 results
    (the reference's, the device's) are rounding-error perturbations of."""
    am = np.ascontiguousarray(a_mean, dtype=np.float32)
    H = am.shape[0]
    st = o.state_to_vec24(ns)
    pt = np.ascontiguousarray(ns.pos_traj, dtype=np.float32)
    vt = np.ascontiguousarray(ns.vel_traj, dtype=np.float32)
    R = np.zeros((4 * H, 4 * H), dtype=np.float64)
    envp = _envp(p)
    _load().oracle_hessian_fof_f64(_p(envp), _p(st), int(ns.time), _p(pt), _p(vt), pt.shape[0], _p(am), H,
                                   R.ctypes.data_as(C.POINTER(C.c_double)))
    return R


def covo_step(ns: o.QuadState, a_mean_prev: np.ndarray, eps: np.ndarray, p: o.EnvParams, lam: float, online: bool = True,
              a_cov=None, sigma: float = 0.5):
    """CoVOController.__call__ (controllers/covo.py:187-283), heavy loops in C/OpenMP."""
    a_mean = o.shift_mean(a_mean_prev.astype(np.float32))
    if online or a_cov is None:
        R = hessian(ns, a_mean, p)
        a_cov = o.optimize_sigma(R, sigma, dtype=np.float32)
    L = np.linalg.cholesky(a_cov.astype(np.float64)).astype(np.float32)
    a_s = o.sample_actions(a_mean, L, eps)
    cost = rollout_costs(ns, a_s, p)
    new_mean, _ = o.softmax_update(a_mean, a_s, cost, lam)
    return new_mean[0].copy(), new_mean
