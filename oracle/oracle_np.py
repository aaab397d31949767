"""CPU restatement (NumPy) of the CoVO-MPC / MPPI hot path of LeCAR-Lab/CoVO-MPC.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``covo_mpc_b200``)
may import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as
the checker / the thing timed as the CPU baseline.

PARITY: pinned by the reference's own source executed in this container under a
NumPy stand-in for jax.numpy (tests/golden/make_reference_golden.py writes
tests/golden/reference_*.npz; tests/test_reference_golden.py holds this module
to them: step_env chains, geometry, reward, optimize_sigma, PID, get_controller
defaults, whole CoVO-online / MPPI controller calls including the Hessian of the
reference's own cost function, and the head of the covo-offline schedule).
UNPINNED: the arithmetic that lives in the un-vendored, un-pinned ``jax`` /
``jaxlib`` dependency -- Threefry PRNG bit-streams, XLA's float32 rounding and
JAX's forward-mode AD itself (jax is not installable here: no wheel, no
network).  Also checked (``tests/test_oracle.py``): analytic invariants of the
reference's maths, finite-difference agreement of the exact Hessian in float64,
Random123 known-answer vectors for the counter RNG.

Every function cites the reference file:line (relative to /root/reference)
whose behaviour it follows.  All maths is written once, generically, over
"scalars" that may be Python floats, NumPy arrays (a batch of N rollouts) or
second-order ``Jet`` objects (value, gradient, Hessian) -- the latter is the
restatement of ``jax.jacfwd(jax.jacfwd(cost))`` (controllers/covo.py:183-185):
forward-over-forward propagation of all tangent lanes through the unrolled
H-step rollout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# constants  (quadjax/dynamics/dataclass.py:40-100)
# --------------------------------------------------------------------------


@dataclass
class EnvParams:
    """Subset of EnvParams3D (dynamics/dataclass.py:40-100) used on the hot path."""

    max_torque: Tuple[float, float, float] = (9e-3, 9e-3, 2e-3)  # :43
    max_omega: Tuple[float, float, float] = (10.0, 10.0, 3.0)  # :44
    max_thrust: float = 0.8  # :45
    dt: float = 0.02  # :46
    g: float = 9.81  # :47
    m: float = 0.027  # :49
    action_scale: float = 1.0  # :71
    alpha_bodyrate: float = 0.5  # :76
    max_steps_in_episode: int = 300  # :81
    disturb_scale: float = 0.2  # :89
    dyn_noise_scale: float = 0.05  # :99
    obs_noise_scale: float = 0.05  # :100
    pos_limit: float = 3.0  # envs/quadrotor.py:484


# --------------------------------------------------------------------------
# second-order jets: forward-over-forward AD  (controllers/covo.py:183-185)
# --------------------------------------------------------------------------


class Jet:
    """value + gradient (n,) + Hessian (n,n) of a scalar w.r.t. n inputs.

    ``g is None`` / ``h is None`` mean "identically zero" (JAX's symbolic-zero
    tangents); this keeps constants cheap and reproduces the fact that JAX does
    not evaluate JVP rules on values that do not depend on the inputs.
    """

    __slots__ = ("v", "g", "h")

    def __init__(self, v, g=None, h=None):
        self.v = v
        self.g = g
        self.h = h

    # -- helpers ------------------------------------------------------------
    @staticmethod
    def _lift(x):
        return x if isinstance(x, Jet) else Jet(x)

    @staticmethod
    def _axpy(a, x, b, y):
        """a*x + b*y for arrays that may be None."""
        if x is None and y is None:
            return None
        if x is None:
            return b * y
        if y is None:
            return a * x
        return a * x + b * y

    def _unary(self, f0, f1, f2):
        """Chain rule for out = f(self): f0=f(v), f1=f'(v), f2=f''(v)."""
        g = None if self.g is None else f1 * self.g
        h = None
        if self.h is not None:
            h = f1 * self.h
        if self.g is not None and f2 is not None:
            o = f2 * np.outer(self.g, self.g)
            h = o if h is None else h + o
        return Jet(f0, g, h)

    # -- arithmetic ---------------------------------------------------------
    def __add__(self, o):
        o = Jet._lift(o)
        return Jet(self.v + o.v, Jet._axpy(1.0, self.g, 1.0, o.g), Jet._axpy(1.0, self.h, 1.0, o.h))

    __radd__ = __add__

    def __neg__(self):
        return Jet(-self.v, None if self.g is None else -self.g, None if self.h is None else -self.h)

    def __sub__(self, o):
        o = Jet._lift(o)
        return Jet(self.v - o.v, Jet._axpy(1.0, self.g, -1.0, o.g), Jet._axpy(1.0, self.h, -1.0, o.h))

    def __rsub__(self, o):
        return Jet._lift(o) - self

    def __mul__(self, o):
        o = Jet._lift(o)
        g = Jet._axpy(o.v, self.g, self.v, o.g)
        h = Jet._axpy(o.v, self.h, self.v, o.h)
        if self.g is not None and o.g is not None:
            c = np.outer(self.g, o.g)
            c = c + c.T
            h = c if h is None else h + c
        return Jet(self.v * o.v, g, h)

    __rmul__ = __mul__

    def recip(self):
        r = 1.0 / self.v
        return self._unary(r, -r * r, 2.0 * r * r * r)

    def __truediv__(self, o):
        o = Jet._lift(o)
        return self * o.recip()

    def __rtruediv__(self, o):
        return Jet._lift(o) * self.recip()


def _is_jet(x):
    return isinstance(x, Jet)


def m_sqrt(x):
    if _is_jet(x):
        s = math.sqrt(x.v) if np.ndim(x.v) == 0 else np.sqrt(x.v)
        if np.ndim(s) == 0 and s == 0.0:
            # d/dx sqrt at 0 is singular; the reference would emit NaN (0*inf) only if
            # the tangent were instantiated.  EXTENSION: treat as locally constant.
            return Jet(s)
        return x._unary(s, 0.5 / s, -0.25 / (s * x.v))
    return np.sqrt(x)


def m_log(x):
    if _is_jet(x):
        return x._unary(math.log(x.v), 1.0 / x.v, -1.0 / (x.v * x.v))
    return np.log(x)


def m_abs(x):
    if _is_jet(x):
        # jax.numpy.abs JVP = sign(x) * tangent  (sign(0) = 0)
        return x._unary(abs(x.v), float(np.sign(x.v)), None)
    return np.abs(x)


def m_atan2(y, x):
    if _is_jet(y) or _is_jet(x):
        y = Jet._lift(y)
        x = Jet._lift(x)
        r = x.v * x.v + y.v * y.v
        # d atan2(y, x) = (x dy - y dx) / r
        g = Jet._axpy(x.v / r, y.g, -y.v / r, x.g)
        h = Jet._axpy(x.v / r, y.h, -y.v / r, x.h)
        if g is not None:
            n = g.shape[0]
            yg = y.g if y.g is not None else np.zeros(n, dtype=g.dtype)
            xg = x.g if x.g is not None else np.zeros(n, dtype=g.dtype)
            # d/dj of (x y_i - y x_i)/r with x_ij, y_ij handled above:
            #   (x_j y_i - y_j x_i)/r  -  (x y_i - y x_i) (2 x x_j + 2 y y_j) / r^2
            cross = (np.outer(yg, xg) - np.outer(xg, yg)) / r
            w = x.v * yg - y.v * xg
            dr = 2.0 * (x.v * xg + y.v * yg)
            second = cross - np.outer(w, dr) / (r * r)
            h = second if h is None else h + second
        return Jet(math.atan2(y.v, x.v), g, h)
    return np.arctan2(y, x)


def m_clip(x, lo, hi):
    """jnp.clip == minimum(maximum(x, lo), hi) (jax.numpy, 0.4.x).  lax.max/min JVPs
    split ties evenly, so the derivative is 1 inside, 0 outside and 0.5 exactly on a bound."""
    if _is_jet(x):
        v = x.v
        if v < lo or v > hi:
            return Jet(min(max(v, lo), hi))
        w = 0.5 if (v == lo or v == hi) else 1.0
        return Jet(v, None if x.g is None else w * x.g, None if x.h is None else w * x.h)
    return np.minimum(np.maximum(x, lo), hi)


def m_norm(c: Sequence):
    """jnp.linalg.norm of a short vector = sqrt(sum(x*x))."""
    s = c[0] * c[0]
    for k in range(1, len(c)):
        s = s + c[k] * c[k]
    return m_sqrt(s)


# --------------------------------------------------------------------------
# state container (SoA: every component is a "scalar" in the sense above)
# --------------------------------------------------------------------------


@dataclass
class QuadState:
    """Hot-path subset of EnvState3D (dynamics/dataclass.py:10-37)."""

    pos: list
    quat: list  # (x, y, z, w)  dataclass.py:14
    vel: list
    omega: list
    f_disturb: list
    time: object  # int or int array
    pos_tar: list
    vel_tar: list
    pos_traj: np.ndarray  # (T,3)
    vel_traj: np.ndarray  # (T,3)

    def copy(self):
        return replace(
            self,
            pos=list(self.pos),
            quat=list(self.quat),
            vel=list(self.vel),
            omega=list(self.omega),
            f_disturb=list(self.f_disturb),
            pos_tar=list(self.pos_tar),
            vel_tar=list(self.vel_tar),
        )


def make_state(pos, quat, vel, omega, f_disturb, time, pos_traj, vel_traj, pos_tar=None, vel_tar=None, dtype=np.float64):
    pos_traj = np.asarray(pos_traj, dtype=dtype)
    vel_traj = np.asarray(vel_traj, dtype=dtype)
    T = pos_traj.shape[0]
    ti = min(int(time), T - 1)
    c = lambda a: [dtype(x) for x in a]
    return QuadState(
        pos=c(pos),
        quat=c(quat),
        vel=c(vel),
        omega=c(omega),
        f_disturb=c(f_disturb),
        time=int(time),
        pos_tar=c(pos_traj[ti] if pos_tar is None else pos_tar),
        vel_tar=c(vel_traj[ti] if vel_tar is None else vel_tar),
        pos_traj=pos_traj,
        vel_traj=vel_traj,
    )


def state_to_vec24(s: QuadState) -> np.ndarray:
    """Pack into the C-ABI layout (include/covo_b200.h: covo_state24)."""
    out = np.zeros(24, dtype=np.float32)
    out[0:3] = s.pos
    out[3:7] = s.quat
    out[7:10] = s.vel
    out[10:13] = s.omega
    out[13:16] = s.f_disturb
    out[16:19] = s.pos_tar
    out[19:22] = s.vel_tar
    return out


# --------------------------------------------------------------------------
# reward / termination   (dynamics/utils.py:266-294, envs/quadrotor.py:479-503)
# --------------------------------------------------------------------------


def log_pos_fn(err_pos):
    """dynamics/utils.py:266-274."""
    l = m_log(err_pos + 1.0)
    return (
        err_pos * 0.4
        + m_clip(l * 4.0, 0.0, 1.0) * 0.4
        + m_clip(l * 8.0, 0.0, 1.0) * 0.2
        + m_clip(l * 16.0, 0.0, 1.0) * 0.1
        + m_clip(l * 32.0, 0.0, 1.0) * 0.1
    )


def tracking_penyaw_reward(s: QuadState):
    """dynamics/utils.py:285-294 -- bound to tracking / tracking_zigzag / hovering
    (envs/quadrotor.py:56,73,81).  Uses the *stored* (possibly un-normalised) quaternion."""
    err_pos = m_norm([s.pos_tar[k] - s.pos[k] for k in range(3)])
    err_vel = m_norm([s.vel_tar[k] - s.vel[k] for k in range(3)])
    q = s.quat
    yaw = m_atan2(2.0 * (q[3] * q[2] + q[0] * q[1]), 1.0 - 2.0 * (q[1] * q[1] + q[2] * q[2]))
    return 1.3 - 0.05 * err_vel - log_pos_fn(err_pos) - m_abs(yaw) * 0.2


def is_terminal(s: QuadState, p: EnvParams):
    """envs/quadrotor.py:479-503 with disable_rollover_terminate=True (main, :779)."""
    pv = [x.v if _is_jet(x) else x for x in s.pos]
    out = np.asarray(s.time) >= p.max_steps_in_episode
    for k in range(3):
        out = out | (np.abs(pv[k]) > p.pos_limit)
    return out


# --------------------------------------------------------------------------
# dynamics   (envs/quadrotor.py:215-263, dynamics/free.py:74-202, geom.py:35-77)
# --------------------------------------------------------------------------


def _gather_clamped(traj: np.ndarray, t):
    """JAX out-of-range gather clamps (dynamics/free.py:153-155)."""
    T = traj.shape[0]
    ti = np.minimum(np.asarray(t), T - 1)
    return traj[ti]


def step_env(s: QuadState, action: Sequence, p: EnvParams, f_disturb_next=None) -> QuadState:
    """One ``Quad3D.step_env`` transition of the dynamic state (the reward/done of the
    PRE-step state are evaluated by the caller, envs/quadrotor.py:243-244).

    ``f_disturb_next`` is the value ``disturb_func`` returns (dynamics/free.py:144-147):
    zeros for ``disturb_type == 'none'`` and for ``'gaussian'`` with deterministic=True
    (envs/quadrotor.py:234-235); the caller supplies the draw otherwise.
    """
    dt = p.dt
    # envs/quadrotor.py:223 and :257 -- two clips
    a = [m_clip(m_clip(action[k], -1.0, 1.0), -1.0, 1.0) for k in range(4)]
    thrust = (a[0] + 1.0) / 2.0 * p.max_thrust  # :258
    torque = [a[1 + k] * p.max_torque[k] for k in range(3)]  # :259
    omega_tar = [torque[k] / p.max_torque[k] * p.max_omega[k] for k in range(3)]  # free.py:122
    thrust = thrust * p.action_scale  # free.py:82
    omega_tar = [w * p.action_scale for w in omega_tar]

    qn = m_norm(s.quat)  # free.py:88
    q = [s.quat[k] / qn for k in range(4)]
    x, y, z, w = q
    # third column of qtoQ(q) = H^T T L T L H  (geom.py:68-77)
    qe3 = [2.0 * (x * z + y * w), 2.0 * (y * z - x * w), 1.0 - 2.0 * (x * x + y * y)]
    om = s.omega
    # 0.5 * L(q) @ H @ omega  (geom.py:41-55, free.py:96)
    qdot = [
        0.5 * (w * om[0] + (y * om[2] - z * om[1])),
        0.5 * (w * om[1] + (z * om[0] - x * om[2])),
        0.5 * (w * om[2] + (x * om[1] - y * om[0])),
        0.5 * (-(x * om[0] + y * om[1] + z * om[2])),
    ]
    inv_m = 1.0 / p.m
    vdot = [
        inv_m * (qe3[0] * thrust + s.f_disturb[0]),
        inv_m * (qe3[1] * thrust + s.f_disturb[1]),
        -p.g + inv_m * (qe3[2] * thrust + s.f_disturb[2]),
    ]  # free.py:97-99
    pos_new = [s.pos[k] + s.vel[k] * dt for k in range(3)]  # explicit Euler, old v (free.py:102)
    q_new = [q[k] + qdot[k] * dt for k in range(4)]
    vel_new = [s.vel[k] + vdot[k] * dt for k in range(3)]
    al = p.alpha_bodyrate
    om_new = [al * om[k] + (1.0 - al) * omega_tar[k] for k in range(3)]  # free.py:105-107
    qn2 = m_norm(q_new)  # free.py:139
    q_new = [q_new[k] / qn2 for k in range(4)]

    time = s.time + 1  # free.py:150
    ptar = _gather_clamped(s.pos_traj, time)
    vtar = _gather_clamped(s.vel_traj, time)
    if f_disturb_next is None:
        zero = 0.0 * (pos_new[0].v if _is_jet(pos_new[0]) else pos_new[0])
        f_disturb_next = [zero, zero, zero]
    return QuadState(
        pos=pos_new,
        quat=q_new,
        vel=vel_new,
        omega=om_new,
        f_disturb=list(f_disturb_next),
        time=time,
        pos_tar=[ptar[..., k] for k in range(3)],
        vel_tar=[vtar[..., k] for k in range(3)],
        pos_traj=s.pos_traj,
        vel_traj=s.vel_traj,
    )


# --------------------------------------------------------------------------
# sampling + rollout + softmax update   (controllers/covo.py:201-283, mppi.py:46-134)
# --------------------------------------------------------------------------


def shift_mean(a_mean: np.ndarray) -> np.ndarray:
    """controllers/covo.py:201-203."""
    return np.concatenate([a_mean[1:], a_mean[-1:]], axis=0)


def sample_actions(a_mean: np.ndarray, L: np.ndarray, eps: np.ndarray) -> np.ndarray:
    """mean + chol(cov) @ eps, clipped  (jax.random.multivariate_normal, method='cholesky';
    controllers/covo.py:215-224).  a_mean (H,4); L (n,n) lower; eps (N,n) -> (N,H,4)."""
    dt = a_mean.dtype
    a = a_mean.reshape(1, -1) + (eps.astype(dt) @ L.astype(dt).T)
    return np.clip(a, -1.0, 1.0).reshape(eps.shape[0], a_mean.shape[0], a_mean.shape[1]).astype(dt)


def sample_actions_blockdiag(a_mean: np.ndarray, Lblk: np.ndarray, eps: np.ndarray) -> np.ndarray:
    """MPPI: per-step 4x4 Gaussians (controllers/mppi.py:56-66).  Lblk (H,4,4) lower; eps (N,H,4)."""
    dt = a_mean.dtype
    a = a_mean[None] + np.einsum("hab,nhb->nha", Lblk.astype(dt), eps.astype(dt))
    return np.clip(a, -1.0, 1.0).astype(dt)


def broadcast_state(s: QuadState, N: int, dtype) -> QuadState:
    """jax.tree_map(repeat N) (controllers/covo.py:240-245), without copying the trajectories."""
    rep = lambda c: [np.full(N, x, dtype=dtype) for x in c]
    return QuadState(
        pos=rep(s.pos),
        quat=rep(s.quat),
        vel=rep(s.vel),
        omega=rep(s.omega),
        f_disturb=rep(s.f_disturb),
        time=np.full(N, s.time, dtype=np.int64),
        pos_tar=rep(s.pos_tar),
        vel_tar=rep(s.vel_tar),
        pos_traj=s.pos_traj.astype(dtype),
        vel_traj=s.vel_traj.astype(dtype),
    )


def rollout_costs(s0: QuadState, a_sampled: np.ndarray, p: EnvParams, discount: float = 1.0,
                  f_disturb_seq: Optional[np.ndarray] = None, return_pos: bool = False):
    """lax.scan over H of vmap over N of step_env (controllers/covo.py:227-263).

    Reward/done are those of the pre-step state; once done, the reward freezes at the last
    pre-termination value (``where(done_before, reward_before, reward)``, covo.py:233)."""
    N, H, _ = a_sampled.shape
    dt = a_sampled.dtype.type
    s = broadcast_state(s0, N, dt)
    reward_before = np.zeros(N, dtype=dt)
    done_before = np.zeros(N, dtype=bool)
    disc_sum = np.zeros(N, dtype=dt)
    poses = []
    for h in range(H):
        a = [a_sampled[:, h, k] for k in range(4)]
        reward = tracking_penyaw_reward(s).astype(dt)
        done = is_terminal(s, p)
        fd = None if f_disturb_seq is None else [np.full(N, f_disturb_seq[h, k], dtype=dt) for k in range(3)]
        s = step_env(s, a, p, fd)
        reward = np.where(done_before, reward_before, reward)
        reward_before = reward
        done_before = done | done_before
        disc_sum = disc_sum + reward * dt(discount) ** h
        if return_pos:
            poses.append(np.stack(s.pos, axis=-1))
    cost = -disc_sum
    if return_pos:
        return cost, np.stack(poses, axis=0)  # (H,N,3): env_state.pos AFTER each step (covo.py:236)
    return cost


def softmax_update(a_mean: np.ndarray, a_sampled: np.ndarray, cost: np.ndarray, lam: float, gamma_mean: float = 1.0):
    """controllers/covo.py:266-275."""
    dt = a_sampled.dtype.type
    cost_exp = np.exp(-(cost - np.min(cost)) / dt(lam))
    weight = cost_exp / np.sum(cost_exp)
    new = np.sum(weight[:, None, None] * a_sampled, axis=0) * dt(gamma_mean) + a_mean * dt(1.0 - gamma_mean)
    return new.astype(a_sampled.dtype), weight


def softmax_partials(a_sampled: np.ndarray, cost: np.ndarray, lam: float):
    """(m, s, v) triple of one shard: m=min cost, s=sum exp(-(c-m)/lam), v=sum exp(.)*a."""
    m = np.min(cost)
    e = np.exp(-(cost - m) / lam)
    return m, np.sum(e), np.sum(e[:, None, None] * a_sampled, axis=0)


def merge_partials(parts, lam: float):
    """Overflow-safe merge of shard triples in rank order."""
    M = min(p[0] for p in parts)
    S = 0.0
    V = 0.0
    for m, s, v in parts:
        sc = np.exp(-(m - M) / lam)
        S = S + s * sc
        V = V + v * sc
    return M, S, V


# --------------------------------------------------------------------------
# exact Hessian   (controllers/covo.py:134-185)
# --------------------------------------------------------------------------


def cumulated_cost(s0: QuadState, a_flat, p: EnvParams, H: int):
    """get_cumulated_cost (controllers/covo.py:165-180): -(sum_h r(x_h)) over the Python-unrolled
    rollout with deterministic=True; no termination freeze, no discount.  The extra
    ``reward_fn(initial state)`` term (:176-178) is constant in the controls and kept for fidelity."""
    s = s0
    total = 0.0
    for h in range(H):
        r = tracking_penyaw_reward(s)
        total = total + r
        s = step_env(s, [a_flat[4 * h + k] for k in range(4)], p)
    total = total + tracking_penyaw_reward(s0)
    return -total


def get_hessian(s0: QuadState, a_mean: np.ndarray, p: EnvParams, dtype=np.float64) -> np.ndarray:
    """jacfwd(jacfwd(get_cumulated_cost)) (controllers/covo.py:183-185) by second-order jets."""
    H = a_mean.shape[0]
    n = 4 * H
    flat = a_mean.reshape(-1).astype(dtype)
    eye = np.eye(n, dtype=dtype)
    seeds = [Jet(dtype(flat[i]), eye[i].copy(), None) for i in range(n)]
    cast = lambda c: [dtype(x) for x in c]
    s = QuadState(
        pos=cast(s0.pos), quat=cast(s0.quat), vel=cast(s0.vel), omega=cast(s0.omega),
        f_disturb=cast(s0.f_disturb), time=int(s0.time), pos_tar=cast(s0.pos_tar), vel_tar=cast(s0.vel_tar),
        pos_traj=s0.pos_traj.astype(dtype), vel_traj=s0.vel_traj.astype(dtype),
    )
    c = cumulated_cost(s, seeds, p, H)
    Hm = c.h if (_is_jet(c) and c.h is not None) else np.zeros((n, n), dtype=dtype)
    return np.asarray(Hm, dtype=dtype)


def hessian_fd(s0: QuadState, a_mean: np.ndarray, p: EnvParams, step: float = 1e-4, chunk: int = 40000) -> np.ndarray:
    """Central finite differences of the float64 cost (independent check of the jets)."""
    H = a_mean.shape[0]
    n = 4 * H
    x0 = a_mean.reshape(-1).astype(np.float64)
    s64 = make_state(s0.pos, s0.quat, s0.vel, s0.omega, s0.f_disturb, s0.time, s0.pos_traj, s0.vel_traj,
                     s0.pos_tar, s0.vel_tar, dtype=np.float64)

    def f(X):  # X (B,n) -> (B,)
        B = X.shape[0]
        s = broadcast_state(s64, B, np.float64)
        tot = np.zeros(B)
        for hh in range(H):
            tot = tot + tracking_penyaw_reward(s)
            s = step_env(s, [X[:, 4 * hh + k] for k in range(4)], p)
        return -tot

    ii, jj = np.triu_indices(n)
    R = np.zeros((n, n))
    for c0 in range(0, ii.size, chunk):
        i = ii[c0:c0 + chunk]
        j = jj[c0:c0 + chunk]
        B = i.size
        acc = np.zeros(B)
        for si, sj, sg in ((1, 1, 1.0), (1, -1, -1.0), (-1, 1, -1.0), (-1, -1, 1.0)):
            X = np.repeat(x0[None], B, axis=0)
            np.add.at(X, (np.arange(B), i), si * step)
            np.add.at(X, (np.arange(B), j), sj * step)
            acc += sg * f(X)
        R[i, j] = acc / (4 * step * step)
    R = np.triu(R) + np.triu(R, 1).T
    return R


# --------------------------------------------------------------------------
# CoVO covariance   (controllers/covo.py:116-132) and Cholesky (inside covo.py:216)
# --------------------------------------------------------------------------


def optimize_sigma(R: np.ndarray, sample_sigma: float, dtype=np.float32) -> np.ndarray:
    """controllers/covo.py:116-132 (jnp.linalg.eigh -> LAPACK syevd; numpy.linalg.eigh is the
    same LAPACK driver)."""
    R = np.asarray(R, dtype=dtype)
    R = (R + R.T) / dtype(2.0)
    eigns, u = np.linalg.eigh(R)
    min_eign = np.min(eigns)
    offset = -min_eign + dtype(1e-2)
    eigns = eigns + offset
    log_o = np.log(eigns)
    n = R.shape[0]
    log_det_a_cov = dtype(n) * (np.log(dtype(sample_sigma)) * dtype(2))
    log_const = (log_det_a_cov * dtype(2) + np.sum(log_o)) / dtype(n)
    log_s = dtype(0.5) * log_const - dtype(0.5) * log_o
    a_cov = (u * np.exp(log_s)[None, :]) @ u.T
    return ((a_cov + a_cov.T) / dtype(2.0)).astype(dtype)


def cholesky_lower(cov: np.ndarray) -> np.ndarray:
    """Lower Cholesky factor used by jax.random.multivariate_normal(method='cholesky')."""
    return np.linalg.cholesky(cov)


# --------------------------------------------------------------------------
# controllers   (controllers/covo.py:187-283, mppi.py:28-134)
# --------------------------------------------------------------------------


def hover_mean(H: int, p: EnvParams, dtype=np.float32) -> np.ndarray:
    """get_sample_mean (envs/quadrotor.py:685-690)."""
    th = (p.m * p.g / p.max_thrust) * 2.0 - 1.0
    return np.tile(np.array([th, 0.0, 0.0, 0.0], dtype=dtype), (H, 1))


def covo_call(noisy_state: QuadState, a_mean_prev: np.ndarray, eps: np.ndarray, p: EnvParams, *, lam: float,
              sigma: float = 0.5, gamma_mean: float = 1.0, discount: float = 1.0, a_cov: Optional[np.ndarray] = None,
              dtype=np.float32, hessian_dtype=np.float64, return_debug: bool = False):
    """CoVOController.__call__ (controllers/covo.py:187-283) with the Gaussian draws ``eps`` (N,4H)
    supplied explicitly.  ``a_cov`` given  -> offline mode (table lookup, :107-108);
    ``a_cov`` None -> online mode (Hessian at the shifted mean + optimize_sigma, :36-41)."""
    a_mean = shift_mean(a_mean_prev.astype(dtype))
    dbg = {}
    if a_cov is None:
        R = get_hessian(noisy_state, a_mean, p, dtype=hessian_dtype)
        a_cov = optimize_sigma(R, sigma, dtype=dtype)
        dbg["R"] = R
    a_cov = np.asarray(a_cov, dtype=dtype)
    L = cholesky_lower(a_cov.astype(np.float64)).astype(dtype) if dtype == np.float32 else cholesky_lower(a_cov)
    a_s = sample_actions(a_mean, L, eps)
    s0 = make_state(noisy_state.pos, noisy_state.quat, noisy_state.vel, noisy_state.omega, noisy_state.f_disturb,
                    noisy_state.time, noisy_state.pos_traj, noisy_state.vel_traj, noisy_state.pos_tar,
                    noisy_state.vel_tar, dtype=dtype)
    cost, poses = rollout_costs(s0, a_s, p, discount, return_pos=True)
    new_mean, weight = softmax_update(a_mean, a_s, cost, lam, gamma_mean)
    u = new_mean[0].copy()
    info = {"pos_mean": poses.mean(axis=1), "pos_std": poses.std(axis=1)}
    if return_debug:
        dbg.update(a_cov=a_cov, L=L, a_sampled=a_s, cost=cost, weight=weight, a_mean_shifted=a_mean)
        return u, new_mean, a_cov, info, dbg
    return u, new_mean, a_cov, info


def mppi_call(noisy_state: QuadState, a_mean_prev: np.ndarray, a_cov_prev: np.ndarray, eps: np.ndarray, p: EnvParams, *,
              lam: float, gamma_mean: float = 1.0, gamma_sigma: float = 0.0, discount: float = 1.0,
              f_disturb_seq: Optional[np.ndarray] = None, dtype=np.float32, return_debug: bool = False):
    """MPPIController.__call__ (controllers/mppi.py:28-134).  eps (N,H,4).  ``f_disturb_seq`` (H,3) is the
    force every sample sees after step h under a non-deterministic 'gaussian' disturbance (mppi.py:74:
    all samples share ``step_key``); None == disturb_type 'none'."""
    a_mean = shift_mean(a_mean_prev.astype(dtype))
    a_cov = np.concatenate([a_cov_prev[1:], a_cov_prev[-1:]], axis=0).astype(dtype)  # mppi.py:46-49
    Lblk = np.linalg.cholesky(a_cov.astype(np.float64)).astype(dtype)
    a_s = sample_actions_blockdiag(a_mean, Lblk, eps)
    s0 = make_state(noisy_state.pos, noisy_state.quat, noisy_state.vel, noisy_state.omega, noisy_state.f_disturb,
                    noisy_state.time, noisy_state.pos_traj, noisy_state.vel_traj, noisy_state.pos_tar,
                    noisy_state.vel_tar, dtype=dtype)
    cost, poses = rollout_costs(s0, a_s, p, discount, f_disturb_seq=f_disturb_seq, return_pos=True)
    new_mean, weight = softmax_update(a_mean, a_s, cost, lam, gamma_mean)
    d = a_s - new_mean[None]
    new_cov = np.sum(weight[:, None, None, None] * (d[..., None] * d[:, :, None, :]), axis=0) * dtype(gamma_sigma) \
        + a_cov * dtype(1.0 - gamma_sigma)  # mppi.py:119-125
    u = new_mean[0].copy()
    info = {"pos_mean": poses.mean(axis=1), "pos_std": poses.std(axis=1)}
    if return_debug:
        return u, new_mean, new_cov.astype(dtype), info, dict(a_sampled=a_s, cost=cost, weight=weight, Lblk=Lblk)
    return u, new_mean, new_cov.astype(dtype), info


# --------------------------------------------------------------------------
# PID expansion policy + offline schedule   (controllers/pid.py:38-83, covo.py:44-112)
# --------------------------------------------------------------------------


def _qtoQ(q):
    """geom.qtoQ (geom.py:68-77) for an un-normalised quaternion: |q|^2 * R(q/|q|)."""
    x, y, z, w = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def _hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def pid_action(s: QuadState, p: EnvParams, Kp=10.0, Kd=5.0, Ki=0.0, Kp_att=10.0, integral=None, acc_tar=None):
    """PIDController.__call__ (controllers/pid.py:38-83) with the gains CoVO-offline uses
    (controllers/covo.py:48-53)."""
    pos = np.array(s.pos, dtype=np.float64)
    vel = np.array(s.vel, dtype=np.float64)
    q = np.array(s.quat, dtype=np.float64)
    integral = np.zeros(3) if integral is None else integral
    acc_tar = np.zeros(3) if acc_tar is None else acc_tar
    Q = _qtoQ(q)
    f_d = p.m * (np.array([0.0, 0.0, p.g]) - Kp * (pos - np.array(s.pos_tar)) - Kd * (vel - np.array(s.vel_tar))
                 - Ki * integral + acc_tar)
    thrust = float(np.clip((Q.T @ f_d)[2], 0.0, p.max_thrust))
    nrm = np.linalg.norm(f_d)
    nrm = 1e-3 if nrm < 1e-3 else nrm
    z_d = f_d / nrm
    axis_angle = np.cross(np.array([0.0, 0.0, 1.0]), z_d)
    angle = np.linalg.norm(axis_angle)
    small = angle < 1e-3
    angle = 5e-4 if small else angle  # pid.py:59
    # pid.py:60 tests the *updated* angle, which is never < 1e-3 -> axis = axis_angle / angle always
    axis = axis_angle / angle
    an = np.linalg.norm(axis)
    if an > 0:
        axis_n = axis / an  # geom.axisangletoR normalises (geom.py:111)
        Hx = _hat(axis_n)
        R_d = np.eye(3) + np.sin(angle) * Hx + (1 - np.cos(angle)) * Hx @ Hx
    else:
        # EXTENSION shared with csrc/pid.cuh: the reference divides 0/0 here (f_d exactly vertical) and
        # returns NaN; the zero-angle limit R_d = I is used instead.
        R_d = np.eye(3)
    R_e = R_d.T @ Q
    E = R_e - R_e.T
    angle_err = np.array([E[2, 1], E[0, 2], E[1, 0]])
    omega_d = -Kp_att * angle_err
    act = np.concatenate([[thrust / p.max_thrust * 2.0 - 1.0], omega_d / np.array(p.max_omega)])
    return act


# --------------------------------------------------------------------------
# environment side (caller of the hot path): reference trajectories, reset, env.step, noisy state
# --------------------------------------------------------------------------


def generate_zigzag_traj(max_steps: int, dt: float, rng: np.random.Generator):
    """Distribution of dynamics/utils.py:183-251 restated with a NumPy generator (the JAX
    Threefry stream itself is unpinned here, see module docstring).  8 segments x 40 points."""
    point_per_seg = 40
    num_seg = max_steps // point_per_seg + 1
    prev = rng.uniform(-1.0, 1.0, size=3)
    prev = prev / np.linalg.norm(prev) * 0.1
    pos_segs, vel_segs = [], []
    for _ in range(num_seg):
        to_c = -prev / np.linalg.norm(prev)
        dth, dph = rng.uniform(-np.pi / 3, np.pi / 3, size=2)
        theta = np.arccos(to_c[2]) + dth
        phi = np.arctan2(to_c[1], to_c[0]) + dph
        new_dir = np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
        dist = rng.uniform(1.0, 1.5)
        nxt = prev + dist * new_dir
        seg = np.stack([np.linspace(prev[k], nxt[k], point_per_seg, endpoint=False) for k in range(3)], axis=-1)
        vseg = (nxt - prev) / (point_per_seg + 1) * np.ones((point_per_seg, 3)) / dt  # sic, utils.py:231-236
        pos_segs.append(seg)
        vel_segs.append(vseg)
        prev = nxt
    pos = np.concatenate(pos_segs, axis=0)
    pos = pos - pos[0]
    vel = np.concatenate(vel_segs, axis=0)
    return pos, vel, np.zeros_like(pos)


def generate_lissa_traj(max_steps: int, dt: float, rng: np.random.Generator):
    """dynamics/utils.py:87-130 (task 'tracking')."""
    amp = rng.uniform(-1.0, 1.0, size=(3, 2))
    ph = rng.uniform(-np.pi, np.pi, size=(3, 2))
    ts = np.arange(0, max_steps + 50) * dt
    w1, w2 = 2 * np.pi * 0.2, 2 * np.pi * 0.4
    pos = np.stack([amp[i, 0] * np.sin(w1 * ts + ph[i, 0]) + amp[i, 1] * np.sin(w2 * ts + ph[i, 1]) for i in range(3)], axis=1)
    pos = pos - pos[0]
    vel = np.stack([amp[i, 0] * w1 * np.cos(w1 * ts + ph[i, 0]) + amp[i, 1] * w2 * np.cos(w2 * ts + ph[i, 1]) for i in range(3)], axis=1)
    acc = np.stack([-amp[i, 0] * w1 ** 2 * np.sin(w1 * ts + ph[i, 0]) - amp[i, 1] * w2 ** 2 * np.sin(w2 * ts + ph[i, 1]) for i in range(3)], axis=1)
    return pos, vel, acc


def generate_fixed_traj(max_steps: int, dt: float, rng=None):
    """dynamics/utils.py:49-53 (task 'hovering')."""
    z = np.zeros((max_steps, 3))
    return z, z.copy(), z.copy()


TRAJ_GENERATORS = {"tracking": generate_lissa_traj, "tracking_zigzag": generate_zigzag_traj, "hovering": generate_fixed_traj}


def reset_env(task: str, p: EnvParams, rng: np.random.Generator, dtype=np.float32, zero_disturb: bool = False) -> QuadState:
    """Quad3D.get_zero_state + reset_env (envs/quadrotor.py:265-312, 363-370)."""
    pos_traj, vel_traj, _ = TRAJ_GENERATORS[task](p.max_steps_in_episode, p.dt, rng)
    fd = rng.uniform(-p.disturb_scale, p.disturb_scale, size=3)  # :300-305 (drawn even for 'none')
    if zero_disturb:
        fd = np.zeros(3)
    return make_state(np.zeros(3), [0, 0, 0, 1.0], np.zeros(3), np.zeros(3), fd, 0, pos_traj, vel_traj, dtype=dtype)


def noisy_state(s: QuadState, p: EnvParams, rng: np.random.Generator) -> QuadState:
    """get_info's noisy_state (envs/quadrotor.py:323-351)."""
    sc = p.obs_noise_scale
    dt = type(s.pos[0])
    out = s.copy()
    out.pos = [dt(s.pos[k] + rng.standard_normal() * sc * 0.25) for k in range(3)]
    out.vel = [dt(s.vel[k] + rng.standard_normal() * sc * 0.5) for k in range(3)]
    out.quat = [dt(s.quat[k] + rng.standard_normal() * sc * 0.02) for k in range(4)]
    out.omega = [dt(s.omega[k] + rng.standard_normal() * sc * 0.5) for k in range(3)]
    return out


def env_step(s: QuadState, action: np.ndarray, p: EnvParams, rng: np.random.Generator, disturb_type: str = "none"):
    """BaseEnvironment.step -> Quad3D.step_env (envs/base.py:15-40, envs/quadrotor.py:215-248) for one
    environment, without the auto-reset branch (callers stop at ``done``).  Returns
    (next_state, reward, done, err_pos) with reward/done/err_pos of the PRE-step state."""
    dt = type(s.pos[0])
    reward = float(tracking_penyaw_reward(s))
    done = bool(is_terminal(s, p))
    err_pos = float(m_norm([s.pos_tar[k] - s.pos[k] for k in range(3)]))
    if disturb_type == "none":
        fd = [dt(0.0)] * 3
    elif disturb_type == "gaussian":
        fd = [dt(p.dyn_noise_scale * rng.standard_normal()) for _ in range(3)]
    else:
        raise NotImplementedError(disturb_type)
    nxt = step_env(s, [dt(a) for a in action], p, fd)
    nxt.pos = [dt(x) for x in nxt.pos]
    nxt.quat = [dt(x) for x in nxt.quat]
    nxt.vel = [dt(x) for x in nxt.vel]
    nxt.omega = [dt(x) for x in nxt.omega]
    nxt.pos_tar = [dt(x) for x in nxt.pos_tar]
    nxt.vel_tar = [dt(x) for x in nxt.vel_tar]
    nxt.time = int(nxt.time)
    return nxt, reward, done, err_pos


def covo_offline_schedule(s0: QuadState, p: EnvParams, H: int, sigma: float, rng: np.random.Generator,
                          n_steps: Optional[int] = None, disturb_type: str = "none", dtype=np.float32,
                          hessian_dtype=np.float64):
    """reset_a_cov_offline (controllers/covo.py:58-104): for each episode step, PID-policy H-step
    deterministic nominal rollout -> Hessian at that nominal -> optimize_sigma; then advance the env one
    (stochastic) PID step.  Returns a_cov_offline (T, 4H, 4H)."""
    T = p.max_steps_in_episode if n_steps is None else n_steps
    out = []
    s = s0.copy()
    for _ in range(T):
        sr = s.copy()
        nominal = []
        for _h in range(H):
            a = pid_action(sr, p)
            nominal.append(a)
            sr, _, _, _ = env_step(sr, a, p, rng, "none")  # deterministic=True (covo.py:67-69)
        a_mean = np.asarray(nominal)
        R = get_hessian(s, a_mean, p, dtype=hessian_dtype)
        out.append(optimize_sigma(R, sigma, dtype=dtype))
        a = pid_action(s, p)
        s, _, _, _ = env_step(s, a, p, rng, disturb_type)
    return np.stack(out, axis=0)


# --------------------------------------------------------------------------
# counter RNG used by the product's "production mode" (NOT the reference's Threefry stream)
# --------------------------------------------------------------------------

_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox-4x32-10 (Salmon et al., SC'11; Random123).  ctr (...,4) uint32, key (2,) uint32."""
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k0 = np.uint32(key[0])
    k1 = np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c[..., 0].astype(np.uint64) * _PHILOX_M0
            p1 = c[..., 2].astype(np.uint64) * _PHILOX_M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = p1.astype(np.uint32)
            n0 = hi1 ^ c[..., 1] ^ k0
            n1 = lo1
            n2 = hi0 ^ c[..., 3] ^ k1
            n3 = lo0
            c = np.stack([n0, n1, n2, n3], axis=-1)
            k0 = np.uint32(k0 + _PHILOX_W0)
            k1 = np.uint32(k1 + _PHILOX_W1)
    return c


def philox_normals(seed: int, stream: int, n_samples: int, n_cols: int, sample_offset: int = 0) -> np.ndarray:
    """The product's production-mode Gaussian field eps[i, c] (covo_mpc_b200/csrc/rng.cuh):
    counter = (global sample index i, c // 4, stream, 0), key = (seed lo, seed hi); the four
    outputs go through Box-Muller pairs -> four normals for columns 4*(c//4) .. +3."""
    assert n_cols % 4 == 0
    i = (np.arange(n_samples, dtype=np.uint32) + np.uint32(sample_offset))[:, None]
    b = np.arange(n_cols // 4, dtype=np.uint32)[None, :]
    ctr = np.stack(np.broadcast_arrays(i, b, np.uint32(stream), np.uint32(0)), axis=-1).astype(np.uint32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    r = philox4x32_10(ctr, key)
    u = (r.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)  # (0,1)
    r0 = np.sqrt(-2.0 * np.log(u[..., 0]))
    r1 = np.sqrt(-2.0 * np.log(u[..., 2]))
    t0 = 2.0 * np.pi * u[..., 1]
    t1 = 2.0 * np.pi * u[..., 3]
    # the kernel evaluates cos/sin at theta - pi (argument range of the fast intrinsics): a sign flip
    z = -np.stack([r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)], axis=-1)
    return z.reshape(n_samples, n_cols).astype(np.float32)
