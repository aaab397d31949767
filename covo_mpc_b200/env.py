"""Host-side quadrotor environment: the CALLER of the controller hot path.

Mirrors the surface the reference's harness uses around the controllers --
``Quad3D(task, ..., disturb_type, disable_rollover_terminate, generate_noisy_state)``,
``default_params``, ``reset``, ``step`` and ``info["noisy_state"]``
(quadjax/envs/quadrotor.py:23-370, quadjax/envs/base.py:15-50) -- in float32 NumPy for ONE
environment.  It is not on the hot path (SURVEY 8f ranks the device-resident version "next"); it exists
so that episodes can be driven through the drop-in controllers end to end.

Every ``rng`` argument accepts either a ``numpy.random.Generator`` or a JAX PRNGKey (two uint32 words, see
``jaxrng``).  With a key the function consumes it exactly as the reference does -- the same ``split`` tree, the same
``uniform`` / ``normal`` calls in the same order -- so trajectories, initial disturbances and
observation noise are those the reference generates from that key (up to float32 rounding of libm).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Optional, Tuple

import numpy as np

from . import jaxrng as jr

F = np.float32


@dataclass
class EnvParams3D:
    """Hot-path subset of quadjax/dynamics/dataclass.py:40-100 (same names, same defaults)."""

    max_speed: float = 8.0
    max_torque: Tuple[float, float, float] = (9e-3, 9e-3, 2e-3)
    max_omega: Tuple[float, float, float] = (10.0, 10.0, 3.0)
    max_thrust: float = 0.8
    dt: float = 0.02
    g: float = 9.81
    m: float = 0.027
    action_scale: float = 1.0
    alpha_bodyrate: float = 0.5
    max_steps_in_episode: int = 300
    disturb_period: int = 50
    disturb_scale: float = 0.2
    dyn_noise_scale: float = 0.05
    obs_noise_scale: float = 0.05

    def replace(self, **kw):
        return replace(self, **kw)


@dataclass
class EnvState3D:
    """Hot-path subset of quadjax/dynamics/dataclass.py:10-37."""

    pos: np.ndarray
    vel: np.ndarray
    quat: np.ndarray  # (x, y, z, w)
    omega: np.ndarray
    pos_traj: np.ndarray
    vel_traj: np.ndarray
    acc_traj: np.ndarray
    pos_tar: np.ndarray
    vel_tar: np.ndarray
    acc_tar: np.ndarray
    time: int
    f_disturb: np.ndarray
    last_thrust: float = 0.0
    last_torque: np.ndarray = field(default_factory=lambda: np.zeros(3, F))

    def replace(self, **kw):
        return replace(self, **kw)

    def pack_into(self, o: np.ndarray) -> None:
        """Write the C-ABI record into the first 24 floats of ``o`` (no allocation: the per-step call path)."""
        o[0:3] = self.pos
        o[3:7] = self.quat
        o[7:10] = self.vel
        o[10:13] = self.omega
        o[13:16] = self.f_disturb
        o[16:19] = self.pos_tar
        o[19:22] = self.vel_tar
        o[22:24] = 0.0

    def to_state24(self) -> np.ndarray:
        """Pack into the C-ABI record (include/covo_b200.h)."""
        o = np.zeros(24, F)
        o[0:3] = self.pos
        o[3:7] = self.quat
        o[7:10] = self.vel
        o[10:13] = self.omega
        o[13:16] = self.f_disturb
        o[16:19] = self.pos_tar
        o[19:22] = self.vel_tar
        return o


# --------------------------------------------------------------------------------------------------
# reference-trajectory generators (dynamics/utils.py:49-53, 87-130, 183-251)
# --------------------------------------------------------------------------------------------------


def generate_fixed_traj(max_steps: int, dt: float, rng=None):
    z = np.zeros((max_steps, 3), F)
    return z, z.copy(), z.copy()


def generate_lissa_traj(max_steps: int, dt: float, rng):
    if jr.is_key(rng):  # dynamics/utils.py:89-94
        key_amp, key_phase = jr.split(rng, 2)
        amp = jr.uniform(key_amp, (3, 2), -1.0, 1.0).astype(np.float64)
        ph = jr.uniform(key_phase, (3, 2), -np.pi, np.pi).astype(np.float64)
    else:
        amp = rng.uniform(-1.0, 1.0, size=(3, 2))
        ph = rng.uniform(-np.pi, np.pi, size=(3, 2))
    ts = np.arange(0, max_steps + 50) * dt
    w1, w2 = 2 * np.pi * 0.2, 2 * np.pi * 0.4
    pos = np.stack([amp[i, 0] * np.sin(w1 * ts + ph[i, 0]) + amp[i, 1] * np.sin(w2 * ts + ph[i, 1]) for i in range(3)], 1)
    pos = pos - pos[0]
    vel = np.stack([amp[i, 0] * w1 * np.cos(w1 * ts + ph[i, 0]) + amp[i, 1] * w2 * np.cos(w2 * ts + ph[i, 1]) for i in range(3)], 1)
    acc = np.stack([-amp[i, 0] * w1 ** 2 * np.sin(w1 * ts + ph[i, 0]) - amp[i, 1] * w2 ** 2 * np.sin(w2 * ts + ph[i, 1]) for i in range(3)], 1)
    return pos.astype(F), vel.astype(F), acc.astype(F)


def _zigzag_from_key(max_steps: int, dt: float, key):
    """dynamics/utils.py:183-251 with the reference's key plumbing, float32 throughout: ``key_keypoints`` and ``key_angles``
    are the same split of the same key (:187-188); the scan carry starts at keys[1] and is refreshed with keys[i + 1]
    (clamped at the end), so segments 0 and 1 share a key (:238, :241)."""
    point_per_seg = 40
    num_seg = max_steps // point_per_seg + 1
    keys = jr.split(key, num_seg)
    prev = jr.uniform(keys[0], (3,), -1.0, 1.0)
    prev = (prev / _norm(prev) * F(0.1)).astype(F)
    carry = keys[1]
    third = F(np.pi / 3)
    ps, vs = [], []
    for i in range(num_seg):
        to_c = (-prev / _norm(prev)).astype(F)
        dth, dph = jr.uniform(carry, (2,), -third, third)
        theta = np.arccos(to_c[2]) + dth
        phi = np.arctan2(to_c[1], to_c[0]) + dph
        d = np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], F)
        dist = jr.uniform(carry, (), 1.0, 1.5)
        nxt = (prev + dist * d).astype(F)
        ps.append(np.stack([np.linspace(prev[k], nxt[k], point_per_seg, endpoint=False) for k in range(3)], -1).astype(F))
        vs.append(((nxt - prev) / F(point_per_seg + 1) * np.ones((point_per_seg, 3), F) / F(dt)).astype(F))
        carry = keys[min(i + 1, num_seg - 1)]
        prev = nxt
    pos = np.concatenate(ps, 0)
    pos = (pos - pos[0]).astype(F)
    return pos, np.concatenate(vs, 0), np.zeros_like(pos)


def generate_zigzag_traj(max_steps: int, dt: float, rng):
    if jr.is_key(rng):
        return _zigzag_from_key(max_steps, dt, rng)
    point_per_seg = 40
    num_seg = max_steps // point_per_seg + 1
    prev = rng.uniform(-1.0, 1.0, size=3)
    prev = prev / np.linalg.norm(prev) * 0.1
    ps, vs = [], []
    for _ in range(num_seg):
        to_c = -prev / np.linalg.norm(prev)
        dth, dph = rng.uniform(-np.pi / 3, np.pi / 3, size=2)
        theta = np.arccos(to_c[2]) + dth
        phi = np.arctan2(to_c[1], to_c[0]) + dph
        d = np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
        nxt = prev + rng.uniform(1.0, 1.5) * d
        ps.append(np.stack([np.linspace(prev[k], nxt[k], point_per_seg, endpoint=False) for k in range(3)], -1))
        vs.append((nxt - prev) / (point_per_seg + 1) * np.ones((point_per_seg, 3)) / dt)  # sic, utils.py:231-236
        prev = nxt
    pos = np.concatenate(ps, 0)
    pos = pos - pos[0]
    return pos.astype(F), np.concatenate(vs, 0).astype(F), np.zeros_like(pos, dtype=F)


_TASKS = {"tracking": generate_lissa_traj, "tracking_zigzag": generate_zigzag_traj, "hovering": generate_fixed_traj}


def _norm(v):
    return np.sqrt(np.sum(v * v, dtype=F), dtype=F)


def _log_pos(e):
    l = np.log(e + F(1.0))
    c = lambda x: np.minimum(np.maximum(x, F(0)), F(1))
    return e * F(0.4) + c(l * F(4)) * F(0.4) + c(l * F(8)) * F(0.2) + c(l * F(16)) * F(0.1) + c(l * F(32)) * F(0.1)


class Quad3D:
    """Quad3D with the MPC harness's configuration (envs/quadrotor.py:773-781)."""

    def __init__(self, task: str = "tracking", obs_type: str = "quad", enable_randomizer: bool = False,
                 lower_controller: str = "base", disturb_type: str = "none", disable_rollover_terminate: bool = True,
                 generate_noisy_state: bool = True):
        if task not in _TASKS:
            raise NotImplementedError(task)  # envs/quadrotor.py:83-84
        if lower_controller != "base":
            raise NotImplementedError(lower_controller)
        if disturb_type not in ("none", "gaussian"):
            raise NotImplementedError(f"disturb_type {disturb_type!r}")
        if enable_randomizer:
            raise NotImplementedError("domain randomisation is RL-only and out of scope")
        self.task = task
        self.disturb_type = disturb_type
        self.disable_rollover_terminate = disable_rollover_terminate
        self.generate_noisy_state = generate_noisy_state
        self.generate_traj = _TASKS[task]
        self.action_dim = 4

    @property
    def default_params(self) -> EnvParams3D:
        return EnvParams3D()

    def sample_params(self, rng=None) -> EnvParams3D:
        return EnvParams3D()

    # -- reward / termination on a single state (dynamics/utils.py:285-294, quadrotor.py:479-503) ----
    @staticmethod
    def reward_fn(s: EnvState3D, params=None) -> float:
        err_pos = _norm(s.pos_tar - s.pos)
        err_vel = _norm(s.vel_tar - s.vel)
        q = s.quat
        yaw = np.arctan2(F(2) * (q[3] * q[2] + q[0] * q[1]), F(1) - F(2) * (q[1] ** 2 + q[2] ** 2))
        return float(F(1.3) - F(0.05) * err_vel - _log_pos(err_pos) - np.abs(yaw) * F(0.2))

    def is_terminal(self, s: EnvState3D, p: EnvParams3D) -> bool:
        done = (s.time >= p.max_steps_in_episode) or bool(np.any(np.abs(s.pos) > 3.0))
        if not self.disable_rollover_terminate:
            done = done or bool(s.quat[3] < np.cos(np.pi / 4.0)) or bool(np.any(np.abs(s.omega) > 100.0))
        return done

    # -- reset / step ----------------------------------------------------------------------------------
    def reset(self, rng, params: Optional[EnvParams3D] = None):
        p = params or self.default_params
        info_rng = rng
        if jr.is_key(rng):
            # get_zero_state: traj_key, disturb_key, key = split(key, 3) (quadrotor.py:267); reset_env then splits the
            # ORIGINAL key once more for get_info (:368)
            traj_key, disturb_key, _ = jr.split(rng, 3)
            pos_traj, vel_traj, acc_traj = self.generate_traj(p.max_steps_in_episode, p.dt, traj_key)
            fd = jr.uniform(disturb_key, (3,), -p.disturb_scale, p.disturb_scale)
            info_rng = jr.split(rng)[0]
        else:
            pos_traj, vel_traj, acc_traj = self.generate_traj(p.max_steps_in_episode, p.dt, rng)
            fd = rng.uniform(-p.disturb_scale, p.disturb_scale, size=3).astype(F)  # quadrotor.py:300-305
        z = np.zeros(3, F)
        state = EnvState3D(pos=z.copy(), vel=z.copy(), quat=np.array([0, 0, 0, 1], F), omega=z.copy(),
                           pos_traj=pos_traj, vel_traj=vel_traj, acc_traj=acc_traj, pos_tar=pos_traj[0].copy(),
                           vel_tar=vel_traj[0].copy(), acc_tar=acc_traj[0].copy(), time=0, f_disturb=fd)
        info = self.get_info(info_rng, state, state, p)
        return None, info, state  # obs is unused by the MPC controllers (controllers/covo.py:198)

    def get_info(self, rng, state: EnvState3D, next_state: EnvState3D, p: EnvParams3D) -> dict:
        noisy = None
        if self.generate_noisy_state and jr.is_key(rng):  # quadrotor.py:323-351
            k_pos, k_vel, k_quat, k_omega, _ = jr.split(rng, 5)
            sc = F(p.obs_noise_scale)
            noisy = next_state.replace(
                pos=(next_state.pos + jr.normal(k_pos, (3,)) * sc * F(0.25)).astype(F),
                vel=(next_state.vel + jr.normal(k_vel, (3,)) * sc * F(0.5)).astype(F),
                quat=(next_state.quat + jr.normal(k_quat, (4,)) * sc * F(0.02)).astype(F),
                omega=(next_state.omega + jr.normal(k_omega, (3,)) * sc * F(0.5)).astype(F),
            )
        elif self.generate_noisy_state:
            sc = p.obs_noise_scale
            noisy = next_state.replace(
                pos=(next_state.pos + rng.standard_normal(3) * sc * 0.25).astype(F),
                vel=(next_state.vel + rng.standard_normal(3) * sc * 0.5).astype(F),
                quat=(next_state.quat + rng.standard_normal(4) * sc * 0.02).astype(F),
                omega=(next_state.omega + rng.standard_normal(3) * sc * 0.5).astype(F),
            )
        return {"err_pos": float(_norm(state.pos_tar - state.pos)), "err_vel": float(_norm(state.vel_tar - state.vel)),
                "noisy_state": noisy}

    def step_env(self, rng, state: EnvState3D, action, p: EnvParams3D, deterministic: bool = False):
        a = np.clip(np.asarray(action, F), -1.0, 1.0)
        thrust = (a[0] + F(1)) / F(2) * F(p.max_thrust) * F(p.action_scale)
        omega_tar = a[1:] * np.asarray(p.max_omega, F) * F(p.action_scale)
        q = state.quat / _norm(state.quat)
        x, y, z, w = q
        qe3 = np.array([2 * (x * z + y * w), 2 * (y * z - x * w), 1 - 2 * (x * x + y * y)], F)
        om = state.omega
        qdot = F(0.5) * np.array([w * om[0] + (y * om[2] - z * om[1]), w * om[1] + (z * om[0] - x * om[2]),
                                  w * om[2] + (x * om[1] - y * om[0]), -(x * om[0] + y * om[1] + z * om[2])], F)
        vdot = np.array([0, 0, -p.g], F) + F(1.0 / p.m) * (qe3 * thrust + state.f_disturb)
        dt = F(p.dt)
        pos = state.pos + state.vel * dt
        qn = q + qdot * dt
        qn = qn / _norm(qn)
        vel = state.vel + vdot * dt
        omega = F(p.alpha_bodyrate) * om + F(1 - p.alpha_bodyrate) * omega_tar
        info_rng = rng
        if jr.is_key(rng):
            # raw_step: key, step_key = split(key) (quadrotor.py:262); step_fn: key, key_dyn = split(step_key);
            # disturb_key, key = split(key) (dynamics/free.py:136,144); step_env: info_key, key = split(key) (:245)
            disturb_key = jr.split(jr.split(jr.split(rng)[1])[0])[0]
            info_rng = jr.split(rng)[0]
        if self.disturb_type == "gaussian" and not deterministic:
            z3 = jr.normal(disturb_key, (3,)) if jr.is_key(rng) else rng.standard_normal(3)
            fd = (F(p.dyn_noise_scale) * z3).astype(F)
        else:
            fd = np.zeros(3, F)
        time = state.time + 1
        ti = min(time, state.pos_traj.shape[0] - 1)  # clamped gather, dynamics/free.py:153-155
        nxt = state.replace(pos=pos.astype(F), vel=vel.astype(F), quat=qn.astype(F), omega=omega.astype(F), time=time,
                            f_disturb=fd, pos_tar=state.pos_traj[ti].copy(), vel_tar=state.vel_traj[ti].copy(),
                            acc_tar=state.acc_traj[ti].copy(), last_thrust=float(thrust))
        reward = self.reward_fn(state)  # PRE-step state, quadrotor.py:243
        done = self.is_terminal(state, p)
        info = self.get_info(info_rng, state, nxt, p)
        return None, nxt, reward, done, info

    def step(self, rng, state: EnvState3D, action, params: Optional[EnvParams3D] = None):
        """BaseEnvironment.step (envs/base.py:15-40) including the auto-reset on ``done``."""
        p = params or self.default_params
        key_reset = rng
        if jr.is_key(rng):
            rng, key_reset = jr.split(rng)  # base.py:27
        obs, nxt, reward, done, info = self.step_env(rng, state, action, p)
        if done:
            obs, info, nxt = self.reset(key_reset, p)
        return obs, nxt, reward, done, info
