"""Host-side mirror of the reference's controller plugin surface.

    controller, control_params = get_controller(env, "covo-online", "N8192_H50_lam0.01")
    control_params = controller.reset(env_state, env_params, controller.init_control_params, key)
    action, control_params, info = controller(obs, env_state, env_params, rng_act, control_params, env_info)

is the reference's contract (quadjax/controllers/base.py:5-19; dispatch quadjax/envs/quadrotor.py:670-752;
call sites :523-525, :548-550, :609-611, :618-620) and works unchanged here for the names
"mppi", "covo-online" / "covo_online", "covo-offline" / "covo_offline".  Everything numeric happens in
libcovo_b200.so (hand-written sm_100a CUDA); this module only moves a 24-float state record in and a
4-float action out per step.  There is no CPU fallback.

``rng_act`` (a JAX PRNG key in the reference):
  * ``None`` -> production mode, Gaussian draws from the in-kernel Philox counter RNG;
  * a 2-word uint32 key (``jaxrng.PRNGKey`` / ``jaxrng.split``) -> the kernel draws what ``jax.random`` would draw from
    that key (Threefry-2x32-20, legacy layout): same split / normal sequence as the reference;
  * a float array of standard normals, shape (N, 4H) (CoVO) or (N, H, 4) (MPPI) -> "parity mode": the
    caller supplies exactly what ``jax.random.normal`` would have drawn (the JAX Threefry stream itself is
    un-pinned third-party arithmetic);
  * a ``numpy.random.Generator`` -> the draws are taken from it on the host.

``control_params`` keeps the reference's functional style: the call returns a new params object.  The
mean / covariance stay resident in HBM; the returned object materialises them lazily on attribute access.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from . import _lib, jaxrng
from .env import EnvParams3D, EnvState3D, Quad3D


class _DeviceArray:
    """Lazy host view of a controller-resident array, valid for one controller generation."""

    def __init__(self, owner: "_SamplingController", generation: int, what: str):
        self._owner, self._gen, self._what, self._cache = owner, generation, what, None

    def _get(self) -> np.ndarray:
        if self._cache is None:
            if self._owner._generation != self._gen:
                raise RuntimeError("stale control_params: the controller state has advanced past this object")
            self._cache = self._owner._download(self._what)
        return self._cache

    def __array__(self, dtype=None, copy=None):
        a = self._get()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, i):
        return self._get()[i]

    @property
    def shape(self):
        return self._get().shape


@dataclass
class MPPIParams:
    """quadjax/controllers/mppi.py:11-19"""

    gamma_mean: float
    gamma_sigma: float
    discount: float
    sample_sigma: float
    a_mean: Any
    a_cov: Any
    _gen: int = field(default=-1, repr=False)

    def replace(self, **kw):
        d = dict(gamma_mean=self.gamma_mean, gamma_sigma=self.gamma_sigma, discount=self.discount,
                 sample_sigma=self.sample_sigma, a_mean=self.a_mean, a_cov=self.a_cov)
        d.update(kw)
        return MPPIParams(**d)


@dataclass
class CoVOParams:
    """quadjax/controllers/covo.py:13-22"""

    gamma_mean: float
    gamma_sigma: float
    discount: float
    sample_sigma: float
    a_mean: Any
    a_cov: Any
    a_cov_offline: Any
    _gen: int = field(default=-1, repr=False)

    def replace(self, **kw):
        d = dict(gamma_mean=self.gamma_mean, gamma_sigma=self.gamma_sigma, discount=self.discount,
                 sample_sigma=self.sample_sigma, a_mean=self.a_mean, a_cov=self.a_cov, a_cov_offline=self.a_cov_offline)
        d.update(kw)
        return CoVOParams(**d)


class BaseController:
    """quadjax/controllers/base.py:5-19"""

    def __init__(self, env, control_params) -> None:
        self.env = env
        self.init_control_params = control_params

    def update_params(self, env_params, control_params):
        return control_params

    def reset(self, env_state=None, env_params=None, control_params=None, key=None):
        return self.init_control_params

    def __call__(self, obs, state, env_params, rng_act, control_params, env_info=None):
        raise NotImplementedError


class _SamplingController(BaseController):
    _mode = -1

    def __init__(self, env, control_params, N: int, H: int, lam: float, *, device: int = 0, seed: int = 0,
                 rank: int = 0, world: int = 1) -> None:
        super().__init__(env, control_params)
        self.N, self.H, self.lam = int(N), int(H), float(lam)
        self.action_dim = getattr(env, "action_dim", 4)
        assert self.action_dim == 4, "only support 4D action space Quadrotor environment for now"  # covo.py:45-47
        p: EnvParams3D = env.default_params
        cfg = _lib.default_config()
        cfg.mode = self._mode
        cfg.n_samples, cfg.horizon, cfg.n_env = self.N, self.H, 1
        cfg.device, cfg.rank, cfg.world = device, rank, world
        cfg.lam = self.lam
        cfg.sample_sigma = float(control_params.sample_sigma)
        cfg.gamma_mean = float(control_params.gamma_mean)
        cfg.gamma_sigma = float(control_params.gamma_sigma)
        cfg.discount = float(control_params.discount)
        cfg.m, cfg.g, cfg.max_thrust, cfg.dt = p.m, p.g, p.max_thrust, p.dt
        cfg.alpha_bodyrate, cfg.action_scale, cfg.pos_limit = p.alpha_bodyrate, p.action_scale, 3.0
        for k in range(3):
            cfg.max_omega[k] = p.max_omega[k]
        cfg.max_steps_in_episode = p.max_steps_in_episode
        cfg.seed = seed
        self._cfg = cfg
        self._handle: Optional[_lib.Handle] = None
        self._traj_id = None
        self._generation = 0
        self._env_params_seen = None
        self.want_info = False  # pos_mean / pos_std (covo.py:281) are computed only on request
        self._noise_rng = None
        self._fdist_set = False

    # -- plumbing ---------------------------------------------------------------------------------------
    def _sync_control_consts(self, control_params):
        """gamma_mean / gamma_sigma / discount / sample_sigma are fields of control_params in the reference and read at CALL time
        (covo.py:270-278, mppi.py:100-125): a caller that passes cp.replace(gamma_sigma=0.3) gets a handle configured that way."""
        vals = tuple(float(getattr(control_params, f)) for f in ("gamma_mean", "gamma_sigma", "discount", "sample_sigma"))
        cur = (float(self._cfg.gamma_mean), float(self._cfg.gamma_sigma), float(self._cfg.discount), float(self._cfg.sample_sigma))
        if any(abs(a - b) > 1e-7 * max(1.0, abs(b)) for a, b in zip(vals, cur)):
            self._cfg.gamma_mean, self._cfg.gamma_sigma, self._cfg.discount, self._cfg.sample_sigma = vals
            self._cfg_dirty = True

    def _ensure_handle(self, traj_len: int) -> _lib.Handle:
        if self._handle is None or self._cfg.traj_len != traj_len or getattr(self, "_cfg_dirty", False):
            self._cfg_dirty = False
            carried = None
            if self._handle is not None:
                # a reference trajectory of another length: the resident controller state moves to the new handle, so that
                # the params object returned by the last call stays valid (its _gen still matches)
                carried = (self._handle.get_mean(), self._handle.get_cov() if self._mode == _lib.MODE_MPPI else None)
                self._handle.close()
            self._cfg.traj_len = traj_len
            self._handle = _lib.Handle(self._cfg)
            self._traj_id = None
            self._env_params_seen = None
            if carried is not None:
                self._handle.set_mean(carried[0])
                if carried[1] is not None:
                    self._handle.set_cov(carried[1])
            self._on_new_handle()
        return self._handle

    _MODEL_FIELDS = ("m", "g", "max_thrust", "dt", "alpha_bodyrate", "action_scale", "max_omega", "max_steps_in_episode")

    def _sync_env_params(self, env_params):
        """The reference plans with the env_params of the CALL (controllers/covo.py:187-283, mppi.py:28-134), e.g. a mass drawn by
        sample_params: forward the model constants to the handle whenever the caller passes a different params object."""
        if env_params is None or env_params is self._env_params_seen:
            return
        self._env_params_seen = env_params
        vals = tuple(getattr(env_params, f, getattr(self.env.default_params, f)) for f in self._MODEL_FIELDS)
        key = tuple(tuple(float(x) for x in v) if isinstance(v, (tuple, list, np.ndarray)) else float(v) for v in vals)
        if key != getattr(self, "_model_key", None):
            m, g, mt, dt, al, sc, mo, ms = vals
            self._handle.set_env_params(float(m), float(g), float(mt), float(dt), float(al), float(sc), mo, int(ms))
            self._cfg.m, self._cfg.g, self._cfg.max_thrust, self._cfg.dt = float(m), float(g), float(mt), float(dt)
            self._cfg.alpha_bodyrate, self._cfg.action_scale, self._cfg.max_steps_in_episode = float(al), float(sc), int(ms)
            for k in range(3):
                self._cfg.max_omega[k] = float(mo[k])
            self._model_key = key

    def _on_new_handle(self):
        pass

    def _sync_reference(self, state: EnvState3D):
        h = self._ensure_handle(int(state.pos_traj.shape[0]))
        tid = (id(state.pos_traj), id(state.vel_traj))
        if tid != self._traj_id:  # trajectories change only at reset
            h.set_reference(state.pos_traj, state.vel_traj, state.acc_traj)
            self._traj_id = tid
            self._ref_keepalive = (state.pos_traj, state.vel_traj)
        return h

    def _download(self, what: str) -> np.ndarray:
        h = self._handle
        if what == "a_mean":
            return h.get_mean()[0]
        if what == "a_cov":
            return h.get_cov()[0]
        raise KeyError(what)

    def _upload_params(self, control_params):
        if getattr(control_params, "_gen", -1) == self._generation:
            return  # the object we returned last time: the state is already resident
        self._handle.set_mean(np.asarray(control_params.a_mean, dtype=np.float32).reshape(1, self.H, 4))
        self._upload_cov(control_params)

    def _upload_cov(self, control_params):
        pass

    def _eps(self, rng_act, shape):
        if rng_act is None:
            return None
        if isinstance(rng_act, np.random.Generator):
            return rng_act.standard_normal(shape).astype(np.float32)
        a = np.asarray(rng_act)
        if a.dtype.kind == "u" and a.size == 2:
            # a JAX PRNGKey: draw what the reference draws from it -- rng_act, act_key = split(rng_act); act_keys =
            # split(act_key, N); normal(act_keys[i], ...) (controllers/covo.py:212-217, mppi.py:53-60) -- in the kernel
            self._handle.set_jax_key(jaxrng.split(a.astype(np.uint32).reshape(2))[1])
            return None
        if a.size != int(np.prod(shape)):
            raise ValueError(f"explicit normal draws must have shape {shape}")
        return a.astype(np.float32).reshape(shape)

    def _finish(self, control_params, action):
        self._generation += 1
        gen = self._generation
        # control_params.replace(a_mean=..., a_cov=...) without re-running the dataclass constructor (per-step path)
        new = object.__new__(type(control_params))
        new.__dict__.update(control_params.__dict__)
        new.a_mean, new.a_cov, new._gen = _DeviceArray(self, gen, "a_mean"), _DeviceArray(self, gen, "a_cov"), gen
        info = None
        if self.want_info:
            m, s = self._handle.pos_stats()
            info = {"pos_mean": m[0], "pos_std": s[0]}
        return action, new, info

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None


class MPPIController(_SamplingController):
    """quadjax/controllers/mppi.py:21-134"""

    _mode = _lib.MODE_MPPI

    def _upload_cov(self, control_params):
        self._handle.set_cov(np.asarray(control_params.a_cov, dtype=np.float32).reshape(1, self.H, 4, 4))

    def __call__(self, obs, env_state, env_params, rng_act, control_params, info=None):
        state: EnvState3D = info["noisy_state"]  # mppi.py:40
        self._sync_control_consts(control_params)
        h = self._sync_reference(state)
        self._sync_env_params(env_params)
        self._upload_params(control_params)
        if self.want_info:
            h.enable_pos_stats(True)
        if getattr(self.env, "disturb_type", "none") == "gaussian":
            # mppi.py:74 rolls out with a STOCHASTIC step_env, and every sample and horizon step gets the same step_key: one force
            # dyn_noise_scale * N(0, I)^3 for the whole call.  With a JAX key: rng_act, act_key = split(rng_act) (:53);
            # rng_act, step_key = split(rng_act) (:69); inside step_env the disturbance key is three splits down (quadrotor.py:262,
            # dynamics/free.py:136-144).  Without one: the controller's own generator.
            scale = np.float32((env_params or self.env.default_params).dyn_noise_scale)
            if jaxrng.is_key(rng_act):
                step_key = jaxrng.split(jaxrng.split(np.asarray(rng_act, dtype=np.uint32))[0])[1]
                z = jaxrng.normal(jaxrng.split(jaxrng.split(jaxrng.split(step_key)[1])[0])[0], (3,))
            else:
                if self._noise_rng is None:
                    self._noise_rng = np.random.default_rng(int(self._cfg.seed) + 7919)
                z = self._noise_rng.standard_normal(3)
            h.set_rollout_disturbance(np.tile((scale * np.asarray(z, np.float32))[None, None, :], (1, self.H, 1)))
            self._fdist_set = True
        elif self._fdist_set:
            h.set_rollout_disturbance(None)
            self._fdist_set = False
        eps = self._eps(rng_act, (1, h.n_local, self.H * 4))
        action = h.step_state(state) if eps is None else h.step(state.to_state24(), [state.time], eps)[0]
        return self._finish(control_params, action)


def offline_disturbance_normals(key, T: int) -> np.ndarray:
    """The [T][3] standard normals reset_a_cov_offline (controllers/covo.py:77-99) draws for the disturbance of its T state advances,
    from the reference's key schedule: per schedule step `rng_step, key = split(key)` (expansion controller, unused by the PID policy),
    `rng_step, key = split(key)` (step_env); inside step_env the disturbance key is split(split(split(rng_step)[1])[0])[0]
    (envs/quadrotor.py:262, dynamics/free.py:136-144) -- the chain env.Quad3D.step_env follows for the same key."""
    k = np.asarray(key, dtype=np.uint32)
    z = np.empty((T, 3), np.float32)
    for t in range(T):
        k = jaxrng.split(k)[1]
        rs, k = jaxrng.split(k)
        z[t] = jaxrng.normal(jaxrng.split(jaxrng.split(jaxrng.split(rs)[1])[0])[0], (3,))
    return z


class CoVOController(_SamplingController):
    """quadjax/controllers/covo.py:25-283"""

    def __init__(self, env, control_params, N: int, H: int, lam: float, mode: str = "online", **kw) -> None:
        if mode == "online":
            self._mode = _lib.MODE_COVO_ONLINE
        elif mode == "offline":
            self._mode = _lib.MODE_COVO_OFFLINE
        else:
            raise NotImplementedError(mode)  # covo.py:113-114
        self.mode = mode
        super().__init__(env, control_params, N, H, lam, **kw)
        self._table = None

    def _on_new_handle(self):
        if self._mode == _lib.MODE_COVO_OFFLINE and self._table is not None:
            self._handle.set_cov_offline(self._table)

    def reset(self, env_state=None, env_params=None, control_params=None, key=None):
        if self._mode != _lib.MODE_COVO_OFFLINE:
            return self.init_control_params
        # reset_a_cov_offline, covo.py:101-104: rebuild the covariance schedule from this state (on device)
        disturb_type = getattr(self.env, "disturb_type", "none")
        if disturb_type not in ("none", "gaussian"):
            raise NotImplementedError(f"covo-offline schedule under disturb_type {disturb_type!r}")
        h = self._sync_reference(env_state)
        self._sync_env_params(env_params)
        T = int(self.env.default_params.max_steps_in_episode)
        f_disturb = None
        if disturb_type == "gaussian":
            # The state advance between schedule entries is stochastic (covo.py:80-89).  Key schedule of get_single_a_cov_offline:
            # `rng_step, key = split(key)` for the expansion controller, `rng_step, key = split(key)` for step_env (:81-88); inside
            # step_env the disturbance key is split(split(split(rng_step)[1])[0])[0] (quadrotor.py:262, dynamics/free.py:136-144).
            scale = float((env_params or self.env.default_params).dyn_noise_scale)
            z = np.empty((T, 3), np.float32)
            if jaxrng.is_key(key):
                z[:] = offline_disturbance_normals(key, T)
            else:
                gen = key if isinstance(key, np.random.Generator) else np.random.default_rng(int(self._cfg.seed))
                z[:] = gen.standard_normal((T, 3))
            f_disturb = scale * z
        h.reset_offline(env_state.to_state24(), [env_state.time], T, f_disturb)
        self._table = None
        # The reference keeps a_mean / a_cov across this reset (covo.py:101-104 replaces a_cov_offline only) and render_env calls it
        # with the CURRENT params after `done` (envs/quadrotor.py:637-639): the resident mean is untouched by the schedule build, so
        # the returned object stays in the generation of the one passed in.
        cp = control_params if control_params is not None else self.init_control_params
        new = cp.replace(a_cov_offline=_OfflineTable(self, T))
        new._gen = getattr(cp, "_gen", -1)
        return new

    def _upload_cov(self, control_params):
        if self._mode == _lib.MODE_COVO_OFFLINE:
            tab = control_params.a_cov_offline
            if isinstance(tab, _OfflineTable) and tab.owner is self:
                return
            tab = np.asarray(tab, dtype=np.float32)
            if tab.ndim != 3 or tab.shape[1] != 4 * self.H:
                raise ValueError("a_cov_offline must be (T, 4H, 4H); call controller.reset first")
            self._table = tab
            self._handle.set_cov_offline(tab)

    def __call__(self, obs, env_state, env_params, rng_act, control_params, info=None):
        state: EnvState3D = info["noisy_state"]  # covo.py:198
        self._sync_control_consts(control_params)
        h = self._sync_reference(state)
        self._sync_env_params(env_params)
        self._upload_params(control_params)
        if self.want_info:
            h.enable_pos_stats(True)
        eps = self._eps(rng_act, (1, h.n_local, self.H * 4))
        action = h.step_state(state) if eps is None else h.step(state.to_state24(), [state.time], eps)[0]
        return self._finish(control_params, action)


class _OfflineTable:
    """Handle to the device-resident a_cov_offline schedule (materialised on np.asarray)."""

    def __init__(self, owner: CoVOController, T: int):
        self.owner, self.T = owner, T

    def __array__(self, dtype=None, copy=None):
        a = self.owner._handle.get_cov_offline(self.T)
        return a if dtype is None else a.astype(dtype)

    @property
    def shape(self):
        return (self.T, 4 * self.owner.H, 4 * self.owner.H)



@dataclass
class PIDParams:
    """quadjax/controllers/pid.py:11-22 (same names, same defaults)."""

    Kp: float = 4.0
    Kd: float = 4.0
    Ki: float = 1.0
    Kp_att: float = 4.0
    Ki_att: float = 1.0
    integral: Any = field(default_factory=lambda: np.zeros(3, np.float32))
    quat_desired: Any = field(default_factory=lambda: np.array([0.0, 0.0, 0.0, 1.0], np.float32))
    att_integral: Any = field(default_factory=lambda: np.zeros(3, np.float32))

    def replace(self, **kw):
        import dataclasses

        return dataclasses.replace(self, **kw)


class PIDController(BaseController):
    """quadjax/controllers/pid.py:24-83 -- ``--controller pid`` and the expansion policy of CoVO-offline.  The action is
    computed by the device PID kernel (the same ``pid_action`` the offline schedule uses); the integral bookkeeping
    (:77-81) stays on the host.  ``quat_desired`` is not on any caller's path and is left unchanged."""

    def __init__(self, env, control_params, *, device: int = 0) -> None:
        super().__init__(env, control_params)
        self.param = env.default_params
        self._device = device
        self._handle: Optional[_lib.Handle] = None
        self._traj_id = None

    def _sync(self, state: EnvState3D) -> _lib.Handle:
        T = int(state.pos_traj.shape[0])
        if self._handle is None or self._handle.cfg.traj_len != T:
            if self._handle is not None:
                self._handle.close()
            p = self.param
            cfg = _lib.default_config()
            cfg.mode, cfg.n_samples, cfg.horizon, cfg.n_env, cfg.traj_len, cfg.device = _lib.MODE_MPPI, 64, 2, 1, T, self._device
            cfg.m, cfg.g, cfg.max_thrust, cfg.dt = p.m, p.g, p.max_thrust, p.dt
            for k in range(3):
                cfg.max_omega[k] = p.max_omega[k]
            self._handle = _lib.Handle(cfg)
            self._traj_id = None
        tid = (id(state.pos_traj), id(state.acc_traj))
        if tid != self._traj_id:
            self._handle.set_reference(state.pos_traj, state.vel_traj, state.acc_traj)
            self._traj_id, self._keep = tid, (state.pos_traj, state.acc_traj)
        return self._handle

    def __call__(self, obs, state, env_param, rng_act, control_params, info=None):
        h = self._sync(state)
        integral = np.asarray(control_params.integral, np.float32)
        action = h.pid_action(state.to_state24(), [state.time], control_params.Kp, control_params.Kd, control_params.Ki,
                              control_params.Kp_att, integral)[0]
        dt = np.float32(getattr(env_param, "dt", self.param.dt))
        new = control_params.replace(integral=(integral + (np.asarray(state.pos, np.float32) - np.asarray(state.pos_tar, np.float32)) * dt))
        return action, new, None

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None


class RandomController(BaseController):
    """quadjax/controllers/random.py:8-16: 0.3 * normal(rng_act, (4,)); rng_act a JAX key or a NumPy generator."""

    def __call__(self, obs, state, env_params, rng_act, control_params, env_info=None):
        if jaxrng.is_key(rng_act):
            z = jaxrng.normal(rng_act, (4,))
        else:
            z = (rng_act if rng_act is not None else np.random.default_rng()).standard_normal(4).astype(np.float32)
        return z * np.float32(0.3), control_params, None


def get_controller(env, controller_name: str, controller_params: Optional[str] = None, debug: bool = False, **kw):
    """quadjax/envs/quadrotor.py:670-752 -- same names, same parameter string, same defaults."""

    def parse_sample_params(param_text):
        if not param_text:
            return 8192, 32, 0.01, 0.5
        parts = param_text.split("_")
        return int(parts[0][1:]), int(parts[1][1:]), float(parts[2][3:]), 0.5

    p = env.default_params

    def get_sample_mean(H):
        th = (p.m * p.g / p.max_thrust) * 2.0 - 1.0
        return np.tile(np.array([th, 0.0, 0.0, 0.0], np.float32), (H, 1))

    if controller_name == "mppi":
        N, H, lam, sigma = parse_sample_params(controller_params)
        if debug:
            N, H = 4, 2
        a_cov = np.tile(np.diag(np.full(4, sigma ** 2, np.float32)), (H, 1, 1))
        control_params = MPPIParams(gamma_mean=1.0, gamma_sigma=0.0, discount=1.0, sample_sigma=sigma,
                                    a_mean=get_sample_mean(H), a_cov=a_cov)
        controller = MPPIController(env=env, control_params=control_params, N=N, H=H, lam=lam, **kw)
    elif "covo" in controller_name:
        N, H, lam, sigma = parse_sample_params(controller_params)
        if debug:
            N, H = 4, 2
        mode = "online" if "online" in controller_name else ("offline" if "offline" in controller_name else "online")  # :731-737
        control_params = CoVOParams(gamma_mean=1.0, gamma_sigma=0.0, discount=1.0, sample_sigma=sigma,
                                    a_mean=get_sample_mean(H), a_cov=np.diag(np.full(H * 4, sigma ** 2, np.float32)),
                                    a_cov_offline=np.zeros((H, 4, 4), np.float32))
        controller = CoVOController(env=env, control_params=control_params, N=N, H=H, lam=lam, mode=mode, **kw)
    elif controller_name == "pid":  # quadrotor.py:692-699
        control_params = PIDParams(Kp=10.0, Kd=5.0, Ki=0.0, Kp_att=10.0)
        controller = PIDController(env, control_params=control_params, **kw)
    elif controller_name == "random":  # quadrotor.py:700-702
        control_params = None
        controller = RandomController(env, control_params)
    else:
        raise NotImplementedError(controller_name)  # quadrotor.py:750-751
    return controller, control_params
