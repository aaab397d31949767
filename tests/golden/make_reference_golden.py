"""Golden vectors produced by the REFERENCE'S OWN PYTHON SOURCE (/root/reference/quadjax), executed in this container
with a NumPy shim standing in for `jax.numpy` (JAX itself is not installable here: no wheel, no network).

What the shim is: `jax.numpy` -> NumPy with float32 array creation (JAX's default precision), `jax.jit` -> identity,
`lax.scan` / `lax.select` / `lax.cond` -> their Python meaning, `flax.struct.dataclass` -> a frozen-style dataclass with
`.replace`, `jax.random.*` -> a logged NumPy generator (the draws are stored next to the outputs so that the oracle and
the kernels can be fed the same numbers), `jax.vmap` -> a Python loop over the leading axis of every pytree leaf,
`jax.random.multivariate_normal` -> mean + cholesky(cov) @ normal (JAX's default method) with the normals logged,
`jax.jacfwd(jax.jacfwd(f))` -> Richardson-extrapolated central second differences of the reference's own `f` evaluated
in float64 (the result is rounded to float32, JAX's output precision), empty stand-ins for chex / gymnax.  What it is
NOT: XLA arithmetic, JAX's Threefry streams, JAX's forward-mode AD.  So the vectors pin the maths of the hot path as the
reference wrote it -- per step and for a whole controller call (sections 6-8: the
Hessian of the reference's own cost function, optimize_sigma, sampling, rollouts with reward freeze, softmax update)
-- but not the bits of the random streams or of XLA's float32 rounding.

Functions executed from the reference, unmodified:
  Quad3D.__init__ / step_env / raw_step / is_terminal / get_info   envs/quadrotor.py:29-200, 215-263, 314-361, 479-503
  free_dynamics_3d_bodyrate, quad_dynamics_bodyrate                dynamics/free.py:74-202
  geom.L, geom.H, geom.qtoQ, geom.hat, ...                          dynamics/geom.py
  tracking_penyaw_reward_fn, log_pos_fn                             dynamics/utils.py:266-294
  CoVOController.optimize_sigma                                     controllers/covo.py:116-132
  PIDController.__call__                                            controllers/pid.py:38-83
  get_controller (defaults, hover warm start)                       envs/quadrotor.py:670-752

Run from the repo root (needs /root/reference):   python tests/golden/make_reference_golden.py
Writes tests/golden/reference_*.npz."""
import dataclasses
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
F = np.float32


# ------------------------------------------------------------------------------------------------------------------
# the shim
# ------------------------------------------------------------------------------------------------------------------
CREATE = [F]  # dtype of newly created floating arrays: float32 (JAX's default); float64 while differencing the cost


def _f32(a):
    a = np.asarray(a)
    return a.astype(CREATE[0]) if a.dtype == np.float64 else a


# pytrees: dataclasses (flax.struct), tuples, lists, dicts; everything else is a leaf
def _is_dc(x):
    return dataclasses.is_dataclass(x) and not isinstance(x, type)


def tree_map(f, t, *rest):
    if _is_dc(t):
        return type(t)(**{fl.name: tree_map(f, getattr(t, fl.name), *[getattr(r, fl.name) for r in rest]) for fl in dataclasses.fields(t)})
    if isinstance(t, (tuple, list)):
        return type(t)(tree_map(f, x, *[r[i] for r in rest]) for i, x in enumerate(t))
    if isinstance(t, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in t.items()}
    if t is None:
        return None
    return f(t, *rest)


def tree_leaves(t):
    out = []
    tree_map(lambda x: out.append(x), t)
    return out


def tree_stack(items):
    return tree_map(lambda *xs: np.stack([np.asarray(x) for x in xs]), items[0], *items[1:])


def vmap(f, in_axes=0, out_axes=0):
    def g(*args):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(len(tree_leaves(a)[0]) for a, ax in zip(args, axes) if ax == 0)
        return tree_stack([f(*[tree_map(lambda x: x[i], a) if ax == 0 else a for a, ax in zip(args, axes)]) for i in range(n)])

    return g


class ClampArr(np.ndarray):
    """JAX clamps out-of-range integer indices (dynamics/utils.py:238 relies on it); NumPy raises."""

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            idx = min(max(int(idx), -self.shape[0]), self.shape[0] - 1)
        return np.asarray(np.ndarray.__getitem__(self, idx))

    def __iter__(self):
        return iter(np.asarray(self))


class _Jac:
    def __init__(self, f):
        self.f = f


def fd_hessian(f, x, *rest, h=1e-3):
    """Hessian of the scalar f at x by central second differences in float64, steps h and 2h, Richardson-extrapolated."""
    CREATE[0] = np.float64
    try:
        x = np.asarray(x, dtype=np.float64)
        n = x.size

        def ev(d):
            return float(f(x + d, *rest))

        def one(hh):
            Hm = np.zeros((n, n))
            f0 = ev(np.zeros(n))
            E = np.eye(n) * hh
            fp = [ev(E[i]) for i in range(n)]
            fm = [ev(-E[i]) for i in range(n)]
            for i in range(n):
                Hm[i, i] = (fp[i] - 2 * f0 + fm[i]) / hh ** 2
                for j in range(i):
                    v = (ev(E[i] + E[j]) - ev(E[i] - E[j]) - ev(E[j] - E[i]) + ev(-E[i] - E[j])) / (4 * hh ** 2)
                    Hm[i, j] = Hm[j, i] = v
            return Hm

        return ((4.0 * one(h) - one(2 * h)) / 3.0).astype(F)
    finally:
        CREATE[0] = F


def jacfwd(f, argnums=0):
    if isinstance(f, _Jac):
        return lambda x, *rest: fd_hessian(f.f, x, *rest)
    return _Jac(f)


class JArr(np.ndarray):
    """ndarray with JAX's functional update syntax x.at[idx].set(v)."""

    @property
    def at(self):
        outer = self

        class _At:
            def __getitem__(self, idx):
                class _Ref:
                    def set(self, v):
                        out = np.array(outer)
                        out[idx] = v
                        return out.view(JArr)

                return _Ref()

        return _At()


class _JNP(types.ModuleType):
    """NumPy with JAX's default dtypes for array creation."""

    def __getattr__(self, name):
        return getattr(np, name)


def make_jnp():
    m = _JNP("jax.numpy")
    m.array = lambda x, dtype=None: _f32(np.array(x, dtype=dtype))
    m.asarray = lambda x, dtype=None: _f32(np.asarray(x, dtype=dtype))
    m.zeros = lambda shape, dtype=F: np.zeros(shape, dtype=dtype).view(JArr)
    m.ones = lambda shape, dtype=F: np.ones(shape, dtype=dtype)
    m.eye = lambda n, dtype=F: np.eye(n, dtype=dtype)
    m.full = lambda shape, v, dtype=None: np.full(shape, v, dtype=(bool if isinstance(v, (bool, np.bool_)) else CREATE[0]) if dtype is None else dtype)
    m.linspace = lambda *a, **k: _f32(np.linspace(*a, **k))
    m.zeros_like = np.zeros_like
    m.ndarray = np.ndarray
    m.float32 = np.float32
    m.pi = np.pi
    m.newaxis = None
    m.linalg = np.linalg
    return m


class RandomLog:
    """jax.random stand-in: keys are opaque, every draw is logged in call order."""

    def __init__(self, seed=0):
        self.rng = np.random.default_rng(seed)
        self.normals, self.uniforms, self.mvn = [], [], []

    def PRNGKey(self, seed):
        return np.array([0, seed], dtype=np.uint32)

    def split(self, key, num=2):
        k = int(np.asarray(key).ravel()[-1])
        return np.array([[(k + 1 + i) & 0xFFFFFFFF, (k * 7 + i) & 0xFFFFFFFF] for i in range(num)], dtype=np.uint32)

    def multivariate_normal(self, key, mean, cov):
        # jax.random.multivariate_normal, method='cholesky' (the default): mean + L z
        mean, cov = np.asarray(mean, F), np.asarray(cov, F)
        z = self.rng.standard_normal(mean.shape[-1]).astype(F)
        self.mvn.append(z)
        return (mean + np.linalg.cholesky(cov) @ z).astype(F)

    def normal(self, key, shape=()):
        z = self.rng.standard_normal(shape).astype(F)
        self.normals.append(np.atleast_1d(z).ravel())
        return z

    def uniform(self, key, shape=(), minval=0.0, maxval=1.0, dtype=F):
        u = (self.rng.uniform(size=shape) * (maxval - minval) + minval).astype(F)
        self.uniforms.append(np.atleast_1d(u).ravel())
        return u


def install_shim(rand: RandomLog):
    jnp = make_jnp()
    jax = types.ModuleType("jax")
    jax.numpy = jnp

    def jit(f=None, **kw):
        if f is None:
            return lambda g: g
        return f

    jax.jit = jit
    lax = types.ModuleType("jax.lax")

    def scan(f, init, xs, length=None):
        carry, ys = init, []
        for x in (xs if xs is not None else [None] * length):
            carry, y = f(carry, x)
            ys.append(y)
        return carry, (tree_stack(ys) if ys and ys[0] is not None else None)

    lax.scan = scan
    lax.select = lambda c, a, b: a if bool(c) else b
    lax.cond = lambda c, t, f, *ops: t(*ops) if bool(c) else f(*ops)
    jax.lax = lax
    rnd = types.ModuleType("jax.random")
    for k in ("PRNGKey", "split", "normal", "uniform", "multivariate_normal"):
        setattr(rnd, k, getattr(rand, k))
    jax.random = rnd
    jax.tree_map = tree_map
    jax.vmap = vmap
    jax.jacfwd = jacfwd
    lax.stop_gradient = lambda x: x
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.debug = types.SimpleNamespace(print=lambda *a, **k: None)
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.random": rnd})

    chex = types.ModuleType("chex")
    chex.PRNGKey = chex.Array = chex.ArrayTree = object
    sys.modules["chex"] = chex

    flax = types.ModuleType("flax")
    struct = types.ModuleType("flax.struct")

    def dataclass(cls):
        cls = dataclasses.dataclass(cls)
        cls.replace = lambda self, **kw: dataclasses.replace(self, **kw)
        return cls

    struct.dataclass = dataclass
    struct.field = lambda **kw: dataclasses.field(**{k: v for k, v in kw.items() if k in ("default", "default_factory")})
    flax.struct = struct
    sys.modules.update({"flax": flax, "flax.struct": struct})

    gymnax = types.ModuleType("gymnax")
    envs = types.ModuleType("gymnax.environments")
    envmod = types.ModuleType("gymnax.environments.environment")

    class Environment:  # gymnax.environments.environment.Environment: only what the hot path touches
        def __init__(self, *a, **k):
            pass

        def discount(self, state, params):  # gymnax: select(is_terminal, 0.0, 1.0)
            return 0.0 if bool(self.is_terminal(state, params)) else 1.0

    envmod.Environment, envmod.EnvParams, envmod.EnvState = Environment, object, object
    wr = types.ModuleType("gymnax.wrappers")
    purerl = types.ModuleType("gymnax.wrappers.purerl")
    purerl.GymnaxWrapper = type("GymnaxWrapper", (), {"__init__": lambda self, env: setattr(self, "_env", env)})
    spaces = types.ModuleType("gymnax.environments.spaces")
    spaces.Box = spaces.Discrete = lambda *a, **k: None
    sys.modules.update({"gymnax": gymnax, "gymnax.environments": envs, "gymnax.environments.environment": envmod,
                        "gymnax.environments.spaces": spaces, "gymnax.wrappers": wr, "gymnax.wrappers.purerl": purerl})
    sys.path.insert(0, REF)


# ------------------------------------------------------------------------------------------------------------------
def main():
    rand = RandomLog(123)
    install_shim(rand)
    sys.path.insert(0, ROOT)
    import quadjax  # the reference package, unmodified
    from quadjax.controllers.covo import CoVOController, CoVOParams
    from quadjax.controllers.pid import PIDController, PIDParams
    from quadjax.dynamics import geom
    from quadjax.dynamics import utils as rutils
    from quadjax.dynamics.dataclass import EnvParams3D, EnvState3D
    from quadjax.envs.quadrotor import Quad3D, get_controller
    from oracle import oracle_np as o  # only its trajectory generator / scenario builder: INPUTS, not outputs

    out_dir = os.path.join(ROOT, "tests", "golden")
    env = Quad3D(task="tracking_zigzag", dynamics="bodyrate", obs_type="quad", lower_controller="base", enable_randomizer=False,
                 disturb_type="none", disable_rollover_terminate=True, generate_noisy_state=True) \
        if "dynamics" in Quad3D.__init__.__code__.co_varnames else \
        Quad3D(task="tracking_zigzag", obs_type="quad", lower_controller="base", enable_randomizer=False, disturb_type="none",
               disable_rollover_terminate=True, generate_noisy_state=True)
    params = env.default_params
    # observations are not on the hot path (the MPC controllers ignore `obs`, controllers/covo.py:198) and get_obs relies
    # on JAX's clamped out-of-range gather, which NumPy does not have: stub it out
    env.get_obs = lambda *a, **k: None

    def to_ref_state(s: o.QuadState):
        z3 = np.zeros(3, F)
        return EnvState3D(pos=np.array(s.pos, F), vel=np.array(s.vel, F), quat=np.array(s.quat, F), omega=np.array(s.omega, F),
                          omega_tar=z3.copy(), pos_traj=s.pos_traj.astype(F), vel_traj=s.vel_traj.astype(F),
                          acc_traj=np.zeros_like(s.pos_traj, dtype=F), pos_tar=np.array(s.pos_tar, F), vel_tar=np.array(s.vel_tar, F),
                          acc_tar=z3.copy(), last_thrust=0.0, last_torque=z3.copy(), time=int(s.time), f_disturb=np.array(s.f_disturb, F),
                          vel_hist=np.zeros((env.default_params.adapt_horizon + 2, 3), F) if hasattr(params, "adapt_horizon") else np.zeros((4, 3), F),
                          omega_hist=np.zeros((env.default_params.adapt_horizon + 2, 3), F) if hasattr(params, "adapt_horizon") else np.zeros((4, 3), F),
                          action_hist=np.zeros((env.default_params.adapt_horizon + 2, 4), F) if hasattr(params, "adapt_horizon") else np.zeros((4, 4), F))

    # ---- 1. step_env chains: reward / done of the pre-step state, transition, targets, noisy state -------------------
    p = o.EnvParams()
    rng = np.random.default_rng(7)
    rec = {k: [] for k in ("state24", "time", "action", "next24", "next_time", "reward", "done", "err_pos", "noisy24", "noise13")}
    trajs = []
    for ep in range(4):
        s = o.reset_env("tracking_zigzag", p, np.random.default_rng(50 + ep), dtype=np.float32, zero_disturb=(ep % 2 == 0))
        if ep == 3:
            s.time = 296  # runs into max_steps_in_episode and the clamped end of the trajectory
            s.pos_tar, s.vel_tar = list(s.pos_traj[296]), list(s.vel_traj[296])
            s.pos = [F(x) for x in (s.pos_traj[296] + np.array([0.05, -0.02, 0.03]))]
        if ep == 2:
            s.pos = [F(2.99), F(0.0), F(-0.5)]  # about to leave the |pos| <= 3 box
            s.vel = [F(1.5), F(0.0), F(0.0)]
        trajs.append((s.pos_traj.astype(F), s.vel_traj.astype(F)))
        st = to_ref_state(s)
        for k in range(8):
            act = rng.uniform(-1.25, 1.25, size=4).astype(F)  # some components outside the clip box
            n0 = len(rand.normals)
            key = rand.PRNGKey(ep * 100 + k)
            _, nxt, reward, done, info = env.step_env(key, st, act, params, True)
            drawn = np.concatenate(rand.normals[n0:]) if len(rand.normals) > n0 else np.zeros(0, F)
            s24 = np.zeros(24, F)
            for dst, src in ((s24, st),):
                dst[0:3], dst[3:7], dst[7:10], dst[10:13], dst[13:16], dst[16:19], dst[19:22] = src.pos, src.quat, src.vel, src.omega, \
                    src.f_disturb, src.pos_tar, src.vel_tar
            n24 = np.zeros(24, F)
            n24[0:3], n24[3:7], n24[7:10], n24[10:13], n24[13:16], n24[16:19], n24[19:22] = nxt.pos, nxt.quat, nxt.vel, nxt.omega, \
                nxt.f_disturb, nxt.pos_tar, nxt.vel_tar
            ns = info["noisy_state"]
            z24 = np.zeros(24, F)
            z24[0:3], z24[3:7], z24[7:10], z24[10:13], z24[13:16], z24[16:19], z24[19:22] = ns.pos, ns.quat, ns.vel, ns.omega, \
                ns.f_disturb, ns.pos_tar, ns.vel_tar
            rec["state24"].append(s24); rec["time"].append(int(st.time)); rec["action"].append(act)
            rec["next24"].append(n24); rec["next_time"].append(int(nxt.time)); rec["reward"].append(F(reward)); rec["done"].append(bool(done))
            rec["err_pos"].append(F(info["err_pos"])); rec["noisy24"].append(z24)
            # get_info draws pos(3) vel(3) quat(4) omega(3) in this order (quadrotor.py:323-344); keep the last 13 draws
            rec["noise13"].append(drawn[-13:] if drawn.size >= 13 else np.zeros(13, F))
            st = nxt
        rec.setdefault("episode_len", []).append(8)
    np.savez_compressed(os.path.join(out_dir, "reference_step_env.npz"), pos_traj=np.stack([t[0] for t in trajs]),
                        vel_traj=np.stack([t[1] for t in trajs]), **{k: np.array(v) for k, v in rec.items()})

    # ---- 2. geometry / reward primitives ---------------------------------------------------------------------------
    qs = rng.standard_normal((16, 4)).astype(F)
    qs[:8] /= np.linalg.norm(qs[:8], axis=1, keepdims=True)  # the rest stay un-normalised on purpose
    es = np.abs(rng.standard_normal(32)).astype(F) * F(0.7)
    np.savez_compressed(os.path.join(out_dir, "reference_geom_reward.npz"), quat=qs,
                        qtoQ=np.stack([geom.qtoQ(q) for q in qs]), L=np.stack([geom.L(q) for q in qs]), H=np.asarray(geom.H),
                        err=es, log_pos=np.array([rutils.log_pos_fn(e) for e in es], F))

    # ---- 3. optimize_sigma on Hessians of the hot path (float32 in, float32 LAPACK eigh as in JAX-CPU) --------------
    Rs, Ss = [], []
    for H, seed in ((8, 1), (16, 2), (32, 3)):
        pp, ns_, a_mean, _ = __import__("tests.util", fromlist=["scenario"]).scenario("tracking_zigzag", seed=seed, H=H, warm_steps=6)
        R = o.get_hessian(ns_, a_mean, pp, dtype=np.float64).astype(F)
        ctl = CoVOController.__new__(CoVOController)
        ctl.action_dim, ctl.H = 4, H
        cp = types.SimpleNamespace(sample_sigma=F(0.5))
        S = CoVOController.optimize_sigma(ctl, R, cp)
        Rs.append(R.ravel()); Ss.append(np.asarray(S, F).ravel())
    np.savez_compressed(os.path.join(out_dir, "reference_optimize_sigma.npz"), H=np.array([8, 16, 32]),
                        R=np.concatenate(Rs), Sigma=np.concatenate(Ss))

    # ---- 4. PID expansion policy (covo-offline) --------------------------------------------------------------------
    pid = PIDController.__new__(PIDController)
    pid.env, pid.param = env, params
    pid_in, pid_out = [], []
    for k in range(12):
        s = o.reset_env("tracking_zigzag", p, np.random.default_rng(90 + k), dtype=np.float32, zero_disturb=True)
        s.pos = [F(x) for x in rng.normal(0, 0.3, 3)]
        s.vel = [F(x) for x in rng.normal(0, 0.5, 3)]
        q = rng.normal(0, 0.2, 4); q[3] = 1.0; q /= np.linalg.norm(q)
        s.quat = [F(x) for x in q]
        st = to_ref_state(s)
        cpp = PIDParams(Kp=10.0, Kd=5.0, Ki=0.0, Kp_att=10.0)
        act, _, _ = pid(None, st, params, None, cpp, None)
        s24 = o.state_to_vec24(s)
        pid_in.append(s24); pid_out.append(np.asarray(act, F))
    np.savez_compressed(os.path.join(out_dir, "reference_pid.npz"), state24=np.array(pid_in), action=np.array(pid_out))

    # ---- 5. get_controller defaults ---------------------------------------------------------------------------------
    ctl, cp = get_controller(env, "covo-online", "N64_H8_lam0.01")
    np.savez_compressed(os.path.join(out_dir, "reference_controller_defaults.npz"), a_mean=np.asarray(cp.a_mean, F),
                        N=ctl.N, H=ctl.H, lam=ctl.lam, sample_sigma=F(cp.sample_sigma), gamma_mean=F(cp.gamma_mean),
                        discount=F(getattr(cp, "discount", 1.0)))

    # ---- 6-8. whole controller calls executed from the reference source ---------------------------------------------
    scenario = __import__("tests.util", fromlist=["scenario"]).scenario

    def vec24(st):
        v = np.zeros(24, F)
        v[0:3], v[3:7], v[7:10], v[10:13], v[13:16], v[16:19], v[19:22] = st.pos, st.quat, st.vel, st.omega, st.f_disturb, st.pos_tar, st.vel_tar
        return v

    def interior(a, what):
        assert np.abs(np.asarray(a)).max() < 0.98, f"{what}: a nominal action sits on the clip bound, differences would straddle the kink"

    # 6. CoVOController.__call__ (online): Hessian of the reference's cost -> optimize_sigma -> sample -> rollouts -> update
    for tag, seed, text in (("lam0.01", 11, "N64_H8_lam0.01"), ("lam1.0", 12, "N96_H6_lam1.0")):
        ctl, cp = get_controller(env, "covo-online", text)
        H = ctl.H
        pp, ns_, a_prev, _ = scenario("tracking_zigzag", seed=seed, H=H, warm_steps=5)
        a_prev = np.clip(np.asarray(a_prev, F), -0.9, 0.9)
        interior(a_prev, "covo-online")
        st = to_ref_state(ns_)
        grabbed = {}
        orig = ctl.get_hessian
        ctl.get_hessian = lambda *a, **k: grabbed.setdefault("R", orig(*a, **k))  # instrumentation only
        m0 = len(rand.mvn)
        u, cp2, info = ctl(None, st, params, rand.PRNGKey(seed), cp.replace(a_mean=a_prev), {"noisy_state": st})
        eps = np.stack(rand.mvn[m0:])
        assert eps.shape == (ctl.N, 4 * H)
        np.savez_compressed(os.path.join(out_dir, f"reference_call_covo_online_{tag}.npz"), state24=vec24(st), time=int(st.time),
                            pos_traj=np.asarray(st.pos_traj, F), vel_traj=np.asarray(st.vel_traj, F), a_mean=a_prev, eps=eps,
                            R=np.asarray(grabbed["R"], F), a_cov=np.asarray(cp2.a_cov, F), a_mean_new=np.asarray(cp2.a_mean, F),
                            action=np.asarray(u, F), pos_mean=np.asarray(info["pos_mean"], F), pos_std=np.asarray(info["pos_std"], F),
                            lam=ctl.lam, N=ctl.N, H=H)
        R_o = o.get_hessian(ns_, o.shift_mean(a_prev), pp, dtype=np.float64)
        print(f"covo-online {tag}: |R_ref(fd of reference cost) - R_oracle(jets)| / |R| =",
              np.abs(grabbed["R"] - R_o).max() / np.abs(R_o).max())

    # 7. MPPIController.__call__
    ctl, cp = get_controller(env, "mppi", "N128_H8_lam0.01")
    pp, ns_, a_prev, _ = scenario("tracking_zigzag", seed=13, H=ctl.H, warm_steps=5)
    st = to_ref_state(ns_)
    m0 = len(rand.mvn)
    u, cp2, info = ctl(None, st, params, rand.PRNGKey(13), cp.replace(a_mean=np.asarray(a_prev, F)), {"noisy_state": st})
    eps = np.stack(rand.mvn[m0:]).reshape(ctl.N, ctl.H, 4)
    np.savez_compressed(os.path.join(out_dir, "reference_call_mppi.npz"), state24=vec24(st), time=int(st.time),
                        pos_traj=np.asarray(st.pos_traj, F), vel_traj=np.asarray(st.vel_traj, F), a_mean=np.asarray(a_prev, F),
                        a_cov=np.asarray(cp.a_cov, F), eps=eps, a_mean_new=np.asarray(cp2.a_mean, F), a_cov_new=np.asarray(cp2.a_cov, F),
                        action=np.asarray(u, F), pos_mean=np.asarray(info["pos_mean"], F), pos_std=np.asarray(info["pos_std"], F),
                        lam=ctl.lam, N=ctl.N, H=ctl.H)

    # 8. covo-offline schedule (reset_a_cov_offline, covo.py:58-104).  The full table is max_steps_in_episode = 300 Hessians;
    # the scan length is read from env.default_params at call time, so an env whose default_params says 3 steps yields the
    # first 3 entries through the untouched code.
    class ShortEpisode(Quad3D):
        default_params = property(lambda self: EnvParams3D(max_steps_in_episode=3))

    env_short = Quad3D(task="tracking_zigzag", obs_type="quad", lower_controller="base", enable_randomizer=False, disturb_type="none",
                       disable_rollover_terminate=True, generate_noisy_state=True)
    env_short.get_obs = lambda *a, **k: None
    env_short.__class__ = ShortEpisode
    ctl, cp = get_controller(env_short, "covo-offline", "N64_H6_lam0.01")
    pp, ns_, _, _ = scenario("tracking_zigzag", seed=14, H=ctl.H, warm_steps=0)
    st = to_ref_state(ns_)
    cp2 = ctl.reset(st, params, cp, rand.PRNGKey(14))
    table = np.asarray(cp2.a_cov_offline, F)
    assert table.shape == (3, 24, 24), table.shape
    np.savez_compressed(os.path.join(out_dir, "reference_covo_offline_schedule.npz"), state24=vec24(st), time=int(st.time),
                        pos_traj=np.asarray(st.pos_traj, F), vel_traj=np.asarray(st.vel_traj, F), a_cov_offline=table, H=ctl.H)

    # ---- 9. key plumbing: the reference's reset / step / trajectory generators driven by REAL jax.random semantics --------------
    # jax.random is now backed by covo_mpc_b200/jaxrng.py (Threefry-2x32-20, legacy layout; pinned by Random123 and
    # JAX-documentation known answers in tests/test_jaxrng.py).  What this section pins is the reference's USE of its keys:
    # which key is split how often and which draw feeds what (utils.py:87-130, 183-251; quadrotor.py:265-312, 314-370;
    # free.py:136-146; base.py:27-40) -- the product's host environment must consume a key the same way.
    from covo_mpc_b200 import jaxrng as jr

    rnd = sys.modules["jax.random"]
    rnd.PRNGKey = jr.PRNGKey
    rnd.split = lambda key, num=2: jr.split(np.asarray(key, np.uint32), num).view(ClampArr)
    rnd.uniform = lambda key, shape=(), dtype=F, minval=0.0, maxval=1.0: jr.uniform(np.asarray(key, np.uint32), shape, minval, maxval).view(JArr)
    rnd.normal = lambda key, shape=(), dtype=F: jr.normal(np.asarray(key, np.uint32), shape)
    keyed = {}
    for task in ("tracking", "tracking_zigzag", "hovering"):
        for disturb in ("none", "gaussian"):
            e = Quad3D(task=task, obs_type="quad", lower_controller="base", enable_randomizer=False, disturb_type=disturb,
                       disable_rollover_terminate=True, generate_noisy_state=True)
            e.get_obs = lambda *a, **k: None
            key = jr.PRNGKey(100 + len(keyed))
            _, info, st = e.reset_env(key, params)
            tag = f"{task}__{disturb}"
            keyed[tag + "__key"] = key
            keyed[tag + "__pos_traj"], keyed[tag + "__vel_traj"], keyed[tag + "__acc_traj"] = (np.asarray(x, F) for x in (st.pos_traj, st.vel_traj, st.acc_traj))
            keyed[tag + "__reset24"] = vec24(st)
            keyed[tag + "__reset_noisy24"] = vec24(info["noisy_state"])
            cur, k = st, jr.PRNGKey(7)
            steps_next, steps_noisy, steps_key = [], [], []
            for j in range(3):
                k, k_step = jr.split(k)
                act = np.array([0.2 * j - 0.3, 0.1, -0.05 * j, 0.02], F)
                # BaseEnvironment.step (base.py:15-40): key, key_reset = split(key); step_env(key, ...); auto-reset unused here
                k_env, _ = jr.split(k_step)
                _, cur, reward, done, info = e.step_env(k_env, cur, act, params)
                steps_key.append(k_step); steps_next.append(vec24(cur)); steps_noisy.append(vec24(info["noisy_state"]))
            keyed[tag + "__step_keys"], keyed[tag + "__step_next24"], keyed[tag + "__step_noisy24"] = np.array(steps_key), np.array(steps_next), np.array(steps_noisy)
    np.savez_compressed(os.path.join(out_dir, "reference_keyed_env.npz"), **keyed)

    # ---- 10. whole controller calls driven by a PRNGKey (jax.random = the Threefry twin of section 9) -----------------------
    # Same key in -> same action out is the drop-in claim: the product, handed this rng_act, must draw in its kernel what the
    # reference drew here (rng_act, act_key = split(rng_act); split(act_key, N); normal(...)) and end at the same action.
    rnd.multivariate_normal = lambda key, mean, cov: (np.asarray(mean, F) + np.linalg.cholesky(np.asarray(cov, F))
                                                      @ jr.normal(np.asarray(key, np.uint32), (np.asarray(mean).shape[-1],))).astype(F)
    for name, text, seed in (("covo-online", "N128_H8_lam0.01", 21), ("mppi", "N128_H8_lam0.01", 22)):
        ctl, cp = get_controller(env, name, text)
        pp, ns_, a_prev, _ = scenario("tracking_zigzag", seed=seed, H=ctl.H, warm_steps=5)
        a_prev = np.clip(np.asarray(a_prev, F), -0.9, 0.9)
        st = to_ref_state(ns_)
        key = jr.PRNGKey(1000 + seed)
        u, cp2, info = ctl(None, st, params, key, cp.replace(a_mean=a_prev), {"noisy_state": st})
        np.savez_compressed(os.path.join(out_dir, f"reference_call_{name.replace('-', '_')}_keyed.npz"), state24=vec24(st), time=int(st.time),
                            pos_traj=np.asarray(st.pos_traj, F), vel_traj=np.asarray(st.vel_traj, F), a_mean=a_prev, rng_act=key,
                            a_cov_in=np.asarray(cp.a_cov, F), a_cov=np.asarray(cp2.a_cov, F), a_mean_new=np.asarray(cp2.a_mean, F),
                            action=np.asarray(u, F), lam=ctl.lam, N=ctl.N, H=ctl.H)

    # ---- 10b. MPPI behaviours beyond the defaults, executed from the reference with a PRNGKey ------------------------------------
    #  * disturb_type "gaussian" (the reference's default): the rollouts call step_env WITHOUT deterministic=True (mppi.py:74), every
    #    sample and horizon step with the same step_key -> one shared random force for the whole call;
    #  * gamma_sigma != 0: the per-step covariance update of mppi.py:119-125.
    env_g = Quad3D(task="tracking_zigzag", obs_type="quad", lower_controller="base", enable_randomizer=False, disturb_type="gaussian",
                   disable_rollover_terminate=True, generate_noisy_state=True)
    env_g.get_obs = lambda *a, **k: None
    for tag, e_, gs, seed in (("gaussian", env_g, 0.0, 23), ("gamma_sigma", env, 0.3, 24)):
        ctl, cp = get_controller(e_, "mppi", "N128_H8_lam0.5" if gs else "N128_H8_lam0.01")
        pp, ns_, a_prev, _ = scenario("tracking_zigzag", seed=seed, H=ctl.H, warm_steps=5, zero_disturb=False)
        a_prev = np.clip(np.asarray(a_prev, F), -0.9, 0.9)
        st = to_ref_state(ns_)
        key = jr.PRNGKey(1000 + seed)
        cp_in = cp.replace(a_mean=a_prev, gamma_sigma=gs)
        u, cp2, info = ctl(None, st, params, key, cp_in, {"noisy_state": st})
        np.savez_compressed(os.path.join(out_dir, f"reference_call_mppi_keyed_{tag}.npz"), state24=vec24(st), time=int(st.time),
                            pos_traj=np.asarray(st.pos_traj, F), vel_traj=np.asarray(st.vel_traj, F), a_mean=a_prev, rng_act=key,
                            a_cov_in=np.asarray(cp.a_cov, F), a_cov=np.asarray(cp2.a_cov, F), a_mean_new=np.asarray(cp2.a_mean, F),
                            action=np.asarray(u, F), lam=ctl.lam, N=ctl.N, H=ctl.H, gamma_sigma=gs,
                            dyn_noise_scale=float(params.dyn_noise_scale))

    # ---- 11. the whole evaluation protocol: eval_env executed from the reference (envs/quadrotor.py:506-591) -------------------
    # Controller = the reference's RandomController (0.3 * normal(rng_act, (4,))): CPU-only, and wild enough to fly out of the
    # |pos| <= 3 box, so BaseEnvironment.step's auto-reset (base.py:27-38) is exercised.  Pins the key schedule of the
    # protocol end to end: PRNGKey(1), reset keys, per-step split(rng, 4), env.step, the extra split after every step.
    import tempfile

    import quadjax as _q
    from quadjax.controllers import RandomController
    from quadjax.envs import quadrotor as ref_quadrotor

    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "pkg"))
    _q.get_package_path = lambda: os.path.join(tmp, "pkg")  # results/ lands under tmp instead of the read-only reference tree
    ev = {}
    for disturb in ("none", "gaussian"):
        e = Quad3D(task="tracking_zigzag", obs_type="quad", lower_controller="base", enable_randomizer=False, disturb_type=disturb,
                   disable_rollover_terminate=True, generate_noisy_state=True)
        e.get_obs = lambda *a, **k: None
        ref_quadrotor.eval_env(e, RandomController(e, None), total_steps=300 * 4, filename=f"random_{disturb}")
        import pickle

        with open(os.path.join(tmp, "results", f"eval_err_pos_random_{disturb}.pkl"), "rb") as f:
            ev[disturb] = np.asarray(pickle.load(f), np.float64)
    np.savez_compressed(os.path.join(out_dir, "reference_eval_env_random.npz"), **ev)

    # ---- 12. render_env executed from the reference (envs/quadrotor.py:594-667), RandomController, one episode -----------------
    # utils.plot_states (matplotlib figures) is out of scope and stubbed; everything else, including the state_seq pickle, is the
    # reference's own code.
    from quadjax.dynamics import utils as ref_utils

    ref_utils.plot_states = lambda *a, **k: None
    e = Quad3D(task="tracking_zigzag", obs_type="quad", lower_controller="base", enable_randomizer=False, disturb_type="gaussian",
               disable_rollover_terminate=True, generate_noisy_state=True)
    e.get_obs = lambda *a, **k: None
    rc = RandomController(e, None)
    ref_quadrotor.render_env(e, rc, None, repeat_times=1, filename="random")
    with open(os.path.join(tmp, "results", "state_seq_random.pkl"), "rb") as f:
        seq = pickle.load(f)
    np.savez_compressed(os.path.join(out_dir, "reference_render_env_random.npz"),
                        keys=np.array(sorted(seq[0].keys())), n_steps=len(seq),
                        **{k: np.stack([np.asarray(d[k], np.float32) for d in seq]) for k in ("pos", "vel", "quat", "omega", "f_disturb", "pos_tar", "vel_tar")},
                        time=np.array([int(d["time"]) for d in seq]))
    print("reference goldens written to", out_dir)


if __name__ == "__main__":
    main()
