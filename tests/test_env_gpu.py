"""Device-resident environment step and closed loop against the oracle's env_step / noisy_state
(envs/quadrotor.py:215-248, 314-361) and against the host-driven loop."""
import numpy as np
import pytest

from oracle import oracle_np as o
from tests.util import scenario

pytestmark = pytest.mark.gpu


class SeqRng:
    """Hands out pre-drawn standard normals in the order the oracle asks for them."""

    def __init__(self, values):
        self.v = list(values)

    def standard_normal(self):
        return self.v.pop(0)


def _handle(mode, N, H, T, E=1, seed=0):
    from covo_mpc_b200 import _lib

    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.n_env, cfg.seed = mode, N, H, T, E, seed
    return _lib.Handle(cfg)


@pytest.mark.parametrize("disturb", ["none", "gaussian"])
def test_env_step_matches_oracle(disturb):
    from covo_mpc_b200 import _lib

    E = 3
    scen = [scenario("tracking_zigzag", seed=40 + e, H=8, warm_steps=3 + 4 * e, zero_disturb=(e == 0)) for e in range(E)]
    p = scen[0][0]
    T = scen[0][1].pos_traj.shape[0]
    h = _handle(_lib.MODE_MPPI, 64, 8, T, E=E)
    h.set_reference(np.stack([s[1].pos_traj for s in scen]), np.stack([s[1].vel_traj for s in scen]))
    states = [s[1] for s in scen]
    states[2].time = 318  # next targets come from the clamped last row of the trajectory
    h.env_reset(np.stack([o.state_to_vec24(s) for s in states]), [s.time for s in states])
    rng = np.random.default_rng(3)
    for step in range(3):
        act = rng.uniform(-1.3, 1.3, size=(E, 4)).astype(np.float32)  # some components outside the clip box
        z = rng.standard_normal((E, 16)).astype(np.float32)
        noisy, rew, err, done = h.env_step(act, noise=z, gaussian=(disturb == "gaussian"))
        s24, tm = h.env_state()
        for e in range(E):
            seq = ([float(x) for x in z[e, 13:16]] if disturb == "gaussian" else []) + [float(x) for x in z[e, :13]]
            srng = SeqRng(seq)
            nxt, r_o, d_o, e_o = o.env_step(states[e], act[e], p, srng, disturb)
            ns_o = o.noisy_state(nxt, p, srng)
            assert abs(rew[e] - r_o) < 2e-6 * max(1.0, abs(r_o)) and abs(err[e] - e_o) < 2e-6 and bool(done[e]) == d_o
            assert np.abs(s24[e] - o.state_to_vec24(nxt)).max() < 3e-6
            assert tm[e] == nxt.time
            assert np.abs(noisy[e] - o.state_to_vec24(ns_o)).max() < 3e-6
            states[e] = nxt


@pytest.mark.parametrize("mode_name", ["mppi", "covo-online", "mppi-gaussian"])
def test_closed_loop_on_device_equals_step_by_step(mode_name):
    """covo_closed_loop (no host round trip) == the same handle type driven one call at a time through
    covo_env_step + covo_step with the same seeds: identical kernels, so the actions agree bit for bit.
    mppi-gaussian: disturb_type gaussian -- the environment's force comes from the supplied normals, and the force MPPI's rollouts plan
    against (mppi.py:74, one draw per call) is written by the environment kernel for the next controller call in both loops."""
    from covo_mpc_b200 import _lib

    gauss = mode_name.endswith("gaussian")
    mode = _lib.MODE_MPPI if mode_name.startswith("mppi") else _lib.MODE_COVO_ONLINE
    N, H, steps = 256, 10, 6
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=9, H=H, warm_steps=4)
    T = ns.pos_traj.shape[0]
    noise = rng.standard_normal((steps + 1, 1, 16)).astype(np.float32)

    def mk():
        h = _handle(mode, N, H, T, seed=21)
        h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
        h.set_mean(a_mean[None])
        h.env_reset(o.state_to_vec24(ns)[None], [ns.time])
        return h

    ha = mk()
    act_a, rew_a, err_a = ha.closed_loop(steps, noise=noise, gaussian=gauss)
    hb = mk()
    if gauss:
        hb.set_rollout_disturbance(np.zeros((1, H, 3), np.float32))  # switches the planning force on; the kernel overwrites it
    noisy, _, _, _ = hb.env_step(None, noise=noise[0], noise_step=0, gaussian=gauss)
    _, tm = hb.env_state()
    for i in range(steps):
        a = hb.step(noisy, tm)
        assert np.array_equal(a, act_a[i])
        noisy, rew, err, _ = hb.env_step(a, noise=noise[i + 1], noise_step=i + 1, gaussian=gauss)
        _, tm = hb.env_state()
        assert rew[0] == rew_a[i, 0] and err[0] == err_a[i, 0]
    sa, ta = ha.env_state()
    sb, tb = hb.env_state()
    assert np.array_equal(sa, sb) and np.array_equal(ta, tb)
    assert np.isfinite(act_a).all() and np.abs(act_a).max() <= 1.0 + 1e-6
    if gauss:  # the planning force does something: the same loop without it takes other actions
        hc = mk()
        act_c, _, _ = hc.closed_loop(steps, noise=noise, gaussian=False)
        assert not np.array_equal(act_c, act_a)


def test_run_episode_device_tracks_like_the_host_loop():
    """A short episode through the harness with the environment on the device: the tracking error stays in the
    range the host-driven episode reaches (different noise streams, same protocol)."""
    import covo_mpc_b200 as cm

    env = cm.Quad3D("tracking_zigzag")
    ctl, _ = cm.get_controller(env, "covo-offline", "N512_H16_lam0.01", seed=5)
    err_h, _ = cm.run_episode(env, ctl, np.random.default_rng(1), n_steps=40, reset_rng=np.random.default_rng(2))
    err_d, rew_d = cm.run_episode_device(env, ctl, np.random.default_rng(1), n_steps=40, reset_rng=np.random.default_rng(2))
    assert err_d.shape == (40,) and np.isfinite(err_d).all() and np.isfinite(rew_d).all()
    assert err_d.mean() < 3.0 * err_h.mean() + 0.05
    ctl.close()


# ---- golden vectors produced by the reference's own source (tests/golden/make_reference_golden.py) ---------------------------
def _golden(name):
    import os

    return np.load(os.path.join(os.path.dirname(__file__), "golden", name))


def test_env_step_kernel_matches_the_reference_source():
    """Quad3D.step_env + get_info executed from /root/reference (NumPy standing in for jax.numpy) vs env_step_kernel:
    4 chained episodes x 8 steps, crossing the |pos| > 3 box (episode 2) and time >= max_steps (episode 3)."""
    from covo_mpc_b200 import _lib

    g = _golden("reference_step_env.npz")
    for ep in range(4):
        h = _handle(_lib.MODE_MPPI, 64, 8, 320)
        h.set_reference(g["pos_traj"][ep][None], g["vel_traj"][ep][None])
        i0 = 8 * ep
        h.env_reset(g["state24"][i0][None], [int(g["time"][i0])])
        for i in range(i0, i0 + 8):
            z = np.zeros((1, 16), np.float32)
            z[0, :13] = g["noise13"][i]
            noisy, rew, err, done = h.env_step(g["action"][i][None], noise=z)
            s24, tm = h.env_state()
            assert abs(rew[0] - g["reward"][i]) < 3e-6 * max(1.0, abs(g["reward"][i])), i
            assert abs(err[0] - g["err_pos"][i]) < 3e-6 and bool(done[0]) == bool(g["done"][i]), i
            assert int(tm[0]) == int(g["next_time"][i])
            assert np.abs(s24[0] - g["next24"][i]).max() < 3e-6, i
            assert np.abs(noisy[0] - g["noisy24"][i]).max() < 3e-6, i


def test_rollout_kernel_cost_matches_the_reference_source_chain():
    """A rollout of the recorded action sequence (zero perturbation) must cost what the reference's step_env chain paid:
    -sum_h reward_h, the reward freezing after the first terminal pre-step state (controllers/covo.py:233)."""
    from covo_mpc_b200 import _lib

    g = _golden("reference_step_env.npz")
    for ep in range(4):
        h = _handle(_lib.MODE_MPPI, 64, 8, 320)
        h.set_reference(g["pos_traj"][ep][None], g["vel_traj"][ep][None])
        sl = slice(8 * ep, 8 * ep + 8)
        a = np.clip(g["action"][sl], -1.0, 1.0)
        _, _, costs, _ = h.rollout(g["state24"][8 * ep][None], [int(g["time"][8 * ep])], a[None],
                                   eps=np.zeros((1, 64, 32), np.float32), want_costs=True)
        r, d = g["reward"][sl].astype(np.float64), g["done"][sl]
        total, frozen, r_before = 0.0, False, 0.0
        for k in range(8):
            rk = r_before if frozen else r[k]
            total, r_before, frozen = total + rk, rk, frozen or bool(d[k])
        assert np.abs(costs[0] + total).max() < 2e-5 * max(1.0, abs(total)), (ep, costs[0][:2], -total)


def test_optimize_sigma_kernels_match_the_reference_source():
    """optimize_sigma as executed from the reference (float32 LAPACK eigh) vs E1-E3."""
    from covo_mpc_b200 import _lib

    g = _golden("reference_optimize_sigma.npz")
    off = 0
    for H in g["H"]:
        n = 4 * int(H)
        R = g["R"][off:off + n * n].reshape(n, n)
        S_ref = g["Sigma"][off:off + n * n].reshape(n, n)
        off += n * n
        h = _handle(_lib.MODE_COVO_ONLINE, 64, int(H), 320)
        S = h.optimize_sigma(R[None])[0]
        assert np.linalg.norm(S - S_ref) / np.linalg.norm(S_ref) < 5e-5, H


def test_pid_controller_matches_the_reference_source():
    """get_controller(env, "pid"): the device PID policy vs PIDController.__call__ executed from /root/reference
    (tests/golden/reference_pid.npz), through the controller object; integral bookkeeping as pid.py:77-81."""
    import covo_mpc_b200 as cm

    g = _golden("reference_pid.npz")
    env = cm.Quad3D("tracking_zigzag")
    ctl, cp = cm.get_controller(env, "pid")
    assert (cp.Kp, cp.Kd, cp.Ki, cp.Kp_att) == (10.0, 5.0, 0.0, 10.0)
    z = np.zeros((320, 3), np.float32)
    for s, act in zip(g["state24"], g["action"]):
        st = cm.EnvState3D(pos=s[0:3], quat=s[3:7], vel=s[7:10], omega=s[10:13], f_disturb=s[13:16], pos_tar=s[16:19], vel_tar=s[19:22],
                           acc_tar=z[0], pos_traj=z, vel_traj=z, acc_traj=z, time=0)
        a, cp2, info = ctl(None, st, env.default_params, None, cp, None)
        assert info is None and np.abs(a - act).max() < 2e-5
        assert np.allclose(cp2.integral, (s[0:3] - s[16:19]) * np.float32(0.02), atol=1e-7)
    # Ki * integral term (pid.py:48) against the oracle's restatement
    integ = np.array([0.05, -0.1, -0.2], np.float32)
    cpi = cm.PIDParams(Kp=10.0, Kd=5.0, Ki=2.0, Kp_att=10.0, integral=integ)
    for s in g["state24"][:6]:
        st = cm.EnvState3D(pos=s[0:3], quat=s[3:7], vel=s[7:10], omega=s[10:13], f_disturb=s[13:16], pos_tar=s[16:19], vel_tar=s[19:22],
                           acc_tar=z[0], pos_traj=z, vel_traj=z, acc_traj=z, time=0)
        a_i, _, _ = ctl(None, st, env.default_params, None, cpi, None)
        so = o.make_state(s[0:3], s[3:7], s[7:10], s[10:13], s[13:16], 0, z, z, s[16:19], s[19:22], dtype=np.float32)
        a_o = o.pid_action(so, o.EnvParams(), Kp=10.0, Kd=5.0, Ki=2.0, Kp_att=10.0, integral=integ)
        a_o = a_o[0] if isinstance(a_o, tuple) else a_o
        assert np.abs(a_i - np.asarray(a_o, np.float32)).max() < 2e-5
    # a PID episode flies the reference trajectory
    errs, _ = cm.run_episode(env, ctl, np.random.default_rng(0), n_steps=80)
    assert np.isfinite(errs).all() and errs.mean() < 0.5
    ctl_r, cp_r = cm.get_controller(env, "random")
    a_r, _, _ = ctl_r(None, st, env.default_params, cm.jaxrng.PRNGKey(0), cp_r, None)
    assert a_r.shape == (4,) and np.allclose(a_r, 0.3 * cm.jaxrng.normal(cm.jaxrng.PRNGKey(0), (4,)))


def test_auto_reset_on_device():
    """BaseEnvironment.step's auto-reset (envs/base.py:27-38) on the device: when the PRE-step state is terminal (here: time >= a
    shortened max_steps_in_episode, and one environment flown out of the |pos| <= 3 box), the stepped state is discarded, the
    environment continues from its next reset-pool entry with that entry's reference trajectory, and the controller's mean goes back
    to its initial value (what render_env's controller.reset does after `done`, envs/quadrotor.py:637-639)."""
    import dataclasses

    from covo_mpc_b200 import _lib

    E, H, P, steps = 2, 6, 2, 17
    scen = [scenario("tracking_zigzag", seed=60 + e, H=H, warm_steps=2) for e in range(E)]
    p = dataclasses.replace(scen[0][0], max_steps_in_episode=4)
    T = scen[0][1].pos_traj.shape[0]
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.n_env, cfg.max_steps_in_episode = _lib.MODE_MPPI, 64, H, T, E, 4
    h = _lib.Handle(cfg)
    h.set_reference(np.stack([s[1].pos_traj for s in scen]), np.stack([s[1].vel_traj for s in scen]))
    states = [s[1].copy() for s in scen]
    for s in states:
        s.time = 1
    states[1].pos = [np.float32(2.9), np.float32(0.0), np.float32(0.0)]
    states[1].vel = [np.float32(8.0), np.float32(0.0), np.float32(0.0)]  # leaves the box after one step
    pool = [[scenario("tracking_zigzag", seed=80 + 7 * k + e, H=H, warm_steps=0)[1] for e in range(E)] for k in range(P)]
    for k in range(P):
        for e in range(E):
            pool[k][e].time = 0
    mean_init = o.hover_mean(H, p)
    h.set_mean(np.stack([mean_init + 0.1 * (e + 1) for e in range(E)]))
    h.env_set_reset_pool(np.stack([[o.state_to_vec24(pool[k][e]) for e in range(E)] for k in range(P)]),
                         np.zeros((P, E), np.int32),
                         np.stack([[pool[k][e].pos_traj for e in range(E)] for k in range(P)]),
                         np.stack([[pool[k][e].vel_traj for e in range(E)] for k in range(P)]), mean_init)
    h.env_reset(np.stack([o.state_to_vec24(s) for s in states]), [s.time for s in states])
    rng = np.random.default_rng(5)
    count = [0] * E
    n_resets = 0
    for step in range(steps):
        act = rng.uniform(-0.5, 0.5, size=(E, 4)).astype(np.float32)
        z = rng.standard_normal((E, 16)).astype(np.float32)
        noisy, rew, err, done = h.env_step(act, noise=z)
        s24, tm = h.env_state()
        for e in range(E):
            srng = SeqRng([float(x) for x in z[e, :13]])
            nxt, r_o, d_o, e_o = o.env_step(states[e], act[e], p, srng, "none")
            assert bool(done[e]) == d_o and abs(rew[e] - r_o) < 2e-6 * max(1.0, abs(r_o))
            if d_o:
                nxt = pool[count[e] % P][e].copy()
                count[e] += 1
                n_resets += 1
                assert np.array_equal(h.get_mean()[e].reshape(H, 4), mean_init)
            ns_o = o.noisy_state(nxt, p, srng)
            assert np.abs(s24[e] - o.state_to_vec24(nxt)).max() < 3e-6 and tm[e] == nxt.time, (step, e)
            assert np.abs(noisy[e] - o.state_to_vec24(ns_o)).max() < 3e-6
            states[e] = nxt
    assert n_resets >= 6 and count[1] >= 3  # both environments went through the pool more than once (round robin)
    # switched off again: a terminal pre-step state is reported but nothing is replaced
    h.env_set_reset_pool(None, None, None, None)
    for step in range(5):
        noisy, rew, err, done = h.env_step(np.zeros((E, 4), np.float32), noise=np.zeros((E, 16), np.float32))
    assert done.all() and (h.env_state()[1] >= 4).all()
    h.close()
