"""The oracle against golden vectors produced by the REFERENCE'S OWN SOURCE (tests/golden/make_reference_golden.py runs
/root/reference/quadjax with NumPy standing in for jax.numpy; see its header for what that does and does not cover).
These pin the per-step maths of the hot path -- step_env (dynamics, reward and done of the pre-step state, clamped
target gather), get_info's noisy state, geometry, log_pos, optimize_sigma, the PID expansion policy and get_controller's
defaults -- as the reference wrote them.  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle_np as o

G = os.path.join(os.path.dirname(__file__), "golden")
F = np.float32


class SeqRng:
    def __init__(self, values):
        self.v = [float(x) for x in values]

    def standard_normal(self):
        return self.v.pop(0)


def _state(s24, t, pos_traj, vel_traj):
    return o.make_state(s24[0:3], s24[3:7], s24[7:10], s24[10:13], s24[13:16], int(t), pos_traj, vel_traj, s24[16:19], s24[19:22],
                        dtype=np.float32)


def test_step_env_matches_the_reference_source():
    g = np.load(os.path.join(G, "reference_step_env.npz"))
    p = o.EnvParams()
    n = len(g["time"])
    assert n == 32 and g["done"].any() and not g["done"].all()  # the fixtures cross both termination conditions
    for i in range(n):
        ep = i // 8
        s = _state(g["state24"][i], g["time"][i], g["pos_traj"][ep], g["vel_traj"][ep])
        nxt, r, d, e = o.env_step(s, g["action"][i], p, None, "none")
        assert abs(r - g["reward"][i]) < 3e-6 * max(1.0, abs(g["reward"][i])), i
        assert d == bool(g["done"][i]) and abs(e - g["err_pos"][i]) < 3e-6, i
        assert nxt.time == g["next_time"][i]
        assert np.abs(o.state_to_vec24(nxt) - g["next24"][i]).max() < 3e-6, i
        ns = o.noisy_state(nxt, p, SeqRng(g["noise13"][i]))
        assert np.abs(o.state_to_vec24(ns) - g["noisy24"][i]).max() < 3e-6, i


def test_geometry_and_log_pos_match_the_reference_source():
    g = np.load(os.path.join(G, "reference_geom_reward.npz"))
    for q, Q in zip(g["quat"], g["qtoQ"]):
        assert np.abs(np.asarray(o._qtoQ([F(x) for x in q]), dtype=np.float64) - Q).max() < 3e-6 * max(1.0, float(q @ q))
    for e, lp in zip(g["err"], g["log_pos"]):
        assert abs(float(o.log_pos_fn(F(e))) - lp) < 2e-6


def test_optimize_sigma_matches_the_reference_source():
    g = np.load(os.path.join(G, "reference_optimize_sigma.npz"))
    off = 0
    for H in g["H"]:
        n = 4 * int(H)
        R = g["R"][off:off + n * n].reshape(n, n)
        S_ref = g["Sigma"][off:off + n * n].reshape(n, n)
        off += n * n
        S = o.optimize_sigma(R, 0.5, np.float32)
        assert np.linalg.norm(S - S_ref) / np.linalg.norm(S_ref) < 2e-6  # same formula, same LAPACK class of eigh
        S64 = o.optimize_sigma(R.astype(np.float64), 0.5, np.float64)
        assert np.linalg.norm(S64 - S_ref) / np.linalg.norm(S_ref) < 1e-4


def test_pid_policy_matches_the_reference_source():
    g = np.load(os.path.join(G, "reference_pid.npz"))
    p = o.EnvParams()
    z = np.zeros((320, 3), np.float32)
    for s24, act in zip(g["state24"], g["action"]):
        s = _state(s24, 0, z, z)
        a = o.pid_action(s, p)
        a = a[0] if isinstance(a, tuple) else a
        assert np.abs(np.asarray(a, dtype=np.float64) - act).max() < 1e-5


def test_controller_defaults_match_the_reference_source():
    g = np.load(os.path.join(G, "reference_controller_defaults.npz"))
    p = o.EnvParams()
    assert np.abs(o.hover_mean(int(g["H"]), p) - g["a_mean"]).max() < 1e-7
    assert int(g["N"]) == 64 and int(g["H"]) == 8 and abs(float(g["lam"]) - 0.01) < 1e-9
    assert abs(float(g["sample_sigma"]) - 0.5) < 1e-7 and abs(float(g["gamma_mean"]) - 1.0) < 1e-7
    # the product's own get_controller (host side, no GPU needed until the first call)
    import covo_mpc_b200 as cm

    ctl, cp = cm.get_controller(cm.Quad3D("tracking_zigzag"), "covo-online", "N64_H8_lam0.01")
    assert np.abs(np.asarray(cp.a_mean, np.float32).reshape(-1, 4) - g["a_mean"]).max() < 1e-7
    assert ctl.N == 64 and ctl.H == 8 and abs(ctl.lam - 0.01) < 1e-9


# ---- whole controller calls executed from the reference source (generator sections 6-8) ---------------------------------------
def _call_state(g):
    return _state(g["state24"], g["time"], g["pos_traj"], g["vel_traj"])


@pytest.mark.parametrize("tag", ["lam0.01", "lam1.0"])
def test_covo_online_call_matches_the_reference_source(tag):
    """CoVOController.__call__ run from /root/reference: R is the Hessian of the reference's own cost function (float64
    differences), Sigma its optimize_sigma, the update its sampling + vmapped step_env rollouts + softmax."""
    g = np.load(os.path.join(G, f"reference_call_covo_online_{tag}.npz"))
    p = o.EnvParams()
    u, new_mean, a_cov, info, dbg = o.covo_call(_call_state(g), g["a_mean"], g["eps"], p, lam=float(g["lam"]), return_debug=True)
    scale = np.abs(g["R"]).max()
    assert np.abs(dbg["R"] - g["R"]).max() < 5e-7 * scale  # float32 rounding of the stored matrix is 6e-8
    assert np.abs(dbg["R"][-4:, -4:]).max() == 0.0 and np.abs(g["R"][-4:, -4:]).max() < 1e-6 * scale
    assert np.linalg.norm(a_cov - g["a_cov"]) / np.linalg.norm(g["a_cov"]) < 2e-5
    assert np.abs(new_mean - g["a_mean_new"]).max() < 2e-5 and np.abs(u - g["action"]).max() < 2e-5
    assert np.abs(info["pos_mean"] - g["pos_mean"]).max() < 1e-5 and np.abs(info["pos_std"] - g["pos_std"]).max() < 1e-5


def test_mppi_call_matches_the_reference_source():
    g = np.load(os.path.join(G, "reference_call_mppi.npz"))
    u, new_mean, new_cov, info = o.mppi_call(_call_state(g), g["a_mean"], g["a_cov"], g["eps"], o.EnvParams(), lam=float(g["lam"]))
    assert np.abs(new_mean - g["a_mean_new"]).max() < 1e-5 and np.abs(u - g["action"]).max() < 1e-5
    assert np.abs(new_cov - g["a_cov_new"]).max() < 1e-6
    assert np.abs(info["pos_mean"] - g["pos_mean"]).max() < 1e-5 and np.abs(info["pos_std"] - g["pos_std"]).max() < 1e-5


def test_covo_offline_schedule_matches_the_reference_source():
    """reset_a_cov_offline: PID nominal rollout -> Hessian -> optimize_sigma -> one PID step, first 3 table entries."""
    g = np.load(os.path.join(G, "reference_covo_offline_schedule.npz"))
    tab = o.covo_offline_schedule(_call_state(g), o.EnvParams(), int(g["H"]), 0.5, np.random.default_rng(0), n_steps=3)
    for k in range(3):
        assert np.linalg.norm(tab[k] - g["a_cov_offline"][k]) / np.linalg.norm(g["a_cov_offline"][k]) < 2e-5, k


@pytest.mark.parametrize("name", ["covo_online", "mppi"])
def test_keyed_call_matches_the_reference_source(name):
    """The reference's __call__ handed a PRNGKey (jax.random = the Threefry twin, generator section 10): the oracle fed the
    normals that key produces lands on the same update."""
    from covo_mpc_b200 import jaxrng as jr

    g = np.load(os.path.join(G, f"reference_call_{name}_keyed.npz"))
    N, H = int(g["N"]), int(g["H"])
    act_key = jr.split(g["rng_act"])[1]  # rng_act, act_key = split(rng_act)   (covo.py:212, mppi.py:53)
    if name == "mppi":
        u, new_mean, _, _ = o.mppi_call(_call_state(g), g["a_mean"], g["a_cov_in"], jr.mppi_normals(act_key, N, H), o.EnvParams(),
                                        lam=float(g["lam"]))
    else:
        u, new_mean, a_cov, _ = o.covo_call(_call_state(g), g["a_mean"], jr.covo_normals(act_key, N, 4 * H), o.EnvParams(), lam=float(g["lam"]))
        assert np.linalg.norm(a_cov - g["a_cov"]) / np.linalg.norm(g["a_cov"]) < 2e-5
    assert np.abs(new_mean - g["a_mean_new"]).max() < 2e-5 and np.abs(u - g["action"]).max() < 2e-5


@pytest.mark.parametrize("tag", ["gaussian", "gamma_sigma"])
def test_mppi_beyond_the_defaults_matches_the_reference_source(tag):
    """Generator section 10b: MPPI under disturb_type gaussian (stochastic rollouts, one step_key for every sample and horizon step:
    mppi.py:69-74) and with gamma_sigma = 0.3 (covariance update, mppi.py:119-125), both executed from the reference with a PRNGKey.
    The oracle, fed the normals and the force that key produces, lands on the same update."""
    from covo_mpc_b200 import jaxrng as jr

    g = np.load(os.path.join(G, f"reference_call_mppi_keyed_{tag}.npz"))
    N, H = int(g["N"]), int(g["H"])
    rng_act, act_key = jr.split(g["rng_act"])        # mppi.py:53
    step_key = jr.split(rng_act)[1]                   # mppi.py:69
    fd = None
    if tag == "gaussian":  # step_env -> raw_step -> step_fn: the disturbance key is three splits down (quadrotor.py:262, free.py:136-144)
        z = jr.normal(jr.split(jr.split(jr.split(step_key)[1])[0])[0], (3,))
        fd = np.tile((np.float32(g["dyn_noise_scale"]) * z)[None], (H, 1)).astype(np.float32)
    u, new_mean, new_cov, _ = o.mppi_call(_call_state(g), g["a_mean"], g["a_cov_in"], jr.mppi_normals(act_key, N, H), o.EnvParams(),
                                          lam=float(g["lam"]), gamma_sigma=float(g["gamma_sigma"]), f_disturb_seq=fd)
    assert np.abs(new_mean - g["a_mean_new"]).max() < 2e-5 and np.abs(u - g["action"]).max() < 2e-5
    assert np.abs(new_cov - g["a_cov"]).max() < 2e-5
