"""Host environment driven by JAX PRNG keys vs the reference's own reset / step / trajectory generators executed with the
same keys (tests/golden/make_reference_golden.py section 9; jax.random backed by covo_mpc_b200/jaxrng.py, which is pinned
separately in tests/test_jaxrng.py).  Pins the key plumbing: split tree, draw order, which draw feeds what."""
import os

import numpy as np
import pytest

import covo_mpc_b200 as cm
from covo_mpc_b200 import jaxrng as jr

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_keyed_env.npz"))


def _vec24(st):
    return st.to_state24()


@pytest.mark.parametrize("task", ["tracking", "tracking_zigzag", "hovering"])
@pytest.mark.parametrize("disturb", ["none", "gaussian"])
def test_reset_and_step_consume_keys_like_the_reference(task, disturb):
    tag = f"{task}__{disturb}"
    env = cm.Quad3D(task, disturb_type=disturb)
    key = G[tag + "__key"]
    assert jr.is_key(key)
    _, info, st = env.reset(key)
    # trajectories: float32 here vs the shim's mixed float32/float64 evaluation of the same expressions
    assert st.pos_traj.shape == G[tag + "__pos_traj"].shape
    assert np.abs(st.pos_traj - G[tag + "__pos_traj"]).max() < 2e-5
    assert np.abs(st.vel_traj - G[tag + "__vel_traj"]).max() < 2e-5
    assert np.abs(st.acc_traj - G[tag + "__acc_traj"]).max() < 1e-4
    assert np.abs(_vec24(st) - G[tag + "__reset24"]).max() < 2e-5
    assert np.array_equal(st.f_disturb, G[tag + "__reset24"][13:16])  # uniform(disturb_key): bit-identical
    assert np.abs(_vec24(info["noisy_state"]) - G[tag + "__reset_noisy24"]).max() < 2e-5
    noise = _vec24(info["noisy_state"])[:13] - _vec24(st)[:13]
    assert np.abs(noise - (G[tag + "__reset_noisy24"][:13] - G[tag + "__reset24"][:13])).max() < 1e-7  # the same draws
    cur = st
    for j in range(3):
        act = np.array([0.2 * j - 0.3, 0.1, -0.05 * j, 0.02], np.float32)
        _, cur, reward, done, info = env.step(G[tag + "__step_keys"][j], cur, act)
        assert not done
        assert np.abs(_vec24(cur) - G[tag + "__step_next24"][j]).max() < 2e-5, j
        assert np.abs(_vec24(info["noisy_state"]) - G[tag + "__step_noisy24"][j]).max() < 2e-5, j
        if disturb == "gaussian":
            assert np.abs(cur.f_disturb).max() > 0 and np.abs(cur.f_disturb - G[tag + "__step_next24"][j][13:16]).max() < 1e-7
        else:
            assert np.abs(cur.f_disturb).max() == 0


def test_zigzag_key_quirks():
    """Segments 0 and 1 are drawn from the same key (dynamics/utils.py:238, 241): same turn angles relative to the centre
    direction and the same length."""
    pos, vel, acc = cm.env.generate_zigzag_traj(300, 0.02, jr.PRNGKey(5))
    assert pos.shape == (320, 3) and np.all(pos[0] == 0) and np.all(acc == 0)
    seg = [np.linalg.norm(vel[40 * i]) * 41 * 0.02 for i in range(8)]  # |next - prev| per segment
    assert abs(seg[0] - seg[1]) < 1e-5 and 1.0 <= min(seg) and max(seg) <= 1.5
    assert len({round(float(x), 5) for x in seg[1:]}) == 7


@pytest.mark.parametrize("disturb", ["none", "gaussian"])
def test_eval_env_protocol_matches_the_reference_source(disturb):
    """harness.eval_env(keyed=True) vs the reference's eval_env executed from its own source (generator section 11) with the
    reference's RandomController: PRNGKey(1), four reset keys, per-step split(rng, 4), env.step with auto-reset (the random
    policy leaves the |pos| <= 3 box), the extra split after each step -- per-episode mean err_pos must coincide."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_eval_env_random.npz"))[disturb]
    env = cm.Quad3D("tracking_zigzag", disturb_type=disturb)
    ctl, _ = cm.get_controller(env, "random")
    mean, std, per_ep = cm.eval_env(env, ctl, total_steps=300 * 4, num_trajs=4, seed=1, keyed=True)
    assert per_ep.shape == g.shape == (4,)
    assert per_ep.mean() > 0.5  # the episodes do crash and get reset: the auto-reset path is part of what is compared
    assert np.abs(per_ep - g).max() < 1e-5, (per_ep, g)  # observed 1e-7
    assert abs(mean - g.mean()) < 1e-5 and abs(std - g.std()) < 1e-5


def test_render_env_matches_the_reference_source(tmp_path):
    """harness.render_env vs the reference's render_env executed from its own source (generator section 12): same key
    schedule, same states step for step, same pickle layout (list of per-step state dicts, envs/quadrotor.py:655-666)."""
    import pickle

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_render_env_random.npz"))
    env = cm.Quad3D("tracking_zigzag", disturb_type="gaussian")
    ctl, cp = cm.get_controller(env, "random")
    seq, rewards = cm.render_env(env, ctl, cp, repeat_times=1, filename="random", results_dir=str(tmp_path))
    assert len(seq) == int(g["n_steps"]) == len(rewards)
    for k in ("pos", "vel", "quat", "omega", "f_disturb", "pos_tar", "vel_tar"):
        mine = np.stack([np.asarray(d[k], np.float32) for d in seq])
        assert np.abs(mine - g[k]).max() < 1e-4 * max(1.0, np.abs(g[k]).max()), k
    assert [int(d["time"]) for d in seq] == g["time"].tolist()
    with open(os.path.join(str(tmp_path), "state_seq_random.pkl"), "rb") as f:
        stored = pickle.load(f)
    assert isinstance(stored, list) and len(stored) == len(seq) and isinstance(stored[0], dict)
    # every field of the reference's dict that scripts/vis.py (:70-95) reads is there under the same name
    assert {"pos", "quat", "pos_tar", "f_disturb", "pos_traj", "vel", "omega", "time"} <= set(stored[0].keys())
    assert {"pos", "quat", "pos_tar", "f_disturb", "pos_traj", "vel", "omega", "time"} <= set(g["keys"].tolist())


def test_offline_schedule_disturbance_follows_the_reference_key_schedule():
    """reset_a_cov_offline under disturb_type gaussian (controllers/covo.py:77-99): per schedule step two splits of the carried key,
    the second one's first half goes to step_env.  The controller hands the device the normals that chain produces; here the same chain
    is walked with the key-driven host environment (pinned to the reference's own step_env above)."""
    from covo_mpc_b200.controllers import offline_disturbance_normals

    env = cm.Quad3D("tracking_zigzag", disturb_type="gaussian")
    key = jr.PRNGKey(5)
    _, _, st = env.reset(jr.PRNGKey(1))
    T = 6
    z = offline_disturbance_normals(key, T)
    k = key
    act = np.array([0.1, 0.0, 0.0, 0.0], np.float32)
    for t in range(T):
        _, k = jr.split(k)       # rng_step for the expansion controller (covo.py:81)
        rs, k = jr.split(k)      # rng_step for step_env (covo.py:87)
        _, st, _, _, _ = env.step_env(rs, st, act, env.default_params)
        assert np.abs(st.f_disturb - np.float32(env.default_params.dyn_noise_scale) * z[t]).max() < 1e-7
    assert np.abs(z).max() > 0.1 and np.abs(z.mean()) < 1.0
