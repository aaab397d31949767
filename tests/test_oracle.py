"""Pins the oracle itself: analytic invariants the reference's maths guarantees, independent
finite differences for the exact Hessian, Random123 known-answer vectors for the counter RNG.
The reference ships no tests or golden vectors, so these are what stands behind `oracle/`."""
import numpy as np
import pytest

from oracle import oracle_np as o
from tests.util import scenario


def test_hover_equilibrium_and_euler_order():
    p = o.EnvParams()
    z = np.zeros((10, 3))
    s = o.make_state([0.1, 0.2, 0.3], [0, 0, 0, 1.0], [1.0, 0, 0], [0, 0, 0], [0, 0, 0], 0, z, z)
    hover = [(p.m * p.g / p.max_thrust) * 2 - 1, 0, 0, 0]
    s1 = o.step_env(s, hover, p)
    # explicit Euler with the OLD velocity (dynamics/free.py:102); hover thrust cancels gravity exactly
    assert np.allclose(s1.pos, [0.1 + 1.0 * p.dt, 0.2, 0.3])
    assert np.allclose(s1.vel, [1.0, 0, 0], atol=1e-12)
    assert np.allclose(s1.quat, [0, 0, 0, 1])
    assert s1.time == 1


def test_quaternion_stays_normalised_and_bodyrate_filter():
    p = o.EnvParams()
    z = np.zeros((10, 3))
    s = o.make_state([0, 0, 0], [0.02, -0.01, 0.03, 0.97], [0, 0, 0], [1.0, -2.0, 0.5], [0, 0, 0], 0, z, z)
    s1 = o.step_env(s, [0.0, 1.0, -1.0, 0.5], p)
    assert abs(sum(x * x for x in s1.quat) - 1) < 1e-12
    exp = 0.5 * np.array([1.0, -2.0, 0.5]) + 0.5 * np.array([10.0, -10.0, 1.5])  # alpha w + (1-alpha) a*max_omega
    assert np.allclose(s1.omega, exp)


def test_clamped_gather_and_termination():
    p = o.EnvParams()
    traj = np.arange(30, dtype=float).reshape(10, 3)
    s = o.make_state([0, 0, 0], [0, 0, 0, 1.0], [0, 0, 0], [0, 0, 0], [0, 0, 0], 8, traj, traj)
    s1 = o.step_env(s, [0, 0, 0, 0], p)  # time 9 -> last row
    s2 = o.step_env(s1, [0, 0, 0, 0], p)  # time 10 -> clamps to last row
    assert np.allclose(s1.pos_tar, traj[9]) and np.allclose(s2.pos_tar, traj[9])
    s3 = o.make_state([0, 3.1, 0], [0, 0, 0, 1.0], [0, 0, 0], [0, 0, 0], [0, 0, 0], 0, traj, traj)
    assert bool(o.is_terminal(s3, p)) and not bool(o.is_terminal(s, p))
    s4 = o.make_state([0, 0, 0], [0, 0, 0, 1.0], [0, 0, 0], [0, 0, 0], [0, 0, 0], 300, traj, traj)
    assert bool(o.is_terminal(s4, p))


def test_reward_freeze_after_done():
    """covo.py:233: after termination the reward repeats the last pre-termination value."""
    p, ns, a_mean, rng = scenario(H=6, time=297)
    a = np.repeat(a_mean[None], 3, axis=0)
    cost = o.rollout_costs(ns, a, p)
    # manual: rewards of steps with time 297..299, step at time 300 is `done` but its reward is still its own;
    # later ones are frozen to it
    s = o.broadcast_state(ns, 1, np.float32)
    rs = []
    for h in range(6):
        rs.append(float(o.tracking_penyaw_reward(s)[0]))
        s = o.step_env(s, [np.full(1, a_mean[h, k], np.float32) for k in range(4)], p)
    expect = -(rs[0] + rs[1] + rs[2] + rs[3] + rs[3] + rs[3])
    assert np.allclose(cost, expect, rtol=1e-5)


def test_softmax_invariants():
    p, ns, a_mean, rng = scenario(H=8)
    N = 64
    L = np.linalg.cholesky(0.25 * np.eye(32)).astype(np.float32)
    eps = rng.standard_normal((N, 32)).astype(np.float32)
    a = o.sample_actions(a_mean, L, eps)
    assert a.min() >= -1 and a.max() <= 1
    cost = o.rollout_costs(ns, a, p)
    new, w = o.softmax_update(a_mean, a, cost, 0.01)
    assert abs(w.sum() - 1) < 1e-5
    # N = 1: the update returns the clipped sample; Sigma -> 0: returns clip(mu)
    new1, _ = o.softmax_update(a_mean, a[:1], cost[:1], 0.01)
    assert np.array_equal(new1, a[0])
    a0 = o.sample_actions(a_mean, 0 * L, eps)
    assert np.allclose(a0, np.clip(a_mean, -1, 1)[None])
    # shard merge == global softmax, also when exp(-cost/lam) would overflow without the min shift (fact 4)
    parts = [o.softmax_partials(a[i::4].astype(np.float64), cost[i::4].astype(np.float64), 0.01) for i in range(4)]
    M, S, V = o.merge_partials(parts, 0.01)
    assert np.allclose(V / S, new, atol=1e-5)
    assert -cost.min() / 0.01 > 88.0  # exp() of the un-shifted exponent overflows float32


def test_hessian_jets_vs_finite_differences():
    p, ns, a_mean, rng = scenario(H=6, warm_steps=5, zero_disturb=False, dtype=np.float64)
    R = o.get_hessian(ns, a_mean.astype(np.float64), p)
    Rfd = o.hessian_fd(ns, a_mean.astype(np.float64), p, 1e-4)
    assert np.abs(R - Rfd).max() < 1e-6 * max(1.0, np.abs(R).max())
    assert np.abs(R - R.T).max() < 1e-14
    assert np.abs(R[-4:]).max() == 0.0  # last action never reaches a costed state


def test_hessian_spectrum_matches_survey_probe():
    """SURVEY App. C probed an independent throw-away implementation: lam in [-0.824, 42.38], 59 tiny."""
    p = o.EnvParams()
    rng = np.random.default_rng(0)
    s = o.reset_env("tracking_zigzag", p, rng, dtype=np.float64, zero_disturb=True)
    ns = o.noisy_state(s, p, rng)
    R = o.get_hessian(ns, o.hover_mean(50, p, np.float64), p)
    w = np.linalg.eigvalsh(R)
    assert -0.9 < w.min() < -0.7 and 40 < w.max() < 45
    assert (np.abs(w) < 1e-3).sum() >= 50
    S = o.optimize_sigma(R, 0.5, np.float64)
    ws = np.linalg.eigvalsh(S)
    assert 0.03 < ws.min() < 0.04 and 2.2 < ws.max() < 2.4


def test_clip_tie_subgradient_convention():
    """jnp.clip = minimum(maximum()) with balanced ties: 0.5 per clip on the bound, two clips in step_env."""
    p, ns, a_mean, rng = scenario(H=3, dtype=np.float64)
    am = a_mean.astype(np.float64).copy()
    am[0, 1] = 1.0
    R1 = o.get_hessian(ns, am, p)
    am2 = am.copy()
    am2[0, 1] = 0.999999
    R2 = o.get_hessian(ns, am2, p)
    # d/du scales by 0.25 on the tie -> the (u01, u01) entry by 0.0625, mixed entries by 0.25
    i = 1
    assert np.isclose(R1[i, i], 0.0625 * R2[i, i], rtol=1e-3, atol=1e-12)
    j = 4
    assert np.isclose(R1[i, j], 0.25 * R2[i, j], rtol=1e-3, atol=1e-12)
    am3 = am.copy()
    am3[0, 1] = 1.5  # outside: no influence at all
    assert np.abs(o.get_hessian(ns, am3, p)[i]).max() == 0.0


def test_optimize_sigma_invariants():
    rng = np.random.default_rng(3)
    n = 24
    A = rng.standard_normal((n, n))
    R = A + A.T
    R[-4:, :] = 0
    R[:, -4:] = 0
    S = o.optimize_sigma(R, 0.5, np.float64)
    assert np.abs(S - S.T).max() == 0
    assert np.isclose(np.linalg.slogdet(S)[1], 2 * n * np.log(0.5))  # det Sigma = sigma^(2n)
    assert np.abs(S @ R - R @ S).max() < 1e-10  # same eigenvectors
    w = np.linalg.eigvalsh(R)
    ws = np.linalg.eigvalsh(S)
    c = np.exp(0.5 * (4 * n * np.log(0.5) + np.log(w - w.min() + 1e-2).sum()) / n)
    assert np.allclose(np.sort(ws), np.sort(c / np.sqrt(w - w.min() + 1e-2)))
    L = o.cholesky_lower(S)
    assert np.allclose(L @ L.T, S) and np.allclose(L, np.tril(L))


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, exp in kat:
        out = o.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(x) for x in out) == exp


def test_philox_normals_moments():
    z = o.philox_normals(seed=7, stream=3, n_samples=4096, n_cols=64)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.05
    # N-sharding invariance: the field depends on the GLOBAL sample index only
    z2 = o.philox_normals(seed=7, stream=3, n_samples=1024, n_cols=64, sample_offset=1024)
    assert np.array_equal(z2, z[1024:2048])


def test_zigzag_generator_shape_and_velocity_quirk():
    pos, vel, acc = o.generate_zigzag_traj(300, 0.02, np.random.default_rng(0))
    assert pos.shape == (320, 3) and np.allclose(pos[0], 0)
    seg = (pos[40] - pos[0])
    assert np.allclose(vel[0], seg / 41 / 0.02)  # (next - prev)/41/dt, dynamics/utils.py:231-236
    assert np.allclose(acc, 0)


def test_pid_matches_small_angle_limit():
    p = o.EnvParams()
    z = np.zeros((10, 3))
    s = o.make_state([0.0, 0.0, 0.0], [0, 0, 0, 1.0], [0, 0, 0], [0, 0, 0], [0, 0, 0], 0, z + [0.1, 0, 0], z)
    a = o.pid_action(s, p)
    assert abs(a[0] - ((p.m * p.g / p.max_thrust) * 2 - 1)) < 1e-6 and abs(a[3]) < 1e-9  # hover thrust, no yaw demand
    assert a[2] > 0  # pitch towards +x target
