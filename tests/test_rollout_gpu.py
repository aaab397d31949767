"""K1/K2 parity (T1 in SURVEY 7.3): sample -> rollout -> softmax update through the C-ABI against the
oracle on identical (state, mean, factor, eps).  Tolerances: per-sample cost 1e-5 relative, updated mean
1e-5 absolute."""
import numpy as np
import pytest

from oracle import oracle_np as o
from tests.util import scenario

pytestmark = pytest.mark.gpu


def _handle(mode, N, H, T, **kw):
    from covo_mpc_b200 import _lib

    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len = mode, N, H, T
    for k, v in kw.items():
        setattr(cfg, k, v)
    return _lib.Handle(cfg)


def _check(h, p, ns, a_mean, a_s_oracle, eps, shift, lam=0.01, fd=None, gamma=1.0):
    am_in = a_mean
    mean_used = o.shift_mean(a_mean) if shift else a_mean
    cost_o = o.rollout_costs(ns, a_s_oracle, p, f_disturb_seq=fd)
    new_o, w = o.softmax_update(mean_used, a_s_oracle, cost_o, lam, gamma)
    a_out, act, costs, samples = h.rollout(o.state_to_vec24(ns), [ns.time], am_in[None], shift=shift, eps=eps[None],
                                           fdist_seq=None if fd is None else fd[None], want_costs=True, want_samples=True)
    assert np.abs(samples[0] - a_s_oracle).max() < 2e-6
    assert np.abs(costs[0] - cost_o).max() <= 1e-5 * max(1.0, np.abs(cost_o).max())
    assert np.abs(a_out[0] - new_o).max() < 1e-5
    assert np.array_equal(act[0], a_out[0, 0])
    return costs[0], a_out[0]


def test_mppi_hover_config_c():
    """BASELINE config 1: MPPI, hovering, N=128, H=32."""
    from covo_mpc_b200 import _lib

    N, H = 128, 32
    p, ns, a_mean, rng = scenario("hovering", seed=3, H=H)
    a_mean = o.hover_mean(H, p)
    h = _handle(_lib.MODE_MPPI, N, H, ns.pos_traj.shape[0])
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    eps = rng.standard_normal((N, H, 4)).astype(np.float32)
    Lblk = np.tile(0.5 * np.eye(4, dtype=np.float32), (H, 1, 1))
    a_s = o.sample_actions_blockdiag(o.shift_mean(a_mean), Lblk, eps)
    _check(h, p, ns, a_mean, a_s, eps.reshape(N, 4 * H), shift=True)


def test_mppi_general_blocks_and_disturbance():
    from covo_mpc_b200 import _lib

    N, H = 200, 16  # ragged last tile (200 = 3*64 + 8)
    p, ns, a_mean, rng = scenario("tracking", seed=5, H=H, warm_steps=20, zero_disturb=False)
    h = _handle(_lib.MODE_MPPI, N, H, ns.pos_traj.shape[0])
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    A = rng.standard_normal((H, 4, 4)) * 0.2
    cov = (A @ A.transpose(0, 2, 1) + 0.05 * np.eye(4)).astype(np.float32)
    h.set_cov(cov[None])
    Lblk = np.linalg.cholesky(cov.astype(np.float64)).astype(np.float32)
    eps = rng.standard_normal((N, H, 4)).astype(np.float32)
    a_s = o.sample_actions_blockdiag(a_mean, Lblk, eps)
    fd = (0.05 * rng.standard_normal((H, 3))).astype(np.float32)  # mppi.py:74 under 'gaussian'
    _check(h, p, ns, a_mean, a_s, eps.reshape(N, 4 * H), shift=False, fd=fd)


@pytest.mark.parametrize("N,H,task", [(1024, 50, "tracking_zigzag"), (256, 32, "tracking"), (64, 8, "hovering")])
def test_covo_dense_factor(N, H, task):
    from covo_mpc_b200 import _lib

    p, ns, a_mean, rng = scenario(task, seed=11, H=H, warm_steps=10)
    n = 4 * H
    h = _handle(_lib.MODE_COVO_ONLINE, N, H, ns.pos_traj.shape[0])
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    A = rng.standard_normal((n, n)) / np.sqrt(n)
    cov = (0.2 * A @ A.T + 0.1 * np.eye(n)).astype(np.float32)
    h.set_cov(cov[None])
    cov_sym = h.get_cov()[0]
    L = np.linalg.cholesky(cov_sym.astype(np.float64)).astype(np.float32)
    eps = rng.standard_normal((N, n)).astype(np.float32)
    a_s = o.sample_actions(o.shift_mean(a_mean), L, eps)
    # samples that sit within float round-off of the clip bound may clip differently; compare via the device samples
    a_out, act, costs, samples = h.rollout(o.state_to_vec24(ns), [ns.time], a_mean[None], shift=True, eps=eps[None],
                                           want_costs=True, want_samples=True)
    assert np.abs(samples[0] - a_s).max() < 5e-6
    cost_o = o.rollout_costs(ns, samples[0], p)
    assert np.abs(costs[0] - cost_o).max() <= 1e-5 * max(1.0, np.abs(cost_o).max())
    new_o, _ = o.softmax_update(o.shift_mean(a_mean), samples[0], cost_o, 0.01)
    assert np.abs(a_out[0] - new_o).max() < 1e-5


def test_termination_freeze_and_clamped_reference():
    """time close to max_steps: rewards freeze (covo.py:233) and the target gather clamps (free.py:153-155)."""
    from covo_mpc_b200 import _lib

    N, H = 128, 12
    for time in (292, 299, 305, 330):
        p, ns, a_mean, rng = scenario("tracking_zigzag", seed=2, H=H, time=time)
        h = _handle(_lib.MODE_MPPI, N, H, ns.pos_traj.shape[0])
        h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
        eps = rng.standard_normal((N, H, 4)).astype(np.float32)
        Lblk = np.tile(0.5 * np.eye(4, dtype=np.float32), (H, 1, 1))
        a_s = o.sample_actions_blockdiag(a_mean, Lblk, eps)
        _check(h, p, ns, a_mean, a_s, eps.reshape(N, 4 * H), shift=False)
        h.close()


def test_out_of_bounds_position_terminates():
    from covo_mpc_b200 import _lib

    N, H = 64, 10
    p, ns, a_mean, rng = scenario("hovering", seed=4, H=H)
    ns.pos[2] = np.float32(2.99)
    ns.vel[2] = np.float32(2.0)  # crosses |z| > 3 after one step
    h = _handle(_lib.MODE_MPPI, N, H, ns.pos_traj.shape[0])
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    eps = rng.standard_normal((N, H, 4)).astype(np.float32)
    Lblk = np.tile(0.5 * np.eye(4, dtype=np.float32), (H, 1, 1))
    a_s = o.sample_actions_blockdiag(a_mean, Lblk, eps)
    _check(h, p, ns, a_mean, a_s, eps.reshape(N, 4 * H), shift=False)


def test_gamma_mean_and_discount():
    from covo_mpc_b200 import _lib

    N, H = 128, 8
    p, ns, a_mean, rng = scenario("tracking", seed=9, H=H, warm_steps=5)
    h = _handle(_lib.MODE_MPPI, N, H, ns.pos_traj.shape[0], gamma_mean=0.7, discount=0.9, lam=0.05)
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    eps = rng.standard_normal((N, H, 4)).astype(np.float32)
    Lblk = np.tile(0.5 * np.eye(4, dtype=np.float32), (H, 1, 1))
    a_s = o.sample_actions_blockdiag(a_mean, Lblk, eps)
    cost_o = o.rollout_costs(ns, a_s, p, discount=0.9)
    new_o, _ = o.softmax_update(a_mean, a_s, cost_o, 0.05, 0.7)
    a_out, act, costs, _ = h.rollout(o.state_to_vec24(ns), [ns.time], a_mean[None], eps=eps.reshape(1, N, 4 * H), want_costs=True)
    assert np.abs(costs[0] - cost_o).max() < 1e-5 * max(1, np.abs(cost_o).max())
    assert np.abs(a_out[0] - new_o).max() < 1e-5


def test_production_rng_field_and_sharding_invariance():
    """The in-kernel Gaussian field is a function of the GLOBAL sample index: 2 shards == 1 device."""
    from covo_mpc_b200 import _lib

    N, H = 512, 16
    p, ns, a_mean, rng = scenario("tracking", seed=1, H=H)
    h = _handle(_lib.MODE_MPPI, N, H, ns.pos_traj.shape[0], seed=1234)
    z = h.debug_eps(7)
    zo = o.philox_normals(1234, 7, N, 4 * H)
    # fast-intrinsic Box-Muller on the device vs float64 on the host: the same field up to ~1e-4 near z = 0
    assert np.abs(z - zo).max() < 1e-3 and np.abs(z - zo).mean() < 1e-5
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    parts = []
    for r in range(2):
        hr = _handle(_lib.MODE_MPPI, N, H, 300, seed=1234, rank=r, world=2)
        parts.append(hr.debug_eps(7))
    assert np.array_equal(np.concatenate(parts), z)
    # production-mode step == parity-mode step fed with the dumped field
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    sid = h.rng_step()
    a1, act1, c1, _ = h.rollout(o.state_to_vec24(ns), [ns.time], a_mean[None], want_costs=True)
    a2, act2, c2, _ = h.rollout(o.state_to_vec24(ns), [ns.time], a_mean[None], eps=h.debug_eps(sid)[None], want_costs=True)
    assert np.array_equal(c1, c2) and np.array_equal(a1, a2)


def test_pos_stats_info_dict():
    from covo_mpc_b200 import _lib

    N, H = 256, 10
    p, ns, a_mean, rng = scenario("tracking", seed=6, H=H, warm_steps=3)
    h = _handle(_lib.MODE_MPPI, N, H, ns.pos_traj.shape[0])
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    h.enable_pos_stats(True)
    eps = rng.standard_normal((N, H, 4)).astype(np.float32)
    Lblk = np.tile(0.5 * np.eye(4, dtype=np.float32), (H, 1, 1))
    a_s = o.sample_actions_blockdiag(a_mean, Lblk, eps)
    _, poses = o.rollout_costs(ns, a_s, p, return_pos=True)
    h.rollout(o.state_to_vec24(ns), [ns.time], a_mean[None], eps=eps.reshape(1, N, 4 * H))
    m, s = h.pos_stats()
    assert np.abs(m[0] - poses.mean(1)).max() < 1e-5
    assert np.abs(s[0] - poses.std(1)).max() < 2e-4


def test_full_size_properties():
    """BASELINE headline size (N=8192, H=50): size-independent properties instead of an element-wise oracle."""
    from covo_mpc_b200 import _lib

    N, H = 8192, 50
    n = 4 * H
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=0, H=H, warm_steps=30)
    h = _handle(_lib.MODE_COVO_ONLINE, N, H, ns.pos_traj.shape[0])
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    h.set_cov((0.25 * np.eye(n, dtype=np.float32))[None])
    eps = rng.standard_normal((N, n)).astype(np.float32)
    st = o.state_to_vec24(ns)
    a_out, act, costs, samples = h.rollout(st, [ns.time], a_mean[None], eps=eps[None], want_costs=True, want_samples=True)
    # (1) the update is the softmax-weighted mean of the device's own samples/costs (float64 re-reduction)
    c = costs[0].astype(np.float64)
    w = np.exp(-(c - c.min()) / 0.01)
    w /= w.sum()
    assert np.abs((w[:, None, None] * samples[0]).sum(0) - a_out[0]).max() < 2e-6
    # (2) permutation invariance of the reduction, (3) a spot check of 512 samples against the oracle
    perm = rng.permutation(N)
    a_out2, _, costs2, _ = h.rollout(st, [ns.time], a_mean[None], eps=eps[perm][None], want_costs=True)
    assert np.array_equal(costs2[0], costs[0][perm])
    assert np.abs(a_out2[0] - a_out[0]).max() < 2e-6
    idx = rng.choice(N, 512, replace=False)
    co = o.rollout_costs(ns, samples[0][idx], p)
    assert np.abs(co - costs[0][idx]).max() < 1e-5 * max(1, np.abs(co).max())
    # (4) Sigma -> 0 returns clip(mu); N-sample answers stay inside the clip box
    h.set_cov((1e-12 * np.eye(n, dtype=np.float32))[None])
    a0, _, _, _ = h.rollout(st, [ns.time], a_mean[None], eps=eps[None])
    assert np.abs(a0[0] - np.clip(a_mean, -1, 1)).max() < 1e-5
    assert a_out.min() >= -1 and a_out.max() <= 1


@pytest.mark.parametrize("mode_name", ["covo", "mppi"])
def test_jax_compatible_stream_in_kernel(mode_name):
    """covo_set_jax_key: the kernel's Threefry draws == the host twin of jax.random (covo_mpc_b200/jaxrng.py, pinned by the
    Random123 / JAX-documentation known answers in tests/test_jaxrng.py), for both samplers, with a ragged last tile, and
    independent of N-sharding."""
    from covo_mpc_b200 import _lib, jaxrng

    N, H = 160, 10  # 2.5 tiles of 64
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=4, H=H, warm_steps=3)
    mode = _lib.MODE_MPPI if mode_name == "mppi" else _lib.MODE_COVO_OFFLINE
    act_key = jaxrng.split(jaxrng.PRNGKey(2024))[1]
    eps = (jaxrng.mppi_normals(act_key, N, H).reshape(N, 4 * H) if mode_name == "mppi" else jaxrng.covo_normals(act_key, N, 4 * H))
    st = o.state_to_vec24(ns)

    def mk(**kw):
        h = _handle(mode, N, H, ns.pos_traj.shape[0], **kw)
        h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
        if mode_name == "covo":
            A = rng_cov.standard_normal((4 * H, 4 * H)) * 0.1
            h.set_cov_offline((A @ A.T + 0.2 * np.eye(4 * H)).astype(np.float32)[None])
        return h

    rng_cov = np.random.default_rng(0)
    h = mk()
    h.set_jax_key(act_key)
    a1, act1, c1, s1 = h.rollout(st, [ns.time], a_mean[None], want_costs=True, want_samples=True)
    a2, act2, c2, s2 = h.rollout(st, [ns.time], a_mean[None], eps=eps[None], want_costs=True, want_samples=True)
    assert np.abs(s1 - s2).max() < 5e-6  # logf / sqrtf vs NumPy: last-ulp differences only
    assert np.abs(c1 - c2).max() < 1e-4 and np.abs(a1 - a2).max() < 1e-4
    # the key is one-shot: the next call is back on the Philox field
    a3, _, c3, _ = h.rollout(st, [ns.time], a_mean[None], want_costs=True)
    assert np.abs(c3 - c1).max() > 1e-3
    # N-sharding: rank r draws rows [r N/2, (r+1) N/2) of the same stream
    for r in range(2):
        rng_cov = np.random.default_rng(0)
        hr = mk(rank=r, world=2)
        hr.set_jax_key(act_key)
        _, _, cr, sr = hr.rollout(st, [ns.time], a_mean[None], want_costs=True, want_samples=True)
        assert np.array_equal(sr[0], s1[0][r * N // 2:(r + 1) * N // 2])


@pytest.mark.parametrize("mode_name", ["mppi", "covo"])
@pytest.mark.parametrize("N,H", [(4, 2), (4, 3), (8192, 2), (8192, 3), (65536, 2)])
def test_tiny_horizons_merge_scratch(mode_name, N, H):
    """get_controller(..., debug=True) forces N = 4, H = 2 (envs/quadrotor.py:705-707, 726-728).  The grid-wide merge keeps its
    scale table and partial sums in the dead part of the dynamic shared memory, which at H = 2 is far smaller than the layout
    tuned for n = 200 (out-of-bounds shared-memory write).  gamma_mean != 1 reads the staged mean AFTER the merge
    scratch was written.  N = 65536 at H = 2: the scale table no longer fits and is recomputed from the record headers."""
    from covo_mpc_b200 import _lib

    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=17, H=H, warm_steps=5)
    n = 4 * H
    mode = _lib.MODE_MPPI if mode_name == "mppi" else _lib.MODE_COVO_ONLINE
    h = _handle(mode, N, H, ns.pos_traj.shape[0], gamma_mean=0.7)
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    if mode_name == "mppi":
        eps = rng.standard_normal((N, H, 4)).astype(np.float32)
        Lblk = np.tile(0.5 * np.eye(4, dtype=np.float32), (H, 1, 1))
        a_s = o.sample_actions_blockdiag(o.shift_mean(a_mean), Lblk, eps)
    else:
        A = rng.standard_normal((n, n)) / np.sqrt(n)
        cov = (0.2 * A @ A.T + 0.1 * np.eye(n)).astype(np.float32)
        h.set_cov(cov[None])
        L = np.linalg.cholesky(h.get_cov()[0].astype(np.float64)).astype(np.float32)
        eps = rng.standard_normal((N, n)).astype(np.float32)
        a_s = o.sample_actions(o.shift_mean(a_mean), L, eps)
    a_out, act, costs, samples = h.rollout(o.state_to_vec24(ns), [ns.time], a_mean[None], shift=True, eps=eps.reshape(1, N, n),
                                           want_costs=True, want_samples=True)
    assert np.abs(samples[0] - a_s).max() < 5e-6
    cost_o = o.rollout_costs(ns, samples[0], p)
    assert np.abs(costs[0] - cost_o).max() <= 1e-5 * max(1.0, np.abs(cost_o).max())
    new_o, _ = o.softmax_update(o.shift_mean(a_mean), samples[0], cost_o, 0.01, 0.7)
    assert np.abs(a_out[0] - new_o).max() < 2e-5
    h.close()


@pytest.mark.parametrize("gamma_sigma,lam", [(0.3, 0.5), (0.1, 0.01)])
def test_mppi_covariance_update(gamma_sigma, lam):
    """MPPI with gamma_sigma != 0 (controllers/mppi.py:119-125): the per-step 4 x 4 covariances follow the weighted second moments of
    the clipped samples about the UPDATED mean, blended with the shifted covariance the samples were drawn from; the next step samples
    from their Cholesky factors.  Three consecutive steps against the oracle on identical eps (lam = 0.5: many samples carry weight;
    lam = 0.01 with a small gamma: the arg-min regime, the blend keeps the covariance positive definite)."""
    from covo_mpc_b200 import _lib

    N, H = 512, 12
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=9, H=H, warm_steps=10)
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.lam = _lib.MODE_MPPI, N, H, ns.pos_traj.shape[0], lam
    cfg.gamma_sigma = gamma_sigma
    h = _lib.Handle(cfg)
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    h.set_mean(a_mean[None])
    cov = np.tile(0.25 * np.eye(4, dtype=np.float32), (H, 1, 1))
    mean = a_mean.copy()
    for step in range(3):
        eps = rng.standard_normal((N, H, 4)).astype(np.float32)
        act = h.step(o.state_to_vec24(ns), [ns.time], eps.reshape(1, N, 4 * H))[0]
        u_o, mean, cov, _ = o.mppi_call(ns, mean, cov, eps, p, lam=lam, gamma_sigma=gamma_sigma)
        cov_d = h.get_cov()[0].reshape(H, 4, 4)
        tol = 2e-5 if lam > 0.1 else 1e-4  # lam = 0.01: d w / d cost = w (1 - w) / lam amplifies the float32 cost rounding (carried over the steps)
        assert np.abs(act - u_o).max() < tol, step
        assert np.abs(h.get_mean()[0].reshape(H, 4) - mean).max() < tol, step
        assert np.abs(cov_d - cov).max() < tol * max(1.0, np.abs(cov).max()), step
        assert np.array_equal(cov_d, cov_d.transpose(0, 2, 1))
        assert (h.status() == 0).all()
    assert np.abs(cov - 0.25 * np.eye(4)).max() > 1e-3  # the update did something
    h.close()
