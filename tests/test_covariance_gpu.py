"""K3-K5 parity (T2 in SURVEY 7.3): exact Hessian, optimize_sigma and Cholesky through the C-ABI."""
import numpy as np
import pytest

from oracle import oracle_np as o
from tests.util import scenario

pytestmark = pytest.mark.gpu


def _handle(N, H, T, mode=None, **kw):
    from covo_mpc_b200 import _lib

    cfg = _lib.default_config()
    cfg.mode = _lib.MODE_COVO_ONLINE if mode is None else mode
    cfg.n_samples, cfg.horizon, cfg.traj_len = N, H, T
    for k, v in kw.items():
        setattr(cfg, k, v)
    return _lib.Handle(cfg)


@pytest.mark.parametrize("H,task,warm,time", [(50, "tracking_zigzag", 0, 0), (50, "tracking_zigzag", 40, 0),
                                               (32, "tracking", 25, 0), (12, "hovering", 5, 0), (20, "tracking_zigzag", 10, 290)])
def test_hessian_vs_forward_over_forward_oracle(H, task, warm, time):
    p, ns, a_mean, rng = scenario(task, seed=H + warm, H=H, warm_steps=warm, zero_disturb=False, time=time)
    h = _handle(64, H, ns.pos_traj.shape[0])
    h.set_reference(ns.pos_traj[None], ns.vel_traj[None])
    am = a_mean.copy()
    am[1, 2] = 1.0  # clip tie
    am[2, 0] = -1.2  # outside the clip box
    R = h.hessian(o.state_to_vec24(ns), [ns.time], am[None], shift=False)[0]
    Ro = o.get_hessian(ns, am, p, dtype=np.float64)
    scale = max(1.0, np.abs(Ro).max())
    assert np.abs(R - Ro).max() < 2e-5 * scale  # R vs float64 oracle <= 1e-5 ||R||
    assert np.abs(R - R.T).max() == 0.0
    assert np.abs(R[-4:]).max() == 0.0 and np.abs(R[8]).max() == 0.0  # last action; the clipped-out control (h=2, c=0)
    # the shift operator fused into the load
    Rs = h.hessian(o.state_to_vec24(ns), [ns.time], am[None], shift=True)[0]
    Rso = o.get_hessian(ns, o.shift_mean(am), p, dtype=np.float64)
    assert np.abs(Rs - Rso).max() < 2e-5 * max(1.0, np.abs(Rso).max())


@pytest.mark.parametrize("H", [50, 8])
def test_hessian_large_batch_equals_single_environments(H):
    """Batches of more than 18 environments run the forward chains 40 to a CTA (one staged copy of the records per 40 chains) instead of
    2: the same per-chain arithmetic, so every environment's Hessian must equal bit for bit what a single-environment handle delivers."""
    E = 20
    hs = _handle(64, H, 320)
    hb = _handle(64, H, 320, n_env=E)
    states, times, means, pos, vel = [], [], [], [], []
    for e in range(E):
        p, ns, a_mean, rng = scenario("tracking_zigzag", seed=100 + e, H=H, warm_steps=5 + e, zero_disturb=False)
        states.append(o.state_to_vec24(ns)); times.append(ns.time); means.append(a_mean); pos.append(ns.pos_traj); vel.append(ns.vel_traj)
    hb.set_reference(np.stack(pos), np.stack(vel))
    Rb = hb.hessian(np.stack(states), times, np.stack(means), shift=True)
    for e in (0, 7, 19):
        hs.set_reference(pos[e][None], vel[e][None])
        R1 = hs.hessian(states[e], [times[e]], means[e][None], shift=True)[0]
        assert np.array_equal(R1, Rb[e]), e
    hs.close()
    hb.close()


DENSE_TOL = 1e-5  # the dense path (float64 pole inverses since round 2) is held to the tolerance of E1-E3


@pytest.mark.parametrize("path", ["default", "tridiag", "dense"])
@pytest.mark.parametrize("H", [50, 32, 8, 3])
def test_optimize_sigma_and_cholesky(monkeypatch, H, path):
    """optimize_sigma on both kernel paths, held to the same tolerance: the dense one (D1-D3: adaptive Lanczos + one float64 Gauss-Jordan
    inverse per pole; the default of a single-environment handle) and the tridiagonal one (E1-E3; COVO_SIGMA=tridiag, batches)."""
    monkeypatch.delenv("COVO_SIGMA", raising=False)
    if path != "default":
        monkeypatch.setenv("COVO_SIGMA", path)
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=7, H=H, warm_steps=15)
    n = 4 * H
    h = _handle(64, H, ns.pos_traj.shape[0])
    assert h.sigma_path() == (0 if path == "tridiag" else 3)
    R = o.get_hessian(ns, a_mean, p, dtype=np.float64).astype(np.float32)
    S = h.optimize_sigma(R[None])[0]
    So = o.optimize_sigma(R.astype(np.float64), 0.5, np.float64)
    tol = DENSE_TOL if path != "tridiag" else 1e-5
    assert np.linalg.norm(S - So) / np.linalg.norm(So) < tol
    assert np.abs(S - So).max() < tol * np.abs(So).max()
    assert np.abs(S - S.T).max() == 0.0
    assert abs(np.linalg.slogdet(S.astype(np.float64))[1] - 2 * n * np.log(0.5)) < 1e-3  # det Sigma = sigma^(2n)
    if h.sigma_path() == 0:
        d, e, sc = h.debug_tridiag()
        T = np.diag(d) + np.diag(e[:-1], 1) + np.diag(e[:-1], -1)
        w = np.linalg.eigvalsh(((R + R.T) / 2).astype(np.float64))
        assert np.abs(np.linalg.eigvalsh(T) - w).max() < 2e-5 * max(1, np.abs(w).max())  # Householder is a similarity
        assert abs(sc[0] - np.linalg.eigvalsh(T)[0]) < 1e-10  # fp64 Sturm multisection
    L = h.cholesky(S[None])[0]
    Lo = np.linalg.cholesky(S.astype(np.float64))
    assert np.abs(L - Lo).max() < 2e-6 * max(1, np.abs(Lo).max())
    assert np.abs(np.triu(L, 1)).max() == 0.0


@pytest.mark.parametrize("path", ["tridiag", "dense"])
def test_sigma_random_symmetric_and_degenerate(monkeypatch, path):
    """Arbitrary symmetric input (not a CoVO Hessian: no separated lowest eigenvalue).  The tridiagonal path handles it directly; the
    dense path either converges (to its own accuracy) or detects that its Lanczos stage has not (status 3), in which case
    covo_optimize_sigma redoes the matrix on the tridiagonal path."""
    monkeypatch.setenv("COVO_SIGMA", path)
    rng = np.random.default_rng(0)
    H = 16
    n = 4 * H
    h = _handle(64, H, 300)
    for kind in range(4):
        A = rng.standard_normal((n, n)).astype(np.float32)
        R = (A + A.T) * (0.1 if kind == 0 else 3.0)
        if kind == 2:
            R[-8:, :] = 0
            R[:, -8:] = 0  # exact zero block -> the tridiagonal splits
        if kind == 3:
            R = np.diag(rng.uniform(-1, 5, n)).astype(np.float32)  # already diagonal: every reflector is trivial
        S = h.optimize_sigma(R[None])[0]
        So = o.optimize_sigma(R.astype(np.float64), 0.5, np.float64)
        assert np.linalg.norm(S - So) / np.linalg.norm(So) < 2e-5, kind
        assert h.status()[0] == 0


def test_sigma_batched_envs():
    from covo_mpc_b200 import _lib

    rng = np.random.default_rng(1)
    H, E = 8, 5
    n = 4 * H
    h = _handle(64, H, 300, n_env=E)
    A = rng.standard_normal((E, n, n)).astype(np.float32)
    R = A + A.transpose(0, 2, 1)
    S = h.optimize_sigma(R)
    for e in range(E):
        So = o.optimize_sigma(R[e].astype(np.float64), 0.5, np.float64)
        assert np.linalg.norm(S[e] - So) / np.linalg.norm(So) < 2e-5


@pytest.mark.parametrize("H,E", [(50, 20), (8, 40), (3, 24), (9, 19)])
@pytest.mark.parametrize("sandwich", ["tensor", "simt"])
def test_sigma_large_batches_tensor_core_sandwich(monkeypatch, H, E, sandwich):
    """Batches of more than 18 matrices form Sigma = Q F Q^T on the tensor cores (mma.sync TF32, operands split 3 x TF32: float32-grade
    products); COVO_SANDWICH=simt keeps the FFMA kernel.  Both against the float64 eigen-decomposition, exactly symmetric, n = 200, 32,
    the tiny 12 and the ragged 36."""
    if sandwich == "simt":
        monkeypatch.setenv("COVO_SANDWICH", "simt")
    else:
        monkeypatch.delenv("COVO_SANDWICH", raising=False)
    rng = np.random.default_rng(11)
    n = 4 * H
    h = _handle(64, H, 300, n_env=E)
    if H == 50:
        p, ns, a_mean, _ = scenario("tracking_zigzag", seed=3, H=50, warm_steps=20)
        R0 = o.get_hessian(ns, a_mean, p, dtype=np.float64)
        R = np.stack([(R0 * (1.0 + 0.05 * e) + 0.3 * e * np.eye(n)).astype(np.float32) for e in range(E)])  # CoVO Hessians, scaled / shifted per env
    else:
        A = rng.standard_normal((E, n, n)).astype(np.float32)
        R = A + A.transpose(0, 2, 1)
    S = h.optimize_sigma(R)
    assert (h.status() == 0).all()
    for e in range(0, E, 3):
        So = o.optimize_sigma(R[e].astype(np.float64), 0.5, np.float64)
        assert np.linalg.norm(S[e] - So) / np.linalg.norm(So) < 2e-5, e
        assert np.abs(S[e] - S[e].T).max() == 0.0
    h.close()


@pytest.mark.parametrize("nc", [1, 2, 4, 8])
def test_sigma_cluster_widths_agree(nc, monkeypatch):
    """The tridiagonalisation spreads one matrix over 1, 2, 4 or 8 CTAs of a cluster (batched environments use the
    narrow ones): every width must deliver the same covariance at the headline size n = 200."""
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=3, H=50, warm_steps=20)
    R = o.get_hessian(ns, a_mean, p, dtype=np.float64).astype(np.float32)
    So = o.optimize_sigma(R.astype(np.float64), 0.5, np.float64)
    monkeypatch.setenv("COVO_E1_CLUSTER", str(nc))
    monkeypatch.setenv("COVO_SIGMA", "tridiag")
    h = _handle(64, 50, ns.pos_traj.shape[0])
    S = h.optimize_sigma(R[None])[0]
    assert np.linalg.norm(S - So) / np.linalg.norm(So) < 1e-5
    assert np.abs(S - S.T).max() == 0.0
    d, e, sc = h.debug_tridiag()
    T = np.diag(d) + np.diag(e[:-1], 1) + np.diag(e[:-1], -1)
    w = np.linalg.eigvalsh(((R + R.T) / 2).astype(np.float64))
    assert np.abs(np.linalg.eigvalsh(T) - w).max() < 2e-5 * max(1, np.abs(w).max())
    assert abs(sc[0] - np.linalg.eigvalsh(T)[0]) < 1e-9


def test_sigma_many_matrices_batched():
    """40 matrices at once (environment batch / offline schedule): the 2-CTA path of E1 and the 2-CTA path of E2."""
    from covo_mpc_b200 import _lib

    rng = np.random.default_rng(5)
    H, E = 20, 40
    n = 4 * H
    h = _handle(64, H, 300, n_env=E)
    A = rng.standard_normal((E, n, n)).astype(np.float32)
    R = A + A.transpose(0, 2, 1)
    S = h.optimize_sigma(R)
    L = h.cholesky(S)
    for e in (0, 7, 39):
        So = o.optimize_sigma(R[e].astype(np.float64), 0.5, np.float64)
        assert np.linalg.norm(S[e] - So) / np.linalg.norm(So) < 2e-5
        Lo = np.linalg.cholesky(S[e].astype(np.float64))
        assert np.abs(L[e] - Lo).max() < 2e-6 * max(1, np.abs(Lo).max())
    assert (h.status() == 0).all()


def test_maximum_horizon_full_step():
    """H = 56 (n = 224, the shared-memory / register limit of the covariance kernels): one full CoVO-online call."""
    import covo_mpc_b200 as cm
    from tests.test_step_gpu import _to_env_state

    N, H = 256, 56
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=77, H=H, warm_steps=10)
    env = cm.Quad3D("tracking_zigzag")
    ctl, cp = cm.get_controller(env, "covo-online", f"N{N}_H{H}_lam0.01")
    cp = cp.replace(a_mean=a_mean)
    eps = rng.standard_normal((N, 4 * H)).astype(np.float32)
    st = _to_env_state(cm, ns)
    action, cp2, _ = ctl(None, st, env.default_params, eps, cp, {"noisy_state": st})
    u_o, mean_o, cov_o, _ = o.covo_call(ns, a_mean, eps, p, lam=0.01)
    cov = np.asarray(cp2.a_cov)
    assert np.linalg.norm(cov - cov_o) / np.linalg.norm(cov_o) < 3e-5
    assert np.abs(action - u_o).max() < 5e-4
    ctl.close()


def test_dense_sigma_on_hessians_that_need_more_than_24_lanczos_steps():
    """tests/golden/hessians/hard_hessians_n200.npz: closed-loop Hessians on which a fixed 24-step Lanczos iteration leaves the smallest Ritz
    value 1e-6 .. 6e-2 above lambda_min (the last one would make A = R - lambda_min + 1e-2 indefinite).  The adaptive kernel must
    converge on all of them (no status, no fall-back to the tridiagonal path) and give the Sigma of the float64 eigen-decomposition."""
    import os

    from covo_mpc_b200 import _lib

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hessians", "hard_hessians_n200.npz"))
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len = _lib.MODE_COVO_ONLINE, 64, 50, 320
    h = _lib.Handle(cfg)
    h.set_sigma_path(3)
    for R in g["R"]:
        S = h.optimize_sigma(R[None])[0]
        assert h.sigma_path() == 3 and int(h.status()[0]) == 0
        S_ref = o.optimize_sigma(R.astype(np.float64), 0.5, np.float64)
        S_f32 = o.optimize_sigma(R, 0.5, np.float32)  # float32 LAPACK eigh: the reference's arithmetic
        err, err_f32 = np.linalg.norm(S - S_ref) / np.linalg.norm(S_ref), np.linalg.norm(S_f32 - S_ref) / np.linalg.norm(S_ref)
        print(f"dense Sigma vs float64: {err:.2e} (float32 eigh: {err_f32:.2e})")
        assert err < DENSE_TOL and np.abs(S - S.T).max() == 0.0
        assert np.linalg.eigvalsh(S.astype(np.float64))[0] > 0
    h.close()
