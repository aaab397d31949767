"""Reference-executed fixtures at the BASELINE sizes: the UNMODIFIED /root/reference/quadjax source, run under the
NumPy shim of make_reference_golden.py, for

  * CoVOController.optimize_sigma at n = 200 (H = 50)                          -> reference_fullsize_optimize_sigma_n200.npz
  * one MPPIController.__call__ at N = 8192, H = 50                            -> reference_fullsize_mppi_N8192_H50.npz
  * one CoVOController.__call__ in OFFLINE mode at N = 8192, H = 50 (the table entry it looks up is supplied: the schedule itself
    is covered at H = 6 by reference_covo_offline_schedule.npz)                -> reference_fullsize_covo_offline_N8192_H50.npz

The N x 4H Gaussian draws (6.5 MB) are NOT stored: the shim's jax.random draws from numpy default_rng(seed) in call order, and the
sampling of a controller call happens before anything else draws, so the tests regenerate them as
default_rng(seed).standard_normal((N, 4H)) (asserted here against the logged draws).  Stored: inputs, the updated mean, the action, the
arg-min sample, and the 64 smallest / first 256 per-sample costs recomputed through the reference's own step_env.

Runs only where /root/reference exists (this container); ~10 minutes of shim time.  TEST INFRASTRUCTURE.

    python tests/golden/make_reference_golden_fullsize.py"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_reference_golden as g  # noqa: E402  (the shim)

F = np.float32
N, H, LAM = 8192, 50, 0.01


def main():
    rand = g.RandomLog(123)
    g.install_shim(rand)
    sys.path.insert(0, g.REF)
    from quadjax.controllers.covo import CoVOController  # noqa: F401
    from quadjax.dynamics.dataclass import EnvState3D
    from quadjax.envs.quadrotor import Quad3D, get_controller

    from oracle import oracle_np as o
    from tests.util import scenario

    env = Quad3D(task="tracking_zigzag", obs_type="quad", lower_controller="base", enable_randomizer=False, disturb_type="none",
                 disable_rollover_terminate=True, generate_noisy_state=True)
    env.get_obs = lambda *a, **k: None
    params = env.default_params

    def to_ref_state(s):
        z3 = np.zeros(3, F)
        hist = np.zeros((4, 3), F)
        kw = dict(pos=np.array(s.pos, F), vel=np.array(s.vel, F), quat=np.array(s.quat, F), omega=np.array(s.omega, F), omega_tar=z3.copy(),
                  pos_traj=s.pos_traj.astype(F), vel_traj=s.vel_traj.astype(F), acc_traj=np.zeros_like(s.pos_traj, dtype=F),
                  pos_tar=np.array(s.pos_tar, F), vel_tar=np.array(s.vel_tar, F), acc_tar=z3.copy(), last_thrust=0.0, last_torque=z3.copy(),
                  time=int(s.time), f_disturb=np.array(s.f_disturb, F))
        fields = getattr(EnvState3D, "__dataclass_fields__", {})
        if "vel_hist" in fields:
            kw.update(vel_hist=hist.copy(), omega_hist=hist.copy(), action_hist=np.zeros((4, 4), F))
        return EnvState3D(**kw)

    def vec24(st):
        v = np.zeros(24, F)
        v[0:3], v[3:7], v[7:10], v[10:13], v[13:16], v[16:19], v[19:22] = st.pos, st.quat, st.vel, st.omega, st.f_disturb, st.pos_tar, st.vel_tar
        return v

    # ---- optimize_sigma at n = 200 ------------------------------------------------------------------------------------------
    t0 = time.time()
    pp, ns_, a_mean, _ = scenario("tracking_zigzag", seed=5, H=H, warm_steps=25)
    from oracle import oracle_c

    R = oracle_c.hessian_f64(ns_, o.shift_mean(a_mean), pp).astype(F)
    ctl0 = CoVOController.__new__(CoVOController)
    ctl0.action_dim, ctl0.H = 4, H
    S = np.asarray(CoVOController.optimize_sigma(ctl0, R, types.SimpleNamespace(sample_sigma=F(0.5))), F)
    np.savez_compressed(os.path.join(HERE, "reference_fullsize_optimize_sigma_n200.npz"), R=R, Sigma=S, H=H)
    print(f"optimize_sigma n=200 done [{time.time() - t0:.0f} s]", flush=True)

    def run_call(name, ctl, cp, seed, table_entry=None):
        t0 = time.time()
        rand.rng = np.random.default_rng(seed)  # fresh stream: the first N x 4H normals are the samples' draws
        rand.mvn.clear()
        pp, ns_, a_prev, _ = scenario("tracking_zigzag", seed=seed, H=H, warm_steps=40)
        a_prev = np.clip(np.asarray(a_prev, F), -0.9, 0.9)
        st = to_ref_state(ns_)
        u, cp2, info = ctl(None, st, params, rand.PRNGKey(seed), cp.replace(a_mean=a_prev), {"noisy_state": st})
        eps = np.stack(rand.mvn).reshape(N, 4 * H)
        regen = np.random.default_rng(seed).standard_normal((N, 4 * H) if name != "mppi" else (N * H, 4)).astype(F).reshape(N, 4 * H)
        assert np.array_equal(eps, regen), "the tests' regeneration rule does not reproduce the logged draws"
        out = dict(state24=vec24(st), time=int(st.time), pos_traj=np.asarray(st.pos_traj, F), vel_traj=np.asarray(st.vel_traj, F),
                   a_mean=a_prev, a_mean_new=np.asarray(cp2.a_mean, F), action=np.asarray(u, F), pos_mean=np.asarray(info["pos_mean"], F),
                   pos_std=np.asarray(info["pos_std"], F), seed=seed, lam=LAM, N=N, H=H)
        if table_entry is not None:
            out["a_cov"] = table_entry
        np.savez_compressed(os.path.join(HERE, f"reference_fullsize_{name}_N{N}_H{H}.npz"), **out)
        print(f"{name} call done [{time.time() - t0:.0f} s]  action {np.asarray(u)}", flush=True)

    # ---- MPPI call --------------------------------------------------------------------------------------------------------------
    ctl, cp = get_controller(env, "mppi", f"N{N}_H{H}_lam{LAM}")
    run_call("mppi", ctl, cp, seed=31)

    # ---- CoVO-offline call: table lookup (covo.py:107-108) with a supplied entry at every time index ---------------------------
    ctl, cp = get_controller(env, "covo-offline", f"N{N}_H{H}_lam{LAM}")
    pp, ns_, a_mean, _ = scenario("tracking_zigzag", seed=32, H=H, warm_steps=40)
    Rt = oracle_c.hessian_f64(ns_, o.shift_mean(np.clip(np.asarray(a_mean, F), -0.9, 0.9)), pp).astype(F)
    entry = np.asarray(CoVOController.optimize_sigma(ctl0, Rt, types.SimpleNamespace(sample_sigma=F(0.5))), F)

    class _Table:  # a_cov_offline[time] for any time: one entry (the reference indexes with env_state.time)
        def __getitem__(self, idx):
            return entry

    run_call("covo_offline", ctl, cp.replace(a_cov_offline=_Table()), seed=32, table_entry=entry)


if __name__ == "__main__":
    main()
