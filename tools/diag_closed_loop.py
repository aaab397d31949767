"""Step-by-step comparison of the device controller and the oracle on the headline closed loop (development tool,
GPU box): where do the two first differ -- covariance, Cholesky factor, mean, action?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as o  # noqa: E402
import covo_mpc_b200 as cm  # noqa: E402
from tests.test_step_gpu import _to_env_state  # noqa: E402


def main():
    N, H, steps = 8192, 50, 6
    p = o.EnvParams()
    rng = np.random.default_rng(11)
    s_dev = o.reset_env("tracking_zigzag", p, rng, dtype=np.float32, zero_disturb=True)
    s_ora = s_dev.copy()
    env = cm.Quad3D("tracking_zigzag")
    ctl, cp = cm.get_controller(env, "covo-online", f"N{N}_H{H}_lam0.01")
    mean_o = o.hover_mean(H, p)
    for i in range(steps):
        ns_dev = o.noisy_state(s_dev, p, np.random.default_rng(1000 + i))
        ns_ora = o.noisy_state(s_ora, p, np.random.default_rng(1000 + i))
        eps = np.random.default_rng(2000 + i).standard_normal((N, 4 * H)).astype(np.float32)
        st = _to_env_state(cm, ns_dev)
        a_dev, cp, _ = ctl(None, st, env.default_params, eps, cp, {"noisy_state": st})
        a_ora, mean_o, cov_o, _, dbg = o.covo_call(ns_ora, mean_o, eps, p, lam=0.01, return_debug=True)
        cov = np.asarray(cp.a_cov)
        mean_d = np.asarray(cp.a_mean)
        c = np.sort(dbg["cost"].astype(np.float64))
        Ld = np.linalg.cholesky(cov.astype(np.float64))
        print(f"step {i}: |a_dev-a_ora| {np.abs(a_dev - a_ora).max():.2e}  |mean_dev-mean_ora| {np.abs(mean_d - mean_o).max():.2e} "
              f"cov rel {np.linalg.norm(cov - cov_o) / np.linalg.norm(cov_o):.2e}  |L44 diff| "
              f"{np.abs(Ld[:4, :4] - dbg['L'][:4, :4]).max():.2e}  gap {(c[1] - c[0]) / 0.01:.1f}  state diff "
              f"{np.abs(o.state_to_vec24(ns_dev) - o.state_to_vec24(ns_ora)).max():.2e} argmin_ora {int(np.argmin(dbg['cost']))}")
        s_dev, _, _, _ = o.env_step(s_dev, a_dev, p, rng, "none")
        s_ora, _, _, _ = o.env_step(s_ora, a_ora, p, rng, "none")


if __name__ == "__main__":
    main()
