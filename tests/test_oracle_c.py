"""oracle/covo_oracle.c (C/OpenMP restatement used as the CPU baseline) against oracle/oracle_np.py."""
import numpy as np
import pytest

from oracle import oracle_c, oracle_np as o
from tests.util import scenario


@pytest.mark.parametrize("task,H,time", [("tracking_zigzag", 12, 0), ("tracking", 8, 0), ("tracking_zigzag", 10, 294)])
def test_c_port_matches_numpy_oracle(task, H, time):
    assert oracle_c.available()
    p, ns, a_mean, rng = scenario(task, seed=3, H=H, warm_steps=6, zero_disturb=False, time=time)
    a = np.clip(a_mean[None] + 0.5 * rng.standard_normal((96, H, 4)).astype(np.float32), -1, 1)
    c_np = o.rollout_costs(ns, a, p)
    c_c = oracle_c.rollout_costs(ns, a, p)
    assert np.abs(c_np - c_c).max() < 2e-5 * max(1, np.abs(c_np).max())
    am = a_mean.copy()
    am[1, 2] = 1.0
    am[2, 0] = -1.2
    R_np = o.get_hessian(ns, am, p, dtype=np.float64)
    R_c = oracle_c.hessian(ns, am, p)
    assert np.abs(R_c - R_np).max() < 5e-5 * max(1, np.abs(R_np).max())
    assert np.abs(R_c - R_c.T).max() == 0 and np.abs(R_c[-4:]).max() == 0


def test_c_port_full_step():
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=5, H=10, warm_steps=4)
    eps = rng.standard_normal((256, 40)).astype(np.float32)
    u_c, mean_c = oracle_c.covo_step(ns, a_mean, eps, p, 0.01)
    u_n, mean_n, _, _ = o.covo_call(ns, a_mean, eps, p, lam=0.01)
    assert np.abs(mean_c - mean_n).max() < 5e-4
