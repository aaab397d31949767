"""Regenerates tests/golden/*.npz from the float64 oracle (oracle/oracle_np.py).

The reference cannot be imported in this container (no jax / flax / gymnax, no network) and ships no golden
vectors of its own, so these fixtures are ORACLE-generated: they pin the oracle against accidental change and
give the CUDA path fixed, versioned inputs/outputs.  They are not reference output; the vectors produced by executing the reference itself are reference_*.npz (make_reference_golden.py).
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_np as o  # noqa: E402
from tests.util import scenario  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, task, seed, N, H, warm, mppi=False, time=0):
    p, ns, a_mean, rng = scenario(task, seed=seed, H=H, warm_steps=warm, zero_disturb=False, time=time)
    out = dict(state24=o.state_to_vec24(ns), time=np.int32(ns.time), pos_traj=ns.pos_traj.astype(np.float32),
               vel_traj=ns.vel_traj.astype(np.float32), a_mean=a_mean)
    if mppi:
        eps = rng.standard_normal((N, H, 4)).astype(np.float32)
        cov = np.tile(0.25 * np.eye(4, dtype=np.float32), (H, 1, 1))
        u, new_mean, new_cov, info, dbg = o.mppi_call(ns, a_mean, cov, eps, p, lam=0.01, return_debug=True)
        out.update(eps=eps.reshape(N, 4 * H), a_cov=cov, cost=dbg["cost"])
    else:
        eps = rng.standard_normal((N, 4 * H)).astype(np.float32)
        u, new_mean, a_cov, info, dbg = o.covo_call(ns, a_mean, eps, p, lam=0.01, return_debug=True)
        out.update(eps=eps, R=dbg["R"].astype(np.float64), a_cov=a_cov, L=dbg["L"], cost=dbg["cost"])
    out.update(action=u, a_mean_new=new_mean, pos_mean=info["pos_mean"].astype(np.float32))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "action", u)


if __name__ == "__main__":
    case("covo_online_zigzag_N64_H8", "tracking_zigzag", 1, 64, 8, 6)
    case("covo_online_lissa_N128_H12", "tracking", 2, 128, 12, 15)
    case("covo_online_zigzag_late_N64_H10", "tracking_zigzag", 3, 64, 10, 4, time=295)
    case("mppi_hover_N128_H32", "hovering", 4, 128, 32, 3, mppi=True)
