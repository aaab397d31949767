"""Shared protocol of the powered closed-loop tracking comparison: who draws what, in which order.

The reference's evaluation (envs/quadrotor.py:564-579) runs 4 reference trajectories x 10 episodes x 300 steps and reports
the mean / std over episodes of the per-episode mean ||pos_tar - pos||.  Here episode k (0..39) uses

  * trajectory seed  k // 10          -> reset_env (zigzag generator + the initial f_disturb ~ U(-0.2, 0.2), quadrotor.py:300-305)
  * noise seed       5000 + k         -> one block of 16 standard normals per environment step: z[0:13] observation noise in the order
                                         pos3, vel3, quat4, omega3 (get_info, quadrotor.py:323-351), z[13:16] the disturbance normals
                                         (unused for disturb_type none); block 0 is the noisy copy of the initial state
  * sample seed      9000 + k         -> eps [N, 4H] float32 per MPC step from numpy default_rng (oracle arm and the device's
                                         "identical eps" arm); the device's production arm draws its own Philox field instead.

Test infrastructure: imported by tools/oracle_tracking_stats.py (CPU, oracle), tools/device_tracking_stats.py and bench.py."""
import numpy as np

N_EPISODES, N_TRAJ, EP_STEPS = 40, 4, 300
TASK = "tracking_zigzag"


def episode_seeds(k: int):
    return k // (N_EPISODES // N_TRAJ), 5000 + k, 9000 + k


def episode_noise(k: int, n_steps: int = EP_STEPS) -> np.ndarray:
    """[n_steps + 1, 16] float32 standard normals of episode k."""
    return np.random.default_rng(episode_seeds(k)[1]).standard_normal((n_steps + 1, 16)).astype(np.float32)


def episode_eps_rng(k: int) -> np.random.Generator:
    return np.random.default_rng(episode_seeds(k)[2])


class SeqRng:
    """Hands out pre-drawn standard normals in the order the oracle asks for them."""

    def __init__(self, values):
        self.v = [float(x) for x in values]
        self.i = 0

    def standard_normal(self):
        x = self.v[self.i]
        self.i += 1
        return x
