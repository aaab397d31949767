"""covo_mpc_b200: B200-native CoVO-MPC / MPPI inner loop behind the reference's controller plugin surface.

The numerical path is libcovo_b200.so (hand-written sm_100a CUDA, C-ABI in include/covo_b200.h); importing
this package loads it and fails loudly if it has not been built.  There is no CPU fallback.
"""
from . import _lib
from ._lib import Handle, CovoConfig, default_config, MODE_MPPI, MODE_COVO_ONLINE, MODE_COVO_OFFLINE
from .env import Quad3D, EnvParams3D, EnvState3D
from .controllers import (BaseController, MPPIController, CoVOController, PIDController, RandomController, MPPIParams, CoVOParams,
                          PIDParams, get_controller)
from . import jaxrng
from .harness import run_episode, run_episode_device, run_episode_keyed, render_env, eval_env, save_eval_results, save_state_seq

_lib.load()

__all__ = ["Handle", "CovoConfig", "default_config", "MODE_MPPI", "MODE_COVO_ONLINE", "MODE_COVO_OFFLINE",
           "Quad3D", "EnvParams3D", "EnvState3D", "BaseController", "MPPIController", "CoVOController",
           "PIDController", "RandomController", "MPPIParams", "CoVOParams", "PIDParams", "get_controller", "jaxrng", "run_episode", "run_episode_device", "run_episode_keyed", "render_env", "eval_env", "save_eval_results", "save_state_seq"]
