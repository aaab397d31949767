"""Episode loops around the controllers: the eager ``render_env`` loop (quadjax/envs/quadrotor.py:594-667)
and the ``eval_env`` protocol (:506-591: 4 reference trajectories x episodes, metric = mean/std over episodes
of the per-episode mean ``err_pos``).  The reference's jitted ``lax.scan`` variant of eval_env cannot host a
non-JAX plugin; this is the same protocol driven eagerly."""
from __future__ import annotations

import os
import pickle
from typing import Callable, Optional

import numpy as np

from . import jaxrng as jr


def run_episode(env, controller, rng: np.random.Generator, n_steps: Optional[int] = None, rng_act_fn: Optional[Callable] = None,
                record: Optional[list] = None, reset_rng: Optional[np.random.Generator] = None, state_seq: Optional[list] = None):
    """One episode: reset -> controller.reset -> n_steps x (controller call, env.step).

    rng_act_fn(step) -> the ``rng_act`` argument of each controller call (None = production RNG).
    state_seq: if a list, one dict per step is appended in the layout ``render_env`` pickles
    (``state.__dict__``, envs/quadrotor.py:655-666) -- the keys scripts/vis.py reads (pos, quat, pos_tar,
    f_disturb, pos_traj) among them.
    Returns (err_pos[n_steps], rewards[n_steps])."""
    params = env.default_params
    n_steps = n_steps or params.max_steps_in_episode
    obs, info, state = env.reset(reset_rng if reset_rng is not None else rng, params)
    control_params = controller.reset(state, params, controller.init_control_params, None)
    errs, rews = [], []
    for i in range(n_steps):
        if record is not None:
            record.append((info["noisy_state"].to_state24(), int(info["noisy_state"].time)))
        rng_act = rng_act_fn(i) if rng_act_fn else None
        if state_seq is not None:
            state_seq.append(dict(state.__dict__))
        action, control_params, _ = controller(obs, state, params, rng_act, control_params, info)
        obs, state, reward, done, info = env.step(rng, state, action, params)
        errs.append(info["err_pos"])
        rews.append(reward)
        if done:
            break
    return np.asarray(errs), np.asarray(rews)


def run_episode_device(env, controller, rng: np.random.Generator, n_steps: Optional[int] = None,
                       reset_rng: Optional[np.random.Generator] = None, noise_seed: Optional[int] = None):
    """The same episode with the environment ON THE DEVICE: reset on the host (trajectory generation), then
    ``n_steps`` x [noisy state -> controller -> Quad3D.step_env] in one call without a host round trip
    (``covo_closed_loop``; production RNG for the samples, Philox field for observation noise / disturbances).
    Returns (err_pos[n_steps], rewards[n_steps]) like run_episode."""
    params = env.default_params
    n_steps = n_steps or params.max_steps_in_episode
    obs, info, state = env.reset(reset_rng if reset_rng is not None else rng, params)
    control_params = controller.reset(state, params, controller.init_control_params, None)
    h = controller._sync_reference(state)
    controller._upload_params(control_params)
    h.env_reset(state.to_state24()[None], [int(state.time)])
    seed = int(rng.integers(0, 2 ** 62)) if noise_seed is None else int(noise_seed)
    _, rew, err = h.closed_loop(n_steps, noise_seed=seed, gaussian=(env.disturb_type == "gaussian"),
                                obs_noise_scale=params.obs_noise_scale if env.generate_noisy_state else 0.0,
                                dyn_noise_scale=params.dyn_noise_scale)
    controller._generation += 1  # the device-resident mean has moved on: stale params objects are refused
    return err[:, 0], rew[:, 0]


def run_episode_keyed(env, controller, rng_reset, rng, n_steps: Optional[int] = None):
    """``run_one_ep`` of eval_env (quadjax/envs/quadrotor.py:542-563) with the reference's key schedule: reset from
    ``rng_reset``; ``rng_control, rng = split(rng)``; per step ``rng, rng_act, rng_step, rng_control = split(rng, 4)``, the
    controller call with ``rng_act``, ``env.step(rng_step, ...)`` (auto-reset on done, the scan never stops early), then
    ``rng, rng_control = split(rng)``.  Returns (rng, err_pos[n_steps], rewards[n_steps])."""
    params = env.default_params
    n_steps = n_steps or params.max_steps_in_episode
    obs, info, state = env.reset(rng_reset, params)
    rng_control, rng = jr.split(rng)
    control_params = controller.reset(state, params, controller.init_control_params, rng_control)
    errs, rews = [], []
    for _ in range(n_steps):
        rng, rng_act, rng_step, rng_control = jr.split(rng, 4)
        action, control_params, _ = controller(obs, state, params, rng_act, control_params, info)
        obs, state, reward, done, info = env.step(rng_step, state, action, params)
        rng, rng_control = jr.split(rng)
        errs.append(info["err_pos"])
        rews.append(reward)
    return rng, np.asarray(errs), np.asarray(rews)


def render_env(env, controller, control_params=None, repeat_times: int = 1, filename: str = "", results_dir: Optional[str] = "results",
               seed: int = 1):
    """The eager loop of the reference's ``render_env`` (quadjax/envs/quadrotor.py:594-667) with its key schedule:
    ``PRNGKey(1)``; one split each for the parameter draw, the reset and the controller reset; per step
    ``rng, rng_act, rng_step = split(rng, 3)``; after a ``done`` two more splits (parameters, controller reset) and the
    controller is reset with its CURRENT params (:636-641).  Runs until ``repeat_times`` episodes have ended and writes
    ``results/state_seq_{filename}.pkl`` (list of per-step state dicts, :655-666); plotting (``utils.plot_states``) is out of
    scope.  Returns (state_seq, reward_seq)."""
    rng = jr.PRNGKey(seed)
    rng, rng_params = jr.split(rng)
    env_params = env.sample_params(rng_params)
    state_seq, reward_seq = [], []
    rng, rng_reset = jr.split(rng)
    obs, info, env_state = env.reset(rng_reset, env_params)
    rng, rng_control = jr.split(rng)
    control_params = controller.reset(env_state, env_params, controller.init_control_params, rng_control)
    n_dones = 0
    while n_dones < repeat_times:
        state_seq.append(dict(env_state.__dict__))
        rng, rng_act, rng_step = jr.split(rng, 3)
        action, control_params, _ = controller(obs, env_state, env_params, rng_act, control_params, info)
        if hasattr(control_params, "quat_desired"):  # :626-627
            state_seq[-1]["quat_desired"] = control_params.quat_desired
        next_obs, next_env_state, reward, done, info = env.step(rng_step, env_state, action, env_params)
        if done:
            rng, rng_params = jr.split(rng)
            env_params = env.sample_params(rng_params)
            rng, rng_control = jr.split(rng)
            control_params = controller.reset(env_state, env_params, control_params, rng_control)
            n_dones += 1
        reward_seq.append(reward)
        obs, env_state = next_obs, next_env_state
    if results_dir is not None:
        save_state_seq(state_seq, filename, results_dir)
    return state_seq, reward_seq


def eval_env(env, controller, total_steps: int = 300 * 4 * 10, num_trajs: int = 4, seed: int = 1, keyed: bool = False):
    """quadjax/envs/quadrotor.py:506-591 (PRNGKey(1); num_trajs reference trajectories, each re-used for
    num_eps // num_trajs episodes).  Returns (mean, std, per-episode array).

    keyed=True threads JAX PRNG keys exactly as the reference does (``rng, rng_reset_meta = split(PRNGKey(seed))``,
    ``split(rng_reset_meta, num_trajs)``, then run_episode_keyed): the four reference trajectories, initial disturbances,
    observation noise and the controllers' sample streams are then the ones the reference's eval_env generates."""
    T = env.default_params.max_steps_in_episode
    num_eps = int(total_steps // T)
    if keyed:
        rng = jr.PRNGKey(seed)
        rng, rng_reset_meta = jr.split(rng)
        out = []
        for rng_reset in jr.split(rng_reset_meta, num_trajs):
            for _ in range(num_eps // num_trajs):
                rng, errs, _ = run_episode_keyed(env, controller, rng_reset, rng, T)
                out.append(errs.mean())
        out = np.asarray(out)
        return float(out.mean()), float(out.std()), out
    rng = np.random.default_rng(seed)
    out = []
    for i in range(num_trajs):
        traj_seed = int(rng.integers(0, 2 ** 31))
        for _ in range(num_eps // num_trajs):
            errs, _ = run_episode(env, controller, rng, T, reset_rng=np.random.default_rng(traj_seed))
            out.append(errs.mean())
    out = np.asarray(out)
    return float(out.mean()), float(out.std()), out


def save_eval_results(err_pos_ep: np.ndarray, filename: str, results_dir: str = "results") -> str:
    """``results/eval_err_pos_{filename}.pkl`` exactly as eval_env writes it (envs/quadrotor.py:581-591): one pickled
    NumPy array of per-episode mean err_pos."""
    os.makedirs(results_dir, exist_ok=True)
    path = os.path.join(results_dir, f"eval_err_pos_{filename}.pkl")
    with open(path, "wb") as f:
        pickle.dump(np.asarray(err_pos_ep), f)
    return path


def save_state_seq(state_seq: list, filename: str, results_dir: str = "results") -> str:
    """``results/state_seq_{filename}.pkl`` as render_env writes it (envs/quadrotor.py:655-666): a pickled list of
    per-step state dicts, the input of scripts/vis.py:70-95."""
    os.makedirs(results_dir, exist_ok=True)
    path = os.path.join(results_dir, f"state_seq_{filename}.pkl")
    with open(path, "wb") as f:
        pickle.dump(list(state_seq), f)
    return path
