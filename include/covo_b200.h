/* covo_b200.h -- C-ABI of the B200-native CoVO-MPC / MPPI inner loop.
 *
 * This is the drop-in boundary for the controller hot path of LeCAR-Lab/CoVO-MPC
 * (reference paths are relative to the reference checkout):
 *
 *   covo_create / covo_destroy     <-> CoVOController.__init__ / MPPIController.__init__
 *                                      quadjax/controllers/covo.py:26-114, mppi.py:22-26,
 *                                      constants chosen by get_controller, envs/quadrotor.py:670-752
 *   covo_step / covo_step_device   <-> CoVOController.__call__  controllers/covo.py:187-283
 *                                      MPPIController.__call__  controllers/mppi.py:28-134
 *   covo_reset_offline[_disturbed] <-> reset_a_cov_offline      controllers/covo.py:58-104
 *   covo_hessian                   <-> CoVOController.get_hessian      controllers/covo.py:134-185
 *   covo_optimize_sigma            <-> CoVOController.optimize_sigma   controllers/covo.py:116-132
 *   covo_cholesky                  <-> the factorisation inside jax.random.multivariate_normal
 *                                      (controllers/covo.py:216, mppi.py:59)
 *   covo_rollout                   <-> sample + rollout + softmax update, controllers/covo.py:212-278
 *
 * Plain pointers and sizes only; no torch / CUDA types in the signatures (a stream is a void*).
 * Every function returns 0 on success, non-zero on failure; covo_last_error() gives the reason.
 * The reference raises NotImplementedError / AssertionError for bad arguments
 * (envs/quadrotor.py:751, controllers/covo.py:45-47, :114); the Python host maps the codes below
 * to the same exception types.
 *
 * Ownership: the handle owns its device workspace.  Buffers passed in are caller-owned and never
 * freed or retained beyond the call (except by the *_device variants, which read them on the given
 * stream).  One handle <-> one device <-> one stream at a time; calls on one handle are not
 * thread-safe, distinct handles are independent.
 *
 * Layouts (all float32, C order):
 *   state24   [E][24]: pos(3) quat xyzw(4) vel(3) omega(3) f_disturb(3) pos_tar(3) vel_tar(3) pad(2)
 *             -- the "noisy_state" the reference controllers plan from (controllers/covo.py:198)
 *   time      [E] int32: env_state.time
 *   a_mean    [E][H][4]      a_cov (CoVO) [E][4H][4H]      a_cov (MPPI) [E][H][4][4]
 *   eps       [E][N_local][4H]  standard normal draws ("parity mode": the caller supplies what
 *             jax.random.normal would have drawn); NULL = in-kernel counter RNG ("production mode")
 *   reference trajectory  pos/vel/acc [E][T][3]   (dynamics/utils.py:49-53, 87-130, 183-251)
 */
#ifndef COVO_B200_H
#define COVO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define COVO_OK 0
#define COVO_ERR_INVALID 1        /* bad argument (reference: AssertionError / ValueError) */
#define COVO_ERR_NOT_IMPLEMENTED 2 /* unsupported mode / option (reference: NotImplementedError) */
#define COVO_ERR_CUDA 3           /* CUDA runtime failure; message carries cudaGetErrorString */
#define COVO_ERR_NUMERIC 4        /* spectral range beyond the rational table, Cholesky breakdown */

#define COVO_MODE_MPPI 0
#define COVO_MODE_COVO_ONLINE 1
#define COVO_MODE_COVO_OFFLINE 2

typedef struct covo_handle covo_handle;

typedef struct covo_config {
    int mode;       /* COVO_MODE_* */
    int n_samples;  /* N, global number of sampled control sequences (reference default 8192) */
    int horizon;    /* H (reference default 32); u_dim is 4 */
    int n_env;      /* environments batched behind this handle (reference: 1) */
    int traj_len;   /* rows T of the reference trajectory (300 / 320 / 350) */
    int device;     /* CUDA device ordinal */
    int rank;       /* N-sharding: this handle rolls samples [rank*N/world, (rank+1)*N/world) */
    int world;
    float lam;           /* temperature (reference default 0.01) */
    float sample_sigma;  /* sigma (0.5) */
    float gamma_mean;    /* 1.0 */
    float gamma_sigma;   /* 0.0 (reference default); MPPI: != 0 enables the covariance update of mppi.py:119-125 (world == 1); ignored by CoVO */
    float discount;      /* 1.0 */
    /* EnvParams3D, quadjax/dynamics/dataclass.py:40-100 */
    float m, g, max_thrust, dt, alpha_bodyrate, action_scale, pos_limit;
    float max_omega[3];
    int max_steps_in_episode;
    unsigned long long seed; /* production-mode RNG key */
} covo_config;

/* Fill *cfg with the reference defaults (get_controller with controller_params == "",
 * envs/quadrotor.py:671-683, and EnvParams3D defaults). */
int covo_default_config(covo_config* cfg);

int covo_create(const covo_config* cfg, covo_handle** out);
int covo_destroy(covo_handle* h);
const char* covo_last_error(void);
const char* covo_version(void);

/* Reference trajectory held in the env state (EnvState3D.pos_traj / vel_traj / acc_traj). Host
 * pointers, [E][T][3]; acc may be NULL (zeros). */
int covo_set_reference(covo_handle* h, const float* pos_traj, const float* vel_traj, const float* acc_traj);

/* get_controller(env, "pid") -- PIDController.__call__ (controllers/pid.py:38-83; gains of envs/quadrotor.py:692-699) for the
 * handle's n_env environments: position PD(+I) -> desired thrust vector -> attitude error -> (thrust, body rates) in the
 * normalised action box.  state24/time as in covo_step; acc_tar is looked up in the reference's acceleration table
 * (covo_set_reference) at `time`; integral: [n_env][3] or NULL (Ki term skipped).  All pointers are HOST pointers. */
int covo_pid_action(covo_handle* h, const float* state24, const int* time, float Kp, float Kd, float Ki, float Kp_att,
                    const float* integral, float* action);

/* control_params.a_mean / a_cov accessors (host pointers). */
int covo_set_mean(covo_handle* h, const float* a_mean);
int covo_get_mean(covo_handle* h, float* a_mean);
int covo_set_cov(covo_handle* h, const float* a_cov); /* MPPI: [E][H][4][4]; CoVO: [E][4H][4H] (factorised on device) */
int covo_get_cov(covo_handle* h, float* a_cov);
/* CoVO-offline table a_cov_offline [T_sched][4H][4H] (E == 1); factorised on device in one batch. */
int covo_set_cov_offline(covo_handle* h, const float* table, int t_sched);
int covo_get_cov_offline(covo_handle* h, float* table, int t_sched);
/* Build the table on device: PID expansion policy closed loop + nominal rollouts + Hessian + sigma,
 * as controllers/covo.py:58-104 with disturb_type == "none".  state24/time: the reset state. */
int covo_reset_offline(covo_handle* h, const float* state24, const int* time, int t_sched);
/* The same under disturb_type == "gaussian" (the reference's default, envs/quadrotor.py:765): the state advance between schedule
 * entries (controllers/covo.py:86-89) is stochastic.  f_disturb [t_sched][3] (host): the force dyn_noise_scale * N(0, I)
 * (dynamics/free.py:66-70) the state carries after path step t -- drawn by the caller, with the reference's key schedule when it
 * holds a JAX key (covo_mpc_b200/controllers.py does).  The H-step nominal rollouts stay deterministic (covo.py:67-69).
 * f_disturb == NULL: covo_reset_offline. */
int covo_reset_offline_disturbed(covo_handle* h, const float* state24, const int* time, int t_sched, const float* f_disturb);

/* One MPC step for all E environments.  Host buffers; H2D of (state24, time[, eps]) and D2H of the
 * action happen inside the call, which returns after the stream is idle.  eps may be NULL. */
int covo_step(covo_handle* h, const float* state24, const int* time, const float* eps, float* action);
/* Numeric status: the covariance step of every CoVO-online call starts by clearing the per-environment status (it is per step,
 * not sticky).  covo_step reads it back with the action and returns COVO_ERR_NUMERIC (action still written) when the spectral
 * range left the rational-approximation ladder (1) or a Cholesky pivot was not positive (2) -- where the reference would surface
 * NaNs (controllers/covo.py:116-132, :216); status 3 (dense path only, see covo_set_sigma_path) is not an error.  The asynchronous entry points (covo_step_device, covo_step_partial_device,
 * covo_closed_loop) cannot: their callers poll covo_get_status(). */
/* MPPI under disturb_type "gaussian" (the reference's default): its rollouts call step_env WITHOUT deterministic=True
 * (controllers/mppi.py:74), every sample and every horizon step with the same step_key, i.e. all of them see the same force
 * dyn_noise_scale * N(0, I)^3 from the second step on (dynamics/free.py:66-70, 144-147).  fdist_seq [E][H][3] (host): the force produced
 * by rollout step h (it acts during step h + 1); used by every following covo_step* of the handle until replaced; NULL switches it off
 * (disturb_type "none").  CoVO's rollouts are deterministic (controllers/covo.py:229) and do not take one. */
int covo_set_rollout_disturbance(covo_handle* h, const float* fdist_seq);
/* The physical model a handle plans with (dynamics/dataclass.py:43-49, 71, 76, 81), replaceable after creation: the reference rolls
 * out and differentiates with the env_params of the CALL (controllers/covo.py:187-283 `env_params`), e.g. a mass sampled by
 * Quad3D.sample_params.  Takes effect with the next launch (the constants travel in the kernel argument blocks). */
int covo_set_env_params(covo_handle* h, float m, float g, float max_thrust, float dt, float alpha_bodyrate, float action_scale,
                        const float* max_omega3, int max_steps_in_episode);
/* Same, device pointers, asynchronous on `stream` (a cudaStream_t passed as void*). */
int covo_step_device(covo_handle* h, const float* state24_dev, const int* time_dev, const float* eps_dev,
                     float* action_dev, void* stream);

/* N-sharded step (world > 1): phase A leaves this rank's (min cost, sum w, sum w*u) record, 4 + n_pad
 * floats per environment, in the buffer returned by covo_partial_buffer(); the caller all-gathers the
 * records of all ranks (NCCL, rank order) and calls phase B on every rank. */
int covo_step_partial_device(covo_handle* h, const float* state24_dev, const int* time_dev, const float* eps_dev,
                             void* stream);
int covo_partial_buffer(covo_handle* h, float** dev_ptr, int* n_floats);
int covo_step_merge_device(covo_handle* h, const float* gathered_dev, float* action_dev, void* stream);
/* Fused exchange for the same step (world <= 8 ranks on one NVLink domain): instead of kernel -> all-gather -> kernel, the rollout
 * kernel's finalising CTA writes the rank record (832 B at H = 50) straight into the exchange buffer of EVERY rank -- peer device
 * memory mapped through CUDA IPC -- and raises a flag there; the merge kernel, launched right behind it, waits for the flags of all
 * ranks and merges in rank order (bit-identical to the all-gather path).  No collective call, no host synchronisation.
 *   covo_exchange_info   : this rank's buffer as a 64-byte cudaIpcMemHandle_t (ship it to the other processes, e.g. with
 *                          torch.distributed.all_gather_object) and/or as a raw device pointer (handles of one process);
 *   covo_exchange_attach : make rank `peer_rank`'s buffer known, by IPC handle or by pointer (exactly one non-NULL);
 *   covo_step_sharded_device : one MPC step; every rank must call it the same number of times (slots are keyed by the call count).
 * A peer that never delivers is reported as status 4 after a 4 s watchdog instead of hanging the device. */
int covo_exchange_info(covo_handle* h, void* ipc_handle64, void** dev_ptr);
int covo_exchange_attach(covo_handle* h, int peer_rank, const void* ipc_handle64, void* dev_ptr);
int covo_step_sharded_device(covo_handle* h, const float* state24_dev, const int* time_dev, const float* eps_dev, float* action_dev, void* stream);

/* Operators, exposed on their own for parity tests (host pointers, synchronous). */
int covo_hessian(covo_handle* h, const float* state24, const int* time, const float* a_mean, int shift, float* R);
int covo_optimize_sigma(covo_handle* h, const float* R, float* a_cov);
int covo_cholesky(covo_handle* h, const float* a_cov, float* L);
/* sample + rollout + softmax update with an explicit covariance factor source:
 *   uses the handle's current factor (set by covo_set_cov / covo_optimize_sigma+covo_cholesky).
 * Optional outputs may be NULL: costs [E][N_local], samples [E][N_local][4H]. */
int covo_rollout(covo_handle* h, const float* state24, const int* time, const float* a_mean, int shift,
                 const float* eps, const float* fdist_seq, float* a_mean_out, float* action, float* costs,
                 float* samples);

/* JAX-compatible sampling stream.  Replaces, for ONE following sampling call (covo_step,
 * covo_step_device, covo_step_partial_device or covo_rollout without explicit eps), the Philox field by the draws the
 * reference itself would make from `act_key`:  act_keys = jax.random.split(act_key, N);  sample i draws
 * jax.random.normal(act_keys[i], (4H,))  (controllers/covo.py:212-221)  or, in MPPI mode, normal(split(act_keys[i], H)[h], (4,))
 * (controllers/mppi.py:53-61) -- Threefry-2x32-20, legacy counter layout, erfinv-based normal.  `act_key` = two uint32 words
 * on the host.  The caller derives act_key from its rng_act exactly as the reference does (rng_act, act_key = split(rng_act));
 * covo_mpc_b200/jaxrng.py is the host twin.  N-sharded handles index the stream by GLOBAL sample number. */
int covo_set_jax_key(covo_handle* h, const unsigned int* act_key);

/* Debug / introspection. */
int covo_get_pos_stats(covo_handle* h, float* pos_mean, float* pos_std); /* info dict, controllers/covo.py:281; [E][H][3] each */
int covo_enable_pos_stats(covo_handle* h, int on);
int covo_debug_eps(covo_handle* h, unsigned int stream_id, float* eps);  /* the production-mode field, [N_local][4H] */
int covo_debug_tridiag(covo_handle* h, double* d, double* e, double* scalars5); /* after covo_optimize_sigma / a step, env 0; dense path: d = e = 0, scalars5 = lambda_min, spectrum upper bound, log det A, Lanczos steps taken, 0 */
int covo_zolotarev_nodes(double m, double M, int n_poles, double* shifts, double* weights); /* host only, no GPU */
int covo_get_status(covo_handle* h, int* status); /* [E] numeric status of the last covariance step */
/* ---- device-resident environment and closed loop -------------------------------------------
 * The caller side of the hot path on the device, so that noisy state -> controller -> env step runs without a
 * host round trip.  Replaces Quad3D.step_env (envs/quadrotor.py:215-248: reward/done of the PRE-step state, then
 * free_dynamics_3d_bodyrate, dynamics/free.py:114-202) and Quad3D.get_info (envs/quadrotor.py:314-361: the noisy
 * state the next controller call plans from).  No auto-reset (envs/base.py:27-38): episodes are bounded by the caller.
 * `noise`: 16 standard normals per environment and step (13 observation noise: pos3 vel3 quat4 omega3; 3 disturbance
 * force) supplied by the caller, or NULL -> Philox field keyed by (noise_seed, step, environment). */
int covo_env_reset(covo_handle* h, const float* state24, const int* time);   /* [E][24], [E]: the TRUE state */
int covo_env_get_state(covo_handle* h, float* state24, int* time);
/* one transition; action == NULL only draws the noisy copy of the current state (what reset_env's get_info does) */
int covo_env_step(covo_handle* h, const float* action, const float* noise, unsigned long long noise_seed,
                  unsigned int noise_step, int gaussian_disturbance, float obs_noise_scale, float dyn_noise_scale,
                  float* noisy24, float* reward, float* err_pos, int* done);
/* n_steps x [controller call (production RNG) -> env step], all on the device; noise: [n_steps+1][E][16] or NULL.
 * Outputs (host, optional): actions [n_steps][E][4], rewards / err_pos [n_steps][E] (of the pre-step states). */
int covo_closed_loop(covo_handle* h, int n_steps, unsigned long long noise_seed, int gaussian_disturbance,
                     float obs_noise_scale, float dyn_noise_scale, const float* noise, float* actions, float* rewards,
                     float* err_pos);
/* Auto-reset of BaseEnvironment.step (envs/base.py:27-38) on the device.  The host prepares n_pool reset_env draws per environment
 * (state24 [n_pool][E][24], time [n_pool][E], pos_traj / vel_traj [n_pool][E][T][3]: envs/quadrotor.py:265-312 with its own key
 * schedule); from then on covo_env_step / covo_closed_loop replace the state of an environment whose PRE-step state was terminal by
 * its next pool entry (round robin), overwrite its reference trajectory in place, and -- when a_mean_init [4H] is given -- put the
 * controller's resident mean back to it, which is what the reference's harness does after `done` (envs/quadrotor.py:637-639:
 * controller.reset returns the initial parameters).  n_pool == 0 switches it off.  Not for covo-offline (schedule per episode). */
int covo_env_set_reset_pool(covo_handle* h, int n_pool, const float* state24, const int* time, const float* pos_traj, const float* vel_traj,
                            const float* a_mean_init);
/* Which optimize_sigma (controllers/covo.py:116-132) kernels the handle runs:
 *   3 = dense path (default of a single-environment handle): adaptive Lanczos lambda_min (float64) -> one float64 Gauss-Jordan inverse
 *       per pole of a 13-pole rational approximation of x^(-1/2), each on an 8-CTA cluster -> combine.  Sigma within 1e-7 .. 2e-7
 *       (relative Frobenius) of exact arithmetic on the same float32 Hessian (the reference's float32 eigh: 4e-7 .. 1e-5).
 *   0 = tridiagonal path (default of environment batches, 8 CTAs per matrix instead of 112): Householder -> tridiagonal matrix
 *       function in float64 -> Q F Q^T.  Sigma within 1e-6 .. 1e-5 of exact arithmetic; 0.1 ms more per step at H = 50.
 * COVO_SIGMA=dense / COVO_SIGMA=tridiag choose at handle creation. */
int covo_get_sigma_path(covo_handle* h, int* path);
/* Select the path (0 or 3).  The dense path needs a lowest eigenvalue that a Lanczos iteration finds within 64 steps (16 .. 52 on CoVO
 * Hessians; the kernel runs as many as the residual of the Ritz pair asks for).  If it does not converge, lambda_min is replaced by a
 * lower bound (Sigma stays positive definite) and the step ends with status 3: covo_optimize_sigma then redoes the matrix on the
 * tridiagonal path by itself; covo_step switches the handle to the tridiagonal path for the following steps; callers of the
 * asynchronous entry points see status 3 in covo_get_status() and can switch with this function. */
int covo_set_sigma_path(covo_handle* h, int path);
/* Per-kernel device time of the last instrumented step, CUDA events on the launch stream.
 * slots: 0 hessian (local + assemble + forward chains), 4 cholesky, 5 rollout, and
 *   tridiagonal path: 1 tridiagonalisation (cluster), 2 tridiagonal matrix function, 3 Sigma = Q F Q^T
 *   dense path:       1 Lanczos (cluster),            2 shifted inverses (one cluster per pole), 3 combine */
int covo_set_profiling(covo_handle* h, int on);
int covo_get_kernel_ms(covo_handle* h, float* ms6);
/* Debug: switch in-kernel clock64() phase stamps on/off and read the 64 slots of the last step (may be NULL). */
int covo_debug_phase_clocks(covo_handle* h, int on, long long* out64);
int covo_rng_step(covo_handle* h, unsigned int* stream_id); /* counter of production-mode draws so far */
int covo_local_samples(covo_handle* h, int* n_local, int* offset);

#ifdef __cplusplus
}
#endif
#endif /* COVO_B200_H */
