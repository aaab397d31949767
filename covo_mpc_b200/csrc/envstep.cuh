// Device-resident environment step (the caller side of the hot path): internal declarations.
#pragma once
#include "common.cuh"

namespace covo {

constexpr int kEnvNoiseFloats = 16;  // per environment and step: 13 observation-noise normals + 3 disturbance normals

struct EnvStepArgs {
    EnvConsts env;
    int n_env, traj_len;
    long long traj_stride;   // floats between environments in pos_traj / vel_traj (0: shared)
    float obs_noise_scale;   // EnvParams3D.obs_noise_scale (0.05)
    float dyn_noise_scale;   // EnvParams3D.dyn_noise_scale (0.05), used when gaussian != 0
    int gaussian;            // disturb_type == "gaussian"
    int do_step;             // 0: only draw the noisy copy of the current state (after a reset)
    unsigned long long seed;
    unsigned int stream;     // step counter of the noise field
    float* state24;          // [E][24] TRUE state, in/out
    int* time;               // [E] in/out
    const float* pos_traj;   // [E][T][3]
    const float* vel_traj;   // [E][T][3]
    const float* action;     // [E][4]
    const float* noise_in;   // optional [E][16] normals supplied by the caller (parity mode), else Philox
    float* noisy24;          // [E][24] out: info["noisy_state"] for the next controller call
    int* noisy_time;         // [E] out
    float* reward;           // [E] out, of the PRE-step state
    float* err_pos;          // [E] out, of the PRE-step state
    int* done;               // [E] out, of the PRE-step state
    // CUDA-graph replay of a closed loop: when set, the step index i = *step_ctr selects the noise stream (stream + i), the noise
    // block (noise_in + i * E * 16) and the log rows (reward / err_pos + i * E, action_log + i * E * 4 <- action)
    const unsigned int* step_ctr = nullptr;
    float* action_log = nullptr;
    // Auto-reset of BaseEnvironment.step (envs/base.py:27-38): when the PRE-step state is terminal, the stepped state is discarded
    // and environment e continues from the next entry of its reset pool (reset_env draws prepared by the host: state, time, reference
    // trajectory), the trajectory tables are overwritten in place (the controller kernels read the same buffers), and -- what the
    // reference's harness does next (envs/quadrotor.py:637-639: controller.reset returns the initial parameters) -- the
    // controller's resident mean goes back to its initial value.  reset_pool == 0: no auto-reset (episodes bounded by the caller).
    // MPPI under disturb_type gaussian (mppi.py:74): the force the NEXT controller call's rollouts see (one draw per call, shared by all
    // samples and horizon steps), written here so that the device-resident loop needs no host draw: [E][mppi_H][3]
    float* mppi_fdist = nullptr;
    int mppi_H = 0;
    int reset_pool = 0;
    const float* reset_state24 = nullptr;  // [P][E][24]
    const int* reset_time = nullptr;       // [P][E]
    const float* reset_pos_traj = nullptr; // [P][E][T][3]
    const float* reset_vel_traj = nullptr; // [P][E][T][3]
    int* reset_count = nullptr;            // [E]: resets taken so far (entry = count % P)
    float* traj_pos_rw = nullptr;          // = pos_traj / vel_traj, writable
    float* traj_vel_rw = nullptr;
    float* a_mean = nullptr;               // optional [E][n_mean]: the controller's mean, reset with the environment
    const float* a_mean_init = nullptr;    // [n_mean]
    int n_mean = 0;
};

// one thread: advances the device step counters at the end of a replayed step
cudaError_t launch_bump(unsigned int* ctr_a, unsigned int* ctr_b, cudaStream_t st, const int* status_src = nullptr, int* status_dst = nullptr,
                        int n_status = 0, unsigned int* flag = nullptr, const float* action_src = nullptr, float* action_dst = nullptr);

cudaError_t launch_env_step(const EnvStepArgs& a, cudaStream_t st);

}  // namespace covo
