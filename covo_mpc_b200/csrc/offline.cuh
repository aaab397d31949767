// Internal declarations for the CoVO-offline schedule helpers (offline.cu).
#pragma once
#include "common.cuh"

namespace covo {

struct OfflineArgs {
    int H, traj_len, t_sched;
    EnvConsts env;
    float max_thrust, Kp, Kd, Kp_att;
    const float* state24;   // [24] reset state
    const int* time;        // [1]
    const float* pos_traj;  // [T][3]
    const float* vel_traj;  // [T][3]
    const float* acc_traj;  // [T][3] or nullptr
    const float* f_disturb = nullptr;  // [t_sched][3] or nullptr: the disturbance force the state carries AFTER path step t (gaussian)
    float* states24;        // [t_sched][24]
    int* times;             // [t_sched]
    float* a_nom;           // [t_sched][H][4]
};

cudaError_t launch_offline_paths(const OfflineArgs& a, cudaStream_t st);

// standalone PID policy (get_controller("pid"), controllers/pid.py:38-83): one action per environment
struct PidArgs {
    int n_env, traj_len;
    EnvConsts env;
    float max_thrust, Kp, Kd, Ki, Kp_att;
    const float* state24;   // [E][24]
    const int* time;        // [E]
    const float* acc_traj;  // [E][T][3] or nullptr (acc_tar = 0)
    const float* integral;  // [E][3] or nullptr
    float* action;          // [E][4]
};
cudaError_t launch_pid(const PidArgs& a, cudaStream_t st);

}  // namespace covo
