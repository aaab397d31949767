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
    float* states24;        // [t_sched][24]
    int* times;             // [t_sched]
    float* a_nom;           // [t_sched][H][4]
};

cudaError_t launch_offline_paths(const OfflineArgs& a, cudaStream_t st);

}  // namespace covo
