// Internal declarations for the Hessian kernels (hessian.cu).
#pragma once
#include "common.cuh"

namespace covo {

struct HessianArgs {
    long long* prof = nullptr;  // optional: clock64() stamps at phase boundaries (debug)
    int H, traj_len, shift;
    long long traj_stride;  // floats between environments in pos_traj / vel_traj (0: shared)
    EnvConsts env;
    const float* state24;   // [E][24]
    const int* time;        // [E]
    const float* pos_traj;  // [E][T][3]
    const float* vel_traj;  // [E][T][3]
    const float* a_mean;    // [E][H][4]   nominal controls (shift applied on load if `shift`)
    float* workspace;       // [E][H][14*153 + 14*17]
    float* R;               // [E][n][n]
    int* status = nullptr;  // [E] numeric status of the covariance step: cleared here, set by the sigma / Cholesky kernels
    int* progress = nullptr;  // [E] column-block counter of the Cholesky -> rollout pipeline: cleared here (first kernel of the step)
};

size_t hessian_workspace_floats(int H);
size_t hessian_assemble_smem(int H);
cudaError_t launch_hessian(const HessianArgs& a, int n_env, cudaStream_t st);
const void* hess_local_kernel_address();  // for CUDA-graph node lookup (capi.cu)

}  // namespace covo
