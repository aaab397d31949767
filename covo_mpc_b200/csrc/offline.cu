// K6 helpers: the PID "expansion policy" CoVO-offline uses to lay down nominal trajectories.
//
// Replaces controllers/pid.py:38-83 (PIDController.__call__ with the gains of controllers/covo.py:48-53)
// and the two scans of controllers/covo.py:58-104: the closed-loop PID path over the episode
// (get_single_a_cov_offline's state advance, :80-89) and, from every state on that path, the H-step
// deterministic PID rollout that becomes the nominal control sequence (:58-76).  The Hessian / sigma /
// Cholesky of all schedule steps then run as ONE batched launch each (capi.cu: covo_reset_offline),
// schedule step t playing the role of "environment t".
// disturb_type "none" (deterministic state advance) or "gaussian" (the caller supplies the force the state carries after each
// path step; the H-step nominal rollouts are deterministic=True in the reference, controllers/covo.py:67-69, i.e. force 0).
#include <cuda_runtime.h>

#include "common.cuh"
#include "offline.cuh"
#include "pid.cuh"

namespace covo {

namespace {

__device__ inline void gather3(const float* traj, int row, float out[3]) {
    for (int k = 0; k < 3; ++k) out[k] = traj ? traj[(long long)row * 3 + k] : 0.f;
}

}  // namespace

// closed-loop PID path over the schedule: states24[t], times[t] for t = 0 .. T_sched-1
__global__ void pid_path_kernel(const OfflineArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    QState<float> s;
    float fd[3], pt[3], vt[3], at[3];
    load_state24(a.state24, s, fd, pt, vt);
    int time = a.time[0];
    gather3(a.acc_traj, min(time, a.traj_len - 1), at);
    for (int t = 0; t < a.t_sched; ++t) {
        float* o = a.states24 + (long long)t * kStateFloats;
        for (int k = 0; k < 3; ++k) o[k] = s.p[k];
        for (int k = 0; k < 4; ++k) o[3 + k] = s.q[k];
        for (int k = 0; k < 3; ++k) o[7 + k] = s.v[k];
        for (int k = 0; k < 3; ++k) o[10 + k] = s.w[k];
        for (int k = 0; k < 3; ++k) o[13 + k] = fd[k];
        for (int k = 0; k < 3; ++k) o[16 + k] = pt[k];
        for (int k = 0; k < 3; ++k) o[19 + k] = vt[k];
        o[22] = o[23] = 0.f;
        a.times[t] = time;
        float act[4];
        pid_action(s, pt, vt, at, a.env, a.max_thrust, a.Kp, a.Kd, a.Kp_att, act);
        quad_step(s, act, fd, a.env);
        // the stochastic state advance of get_single_a_cov_offline (controllers/covo.py:86-89): disturb_type none -> 0,
        // gaussian -> dyn_noise_scale * N(0, I) drawn by the caller for every path step (dynamics/free.py:66-70)
        for (int k = 0; k < 3; ++k) fd[k] = a.f_disturb ? a.f_disturb[(long long)t * 3 + k] : 0.f;
        ++time;
        int row = min(time, a.traj_len - 1);
        gather3(a.pos_traj, row, pt);
        gather3(a.vel_traj, row, vt);
        gather3(a.acc_traj, row, at);
    }
}

// from every schedule state: H-step deterministic PID rollout -> nominal controls a_nom[t][H][4]
__global__ void pid_nominal_kernel(const OfflineArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.t_sched) return;
    QState<float> s;
    float fd[3], pt[3], vt[3], at[3];
    load_state24(a.states24 + (long long)t * kStateFloats, s, fd, pt, vt);
    int time = a.times[t];
    gather3(a.acc_traj, min(time, a.traj_len - 1), at);
    for (int h = 0; h < a.H; ++h) {
        float act[4];
        pid_action(s, pt, vt, at, a.env, a.max_thrust, a.Kp, a.Kd, a.Kp_att, act);
        for (int k = 0; k < 4; ++k) a.a_nom[((long long)t * a.H + h) * 4 + k] = act[k];
        quad_step(s, act, fd, a.env);
        fd[0] = fd[1] = fd[2] = 0.f;  // deterministic=True (controllers/covo.py:67-69)
        ++time;
        int row = min(time, a.traj_len - 1);
        gather3(a.pos_traj, row, pt);
        gather3(a.vel_traj, row, vt);
        gather3(a.acc_traj, row, at);
    }
}

__global__ void pid_policy_kernel(const PidArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n_env) return;
    QState<float> s;
    float fd[3], pt[3], vt[3], at[3] = {0.f, 0.f, 0.f};
    load_state24(a.state24 + (long long)e * kStateFloats, s, fd, pt, vt);
    if (a.acc_traj) gather3(a.acc_traj + (long long)e * a.traj_len * 3, min(max(a.time[e], 0), a.traj_len - 1), at);
    float act[4];
    pid_action(s, pt, vt, at, a.env, a.max_thrust, a.Kp, a.Kd, a.Kp_att, act, a.Ki, a.integral ? a.integral + 3 * e : nullptr);
    for (int k = 0; k < 4; ++k) a.action[4 * e + k] = act[k];
}

cudaError_t launch_pid(const PidArgs& a, cudaStream_t st) {
    pid_policy_kernel<<<(a.n_env + 63) / 64, 64, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_offline_paths(const OfflineArgs& a, cudaStream_t st) {
    pid_path_kernel<<<1, 32, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    pid_nominal_kernel<<<(a.t_sched + 63) / 64, 64, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace covo
