// Device-resident environment step: the caller side of the hot path, so that a whole closed loop
// (noisy state -> controller -> env step) runs without a host round trip.
//
// Replaces, per environment,
//   Quad3D.step_env            envs/quadrotor.py:215-248   reward and done of the PRE-step state, then the transition
//   free_dynamics_3d_bodyrate  dynamics/free.py:114-202    (the same templated quad_step the rollout kernel runs)
//   disturbances               dynamics/free.py:58-72      "none" -> 0, "gaussian" -> dyn_noise_scale * N(0, I)
//   Quad3D.get_info            envs/quadrotor.py:314-361   noisy_state = next_state + N(0, (obs_noise_scale * k)^2),
//                                                          k = 0.25 / 0.5 / 0.02 / 0.5 for pos / vel / quat / omega
// Not replicated: the auto-reset of BaseEnvironment.step (envs/base.py:27-38); episodes are bounded by the caller
// (the MPC harness runs exactly max_steps_in_episode steps per episode).
// Noise: Philox field keyed by (seed, step, environment), or normals supplied by the caller (parity tests).
#include <cuda_runtime.h>

#include "envstep.cuh"
#include "rng.cuh"

namespace covo {

__global__ void __launch_bounds__(128) env_step_kernel(const EnvStepArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n_env) return;
    float* sg = a.state24 + (long long)e * kStateFloats;
    QState<float> s;
    float fd[3], pt[3], vt[3];
    load_state24(sg, s, fd, pt, vt);
    int t = a.time[e];
    const unsigned int step = a.step_ctr ? __ldg(a.step_ctr) : 0u;
    const long long log_off = (long long)step * a.n_env;
    float z[kEnvNoiseFloats];
    if (a.noise_in) {
        const float* zin = a.noise_in + (log_off + e) * kEnvNoiseFloats;
#pragma unroll
        for (int k = 0; k < kEnvNoiseFloats; ++k) z[k] = zin[k];
    } else {
#pragma unroll
        for (int b = 0; b < kEnvNoiseFloats / 4; ++b) philox_normal4(a.seed, a.stream + step, (uint32_t)e, (uint32_t)b, z + 4 * b);
    }
    if (a.do_step) {
        // reward / done / err_pos of the PRE-step state (envs/quadrotor.py:243-244)
        const float ex = pt[0] - s.p[0], ey = pt[1] - s.p[1], ez = pt[2] - s.p[2];
        if (a.err_pos) a.err_pos[log_off + e] = sqrtf(ex * ex + ey * ey + ez * ez);
        if (a.reward) a.reward[log_off + e] = quad_reward(s, pt, vt);
        const bool quad_terminal_prestep = quad_terminal(s, t, a.env);
        if (a.done) a.done[e] = quad_terminal_prestep ? 1 : 0;
        const float* ag = a.action + (long long)e * 4;
        const float u[4] = {ag[0], ag[1], ag[2], ag[3]};
        if (a.action_log) {
            float* al = a.action_log + (log_off + e) * 4;
            al[0] = u[0];
            al[1] = u[1];
            al[2] = u[2];
            al[3] = u[3];
        }
        quad_step(s, u, fd, a.env);
        // f_disturb <- disturb_func (dynamics/free.py:144-147)
        for (int k = 0; k < 3; ++k) fd[k] = a.gaussian ? a.dyn_noise_scale * z[13 + k] : 0.f;
        t += 1;
        const int row = min(t, a.traj_len - 1);  // clamped gather, dynamics/free.py:153-155
        const float* pr = a.pos_traj + (long long)e * a.traj_stride + (long long)row * 3;
        const float* vr = a.vel_traj + (long long)e * a.traj_stride + (long long)row * 3;
        for (int k = 0; k < 3; ++k) {
            pt[k] = pr[k];
            vt[k] = vr[k];
        }
        if (a.reset_pool > 0 && quad_terminal_prestep) {
            // envs/base.py:27-38: done (of the pre-step state) -> the next reset_env draw replaces the stepped state
            const int k = a.reset_count[e] % a.reset_pool;
            a.reset_count[e] += 1;
            const float* rs = a.reset_state24 + ((long long)k * a.n_env + e) * kStateFloats;
            load_state24(rs, s, fd, pt, vt);
            t = a.reset_time[(long long)k * a.n_env + e];
            const long long tl = (long long)a.traj_len * 3;
            const float* rp = a.reset_pos_traj + ((long long)k * a.n_env + e) * tl;
            const float* rv = a.reset_vel_traj + ((long long)k * a.n_env + e) * tl;
            float* wp = a.traj_pos_rw + (long long)e * a.traj_stride;
            float* wv = a.traj_vel_rw + (long long)e * a.traj_stride;
            for (long long i = 0; i < tl; ++i) {
                wp[i] = rp[i];
                wv[i] = rv[i];
            }
            if (a.a_mean)
                for (int i = 0; i < a.n_mean; ++i) a.a_mean[(long long)e * a.n_mean + i] = a.a_mean_init[i];
        }
        for (int k = 0; k < 3; ++k) sg[k] = s.p[k];
        for (int k = 0; k < 4; ++k) sg[3 + k] = s.q[k];
        for (int k = 0; k < 3; ++k) sg[7 + k] = s.v[k];
        for (int k = 0; k < 3; ++k) sg[10 + k] = s.w[k];
        for (int k = 0; k < 3; ++k) sg[13 + k] = fd[k];
        for (int k = 0; k < 3; ++k) sg[16 + k] = pt[k];
        for (int k = 0; k < 3; ++k) sg[19 + k] = vt[k];
        a.time[e] = t;
    }
    if (a.mppi_fdist) {  // planning force of the next MPPI call: its own block of the noise field (never the caller-supplied normals)
        float zp[4];
        philox_normal4(a.seed, a.stream + step, (uint32_t)e, 4u, zp);
        float* fo = a.mppi_fdist + (long long)e * a.mppi_H * 3;
        for (int hh = 0; hh < a.mppi_H; ++hh)
            for (int k = 0; k < 3; ++k) fo[hh * 3 + k] = a.dyn_noise_scale * zp[k];
    }
    // info["noisy_state"] (envs/quadrotor.py:323-351)
    float* ng = a.noisy24 + (long long)e * kStateFloats;
    const float sc = a.obs_noise_scale;
    for (int k = 0; k < 3; ++k) ng[k] = s.p[k] + z[k] * (sc * 0.25f);
    for (int k = 0; k < 3; ++k) ng[7 + k] = s.v[k] + z[3 + k] * (sc * 0.5f);
    for (int k = 0; k < 4; ++k) ng[3 + k] = s.q[k] + z[6 + k] * (sc * 0.02f);
    for (int k = 0; k < 3; ++k) ng[10 + k] = s.w[k] + z[10 + k] * (sc * 0.5f);
    for (int k = 0; k < 3; ++k) ng[13 + k] = fd[k];
    for (int k = 0; k < 3; ++k) ng[16 + k] = pt[k];
    for (int k = 0; k < 3; ++k) ng[19 + k] = vt[k];
    ng[22] = ng[23] = 0.f;
    a.noisy_time[e] = t;
}

// Last node of a replayed step: advances the device step counters and -- for the host entry point -- hands the result over without
// a copy node or a stream synchronisation: the actions and the status words go to mapped pinned memory and, behind a system-scope
// fence in the same thread, the flag the host spins on is raised to the new step number.
__global__ void bump_kernel(unsigned int* ctr_a, unsigned int* ctr_b, const int* status_src, int* status_dst, int n_status, unsigned int* flag,
                            const float* action_src, float* action_dst) {
    unsigned int v = 0u;
    if (ctr_a) {
        v = *ctr_a + 1u;
        *ctr_a = v;
    }
    if (ctr_b) *ctr_b += 1u;
    if (flag) {
        for (int e = 0; e < n_status; ++e) {
            if (status_dst) status_dst[e] = status_src ? status_src[e] : 0;
            if (action_dst)
                for (int k = 0; k < 4; ++k) action_dst[4 * e + k] = action_src[4 * e + k];
        }
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(flag) = v;
    }
}

cudaError_t launch_bump(unsigned int* ctr_a, unsigned int* ctr_b, cudaStream_t st, const int* status_src, int* status_dst, int n_status,
                        unsigned int* flag, const float* action_src, float* action_dst) {
    bump_kernel<<<1, 1, 0, st>>>(ctr_a, ctr_b, status_src, status_dst, n_status, flag, action_src, action_dst);
    return cudaGetLastError();
}

cudaError_t launch_env_step(const EnvStepArgs& a, cudaStream_t st) {
    env_step_kernel<<<(a.n_env + 127) / 128, 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace covo
