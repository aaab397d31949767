// C-ABI (include/covo_b200.h): handle, device workspace and the per-mode step schedules.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/covo_b200.h"
#include "common.cuh"
#include "hessian.cuh"
#include "offline.cuh"
#include "rng.cuh"
#include "sigma.cuh"
#include "envstep.cuh"

using namespace covo;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) return fail(COVO_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        n = count;
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) e = cudaMemset(p, 0, count * sizeof(T));
        if (e == cudaSuccess) e = cudaDeviceSynchronize();  // the handle's stream is non-blocking
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

__global__ void shift_blocks_kernel(float* Lblk, float* cov, int H) {
    // a_cov <- concat(a_cov[1:], a_cov[-1:])  (controllers/mppi.py:46-49), same for its factor
    const int env = blockIdx.x;
    float* L = Lblk + (long long)env * H * 16;
    float* C = cov + (long long)env * H * 16;
    extern __shared__ float sh[];
    for (int i = threadIdx.x; i < H * 16; i += blockDim.x) {
        sh[i] = L[i];
        sh[H * 16 + i] = C[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H * 16; i += blockDim.x) {
        int h = i >> 4, r = i & 15;
        int hs = min(h + 1, H - 1);
        L[i] = sh[hs * 16 + r];
        C[i] = sh[H * 16 + hs * 16 + r];
    }
}

// MPPI covariance update (controllers/mppi.py:119-125, gamma_sigma != 0):
//     a_cov[h] <- gamma_sigma * sum_i w_i (a_i[h] - a_mean[h]) (a_i[h] - a_mean[h])^T + (1 - gamma_sigma) * a_cov[h]
// with w the normalised softmax weights of this step, a_i the CLIPPED samples (:66), a_mean the UPDATED (blended) mean and a_cov the
// shifted covariance the samples were drawn from; then the 4 x 4 Cholesky factor the next step's sampler uses (:59).  One CTA per
// (horizon step, environment) over the per-sample costs and samples the rollout kernel leaves behind in this mode.  A pivot that is
// not positive (the weighted covariance has rank ~ESS: with lambda = 0.01 the reference's own Cholesky returns NaN there) sets status 2.
__global__ void __launch_bounds__(256) mppi_cov_update_kernel(const float* __restrict__ costs, const float* __restrict__ samples,
                                                              const float* __restrict__ a_mean, float* cov, float* Lblk, int* status, int N, int H,
                                                              float inv_lam, float gamma_sigma) {
    const int h = blockIdx.x, env = blockIdx.y, tid = threadIdx.x, n = 4 * H;
    const float* c_e = costs + (long long)env * N;
    const float* s_e = samples + (long long)env * N * n + 4 * h;
    __shared__ float red[8][12];
    float m = CUDART_INF_F;
    for (int i = tid; i < N; i += 256) m = fminf(m, c_e[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) red[tid >> 5][0] = m;
    __syncthreads();
    m = red[0][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fminf(m, red[w][0]);
    __syncthreads();
    const float4 mu = *reinterpret_cast<const float4*>(a_mean + (long long)env * n + 4 * h);
    float acc[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) acc[k] = 0.f;
    for (int i = tid; i < N; i += 256) {
        const float c = c_e[i];
        const float w = (c < CUDART_INF_F) ? expf(-(c - m) * inv_lam) : 0.f;
        const float4 a = *reinterpret_cast<const float4*>(s_e + (long long)i * n);
        const float d[4] = {a.x - mu.x, a.y - mu.y, a.z - mu.z, a.w - mu.w};
        acc[0] += w;
        int k = 1;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q <= r; ++q) acc[k++] = fmaf(w * d[r], d[q], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5][k] = v;
    }
    __syncthreads();
    if (tid == 0) {
        float tot[11];
        for (int k = 0; k < 11; ++k) {
            float v = 0.f;
            for (int w = 0; w < 8; ++w) v += red[w][k];
            tot[k] = v;
        }
        float* C = cov + ((long long)env * H + h) * 16;
        float* L = Lblk + ((long long)env * H + h) * 16;
        const float inv_s = tot[0] > 0.f ? 1.f / tot[0] : 0.f;
        float Cn[4][4];
        int k = 1;
        for (int r = 0; r < 4; ++r)
            for (int q = 0; q <= r; ++q) {
                const float v = gamma_sigma * tot[k++] * inv_s + (1.f - gamma_sigma) * C[r * 4 + q];
                Cn[r][q] = Cn[q][r] = v;
            }
        float Lf[4][4] = {};
        bool bad = false;
        for (int r = 0; r < 4; ++r)
            for (int q = 0; q <= r; ++q) {
                float v = Cn[r][q];
                for (int t = 0; t < q; ++t) v -= Lf[r][t] * Lf[q][t];
                if (r == q) {
                    if (!(v > 0.f)) {
                        bad = true;
                        v = 1e-30f;
                    }
                    Lf[r][r] = sqrtf(v);
                } else {
                    Lf[r][q] = v / Lf[q][q];
                }
            }
        for (int r = 0; r < 4; ++r)
            for (int q = 0; q < 4; ++q) {
                C[r * 4 + q] = Cn[r][q];
                L[r * 4 + q] = Lf[r][q];
            }
        if (bad && status) status[env] = 2;
    }
}

__global__ void debug_eps_kernel(float* out, int n_local, int offset, int n, unsigned long long seed, unsigned int stream) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int blocks4 = n >> 2;
    if (idx >= n_local * blocks4) return;
    int s = idx / blocks4, b = idx % blocks4;
    float z[4];
    philox_normal4(seed, stream, (uint32_t)(offset + s), (uint32_t)b, z);
    for (int j = 0; j < 4; ++j) out[(long long)s * n + 4 * b + j] = z[j];
}

}  // namespace

struct covo_handle {
    covo_config cfg;
    EnvConsts env;
    int H, n, n_pad, E, T;
    int n_local, sample_offset, n_cta, rec;
    size_t lt_floats;
    cudaStream_t own_stream = nullptr;
    cudaStream_t aux_stream = nullptr;  // side stream of the covariance step (Q accumulation next to E2)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // Cholesky -> rollout pipeline: the rollout kernel runs next to the factorisation and consumes the factor column block
    // by column block (RolloutArgs::lfac_progress)
    DevBuf<int> chol_progress;  // [E] monotone counter: epoch + finished column blocks
    unsigned int chol_epoch = 64;  // wraps; compared as a signed difference on the device
    int num_sms = 0;
    bool pipeline_enabled = true;
    bool pipeline_forced = false;
    int sigma_dense = 0;  // optimize_sigma path: 0 tridiagonal (E1-E3, sigma.cu); 3 the dense kernels D1-D3 of sigma_dense.cu
    DevBuf<double> dense_scal;
    DevBuf<float> dense_X;  // COVO_PIPELINE=2 (development): keep the pipeline on while per-kernel timings are taken
    unsigned int rng_stream = 0;
    bool jax_key_pending = false;  // covo_set_jax_key: the next sampling launch uses the JAX-compatible stream
    unsigned int jax_key[2] = {0, 0};
    bool pos_stats_on = false;
    bool profiling = false;
    bool have_factor = false;
    int t_sched = 0;
    cudaEvent_t ev[8] = {};
    float kernel_ms[6] = {0, 0, 0, 0, 0, 0};
    // inputs
    DevBuf<float> state24, pos_traj, vel_traj, acc_traj, a_mean, eps, fdist;
    bool fdist_on = false;  // MPPI under disturb_type gaussian: the rollouts of the step see the force set by covo_set_rollout_disturbance
    DevBuf<int> time;
    // covariance pipeline
    DevBuf<float> R, Vh, tau, Qt, F, cov, Lfull, Lt, Lblk, hess_ws;
    DevBuf<double> diag, zolo;
    DevBuf<int> status;
    // offline schedule (batched over schedule steps)
    DevBuf<float> cov_table, Lt_table, sched_states, sched_anom, sched_R, sched_Vh, sched_tau, sched_Qt, sched_F, sched_ws, sched_disturb;
    DevBuf<double> sched_diag;
    DevBuf<int> sched_times, sched_status;
    // rollout
    DevBuf<float> partials, rank_partial, action, costs, samples, pos_stats, gathered_scratch;
    // fused exchange of the N-sharded step (world > 1): [2][world][E][rec] floats followed by [2][world][E] flags, one allocation
    // (one CUDA IPC handle); xpeer[w] = rank w's buffer as this process sees it (IPC mapping, or a pointer of the same process)
    DevBuf<float> xchg;
    float* xpeer[kMaxPeers] = {};
    bool xpeer_ipc[kMaxPeers] = {};
    unsigned int xchg_step = 0;  // steps taken through covo_step_sharded_device
    DevBuf<unsigned int> counters;
    DevBuf<long long> prof;
    bool phase_clocks = false;
    // device-resident environment (caller side of the hot path)
    DevBuf<float> env_state24, env_noisy24, env_noise, env_log_f, env_action, pid_integral;
    DevBuf<float> pool_state24, pool_pos, pool_vel, mean_init;  // auto-reset pool (covo_env_set_reset_pool)
    DevBuf<int> pool_time, pool_count;
    int pool_n = 0;
    bool pool_reset_mean = false;
    DevBuf<int> env_time, env_noisy_time, env_done;
    bool env_ready = false;
    // CUDA-graph replay of the step (one graph launch per MPC step instead of 8-10 kernel launches, event records and waits).
    // Everything that used to change from launch to launch now lives on the device: the sample-field step counter (rng_ctr), the
    // Cholesky -> rollout progress counter (cleared by the first kernel of the step), the closed loop's step index (loop_ctr).
    // What still differs between calls -- the caller's state / time / action pointers of covo_step_device -- is patched into the two
    // kernel nodes that carry it.
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        cudaGraph_t graph = nullptr;  // kept: its node handles are the keys for patching the executable graph
        cudaGraphNode_t n_hess = nullptr, n_roll = nullptr;
        HessianArgs ha;
        RolloutArgs ra;
        const float* st_d = nullptr;
        const int* tm_d = nullptr;
        float* act_d = nullptr;
        bool valid = false;
        bool zero_copy = false;
    };
    StepGraph g_dev, g_host, g_loop;
    bool graphs_enabled = true;
    int direct_steps = 0;
    DevBuf<unsigned int> dev_ctr;      // [0] sample-field step counter, [1] closed-loop step index
    unsigned int rng_ctr_shadow = 0;   // what dev_ctr[0] holds (host mirror)
    EnvStepArgs loop_env_args;         // the environment-step node of g_loop
    bool loop_args_valid = false;
    // pinned staging
    float* h_state = nullptr;
    int* h_time = nullptr;
    float* h_action = nullptr;
    int* h_status = nullptr;
    // The staging buffers are MAPPED pinned memory.  In the replayed graph of covo_step the kernels read the state / time straight
    // through these device aliases (zero-copy over PCIe: 100 bytes), the last node stores the action and the status words there and
    // raises h_flag to the step number, and the host spins on h_flag instead of synchronising the stream: no copy nodes, no wake-up
    // through the driver.
    float* dh_state = nullptr;
    int* dh_time = nullptr;
    float* dh_action = nullptr;
    int* dh_status = nullptr;
    unsigned int* h_flag = nullptr;
    unsigned int* dh_flag = nullptr;
    unsigned int host_epoch = 0;  // value the flag will hold when the step in flight is complete
    bool zero_copy = true;
    bool flag_pending = false;  // the step in flight signals its completion through h_flag
};


// All host<->device traffic goes through the handle's own (non-blocking) stream: a plain cudaMemcpy on the
// legacy default stream is NOT ordered against it (a staged pageable H2D copy may still be in flight).
static cudaError_t h2d(covo_handle* h, void* dst, const void* src, size_t bytes);
static cudaError_t d2h(covo_handle* h, void* dst, const void* src, size_t bytes);
static cudaError_t dzero(covo_handle* h, void* dst, size_t bytes);

namespace {

void release_all(covo_handle* h) {
    h->state24.release(); h->pos_traj.release(); h->vel_traj.release(); h->acc_traj.release();
    h->a_mean.release(); h->eps.release(); h->fdist.release(); h->time.release();
    h->R.release(); h->Vh.release(); h->tau.release(); h->Qt.release(); h->F.release(); h->cov.release();
    h->Lfull.release(); h->Lt.release(); h->Lblk.release(); h->hess_ws.release(); h->diag.release();
    h->zolo.release(); h->status.release();
    h->cov_table.release(); h->Lt_table.release(); h->sched_states.release(); h->sched_anom.release(); h->sched_disturb.release();
    h->sched_R.release(); h->sched_Vh.release(); h->sched_tau.release(); h->sched_Qt.release(); h->sched_F.release();
    h->sched_ws.release(); h->sched_diag.release(); h->sched_times.release(); h->sched_status.release();
    for (int w = 0; w < kMaxPeers; ++w)
        if (h->xpeer[w] && h->xpeer_ipc[w]) cudaIpcCloseMemHandle(h->xpeer[w]);
    h->xchg.release();
    h->partials.release(); h->rank_partial.release(); h->action.release(); h->costs.release();
    h->env_state24.release(); h->env_noisy24.release(); h->env_noise.release(); h->env_log_f.release(); h->env_action.release();
    h->env_time.release(); h->env_noisy_time.release(); h->env_done.release();
    h->prof.release(); h->samples.release(); h->pos_stats.release(); h->gathered_scratch.release(); h->counters.release();
    if (h->h_state) cudaFreeHost(h->h_state);
    if (h->h_time) cudaFreeHost(h->h_time);
    if (h->h_action) cudaFreeHost(h->h_action);
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->h_flag) cudaFreeHost(h->h_flag);
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    h->chol_progress.release();
    h->dense_scal.release();
    h->dense_X.release();
    h->pid_integral.release();
    h->pool_state24.release(); h->pool_pos.release(); h->pool_vel.release(); h->mean_init.release(); h->pool_time.release(); h->pool_count.release();
    h->dev_ctr.release();
    for (covo_handle::StepGraph* g : {&h->g_dev, &h->g_host, &h->g_loop}) {
        if (g->exec) cudaGraphExecDestroy(g->exec);
        if (g->graph) cudaGraphDestroy(g->graph);
    }
}

HessianArgs hess_args(covo_handle* h, const float* st, const int* tm, const float* a_mean, int shift, float* R, float* ws,
                      long long traj_stride) {
    HessianArgs a;
    a.H = h->H;
    a.traj_len = h->T;
    a.shift = shift;
    a.traj_stride = traj_stride;
    a.env = h->env;
    a.state24 = st;
    a.time = tm;
    a.pos_traj = h->pos_traj.p;
    a.vel_traj = h->vel_traj.p;
    a.a_mean = a_mean;
    a.workspace = ws;
    a.R = R;
    a.status = (R == h->R.p) ? h->status.p : nullptr;  // the offline schedule keeps its own status array
    a.progress = (R == h->R.p) ? h->chol_progress.p : nullptr;
    a.prof = h->phase_clocks ? h->prof.p : nullptr;
    return a;
}

SigmaArgs sigma_args(covo_handle* h) {
    SigmaArgs a;
    a.n = h->n;
    a.n_pad = h->n_pad;
    a.sample_sigma = h->cfg.sample_sigma;
    a.R = h->R.p;
    a.Vh = h->Vh.p;
    a.tau = h->tau.p;
    a.Qt = h->Qt.p;
    a.F = h->F.p;
    a.cov = h->cov.p;
    a.L = h->Lfull.p;
    a.Lt = h->Lt.p;
    a.diag = h->diag.p;
    a.zolo = h->zolo.p;
    a.status = h->status.p;
    a.lt_stride = (long long)h->lt_floats;
    a.prof = h->phase_clocks ? h->prof.p : nullptr;
    return a;
}

RolloutArgs rollout_args(covo_handle* h, const float* st, const int* tm, const float* a_in, int shift, const float* eps,
                         const float* fdist, float* a_out, float* act, float* costs, float* samples, int finalize) {
    RolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.n_samples = h->n_local;
    a.sample_offset = h->sample_offset;
    a.H = h->H;
    a.n = h->n;
    a.n_pad = h->n_pad;
    a.mode = (h->cfg.mode == COVO_MODE_MPPI) ? 1 : 0;
    a.shift = shift;
    a.finalize = finalize;
    a.traj_len = h->T;
    a.traj_stride = (long long)h->T * 3;
    a.lam = h->cfg.lam;
    a.gamma_mean = h->cfg.gamma_mean;
    a.discount = h->cfg.discount;
    a.env = h->env;
    a.seed = h->cfg.seed;
    a.stream = h->rng_stream;
    if (h->jax_key_pending && !eps) {  // one-shot: consumed by this launch
        a.rng_kind = 1;
        a.seed = (unsigned long long)h->jax_key[0] | ((unsigned long long)h->jax_key[1] << 32);
        a.n_total = h->cfg.n_samples;
    }
    h->jax_key_pending = false;
    a.state24 = st;
    a.time = tm;
    a.pos_traj = h->pos_traj.p;
    a.vel_traj = h->vel_traj.p;
    a.a_mean_in = a_in;
    a.eps = eps;
    a.fdist_seq = fdist;
    a.partials = h->partials.p;
    a.counters = h->counters.p;
    a.rank_partial = h->rank_partial.p;
    a.a_mean_out = a_out;
    a.action_out = act;
    a.costs_out = costs;
    a.samples_out = samples;
    a.pos_stats = h->pos_stats_on ? h->pos_stats.p : nullptr;
    a.prof = h->phase_clocks ? h->prof.p : nullptr;
    if (a.mode == 1) {
        a.Lfac = h->Lblk.p;
        a.lfac_stride = (long long)h->H * 16;
    } else if (h->cfg.mode == COVO_MODE_COVO_OFFLINE && h->t_sched > 0) {
        a.Lfac = h->Lt_table.p;
        a.lfac_stride = 0;
        a.lfac_time_stride = (long long)h->lt_floats;
        a.lfac_time_max = h->t_sched - 1;
    } else {
        a.Lfac = h->Lt.p;
        a.lfac_stride = (long long)h->lt_floats;
    }
    return a;
}

struct Prof {
    covo_handle* h;
    cudaStream_t st;
    int slot = 0;
    Prof(covo_handle* h_, cudaStream_t s) : h(h_), st(s) {
        if (h->profiling) cudaEventRecord(h->ev[0], st);
    }
    void mark(int idx) {  // idx-th boundary (1..7)
        if (h->profiling) cudaEventRecord(h->ev[idx], st);
    }
};

// covariance step for the online mode: R (already in h->R) -> cov -> factor
// The rollout kernel may run NEXT TO the Cholesky kernel (programmatic dependent launch: it starts once every Cholesky CTA is
// resident and has released it) when both grids fit on the device together with room to spare -- every CTA of either kernel on
// an SM of its own in the worst case -- and nothing asks for per-kernel timings.
bool pipeline_ok(covo_handle* h) {
    return h->pipeline_enabled && ((!h->profiling && !h->phase_clocks) || h->pipeline_forced) && h->cfg.world == 1 &&
           rollout_is_overlapped(h->n_pad, 0, h->H) && (long long)(h->n_cta + 1) * h->E <= h->num_sms;
}

int run_sigma_chol(covo_handle* h, cudaStream_t st, Prof* pf, bool want_L = false, bool pipelined = false) {
    SigmaArgs sa = sigma_args(h);
    if (!want_L) sa.L = nullptr;  // the sampler only needs the packed factor
    if (h->sigma_dense) {  // Lanczos + 17 shifted inverses + combine instead of E1-E3
        const bool tm = pf && h->profiling;
        CK(launch_sigma_dense(sa, h->dense_scal.p, h->dense_X.p, h->E, st, tm ? h->ev[2] : nullptr, tm ? h->ev[3] : nullptr));
        if (pf) pf->mark(4);  // slots: 1 Lanczos, 2 inverses, 3 combine
    } else {
        CK(launch_tridiag(sa, h->E, st));
        if (pf) pf->mark(2);
        // Q^T from the reflectors does not depend on the tridiagonal matrix function: it runs on a side stream, in
        // the shadow of E2, and joins before the sandwich kernel
        CK(cudaEventRecord(h->ev_fork, st));
        CK(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
        CK(launch_qacc(sa, h->E, h->aux_stream));
        CK(cudaEventRecord(h->ev_join, h->aux_stream));
        CK(launch_trifunc(sa, h->E, st));
        if (pf) pf->mark(3);
        CK(cudaStreamWaitEvent(st, h->ev_join, 0));
        CK(launch_sandwich(sa, h->E, st));
        if (pf) pf->mark(4);
    }
    sa.cov_symmetric = 1;
    if (pipelined) {  // the rollout kernel follows in the same stream as a programmatic dependent launch (step_common)
        sa.progress = h->chol_progress.p;
        sa.epoch = 0;  // the counter is cleared by the first kernel of the step (hess_local)
    }
    CK(launch_cholesky(sa, h->E, st));
    if (pf) pf->mark(5);
    h->have_factor = true;
    return COVO_OK;
}

// steps whose covariance work can end with a numeric status: CoVO-online (optimize_sigma, Cholesky), MPPI with gamma_sigma != 0
static bool step_has_status(const covo_handle* h) {
    return h->cfg.mode == COVO_MODE_COVO_ONLINE || (h->cfg.mode == COVO_MODE_MPPI && h->cfg.gamma_sigma != 0.f);
}

// The launch sequence of one MPC step on stream `st`.  rec != nullptr: the step is being captured into a CUDA graph -- the sample
// field is indexed by the device counter and the arguments of the two kernels that carry caller pointers are kept for patching.
int step_launch(covo_handle* h, const float* st_d, const int* tm_d, const float* eps_d, float* act_d, cudaStream_t st, int finalize,
                covo_handle::StepGraph* rec) {
    Prof pf(h, st);
    const int mode = h->cfg.mode;
    bool pipelined = false;
    if (h->pos_stats_on) CK(cudaMemsetAsync(h->pos_stats.p, 0, h->pos_stats.n * sizeof(float), st));
    if (mode == COVO_MODE_MPPI) {
        shift_blocks_kernel<<<h->E, 128, 2 * h->H * 16 * sizeof(float), st>>>(h->Lblk.p, h->cov.p, h->H);
        CK(cudaGetLastError());
        for (int i = 1; i <= 5; ++i) pf.mark(i);
    } else if (mode == COVO_MODE_COVO_ONLINE) {
        HessianArgs ha = hess_args(h, st_d, tm_d, h->a_mean.p, 1, h->R.p, h->hess_ws.p, (long long)h->T * 3);
        if (rec) rec->ha = ha;
        CK(launch_hessian(ha, h->E, st));
        pf.mark(1);
        pipelined = pipeline_ok(h);
        int rc = run_sigma_chol(h, st, &pf, false, pipelined);
        if (rc) return rc;
    } else {
        if (h->t_sched <= 0) return fail(COVO_ERR_INVALID, "covo-offline: no schedule; call covo_reset_offline or covo_set_cov_offline first");
        for (int i = 1; i <= 5; ++i) pf.mark(i);
    }
    const bool cov_update = mode == COVO_MODE_MPPI && h->cfg.gamma_sigma != 0.f;  // mppi.py:119-125 needs the samples and their costs
    const float* fdist = (mode == COVO_MODE_MPPI && h->fdist_on) ? h->fdist.p : nullptr;  // mppi.py:74: stochastic step_env in the rollouts
    RolloutArgs ra = rollout_args(h, st_d, tm_d, h->a_mean.p, 1, eps_d, fdist, h->a_mean.p, act_d, cov_update ? h->costs.p : nullptr,
                                  cov_update ? h->samples.p : nullptr, finalize);
    if (pipelined) {
        ra.lfac_progress = h->chol_progress.p;
        ra.lfac_epoch = 0;
    }
    if (rec) {
        ra.stream = 0;
        ra.stream_ctr = h->dev_ctr.p;
        rec->ra = ra;
    }
    CK(launch_rollout(ra, h->E, st));
    if (cov_update) {
        CK(cudaMemsetAsync(h->status.p, 0, (size_t)h->E * sizeof(int), st));
        mppi_cov_update_kernel<<<dim3(h->H, h->E), 256, 0, st>>>(h->costs.p, h->samples.p, h->a_mean.p, h->cov.p, h->Lblk.p, h->status.p,
                                                                  h->n_local, h->H, 1.0f / h->cfg.lam, h->cfg.gamma_sigma);
        CK(cudaGetLastError());
    }
    pf.mark(6);
    return COVO_OK;
}

// node handles of the two kernels whose arguments carry caller pointers
int graph_find_nodes(cudaGraph_t g, covo_handle::StepGraph* sg) {
    size_t n = 0;
    CK(cudaGraphGetNodes(g, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    CK(cudaGraphGetNodes(g, nodes.data(), &n));
    for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType ty;
        CK(cudaGraphNodeGetType(nd, &ty));
        if (ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp;
        CK(cudaGraphKernelNodeGetParams(nd, &kp));
        if (kp.func == hess_local_kernel_address()) sg->n_hess = nd;
        else if (kp.func == rollout_kernel_address()) sg->n_roll = nd;
    }
    return COVO_OK;
}

int graph_patch_node(cudaGraphExec_t exec, cudaGraphNode_t nd, void* args_struct) {
    cudaKernelNodeParams kp;
    CK(cudaGraphKernelNodeGetParams(nd, &kp));
    void* params[1] = {args_struct};
    kp.kernelParams = params;
    kp.extra = nullptr;
    CK(cudaGraphExecKernelNodeSetParams(exec, nd, &kp));
    return COVO_OK;
}

// kernel arguments baked into the graphs changed (model constants, schedule table): rebuild on next use
void graphs_invalidate(covo_handle* h) { h->g_dev.valid = h->g_host.valid = h->g_loop.valid = false; }

bool graph_ok(covo_handle* h, const float* eps_d) {
    return h->graphs_enabled && !h->profiling && !h->phase_clocks && !eps_d && !h->jax_key_pending && !h->pos_stats_on;
}

// keeps the device copy of the sample-field counter equal to the host's (they part only when direct launches were mixed in)
int sync_rng_counter(covo_handle* h, cudaStream_t st) {
    if (h->rng_ctr_shadow != h->rng_stream) {
        CK(cudaMemcpyAsync(h->dev_ctr.p, &h->rng_stream, sizeof(unsigned int), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));  // the source is a host variable that keeps changing
        h->rng_ctr_shadow = h->rng_stream;
    }
    return COVO_OK;
}

// Capture the step (plus, for the host entry point, its H2D / D2H copies; plus, for the closed loop, the environment step) into sg.
int graph_build(covo_handle* h, covo_handle::StepGraph* sg, const float* st_d, const int* tm_d, float* act_d, int finalize, int kind) {
    if (sg->exec) {
        cudaGraphExecDestroy(sg->exec);
        sg->exec = nullptr;
    }
    if (sg->graph) {
        cudaGraphDestroy(sg->graph);
        sg->graph = nullptr;
    }
    sg->n_hess = sg->n_roll = nullptr;
    cudaStream_t cs = h->own_stream;
    CK(cudaStreamSynchronize(cs));
    CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int rc = COVO_OK;
    cudaError_t ce = cudaSuccess;
    const size_t E = (size_t)h->E;
    const bool zc = kind == 1 && h->zero_copy;
    if (kind == 1 && !zc) {  // covo_step: pinned staging -> device
        ce = cudaMemcpyAsync(h->state24.p, h->h_state, E * kStateFloats * sizeof(float), cudaMemcpyHostToDevice, cs);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h->time.p, h->h_time, E * sizeof(int), cudaMemcpyHostToDevice, cs);
    }
    if (zc) {  // the kernels read the staged state / time through the device aliases of the pinned buffers; the last node hands the result back
        st_d = h->dh_state;
        tm_d = h->dh_time;
    }
    if (ce == cudaSuccess) rc = step_launch(h, st_d, tm_d, nullptr, act_d, cs, finalize, sg);
    if (ce == cudaSuccess && rc == COVO_OK && kind == 2) ce = launch_env_step(h->loop_env_args, cs);
    if (ce == cudaSuccess && rc == COVO_OK)
        ce = launch_bump(h->dev_ctr.p, kind == 2 ? h->dev_ctr.p + 1 : nullptr, cs, zc && step_has_status(h) ? h->status.p : nullptr, zc ? h->dh_status : nullptr,
                         (int)E, zc ? h->dh_flag : nullptr, zc ? h->action.p : nullptr, zc ? h->dh_action : nullptr);
    if (ce == cudaSuccess && rc == COVO_OK && kind == 1 && !zc) {
        ce = cudaMemcpyAsync(h->h_action, h->action.p, E * 4 * sizeof(float), cudaMemcpyDeviceToHost, cs);
        if (ce == cudaSuccess && step_has_status(h))
            ce = cudaMemcpyAsync(h->h_status, h->status.p, E * sizeof(int), cudaMemcpyDeviceToHost, cs);
    }
    cudaGraph_t g = nullptr;
    cudaError_t ee = cudaStreamEndCapture(cs, &g);
    if (rc != COVO_OK) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    if (ce != cudaSuccess || ee != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return fail(COVO_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ee));
    }
    rc = graph_find_nodes(g, sg);
    if (rc == COVO_OK) {
        ce = cudaGraphInstantiate(&sg->exec, g, 0);
        if (ce != cudaSuccess) rc = fail(COVO_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
    }
    if (rc != COVO_OK) {
        cudaGraphDestroy(g);
        return rc;
    }
    sg->graph = g;
    sg->st_d = zc ? h->state24.p : st_d;  // (the host entry point is recognised by these markers, see step_common)
    sg->tm_d = zc ? h->time.p : tm_d;
    sg->act_d = zc ? h->action.p : act_d;
    sg->zero_copy = zc;
    sg->valid = true;
    return COVO_OK;
}

int step_common(covo_handle* h, const float* st_d, const int* tm_d, const float* eps_d, float* act_d, cudaStream_t st,
                int finalize) {
    // (the first step of a handle is launched directly: it configures the kernels' shared-memory attributes, which a capture must not do)
    if (finalize == 1 && graph_ok(h, eps_d) && h->direct_steps > 0) {
        covo_handle::StepGraph* sg = (st_d == h->state24.p && act_d == h->action.p) ? &h->g_host : &h->g_dev;
        const int kind = (sg == &h->g_host) ? 1 : 0;
        if (!sg->valid) {
            int rc = graph_build(h, sg, st_d, tm_d, act_d, finalize, kind);
            if (rc == COVO_ERR_CUDA) {  // no graph on this driver / configuration: keep launching directly
                h->graphs_enabled = false;
                cudaGetLastError();
                return step_common(h, st_d, tm_d, eps_d, act_d, st, finalize);
            }
            if (rc) return rc;
        } else if (sg->st_d != st_d || sg->tm_d != tm_d || sg->act_d != act_d) {
            // other caller buffers than last time: patch the two kernel nodes that read / write them
            if (sg->n_hess) {
                sg->ha.state24 = st_d;
                sg->ha.time = tm_d;
                int rc = graph_patch_node(sg->exec, sg->n_hess, &sg->ha);
                if (rc) return rc;
            }
            sg->ra.state24 = st_d;
            sg->ra.time = tm_d;
            sg->ra.action_out = act_d;
            int rc = graph_patch_node(sg->exec, sg->n_roll, &sg->ra);
            if (rc) return rc;
            sg->st_d = st_d;
            sg->tm_d = tm_d;
            sg->act_d = act_d;
        }
        int rc = sync_rng_counter(h, st);
        if (rc) return rc;
        CK(cudaGraphLaunch(sg->exec, st));
        h->rng_stream += 1;
        h->rng_ctr_shadow += 1;
        h->have_factor = true;
        h->flag_pending = sg->zero_copy;
        if (sg->zero_copy) h->host_epoch = h->rng_stream;  // the counter kernel raises the flag to the new step number
        return COVO_OK;
    }
    const bool host_io = (st_d == h->state24.p && act_d == h->action.p && finalize == 1);
    if (host_io) {  // what the covo_step graph does inside: staging copies around the kernels
        const size_t E = (size_t)h->E;
        CK(cudaMemcpyAsync(h->state24.p, h->h_state, E * kStateFloats * sizeof(float), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->time.p, h->h_time, E * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    int rc = step_launch(h, st_d, tm_d, eps_d, act_d, st, finalize, nullptr);
    if (rc) return rc;
    if (host_io) {
        const size_t E = (size_t)h->E;
        CK(cudaMemcpyAsync(h->h_action, h->action.p, E * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (step_has_status(h)) CK(cudaMemcpyAsync(h->h_status, h->status.p, E * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    if (!eps_d) h->rng_stream += 1;
    h->direct_steps += 1;
    return COVO_OK;
}

}  // namespace


static cudaError_t h2d(covo_handle* h, void* dst, const void* src, size_t bytes) {
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->own_stream);
}
static cudaError_t d2h(covo_handle* h, void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->own_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->own_stream);
    return e;
}
static cudaError_t dzero(covo_handle* h, void* dst, size_t bytes) { return cudaMemsetAsync(dst, 0, bytes, h->own_stream); }

// =================================================================================================
extern "C" {

const char* covo_last_error(void) { return g_err.c_str(); }
const char* covo_version(void) { return "covo_b200 0.1 (sm_100a)"; }

int covo_default_config(covo_config* c) {
    if (!c) return fail(COVO_ERR_INVALID, "null config");
    memset(c, 0, sizeof(*c));
    c->mode = COVO_MODE_COVO_ONLINE;
    c->n_samples = 8192;  // envs/quadrotor.py:673-676
    c->horizon = 32;
    c->n_env = 1;
    c->traj_len = 300;
    c->device = 0;
    c->rank = 0;
    c->world = 1;
    c->lam = 0.01f;
    c->sample_sigma = 0.5f;
    c->gamma_mean = 1.0f;
    c->gamma_sigma = 0.0f;
    c->discount = 1.0f;
    c->m = 0.027f;  // dynamics/dataclass.py:43-49, 71, 76, 81
    c->g = 9.81f;
    c->max_thrust = 0.8f;
    c->dt = 0.02f;
    c->alpha_bodyrate = 0.5f;
    c->action_scale = 1.0f;
    c->pos_limit = 3.0f;
    c->max_omega[0] = 10.f;
    c->max_omega[1] = 10.f;
    c->max_omega[2] = 3.f;
    c->max_steps_in_episode = 300;
    c->seed = 0;
    return COVO_OK;
}

int covo_create(const covo_config* cfg, covo_handle** out) {
    if (!cfg || !out) return fail(COVO_ERR_INVALID, "null argument");
    if (cfg->mode < 0 || cfg->mode > 2) return fail(COVO_ERR_NOT_IMPLEMENTED, "unknown mode %d", cfg->mode);
    if (cfg->horizon < 2 || cfg->horizon > kMaxH) return fail(COVO_ERR_INVALID, "horizon must be in [2, %d]", kMaxH);
    if (cfg->mode != COVO_MODE_MPPI && 4 * cfg->horizon > kSigmaMaxN)
        return fail(COVO_ERR_INVALID, "CoVO modes support 4*H <= %d (H <= %d)", kSigmaMaxN, kSigmaMaxN / 4);
    if (cfg->n_samples < 1 || cfg->n_env < 1 || cfg->traj_len < 1) return fail(COVO_ERR_INVALID, "sizes must be positive");
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return fail(COVO_ERR_INVALID, "bad rank/world");
    if (cfg->n_samples % cfg->world) return fail(COVO_ERR_INVALID, "n_samples must divide evenly across world");
    if (cfg->gamma_sigma != 0.f && cfg->mode == COVO_MODE_MPPI && cfg->world != 1)
        return fail(COVO_ERR_NOT_IMPLEMENTED, "gamma_sigma != 0 (MPPI covariance update) with the sample axis sharded (world > 1) is not implemented");
    if (cfg->gamma_sigma < 0.f || cfg->gamma_sigma > 1.f) return fail(COVO_ERR_INVALID, "gamma_sigma must be in [0, 1]");
    if (cfg->mode == COVO_MODE_COVO_OFFLINE && cfg->n_env != 1)
        return fail(COVO_ERR_NOT_IMPLEMENTED, "covo-offline supports n_env == 1");
    if (!(cfg->lam > 0.f)) return fail(COVO_ERR_INVALID, "lam must be positive");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(COVO_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(ce));
    CK(cudaSetDevice(cfg->device));
    covo_handle* h = new covo_handle();
    h->cfg = *cfg;
    h->env.m = cfg->m;
    h->env.g = cfg->g;
    h->env.max_thrust = cfg->max_thrust;
    h->env.dt = cfg->dt;
    h->env.alpha_bodyrate = cfg->alpha_bodyrate;
    h->env.action_scale = cfg->action_scale;
    h->env.pos_limit = cfg->pos_limit;
    for (int k = 0; k < 3; ++k) h->env.max_omega[k] = cfg->max_omega[k];
    h->env.max_steps = cfg->max_steps_in_episode;
    h->H = cfg->horizon;
    h->n = 4 * h->H;
    h->n_pad = round_up8(h->n);
    h->E = cfg->n_env;
    h->T = cfg->traj_len;
    h->n_local = cfg->n_samples / cfg->world;
    h->sample_offset = cfg->rank * h->n_local;
    h->n_cta = (h->n_local + kTileSamples - 1) / kTileSamples;
    h->rec = kPartialHdr + h->n_pad;
    h->lt_floats = (size_t)lt_size(h->n, h->n_pad);
    const size_t E = h->E, n = h->n, nn = n * n;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    A(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    {
        A(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));
        A(h->chol_progress.alloc(E));
        if (e == cudaSuccess) A(cudaMemset(h->chol_progress.p, 0, E * sizeof(int)));
        const char* ge = getenv("COVO_GRAPH");
        h->graphs_enabled = !(ge && ge[0] == '0');
        A(h->dev_ctr.alloc(2));
        const char* pe = getenv("COVO_PIPELINE");
        h->pipeline_enabled = !(pe && pe[0] == '0');
        h->pipeline_forced = pe && pe[0] == '2';
        // optimize_sigma path.  Single environments default to the DENSE kernels D1-D3 of sigma_dense.cu (adaptive Lanczos in float64,
        // A^(-1/2) as 13 float64 Gauss-Jordan pole inverses: Sigma within 1e-7 .. 2e-7 of the float64 eigen-decomposition, 0.1 ms less
        // per step at H = 50); batches of environments use the tridiagonal kernels E1-E3 of sigma.cu (orthogonal reduction in float32,
        // everything on the tridiagonal in float64; 1e-6 .. 1e-5), which need 8 CTAs per matrix instead of 112.  COVO_SIGMA=tridiag /
        // COVO_SIGMA=dense or covo_set_sigma_path(h, 0 / 3) override.  A Lanczos stage that does not converge (status 3: no separated
        // lowest eigenvalue, not a CoVO Hessian) moves the handle to E1-E3.
        const char* se = getenv("COVO_SIGMA");
        h->sigma_dense = 0;
        if (cfg->mode != COVO_MODE_MPPI && h->n <= kSigmaMaxN) {
            if (se && strncmp(se, "dense", 5) == 0) h->sigma_dense = 3;
            else if (se && strncmp(se, "tridiag", 7) == 0) h->sigma_dense = 0;
            else if (E == 1) h->sigma_dense = 3;
        }
        if (h->sigma_dense) {
            A(h->dense_scal.alloc(E * 4));
            A(h->dense_X.alloc(E * sigma_dense_scratch_floats(h->n)));
        }
    }
    for (auto& ev : h->ev) A(cudaEventCreate(&ev));
    A(h->state24.alloc(E * kStateFloats));
    A(h->time.alloc(E));
    A(h->pos_traj.alloc(E * h->T * 3));
    A(h->vel_traj.alloc(E * h->T * 3));
    A(h->acc_traj.alloc(E * h->T * 3));
    A(h->a_mean.alloc(E * n));
    A(h->partials.alloc(E * h->n_cta * h->rec));
    A(h->rank_partial.alloc(E * h->rec));
    if (cfg->world > 1 && cfg->world <= kMaxPeers) {
        A(h->xchg.alloc((size_t)2 * cfg->world * E * (h->rec + 1)));
        h->xpeer[cfg->rank] = h->xchg.p;
    }
    A(h->counters.alloc(E));
    A(h->action.alloc(E * 4));
    A(h->pos_stats.alloc(E * h->H * 6));
    A(h->status.alloc(E));
    A(h->prof.alloc(64));
    if (cfg->mode == COVO_MODE_MPPI) {
        A(h->Lblk.alloc(E * h->H * 16));
        A(h->cov.alloc(E * h->H * 16));
        if (cfg->gamma_sigma != 0.f) {  // the covariance update reads the samples and their costs back (mppi.py:119-125)
            A(h->costs.alloc(E * (size_t)h->n_local));
            A(h->samples.alloc(E * (size_t)h->n_local * h->n));
        }
    } else {
        A(h->R.alloc(E * nn));
        A(h->Vh.alloc(E * nn));
        A(h->tau.alloc(E * n));
        A(h->Qt.alloc(E * nn));
        A(h->F.alloc(E * nn));
        A(h->cov.alloc(E * nn));
        A(h->Lfull.alloc(E * nn));
        A(h->Lt.alloc(E * h->lt_floats));
        A(h->diag.alloc(E * 4 * n));
        A(h->hess_ws.alloc(E * hessian_workspace_floats(h->H)));
        A(h->zolo.alloc(kZoloTableDoubles));
        if (e == cudaSuccess) {
            std::vector<double> tab(kZoloTableDoubles);
            zolotarev_table(tab.data());
            zolotarev_table_dense(tab.data() + (size_t)kZoloLadder * 2 * kZoloPoles);
            A(h2d(h, h->zolo.p, tab.data(), tab.size() * sizeof(double)));
        }
    }
    A(cudaHostAlloc(&h->h_state, E * kStateFloats * sizeof(float), cudaHostAllocMapped));
    A(cudaHostAlloc(&h->h_time, E * sizeof(int), cudaHostAllocMapped));
    A(cudaHostAlloc(&h->h_action, E * 4 * sizeof(float), cudaHostAllocMapped));
    A(cudaHostAlloc(&h->h_status, E * sizeof(int), cudaHostAllocMapped));
    A(cudaHostAlloc(&h->h_flag, sizeof(unsigned int), cudaHostAllocMapped));
    if (e == cudaSuccess) {
        *h->h_flag = 0u;
        memset(h->h_status, 0, E * sizeof(int));
        A(cudaHostGetDevicePointer((void**)&h->dh_state, h->h_state, 0));
        A(cudaHostGetDevicePointer((void**)&h->dh_time, h->h_time, 0));
        A(cudaHostGetDevicePointer((void**)&h->dh_action, h->h_action, 0));
        A(cudaHostGetDevicePointer((void**)&h->dh_status, h->h_status, 0));
        A(cudaHostGetDevicePointer((void**)&h->dh_flag, h->h_flag, 0));
        const char* zc = getenv("COVO_ZEROCOPY");
        h->zero_copy = !(zc && zc[0] == '0');
    }
    if (e != cudaSuccess) {
        release_all(h);
        delete h;
        return fail(COVO_ERR_CUDA, "workspace allocation failed: %s", cudaGetErrorString(e));
    }
    // defaults of get_controller (envs/quadrotor.py:685-690, 703-746)
    {
        std::vector<float> mean(E * n);
        float th = (cfg->m * cfg->g / cfg->max_thrust) * 2.0f - 1.0f;
        for (size_t i = 0; i < E * (size_t)h->H; ++i) {
            mean[i * 4 + 0] = th;
            mean[i * 4 + 1] = mean[i * 4 + 2] = mean[i * 4 + 3] = 0.f;
        }
        h2d(h, h->a_mean.p, mean.data(), mean.size() * sizeof(float));
        if (cfg->mode == COVO_MODE_MPPI) {
            std::vector<float> L(E * h->H * 16, 0.f), C(E * h->H * 16, 0.f);
            for (size_t b = 0; b < E * (size_t)h->H; ++b)
                for (int k = 0; k < 4; ++k) {
                    L[b * 16 + k * 5] = cfg->sample_sigma;
                    C[b * 16 + k * 5] = cfg->sample_sigma * cfg->sample_sigma;
                }
            h2d(h, h->Lblk.p, L.data(), L.size() * sizeof(float));
            h2d(h, h->cov.p, C.data(), C.size() * sizeof(float));
            h->have_factor = true;
        }
        ce = cudaStreamSynchronize(h->own_stream);  // initial uploads complete before any caller-stream launch
        if (ce != cudaSuccess) {
            release_all(h);
            delete h;
            return fail(COVO_ERR_CUDA, "initial upload failed: %s", cudaGetErrorString(ce));
        }
    }
    *out = h;
    return COVO_OK;
}

int covo_destroy(covo_handle* h) {
    if (!h) return COVO_OK;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    release_all(h);
    delete h;
    return COVO_OK;
}

int covo_set_reference(covo_handle* h, const float* pos, const float* vel, const float* acc) {
    if (!h || !pos || !vel) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    size_t bytes = (size_t)h->E * h->T * 3 * sizeof(float);
    CK(h2d(h, h->pos_traj.p, pos, bytes));
    CK(h2d(h, h->vel_traj.p, vel, bytes));
    if (acc) CK(h2d(h, h->acc_traj.p, acc, bytes));
    else CK(dzero(h, h->acc_traj.p, bytes));
    // the step entry points may run on a caller stream: uploads on the handle's own stream are complete on return
    CK(cudaStreamSynchronize(h->own_stream));
    return COVO_OK;
}

int covo_set_mean(covo_handle* h, const float* a_mean) {
    if (!h || !a_mean) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    CK(h2d(h, h->a_mean.p, a_mean, (size_t)h->E * h->n * sizeof(float)));
    CK(cudaStreamSynchronize(h->own_stream));
    return COVO_OK;
}
int covo_get_mean(covo_handle* h, float* a_mean) {
    if (!h || !a_mean) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    CK(d2h(h, a_mean, h->a_mean.p, (size_t)h->E * h->n * sizeof(float)));
    return COVO_OK;
}

int covo_set_cov(covo_handle* h, const float* a_cov) {
    if (!h || !a_cov) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    if (h->cfg.mode == COVO_MODE_MPPI) {
        size_t cnt = (size_t)h->E * h->H * 16;
        std::vector<float> L(cnt, 0.f);
        for (size_t b = 0; b < (size_t)h->E * h->H; ++b) {  // 4x4 lower Cholesky on the host (setup path)
            const float* C = a_cov + b * 16;
            float* Lb = L.data() + b * 16;
            for (int j = 0; j < 4; ++j) {
                double s = C[j * 4 + j];
                for (int k = 0; k < j; ++k) s -= (double)Lb[j * 4 + k] * Lb[j * 4 + k];
                if (!(s > 0.0)) return fail(COVO_ERR_NUMERIC, "a_cov block %zu is not positive definite", b);
                double d = sqrt(s);
                Lb[j * 4 + j] = (float)d;
                for (int i = j + 1; i < 4; ++i) {
                    double t = C[i * 4 + j];
                    for (int k = 0; k < j; ++k) t -= (double)Lb[i * 4 + k] * Lb[j * 4 + k];
                    Lb[i * 4 + j] = (float)(t / d);
                }
            }
        }
        CK(h2d(h, h->Lblk.p, L.data(), cnt * sizeof(float)));
        CK(h2d(h, h->cov.p, a_cov, cnt * sizeof(float)));
        CK(cudaStreamSynchronize(h->own_stream));
    } else {
        CK(h2d(h, h->cov.p, a_cov, (size_t)h->E * h->n * h->n * sizeof(float)));
        CK(dzero(h, h->status.p, h->E * sizeof(int)));
        SigmaArgs sa = sigma_args(h);
        CK(launch_cholesky(sa, h->E, h->own_stream));
        CK(cudaStreamSynchronize(h->own_stream));
    }
    h->have_factor = true;
    return COVO_OK;
}
int covo_get_cov(covo_handle* h, float* a_cov) {
    if (!h || !a_cov) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    size_t cnt = (h->cfg.mode == COVO_MODE_MPPI) ? (size_t)h->E * h->H * 16 : (size_t)h->E * h->n * h->n;
    CK(d2h(h, a_cov, h->cov.p, cnt * sizeof(float)));
    return COVO_OK;
}

static int alloc_schedule(covo_handle* h, int t_sched, bool full) {
    if (h->cfg.mode != COVO_MODE_COVO_OFFLINE) return fail(COVO_ERR_INVALID, "handle is not in covo-offline mode");
    if (t_sched < 1) return fail(COVO_ERR_INVALID, "t_sched must be positive");
    const size_t nn = (size_t)h->n * h->n, S = t_sched;
    if ((int)(h->cov_table.n / nn) < t_sched) {
        h->cov_table.release();
        h->Lt_table.release();
        h->sched_status.release();
        CK(h->cov_table.alloc(S * nn));
        CK(h->Lt_table.alloc(S * h->lt_floats));
        CK(h->sched_status.alloc(S));
    }
    if (full && (int)(h->sched_R.n / nn) < t_sched) {
        h->sched_states.release(); h->sched_times.release(); h->sched_anom.release(); h->sched_R.release();
        h->sched_Vh.release(); h->sched_tau.release(); h->sched_Qt.release(); h->sched_F.release();
        h->sched_ws.release(); h->sched_diag.release();
        CK(h->sched_states.alloc(S * kStateFloats));
        CK(h->sched_times.alloc(S));
        CK(h->sched_anom.alloc(S * h->n));
        CK(h->sched_R.alloc(S * nn));
        CK(h->sched_Vh.alloc(S * nn));
        CK(h->sched_tau.alloc(S * h->n));
        CK(h->sched_Qt.alloc(S * nn));
        CK(h->sched_F.alloc(S * nn));
        CK(h->sched_ws.alloc(S * hessian_workspace_floats(h->H)));
        CK(h->sched_diag.alloc(S * 4 * h->n));
    }
    return COVO_OK;
}

static SigmaArgs sched_sigma_args(covo_handle* h) {
    SigmaArgs a = sigma_args(h);
    a.R = h->sched_R.p;
    a.Vh = h->sched_Vh.p;
    a.tau = h->sched_tau.p;
    a.Qt = h->sched_Qt.p;
    a.F = h->sched_F.p;
    a.cov = h->cov_table.p;
    a.L = nullptr;
    a.Lt = h->Lt_table.p;
    a.diag = h->sched_diag.p;
    a.status = h->sched_status.p;
    return a;
}

int covo_set_cov_offline(covo_handle* h, const float* table, int t_sched) {
    if (!h || !table) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    int rc = alloc_schedule(h, t_sched, false);
    if (rc) return rc;
    CK(h2d(h, h->cov_table.p, table, (size_t)t_sched * h->n * h->n * sizeof(float)));
    SigmaArgs sa = sigma_args(h);
    sa.cov = h->cov_table.p;
    sa.L = nullptr;
    sa.Lt = h->Lt_table.p;
    sa.status = h->sched_status.p;
    CK(dzero(h, h->sched_status.p, t_sched * sizeof(int)));
    CK(launch_cholesky(sa, t_sched, h->own_stream));
    CK(cudaStreamSynchronize(h->own_stream));
    h->t_sched = t_sched;
    h->have_factor = true;
    graphs_invalidate(h);
    return COVO_OK;
}

int covo_get_cov_offline(covo_handle* h, float* table, int t_sched) {
    if (!h || !table) return fail(COVO_ERR_INVALID, "null argument");
    if (t_sched > h->t_sched) return fail(COVO_ERR_INVALID, "schedule has %d steps", h->t_sched);
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    CK(d2h(h, table, h->cov_table.p, (size_t)t_sched * h->n * h->n * sizeof(float)));
    return COVO_OK;
}

int covo_pid_action(covo_handle* h, const float* state24, const int* time, float Kp, float Kd, float Ki, float Kp_att,
                    const float* integral, float* action) {
    if (!h || !state24 || !time || !action) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->own_stream;
    CK(cudaMemcpyAsync(h->state24.p, state24, (size_t)h->E * kStateFloats * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->time.p, time, (size_t)h->E * sizeof(int), cudaMemcpyHostToDevice, st));
    if (integral) {
        if (h->pid_integral.n < (size_t)h->E * 3) {
            h->pid_integral.release();
            CK(h->pid_integral.alloc((size_t)h->E * 3));
        }
        CK(cudaMemcpyAsync(h->pid_integral.p, integral, (size_t)h->E * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    PidArgs pa;
    pa.n_env = h->E;
    pa.traj_len = h->T;
    pa.env = h->env;
    pa.max_thrust = h->cfg.max_thrust;
    pa.Kp = Kp;
    pa.Kd = Kd;
    pa.Ki = Ki;
    pa.Kp_att = Kp_att;
    pa.state24 = h->state24.p;
    pa.time = h->time.p;
    pa.acc_traj = h->acc_traj.p;
    pa.integral = integral ? h->pid_integral.p : nullptr;
    pa.action = h->action.p;
    cudaError_t e = launch_pid(pa, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(action, h->action.p, (size_t)h->E * 4 * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(COVO_ERR_CUDA, "pid_action: %s", cudaGetErrorString(e));
    return COVO_OK;
}

int covo_reset_offline(covo_handle* h, const float* state24, const int* time, int t_sched) {
    return covo_reset_offline_disturbed(h, state24, time, t_sched, nullptr);
}

int covo_reset_offline_disturbed(covo_handle* h, const float* state24, const int* time, int t_sched, const float* f_disturb) {
    if (!h || !state24 || !time) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    int rc = alloc_schedule(h, t_sched, true);
    if (rc) return rc;
    cudaStream_t st = h->own_stream;
    if (f_disturb) {
        if (h->sched_disturb.n < (size_t)t_sched * 3) {
            h->sched_disturb.release();
            CK(h->sched_disturb.alloc((size_t)t_sched * 3));
        }
        CK(cudaMemcpyAsync(h->sched_disturb.p, f_disturb, (size_t)t_sched * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(h->state24.p, state24, kStateFloats * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->time.p, time, sizeof(int), cudaMemcpyHostToDevice, st));
    OfflineArgs oa;
    oa.H = h->H;
    oa.traj_len = h->T;
    oa.t_sched = t_sched;
    oa.env = h->env;
    oa.max_thrust = h->cfg.max_thrust;
    oa.Kp = 10.f;  // controllers/covo.py:48-53
    oa.Kd = 5.f;
    oa.Kp_att = 10.f;
    oa.state24 = h->state24.p;
    oa.time = h->time.p;
    oa.pos_traj = h->pos_traj.p;
    oa.vel_traj = h->vel_traj.p;
    oa.acc_traj = h->acc_traj.p;
    oa.f_disturb = f_disturb ? h->sched_disturb.p : nullptr;
    oa.states24 = h->sched_states.p;
    oa.times = h->sched_times.p;
    oa.a_nom = h->sched_anom.p;
    CK(launch_offline_paths(oa, st));
    // all schedule steps as one batch: "environment" t = schedule step t, shared reference trajectory
    HessianArgs ha = hess_args(h, h->sched_states.p, h->sched_times.p, h->sched_anom.p, 0, h->sched_R.p, h->sched_ws.p, 0);
    CK(launch_hessian(ha, t_sched, st));
    CK(cudaMemsetAsync(h->sched_status.p, 0, t_sched * sizeof(int), st));
    SigmaArgs sa = sched_sigma_args(h);
    CK(launch_sigma(sa, t_sched, st));
    sa.cov_symmetric = 1;
    CK(launch_cholesky(sa, t_sched, st));
    CK(cudaStreamSynchronize(st));
    h->t_sched = t_sched;
    h->have_factor = true;
    graphs_invalidate(h);
    return COVO_OK;
}

int covo_step_device(covo_handle* h, const float* st_d, const int* tm_d, const float* eps_d, float* act_d, void* stream) {
    if (!h || !st_d || !tm_d || !act_d) return fail(COVO_ERR_INVALID, "null argument");
    if (h->cfg.world != 1) return fail(COVO_ERR_INVALID, "world > 1: use covo_step_partial_device + covo_step_merge_device");
    CK(cudaSetDevice(h->cfg.device));
    return step_common(h, st_d, tm_d, eps_d, act_d, (cudaStream_t)stream, 1);
}

// ---------------------------------------------------------------------------------------------------------
// Device-resident environment and closed loop
// ---------------------------------------------------------------------------------------------------------
static int env_alloc(covo_handle* h) {
    const size_t E = (size_t)h->E;
    if (h->env_state24.n < E * kStateFloats) {
        CK(h->env_state24.alloc(E * kStateFloats));
        CK(h->env_noisy24.alloc(E * kStateFloats));
        CK(h->env_time.alloc(E));
        CK(h->env_noisy_time.alloc(E));
        CK(h->env_done.alloc(E));
        CK(h->env_action.alloc(E * 4));
    }
    return COVO_OK;
}

static EnvStepArgs env_args(covo_handle* h, int gaussian, float obs_scale, float dyn_scale, unsigned long long seed) {
    EnvStepArgs a;
    memset(&a, 0, sizeof(a));  // (compared bytewise by the closed-loop graph cache)
    a.env = h->env;
    a.n_env = h->E;
    a.traj_len = h->T;
    a.traj_stride = (long long)h->T * 3;
    a.obs_noise_scale = obs_scale;
    a.dyn_noise_scale = dyn_scale;
    a.gaussian = gaussian;
    a.do_step = 1;
    a.seed = seed;
    a.stream = 0;
    a.state24 = h->env_state24.p;
    a.time = h->env_time.p;
    a.pos_traj = h->pos_traj.p;
    a.vel_traj = h->vel_traj.p;
    a.action = nullptr;
    a.noise_in = nullptr;
    a.noisy24 = h->env_noisy24.p;
    a.noisy_time = h->env_noisy_time.p;
    a.reward = nullptr;
    a.err_pos = nullptr;
    a.done = nullptr;
    if (gaussian && h->cfg.mode == COVO_MODE_MPPI && h->fdist.p && h->fdist_on) {
        a.mppi_fdist = h->fdist.p;
        a.mppi_H = h->H;
    }
    if (h->pool_n > 0) {
        a.reset_pool = h->pool_n;
        a.reset_state24 = h->pool_state24.p;
        a.reset_time = h->pool_time.p;
        a.reset_pos_traj = h->pool_pos.p;
        a.reset_vel_traj = h->pool_vel.p;
        a.reset_count = h->pool_count.p;
        a.traj_pos_rw = h->pos_traj.p;
        a.traj_vel_rw = h->vel_traj.p;
        if (h->pool_reset_mean) {
            a.a_mean = h->a_mean.p;
            a.a_mean_init = h->mean_init.p;
            a.n_mean = h->n;
        }
    }
    return a;
}

int covo_env_set_reset_pool(covo_handle* h, int n_pool, const float* state24, const int* time, const float* pos_traj, const float* vel_traj,
                            const float* a_mean_init) {
    if (!h) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    graphs_invalidate(h);  // the closed-loop graph carries the pool pointers
    h->loop_args_valid = false;
    if (n_pool <= 0) {
        h->pool_n = 0;
        return COVO_OK;
    }
    if (!state24 || !time || !pos_traj || !vel_traj) return fail(COVO_ERR_INVALID, "null argument");
    if (h->cfg.mode == COVO_MODE_COVO_OFFLINE)
        return fail(COVO_ERR_NOT_IMPLEMENTED, "auto-reset on the device for covo-offline (the covariance schedule is rebuilt per episode: covo_reset_offline)");
    const size_t E = (size_t)h->E, P = (size_t)n_pool, tl = (size_t)h->T * 3;
    h->pool_state24.release(); h->pool_time.release(); h->pool_pos.release(); h->pool_vel.release(); h->pool_count.release();
    CK(h->pool_state24.alloc(P * E * kStateFloats));
    CK(h->pool_time.alloc(P * E));
    CK(h->pool_pos.alloc(P * E * tl));
    CK(h->pool_vel.alloc(P * E * tl));
    CK(h->pool_count.alloc(E));
    CK(h2d(h, h->pool_state24.p, state24, P * E * kStateFloats * sizeof(float)));
    CK(h2d(h, h->pool_time.p, time, P * E * sizeof(int)));
    CK(h2d(h, h->pool_pos.p, pos_traj, P * E * tl * sizeof(float)));
    CK(h2d(h, h->pool_vel.p, vel_traj, P * E * tl * sizeof(float)));
    h->pool_reset_mean = a_mean_init != nullptr;
    if (a_mean_init) {
        if (h->mean_init.n < (size_t)h->n) {
            h->mean_init.release();
            CK(h->mean_init.alloc(h->n));
        }
        CK(h2d(h, h->mean_init.p, a_mean_init, (size_t)h->n * sizeof(float)));
    }
    CK(cudaStreamSynchronize(h->own_stream));
    h->pool_n = n_pool;
    return COVO_OK;
}

int covo_env_reset(covo_handle* h, const float* state24, const int* time) {
    if (!h || !state24 || !time) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    if (int rc = env_alloc(h)) return rc;
    CK(h2d(h, h->env_state24.p, state24, (size_t)h->E * kStateFloats * sizeof(float)));
    CK(h2d(h, h->env_time.p, time, (size_t)h->E * sizeof(int)));
    CK(cudaStreamSynchronize(h->own_stream));
    h->env_ready = true;
    return COVO_OK;
}

int covo_env_get_state(covo_handle* h, float* state24, int* time) {
    if (!h || !state24 || !time) return fail(COVO_ERR_INVALID, "null argument");
    if (!h->env_ready) return fail(COVO_ERR_INVALID, "covo_env_reset has not been called");
    CK(cudaSetDevice(h->cfg.device));
    CK(d2h(h, state24, h->env_state24.p, (size_t)h->E * kStateFloats * sizeof(float)));
    CK(d2h(h, time, h->env_time.p, (size_t)h->E * sizeof(int)));
    return COVO_OK;
}

int covo_env_step(covo_handle* h, const float* action, const float* noise, unsigned long long noise_seed, unsigned int noise_step,
                  int gaussian, float obs_noise_scale, float dyn_noise_scale, float* noisy24, float* reward, float* err_pos,
                  int* done) {
    if (!h) return fail(COVO_ERR_INVALID, "null argument");
    if (!h->env_ready) return fail(COVO_ERR_INVALID, "covo_env_reset has not been called");
    CK(cudaSetDevice(h->cfg.device));
    const size_t E = (size_t)h->E;
    if (h->env_log_f.n < 2 * E) {
        h->env_log_f.release();
        CK(h->env_log_f.alloc(2 * E));
    }
    EnvStepArgs a = env_args(h, gaussian, obs_noise_scale, dyn_noise_scale, noise_seed);
    a.stream = noise_step;
    a.do_step = action ? 1 : 0;  // NULL action: only the noisy copy of the current state (after a reset)
    if (action) {
        CK(h2d(h, h->env_action.p, action, E * 4 * sizeof(float)));
        a.action = h->env_action.p;
    }
    if (noise) {
        if (h->env_noise.n < E * kEnvNoiseFloats) {
            h->env_noise.release();
            CK(h->env_noise.alloc(E * kEnvNoiseFloats));
        }
        CK(h2d(h, h->env_noise.p, noise, E * kEnvNoiseFloats * sizeof(float)));
        a.noise_in = h->env_noise.p;
    }
    a.reward = h->env_log_f.p;
    a.err_pos = h->env_log_f.p + E;
    a.done = h->env_done.p;
    CK(launch_env_step(a, h->own_stream));
    if (noisy24) CK(d2h(h, noisy24, h->env_noisy24.p, E * kStateFloats * sizeof(float)));
    if (action) {
        if (reward) CK(d2h(h, reward, h->env_log_f.p, E * sizeof(float)));
        if (err_pos) CK(d2h(h, err_pos, h->env_log_f.p + E, E * sizeof(float)));
        if (done) CK(d2h(h, done, h->env_done.p, E * sizeof(int)));
    }
    CK(cudaStreamSynchronize(h->own_stream));
    return COVO_OK;
}

int covo_closed_loop(covo_handle* h, int n_steps, unsigned long long noise_seed, int gaussian, float obs_noise_scale,
                     float dyn_noise_scale, const float* noise, float* actions, float* rewards, float* err_pos) {
    if (!h || n_steps <= 0) return fail(COVO_ERR_INVALID, "bad argument");
    if (!h->env_ready) return fail(COVO_ERR_INVALID, "covo_env_reset has not been called");
    if (h->cfg.world != 1) return fail(COVO_ERR_INVALID, "closed loop: world must be 1 (shard environments, not samples)");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->own_stream;
    const size_t E = (size_t)h->E, S = (size_t)n_steps;
    // logs: actions [S][E][4] | rewards [S][E] | err_pos [S][E]
    if (h->env_log_f.n < S * E * 6) {
        h->env_log_f.release();
        CK(h->env_log_f.alloc(S * E * 6));
    }
    float* act_d = h->env_log_f.p;
    float* rew_d = act_d + S * E * 4;
    float* err_d = rew_d + S * E;
    if (noise) {
        const size_t cnt = (S + 1) * E * kEnvNoiseFloats;
        if (h->env_noise.n < cnt) {
            h->env_noise.release();
            CK(h->env_noise.alloc(cnt));
        }
        CK(h2d(h, h->env_noise.p, noise, cnt * sizeof(float)));
    }
    if (h->cfg.mode == COVO_MODE_MPPI) {
        // mppi.py:74 under gaussian: the rollouts of every call see one random force; the environment kernel draws it for the next call
        const bool want = gaussian != 0;
        if (want && h->fdist.n < E * h->H * 3) {
            h->fdist.release();
            CK(h->fdist.alloc(E * h->H * 3));
            graphs_invalidate(h);
        }
        if (want != h->fdist_on) graphs_invalidate(h);
        h->fdist_on = want;
    }
    EnvStepArgs a = env_args(h, gaussian, obs_noise_scale, dyn_noise_scale, noise_seed);
    // the noisy copy of the initial state (quadrotor.py:366-370: reset_env ends with get_info)
    a.do_step = 0;
    a.stream = 0;
    a.noise_in = noise ? h->env_noise.p : nullptr;
    CK(launch_env_step(a, st));
    int i0 = 0;
    if (graph_ok(h, nullptr) && h->direct_steps == 0 && n_steps > 0) {
        // first step of a fresh handle: launched directly (configures the kernels), same arithmetic as the replayed ones
        float* ai = act_d;
        int rc = step_common(h, h->env_noisy24.p, h->env_noisy_time.p, nullptr, ai, st, 1);
        if (rc) return rc;
        a.do_step = 1;
        a.stream = 1u;
        a.action = ai;
        a.noise_in = noise ? h->env_noise.p + E * kEnvNoiseFloats : nullptr;
        a.reward = rew_d;
        a.err_pos = err_d;
        a.done = h->env_done.p;
        CK(launch_env_step(a, st));
        i0 = 1;
    }
    if (graph_ok(h, nullptr) && i0 < n_steps) {
        // One graph launch per closed-loop step: [controller kernels -> Quad3D.step_env -> counters].  The step index lives on the
        // device (dev_ctr[1]): it selects the noise stream / block and the log rows; the action travels through h->action.
        EnvStepArgs la = env_args(h, gaussian, obs_noise_scale, dyn_noise_scale, noise_seed);
        la.do_step = 1;
        la.stream = 1u;
        la.action = h->action.p;
        la.action_log = act_d;
        la.noise_in = noise ? h->env_noise.p + E * kEnvNoiseFloats : nullptr;
        la.reward = rew_d;
        la.err_pos = err_d;
        la.done = h->env_done.p;
        la.step_ctr = h->dev_ctr.p + 1;
        if (!h->g_loop.valid || memcmp(&la, &h->loop_env_args, sizeof(la)) != 0) {
            h->loop_env_args = la;
            h->g_loop.valid = false;
            int rc = graph_build(h, &h->g_loop, h->env_noisy24.p, h->env_noisy_time.p, h->action.p, 1, 2);
            if (rc == COVO_ERR_CUDA) {
                h->graphs_enabled = false;
                cudaGetLastError();
            } else if (rc) {
                return rc;
            }
        }
        if (h->g_loop.valid) {
            const unsigned int first = (unsigned int)i0;
            CK(cudaMemcpyAsync(h->dev_ctr.p + 1, &first, sizeof(unsigned int), cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
            int rc = sync_rng_counter(h, st);
            if (rc) return rc;
            for (int i = i0; i < n_steps; ++i) {
                CK(cudaGraphLaunch(h->g_loop.exec, st));
                h->rng_stream += 1;
                h->rng_ctr_shadow += 1;
            }
            i0 = n_steps;
        }
    }
    for (int i = i0; i < n_steps; ++i) {
        float* ai = act_d + (size_t)i * E * 4;
        int rc = step_common(h, h->env_noisy24.p, h->env_noisy_time.p, nullptr, ai, st, 1);
        if (rc) return rc;
        a.do_step = 1;
        a.stream = (unsigned)(i + 1);
        a.action = ai;
        a.noise_in = noise ? h->env_noise.p + (size_t)(i + 1) * E * kEnvNoiseFloats : nullptr;
        a.reward = rew_d + (size_t)i * E;
        a.err_pos = err_d + (size_t)i * E;
        a.done = h->env_done.p;
        CK(launch_env_step(a, st));
    }
    if (actions) CK(d2h(h, actions, act_d, S * E * 4 * sizeof(float)));
    if (rewards) CK(d2h(h, rewards, rew_d, S * E * sizeof(float)));
    if (err_pos) CK(d2h(h, err_pos, err_d, S * E * sizeof(float)));
    CK(cudaStreamSynchronize(st));
    return COVO_OK;
}

int covo_step(covo_handle* h, const float* state24, const int* time, const float* eps, float* action) {
    if (!h || !state24 || !time || !action) return fail(COVO_ERR_INVALID, "null argument");
    if (h->cfg.world != 1) return fail(COVO_ERR_INVALID, "world > 1: use the *_device partial/merge entry points");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->own_stream;
    memcpy(h->h_state, state24, (size_t)h->E * kStateFloats * sizeof(float));
    memcpy(h->h_time, time, (size_t)h->E * sizeof(int));
    const float* eps_d = nullptr;
    if (eps) {
        size_t cnt = (size_t)h->E * h->n_local * h->n;
        if (h->eps.n < cnt) {
            h->eps.release();
            CK(h->eps.alloc(cnt));
        }
        CK(cudaMemcpyAsync(h->eps.p, eps, cnt * sizeof(float), cudaMemcpyHostToDevice, st));
        eps_d = h->eps.p;
    }
    // pinned staging -> device, the kernels, device -> pinned staging: one CUDA graph (or the same sequence launched directly)
    h->flag_pending = false;
    int rc = step_common(h, h->state24.p, h->time.p, eps_d, h->action.p, st, 1);
    if (rc) return rc;
    if (h->flag_pending) {
        // zero-copy graph: spin on the flag the last node raises (mapped pinned memory); look at the stream now and then so that a
        // failed kernel surfaces as an error instead of an endless wait
        volatile unsigned int* flag = h->h_flag;
        unsigned int spins = 0;
        while (*flag != h->host_epoch) {
            if ((++spins & 0x3fffu) == 0u) {
                cudaError_t q = cudaStreamQuery(st);
                if (q != cudaErrorNotReady) {
                    CK(q);
                    if (*flag == h->host_epoch) break;
                    return fail(COVO_ERR_CUDA, "covo_step: the step completed without raising its completion flag");
                }
            }
        }
        __sync_synchronize();
        h->flag_pending = false;
    } else {
        CK(cudaStreamSynchronize(st));
    }
    memcpy(action, h->h_action, (size_t)h->E * 4 * sizeof(float));
    if (step_has_status(h))  // the covariance step of THIS call (its first kernel clears the status)
        for (int e = 0; e < h->E; ++e)
            if (h->h_status[e] == 3) {
                // the dense path's Lanczos stage was not converged to 2e-7: this step's covariance is a valid SPD matrix whose
                // largest eigenvalue is off by up to a few per cent; from now on the handle uses the tridiagonal path
                h->sigma_dense = 0;
                graphs_invalidate(h);
            } else if (h->h_status[e])
                return fail(COVO_ERR_NUMERIC, "covo_step: numeric status %d in environment %d (%s)", h->h_status[e], e,
                            h->h_status[e] == 1 ? "spectral range of the Hessian exceeds the rational-approximation ladder"
                                                : "covariance not positive definite in float32");
    return COVO_OK;
}

int covo_set_rollout_disturbance(covo_handle* h, const float* fdist_seq) {
    if (!h) return fail(COVO_ERR_INVALID, "null argument");
    if (h->cfg.mode != COVO_MODE_MPPI) return fail(COVO_ERR_INVALID, "only MPPI rolls out with a stochastic step_env (mppi.py:74); CoVO's rollouts are deterministic (covo.py:229)");
    CK(cudaSetDevice(h->cfg.device));
    const bool on = fdist_seq != nullptr;
    if (on) {
        const size_t cnt = (size_t)h->E * h->H * 3;
        if (h->fdist.n < cnt) {
            h->fdist.release();
            CK(h->fdist.alloc(cnt));
            graphs_invalidate(h);
        }
        CK(cudaMemcpyAsync(h->fdist.p, fdist_seq, cnt * sizeof(float), cudaMemcpyHostToDevice, h->own_stream));
        CK(cudaStreamSynchronize(h->own_stream));
    }
    if (on != h->fdist_on) graphs_invalidate(h);  // the rollout node carries the pointer (or its absence)
    h->fdist_on = on;
    return COVO_OK;
}

int covo_set_env_params(covo_handle* h, float m, float g, float max_thrust, float dt, float alpha_bodyrate, float action_scale,
                        const float* max_omega3, int max_steps_in_episode) {
    if (!h || !max_omega3) return fail(COVO_ERR_INVALID, "null argument");
    if (!(m > 0.f) || !(dt > 0.f) || !(max_thrust > 0.f)) return fail(COVO_ERR_INVALID, "m, dt and max_thrust must be positive");
    // the model constants travel by value in every kernel's argument block: the next launch uses the new ones
    h->cfg.m = h->env.m = m;
    h->cfg.g = h->env.g = g;
    h->cfg.max_thrust = h->env.max_thrust = max_thrust;
    h->cfg.dt = h->env.dt = dt;
    h->cfg.alpha_bodyrate = h->env.alpha_bodyrate = alpha_bodyrate;
    h->cfg.action_scale = h->env.action_scale = action_scale;
    for (int k = 0; k < 3; ++k) h->cfg.max_omega[k] = h->env.max_omega[k] = max_omega3[k];
    h->cfg.max_steps_in_episode = h->env.max_steps = max_steps_in_episode;
    graphs_invalidate(h);
    return COVO_OK;
}

int covo_step_partial_device(covo_handle* h, const float* st_d, const int* tm_d, const float* eps_d, void* stream) {
    if (!h || !st_d || !tm_d) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    // a_mean is NOT updated here (finalize = 0): the merge kernel applies the shift + update
    Prof pf(h, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = h->cfg.mode;
    if (mode == COVO_MODE_MPPI) {
        shift_blocks_kernel<<<h->E, 128, 2 * h->H * 16 * sizeof(float), st>>>(h->Lblk.p, h->cov.p, h->H);
        CK(cudaGetLastError());
    } else if (mode == COVO_MODE_COVO_ONLINE) {
        HessianArgs ha = hess_args(h, st_d, tm_d, h->a_mean.p, 1, h->R.p, h->hess_ws.p, (long long)h->T * 3);
        CK(launch_hessian(ha, h->E, st));
        int rc = run_sigma_chol(h, st, nullptr);
        if (rc) return rc;
    } else if (h->t_sched <= 0) {
        return fail(COVO_ERR_INVALID, "covo-offline: no schedule");
    }
    RolloutArgs ra = rollout_args(h, st_d, tm_d, h->a_mean.p, 1, eps_d, nullptr, h->a_mean.p, h->action.p, nullptr, nullptr, 0);
    CK(launch_rollout(ra, h->E, st));
    if (!eps_d) h->rng_stream += 1;
    return COVO_OK;
}

int covo_partial_buffer(covo_handle* h, float** dev_ptr, int* n_floats) {
    if (!h || !dev_ptr || !n_floats) return fail(COVO_ERR_INVALID, "null argument");
    *dev_ptr = h->rank_partial.p;
    *n_floats = h->E * h->rec;
    return COVO_OK;
}

int covo_step_merge_device(covo_handle* h, const float* gathered_dev, float* action_dev, void* stream) {
    if (!h || !gathered_dev || !action_dev) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    MergeArgs m;
    m.world = h->cfg.world;
    m.n = h->n;
    m.n_pad = h->n_pad;
    m.n_env = h->E;
    m.lam = h->cfg.lam;
    m.gamma_mean = h->cfg.gamma_mean;
    m.shift = 1;
    m.gathered = gathered_dev;
    // the merge reads the un-shifted mean and writes the updated one; a scratch copy keeps it race-free
    if (h->gathered_scratch.n < (size_t)h->E * h->n) {
        h->gathered_scratch.release();
        CK(h->gathered_scratch.alloc((size_t)h->E * h->n));
    }
    CK(cudaMemcpyAsync(h->gathered_scratch.p, h->a_mean.p, (size_t)h->E * h->n * sizeof(float), cudaMemcpyDeviceToDevice,
                       (cudaStream_t)stream));
    m.a_mean_in = h->gathered_scratch.p;
    m.a_mean_out = h->a_mean.p;
    m.action_out = action_dev;
    CK(launch_merge(m, (cudaStream_t)stream));
    return COVO_OK;
}

// ---- fused exchange of the N-sharded step -----------------------------------------------------------------------------------
int covo_exchange_info(covo_handle* h, void* ipc_handle64, void** dev_ptr) {
    if (!h) return fail(COVO_ERR_INVALID, "null argument");
    if (!h->xchg.p) return fail(COVO_ERR_INVALID, "no exchange buffer: the handle was created with world == 1 (or world > %d)", kMaxPeers);
    CK(cudaSetDevice(h->cfg.device));
    if (ipc_handle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        cudaIpcMemHandle_t hd;
        CK(cudaIpcGetMemHandle(&hd, h->xchg.p));
        memcpy(ipc_handle64, &hd, 64);
    }
    if (dev_ptr) *dev_ptr = h->xchg.p;
    return COVO_OK;
}

int covo_exchange_attach(covo_handle* h, int peer_rank, const void* ipc_handle64, void* dev_ptr) {
    if (!h) return fail(COVO_ERR_INVALID, "null argument");
    if (!h->xchg.p) return fail(COVO_ERR_INVALID, "no exchange buffer: the handle was created with world == 1");
    if (peer_rank < 0 || peer_rank >= h->cfg.world || peer_rank == h->cfg.rank) return fail(COVO_ERR_INVALID, "peer_rank %d out of range", peer_rank);
    if ((ipc_handle64 != nullptr) == (dev_ptr != nullptr)) return fail(COVO_ERR_INVALID, "pass either an IPC handle or a device pointer");
    CK(cudaSetDevice(h->cfg.device));
    if (h->xpeer[peer_rank] && h->xpeer_ipc[peer_rank]) CK(cudaIpcCloseMemHandle(h->xpeer[peer_rank]));
    if (ipc_handle64) {
        cudaIpcMemHandle_t hd;
        memcpy(&hd, ipc_handle64, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        h->xpeer[peer_rank] = (float*)p;
        h->xpeer_ipc[peer_rank] = true;
    } else {
        h->xpeer[peer_rank] = (float*)dev_ptr;
        h->xpeer_ipc[peer_rank] = false;
    }
    return COVO_OK;
}

int covo_step_sharded_device(covo_handle* h, const float* st_d, const int* tm_d, const float* eps_d, float* act_d, void* stream) {
    if (!h || !st_d || !tm_d || !act_d) return fail(COVO_ERR_INVALID, "null argument");
    if (!h->xchg.p) return fail(COVO_ERR_INVALID, "no exchange buffer: the handle was created with world == 1");
    const int world = h->cfg.world;
    for (int w = 0; w < world; ++w)
        if (!h->xpeer[w]) return fail(COVO_ERR_INVALID, "rank %d is not attached (covo_exchange_attach)", w);
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = h->cfg.mode;
    if (mode == COVO_MODE_MPPI) {
        shift_blocks_kernel<<<h->E, 128, 2 * h->H * 16 * sizeof(float), st>>>(h->Lblk.p, h->cov.p, h->H);
        CK(cudaGetLastError());
    } else if (mode == COVO_MODE_COVO_ONLINE) {  // every rank repeats the covariance step (it depends on the merged mean only)
        HessianArgs ha = hess_args(h, st_d, tm_d, h->a_mean.p, 1, h->R.p, h->hess_ws.p, (long long)h->T * 3);
        CK(launch_hessian(ha, h->E, st));
        int rc = run_sigma_chol(h, st, nullptr);
        if (rc) return rc;
    } else if (h->t_sched <= 0) {
        return fail(COVO_ERR_INVALID, "covo-offline: no schedule");
    }
    const size_t flag_off = (size_t)2 * world * h->E * h->rec;  // floats in front of the flags
    RolloutArgs ra = rollout_args(h, st_d, tm_d, h->a_mean.p, 1, eps_d, nullptr, h->a_mean.p, h->action.p, nullptr, nullptr, 0);
    ra.px.world = world;
    ra.px.rank = h->cfg.rank;
    for (int w = 0; w < world; ++w) {
        ra.px.rec[w] = h->xpeer[w];
        ra.px.flag[w] = reinterpret_cast<unsigned int*>(h->xpeer[w] + flag_off);
    }
    ra.px.epoch = h->xchg_step + 1u;  // the exchange is keyed by the step number, which every rank advances alike
    CK(launch_rollout(ra, h->E, st));
    MergeArgs m;
    m.world = world;
    m.n = h->n;
    m.n_pad = h->n_pad;
    m.n_env = h->E;
    m.lam = h->cfg.lam;
    m.gamma_mean = h->cfg.gamma_mean;
    m.shift = 1;
    m.gathered = h->xchg.p;
    m.flags = reinterpret_cast<const unsigned int*>(h->xchg.p + flag_off);
    m.stream = h->xchg_step;
    m.status = h->status.p;
    m.a_mean_in = h->a_mean.p;  // in place: the kernel reads before it writes
    m.a_mean_out = h->a_mean.p;
    m.action_out = act_d;
    CK(launch_merge(m, st));
    h->xchg_step += 1;
    if (!eps_d) h->rng_stream += 1;
    return COVO_OK;
}

// ---- operators ------------------------------------------------------------------------------------
static int upload_state(covo_handle* h, const float* state24, const int* time) {
    CK(h2d(h, h->state24.p, state24, (size_t)h->E * kStateFloats * sizeof(float)));
    CK(h2d(h, h->time.p, time, (size_t)h->E * sizeof(int)));
    return COVO_OK;
}

int covo_hessian(covo_handle* h, const float* state24, const int* time, const float* a_mean, int shift, float* R) {
    if (!h || !state24 || !time || !a_mean || !R) return fail(COVO_ERR_INVALID, "null argument");
    if (h->cfg.mode == COVO_MODE_MPPI) return fail(COVO_ERR_INVALID, "handle is in MPPI mode");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    int rc = upload_state(h, state24, time);
    if (rc) return rc;
    DevBuf<float> am;
    CK(am.alloc((size_t)h->E * h->n));
    CK(h2d(h, am.p, a_mean, (size_t)h->E * h->n * sizeof(float)));
    HessianArgs ha = hess_args(h, h->state24.p, h->time.p, am.p, shift, h->R.p, h->hess_ws.p, (long long)h->T * 3);
    cudaError_t e = launch_hessian(ha, h->E, h->own_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->own_stream);
    am.release();
    if (e != cudaSuccess) return fail(COVO_ERR_CUDA, "hessian: %s", cudaGetErrorString(e));
    CK(d2h(h, R, h->R.p, (size_t)h->E * h->n * h->n * sizeof(float)));
    return COVO_OK;
}

static int check_status(covo_handle* h, const char* what) {
    std::vector<int> s(h->E);
    CK(d2h(h, s.data(), h->status.p, h->E * sizeof(int)));
    for (int e = 0; e < h->E; ++e)
        if (s[e]) return fail(COVO_ERR_NUMERIC, "%s: numeric status %d in environment %d", what, s[e], e);
    return COVO_OK;
}

int covo_optimize_sigma(covo_handle* h, const float* R, float* a_cov) {
    if (!h || !R || !a_cov) return fail(COVO_ERR_INVALID, "null argument");
    if (h->cfg.mode == COVO_MODE_MPPI) return fail(COVO_ERR_INVALID, "handle is in MPPI mode");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    size_t bytes = (size_t)h->E * h->n * h->n * sizeof(float);
    CK(h2d(h, h->R.p, R, bytes));
    CK(dzero(h, h->status.p, h->E * sizeof(int)));
    int rc = run_sigma_chol(h, h->own_stream, nullptr);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->own_stream));
    if (h->sigma_dense) {
        // status 3: the Lanczos stage of the dense path did not converge (no well separated lowest eigenvalue: not a CoVO Hessian).
        // This entry point is synchronous, so it simply redoes the matrices on the tridiagonal path.
        std::vector<int> stv(h->E);
        CK(d2h(h, stv.data(), h->status.p, h->E * sizeof(int)));
        bool redo = false;
        for (int e = 0; e < h->E; ++e) redo = redo || stv[e] == 3;
        if (redo) {
            const int saved = h->sigma_dense;
            h->sigma_dense = 0;
            CK(dzero(h, h->status.p, h->E * sizeof(int)));
            rc = run_sigma_chol(h, h->own_stream, nullptr);
            h->sigma_dense = saved;
            if (rc) return rc;
            CK(cudaStreamSynchronize(h->own_stream));
        }
    }
    CK(d2h(h, a_cov, h->cov.p, bytes));
    return check_status(h, "optimize_sigma");
}

int covo_cholesky(covo_handle* h, const float* a_cov, float* L) {
    if (!h || !a_cov || !L) return fail(COVO_ERR_INVALID, "null argument");
    if (h->cfg.mode == COVO_MODE_MPPI) return fail(COVO_ERR_INVALID, "handle is in MPPI mode");
    int rc = covo_set_cov(h, a_cov);
    if (rc) return rc;
    CK(d2h(h, L, h->Lfull.p, (size_t)h->E * h->n * h->n * sizeof(float)));
    return check_status(h, "cholesky");
}

int covo_rollout(covo_handle* h, const float* state24, const int* time, const float* a_mean, int shift, const float* eps,
                 const float* fdist_seq, float* a_mean_out, float* action, float* costs, float* samples) {
    if (!h || !state24 || !time || !a_mean) return fail(COVO_ERR_INVALID, "null argument");
    if (!h->have_factor) return fail(COVO_ERR_INVALID, "no covariance factor: call covo_set_cov / covo_optimize_sigma first");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    int rc = upload_state(h, state24, time);
    if (rc) return rc;
    const size_t E = h->E, n = h->n, NL = h->n_local;
    DevBuf<float> am, amo;
    CK(am.alloc(E * n));
    CK(amo.alloc(E * n));
    CK(h2d(h, am.p, a_mean, E * n * sizeof(float)));
    const float* eps_d = nullptr;
    if (eps) {
        if (h->eps.n < E * NL * n) {
            h->eps.release();
            CK(h->eps.alloc(E * NL * n));
        }
        CK(h2d(h, h->eps.p, eps, E * NL * n * sizeof(float)));
        eps_d = h->eps.p;
    }
    const float* fd_d = nullptr;
    if (fdist_seq) {
        if (h->fdist.n < E * h->H * 3) {
            h->fdist.release();
            CK(h->fdist.alloc(E * h->H * 3));
        }
        CK(h2d(h, h->fdist.p, fdist_seq, E * h->H * 3 * sizeof(float)));
        fd_d = h->fdist.p;
    }
    if (costs && h->costs.n < E * NL) {
        h->costs.release();
        CK(h->costs.alloc(E * NL));
    }
    if (samples && h->samples.n < E * NL * n) {
        h->samples.release();
        CK(h->samples.alloc(E * NL * n));
    }
    if (h->pos_stats_on) CK(dzero(h, h->pos_stats.p, h->pos_stats.n * sizeof(float)));
    const int finalize = (h->cfg.world == 1) ? 1 : 0;
    RolloutArgs ra = rollout_args(h, h->state24.p, h->time.p, am.p, shift, eps_d, fd_d, amo.p, h->action.p,
                                  costs ? h->costs.p : nullptr, samples ? h->samples.p : nullptr, finalize);
    cudaError_t e = launch_rollout(ra, h->E, h->own_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->own_stream);
    if (!eps_d) h->rng_stream += 1;
    int ret = COVO_OK;
    if (e != cudaSuccess) ret = fail(COVO_ERR_CUDA, "rollout: %s", cudaGetErrorString(e));
    if (!ret && a_mean_out) d2h(h, a_mean_out, amo.p, E * n * sizeof(float));
    if (!ret && action) d2h(h, action, h->action.p, E * 4 * sizeof(float));
    if (!ret && costs) d2h(h, costs, h->costs.p, E * NL * sizeof(float));
    if (!ret && samples) d2h(h, samples, h->samples.p, E * NL * n * sizeof(float));
    am.release();
    amo.release();
    return ret;
}

// ---- debug / introspection ------------------------------------------------------------------------
int covo_enable_pos_stats(covo_handle* h, int on) {
    if (!h) return fail(COVO_ERR_INVALID, "null argument");
    h->pos_stats_on = on != 0;
    return COVO_OK;
}

int covo_get_pos_stats(covo_handle* h, float* pos_mean, float* pos_std) {
    if (!h || !pos_mean || !pos_std) return fail(COVO_ERR_INVALID, "null argument");
    if (!h->pos_stats_on) return fail(COVO_ERR_INVALID, "position statistics are disabled (covo_enable_pos_stats)");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    std::vector<float> s(h->pos_stats.n);
    CK(d2h(h, s.data(), h->pos_stats.p, s.size() * sizeof(float)));
    const double N = (double)h->n_local;
    for (int i = 0; i < h->E * h->H; ++i)
        for (int k = 0; k < 3; ++k) {
            double mean = s[i * 6 + k] / N, sq = s[i * 6 + 3 + k] / N;
            pos_mean[i * 3 + k] = (float)mean;
            pos_std[i * 3 + k] = (float)sqrt(fmax(sq - mean * mean, 0.0));  // jnp.std, ddof = 0
        }
    return COVO_OK;
}

int covo_set_jax_key(covo_handle* h, const unsigned int* act_key) {
    if (!h || !act_key) return fail(COVO_ERR_INVALID, "null argument");
    h->jax_key[0] = act_key[0];
    h->jax_key[1] = act_key[1];
    h->jax_key_pending = true;
    return COVO_OK;
}

int covo_debug_eps(covo_handle* h, unsigned int stream_id, float* eps) {
    if (!h || !eps) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    DevBuf<float> d;
    size_t cnt = (size_t)h->n_local * h->n;
    CK(d.alloc(cnt));
    int total = h->n_local * (h->n >> 2);
    debug_eps_kernel<<<(total + 255) / 256, 256, 0, h->own_stream>>>(d.p, h->n_local, h->sample_offset, h->n, h->cfg.seed, stream_id);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->own_stream);
    if (e == cudaSuccess) e = d2h(h, eps, d.p, cnt * sizeof(float));
    d.release();
    if (e != cudaSuccess) return fail(COVO_ERR_CUDA, "debug_eps: %s", cudaGetErrorString(e));
    return COVO_OK;
}

int covo_debug_tridiag(covo_handle* h, double* d, double* e, double* scalars5) {
    if (!h || !d || !e || !scalars5) return fail(COVO_ERR_INVALID, "null argument");
    if (h->cfg.mode == COVO_MODE_MPPI) return fail(COVO_ERR_INVALID, "handle is in MPPI mode");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    if (h->sigma_dense && h->dense_scal.p) {  // dense path: no tridiagonal matrix; the scalars of its Lanczos / log det stages instead
        memset(d, 0, h->n * sizeof(double));
        memset(e, 0, h->n * sizeof(double));
        scalars5[4] = 0.0;
        CK(d2h(h, scalars5, h->dense_scal.p, 4 * sizeof(double)));  // lambda_min, upper bound of the spectrum, log det A, Lanczos steps
        return COVO_OK;
    }
    CK(d2h(h, d, h->diag.p, h->n * sizeof(double)));
    CK(d2h(h, e, h->diag.p + h->n, h->n * sizeof(double)));
    CK(d2h(h, scalars5, h->diag.p + 2 * h->n, 5 * sizeof(double)));
    return COVO_OK;
}

int covo_zolotarev_nodes(double m, double M, int n_poles, double* shifts, double* weights) {
    if (!shifts || !weights || n_poles < 1 || !(m > 0.0) || !(M > m)) return fail(COVO_ERR_INVALID, "bad argument");
    zolotarev_nodes(m, M, n_poles, shifts, weights);
    return COVO_OK;
}

int covo_get_status(covo_handle* h, int* status) {
    if (!h || !status) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->own_stream));
    CK(d2h(h, status, h->status.p, h->E * sizeof(int)));
    return COVO_OK;
}

int covo_set_profiling(covo_handle* h, int on) {
    if (!h) return fail(COVO_ERR_INVALID, "null argument");
    h->profiling = on != 0;
    return COVO_OK;
}

int covo_get_kernel_ms(covo_handle* h, float* ms6) {
    if (!h || !ms6) return fail(COVO_ERR_INVALID, "null argument");
    if (!h->profiling) return fail(COVO_ERR_INVALID, "profiling is off");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaEventSynchronize(h->ev[6]));
    // boundaries: 0 start | 1 hessian (3 kernels) | 2 tridiagonalisation | 3 tridiagonal function | 4 sandwich |
    // 5 cholesky | 6 rollout
    for (int i = 0; i < 6; ++i) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]));
        ms6[i] = t;
    }
    return COVO_OK;
}

int covo_debug_phase_clocks(covo_handle* h, int on, long long* out64) {
    if (!h) return fail(COVO_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    h->phase_clocks = on != 0;
    if (out64) CK(d2h(h, out64, h->prof.p, 64 * sizeof(long long)));
    else CK(dzero(h, h->prof.p, 64 * sizeof(long long)));  // some slots are accumulators
    return COVO_OK;
}

int covo_set_sigma_path(covo_handle* h, int path) {
    if (!h) return fail(COVO_ERR_INVALID, "null argument");
    if (path != 0 && path != 3) return fail(COVO_ERR_INVALID, "path must be 0 (tridiagonal) or 3 (dense)");
    if (h->cfg.mode == COVO_MODE_MPPI) return fail(COVO_ERR_INVALID, "handle is in MPPI mode");
    CK(cudaSetDevice(h->cfg.device));
    if (path == 3 && !h->dense_X.p) {
        CK(h->dense_scal.alloc((size_t)h->E * 4));
        CK(h->dense_X.alloc((size_t)h->E * sigma_dense_scratch_floats(h->n)));
    }
    h->sigma_dense = path;
    graphs_invalidate(h);
    return COVO_OK;
}

int covo_get_sigma_path(covo_handle* h, int* path) {
    if (!h || !path) return fail(COVO_ERR_INVALID, "null argument");
    *path = h->sigma_dense;
    return COVO_OK;
}

int covo_rng_step(covo_handle* h, unsigned int* stream_id) {
    if (!h || !stream_id) return fail(COVO_ERR_INVALID, "null argument");
    *stream_id = h->rng_stream;
    return COVO_OK;
}

int covo_local_samples(covo_handle* h, int* n_local, int* offset) {
    if (!h || !n_local || !offset) return fail(COVO_ERR_INVALID, "null argument");
    *n_local = h->n_local;
    *offset = h->sample_offset;
    return COVO_OK;
}

}  // extern "C"
