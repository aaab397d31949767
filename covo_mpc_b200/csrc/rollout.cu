// K1 + K2: sample -> rollout -> softmax-reduce, one fused kernel.
//
// Replaces (reference file:line): controllers/covo.py:201-203, 212-278 and controllers/mppi.py:46-116
// (shift, multivariate_normal draw as mean + chol(cov) eps, clip, N x H step_env rollout with reward
// freeze on termination, discounted cost, softmax(-(cost - min)/lam) weighted mean), with
// envs/quadrotor.py:215-263, dynamics/free.py:74-155 and dynamics/utils.py:266-294 inlined
// (quad_model.cuh).
//
// One CTA owns a tile of TS = 64 trajectories of one environment:
//   phase 0  TMA bulk copy of the packed Cholesky factor (k-major, ~83 KB at n = 200) into shared
//            memory, overlapped with staging the mean, the reference-trajectory slice and the eps tile
//            (drawn in-kernel from the counter RNG, or read from HBM in parity mode);
//   phase 1  U = clip(mu + E L^T): a register-tiled fp32 triangular GEMM; every thread owns two
//            4-sample x 8-row tiles (row groups g and G-1-g, so the triangular work is balanced);
//   phase 2  64 threads roll one trajectory each through H steps, state in registers, controls read
//            from the U tile (conflict-free), reference rows broadcast from shared memory;
//   phase 3  tile-local (min, sum exp, sum exp * u) partial, then a single grid-wide merge by the last
//            CTA to finish (overflow-safe: partials carry their own min).
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "common.cuh"
#include "rng.cuh"

namespace covo {

namespace {

constexpr int TS = kTileSamples;
constexpr int TSP = TS + 4;  // padded sample stride of the E/U tile (float4 aligned, conflict-free)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ int warp_id_of(int tid) { return tid >> 5; }
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Smem {
    float* lfac;   // packed Lt (dense) or Lblk [H][16]
    float* tile;   // E (then U when the phases are not overlapped): [n_pad][TSP]
    float* utile;  // U: [n_pad][TSP]; a separate region when GEMM and rollouts overlap, else == tile
    int* prog;     // [8] progress of the GEMM warps (overlap mode)
    float* mu;     // [n_pad]
    float* ref;    // [H][8]: pos_tar(3), vel_tar(3), fdist(2 of 3 -> see fds)
    float* fds;    // [H][4]: disturbance force acting during step h
    float* cost;   // [TS]
    float* wgt;    // [TS]
    float* red;    // [32] scratch
    uint64_t* bar;   // whole-factor staging barrier
    uint64_t* cbar;  // [32] one barrier per 8-column block of the factor (Cholesky -> rollout pipeline)
};

__device__ __forceinline__ Smem carve(unsigned char* base, int n_pad, int H, int lfac_floats, int overlap) {
    Smem s;
    float* f = reinterpret_cast<float*>(base);
    s.lfac = f;
    f += (lfac_floats + 3) & ~3;
    s.tile = f;
    f += n_pad * TSP;
    s.utile = s.tile;
    if (overlap) {
        s.utile = f;
        f += n_pad * TSP;
    }
    s.prog = reinterpret_cast<int*>(f);
    f += 8;
    s.mu = f;
    f += n_pad;
    s.ref = f;
    f += H * 8;
    s.fds = f;
    f += H * 4;
    s.cost = f;
    f += TS;
    s.wgt = f;
    f += TS;
    s.red = f;
    f += 32;
    s.bar = reinterpret_cast<uint64_t*>(f);
    s.cbar = s.bar + 1;
    return s;
}

}  // namespace

static size_t rollout_smem_bytes2(int n_pad, int mode, int H, int overlap) {
    int lf = (mode == 0) ? lt_size(4 * H, n_pad) : H * 16;
    lf = (lf + 3) & ~3;
    size_t floats = (size_t)lf + (size_t)n_pad * TSP * (overlap ? 2 : 1) + 8 + n_pad + H * 8 + H * 4 + TS + TS + 32;
    return floats * sizeof(float) + 8 * 34;
}
// GEMM / rollout overlap needs a second [n_pad][TSP] tile: possible while everything fits into 227 KB (n <= 208)
static int rollout_overlap(int n_pad, int mode, int H) {
    return mode == 0 && rollout_smem_bytes2(n_pad, mode, H, 1) <= (size_t)227 * 1024;
}
int rollout_is_overlapped(int n_pad, int mode, int H) { return rollout_overlap(n_pad, mode, H); }
size_t rollout_smem_bytes(int n_pad, int mode, int H) { return rollout_smem_bytes2(n_pad, mode, H, rollout_overlap(n_pad, mode, H)); }

__global__ void __launch_bounds__(kRolloutThreads, 1) rollout_kernel(const RolloutArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int env = blockIdx.y;
    const int n = a.n, n_pad = a.n_pad, H = a.H;
    const int lfac_floats = (a.mode == 0) ? lt_size(n, n_pad) : H * 16;
    Smem sm = carve(smem_raw, n_pad, H, lfac_floats, a.overlap);
    const int tile0 = blockIdx.x * TS;                       // first local sample of this tile
    const int n_valid = min(TS, a.n_samples - tile0);        // valid samples in this tile
    const float* lfac_g = a.Lfac + (long long)env * a.lfac_stride;
    if (a.lfac_time_stride) lfac_g += (long long)min(max(a.time[env], 0), a.lfac_time_max) * a.lfac_time_stride;

    const unsigned int rng_stream_id = a.stream + (a.stream_ctr ? __ldg(a.stream_ctr) : 0u);
    COVO_STAMP(a, 32);
    // ---------------- phase 0: staging -------------------------------------------------------
    // the factor streams in behind the running Cholesky kernel (see phase 1) instead of being staged here
    const bool pipelined = a.lfac_progress != nullptr && a.mode == 0 && a.overlap;
    if (tid == 0) {
        mbar_init(sm.bar, 1);
        if (pipelined)
            for (int g = 0; g < (n_pad >> 3); ++g) mbar_init(sm.cbar + g, 1);
    }
    if (tid < 8) sm.prog[tid] = -1;
    __syncthreads();
    if (tid == 0 && !pipelined) {
        const uint32_t total = (uint32_t)lfac_floats * 4u;
        mbar_expect_tx(sm.bar, total);
        uint32_t done = 0;
        while (done < total) {  // TMA bulk copies, <= 32 KB each
            uint32_t chunk = min(total - done, 32768u);
            tma_load_1d(reinterpret_cast<unsigned char*>(sm.lfac) + done,
                        reinterpret_cast<const unsigned char*>(lfac_g) + done, chunk, sm.bar);
            done += chunk;
        }
    }
    // mean with the shift operator (controllers/covo.py:201-203) fused into the load
    const float* mu_g = a.a_mean_in + (long long)env * n;
    for (int r = tid; r < n_pad; r += blockDim.x) {
        float v = 0.f;
        if (r < n) {
            int h = r >> 2, c = r & 3;
            int hs = a.shift ? min(h + 1, H - 1) : h;
            v = mu_g[hs * 4 + c];
        }
        sm.mu[r] = v;
    }
    // reference slice: step h sees traj[min(t0 + h, T-1)] (clamped gather, dynamics/free.py:153-155);
    // h = 0 uses the targets stored in the state itself.
    const float* st_g = a.state24 + (long long)env * kStateFloats;
    const int t0 = a.time[env];
    for (int i = tid; i < H * 8; i += blockDim.x) {
        int h = i >> 3, c = i & 7;
        float v = 0.f;
        if (c < 6) {
            if (h == 0) {
                v = st_g[16 + c];
            } else {
                int row = min(t0 + h, a.traj_len - 1);
                const float* src = (c < 3 ? a.pos_traj : a.vel_traj) + ((long long)env * a.traj_stride + (long long)row * 3);
                v = src[c < 3 ? c : c - 3];
            }
        }
        sm.ref[i] = v;
    }
    for (int i = tid; i < H * 4; i += blockDim.x) {
        int h = i >> 2, c = i & 3;
        float v = 0.f;
        if (c < 3) {
            if (h == 0) v = st_g[13 + c];
            else if (a.fdist_seq) v = a.fdist_seq[((long long)env * H + (h - 1)) * 3 + c];
        }
        sm.fds[i] = v;
    }
    // eps tile E[c][s]
    if (a.eps) {
        // parity mode: eps[env][i][c] from HBM; each thread moves 8 consecutive columns of one sample
        const float* eg = a.eps + ((long long)env * a.n_samples + tile0) * n;
        const int cblocks = n_pad >> 3;
        for (int i = tid; i < TS * cblocks; i += blockDim.x) {
            int s = i % TS, cb = i / TS;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            if (s < n_valid) {
                const float* src = eg + (long long)s * n + cb * 8;
                if (cb * 8 + 8 <= n && ((((uintptr_t)src) & 15) == 0)) {
                    float4 x0 = __ldg(reinterpret_cast<const float4*>(src));
                    float4 x1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
                    v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
                    v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (cb * 8 + j < n) v[j] = __ldg(src + j);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) sm.tile[(cb * 8 + j) * TSP + s] = v[j];
        }
    } else if (a.rng_kind == 1) {
        // the reference's own stream: jax.random.split / normal restated (rng.cuh); thread -> (sample tid % TS, part tid / TS)
        const int s = tid % TS, parts = blockDim.x / TS, part = tid / TS;
        for (int c = n + part; c < n_pad; c += parts) sm.tile[c * TSP + s] = 0.f;
        if (s < n_valid) {
            jax_sample_normals((uint32_t)a.seed, (uint32_t)(a.seed >> 32), (uint32_t)(a.sample_offset + tile0 + s), (uint32_t)a.n_total,
                               n, H, a.mode == 1, part, parts, [&](int c, float z) { sm.tile[c * TSP + s] = z; });
        } else {
            for (int c = part; c < n; c += parts) sm.tile[c * TSP + s] = 0.f;
        }
    } else {
        const int blocks4 = n_pad >> 2;
        for (int i = tid; i < TS * blocks4; i += blockDim.x) {
            int s = i % TS, b = i / TS;
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            if (s < n_valid && b * 4 < n)
                philox_normal4(a.seed, rng_stream_id, (uint32_t)(a.sample_offset + tile0 + s), (uint32_t)b, z, (uint32_t)env);
#pragma unroll
            for (int j = 0; j < 4; ++j) sm.tile[(b * 4 + j) * TSP + s] = z[j];
        }
    }
    __syncthreads();
    COVO_STAMP(a, 33);
    if (!pipelined) mbar_wait(sm.bar, 0);
    COVO_STAMP(a, 34);

    // ---------------- phases 1 + 2 overlapped ------------------------------------------------------
    // Four GEMM warps produce U = clip(mu + E L^T) row group by row group (8 rows = two horizon steps; warp j takes the
    // groups j, j+4, ...; a lane owns 4 samples x 4 rows, 8 packed FFMA2 per k) into a SEPARATE tile, while the two
    // rollout warps consume it step by step -- the triangular GEMM (14 us) hides behind the 50 serial rollout steps
    // (18 us) instead of preceding them.  The GEMM warps are 2, 3, 6, 7, i.e. schedulers 2 and 3: the rollout
    // warps (0, 1) keep schedulers 0 and 1 and their FMA pipes to themselves (with GEMM warps next to them the
    // latency-bound rollout chain ran 50 % slower); warps 4 and 5 sit this phase out.
    const bool overlapped = (a.mode == 0) && a.overlap;
    constexpr int kGemmSlots = 4;
    if (overlapped && tid >= TS && ((tid >> 5) & 2)) {
        const int lane = tid & 31, w = tid >> 5, slot = (w & 1) + ((w >> 2) << 1);  // 2,3,6,7 -> 0,1,2,3
        const int sg = lane & 15, rq = lane >> 4;
        const int G = n_pad >> 3;
        volatile int* prog = sm.prog;
        int blocks_seen = 0;  // column blocks of the factor this warp has already waited for
        for (int g = slot; g < G; g += kGemmSlots) {
            const int K = min(8 * g + 8, n);
            if (pipelined)  // row group g reads column blocks 0 .. g
                for (; blocks_seen <= g; ++blocks_seen) mbar_wait(sm.cbar + blocks_seen, 0);
            const float* Erow = sm.tile + 4 * sg;
            float2 acc[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
            int off = 0;  // lt_col_offset(k)
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const float4 e = *reinterpret_cast<const float4*>(Erow + k * TSP);
                const float4 l = *reinterpret_cast<const float4*>(sm.lfac + off - (k & ~7) + 8 * g + 4 * rq);
                const float2 l01 = make_float2(l.x, l.y), l23 = make_float2(l.z, l.w);
                const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = __ffma2_rn(make_float2(ev[i], ev[i]), l01, acc[i][0]);
                    acc[i][1] = __ffma2_rn(make_float2(ev[i], ev[i]), l23, acc[i][1]);
                }
                off += n_pad - (k & ~7);
            }
            const int r0 = 8 * g + 4 * rq;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float m = sm.mu[r0 + j];
                float4 o;
                o.x = clip_(m + ((j & 1) ? acc[0][j >> 1].y : acc[0][j >> 1].x), -1.f, 1.f);
                o.y = clip_(m + ((j & 1) ? acc[1][j >> 1].y : acc[1][j >> 1].x), -1.f, 1.f);
                o.z = clip_(m + ((j & 1) ? acc[2][j >> 1].y : acc[2][j >> 1].x), -1.f, 1.f);
                o.w = clip_(m + ((j & 1) ? acc[3][j >> 1].y : acc[3][j >> 1].x), -1.f, 1.f);
                *reinterpret_cast<float4*>(sm.utile + (r0 + j) * TSP + 4 * sg) = o;
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) prog[slot] = g;
        }
    } else if (pipelined && tid == 4 * 32) {
        // warp 4 is idle in this phase: its first lane follows the Cholesky kernel's progress counter and pulls every
        // finished 8-column block of the packed factor (one contiguous piece) into shared memory with one bulk copy
        const int G = n_pad >> 3;
        const int* flag = a.lfac_progress + env;
        unsigned long long t_start;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
        for (int g = 0; g < G; ++g) {
            bool lost = false;
            while ((int)((unsigned)ld_acquire_gpu(flag) - ((unsigned)a.lfac_epoch + (unsigned)(g + 1))) < 0) {  // wrap-safe
                __nanosleep(64);
                unsigned long long t_now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
                if (t_now - t_start > 4000000000ull) {  // 4 s: the producer kernel is not running (cannot happen under
                    lost = true;                         // pipeline_ok()); release the consumers instead of hanging the device
                    break;
                }
            }
            if (lost) {
                if (blockIdx.x == 0) printf("covo rollout: Cholesky pipeline stalled at block %d (env %d)\n", g, env);
                for (int gg = g; gg < G; ++gg)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sm.cbar + gg)) : "memory");
                break;
            }
            asm volatile("fence.proxy.async;" ::: "memory");  // the block was written by generic-proxy stores of another SM
            const int off = lt_col_offset(8 * g, n_pad);
            const uint32_t bytes = (uint32_t)(lt_col_offset(min(8 * g + 8, n), n_pad) - off) * 4u;
            mbar_expect_tx(sm.cbar + g, bytes);
            tma_load_1d(sm.lfac + off, lfac_g + off, bytes, sm.cbar + g);
        }
    }
    // ---------------- phase 1: U = clip(mu + E L^T) ------------------------------------------
    if (overlapped) {
        // done by the GEMM warps above, concurrently with phase 2
    } else if (a.mode == 0) {
        const int SG = TS / 4;
        const int G = n_pad >> 3;
        const int NP = (G + 1) >> 1;
        float accA[4][8], accB[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) accA[i][j] = accB[i][j] = 0.f;
        const bool active = tid < SG * NP;
        const int sg = tid % SG, pr = tid / SG;
        const int gA = pr, gB = G - 1 - pr;
        const bool hasB = active && (gB != gA);
        if (active) {
            const int kendA = min(8 * gA + 8, n), kendB = hasB ? min(8 * gB + 8, n) : 0;
            const float* Erow = sm.tile + 4 * sg;
            int off = 0;  // lt_col_offset(k)
            int k = 0;
#pragma unroll 2
            for (; k < kendA; ++k) {
                const float4 e = *reinterpret_cast<const float4*>(Erow + k * TSP);
                const float* col = sm.lfac + off - (k & ~7);
                const float4 la0 = *reinterpret_cast<const float4*>(col + 8 * gA);
                const float4 la1 = *reinterpret_cast<const float4*>(col + 8 * gA + 4);
                const float ev[4] = {e.x, e.y, e.z, e.w};
                // packed FFMA2: one instruction per (sample, pair of rows)
                const float2 la2[4] = {make_float2(la0.x, la0.y), make_float2(la0.z, la0.w), make_float2(la1.x, la1.y),
                                       make_float2(la1.z, la1.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 r = __ffma2_rn(make_float2(ev[i], ev[i]), la2[j], make_float2(accA[i][2 * j], accA[i][2 * j + 1]));
                        accA[i][2 * j] = r.x;
                        accA[i][2 * j + 1] = r.y;
                    }
                if (hasB) {
                    const float4 lb0 = *reinterpret_cast<const float4*>(col + 8 * gB);
                    const float4 lb1 = *reinterpret_cast<const float4*>(col + 8 * gB + 4);
                    const float2 lb2[4] = {make_float2(lb0.x, lb0.y), make_float2(lb0.z, lb0.w), make_float2(lb1.x, lb1.y),
                                           make_float2(lb1.z, lb1.w)};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 r = __ffma2_rn(make_float2(ev[i], ev[i]), lb2[j], make_float2(accB[i][2 * j], accB[i][2 * j + 1]));
                            accB[i][2 * j] = r.x;
                            accB[i][2 * j + 1] = r.y;
                        }
                }
                off += n_pad - (k & ~7);
            }
#pragma unroll 4
            for (; k < kendB; ++k) {
                const float4 e = *reinterpret_cast<const float4*>(Erow + k * TSP);
                const float* col = sm.lfac + off - (k & ~7);
                const float4 lb0 = *reinterpret_cast<const float4*>(col + 8 * gB);
                const float4 lb1 = *reinterpret_cast<const float4*>(col + 8 * gB + 4);
                const float ev[4] = {e.x, e.y, e.z, e.w};
                const float2 lb2[4] = {make_float2(lb0.x, lb0.y), make_float2(lb0.z, lb0.w), make_float2(lb1.x, lb1.y),
                                       make_float2(lb1.z, lb1.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 r = __ffma2_rn(make_float2(ev[i], ev[i]), lb2[j], make_float2(accB[i][2 * j], accB[i][2 * j + 1]));
                        accB[i][2 * j] = r.x;
                        accB[i][2 * j + 1] = r.y;
                    }
                off += n_pad - (k & ~7);
            }
        }
        __syncthreads();  // everyone is done reading E; U may now overwrite it
        if (active) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int rA = 8 * gA + j;
                float4 o;
                float m = sm.mu[rA];
                o.x = clip_(m + accA[0][j], -1.f, 1.f);
                o.y = clip_(m + accA[1][j], -1.f, 1.f);
                o.z = clip_(m + accA[2][j], -1.f, 1.f);
                o.w = clip_(m + accA[3][j], -1.f, 1.f);
                *reinterpret_cast<float4*>(sm.tile + rA * TSP + 4 * sg) = o;
                if (hasB) {
                    int rB = 8 * gB + j;
                    float mb = sm.mu[rB];
                    o.x = clip_(mb + accB[0][j], -1.f, 1.f);
                    o.y = clip_(mb + accB[1][j], -1.f, 1.f);
                    o.z = clip_(mb + accB[2][j], -1.f, 1.f);
                    o.w = clip_(mb + accB[3][j], -1.f, 1.f);
                    *reinterpret_cast<float4*>(sm.tile + rB * TSP + 4 * sg) = o;
                }
            }
        }
    } else {
        // MPPI: independent 4x4 Gaussians per horizon step (controllers/mppi.py:56-66), in place
        for (int i = tid; i < TS * H; i += blockDim.x) {
            int s = i % TS, h = i / TS;
            const float* Lb = sm.lfac + h * 16;
            float e[4], u[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) e[c] = sm.tile[(4 * h + c) * TSP + s];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float acc = sm.mu[4 * h + r];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c <= r) acc = fmaf(Lb[r * 4 + c], e[c], acc);
                u[r] = clip_(acc, -1.f, 1.f);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) sm.tile[(4 * h + r) * TSP + s] = u[r];
        }
    }
    if (!overlapped) __syncthreads();

    COVO_STAMP(a, 35);
    // ---------------- phase 2: rollouts --------------------------------------------------------
    if (tid < TS) {
        // lanes past the end of a ragged last tile run the same code on the same state (their U columns
        // are clip(mu)) so the warp stays converged for the shuffles; they contribute nothing.
        const bool valid = tid < n_valid;
        QState<float> s;
        float fd[3], p0[3], v0[3];
        load_state24(st_g, s, fd, p0, v0);
        float reward_before = 0.f, sum = 0.f, disc = 1.f;
        bool done_before = false;
        const EnvConsts env_c = a.env;
        float* ps = a.pos_stats ? a.pos_stats + (long long)env * H * 6 : nullptr;
        for (int h = 0; h < H; ++h) {
            const float4 r0 = *reinterpret_cast<const float4*>(sm.ref + h * 8);
            const float4 r1 = *reinterpret_cast<const float4*>(sm.ref + h * 8 + 4);
            const float4 f4 = *reinterpret_cast<const float4*>(sm.fds + h * 4);
            const float pt[3] = {r0.x, r0.y, r0.z};
            const float vt[3] = {r0.w, r1.x, r1.y};
            const float fdh[3] = {f4.x, f4.y, f4.z};
            // reward / done of the PRE-step state (envs/quadrotor.py:243-244)
            float r = quad_reward(s, pt, vt);
            bool done = quad_terminal(s, t0 + h, env_c);
            if (overlapped && !(h & 1)) {  // rows 4h .. 4h+7 belong to row group h/2: wait for its GEMM warp
                const int g = h >> 1;
                volatile int* prog = sm.prog;
                if ((tid & 31) == 0)
                    while (prog[g % 4] < g) {
                    }
                __syncwarp();
                __threadfence_block();
            }
            float u[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) u[c] = sm.utile[(4 * h + c) * TSP + tid];
            quad_step(s, u, fdh, env_c);
            r = done_before ? reward_before : r;  // controllers/covo.py:233
            reward_before = r;
            done_before = done_before || done;
            sum = fmaf(r, disc, sum);
            disc *= a.discount;
            if (ps) {  // debug statistics of env_state.pos after the step (covo.py:236, :281)
                float vals[6] = {s.p[0], s.p[1], s.p[2], s.p[0] * s.p[0], s.p[1] * s.p[1], s.p[2] * s.p[2]};
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    float t = warp_sum(valid ? vals[c] : 0.f);
                    if ((tid & 31) == 0) atomicAdd(ps + h * 6 + c, t);
                }
            }
        }
        float cost = -sum;
        if (!valid || !(fabsf(cost) <= 3.0e38f)) cost = CUDART_INF_F;  // EXTENSION: non-finite cost -> weight 0
        if (valid && a.costs_out) a.costs_out[(long long)env * a.n_samples + tile0 + tid] = cost;
        sm.cost[tid] = cost;
        float m = warp_min(cost);
        if ((tid & 31) == 0) sm.red[tid >> 5] = m;
    }
    __syncthreads();
    if (a.samples_out) {
        float* og = a.samples_out + ((long long)env * a.n_samples + tile0) * n;
        for (int i = tid; i < TS * n; i += blockDim.x) {
            int s = i / n, r = i % n;
            if (s < n_valid) og[(long long)s * n + r] = sm.utile[r * TSP + s];
        }
    }

    COVO_STAMP(a, 36);
    // ---------------- phase 3: tile partial + grid-wide merge ---------------------------------
    float m_b = sm.red[0];
#pragma unroll
    for (int w = 1; w < TS / 32; ++w) m_b = fminf(m_b, sm.red[w]);
    const float inv_lam = 1.0f / a.lam;
    if (tid < TS) {
        float c = sm.cost[tid];
        float w = (c < CUDART_INF_F) ? expf(-(c - m_b) * inv_lam) : 0.f;
        sm.wgt[tid] = w;
        float sw = warp_sum(w);
        if ((tid & 31) == 0) sm.red[8 + (tid >> 5)] = sw;
    }
    __syncthreads();
    float s_b = 0.f;
#pragma unroll
    for (int w = 0; w < TS / 32; ++w) s_b += sm.red[8 + w];
    const int n_cta = gridDim.x;
    const int rec = kPartialHdr + n_pad;
    float* part = a.partials + ((long long)env * n_cta + blockIdx.x) * rec;
    for (int r = tid; r < n_pad; r += blockDim.x) {
        float acc = 0.f;
        if (r < n) {
            const float* row = sm.utile + r * TSP;
#pragma unroll 4
            for (int j = 0; j < TS; ++j) {
                int i = (j + r) & (TS - 1);  // rotated start: conflict-free across consecutive r
                acc = fmaf(sm.wgt[i], row[i], acc);
            }
        }
        part[kPartialHdr + r] = acc;
    }
    if (tid == 0) {
        part[0] = m_b;
        part[1] = s_b;
        part[2] = 0.f;
        part[3] = 0.f;
    }
    __threadfence();
    __syncthreads();
    __shared__ unsigned int s_ticket;
    if (tid == 0) s_ticket = atomicAdd(a.counters + env, 1u);
    __syncthreads();
    COVO_STAMP(a, 37);
    if (s_ticket != (unsigned)(n_cta - 1)) return;
    if (a.prof && tid == 0) a.prof[38] = clock64();

    // last CTA of this environment: merge all tile partials in tile order (bit-reproducible)
    __threadfence();
    const float* base = a.partials + (long long)env * n_cta * rec;
    // Scratch: everything in front of sm.prog (factor + E/U tiles) is dead now.  [n_cta] scale table (when it fits: huge N with a
    // tiny horizon does not, then the factors are recomputed from the record headers) + [ngrp][n_pad] partial sums, ngrp capped
    // by what is left (MPPI with H = 2 has 576 floats here, not the ~1000 the uncapped layout needs).
    const int scratch_floats = (int)(reinterpret_cast<float*>(sm.prog) - sm.lfac);
    const int n_cta_pad = (n_cta + 3) & ~3;
    const bool have_table = n_cta_pad + n_pad <= scratch_floats;
    float* scale = sm.lfac;
    float lm = CUDART_INF_F;
    for (int b = tid; b < n_cta; b += blockDim.x) lm = fminf(lm, __ldcg(base + (long long)b * rec));
    lm = warp_min(lm);
    if ((tid & 31) == 0) sm.red[16 + (tid >> 5)] = lm;
    __syncthreads();
    float M = CUDART_INF_F;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) M = fminf(M, sm.red[16 + w]);
    auto scale_of = [&](int b) -> float {
        if (have_table) return scale[b];
        const float mb = __ldcg(base + (long long)b * rec);
        return (mb < CUDART_INF_F) ? expf(-(mb - M) * inv_lam) : 0.f;
    };
    if (have_table)
        for (int b = tid; b < n_cta; b += blockDim.x) {
            float mb = __ldcg(base + (long long)b * rec);
            scale[b] = (mb < CUDART_INF_F) ? expf(-(mb - M) * inv_lam) : 0.f;
        }
    __syncthreads();
    // Deterministic sums over the tile partials.  Loads are issued 16 at a time so the L2 latency of the records
    // overlaps (the partials were written by other SMs).  V[r]: thread r, tiles in order.  S: the last warp,
    // lane l takes tiles l, l+32, ... in order, then a fixed butterfly.
    __shared__ float s_S;
    if (warp_id_of(tid) == (int)(blockDim.x >> 5) - 1) {
        float sp = 0.f;
        for (int b = tid & 31; b < n_cta; b += 32) sp = fmaf(__ldcg(base + (long long)b * rec + 1), scale_of(b), sp);
        sp = warp_sum(sp);
        if ((tid & 31) == 0) s_S = sp;
    }
    // V[r] = sum_b scale[b] part[b][r]: thread = (float4 column group, tile group); the tile groups run their
    // tiles in order with up to 32 float4 loads in flight (the records sit in L2, written by other SMs), then the
    // groups are added in order -- a fixed summation tree, bit-reproducible.
    float Vr[2] = {0.f, 0.f};
    {
        const int ncol4 = n_pad >> 2;                       // <= 64
        const int vfloats = scratch_floats - (have_table ? n_cta_pad : 0);
        const int ngrp = max(1, min((int)blockDim.x / ncol4, vfloats / n_pad));  // tile groups (4 at n_pad = 208)
        const int per = (n_cta + ngrp - 1) / ngrp;
        float* vpart = sm.lfac + (have_table ? n_cta_pad : 0);  // [ngrp][n_pad] behind the scale table
        const int cg = tid % ncol4, tg = tid / ncol4;
        if (tg < ngrp) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const int b_lo = tg * per, b_hi = min(b_lo + per, n_cta);
            for (int b0 = b_lo; b0 < b_hi; b0 += 8) {
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    v[j] = (b0 + j < b_hi) ? __ldcg(reinterpret_cast<const float4*>(base + (long long)(b0 + j) * rec + kPartialHdr) + cg)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (b0 + j < b_hi) {
                        const float sc = scale_of(b0 + j);
                        acc.x = fmaf(v[j].x, sc, acc.x);
                        acc.y = fmaf(v[j].y, sc, acc.y);
                        acc.z = fmaf(v[j].z, sc, acc.z);
                        acc.w = fmaf(v[j].w, sc, acc.w);
                    }
            }
            *reinterpret_cast<float4*>(vpart + tg * n_pad + 4 * cg) = acc;
        }
        __syncthreads();
        int slot = 0;
        for (int r = tid; r < n_pad; r += blockDim.x, ++slot) {
            float V = 0.f;
            for (int g = 0; g < ngrp; ++g) V += vpart[g * n_pad + r];
            if (slot < 2) Vr[slot] = V;
        }
    }
    __syncthreads();
    const float S = s_S;
    {
        int slot = 0;
        for (int r = tid; r < n_pad; r += blockDim.x, ++slot) {
            const float V = Vr[slot < 2 ? slot : 1];
            if (a.finalize) {
                if (r < n) {
                    // controllers/covo.py:270-278
                    float mu = sm.mu[r];
                    float nm = (S > 0.f) ? (V / S) * a.gamma_mean + mu * (1.0f - a.gamma_mean) : mu;
                    a.a_mean_out[(long long)env * n + r] = nm;
                    if (r < 4) a.action_out[(long long)env * 4 + r] = nm;
                }
            } else {
                a.rank_partial[(long long)env * rec + kPartialHdr + r] = V;
            }
        }
    }
    if (tid == 0) {
        if (!a.finalize) {
            float* rp = a.rank_partial + (long long)env * rec;
            rp[0] = M;
            rp[1] = S;
            rp[2] = 0.f;
            rp[3] = 0.f;
        }
        a.counters[env] = 0u;  // re-arm for the next launch
        if (a.prof) a.prof[39] = clock64();
    }
    if (!a.finalize && a.px.world > 0) {
        // fused exchange: the record goes straight into slot [step parity][this rank][env] of EVERY rank's buffer (remote stores over
        // NVLink for the peers), then -- behind a system-scope fence and the CTA barrier -- the flag of that slot is raised everywhere
        const unsigned int epoch = a.px.epoch;
        const int n_env = (int)gridDim.y;
        const long long slot = ((long long)(epoch & 1u) * a.px.world + a.px.rank) * n_env + env;
        for (int w = 0; w < a.px.world; ++w) {
            float* dst = a.px.rec[w] + slot * rec;
            int sl = 0;
            for (int r = tid; r < n_pad; r += blockDim.x, ++sl) dst[kPartialHdr + r] = Vr[sl < 2 ? sl : 1];
            if (tid == 0) {
                dst[0] = M;
                dst[1] = S;
                dst[2] = 0.f;
                dst[3] = 0.f;
            }
        }
        __threadfence_system();
        __syncthreads();
        if (tid < a.px.world) {
            unsigned int* f = a.px.flag[tid] + slot;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
        }
    }
}

// C1 epilogue: merge the per-rank (m, s, v) records gathered over NCCL, in rank order.
__global__ void merge_ranks_kernel(const MergeArgs a_in) {
    MergeArgs a = a_in;
    const int env = blockIdx.x;
    const int rec = kPartialHdr + a.n_pad;
    const float inv_lam = 1.0f / a.lam;
    if (a.flags) {
        // fused exchange: wait until every rank's record of this step has landed in this rank's buffer (written by the peers' rollout
        // kernels, see rollout_kernel), then merge in rank order exactly as after an all-gather
        const unsigned int epoch = a.stream + (a.stream_ctr ? __ldg(a.stream_ctr) : 0u) + 1u;
        const long long slot0 = (long long)(epoch & 1u) * a.world * a.n_env;
        if ((int)threadIdx.x < a.world) {
            const unsigned int* f = a.flags + slot0 + (long long)threadIdx.x * a.n_env + env;
            unsigned long long t0 = 0, t1 = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (;;) {
                unsigned int v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if ((int)(v - epoch) >= 0) break;
                __nanosleep(100);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 4000000000ull) {  // 4 s: a peer is gone; report instead of hanging the device
                    if (a.status) a.status[env] = 4;
                    break;
                }
            }
        }
        __syncthreads();
        a.gathered += slot0 * rec;
    }
    float M = CUDART_INF_F;
    for (int w = 0; w < a.world; ++w) M = fminf(M, a.gathered[((long long)w * a.n_env + env) * rec]);
    float S = 0.f;
    for (int w = 0; w < a.world; ++w) {
        const float* g = a.gathered + ((long long)w * a.n_env + env) * rec;
        float sc = (g[0] < CUDART_INF_F) ? expf(-(g[0] - M) * inv_lam) : 0.f;
        S = fmaf(g[1], sc, S);
    }
    const int H = a.n >> 2;
    // a_mean_in may be a_mean_out (in place): every thread reads the shifted entries it blends with before anybody writes
    constexpr int kPer = 4;  // entries per thread: n <= 4 * blockDim.x
    float mu[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int r = threadIdx.x + j * blockDim.x;
        mu[j] = 0.f;
        if (r < a.n) {
            const int h = r >> 2, c = r & 3;
            const int hs = a.shift ? min(h + 1, H - 1) : h;
            mu[j] = a.a_mean_in[(long long)env * a.n + hs * 4 + c];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int r = threadIdx.x + j * blockDim.x;
        if (r >= a.n) continue;
        float V = 0.f;
        for (int w = 0; w < a.world; ++w) {
            const float* g = a.gathered + ((long long)w * a.n_env + env) * rec;
            float sc = (g[0] < CUDART_INF_F) ? expf(-(g[0] - M) * inv_lam) : 0.f;
            V = fmaf(g[kPartialHdr + r], sc, V);
        }
        float nm = (S > 0.f) ? (V / S) * a.gamma_mean + mu[j] * (1.0f - a.gamma_mean) : mu[j];
        a.a_mean_out[(long long)env * a.n + r] = nm;
        if (r < 4) a.action_out[(long long)env * 4 + r] = nm;
    }
}

cudaError_t launch_rollout(const RolloutArgs& a_in, int n_env, cudaStream_t st) {
    RolloutArgs a = a_in;
    a.overlap = rollout_overlap(a.n_pad, a.mode, a.H);
    size_t smem = rollout_smem_bytes(a.n_pad, a.mode, a.H);
    static size_t configured[32] = {};
    cudaError_t e = ensure_smem_attr(rollout_kernel, smem, configured);
    if (e != cudaSuccess) return e;
    int n_cta = (a.n_samples + TS - 1) / TS;
    dim3 grid(n_cta, n_env);
    if (a.lfac_progress) {
        // programmatic dependent launch: this grid may start as soon as every CTA of the preceding kernel in the stream
        // (cholesky_kernel) has executed griddepcontrol.launch_dependents -- i.e. while the factorisation is running and
        // already resident, which is what makes spinning on its progress counter safe.  The kernel never calls
        // griddepcontrol.wait: the counter is its only dependency on the primary.
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(kRolloutThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, rollout_kernel, a);
    }
    rollout_kernel<<<grid, kRolloutThreads, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_merge(const MergeArgs& a, cudaStream_t st) {
    merge_ranks_kernel<<<a.n_env, 256, 0, st>>>(a);
    return cudaGetLastError();
}

const void* rollout_kernel_address() { return reinterpret_cast<const void*>(&rollout_kernel); }

}  // namespace covo
