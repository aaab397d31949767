// Shared declarations of the kernel family (internal; the public boundary is include/covo_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "quad_model.cuh"

namespace covo {

constexpr int kMaxH = 64;        // horizon limit: n = 4H <= 256 control dimensions
constexpr int kStateFloats = 24;  // covo_state24
constexpr int kTileSamples = 64;  // trajectories per CTA in the rollout kernel
constexpr int kRolloutThreads = 256;
constexpr int kPartialHdr = 4;  // (m, s, pad, pad) before v[n_pad] in a softmax partial record

__host__ __device__ inline int round_up8(int x) { return (x + 7) & ~7; }

// Packed, k-major ("transposed") storage of the lower Cholesky factor used by the sampler:
// column k of L holds rows r >= (k & ~7) contiguously (rows (k&~7) .. k-1 are explicit zeros so that
// every 8-row group is float4 aligned).  lt_col_offset(k) = sum_{j<k} (n_pad - (j & ~7)).
__host__ __device__ inline int lt_col_offset(int k, int n_pad) {
    int B = k >> 3, c = k & 7;
    return k * n_pad - 32 * B * (B - 1) - 8 * c * B;
}
__host__ __device__ inline int lt_size(int n, int n_pad) { return lt_col_offset(n, n_pad); }

// Dynamic shared memory, named barriers: spelled through macros so that tests/emu (a CPU stand-in for the CUDA execution model,
// test infrastructure) can run kernel logic without a GPU.
#if defined(COVO_CPU_EMU)
#define COVO_DYN_SMEM(name) unsigned char* name = emu_dyn_smem()
#define COVO_NAMED_BARRIER(id, count) emu_named_barrier(id, count)
#else
#define COVO_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define COVO_NAMED_BARRIER(id, count) asm volatile("bar.sync %0, %1;" ::"n"(id), "n"(count) : "memory")
#endif

#define COVO_STAMP(args, slot)                                              \
    do {                                                                     \
        if ((args).prof && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) (args).prof[slot] = clock64(); \
    } while (0)

// N-sharded step (BASELINE config 4): every rank writes its (min cost, sum w, sum w u) record straight into the exchange buffer of
// every rank -- peer device memory over NVLink (CUDA IPC mappings), or plain device pointers when the "ranks" are handles of one
// process -- and raises a flag there; the merge waits for the flags of all ranks.  Two slots by step parity: a rank can be at most
// one step ahead of a peer (it needs the peer's record of step t before it can finish step t).
constexpr int kMaxPeers = 8;
struct PeerExchange {
    int world = 0, rank = 0;          // world == 0: off
    unsigned int epoch = 0;           // step number + 1 (what the flags are raised to)
    float* rec[kMaxPeers] = {};       // rank w's buffer: [2][world][E][kPartialHdr + n_pad] floats ...
    unsigned int* flag[kMaxPeers] = {};  // ... and [2][world][E] flags (the step number the slot holds, + 1)
};

struct RolloutArgs {
    long long* prof = nullptr;  // optional: clock64() stamps at phase boundaries (debug)
    int n_samples;      // samples of THIS launch (this rank's shard)
    int sample_offset;  // global index of the first one (N-sharding keeps the RNG field global)
    int H, n, n_pad;
    int mode;  // 0: dense packed Lt (CoVO), 1: block-diagonal per-step 4x4 (MPPI)
    int shift;  // apply the shift operator to a_mean_in on load
    int finalize;  // last CTA merges partials and writes a_mean_out / action_out (world == 1)
    int traj_len;
    long long traj_stride;  // floats between environments in pos_traj / vel_traj (0: shared)
    float lam, gamma_mean, discount;
    EnvConsts env;
    unsigned long long seed;
    unsigned int stream;
    const unsigned int* stream_ctr = nullptr;  // optional device counter added to `stream` (CUDA-graph replay: no per-step argument)
    int rng_kind = 0;  // 0: Philox field (seed, stream); 1: JAX-compatible Threefry stream, seed = act_key words (lo = key[0], hi = key[1])
    int n_total = 0;   // rng_kind 1: global sample count N (jax.random.split(act_key, N))
    // inputs (per environment e = blockIdx.y, strides below)
    const float* state24;    // [E][24]
    const int* time;         // [E]
    const float* pos_traj;   // [E][T][3]
    const float* vel_traj;   // [E][T][3]
    const float* a_mean_in;  // [E][n]
    const float* Lfac;       // dense: [E][lt_size]; block-diag: [E][H][16] row-major lower
    const float* eps;        // optional [E][n_samples][n]  (parity mode), else nullptr
    const float* fdist_seq;  // optional [E][H][3]: force produced by step h (acts during step h+1)
    // outputs
    float* partials;         // [E][n_cta][kPartialHdr + n_pad]
    unsigned int* counters;  // [E]
    float* rank_partial;     // [E][kPartialHdr + n_pad]   (when !finalize)
    float* a_mean_out;       // [E][n]
    float* action_out;       // [E][4]
    float* costs_out;        // optional [E][n_samples]
    float* samples_out;      // optional [E][n_samples][n]  clipped samples
    float* pos_stats;        // optional [E][H][6]  sum(pos), sum(pos^2) over the samples
    long long lfac_stride;   // floats between environments in Lfac
    long long lfac_time_stride;  // CoVO-offline: factor table indexed by min(time, lfac_time_max)
    int lfac_time_max;
    // Cholesky -> rollout pipeline (dense mode, overlap layout only): the factor arrives column block by column block while
    // cholesky_kernel is still running; lfac_progress[env] >= lfac_epoch + j + 1 means column block j (8 columns) is in HBM
    const int* lfac_progress = nullptr;
    int lfac_epoch = 0;
    int overlap = 0;  // set by launch_rollout: sampling GEMM and rollouts run concurrently (separate U tile fits in smem)
    PeerExchange px;  // !finalize: also publish the rank record to every peer (fused exchange)
};

struct MergeArgs {
    int world, n, n_pad, n_env;
    float lam, gamma_mean;
    int shift;
    // fused exchange: gathered = this rank's exchange buffer [2][world][E][rec], flags [2][world][E]; the kernel waits until every
    // rank's flag of the slot says step (stream + *stream_ctr)
    const unsigned int* flags = nullptr;
    unsigned int stream = 0;
    const unsigned int* stream_ctr = nullptr;
    int* status = nullptr;   // [E]: 4 = a peer's record did not arrive within the watchdog time
    const float* gathered;   // [world][E][kPartialHdr + n_pad]
    const float* a_mean_in;  // [E][n]
    float* a_mean_out;       // [E][n]
    float* action_out;       // [E][4]
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: remember what was configured on each one.
template <class K>
inline cudaError_t ensure_smem_attr(K kernel, size_t bytes, size_t (&configured)[32]) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev &= 31;
    if (bytes > configured[dev]) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        configured[dev] = bytes;
    }
    return cudaSuccess;
}

cudaError_t launch_rollout(const RolloutArgs& a, int n_env, cudaStream_t st);
cudaError_t launch_merge(const MergeArgs& a, cudaStream_t st);
const void* rollout_kernel_address();  // for CUDA-graph node lookup (capi.cu)
size_t rollout_smem_bytes(int n_pad, int mode, int H);
int rollout_is_overlapped(int n_pad, int mode, int H);  // GEMM / rollout overlap layout in use (needed by the Cholesky pipeline)

}  // namespace covo
