// TEST INFRASTRUCTURE: a minimal CPU stand-in for the CUDA execution model, so that kernel LOGIC (indexing, barriers, warp
// collectives, shared-memory hand-overs) can be exercised without a GPU.  It shadows <cuda_runtime.h> for translation units
// compiled with g++ -DCOVO_CPU_EMU -Itests/emu.  One CTA at a time; every CUDA thread is a cooperative fiber (ucontext), so
// __syncthreads / named barriers / __shfl / __ballot block a fiber until its peers arrive, exactly one fiber runs at any moment and
// the run is deterministic.  What it does NOT model: memory ordering, bank conflicts, timing, concurrency between CTAs.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(x) alignas(x)

using std::max;
using std::min;

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr cudaError_t cudaErrorInvalidValue = 1;
constexpr int cudaFuncAttributeMaxDynamicSharedMemorySize = 0;
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline long long clock64() { return 0; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
using std::exp2f;
using std::log2f;

namespace emu {

struct Barrier {
    int expected = 0, arrived = 0;
    unsigned gen = 0;
};

struct Block {
    dim3 grid, block_dim, block_idx;
    int n_threads = 0, cur = -1;
    std::vector<ucontext_t> ctx;
    std::vector<std::unique_ptr<char[]>> stacks;
    std::vector<char> done;
    ucontext_t sched;
    std::vector<unsigned char> smem;
    Barrier all;
    Barrier named[16];
    std::vector<Barrier> warp;                 // one per warp
    std::vector<unsigned long long> slot;      // [n_threads] exchange slots of the warp collectives
    std::function<void()> body;
};

inline Block*& cur() {
    static Block* b = nullptr;
    return b;
}
inline ucontext_t*& sched_ptr() {  // the scheduler context fibers yield to (one CTA, or all CTAs of a cluster)
    static ucontext_t* p = nullptr;
    return p;
}
struct Cluster;
inline Cluster*& cur_cluster() {
    static Cluster* c = nullptr;
    return c;
}
inline long long& progress() {  // barrier arrivals + thread exits: a scheduler pass without any is a deadlock
    static long long p = 0;
    return p;
}

inline void yield() {
    Block* b = cur();
    swapcontext(&b->ctx[b->cur], sched_ptr());
}

inline void wait(Barrier& bar, int expected) {
    if (bar.expected == 0) bar.expected = expected;
    if (bar.expected != expected) {
        fprintf(stderr, "emu: barrier used with inconsistent participant counts (%d vs %d)\n", bar.expected, expected);
        abort();
    }
    const unsigned g = bar.gen;
    ++progress();
    if (++bar.arrived == bar.expected) {
        bar.arrived = 0;
        bar.expected = 0;
        ++bar.gen;
    } else {
        while (bar.gen == g) yield();
    }
}

inline int tid_linear() { return cur()->cur; }

inline void trampoline() {
    Block* b = cur();
    b->body();
    b->done[b->cur] = 1;
    swapcontext(&b->ctx[b->cur], sched_ptr());
}

// run ONE CTA: body() is executed once per thread (fiber)
inline void prepare_block(Block& b, ucontext_t* sched, size_t stack_bytes) {
    const int n = b.n_threads;
    b.ctx.resize(n);
    b.stacks.clear();
    b.done.assign(n, 0);
    b.warp.assign((n + 31) / 32, Barrier());
    b.slot.assign(n, 0);
    for (int t = 0; t < n; ++t) {
        b.stacks.emplace_back(new char[stack_bytes]);
        getcontext(&b.ctx[t]);
        b.ctx[t].uc_stack.ss_sp = b.stacks[t].get();
        b.ctx[t].uc_stack.ss_size = stack_bytes;
        b.ctx[t].uc_link = sched;
        makecontext(&b.ctx[t], (void (*)())trampoline, 0);
    }
}

inline void run_block(Block& b, size_t stack_bytes = 256 * 1024) {
    cur() = &b;
    sched_ptr() = &b.sched;
    const int n = b.n_threads;
    prepare_block(b, &b.sched, stack_bytes);
    int live = n;
    while (live > 0) {
        const long long before = progress();
        for (int t = 0; t < n; ++t) {
            if (b.done[t]) continue;
            b.cur = t;
            swapcontext(&b.sched, &b.ctx[t]);
            if (b.done[t]) {
                --live;
                ++progress();
            }
        }
        if (live > 0 && progress() == before) {
            fprintf(stderr, "emu: deadlock -- %d threads wait at barriers that can never complete\n", live);
            abort();
        }
    }
    cur() = nullptr;
}

// A thread-block cluster: the CTAs run interleaved (all fibers of all CTAs under one scheduler), each with its own shared memory;
// distributed-shared-memory stores and the cluster barrier go through this object.
struct Cluster {
    std::vector<Block*> blocks;
    Barrier all;
    ucontext_t sched;
};

inline void run_cluster(Cluster& c, size_t stack_bytes = 256 * 1024) {
    cur_cluster() = &c;
    sched_ptr() = &c.sched;
    int live = 0;
    for (Block* b : c.blocks) {
        prepare_block(*b, &c.sched, stack_bytes);
        live += b->n_threads;
    }
    while (live > 0) {
        const long long before = progress();
        for (Block* b : c.blocks)
            for (int t = 0; t < b->n_threads; ++t) {
                if (b->done[t]) continue;
                cur() = b;
                b->cur = t;
                swapcontext(&c.sched, &b->ctx[t]);
                if (b->done[t]) {
                    --live;
                    ++progress();
                }
            }
        if (live > 0 && progress() == before) {
            fprintf(stderr, "emu: cluster deadlock -- %d threads wait at barriers that can never complete\n", live);
            abort();
        }
    }
    cur() = nullptr;
    cur_cluster() = nullptr;
}

struct Idx {
    unsigned x, y, z;
};

}  // namespace emu

// ---- the CUDA built-ins the kernels use ---------------------------------------------------------------------------------------
struct EmuThreadIdx {
    struct P { operator unsigned() const { return (unsigned)emu::tid_linear(); } };
    P x;  // 1-D blocks only
};
static EmuThreadIdx threadIdx;
struct EmuBlockIdx {
    struct PX { operator unsigned() const { return emu::cur()->block_idx.x; } };
    struct PY { operator unsigned() const { return emu::cur()->block_idx.y; } };
    PX x;
    PY y;
};
static EmuBlockIdx blockIdx;
struct EmuBlockDim {
    struct PX { operator unsigned() const { return emu::cur()->block_dim.x; } };
    PX x;
};
static EmuBlockDim blockDim;
struct EmuGridDim {
    struct PX { operator unsigned() const { return emu::cur()->grid.x; } };
    struct PY { operator unsigned() const { return emu::cur()->grid.y; } };
    PX x;
    PY y;
};
static EmuGridDim gridDim;

inline void __syncthreads() { emu::wait(emu::cur()->all, emu::cur()->n_threads); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::wait(emu::cur()->warp[emu::tid_linear() >> 5], 32); }
inline void emu_named_barrier(int id, int count) { emu::wait(emu::cur()->named[id], count); }

template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) <= 8, "emu shuffle: 4- or 8-byte types");
    emu::Block* b = emu::cur();
    const int t = emu::tid_linear();
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    b->slot[t] = raw;
    __syncwarp();
    raw = b->slot[(t & ~31) | ((t ^ lane_mask) & 31)];
    __syncwarp();
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}

template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "emu shuffle: 4- or 8-byte types");
    emu::Block* b = emu::cur();
    const int t = emu::tid_linear();
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    b->slot[t] = raw;
    __syncwarp();
    raw = b->slot[(t & ~31) | (src_lane & 31)];
    __syncwarp();
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}

inline unsigned __ballot_sync(unsigned, bool p) {
    emu::Block* b = emu::cur();
    const int t = emu::tid_linear();
    b->slot[t] = p ? 1ull : 0ull;
    __syncwarp();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l)
        if (b->slot[(t & ~31) | l]) m |= 1u << l;
    __syncwarp();
    return m;
}

inline unsigned char* emu_dyn_smem() { return emu::cur()->smem.data(); }

// launch helper: runs the CTAs of a grid one after the other
template <class Kernel, class Args>
inline void emu_launch(Kernel kernel, dim3 grid, int threads, size_t smem_bytes, const Args& args) {
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            emu::Block b;
            b.grid = grid;
            b.block_dim = dim3(threads);
            b.block_idx = dim3(bx, by);
            b.n_threads = threads;
            b.smem.assign(smem_bytes + 64, 0xCD);  // poison: reads of never-written shared memory show up as garbage
            b.body = [&]() { kernel(args); };
            emu::run_block(b);
        }
}

// ---- thread-block clusters: distributed shared memory, mbarriers with transaction counts, cluster barrier --------------------
struct EmuMbar {  // lives in the 8 bytes of the kernel's mbarrier object
    int tx;                 // outstanding transaction bytes (may go negative: data can arrive before the expectation is posted)
    signed char pending;    // outstanding arrivals of the current phase
    signed char count;      // arrivals per phase
    unsigned short phase;
};
static_assert(sizeof(EmuMbar) == 8, "EmuMbar must fit an mbarrier object");
inline void emu_mbar_check(EmuMbar* m) {
    if (m->pending == 0 && m->tx == 0) {
        ++m->phase;
        m->pending = m->count;
        ++emu::progress();
    }
}
inline void emu_mbar_init(void* bar, int count) {
    EmuMbar* m = reinterpret_cast<EmuMbar*>(bar);
    m->tx = 0;
    m->pending = m->count = (signed char)count;
    m->phase = 0;
}
inline void emu_mbar_expect_tx(void* bar, int bytes) {  // mbarrier.arrive.expect_tx
    EmuMbar* m = reinterpret_cast<EmuMbar*>(bar);
    m->tx += bytes;
    --m->pending;
    emu_mbar_check(m);
}
inline void emu_mbar_wait(void* bar, unsigned parity) {  // mbarrier.try_wait.parity loop
    EmuMbar* m = reinterpret_cast<EmuMbar*>(bar);
    while ((m->phase & 1u) == (parity & 1u)) emu::yield();
}
inline unsigned emu_cluster_rank() { return emu::cur()->block_idx.x % (unsigned)emu::cur_cluster()->blocks.size(); }
// st.async to CTA `rank` of the cluster: the value lands at the same shared-memory offset there and completes 4 bytes on that CTA's mbarrier
inline void emu_dsmem_st_signal(float* local_addr, unsigned rank, float v, void* local_bar) {
    emu::Block* me = emu::cur();
    emu::Block* peer = emu::cur_cluster()->blocks[rank];
    const size_t off = reinterpret_cast<unsigned char*>(local_addr) - me->smem.data();
    const size_t boff = reinterpret_cast<unsigned char*>(local_bar) - me->smem.data();
    memcpy(peer->smem.data() + off, &v, 4);
    EmuMbar* m = reinterpret_cast<EmuMbar*>(peer->smem.data() + boff);
    m->tx -= 4;
    emu_mbar_check(m);
}
inline void emu_dsmem_st_signal64(double* local_addr, unsigned rank, double v, void* local_bar) {
    emu::Block* me = emu::cur();
    emu::Block* peer = emu::cur_cluster()->blocks[rank];
    const size_t off = reinterpret_cast<unsigned char*>(local_addr) - me->smem.data();
    const size_t boff = reinterpret_cast<unsigned char*>(local_bar) - me->smem.data();
    memcpy(peer->smem.data() + off, &v, 8);
    EmuMbar* m = reinterpret_cast<EmuMbar*>(peer->smem.data() + boff);
    m->tx -= 8;
    emu_mbar_check(m);
}
// cp.async.bulk shared::cta -> shared::cluster: `bytes` from this CTA's src to the same-offset-as-dst_local location in CTA `rank`
inline void emu_dsmem_bulk_copy(void* dst_local, const void* src_local, unsigned bytes, unsigned rank, void* local_bar) {
    emu::Block* me = emu::cur();
    emu::Block* peer = emu::cur_cluster()->blocks[rank];
    const size_t off = reinterpret_cast<unsigned char*>(dst_local) - me->smem.data();
    const size_t boff = reinterpret_cast<unsigned char*>(local_bar) - me->smem.data();
    memcpy(peer->smem.data() + off, src_local, bytes);
    EmuMbar* m = reinterpret_cast<EmuMbar*>(peer->smem.data() + boff);
    m->tx -= (int)bytes;
    emu_mbar_check(m);
}
inline void emu_cluster_barrier() {
    emu::Cluster* c = emu::cur_cluster();
    int total = 0;
    for (emu::Block* b : c->blocks) total += b->n_threads;
    emu::wait(c->all, total);
}

// launch helper for clustered grids: clusters of `cluster_x` consecutive CTAs along x, one cluster at a time
template <class Kernel, class Args>
inline void emu_launch_cluster(Kernel kernel, dim3 grid, int cluster_x, int threads, size_t smem_bytes, const Args& args) {
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx0 = 0; bx0 < grid.x; bx0 += cluster_x) {
            std::vector<std::unique_ptr<emu::Block>> owned;
            emu::Cluster c;
            for (int r = 0; r < cluster_x; ++r) {
                owned.emplace_back(new emu::Block());
                emu::Block& b = *owned.back();
                b.grid = grid;
                b.block_dim = dim3(threads);
                b.block_idx = dim3(bx0 + r, by);
                b.n_threads = threads;
                b.smem.assign(smem_bytes + 64, 0xCD);
                b.body = [&]() { kernel(args); };
                c.blocks.push_back(&b);
            }
            emu::run_cluster(c);
        }
}
