// optimize_sigma (controllers/covo.py:116-132) WITHOUT an eigen-decomposition ("D1-D3").  Default for single
// environments since the end of round 2 (COVO_SIGMA=tridiag selects the tridiagonal path E1-E3 of sigma.cu, which batches use);
// tools/studies/ holds the numerical studies behind it.  The reference computes
//     Sigma = U diag(s) U^T,  s_k = exp(c/2) / sqrt(o_k),  o_k = lambda_k - lambda_min + 1e-2,  c = (4 n log sigma + sum log o_k) / n
// which is  Sigma = exp(c/2) * A^(-1/2)  with  A = (R + R^T)/2 - lambda_min I + 1e-2 I  and  sum log o_k = log det A.
// So only lambda_min, log det A and the matrix function A^(-1/2) are needed:
//   D1  lanczos_cluster_kernel  lambda_min of R by Lanczos in float64 on the float32 matrix, 8-CTA cluster, ADAPTIVE length: a
//                               fifth warp per CTA follows the smallest Ritz value and its residual while the recurrence runs
//                               (16 .. 64 steps; along closed loops 24 steps suffice for 56 % of the Hessians, 32 for 94 %, 48 for
//                               all -- a fixed 24 left lambda_min off by up to 6e-2, i.e. A indefinite)
//   D2  gjb_inverse_kernel_t    one 8-CTA cluster per pole: w_j (A + t_j I)^-1 by blocked Gauss-Jordan IN FLOAT64, one more for log det A.
//                               x^(-1/2) ~ sum_j w_j / (x + t_j): Zolotarev's partial fractions on [1e-2, M], 13 poles (exact inverses
//                               need no more: 2e-7), the interval ladder of E2
//   D3  combine_kernel          Sigma = exp(c/2) * sum_j w_j (A + t_j I)^-1, written symmetric
// Measured on B200 (n = 200): D1 ~2.4 us per Lanczos step (55 .. 70 us) + D2 68 us + D3 7 us against 283 us for E1 + E2 + E3.  Hardware
// facts that shaped it (tools/microbench): float64 FMA issues at 64 lanes / clock / SM with 9 cycles between dependent operations
// (mma.sync f64: the same rate), st.async to a cluster peer arrives after ~250 cycles, a 1.8 KB bulk DSMEM copy after ~700.
// Accuracy (GPU tests): Sigma within 1e-7 .. 2e-7 (relative Frobenius) of the float64 eigen-decomposition of the same float32
// Hessian; float32 LAPACK eigh, the reference's arithmetic: 4e-7 .. 1e-5.

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "sigma.cuh"

namespace covo {

namespace {

constexpr double kOffset = 1e-2;  // controllers/covo.py:121

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace

struct DenseArgs {
    int n, n_pad;
    float sample_sigma;
    const float* R;      // [E][n][n]
    double* scal;        // [E][4]: lambda_min, upper bound of the spectrum, log det A, Lanczos steps taken
    float* Xbuf;         // [E][kDensePoles][n][n]  upper triangles of w_j (A + t_j I)^-1
    float* cov;          // [E][n][n]
    const double* zolo;  // the dense path's ladder: [kZoloLadder][2][kDensePoles] (zolotarev_table_dense)
    int* status;         // [E]
    float* Asym = nullptr;       // optional [E][n][n]: (R + R^T)/2 written by the Lanczos kernel for the factorisation kernels
    long long* prof = nullptr;   // optional clock64() stamps (slots 48..)
};

#define DENSE_STAMP(slot)                                                        \
    do {                                                                         \
        if (a.prof && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) a.prof[slot] = clock64(); \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------------------
// Thread-block-cluster primitives shared by D1 and D2: distributed shared memory stores that carry their own completion
// signal (st.async ... mbarrier::complete_tx), mbarrier waits, the cluster barrier.  tests/emu provides CPU stand-ins.
// ---------------------------------------------------------------------------------------------------------------------------
#if defined(COVO_CPU_EMU)
__device__ __forceinline__ unsigned gjb_rank() { return emu_cluster_rank(); }
__device__ __forceinline__ void gjb_cluster_sync() { emu_cluster_barrier(); }
__device__ __forceinline__ void gjb_cluster_arrive() {}
__device__ __forceinline__ void gjb_cluster_wait() { emu_cluster_barrier(); }
__device__ __forceinline__ void gjb_mbar_init(unsigned long long* b, int count) { emu_mbar_init(b, count); }
__device__ __forceinline__ void gjb_mbar_expect(unsigned long long* b, int bytes) { emu_mbar_expect_tx(b, bytes); }
__device__ __forceinline__ void gjb_mbar_wait(unsigned long long* b, unsigned parity) { emu_mbar_wait(b, parity); }
__device__ __forceinline__ void gjb_send(float* dst_local, unsigned rank, float v, unsigned long long* bar_local) {
    emu_dsmem_st_signal(dst_local, rank, v, bar_local);
}
__device__ __forceinline__ void gjb_send64(double* dst_local, unsigned rank, double v, unsigned long long* bar_local) {
    emu_dsmem_st_signal64(dst_local, rank, v, bar_local);
}
__device__ __forceinline__ void gjb_bulk_send(void* dst_local, const void* src_local, unsigned bytes, unsigned rank, unsigned long long* bar_local) {
    emu_dsmem_bulk_copy(dst_local, src_local, bytes, rank, bar_local);
}
__device__ __forceinline__ void gjb_fence_async_proxy() {}
__device__ __forceinline__ float gjb_rcp(float x) { return 1.0f / x; }
#else
__device__ __forceinline__ unsigned gjb_s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned gjb_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void gjb_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// split form: only execution ordering is needed (who has finished reading which ring slot); the data travels with its own mbarrier
__device__ __forceinline__ void gjb_cluster_arrive() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void gjb_cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ void gjb_mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gjb_s32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void gjb_mbar_expect(unsigned long long* b, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(gjb_s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gjb_mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GJB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GJB_DONE_%=;\n"
        "bra GJB_WAIT_%=;\n"
        "GJB_DONE_%=:\n"
        "}\n" ::"r"(gjb_s32(b)),
        "r"(parity)
        : "memory");
}
// one float into the shared memory of CTA `rank` of the cluster (same offset as here), 4 bytes completed on that CTA's mbarrier
__device__ __forceinline__ void gjb_send(float* dst_local, unsigned rank, float v, unsigned long long* bar_local) {
    unsigned d, b;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(gjb_s32(dst_local)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(b) : "r"(gjb_s32(bar_local)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(d), "f"(v), "r"(b) : "memory");
}
__device__ __forceinline__ void gjb_send64(double* dst_local, unsigned rank, double v, unsigned long long* bar_local) {
    unsigned d, b;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(gjb_s32(dst_local)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(b) : "r"(gjb_s32(bar_local)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(d), "l"(__double_as_longlong(v)), "r"(b)
                 : "memory");
}
// One bulk copy (multiple of 16 bytes, 16-byte aligned) from this CTA's shared memory into CTA `rank` (same offset as dst_local),
// counted into that CTA's mbarrier.  Measured: publishing a 224-float row as 224 st.async messages made the DSMEM message rate
// (~2 cycles per message per SM) the bottleneck of D2'' (3584 messages = 7000 cycles per block step); one copy per row and
// destination is 16 messages per step.
__device__ __forceinline__ void gjb_bulk_send(void* dst_local, const void* src_local, unsigned bytes, unsigned rank, unsigned long long* bar_local) {
    unsigned d, b;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(gjb_s32(dst_local)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(b) : "r"(gjb_s32(bar_local)), "r"(rank));
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "r"(gjb_s32(src_local)),
                 "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void gjb_fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// tx-count of an mbarrier of THIS CTA, for data its own threads stored with ordinary stores (release fence first; the waiters acquire)
__device__ __forceinline__ void gjb_mbar_complete_tx_local(unsigned long long* b, unsigned bytes) {
    __threadfence_block();
    asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(gjb_s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ float gjb_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * fmaf(-x, r, 2.0f);
}
#endif

// ---------------------------------------------------------------------------------------------------------------------------
// D1 lanczos_cluster_kernel -- the Lanczos recurrence on an 8-CTA cluster.  Measured on B200: scalar float64 instructions issue at
// ~16 lanes / clock / SM, so a 200 x 200 product is >= 2500 cycles on one SM whatever the layout; the only way down is more SMs.
// CTA c keeps rows [c R, (c + 1) R) of (R + R^T)/2 in REGISTERS as float64, four threads per row (columns 4 k + p), and per step
//     y = A v_k,  u = y - beta_{k-1} v_{k-1},  warp partials of u.v_k and u.u,
// then ONE exchange: every row leader sends its u and every warp its two partials to all 8 CTAs with st.async (the stores count
// themselves into the receiver's mbarrier: no fence, no cluster barrier), everybody waits for its own mbarrier and finishes
// alpha, beta^2 = u.u - alpha^2 and v_{k+1} = (u - alpha v_k) / beta redundantly.  The all-to-all makes every step an implicit
// barrier, so two buffers (step parity) are enough.
//
// How many steps: the smallest Ritz value must be within << 1e-2 * 2e-5 = 2e-7 ABSOLUTE of lambda_min (the offset 1e-2 of
// controllers/covo.py:121 sets the scale, not |R| ~ 1e3), and the number of steps that takes depends on the Hessian: along closed
// loops (tools/studies/lanczos_k_needed.py) 16 .. 48, with 24 enough for only 56 % of them -- and an unconverged Ritz value 1e-2 above
// lambda_min makes A indefinite.  So the length is adaptive: a fifth warp per CTA (the CHECKER) follows the recurrence at the
// checkpoints k = 16, 20, 24, ...: smallest eigenvalue theta of the k x k Lanczos matrix T_k (33-way multisection on Sturm
// counts, three float32 rounds, then float64 with both bracket ends verified, then Newton from below) and the residual of its
// Ritz pair, beta_{k-1} |s_{k-1}| (s = eigenvector of T_k; |theta - lambda| <= residual^2 / gap).  It works in the shadow of the
// next steps; the Lanczos warps look at the verdict for checkpoint c when they have finished step c + 3 and stop there.  Every
// CTA runs its own checker on the same numbers (alpha, beta are bit-identical across the cluster), so all CTAs stop at the same
// step without another exchange.  No convergence within 64 steps (never seen on a CoVO Hessian): lambda_min is reported as
// theta - residual (A stays positive definite), status 3, and the host moves the handle to the tridiagonal path.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int LC_CL = 8;            // CTAs
#ifndef COVO_LC_TPR
#define COVO_LC_TPR 4
#endif
constexpr int LC_TPR = COVO_LC_TPR;  // threads per matrix row (column phases): 4, or 8 (-DCOVO_LC_TPR=8: half the matrix-vector chain per thread, twice the warps)
constexpr int LC_T = 32 * LC_TPR;   // Lanczos threads per CTA: 32 row slots x LC_TPR column phases
constexpr int LC_WPC = LC_T / 32;   // Lanczos warps per CTA
constexpr int LC_TT = LC_T + 32;    // ... plus the checker warp
constexpr int LC_KMAX = 224 / LC_TPR;  // columns per thread (n <= 224)
constexpr int kLanczosMax = 64;     // most Lanczos steps
constexpr int kLanczosFirstCheck = 16, kLanczosCheckEvery = 4;
#ifndef COVO_LANCZOS_LAG
#define COVO_LANCZOS_LAG 3
#endif
constexpr int kLanczosLag = COVO_LANCZOS_LAG;
constexpr double kRitzTol2 = 9e-10;  // residual^2 below which the Ritz value counts as converged: error <= 9e-10 / gap

struct LcSmem {
    double v[2][224];                 // normalised Lanczos vector v_k (by step parity), all n entries, zero beyond n
    double u[2][224];                 // u = A v_k - beta_{k-1} v_{k-1}, gathered from all CTAs (by step parity)
    double part[2][LC_CL * LC_WPC][2];  // (u.v_k, u.u) partial of every warp of the cluster
    double al[kLanczosMax], be[kLanczosMax];  // T: diagonal, off-diagonal (be[k-1] couples step k to the next vector)
    double red[LC_WPC];
    unsigned long long bar[2];
    // the checker's tables: 1 / beta_i, beta_{i-1} / beta_i, float32 copies, the eigenvector recurrence of the last evaluation
    double ib[kLanczosMax], cc[kLanczosMax], qv[kLanczosMax], bv[kLanczosMax];
    float alf[kLanczosMax], ibf[kLanczosMax], ccf[kLanczosMax];
    int qe[kLanczosMax], bve[kLanczosMax];
    // hand-over between the Lanczos warps and the checker (volatile accesses + __threadfence_block)
    int progress;     // Lanczos steps completed: al / be are valid below it
    int final_k;      // != 0: the recurrence has ended after this many steps
    int verdict_k;    // the last checkpoint the checker has judged
    int converged_k;  // != 0: the checkpoint at which the smallest Ritz value had converged
};

#if defined(COVO_CPU_EMU)
__device__ __forceinline__ void lc_main_sync() { emu_named_barrier(1, LC_T); }
__device__ __forceinline__ void lc_pause() { emu::yield(); }
__device__ __forceinline__ void lc_published() { ++emu::progress(); }
__device__ __forceinline__ void lc_fence() {}
#else
__device__ __forceinline__ void lc_main_sync() { asm volatile("bar.sync 1, %0;" ::"n"(LC_T) : "memory"); }
__device__ __forceinline__ void lc_pause() { __nanosleep(20); }
__device__ __forceinline__ void lc_published() {}
__device__ __forceinline__ void lc_fence() { __threadfence_block(); }
#endif
__device__ __forceinline__ int lc_load(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void lc_store(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }

// Remote addresses of one thread's three exchange targets in CTA `rank`: its row entry of u, its warp's partial pair, the barrier
// (all for parity 0; parity 1 is a fixed byte offset away) -- mapa once, outside the loop.
#if defined(COVO_CPU_EMU)
struct LcRemote {
    double* u;
    double* part;
    unsigned long long* bar;
    unsigned rank;
};
__device__ __forceinline__ LcRemote lc_remote(double* u, double* part, unsigned long long* bar, unsigned rank) { return LcRemote{u, part, bar, rank}; }
__device__ __forceinline__ void lc_send(const LcRemote& r, int what, int parity, double v) {
    double* dst = (what == 0) ? r.u + parity * 224 : r.part + parity * (LC_CL * LC_WPC * 2) + (what - 1);
    emu_dsmem_st_signal64(dst, r.rank, v, r.bar + parity);
}
#else
struct LcRemote {
    unsigned u, part, bar;
};
__device__ __forceinline__ LcRemote lc_remote(double* u, double* part, unsigned long long* bar, unsigned rank) {
    LcRemote r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r.u) : "r"(gjb_s32(u)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r.part) : "r"(gjb_s32(part)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r.bar) : "r"(gjb_s32(bar)), "r"(rank));
    return r;
}
__device__ __forceinline__ void lc_send(const LcRemote& r, int what, int parity, double v) {
    const unsigned dst = (what == 0) ? r.u + parity * 224 * 8 : r.part + parity * (LC_CL * LC_WPC * 2 * 8) + (what - 1) * 8;
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst), "l"(__double_as_longlong(v)),
                 "r"(r.bar + parity * 8)
                 : "memory");
}
#endif

// ---- the checker's arithmetic on T_k = tridiag(al[0..k), be[0..k-1)) --------------------------------------------------------
// Q_0 = 1, Q_{i+1} = ((al_i - x) Q_i - be_{i-1} Q_{i-1}) / be_i: the leading principal minors of T - x divided by be_0 ... be_i
// (same signs: the be are positive), i.e. up to alternating signs the components of the solution of (T - x) s = 0 -- the numbers
// stay of the size of an eigenvector instead of a product of k pivots.  #{sign changes in Q_0 .. Q_k} = #{eigenvalues below x}.
// The last member is left undivided (be_{k-1} belongs to the next step).  Only "is any eigenvalue below x" is needed.
__device__ __forceinline__ bool lc_any_below_f32(const LcSmem& sm, int k, float x) {
    float qm = 0.f, q = 1.f;
    bool below = false;
    for (int i = 0; i < k - 1; ++i) {
        const float w = (sm.alf[i] - x) * sm.ibf[i];
        const float qn = fmaf(w, q, -sm.ccf[i] * qm);
        qm = q;
        q = qn;
        if (fabsf(q) > 1e18f) {  // keep the pair inside the float32 range (the signs are all that matters)
            q *= 0x1p-80f;
            qm *= 0x1p-80f;
        } else if (fabsf(q) < 1e-18f && fabsf(qm) < 1e-18f) {
            q *= 0x1p80f;
            qm *= 0x1p80f;
        }
        below = below || !(q > 0.f);  // Q_0 = 1 > 0: no eigenvalue below x <=> every member stays positive
    }
    const float bm = (k >= 2) ? (float)sm.be[k - 2] : 0.f;
    const float ql = fmaf(sm.alf[k - 1] - x, q, -bm * qm);
    return below || !(ql > 0.f);
}

__device__ __forceinline__ bool lc_any_below_f64(const LcSmem& sm, int k, double x) {
    double qm = 0.0, q = 1.0;
    bool below = false;
    for (int i = 0; i < k - 1; ++i) {
        const double w = (sm.al[i] - x) * sm.ib[i];
        const double qn = fma(w, q, -sm.cc[i] * qm);
        qm = q;
        q = qn;
        if ((i & 7) == 7) {
            if (fabs(q) > 1e60) {
                q *= 0x1p-256;
                qm *= 0x1p-256;
            } else if (fabs(q) < 1e-60 && fabs(qm) < 1e-60) {
                q *= 0x1p256;
                qm *= 0x1p256;
            }
        }
        below = below || !(q > 0.0);
    }
    const double bm = (k >= 2) ? sm.be[k - 2] : 0.0;
    const double ql = fma(sm.al[k - 1] - x, q, -bm * qm);
    return below || !(ql > 0.0);
}

// Q_k(x) and its derivative (Newton), the members Q_0 .. Q_{k-1} parked in qv / qe (value, binary exponent of the scale applied)
__device__ __forceinline__ void lc_newton_eval(LcSmem& sm, int k, double x, bool park, double& f, double& df) {
    double qm = 0.0, q = 1.0, dm = 0.0, d = 0.0;
    int sc = 0;
    for (int i = 0; i < k - 1; ++i) {
        if (park) {
            sm.qv[i] = q;
            sm.qe[i] = sc;
        }
        const double w = (sm.al[i] - x) * sm.ib[i];
        const double qn = fma(w, q, -sm.cc[i] * qm);
        const double dn = fma(w, d, -fma(sm.ib[i], q, sm.cc[i] * dm));
        qm = q;
        q = qn;
        dm = d;
        d = dn;
        if ((i & 7) == 7) {
            if (fabs(q) > 1e60 || fabs(d) > 1e60) {
                q *= 0x1p-256, qm *= 0x1p-256, d *= 0x1p-256, dm *= 0x1p-256;
                sc -= 256;
            } else if (fabs(q) < 1e-60 && fabs(qm) < 1e-60) {
                q *= 0x1p256, qm *= 0x1p256, d *= 0x1p256, dm *= 0x1p256;
                sc += 256;
            }
        }
    }
    if (park) {
        sm.qv[k - 1] = q;
        sm.qe[k - 1] = sc;
    }
    const double bm = (k >= 2) ? sm.be[k - 2] : 0.0;
    const double t = sm.al[k - 1] - x;
    f = fma(t, q, -bm * qm);
    df = fma(t, d, -(q + bm * dm));
}

// The checker warp: judges T_k.  Returns theta (smallest eigenvalue of T_k), res2 (squared residual of its Ritz pair), gu (upper
// Gershgorin bound of T_k).  k_prev: the tables are complete below k_prev - 1.
__device__ __forceinline__ void lc_judge(LcSmem& sm, int k, int k_prev, int lane, double& theta, double& res2, double& gu_out) {
    // tables for the new rows (be_{k-1} is not part of T_k, but it will be of the next checkpoint's matrix)
    for (int i = max(k_prev - 1, 0) + lane; i < k; i += 32) {
        const double b = sm.be[i];
        const double ibv = (b > 1e-290) ? 1.0 / b : 0.0;
        const double ccv = (i > 0) ? sm.be[i - 1] * ibv : 0.0;
        sm.ib[i] = ibv;
        sm.cc[i] = ccv;
        sm.alf[i] = (float)sm.al[i];
        sm.ibf[i] = (float)ibv;
        sm.ccf[i] = (float)ccv;
    }
    // Gershgorin interval of T_k
    double gl = 1e300, gu = -1e300;
    for (int i = lane; i < k; i += 32) {
        const double r = ((i > 0) ? fabs(sm.be[i - 1]) : 0.0) + ((i < k - 1) ? fabs(sm.be[i]) : 0.0);
        gl = fmin(gl, sm.al[i] - r);
        gu = fmax(gu, sm.al[i] + r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
        gu = fmax(gu, __shfl_xor_sync(0xffffffffu, gu, o));
    }
    __syncwarp();
    const double pad = 1e-12 * fmax(fabs(gl), fabs(gu)) + 1e-300;
    gl -= pad;
    gu += pad;
    gu_out = gu;
    const double W = gu - gl;
    double lo = gl, hi = gu;
    // three float32 rounds of 33-way multisection: the bracket shrinks to W / 33^3 = 2.8e-5 W
    for (int round = 0; round < 3; ++round) {
        const double step = (hi - lo) / 33.0;
        const bool ge = lc_any_below_f32(sm, k, (float)(lo + step * (double)(lane + 1)));
        const unsigned m = __ballot_sync(0xffffffffu, ge);
        const int f = m ? __ffs(m) - 1 : 32;
        hi = (f < 32) ? lo + step * (double)(f + 1) : hi;
        lo = lo + step * (double)f;
    }
    // float32 rounding moves a crossing by ~1e-6 W: widen by 2e-5 W, then float64 rounds over 32 points INCLUDING both ends, so a
    // bracket that does not hold the crossing is noticed and moved instead of trusted
    lo = fmax(lo - 2e-5 * W, gl);
    hi = fmin(hi + 2e-5 * W, gu);
    for (int round = 0, good = 0; round < 12 && good < 3; ++round) {
        const double step = (hi - lo) / 31.0;
        const bool ge = lc_any_below_f64(sm, k, lo + step * (double)lane);
        const unsigned m = __ballot_sync(0xffffffffu, ge);
        if (m & 1u) {  // an eigenvalue at or below lo: move the bracket down
            const double w = hi - lo;
            hi = lo;
            lo = fmax(lo - 16.0 * w, gl);
            if (!(hi > lo)) break;  // (lo == gl: cannot happen, gl is below the spectrum)
        } else if (m == 0u) {  // none below hi: move it up
            const double w = hi - lo;
            lo = hi;
            hi = fmin(hi + 16.0 * w, gu);
            if (!(hi > lo)) break;
        } else {
            const int f = __ffs(m) - 1;  // >= 1
            hi = lo + step * (double)f;
            lo = lo + step * (double)(f - 1);
            ++good;
        }
    }
    // Newton from the left end (below every root of Q_k: monotone, quadratic for a simple root); every lane the same arithmetic
    double x = lo;
    for (int itn = 0; itn < 4; ++itn) {
        double f, df;
        lc_newton_eval(sm, k, x, lane == 0, f, df);
        if (!(df < 0.0) || !(f > 0.0)) break;
        const double xn = fmin(x - f / df, hi);
        if (!(xn > x)) break;  // converged to rounding
        const bool tiny = (xn - x) < 1e-13 * W;
        x = xn;
        if (tiny && itn >= 1) break;  // the members parked are those of the previous iterate: closer than 1e-13 W
    }
    theta = x;
    __syncwarp();
    // Residual of the Ritz pair: be_{k-1} |s_{k-1}| / |s|, s the eigenvector of T_k.  The members Q_i(theta) ARE that eigenvector as
    // long as they grow; once theta has converged it is (to rounding) an eigenvalue of T_j for all later j as well, Q_j(theta) is
    // noise and everything the forward recurrence builds on it is the other, growing solution.  The recurrence run BACKWARDS from
    // the last row, B_{k-1} = 1, B_{i-1} = ((al_i - theta) B_i - be_i B_{i+1}) / be_{i-1}, is accurate exactly where the forward one
    // is not (it grows towards the top when the last components are small).  The two are joined where the forward members peak.
    {
        double bp = 0.0, b = 1.0;
        int sc = 0;
        for (int i = k - 1; i >= 1; --i) {
            if (lane == 0) {
                sm.bv[i] = b;
                sm.bve[i] = sc;
            }
            const double w = (sm.al[i] - theta) * sm.ib[i - 1];
            const double bn = fma(w, b, -(sm.be[i] * sm.ib[i - 1]) * bp);
            bp = b;
            b = bn;
            if ((i & 7) == 0) {
                if (fabs(b) > 1e60) {
                    b *= 0x1p-256, bp *= 0x1p-256;
                    sc -= 256;
                } else if (fabs(b) < 1e-60 && fabs(bp) < 1e-60) {
                    b *= 0x1p256, bp *= 0x1p256;
                    sc += 256;
                }
            }
        }
        if (lane == 0) {
            sm.bv[0] = b;
            sm.bve[0] = sc;
        }
    }
    __syncwarp();
    double lf[2], lb[2];  // log2 magnitudes of the forward / backward members i = lane, lane + 32
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        lf[h] = (i < k && sm.qv[i] != 0.0) ? log2(fabs(sm.qv[i])) - (double)sm.qe[i] : -1e300;
        lb[h] = (i < k && sm.bv[i] != 0.0) ? log2(fabs(sm.bv[i])) - (double)sm.bve[i] : -1e300;
    }
    // r = where the forward members peak (first index on ties), off = what brings the forward members onto the backward ones there
    double mf = fmax(lf[0], lf[1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mf = fmax(mf, __shfl_xor_sync(0xffffffffu, mf, o));
    const unsigned m0 = __ballot_sync(0xffffffffu, lf[0] == mf), m1 = __ballot_sync(0xffffffffu, lf[1] == mf);
    const int r = m0 ? __ffs(m0) - 1 : 32 + __ffs(m1) - 1;
    const double lbr = __shfl_sync(0xffffffffu, (r < 32) ? lb[0] : lb[1], r & 31);
    const double off = (lbr > -1e299) ? lbr - mf : 0.0;
    double st[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        st[h] = (i >= k) ? -1e300 : ((i <= r) ? ((lf[h] > -1e299) ? lf[h] + off : -1e300) : lb[h]);
    }
    double mx = fmax(st[0], st[1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double s2 = ((st[0] > -1e299) ? exp2(2.0 * (st[0] - mx)) : 0.0) + ((st[1] > -1e299) ? exp2(2.0 * (st[1] - mx)) : 0.0);
    s2 = warp_sum_d(s2);
    const double llast = __shfl_sync(0xffffffffu, ((k - 1) < 32) ? st[0] : st[1], (k - 1) & 31);
    const double bk = sm.be[k - 1];
    res2 = (llast > -1e299) ? bk * bk * exp2(2.0 * (llast - mx)) / s2 : 0.0;
}

__global__ void __launch_bounds__(LC_TT, 1) lanczos_cluster_kernel(const DenseArgs a) {
    COVO_DYN_SMEM(smraw);
    LcSmem& sm = *reinterpret_cast<LcSmem*>(smraw);
    const int n = a.n, tid = threadIdx.x, env = blockIdx.y, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)gjb_rank();
    const int k_max = min(kLanczosMax, n);
#if !defined(COVO_CPU_EMU)
    // programmatic dependent launch: the pole-inverse kernel may become resident now and load its copy of the matrix while this
    // recurrence runs; it waits (griddepcontrol.wait) for lambda_min before it shifts the diagonal
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    DENSE_STAMP(48);
    if (tid == 0) {
        gjb_mbar_init(&sm.bar[0], 1);
        gjb_mbar_init(&sm.bar[1], 1);
        sm.progress = sm.final_k = sm.verdict_k = sm.converged_k = 0;
#if !defined(COVO_CPU_EMU)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    if (warp == LC_T / 32) {
        // ================================================ the checker ================================================
        // (one cluster barrier per role, none of the CTA-wide kind: a __syncthreads() in each branch of the role split is what the
        // hardware executes happily and what synccheck reports as a barrier in divergent code)
        gjb_cluster_sync();
        int c = min(kLanczosFirstCheck, k_max), c_prev = 0;
        for (;;) {
            int fk = 0;
            for (;;) {  // T_c complete, or the recurrence over before it got there (lane 0 looks, so the warp cannot split)
                int go = 0;
                if (lane == 0) {
                    fk = lc_load(&sm.final_k);
                    go = fk || lc_load(&sm.progress) >= c;
                }
                go = __shfl_sync(0xffffffffu, go, 0);
                if (go) break;
                lc_pause();
            }
            fk = __shfl_sync(0xffffffffu, fk, 0);
            if (fk && fk < c) c = fk;
            lc_fence();
            __syncwarp();
            double theta, res2, gu;
            lc_judge(sm, c, c_prev, lane, theta, res2, gu);
            c_prev = c;
#if defined(COVO_CPU_EMU)
            if (lane == 0 && rank == 0 && getenv("COVO_EMU_TRACE")) fprintf(stderr, "checker: k %d theta %.12f res2 %.3e gu %.3f\n", c, theta, res2, gu);
#endif
            const bool conv = res2 <= kRitzTol2;
            fk = __shfl_sync(0xffffffffu, (lane == 0) ? lc_load(&sm.final_k) : 0, 0);
            const bool last = conv || c >= k_max || (fk && c >= fk);
            if (lane == 0) {
                if (conv) lc_store(&sm.converged_k, c);
                lc_fence();
                lc_store(&sm.verdict_k, c);
                lc_published();
                if (last && rank == 0) {
                    // not converged (status 3): theta - residual is a lower bound of lambda_min for a Ritz pair that belongs to the
                    // lowest eigenvalue, so A = R - lambda_min + 1e-2 stays positive definite; the spectrum's upper end moves with it
                    const double r = conv ? 0.0 : sqrt(res2);
                    a.scal[(long long)env * 4 + 0] = theta - r;
                    a.scal[(long long)env * 4 + 1] = gu;  // upper bound of the spectrum of T (selects the approximation interval only)
                    a.scal[(long long)env * 4 + 3] = (double)c;
                    if (!conv) a.status[env] = 3;
                }
            }
            if (last) break;
            c = min(c + kLanczosCheckEvery, k_max);
        }
    } else {
        // ================================================ the recurrence ================================================
        const int R = (n + LC_CL - 1) / LC_CL;      // rows per CTA (25 at n = 200)
        const int rl = tid / LC_TPR, ph = tid % LC_TPR;  // local row, column phase
        const int row = rank * R + rl;
        const bool has_row = rl < R && row < n;
        const bool leader = has_row && ph == 0;
        const int kc = (n + LC_TPR - 1) / LC_TPR;   // columns per thread
        const float* Rg = a.R + (long long)env * n * n;
        float* Asym = a.Asym ? a.Asym + (long long)env * n * n : nullptr;
        // rows of (R + R^T)/2 (float32, controllers/covo.py:117) widened into registers
        double ar[LC_KMAX];
        {
            float xa[LC_KMAX], xb[LC_KMAX];  // all loads in flight before the first store (the compiler must assume Asym aliases R)
#pragma unroll
            for (int k = 0; k < LC_KMAX; ++k) {
                const int j = LC_TPR * k + ph;
                const bool ok = has_row && j < n;
                xa[k] = ok ? __ldg(Rg + (long long)row * n + j) : 0.f;
                xb[k] = ok ? __ldg(Rg + (long long)j * n + row) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < LC_KMAX; ++k) {
                const int j = LC_TPR * k + ph;
                const float x = 0.5f * (xa[k] + xb[k]);
                if (Asym && has_row && j < n) Asym[(long long)row * n + j] = x;
                ar[k] = (double)x;
            }
        }
        // start vector: every CTA builds and normalises all of it (same arithmetic everywhere)
        {
            double x[7], s2 = 0.0;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const int j = lane + 32 * c;
                // fixed start vector with a component along every eigenvector in practice: 1 + a 16-bit multiplicative hash of j (no
                // float64 transcendental: cos() alone cost 4 us here)
                x[c] = (j < n) ? 1.0 + (double)(float)((((unsigned)j + 1u) * 2654435761u >> 8) & 0xffffu) * (1.0 / 65536.0) : 0.0;
                s2 = fma(x[c], x[c], s2);
            }
            const double inv = 1.0 / sqrt(warp_sum_d(s2));
            if (warp == 0) {
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    sm.v[0][lane + 32 * c] = x[c] * inv;
                    sm.v[1][lane + 32 * c] = 0.0;
                }
            }
        }
        gjb_cluster_sync();  // the start vector is in place here, the mbarriers are initialised everywhere: before the first read / send
        double vj = has_row ? sm.v[0][row] : 0.0, vprev = 0.0;  // row leaders keep v_k[row], v_{k-1}[row]
        const int n_warps_total = LC_CL * LC_WPC;
        const int tx_bytes = n * 8 + n_warps_total * 16;
        DENSE_STAMP(49);
        int kdone = 0;
        double beta_prev = 0.0;
        LcRemote rem[LC_CL];  // where this thread's row entry / this warp's partial slot / the barrier live in every CTA of the cluster
#pragma unroll
        for (int r = 0; r < LC_CL; ++r)
            rem[r] = lc_remote(&sm.u[0][has_row ? row : 0], &sm.part[0][rank * LC_WPC + warp][0], &sm.bar[0], (unsigned)r);
        const bool pfl = a.prof && tid == 0 && rank == 0 && blockIdx.y == 0;
        long long pl[5] = {0, 0, 0, 0, 0}, tl0 = pfl ? clock64() : 0;  // phase clocks of the recurrence (registers; dumped once)
        for (int it = 0; it < k_max; ++it) {
            const int pc = it & 1;  // v_k is in v[pc]; this step's exchange uses u[pc], part[pc], bar[pc]
            // y_row = (A v_k)_row: four threads per row, eight independent chains each
            // (dependent float64 operations are ~40 cycles apart on this pipe: eight chains of seven, not four of thirteen)
            double ac[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            const double* vv = sm.v[pc];
#pragma unroll
            for (int k = 0; k < LC_KMAX; k += 8) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (k + e < LC_KMAX && k + e < kc) ac[e] = fma(ar[(k + e < LC_KMAX) ? k + e : 0], vv[LC_TPR * (k + e) + ph], ac[e]);
            }
            double y = ((ac[0] + ac[1]) + (ac[2] + ac[3])) + ((ac[4] + ac[5]) + (ac[6] + ac[7]));
#pragma unroll
            for (int o = 1; o < LC_TPR; o <<= 1) y += __shfl_xor_sync(0xffffffffu, y, o);
            // u = y - beta_{k-1} v_{k-1} (row leaders); partials of u.v_k and u.u over the 8 rows of the warp
            const double u = leader ? y - beta_prev * vprev : 0.0;
            double pa = u * vj, pu = u * u;
#pragma unroll
            for (int o = LC_TPR; o < 32; o <<= 1) {
                pa += __shfl_xor_sync(0xffffffffu, pa, o);
                pu += __shfl_xor_sync(0xffffffffu, pu, o);
            }
            if (tid == 0) gjb_mbar_expect(&sm.bar[pc], tx_bytes);
            long long tq0 = 0;
            if (a.prof && tid == 0 && rank == 0 && blockIdx.y == 0) tq0 = clock64();
            if (pfl) {
                pl[0] += tq0 - tl0;  // matrix-vector product, reductions
                tl0 = tq0;
            }
            // the exchange: u of this row and the warp's partials to every CTA of the cluster (this one included)
#pragma unroll
            for (int r = 0; r < LC_CL; ++r) {
                if (leader) lc_send(rem[r], 0, pc, u);
                if (lane == 0) {
                    lc_send(rem[r], 1, pc, pa);
                    lc_send(rem[r], 2, pc, pu);
                }
            }
            // In the shadow of the exchange: 1 / |v_k|^2, every warp for itself (all of v_k is local; same arithmetic everywhere).
            // v_k is normalised with the COMPUTED beta, so |v_k|^2 = 1 + delta; "beta^2 = u.u - alpha^2" takes delta for zero and hands
            // delta alpha^2 / beta^2 on to the next vector: (alpha / beta)^2 is ~30 for every ghost of lambda_max, and after ~55 steps
            // the recurrence had lost its normalisation (Ritz values outside the spectrum).  With the measured norm the formulas below
            // are exact algebra for any delta.
            double iqv;
            {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int c = 0; c < 6; c += 2) {
                    const double x0 = vv[lane + 32 * c], x1 = vv[lane + 32 * (c + 1)];
                    s0 = fma(x0, x0, s0);
                    s1 = fma(x1, x1, s1);
                }
                const double x6 = vv[lane + 32 * 6];
                s0 = fma(x6, x6, s0);
                const double qv = warp_sum_d(s0 + s1);
                double r = 2.0 - qv;  // 1 / qv for qv = 1 + delta: two Newton steps from the first-order guess
                r = r * (2.0 - qv * r);
                iqv = r * (2.0 - qv * r);
            }
            if (pfl) {
                const long long t1 = clock64();
                pl[1] += t1 - tl0;  // sends + |v|^2
                tl0 = t1;
            }
            gjb_mbar_wait(&sm.bar[pc], (unsigned)((it >> 1) & 1));
            if (pfl) {
                const long long t1 = clock64();
                pl[2] += t1 - tl0;  // waiting for the exchange
                tl0 = t1;
            }
            if (a.prof && tid == 0 && rank == 0 && blockIdx.y == 0) a.prof[52] = (it == 0 ? 0 : a.prof[52]) + (clock64() - tq0);  // send + wait
            // alpha = u.v_k / |v_k|^2, beta_k^2 = |u - alpha v_k|^2 = u.u - alpha (u.v_k): every warp sums the 32 partial pairs with the
            // same butterfly
            double qa = (lane < n_warps_total) ? sm.part[pc][lane][0] : 0.0;
            double qu = (lane < n_warps_total) ? sm.part[pc][lane][1] : 0.0;
            if (LC_CL * LC_WPC > 32) {  // (eight threads per row: 64 warps in the cluster, two partial pairs per lane)
                qa += sm.part[pc][(lane + 32) % (LC_CL * LC_WPC)][0];
                qu += sm.part[pc][(lane + 32) % (LC_CL * LC_WPC)][1];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {  // the two sums interleaved: five levels of (shuffle + add) each
                qa += __shfl_xor_sync(0xffffffffu, qa, o);
                qu += __shfl_xor_sync(0xffffffffu, qu, o);
            }
            const double alpha = qa * iqv;
            double b2 = qu - alpha * qa;
            if (b2 < 1e-3 * qu) {
                // |u|^2 - alpha^2 cancels when beta << |alpha| (Krylov space nearly exhausted, small n): form |u - alpha v_k|^2 directly.
                // Every CTA holds all of u and v_k, so this needs no exchange; the branch is uniform across the cluster.
                double ps = 0.0;
                for (int j = tid; j < n; j += LC_T) {
                    const double wj = sm.u[pc][j] - alpha * vv[j];
                    ps = fma(wj, wj, ps);
                }
                ps = warp_sum_d(ps);
                if (lane == 0) sm.red[warp] = ps;
                lc_main_sync();
                b2 = 0.0;
#pragma unroll
                for (int w = 0; w < LC_WPC; ++w) b2 += sm.red[w];
                lc_main_sync();
            }
            // 1 / beta and beta without float64 sqrt / division (each a chain of ~15 dependent float64 operations at ~40 cycles): float32
            // rsqrt seed, two Newton steps (1e-7 -> 1e-14 -> rounding), six dependent operations
            double beta_new = 0.0, ib = 0.0;
            if (b2 > 1e-280) {
                const double sc = (b2 < 1e-30) ? 1e60 : 1.0;  // keep the seed inside the float32 range
                const double bs = b2 * sc;
                double r = (double)rsqrtf((float)bs);
                r = r * fma(-0.5 * bs, r * r, 1.5);
                r = r * fma(-0.5 * bs, r * r, 1.5);
                ib = r * ((b2 < 1e-30) ? 1e30 : 1.0);
                beta_new = b2 * ib;
            }
            if (pfl) {
                const long long t1 = clock64();
                pl[3] += t1 - tl0;  // alpha, beta
                tl0 = t1;
            }
            if (tid == 0) {  // hand T's new row to the checker
                sm.al[it] = alpha;
                sm.be[it] = beta_new;
                lc_fence();
                lc_store(&sm.progress, it + 1);
                lc_published();
            }
            kdone = it + 1;
            if (!(beta_new > 1e-200)) break;  // invariant subspace (uniform across the cluster: every CTA sees the same data)
            // v_{k+1} = (u - alpha v_k) / beta_k: every CTA forms all of it (two entries per thread)
            for (int j = tid; j < n; j += LC_T) sm.v[pc ^ 1][j] = (sm.u[pc][j] - alpha * vv[j]) * ib;
            if (leader) {
                const double vn = (u - alpha * vj) * ib;
                vprev = vj;
                vj = vn;
            }
            beta_prev = beta_new;
            // the checker's verdict on T_c, c = kdone - lag (the same numbers in every CTA: the whole cluster stops here or nowhere)
            const int c = kdone - kLanczosLag;
            bool stop = false;
            if (c >= kLanczosFirstCheck && (c - kLanczosFirstCheck) % kLanczosCheckEvery == 0) {
                if (lane == 0)
                    while (lc_load(&sm.verdict_k) < c) lc_pause();
                __syncwarp();
                lc_fence();
                const int ck = lc_load(&sm.converged_k);
                stop = ck != 0 && ck <= c;
            }
            lc_main_sync();
            if (pfl) {
                const long long t1 = clock64();
                pl[4] += t1 - tl0;  // next vector, verdict, barrier
                tl0 = t1;
            }
            if (stop) break;
        }
        if (pfl) {
#pragma unroll
            for (int q = 0; q < 5; ++q) a.prof[40 + q] = pl[q];
            a.prof[45] = kdone;
        }
        if (tid == 0) {
            lc_fence();
            lc_store(&sm.final_k, kdone);
            lc_published();
        }
        DENSE_STAMP(50);
    }
    DENSE_STAMP(51);
    gjb_cluster_sync();  // nobody leaves while a peer could still be sending to it (and the checker has written lambda_min)
}

// ---------------------------------------------------------------------------------------------------------------------------
// D2 gjb_inverse_kernel: in-place Gauss-Jordan sweep without pivoting (A + t_j I is SPD), BLOCKED (8 pivots per step), in FLOAT64, spread
// over a cluster of GB_CL (8) CTAs per pole, the matrix resident in registers.  Sweeping the index block K with P = A_KK:
//     A_IJ -= A_IK P^-1 A_KJ,   A_KJ <- P^-1 A_KJ =: G,   A_IK <- -A_IK P^-1,   A_KK <- P^-1.
// With this sign convention the matrix is symmetric on the unswept index set and ANTI-symmetric between swept and unswept
// indices, so the column panel is the row panel again: A_iK = sigma(i) (A_Ki)^T, sigma = -1 for swept i.  Everything a step needs
// therefore follows from the 8 raw pivot rows (8 x n) alone; with the multipliers MP[i][u] = -sigma(i) G[u][i] (P^-1 is symmetric):
//     A_ij += sum_u MP[i][u] raw[u][j],    row K_s <- G[s][:],   column K_s <- MP[i][s],   block KK <- P^-1,
// so that G is only needed at the columns of the CTA's own rows (the multipliers) and, for the owner of the pivot rows, in those rows.
// Why float64 (round 2): A + t_j I has condition up to 1.6e5 for the smallest poles and ANY float32 factorisation has a backward error
// of eps |A| ~ 3e-5, 0.3 % of the eigenvalue 1e-2 that carries the largest direction of Sigma: the float32 version of this kernel left
// Sigma 1e-4 .. 5e-3 from exact arithmetic where the reference's float32 eigh pipeline is at 1e-5 (tools/studies/gj_accuracy*.py).
// In float64 the inverses are exact to ~1e-11 and the accuracy of the path is that of the rational approximation: 13 poles, 2e-7
// (tools/studies/pole_count.py); 13 + 1 clusters of 8 CTAs = 112 of the 148 SMs, all resident at once.  Measured on B200
// (tools/microbench/fp64_rates.cu): DFMA issues at 64 lanes / clock / SM with 9 cycles between dependent operations (DMMA m8n8k4: the
// same 64 FMA / clock / SM), so the arithmetic is cheap; what a step costs is shared-memory traffic and the dependent chain
// "look-ahead rows -> DSMEM -> 8 x 8 inverse -> multipliers -> step barrier".
// Layout: ROW-CYCLIC over the cluster -- row i lives in CTA i mod 8 -- so every 8-row pivot block is ONE row per CTA: each CTA sends
// one row (1.8 KB) to its seven peers per step.  (Tiles of 16 rows per CTA, the first float64 version, made one CTA send all eight rows:
// 115 KB through one SM's DSMEM port, measured 2.2 us per step; tools/microbench/dsmem_latency.cu: a 1792-byte copy alone is 735 cycles.)
// Roles (384 threads per CTA):
//   * 8 UPDATE warps hold the matrix: thread (ty = warp, tx = lane) owns the local rows ty + 8 k, k < 4 (global row rank + 8 (ty + 8 k);
//     28 rows per CTA, 25 of them real at n = 200) and columns tx + 32 b, b < 7: 28 doubles.  Per step and thread: up to 224 DFMA
//     against 56 + 16 shared-memory loads.
//   * 4 SOLVER warps run one block ahead: warp 0 inverts the 8 x 8 pivot block of block m + 1, whose rows' owners send their eight
//     pivot-column entries AHEAD of the rows (64 bytes each, an mbarrier of their own), lane = two entries, 8 shuffle-driven pivots; when
//     the rows themselves have arrived (bulk copies through distributed shared memory, completion counted by an mbarrier -- no fences,
//     no cluster barrier) all four build the multiplier table of this CTA's rows.  All of it happens while the update warps are
//     still applying block m.
//   * look-ahead: after the barrier that opens step m, the warp that owns the CTA's row of block m + 1 applies step m to that row
//     first -- the eight pivot columns before the rest -- and publishes it.
// Flow control: the raw panels live in a ring of 4 slots; a split cluster barrier (arrive after a step is opened, wait before the
// next one) keeps everybody within one step of everybody else, so 4 slots never collide.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int GB_CL = 8;     // CTAs per matrix
constexpr int GB_NR = 4;     // rows per update thread
constexpr int GB_UT = 256;   // update threads
constexpr int GB_ST = 128;   // solver threads
constexpr int GB_T = GB_UT + GB_ST;
constexpr int GB_NP = 224;   // padded order: 28 local rows in each of the 8 CTAs, 7 column slots of 32
constexpr int GB_LR = GB_NP / GB_CL;  // local rows per CTA
constexpr int GB_SLOTS = 4;

template <int V>
struct IntC {
    static constexpr int value = V;
};

// NB = pivots per step (8, the default, or 16): a block is NB / 8 rows per CTA.  16 halves the number of steps and with it the per-step
// fixed cost (DSMEM hop, barrier); the NB x NB inverse stays one warp's dependent chain of NB pivots -- and that is what makes 16 the
// slower choice (launch_sigma_dense).
template <int NB>
struct GjbSmemT {
    static constexpr int RB = NB / GB_CL;  // rows of a block per CTA
    double raw[GB_SLOTS][NB][GB_NP];  // pivot-row panels [s][j]: row s comes from CTA s % 8
    double stage[2][RB][GB_NP];       // this CTA's pivot rows on their way out (by block parity): source of the bulk copies
    double MP[2][8][NB][GB_NR];       // [parity][ty][u][k]: -sigma(i) G[u][i] for the local row ty + 8 k; 0 for pivot rows and padding
    double Pblk[GB_SLOTS][NB][NB];    // the NB x NB pivot blocks, sent ahead of the rows
    double Pinv[2][NB * NB];
    double piv[GB_NP];
    unsigned long long rawbar[GB_SLOTS];
    unsigned long long pbar[GB_SLOTS];
    long long pacc[16];  // phase-clock accumulators (slots 48 + i), dumped once at the end: a global read-modify-write per stamp distorts what it measures
    int bad;
};

#if defined(COVO_CPU_EMU)
__device__ __forceinline__ double gjb_rcp64(double x) { return 1.0 / x; }
__device__ __forceinline__ double gjb_sel(bool c, double x, double y) { return c ? x : y; }
#else
// c ? x : y as an opaque select: written as `(b == idx) ? arr[b] : v` over an unrolled loop, the compiler turns the register array into
// a local-memory array with a dynamic index
__device__ __forceinline__ double gjb_sel(bool c, double x, double y) {
    double r;
    asm("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\nselp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(x), "d"(y), "r"((int)c));
    return r;
}
// 1 / x for a pivot (1e-3 .. 1e5): float32 seed, two Newton steps in float64 (4 dependent DFMA instead of a division)
__device__ __forceinline__ double gjb_rcp64(double x) {
    double r = (double)gjb_rcp((float)x);
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
#endif

template <int NB>
__global__ void __launch_bounds__(GB_T, 1) gjb_inverse_kernel_t(const DenseArgs a) {
    using Smem = GjbSmemT<NB>;
    constexpr int RB = NB / GB_CL;      // rows of a block per CTA
    constexpr int LPR = 32 / NB;        // lanes per row of the pivot block in the inverting warp (4 or 2)
    constexpr int EPL = NB / LPR;       // entries per lane (2 or 8)
    COVO_DYN_SMEM(smraw);
    Smem& sm = *reinterpret_cast<Smem*>(smraw);
    const int n = a.n, tid = threadIdx.x, env = blockIdx.y, pole = blockIdx.x / GB_CL;
    const int rank = (int)gjb_rank();
    const int nblk = (n + NB - 1) / NB;
    if (tid < 16) sm.pacc[tid] = 0;
    if (tid == 0) {
        sm.bad = 0;
        for (int q = 0; q < GB_SLOTS; ++q) {
            gjb_mbar_init(&sm.rawbar[q], 1);
            gjb_mbar_init(&sm.pbar[q], 1);
        }
#if !defined(COVO_CPU_EMU)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    for (int i = tid; i < GB_SLOTS * NB * GB_NP; i += GB_T) (&sm.raw[0][0][0])[i] = 0.0;
    gjb_cluster_sync();  // barriers initialised and panels zeroed everywhere before anybody publishes

    if (tid >= GB_UT) {
        // ================================================ solver warps ================================================
        const int sidx = tid - GB_UT, lane = sidx & 31, swarp = sidx >> 5;
#if !defined(COVO_CPU_EMU)
        asm volatile("griddepcontrol.wait;" ::: "memory");  // (the update warps wait where they need lambda_min; nothing here reads it)
#endif
        for (int m = 0; m < nblk; ++m) {
            const int slot = m & (GB_SLOTS - 1), par = m & 1;
            const unsigned ring_par = (unsigned)((m / GB_SLOTS) & 1);
            if (sidx == 0) {
                gjb_mbar_expect(&sm.pbar[slot], NB * NB * 8);
                gjb_mbar_expect(&sm.rawbar[slot], NB * GB_NP * 8);
            }
            const bool pf = a.prof && sidx == 0 && blockIdx.x == 0 && blockIdx.y == 0;
            long long t0 = pf ? clock64() : 0;
            if (swarp == 0) {
                gjb_mbar_wait(&sm.pbar[slot], ring_par);
                if (pf) {
                    const long long t1 = clock64();
                    sm.pacc[6] += (t1 - t0);
                    t0 = t1;
                }
                // P^-1 by an in-place Gauss-Jordan sweep of the NB x NB block: lane = (row r, columns EPL h .. EPL h + EPL - 1).  The
                // reciprocal of the NEXT pivot is started as soon as this pivot's is known -- P'[s+1][s+1] = P[s+1][s+1] -
                // (P[s+1][s] / p) P[s][s+1] is the very operation the sweep applies to that entry -- so the dependent chain per pivot is
                // multiply, FMA, reciprocal instead of shuffle, reciprocal, multiply, FMA, shuffle.
                const int r = lane / LPR, h = lane % LPR;
                double x[EPL];
#pragma unroll
                for (int j = 0; j < EPL; ++j) x[j] = sm.Pblk[slot][r][EPL * h + j];
                double pvs[NB];  // the scalar pivots (log det, positivity): written out after the chain, not inside it
                pvs[0] = __shfl_sync(0xffffffffu, x[0], 0);
                double rinv = gjb_rcp64(pvs[0]);
#pragma unroll
                for (int sp = 0; sp < NB; ++sp) {
                    const int hs = sp / EPL, js = sp % EPL;  // compile-time after unrolling: column sp is entry js of the lanes with h = hs
                    const double prs = __shfl_sync(0xffffffffu, x[js], r * LPR + hs);  // P[r][sp]
                    double ps[EPL];
#pragma unroll
                    for (int j = 0; j < EPL; ++j) ps[j] = __shfl_sync(0xffffffffu, x[j], sp * LPR + h);  // P[sp][EPL h + j]
                    double rinv_next = 0.0;
                    if (sp + 1 < NB) {
                        const int hs1 = (sp + 1) / EPL, js1 = (sp + 1) % EPL;
                        const double pa = __shfl_sync(0xffffffffu, x[js], (sp + 1) * LPR + hs);    // P[sp + 1][sp]
                        const double pb = __shfl_sync(0xffffffffu, x[js1], sp * LPR + hs1);        // P[sp][sp + 1]
                        const double pc = __shfl_sync(0xffffffffu, x[js1], (sp + 1) * LPR + hs1);  // P[sp + 1][sp + 1]
                        pvs[(sp + 1 < NB) ? sp + 1 : 0] = fma(-(pa * rinv), pb, pc);
                        rinv_next = gjb_rcp64(pvs[(sp + 1 < NB) ? sp + 1 : 0]);
                    }
                    if (r == sp) {
#pragma unroll
                        for (int j = 0; j < EPL; ++j) x[j] = (EPL * h + j == sp) ? rinv : ps[j] * rinv;
                    } else {
                        const double f = prs * rinv;
#pragma unroll
                        for (int j = 0; j < EPL; ++j) x[j] = (EPL * h + j == sp) ? -f : fma(-f, ps[j], x[j]);
                    }
                    rinv = rinv_next;
                }
#pragma unroll
                for (int j = 0; j < EPL; ++j) sm.Pinv[par][r * NB + EPL * h + j] = x[j];
                if (lane < NB) {
                    double pl = pvs[0];
#pragma unroll
                    for (int sp = 1; sp < NB; ++sp) pl = (lane == sp) ? pvs[sp] : pl;
                    sm.piv[NB * m + lane] = pl;
                    if (!(pl > 0.0)) sm.bad = 1;
                }
                if (pf) {
                    const long long t1 = clock64();
                    sm.pacc[7] += (t1 - t0);
                    t0 = t1;
                }
            }
            gjb_mbar_wait(&sm.rawbar[slot], ring_par);  // every solver thread: it reads the rows below
            COVO_NAMED_BARRIER(1, GB_ST);               // P^-1 is in place
            if (pf) {
                const long long t1 = clock64();
                sm.pacc[12] += (t1 - t0);
                t0 = t1;
            }
            {
                // multipliers of this CTA's rows: MP[ty][u][k] = -sigma(i) sum_v P^-1[u][v] raw[v][i]; thread = (ty, k, NB / 4 of the u)
                constexpr int UPT = NB / 4;  // outputs per thread
                const double(*rw)[GB_NP] = sm.raw[slot];
                const int ty = sidx >> 4, k = (sidx >> 2) & 3, u0 = (sidx & 3) * UPT;
                const int l = ty + 8 * k, i = rank + GB_CL * l;
                const bool live = l < GB_LR && !(l >= RB * m && l < RB * m + RB);
                double col[NB];
#pragma unroll
                for (int v = 0; v < NB; ++v) col[v] = live ? rw[v][i] : 0.0;
                const double sgn = (l < RB * m) ? 1.0 : -1.0;
#pragma unroll
                for (int q = 0; q < UPT; ++q) {
                    const double2* pv = reinterpret_cast<const double2*>(&sm.Pinv[par][(u0 + q) * NB]);  // row u0 + q of P^-1
                    double ga = 0.0, gb = 0.0;
#pragma unroll
                    for (int v = 0; v < NB / 2; ++v) {
                        const double2 pu = pv[v];
                        ga = fma(pu.x, col[2 * v], ga);
                        gb = fma(pu.y, col[2 * v + 1], gb);
                    }
                    sm.MP[par][ty][u0 + q][k] = live ? sgn * (ga + gb) : 0.0;
                }
            }
            if (pf) {
                const long long t1 = clock64();
                sm.pacc[8] += (t1 - t0);
                t0 = t1;
            }
            if (m > 0) gjb_cluster_wait();
            COVO_NAMED_BARRIER(2, GB_T);  // opens step m for the update warps (named: the two roles arrive from different code)
            gjb_cluster_arrive();
            if (pf) sm.pacc[9] += (clock64() - t0);
        }
    } else {
        // ================================================ update warps ================================================
        const int tx = tid & 31, ty = tid >> 5;
        const float* Rg = a.R + (long long)env * n * n;
        double acc[GB_NR][7];
#pragma unroll
        for (int k = 0; k < GB_NR; ++k) {
            const int l = ty + 8 * k, i = rank + GB_CL * l;
#pragma unroll
            for (int b = 0; b < 7; ++b) {
                const int j = tx + 32 * b;
                double v = (i == j && l < GB_LR) ? 1.0 : 0.0;  // identity padding: never coupled, pivots 1
                if (l < GB_LR && i < n && j < n) {
                    // position (i, j) holds element (n - 1 - i, n - 1 - j): the sweep eliminates the LAST controls first (the order
                    // that was the most accurate one in float32, tools/studies/gj_accuracy.py; kept: the combine kernel and the
                    // tests know the triangle it produces)
                    const int ir = n - 1 - i, jr = n - 1 - j;
                    // (R + R^T)/2 in float32, controllers/covo.py:117 -- formed here, not read from the Lanczos kernel: it is still running
                    v = (double)(0.5f * (__ldg(Rg + (long long)ir * n + jr) + __ldg(Rg + (long long)jr * n + ir)));
                }
                acc[k][b] = v;
            }
        }
#if !defined(COVO_CPU_EMU)
        asm volatile("griddepcontrol.wait;" ::: "memory");  // lambda_min (the Lanczos kernel has completed and flushed)
#endif
        const double lam_min = a.scal[(long long)env * 4 + 0], lam_max = a.scal[(long long)env * 4 + 1];
        int lad = 0;
        {
            const double Mb = 1.02 * (lam_max - lam_min) + kOffset;
            double Mi = kOffset * (1.0 - 1e-7) * 256.0;
            while (lad < kZoloLadder - 1 && Mi < Mb) {
                Mi *= 4.0;
                ++lad;
            }
            if (Mi < Mb && tid == 0 && rank == 0) a.status[env] = 1;
        }
        const double* zt = a.zolo + (size_t)lad * 2 * kDensePoles;
        const bool want_logdet = pole == kDensePoles;
        const double shift = (kOffset - lam_min) + (want_logdet ? 0.0 : zt[pole]);
        const double wj = want_logdet ? 0.0 : zt[kDensePoles + pole];
        // the shift of this pole on the diagonal: position (i, i) sits in column slot i >> 5, lane i & 31
#pragma unroll
        for (int k = 0; k < GB_NR; ++k) {
            const int l = ty + 8 * k, i = rank + GB_CL * l;
#pragma unroll
            for (int b = 0; b < 7; ++b)
                if (l < GB_LR && i < n && i == tx + 32 * b) acc[k][b] += shift;
        }
        const bool has_row3 = rank + GB_CL * (ty + 24) < n;  // warp-uniform: the fourth row of this warp is a row of the matrix (not padding)
        // A row of a block leaves as ONE bulk copy per destination CTA: staged in shared memory, fenced for the async proxy; its NB
        // entries in the block's own pivot columns travel ahead as st.async stores.  s_row = its row index inside the block.
        auto publish_pblock = [&](int blk, int s_row, double v) {  // called by all lanes of one warp; the lanes holding the block's columns send their entry
            const int K1 = NB * blk;
            if ((tx & ~(NB - 1)) == (K1 & 31)) {
                // straight from the register into every CTA's copy of the block: st.async needs no staging and no proxy fence (the
                // staged 64-byte bulk copy took 0.38 us from here to the copy unit)
                const int slot = blk & (GB_SLOTS - 1);
#pragma unroll
                for (int r = 0; r < GB_CL; ++r) gjb_send64(&sm.Pblk[slot][s_row][tx & (NB - 1)], (unsigned)r, v, &sm.pbar[slot]);
            }
        };
        auto publish_row = [&](int blk, int s_row, const double (&vals)[7], int Kfix, double vfix) {
            const int slot = blk & (GB_SLOTS - 1);
            double* st = sm.stage[blk & 1][s_row / GB_CL];
#pragma unroll
            for (int b = 0; b < 7; ++b) st[tx + 32 * b] = vals[b];
            if (Kfix >= 0 && (tx & ~(NB - 1)) == (Kfix & 31)) st[(Kfix & ~31) + tx] = vfix;  // same thread, same address: program order
#if defined(COVO_CPU_EMU)
            __syncwarp();
            if (tx < GB_CL) gjb_bulk_send(&sm.raw[slot][s_row][0], st, GB_NP * 8, (unsigned)tx, &sm.rawbar[slot]);
#else
            // this CTA's own copy of the row: ordinary stores + a local complete_tx (a bulk copy to the CTA's own shared::cluster address works
            // on the hardware but compute-sanitizer's memcheck rejects it as "not located in remote CTA", tools/microbench/dsmem_latency.cu)
            double* own = &sm.raw[slot][s_row][0];
#pragma unroll
            for (int b = 0; b < 7; ++b) own[tx + 32 * b] = vals[b];
            if (Kfix >= 0 && (tx & ~(NB - 1)) == (Kfix & 31)) own[(Kfix & ~31) + tx] = vfix;
            gjb_fence_async_proxy();
            __syncwarp();
            if (tx < GB_CL) {
                if (tx != rank) gjb_bulk_send(&sm.raw[slot][s_row][0], st, GB_NP * 8, (unsigned)tx, &sm.rawbar[slot]);
                else gjb_mbar_complete_tx_local(&sm.rawbar[slot], GB_NP * 8);
            }
#endif
        };
        // block 0: local rows 0 .. RB - 1 of every CTA (warps 0 .. RB - 1, k = 0)
        if (ty < RB) {
            publish_pblock(0, rank + GB_CL * ty, acc[0][0]);
            publish_row(0, rank + GB_CL * ty, acc[0], -1, 0.0);
        }
        const bool pfu = a.prof && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
        long long tu0 = pfu ? clock64() : 0;
        for (int m = 0; m < nblk; ++m) {
            if (m > 0) gjb_cluster_wait();  // every CTA of the cluster has opened step m - 1 (see "flow control")
            COVO_NAMED_BARRIER(2, GB_T);  // multipliers and P^-1 of block m are in place; everybody is done with step m - 1
            gjb_cluster_arrive();
            if (pfu) {
                const long long t1 = clock64();
                sm.pacc[10] += (t1 - tu0);
                tu0 = t1;
            }
            const int par = m & 1, K0 = NB * m, slot = m & (GB_SLOTS - 1);
            gjb_mbar_wait(&sm.rawbar[slot], (unsigned)((m / GB_SLOTS) & 1));  // complete long ago; makes the async-proxy writes visible HERE
            const double(*rw)[GB_NP] = sm.raw[slot];
            const double(*mp)[GB_NR] = sm.MP[par][ty];  // [u][k]
            // ---- look-ahead: a warp that owns one of this CTA's rows of block m + 1 (local rows RB (m + 1) + q) updates it first and
            // publishes it (straight-line code on a copy of the row: with the row / column slot chosen by predicates inside the FMA
            // loops ptxas serialised every shared-memory load with its FMA, 60 cycles per FMA)
            int k_done = -1;
            const int q1 = (ty - RB * (m + 1)) & 7;  // which of the block's rows in this CTA this warp would own
            if (m + 1 < nblk && q1 < RB) {
                const int l1 = RB * (m + 1) + q1, k1 = l1 >> 3, K1 = K0 + NB, s_row = rank + GB_CL * q1;
                k_done = k1;
                double row[7];
#pragma unroll
                for (int b = 0; b < 7; ++b) {
                    row[b] = acc[0][b];
#pragma unroll
                    for (int k = 1; k < GB_NR; ++k) row[b] = (k == k1) ? acc[k][b] : row[b];
                }
                const bool pfo = a.prof && tx == 0 && blockIdx.x == 0 && blockIdx.y == 0;
                long long to0 = pfo ? clock64() : 0;
                {
                    // the block's own pivot columns first (lanes K1 & 31 .. + NB - 1 of column slot K1 >> 5): they leave ahead of the row.
                    // Same operations in the same order as the row update below, so the two agree bit for bit.
                    const int jp = (K1 & ~31) + tx;
                    double pe = row[0];
#pragma unroll
                    for (int b = 1; b < 7; ++b) pe = gjb_sel(b == (K1 >> 5), row[b], pe);
#pragma unroll
                    for (int u = 0; u < NB; ++u) pe = fma(mp[u][k1], rw[u][jp], pe);
                    if (pfo) {
                        const long long t1 = clock64();
                        sm.pacc[14] += (t1 - to0);  // pivot columns of the row
                        to0 = t1;
                    }
                    publish_pblock(m + 1, s_row, pe);
                    if (pfo) {
                        const long long t1 = clock64();
                        sm.pacc[15] += (t1 - to0);  // pivot block handed to the copy unit
                        to0 = t1;
                    }
                }
#pragma unroll 4
                for (int u = 0; u < NB; ++u) {
                    const double mm = mp[u][k1];
#pragma unroll
                    for (int b = 0; b < 7; ++b) row[b] = fma(mm, rw[u][tx + 32 * b], row[b]);
                }
#pragma unroll
                for (int b = 0; b < 7; ++b) {
#pragma unroll
                    for (int k = 0; k < GB_NR; ++k) acc[k][b] = (k == k1) ? row[b] : acc[k][b];
                }
                if (pfo) {
                    const long long t1 = clock64();
                    sm.pacc[13] += (t1 - to0);  // rest of the row
                    to0 = t1;
                }
                // (its entries in the pivot columns of step m are -G[s][i] (i is unswept) = MP[i][s]: patched into the staged copy)
                publish_row(m + 1, s_row, row, K0, mp[tx & (NB - 1)][k1]);
            }
            // ---- the rank-NB update of everything else this thread owns --------------------------------------------------
            // Straight-line: no branch per row.  Padding rows have zero multipliers in the table and the look-ahead row gets zeros
            // here (x + 0 g = x exactly), so that all 18 shared-memory loads of two panel rows are in flight before the 56 FMAs (with
            // a branch per row ptxas issued every multiplier load right in front of the FMAs that need it: 2.8 us per step).
            const bool kd0 = k_done == 0, kd1 = k_done == 1, kd2 = k_done == 2, kd3 = k_done == 3;
            long long tb0 = pfu ? clock64() : 0;
            auto bulk = [&](auto WITH3) {  // (two straight-line instances: only warp 0 of a CTA has a real fourth row at n = 200)
                constexpr bool with3 = decltype(WITH3)::value != 0;
#pragma unroll 2  // (fully unrolled, ptxas hoists all 56 panel loads and spills)
                for (int u = 0; u < NB; ++u) {
                    double g[7];
                    const double2 ma = *reinterpret_cast<const double2*>(&mp[u][0]), mb = *reinterpret_cast<const double2*>(&mp[u][2]);
#pragma unroll
                    for (int b = 0; b < 7; ++b) g[b] = rw[u][tx + 32 * b];
                    const double m0 = gjb_sel(kd0, 0.0, ma.x), m1 = gjb_sel(kd1, 0.0, ma.y), m2 = gjb_sel(kd2, 0.0, mb.x), m3 = gjb_sel(kd3, 0.0, mb.y);
#pragma unroll
                    for (int b = 0; b < 7; ++b) {
                        acc[0][b] = fma(m0, g[b], acc[0][b]);
                        acc[1][b] = fma(m1, g[b], acc[1][b]);
                        acc[2][b] = fma(m2, g[b], acc[2][b]);
                        if (with3) acc[3][b] = fma(m3, g[b], acc[3][b]);
                    }
                }
            };
            if (has_row3) bulk(IntC<1>());
            else bulk(IntC<0>());
            if (pfu) sm.pacc[5] += clock64() - tb0;  // the bulk update alone
            // ---- fix-ups: pivot rows <- G (P^-1 inside the block), pivot columns <- MP[i][s] -------------------------------
            const int q0 = (ty - RB * m) & 7;
            if (q0 < RB) {  // this warp owns one of the CTA's pivot rows: local row RB m + q0, row rank + 8 q0 of the block
                const int k0 = (RB * m + q0) >> 3, s0 = rank + GB_CL * q0;
                const double* pv = &sm.Pinv[par][s0 * NB];
                double v[7];
#pragma unroll
                for (int b = 0; b < 7; ++b) v[b] = 0.0;
#pragma unroll 4
                for (int u = 0; u < NB; ++u) {
                    const double pu = pv[u];
#pragma unroll
                    for (int b = 0; b < 7; ++b) v[b] = fma(pu, rw[u][tx + 32 * b], v[b]);  // G[s0][j]
                }
                const double pin = pv[tx & (NB - 1)];
#pragma unroll
                for (int b = 0; b < 7; ++b) {
                    const int j = tx + 32 * b;
                    const double vb = gjb_sel(j >= K0 && j < K0 + NB, pin, v[b]);
#pragma unroll
                    for (int k = 0; k < GB_NR; ++k) acc[k][b] = gjb_sel(k == k0, vb, acc[k][b]);
                }
            }
            if ((tx & ~(NB - 1)) == (K0 & 31)) {
                const int sc = tx & (NB - 1), b0 = K0 >> 5;
#pragma unroll
                for (int k = 0; k < GB_NR; ++k) {
                    const int l = ty + 8 * k;
                    const bool fix = l < GB_LR && !(l >= RB * m && l < RB * m + RB);
                    const double v = mp[sc][k];
#pragma unroll
                    for (int b = 0; b < 7; ++b) acc[k][b] = gjb_sel(fix && b == b0, v, acc[k][b]);
                }
            }
            if (pfu) {
                const long long t1 = clock64();
                sm.pacc[11] += (t1 - tu0);
                tu0 = t1;
            }
        }
        // ---- results ---------------------------------------------------------------------------------------------------------
#if !defined(COVO_CPU_EMU)
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the combine kernel may become resident; it waits for this grid to complete
#endif
        if (!want_logdet) {
            float* Xg = a.Xbuf + ((long long)env * kDensePoles + pole) * n * n;
#pragma unroll
            for (int k = 0; k < GB_NR; ++k)
#pragma unroll
                for (int b = 0; b < 7; ++b) {
                    const int j = tx + 32 * b;
                    const int l = ty + 8 * k, i = rank + GB_CL * l;
                    if (l < GB_LR && i < n && j <= i)  // (reversed positions: this is the upper triangle of the inverse)
                        Xg[(long long)(n - 1 - i) * n + (n - 1 - j)] = (float)(wj * acc[k][b]);
                }
        } else if (rank == 0 && ty < 7) {  // log det A = sum of the logarithms of the scalar pivots (every CTA holds all of them)
            double lp = 0.0;
            const int i = tx + 32 * ty;
            if (i < n) lp = log(sm.piv[i]);
            lp = warp_sum_d(lp);
            double* red = &sm.MP[0][0][0][0];  // dead: every step is over for these warps
            COVO_NAMED_BARRIER(3, 224);
            if (tx == 0) red[ty] = lp;
            COVO_NAMED_BARRIER(3, 224);
            if (tid == 0) {
                double sum = 0.0;
                for (int w = 0; w < 7; ++w) sum += red[w];
                a.scal[(long long)env * 4 + 2] = sum;
            }
        }
        if (tid == 0 && sm.bad) a.status[env] = 2;
    }
    gjb_cluster_wait();  // the arrive of the last step
    __syncthreads();
    if (a.prof && tid >= 5 && tid < 16 && blockIdx.x == 0 && blockIdx.y == 0) a.prof[48 + tid] = sm.pacc[tid];
    gjb_cluster_sync();  // nobody leaves while a peer could still be sending to it
}

// ---------------------------------------------------------------------------------------------------------------------------
// D3
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) combine_kernel(const DenseArgs a) {
    const int n = a.n, env = blockIdx.y;
#if !defined(COVO_CPU_EMU)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    const double logdet = a.scal[(long long)env * 4 + 2];
    // controllers/covo.py:123-127: log_const = (2 * n * 2 log(sigma) + sum log o) / n;  Sigma = exp(log_const / 2) A^(-1/2)
    const double log_const = (4.0 * (double)n * log((double)a.sample_sigma) + logdet) / (double)n;
    const float scale = (float)exp(0.5 * log_const);
    const float* Xg = a.Xbuf + (long long)env * kDensePoles * n * n;
    float* cov = a.cov + (long long)env * n * n;
    const int npairs = n * (n + 1) / 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < npairs; q += gridDim.x * blockDim.x) {
        int ia = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
        while (ia * (ia + 1) / 2 > q) --ia;
        while ((ia + 1) * (ia + 2) / 2 <= q) ++ia;
        const int ib = q - ia * (ia + 1) / 2;
        float s = 0.f;
        const int I = n - 1 - ia, J = n - 1 - ib;  // I <= J: the inverse kernels store the upper triangle (they work back to front)
#pragma unroll
        for (int j = 0; j < kDensePoles; ++j) s += Xg[(long long)j * n * n + I * n + J];
        s *= scale;
        cov[I * n + J] = s;
        cov[J * n + I] = s;  // (a_cov + a_cov.T)/2 (:132) holds by construction
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
size_t sigma_dense_scratch_floats(int n) { return (size_t)kDensePoles * n * n; }

#if !defined(COVO_CPU_EMU)
cudaError_t launch_sigma_dense(const SigmaArgs& s, double* scal, float* Xbuf, int n_env, cudaStream_t st, cudaEvent_t ev_mid1, cudaEvent_t ev_mid2) {
    if (s.n > kSigmaMaxN || (s.n & 3)) return cudaErrorInvalidValue;
    DenseArgs a;
    a.n = s.n;
    a.n_pad = s.n_pad;
    a.sample_sigma = s.sample_sigma;
    a.R = s.R;
    a.scal = scal;
    a.Xbuf = Xbuf;
    a.cov = s.cov;
    a.zolo = s.zolo + (size_t)kZoloLadder * 2 * kZoloPoles;  // the 13-pole ladder sits behind E2's 16-pole one (zolotarev_table_all)
    a.status = s.status;
    a.prof = s.prof;
    cudaError_t e;
    a.Asym = nullptr;  // (the pole-inverse kernel symmetrises R itself: it loads the matrix while the Lanczos kernel is still running)
    // COVO_DENSE_PDL=0: plain stream order between the three kernels; also when the batch does not fit the device next to the Lanczos clusters
    static const bool pdl_env = !(getenv("COVO_DENSE_PDL") && getenv("COVO_DENSE_PDL")[0] == '0');
    const bool pdl = pdl_env && n_env == 1;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(LC_CL, n_env);
        cfg.blockDim = dim3(LC_TT);
        cfg.dynamicSmemBytes = sizeof(LcSmem);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = LC_CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, lanczos_cluster_kernel, a);
        if (e != cudaSuccess) return e;
    }
    if (ev_mid1) cudaEventRecord(ev_mid1, st);
    {  // blocked Gauss-Jordan, one cluster per pole (+ one for log det A)
        // COVO_GJB_NB=16: sixteen pivots per elimination step instead of eight (see GjbSmemT).  Measured on B200: 92 us against 70 -- the
        // 16 x 16 inverse on one warp is 440 cycles per pivot (eight entries per lane: the instruction count per thread, not the
        // dependent chain, sets the pace of a lone warp), 3.7 us per block where two 8 x 8 inverses take 1.8
        static const int nb = (getenv("COVO_GJB_NB") && atoi(getenv("COVO_GJB_NB")) == 16) ? 16 : 8;
        static size_t conf8[32] = {}, conf16[32] = {};
        const size_t smem = nb == 16 ? sizeof(GjbSmemT<16>) : sizeof(GjbSmemT<8>);
        e = nb == 16 ? ensure_smem_attr(gjb_inverse_kernel_t<16>, smem, conf16) : ensure_smem_attr(gjb_inverse_kernel_t<8>, smem, conf8);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(GB_CL * (kDensePoles + 1), n_env);
        cfg.blockDim = dim3(GB_T);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        // programmatic dependent launch behind the Lanczos kernel: the clusters become resident next to it (8 + 112 CTAs of one per SM),
        // set up their barriers and load the matrix while the recurrence runs; griddepcontrol.wait in the kernel before lambda_min is read
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = GB_CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 2 : 1;
        e = nb == 16 ? cudaLaunchKernelEx(&cfg, gjb_inverse_kernel_t<16>, a) : cudaLaunchKernelEx(&cfg, gjb_inverse_kernel_t<8>, a);
        if (e != cudaSuccess) return e;
    }
    if (ev_mid2) cudaEventRecord(ev_mid2, st);
    {
        const int npairs = a.n * (a.n + 1) / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((npairs + 255) / 256, n_env);
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 1 : 0;
        e = cudaLaunchKernelEx(&cfg, combine_kernel, a);
    }
    return e;
}

#endif

}  // namespace covo
