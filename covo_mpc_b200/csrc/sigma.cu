// K5: CoVO covariance.  Replaces controllers/covo.py:116-132 (optimize_sigma) and the Cholesky
// factorisation inside jax.random.multivariate_normal (controllers/covo.py:216).
//
//   reference:  R <- (R+R^T)/2;  (lam, U) = eigh(R);  o = lam - lam_min + 1e-2;
//               Sigma = U diag(exp(c/2) o^(-1/2)) U^T,  c = (4 n log sigma + sum log o)/n;  L = chol(Sigma)
//
// Sigma is a matrix function of R: Sigma = exp(c/2) (R - lam_min I + 1e-2 I)^(-1/2), c from log det.
// No eigenvectors are needed, so the serial n = 200 eigensolve (LAPACK ssyevd in the reference; ~1 ms
// of Jacobi sweeps on one SM) is replaced by
//   E1  Householder tridiagonalisation R = Q T Q^T in one CTA, matrix resident in shared memory (fp32,
//       the same backward-stable reduction ssyevd starts with);
//   E2  on the tridiagonal T, in fp64: Gershgorin bounds, lam_min by 1024-way Sturm multisection,
//       T_s = T - lam_min + 1e-2 (SPD, smallest eigenvalue 1e-2 by construction), log det T_s from its
//       LDL^T pivots, and T_s^(-1/2) by the Zolotarev rational approximation of x^(-1/2) on [m, M]
//           x^(-1/2) ~= sum_j w_j / (x + t_j),  t_j = m sc^2(u_j|k), w_j = (2 K sqrt(m) / (pi N)) dn/cn^2,
//           u_j = (j - 1/2) K / N,  k^2 = 1 - m/M                    (Hale, Higham, Trefethen 2008),
//       each (T_s + t_j)^(-1) written down entry by entry from the forward/backward pivots
//       (inverse of a tridiagonal is semiseparable);  error ~ exp(-2 pi N K'/K) < 1e-9 for N = 16;
//   E3  Sigma = Q F Q^T: two passes of "apply the n-2 reflectors to every column", one warp per column;
//   E4  blocked right-looking Cholesky in one CTA; emits L (row-major) and the packed k-major factor the
//       rollout kernel streams with TMA.
// Accuracy vs the float64 oracle on the benchmark Hessian: ||Sigma - Sigma_ref||_F / ||Sigma_ref||_F
// ~ 8e-7 (LAPACK float32: 4e-7); a float32 one-sided Jacobi needs ~10 sweeps and reaches only 2e-4.
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>
#include <string.h>
#include <stdlib.h>

#include "sigma.cuh"

namespace covo {

// ---------------------------------------------------------------------------------------------
// host: Jacobi elliptic functions by the AGM / descending Landen transformation
// ---------------------------------------------------------------------------------------------
namespace {

void ellip_agm(double m, double* a, double* c, int& N) {
    // a[0] = 1, b = sqrt(1-m), c[0] = sqrt(m)
    double an = 1.0, bn = sqrt(1.0 - m);
    a[0] = an;
    c[0] = sqrt(m);
    N = 0;
    while (fabs(c[N]) > 1e-17 && N < 30) {
        double a1 = 0.5 * (an + bn), b1 = sqrt(an * bn), c1 = 0.5 * (an - bn);
        ++N;
        a[N] = a1;
        c[N] = c1;
        an = a1;
        bn = b1;
    }
}

void ellipj_host(double u, double m, double& sn, double& cn, double& dn, double& K) {
    double a[32], c[32];
    int N;
    ellip_agm(m, a, c, N);
    K = M_PI / (2.0 * a[N]);
    double phi = ldexp(a[N] * u, N);
    for (int i = N; i >= 1; --i) phi = 0.5 * (phi + asin(c[i] * sin(phi) / a[i]));
    sn = sin(phi);
    cn = cos(phi);
    dn = sqrt(1.0 - m * sn * sn);
}

}  // namespace

void zolotarev_nodes(double m, double M, int N, double* t, double* w) {
    const double k2 = 1.0 - m / M;
    double sn, cn, dn, K;
    ellipj_host(0.0, k2, sn, cn, dn, K);
    for (int j = 0; j < N; ++j) {
        double u = (j + 0.5) * K / N;
        ellipj_host(u, k2, sn, cn, dn, K);
        t[j] = m * (sn / cn) * (sn / cn);
        w[j] = (2.0 * K * sqrt(m) / (M_PI * N)) * dn / (cn * cn);
    }
}

static double zolo_m() { return kCovoOffset * (1.0 - 1e-7); }
double zolotarev_ladder_M(int i) { return zolo_m() * pow(4.0, 4 + i); }

void zolotarev_table(double* table) {
    for (int i = 0; i < kZoloLadder; ++i)
        zolotarev_nodes(zolo_m(), zolotarev_ladder_M(i), kZoloPoles, table + (size_t)i * 2 * kZoloPoles,
                        table + (size_t)i * 2 * kZoloPoles + kZoloPoles);
}

void zolotarev_table_dense(double* table) {
    for (int i = 0; i < kZoloLadder; ++i)
        zolotarev_nodes(zolo_m(), zolotarev_ladder_M(i), kDensePoles, table + (size_t)i * 2 * kDensePoles,
                        table + (size_t)i * 2 * kDensePoles + kDensePoles);
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int TT = 1024;
constexpr int NPOLE = kZoloPoles;

__device__ __forceinline__ double wsumd(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double wmind(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double wmaxd(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Three-term recurrences on the tridiagonal, evaluated by ONE WARP as a prefix product of 2x2 matrices:
//   p_i = alpha_i p_{i-1} - beta_i p_{i-2},  p_{-1} = 1, p_{-2} = 0      (Sturm sequence / pivot numerators)
//   (p_i, p_{i-1})^T = M_i ... M_0 (1, 0)^T,  M_i = [[alpha_i, -beta_i], [1, 0]].
// Lane L multiplies the matrices of its segment of ceil(n/32) indices, a Kogge-Stone scan combines the segments
// (5 levels), and the lane walks its segment again from the prefix state.  The dependent chain is ~2 x 7 + 5
// matrix products instead of n = 200 steps.  Only sign patterns and the ratios p_i / p_{i-1} are used, so every
// partial product is rescaled by a power of two (no overflow, exact).
// ---------------------------------------------------------------------------------------------
struct M2 {
    double a, b, c, d;  // [[a, b], [c, d]]
};
__device__ __forceinline__ M2 m2_mul(const M2& x, const M2& y) {  // x * y
    M2 r;
    r.a = fma(x.a, y.a, x.b * y.c);
    r.b = fma(x.a, y.b, x.b * y.d);
    r.c = fma(x.c, y.a, x.d * y.c);
    r.d = fma(x.c, y.b, x.d * y.d);
    return r;
}
__device__ __forceinline__ M2 m2_normalise(const M2& x) {
    const int e = max(max(__double2hiint(x.a) & 0x7ff00000, __double2hiint(x.b) & 0x7ff00000),
                      max(__double2hiint(x.c) & 0x7ff00000, __double2hiint(x.d) & 0x7ff00000));
    const double sc = __hiloint2double(0x7fe00000 - e, 0);  // 2^(1023 - biased exponent of the largest entry)
    M2 r;
    r.a = x.a * sc; r.b = x.b * sc; r.c = x.c * sc; r.d = x.d * sc;
    return r;
}
__device__ __forceinline__ M2 m2_shfl_up(const M2& x, int off) {
    M2 r;
    r.a = __shfl_up_sync(0xffffffffu, x.a, off);
    r.b = __shfl_up_sync(0xffffffffu, x.b, off);
    r.c = __shfl_up_sync(0xffffffffu, x.c, off);
    r.d = __shfl_up_sync(0xffffffffu, x.d, off);
    return r;
}
// AB(s, alpha, beta) supplies the coefficients of step s (beta of step 0 is ignored); EMIT(s, p_s, p_{s-1}) consumes
// the sequence (both values carry the same positive scale factor).  All 32 lanes of the warp must call.
template <class AB, class EMIT>
__device__ __forceinline__ void warp_recurrence(int n, AB ab, EMIT emit) {
    const int lane = threadIdx.x & 31;
    const int seg = (n + 31) >> 5, s0 = lane * seg, s1 = min(s0 + seg, n);
    M2 P = {1.0, 0.0, 0.0, 1.0};
    for (int s = s0; s < s1; ++s) {
        double al, be;
        ab(s, al, be);
        if (s == 0) be = 0.0;
        const M2 t = P;
        P.a = fma(al, t.a, -be * t.c);
        P.b = fma(al, t.b, -be * t.d);
        P.c = t.a;
        P.d = t.b;
    }
    P = m2_normalise(P);
    // entries are <= 2 after the rescale; a product at most squares-and-doubles the bound (8, 128, 3e4, ...), so
    // one more rescale in the middle of the five levels is plenty
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const M2 Q = m2_shfl_up(P, off);
        if (lane >= off) P = m2_mul(P, Q);
        if (off == 4) P = m2_normalise(P);
    }
    const M2 Q = m2_shfl_up(P, 1);
    double p = (lane == 0) ? 1.0 : Q.a, pm = (lane == 0) ? 0.0 : Q.c;
    for (int s = s0; s < s1; ++s) {
        double al, be;
        ab(s, al, be);
        if (s == 0) be = 0.0;
        const double pn = fma(al, p, -be * pm);
        emit(s, pn, p);
        pm = p;
        p = pn;
    }
}
// does tridiag(d, e) have an eigenvalue below x?  (sign change in the Sturm sequence; a zero term counts as a
// change, which is the conservative answer for the bracket)
__device__ __forceinline__ bool warp_has_eig_below(const double* d, const double* e2, int n, double x) {
    bool neg = false;
    warp_recurrence(
        n, [&](int s, double& al, double& be) { al = d[s] - x; be = (s > 0) ? e2[s - 1] : 0.0; },
        [&](int, double pn, double p) {
            // signs on the high words; +-0 of either term is treated as a change
            const int hn = __double2hiint(pn), hp = __double2hiint(p);
            neg |= ((hn ^ hp) < 0) || (pn == 0.0) || (p == 0.0);
        });
    return __any_sync(0xffffffffu, neg);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// E1: Householder tridiagonalisation with the matrix resident in REGISTERS of a thread-block CLUSTER
// ---------------------------------------------------------------------------------------------
// NC CTAs x 8 warps.  Element (i, j) of A lives in lane (i mod 32), row slot i / 32 of the warp that owns column
// j; global warp g = 8 rank + warp owns the column PAIRS (2 g + 16 NC c, +1), c = 0..C2-1, each pair held as one
// float2 so that the rank-2 update and the matvec run on the packed FFMA2 pipe of sm_100.  Every warp holds
// complete columns (rows over the lanes):
//   * the matvec q = A v' is 7 FFMA2 per column pair followed by ONE transposition through a private
//     shared-memory tile -- no shared-memory traffic for A itself, ever;
//   * the scalar work of a step -- s = q.v, w = tau q - (tau^2 s / 2) v, the UPDATED row m
//         r_i = A[m][i] - v_i w_m - w_i      (look-ahead: the next Householder vector is made of it),
//     its norm, (beta, tau') and v' -- is done REDUNDANTLY by every warp of every CTA on 7 values per lane: no
//     block-wide reduction, no special warp, ONE (cluster) barrier per step;
//   * what a step exchanges is tiny: every warp writes its entries of q and of row m+1 into the shared memory
//     of ALL CTAs of the cluster (distributed shared memory), 2 x n floats per step in total.
// Iteration m (0 <= m < n), LAPACK ssytd2 recurrences; state: A carries reflectors 0..m-2; (v, tau, q = A v)
// belong to reflector m-1:
//   1. s, w;  r = updated row m -> d_m = r_m, Householder of r[m+1:] -> e_m = beta, tau_m, v_m
//   2. A <- A - v w^T - w v^T     (finished columns have v_j = w_j = 0: no masking, no branches)
//   3. publish row m+1;  q = A v_m  (entries j <= m are stored as zero, which makes w_j = 0 there)  -> barrier
// Why a cluster: one SM is issue-/latency-bound on this loop (two single-CTA versions, 16 warps x 128 registers
// and 8 warps x 255 registers, ran at 0.25 IPC per warp: the register file is full of A, so nothing can be
// software-pipelined).  Spreading the columns over NC SMs leaves ~60 registers of A per thread, straight-line
// code, and the per-step cost becomes the redundant scalar chain plus one cluster barrier.
constexpr int EW = 8;         // warps per CTA
constexpr int ET = EW * 32;   // threads per CTA
constexpr int RS = kSigmaMaxN / 32;  // row slots per lane (7)
constexpr int kTilePitch = 36;       // floats; rows of the transposition tile are float4-aligned and conflict-free

__host__ __device__ inline size_t e1_smem_bytes(int n, int c2) {
    // reflector store Vs [n][n]; q[2][256], v[2][256], row[2][256], tau[256], d[256], e[256]; transposition tiles;
    // mbarrier pair; reduction scratch; landing pad of the per-warp progress signals
    return ((size_t)n * n + 256 * 9 + EW * 2 * c2 * kTilePitch + 4 + EW * 64 + 64) * sizeof(float);
}

__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ unsigned dsmem_addr(unsigned saddr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_st2(unsigned addr, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
// Store that carries its own completion signal: the destination CTA's mbarrier receives 4 bytes of
// transaction count when the value has landed -- no fence, no separate arrive.
__device__ __forceinline__ void dsmem_st_signal(unsigned addr, float v, unsigned mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(addr), "f"(v),
                 "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
    }
}
// MUFU approximations + one Newton step (~1 ulp), without the IEEE slow paths of sqrtf / division
__device__ __forceinline__ float rcp_newton(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * fmaf(-x, r, 2.f);
}
__device__ __forceinline__ float rsqrt_newton(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * fmaf(-0.5f * x * r, r, 1.5f);
}

// Warp all-reduce through shared memory: one STS, eight broadcast LDS.128 and an add tree -- about half the
// latency of the five dependent SHFL + FADD levels of a butterfly.  Every lane adds in the same order, so all
// lanes (and all warps given the same inputs) obtain bit-identical sums.
__device__ __forceinline__ float wsum_smem(float x, float* scratch /* [32], warp-private, 16-byte aligned */) {
    scratch[threadIdx.x & 31] = x;
    __syncwarp();
    const float4* p4 = reinterpret_cast<const float4*>(scratch);
    const float4 a0 = p4[0], a1 = p4[1], a2 = p4[2], a3 = p4[3], a4 = p4[4], a5 = p4[5], a6 = p4[6], a7 = p4[7];
    const float b0 = (a0.x + a0.y) + (a0.z + a0.w), b1 = (a1.x + a1.y) + (a1.z + a1.w);
    const float b2 = (a2.x + a2.y) + (a2.z + a2.w), b3 = (a3.x + a3.y) + (a3.z + a3.w);
    const float b4 = (a4.x + a4.y) + (a4.z + a4.w), b5 = (a5.x + a5.y) + (a5.z + a5.w);
    const float b6 = (a6.x + a6.y) + (a6.z + a6.w), b7 = (a7.x + a7.y) + (a7.z + a7.w);
    return ((b0 + b1) + (b2 + b3)) + ((b4 + b5) + (b6 + b7));
}

template <int C2, int NC>
__global__ void __launch_bounds__(ET, 1) tridiag_reg_kernel(const SigmaArgs a) {
    extern __shared__ __align__(16) float esm[];
    const int n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (NC > 1) ? (int)cluster_rank() : 0;
    const int env = blockIdx.x / NC;
    const int gw = rank * EW + warp;        // global warp: owns columns 2 gw + CS c, 2 gw + CS c + 1
    constexpr int GW = EW * NC;             // warps of the cluster
    constexpr int CS = 2 * GW;              // column stride between pair slots
    const int n_idle = GW - min(GW, n >> 1);  // warps that own no column at all (n < 2 GW)
    float* Vs = esm;                    // [n][n] reflectors (row k = v_k), every CTA keeps a copy
    float* qsm0 = esm + n * n;          // [2][256]
    float* vsm0 = qsm0 + 512;           // [2][256]
    float* row0 = vsm0 + 512;           // [2][256]
    float* taus = row0 + 512;           // [256]
    float* dsm = taus + 256;            // [256]
    float* esm_e = dsm + 256;           // [256]
    float* tile = esm_e + 256 + warp * 2 * C2 * kTilePitch;  // [2 C2][36] private transposition tile
    unsigned long long* mbars = reinterpret_cast<unsigned long long*>(esm_e + 256 + EW * 2 * C2 * kTilePitch);  // [2]
    float* red = esm_e + 256 + EW * 2 * C2 * kTilePitch + 4 + warp * 64;  // [2][32] private reduction scratch
    float* pad = esm_e + 256 + EW * 2 * C2 * kTilePitch + 4 + EW * 64;     // [64] landing pad of the progress signals

    const float* Rg = a.R + (long long)env * n * n;
    float* Vg = a.Vh + (long long)env * n * n;
    float* taug = a.tau + (long long)env * n;
    COVO_STAMP(a, 8);
    for (int i = tid; i < 256 * 9; i += ET) qsm0[i] = 0.f;
    const unsigned mbar_local = (unsigned)__cvta_generic_to_shared(mbars);
    if (tid == 0) {
        mbar_init(mbar_local, 1);
        mbar_init(mbar_local + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // (R + R^T)/2 (controllers/covo.py:117) straight from HBM / L2 into registers
    float2 A2[RS][C2];
#pragma unroll
    for (int c = 0; c < C2; ++c) {
        const int j0 = 2 * gw + CS * c;
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int i = lane + 32 * r;
            float2 x = make_float2(0.f, 0.f);
            if (i < n && j0 < n) {  // n is a multiple of 4 and j0 is even: j0 + 1 < n as well
                const float2 rowv = *reinterpret_cast<const float2*>(Rg + (long long)i * n + j0);
                x.x = 0.5f * (rowv.x + Rg[(long long)j0 * n + i]);
                x.y = 0.5f * (rowv.y + Rg[(long long)(j0 + 1) * n + i]);
            }
            A2[r][c] = x;
        }
    }
    // shared::cluster addresses of the exchange buffers and of the mbarrier pair in every CTA of the cluster
    unsigned q_remote[NC], row_remote[NC], bar_remote[NC], pad_remote[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        pad_remote[k] = dsmem_addr((unsigned)__cvta_generic_to_shared(pad), k);
        q_remote[k] = dsmem_addr((unsigned)__cvta_generic_to_shared(qsm0), k);
        row_remote[k] = dsmem_addr((unsigned)__cvta_generic_to_shared(row0), k);
        bar_remote[k] = dsmem_addr(mbar_local, k);
    }
    cluster_barrier();  // zero-fill and mbarrier init done everywhere before anybody publishes
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < C2; ++c) {
            const int j0 = 2 * gw + CS * c;
            if (j0 < n) {
#pragma unroll
                for (int k = 0; k < NC; ++k) dsmem_st2(row_remote[k] + 4 * j0, A2[0][c]);  // row 0
            }
        }
    }
    cluster_barrier();
    COVO_STAMP(a, 15);

    float vi[RS];
#pragma unroll
    for (int r = 0; r < RS; ++r) vi[r] = 0.f;
    float tau = 0.f;
    long long pa[4] = {0, 0, 0, 0}, pt0 = 0;
#define PH(i) do { if (a.prof && tid == 0 && blockIdx.x == 0) { long long t_ = clock64(); pa[i] += t_ - pt0; pt0 = t_; } } while (0)
    if (a.prof && tid == 0) pt0 = clock64();
    for (int m = 0; m < n; ++m) {
        const int p = (m & 1) << 8, pn = p ^ 256;
        const float* qsm = qsm0 + p;    // q = A v of reflector m-1 (zero for indices < m)
        const float* vsm = vsm0 + p;    // v of reflector m-1 (zeros for m = 0)
        const float* rowm = row0 + p;   // row m of A (reflectors <= m-2 applied)
        // ---- 1. scalar work, redundantly in every warp ------------------------------------------------
        float wi[RS], ri[RS];
        float pe = 0.f, po = 0.f;
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            wi[r] = qsm[lane + 32 * r];  // q_i for now
            ri[r] = rowm[lane + 32 * r];
            if (r & 1) po = fmaf(wi[r], vi[r], po); else pe = fmaf(wi[r], vi[r], pe);
        }
        const int m1 = min(m + 1, n - 1);
        const float q_m = qsm[m], q_m1 = qsm[m1], row_m = rowm[m], row_m1 = rowm[m1], v_m1 = vsm[m1];
        // column values of the rank-2 update, loaded ahead of the reduction
        float2 vc[C2], qc[C2];
#pragma unroll
        for (int c = 0; c < C2; ++c) {
            vc[c] = *reinterpret_cast<const float2*>(vsm + 2 * gw + CS * c);
            qc[c] = *reinterpret_cast<const float2*>(qsm + 2 * gw + CS * c);
        }
        const float s = wsum_smem(pe + po, red);
        const float c2 = 0.5f * tau * tau * s;  // tau = 0 when there is no previous reflector: w = 0
        const float w_m = fmaf(tau, q_m, -c2);  // v_{m-1}[m] = 1
        const float w_m1 = fmaf(tau, q_m1, -c2 * v_m1);
        float sge = 0.f, sgo = 0.f;
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int i = lane + 32 * r;
            const float w = fmaf(tau, wi[r], -c2 * vi[r]);  // zero for i < m (q and v are)
            wi[r] = w;
            const float x = (i >= m + 2 && i < n) ? ri[r] - vi[r] * w_m - w : 0.f;
            ri[r] = x;
            if (r & 1) sgo = fmaf(x, x, sgo); else sge = fmaf(x, x, sge);
        }
        // ---- 2. rank-2 update with reflector m-1, in the shadow of the second reduction ------------------
        // (unconditional: tau = 0 gives w = 0 and a zero update; finished columns have v_j = w_j = 0)
        {
            const float2 tau2 = make_float2(tau, tau), nc2 = make_float2(-c2, -c2);
#pragma unroll
            for (int c = 0; c < C2; ++c) {
                const float2 wc = __ffma2_rn(tau2, qc[c], __fmul2_rn(nc2, vc[c]));
#pragma unroll
                for (int r = 0; r < RS; ++r)
                    A2[r][c] = __ffma2_rn(make_float2(-vi[r], -vi[r]), wc,
                                          __ffma2_rn(make_float2(-wi[r], -wi[r]), vc[c], A2[r][c]));
            }
        }
        const unsigned bar_off = (m & 1) << 3;
        // publish column m+1 (= row m+1): it sits in ONE warp, already in the (lane + 32 r) layout of the vectors
        if (m + 1 < n && (((m + 1) >> 1) & (GW - 1)) == gw) {
            const int cstar = (m + 1) / CS;
            const bool hi = (m + 1) & 1;
#pragma unroll
            for (int c = 0; c < C2; ++c)
                if (c == cstar) {
#pragma unroll
                    for (int r = 0; r < RS; ++r) {
                        const float val = hi ? A2[r][c].y : A2[r][c].x;
#pragma unroll
                        for (int k = 0; k < NC; ++k)
                            dsmem_st_signal(row_remote[k] + 4 * (pn + lane + 32 * r), val, bar_remote[k] + bar_off);
                    }
                }
        }
        const float sg = wsum_smem(sge + sgo, red + 32);
        const float dm = row_m - 2.f * w_m;
        const float x0 = (m + 1 < n) ? row_m1 - v_m1 * w_m - w_m1 : 0.f;
        // Householder scalars, branch-free: norm = ||(x0, r)||, beta = -sign(x0) norm, t = x0 - beta,
        // tau = t / (sign(x0) norm) = |t| / norm, scale = 1 / t   (sigma = 0: tau = scale = 0, beta = x0)
        const bool live = sg > 1e-30f;
        const float nn2 = fmaf(x0, x0, live ? sg : 1.f);
        const float rs = rsqrt_newton(nn2);
        const float nrm = copysignf(nn2 * rs, x0);
        const float t = x0 + nrm;
        const float nbeta = live ? -nrm : x0;
        const float ntau = live ? fabsf(t) * rs : 0.f;
        const float nscale = live ? rcp_newton(t) : 0.f;
        float vn[RS];
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int i = lane + 32 * r;
            vn[r] = (i == m + 1 && i < n) ? 1.f : ri[r] * nscale;  // v_m (ri is zero outside m+2 <= i < n)
        }
        const bool do_matvec = (m <= n - 3) && (ntau != 0.f);
        PH(0);
        // ---- 3. q = A v_m ---------------------------------------------------------------------------------
#pragma unroll
        for (int r = 0; r < RS; ++r) vi[r] = vn[r];
        tau = ntau;
        if (do_matvec) {
#pragma unroll
            for (int c = 0; c < C2; ++c) {
                float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
                for (int r = 0; r < RS; r += 2) sa = __ffma2_rn(A2[r][c], make_float2(vi[r], vi[r]), sa);
#pragma unroll
                for (int r = 1; r < RS; r += 2) sb = __ffma2_rn(A2[r][c], make_float2(vi[r], vi[r]), sb);
                const float2 t = __fadd2_rn(sa, sb);
                // transposition through the warp's tile: partial sums of column k at tile[k][lane]
                tile[(2 * c) * kTilePitch + lane] = t.x;
                tile[(2 * c + 1) * kTilePitch + lane] = t.y;
            }
            __syncwarp();
            if (lane < 2 * C2) {  // lane L adds up row L = the 32 partials of its column k = L
                const float4* rowp = reinterpret_cast<const float4*>(tile + lane * kTilePitch);
                float4 t4 = rowp[0];
#pragma unroll
                for (int t = 1; t < 8; ++t) {
                    const float4 x = rowp[t];
                    t4.x += x.x; t4.y += x.y; t4.z += x.z; t4.w += x.w;
                }
                const float t1 = (t4.x + t4.y) + (t4.z + t4.w);
                const int j = 2 * gw + CS * (lane >> 1) + (lane & 1);
                if (j < n) {
                    const float qv = (j >= m + 1) ? t1 : 0.f;
#pragma unroll
                    for (int k = 0; k < NC; ++k) dsmem_st_signal(q_remote[k] + 4 * (pn + j), qv, bar_remote[k] + bar_off);
                }
            }
            __syncwarp();  // the tile is rewritten in the next step
        }
        // Progress signal.  A phase must complete only when EVERY warp of the cluster is done reading the step's
        // buffers (they are overwritten in the next step).  Normally the q entries a warp sends after its reads say
        // so; when no q is exchanged (tau = 0, the last two steps) or the warp owns no column (tiny n), it sends
        // 4 bytes to every CTA instead.
        if (!do_matvec || 2 * gw >= n) {  // warp-uniform
            __syncwarp();                 // every lane is past its reads
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < NC; ++k) dsmem_st_signal(pad_remote[k] + 4 * gw, 0.f, bar_remote[k] + bar_off);
            }
        }
        if (warp == (m & (EW - 1))) {  // bookkeeping by a rotating warp (every CTA keeps its own copy)
#pragma unroll
            for (int r = 0; r < RS; ++r) {
                const int i = lane + 32 * r;
                if (i < n) {
                    vsm0[pn + i] = vi[r];
                    Vs[m * n + i] = vi[r];
                }
            }
            if (lane == 0) {
                dsm[m] = dm;
                esm_e[m] = nbeta;
                taus[m] = ntau;
            }
            __syncwarp();
            // this step's incoming traffic: column m+1 (7 x 32 floats), the n entries of q if the matvec runs, and one
            // progress signal from every warp of the cluster
            if (lane == 0)
                mbar_expect_tx(mbar_local + bar_off, ((m + 1 < n) ? 4u * 32u * RS : 0u) + (do_matvec ? 4u * (n + n_idle) : 4u * GW));
        }
        PH(2);
        // everything this CTA reads in step m+1 has landed when its mbarrier phase completes
        mbar_wait(mbar_local + bar_off, (m >> 1) & 1);
        PH(3);
    }
#undef PH
    if (a.prof && tid == 0 && blockIdx.x == 0)
        for (int i = 0; i < 4; ++i) a.prof[40 + i] = pa[i];
    __syncthreads();
    if (NC > 1) cluster_barrier();  // nobody leaves while a peer could still be sending to it
    COVO_STAMP(a, 9);
    // every CTA holds the complete result: d, e (fp64, for E2) and tau by rank 0, the reflector rows shared out
    if (rank == 0 && tid < n) {
        double* dg = a.diag + (long long)env * 4 * n;
        dg[tid] = (double)dsm[tid];
        dg[n + tid] = (tid < n - 1) ? (double)esm_e[tid] : 0.0;
        taug[tid] = (tid < n - 2) ? taus[tid] : 0.f;
    }
    {
        const int nv4 = (n * n) >> 2;
        for (int idx = rank * ET + tid; idx < nv4; idx += ET * NC) {
            const int k2 = (4 * idx) / n;
            reinterpret_cast<float4*>(Vg)[idx] = (k2 < n - 2) ? reinterpret_cast<const float4*>(Vs)[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    COVO_STAMP(a, 16);
}

// ---------------------------------------------------------------------------------------------
// E2: everything on the tridiagonal, fp64
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int trifunc_region_floats(int n) {
    // fp64 pivot numerators/denominators [2][2][17][n], then float gd[16][n], and the generator arrays [16][n+1] x 3
    int b = (2 * 2 * (kZoloPoles + 1) * n) * 2 + kZoloPoles * n + 3 * kZoloPoles * (n + 1);
    return (b + 3) & ~3;
}

__global__ void __launch_bounds__(TT, 1) sigma_trifunc_kernel(const SigmaArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    // grid (NB, E): the NB CTAs of one matrix repeat the (cheap, deterministic) scalar stages and share out the rows of F
    const int n = a.n, tid = threadIdx.x, env = blockIdx.y, part = blockIdx.x, nparts = gridDim.x, lane = tid & 31, warp = tid >> 5;
    float* As = reinterpret_cast<float*>(smraw);          // E2 scratch region
    double* dd = reinterpret_cast<double*>(As + trifunc_region_floats(n));  // [256]
    double* ee = dd + 256;                                // [256]
    double* e2s = ee + 256;                               // [256] e^2
    double* sc = e2s + 256;                               // [16] scalars
    double* rdbuf = sc + 16;                              // [64]
    int* ired = reinterpret_cast<int*>(rdbuf + 64);       // [64]
    COVO_STAMP(a, 17);
    {
        const double* dg = a.diag + (long long)env * 4 * n;
        if (tid < n) {
            dd[tid] = dg[tid];
            ee[tid] = dg[n + tid];
        }
    }
    __syncthreads();

    // ---- E2: everything on the tridiagonal, fp64 ---------------------------------------------
    // Gershgorin interval
    {
        double lo = CUDART_INF, hi = -CUDART_INF;
        if (tid < n) {
            double r = ((tid > 0) ? fabs(ee[tid - 1]) : 0.0) + ((tid < n - 1) ? fabs(ee[tid]) : 0.0);
            lo = dd[tid] - r;
            hi = dd[tid] + r;
        }
        if (tid < n) e2s[tid] = ee[tid] * ee[tid];
        lo = wmind(lo);
        hi = wmaxd(hi);
        if (lane == 0) {
            rdbuf[warp] = lo;
            rdbuf[32 + warp] = hi;
        }
        __syncthreads();
        if (tid == 0) {
            double l = rdbuf[0], h = rdbuf[32];
            for (int w = 1; w < TT / 32; ++w) {
                l = fmin(l, rdbuf[w]);
                h = fmax(h, rdbuf[32 + w]);
            }
            double pad = 1e-9 * fmax(1.0, fmax(fabs(l), fabs(h)));
            sc[0] = l - pad;
            sc[1] = h + pad;
            sc[2] = l;
            sc[3] = h;
        }
        __syncthreads();
    }
    COVO_STAMP(a, 10);
    // lam_min by multisection: MS warps evaluate one trial shift each per round with the warp-scan Sturm test; the
    // rounds shrink the Gershgorin bracket by (MS+1)^rounds >= 4e13, i.e. to ~1e-10 absolute.  One barrier per
    // round: the warps post their verdicts, every thread derives the new bracket from them in registers.
    {
        const int MS = a.e2_points;  // 8, 16 or 32
        const int rounds = (MS >= 32) ? 9 : (MS >= 16 ? 12 : 15);
        int* flags = ired;  // [2][32]
        double lo = sc[0], hi = sc[1];
        for (int round = 0; round < rounds; ++round) {
            int* fl = flags + (round & 1) * 32;
            if (warp < MS) {
                const double x = lo + (hi - lo) * ((double)(warp + 1) / (double)(MS + 1));
                const bool below = warp_has_eig_below(dd, e2s, n, x);
                if (lane == 0) fl[warp] = below ? 1 : 0;
            }
            __syncthreads();
            const unsigned m = __ballot_sync(0xffffffffu, lane < MS && fl[lane] != 0);
            const int ts = m ? (__ffs(m) - 1) : MS;  // first trial shift with an eigenvalue below it
            const double step = (hi - lo) / (double)(MS + 1);
            const double nlo = (ts == 0) ? lo : lo + step * ts;
            const double nhi = (ts == MS) ? hi : lo + step * (ts + 1);
            lo = nlo;
            hi = nhi;
        }
        if (tid == 0) {
            sc[0] = lo;
            sc[1] = hi;
        }
        __syncthreads();
    }
    COVO_STAMP(a, 11);
    const double lam_min = 0.5 * (sc[0] + sc[1]);
    const double shift0 = kCovoOffset - lam_min;  // T_s = T + shift0 I
    // ladder index from the Gershgorin upper bound of T_s
    int lad = 0;
    {
        const double Mb = sc[3] + shift0;
        const double m0 = kCovoOffset * (1.0 - 1e-7);
        double Mi = m0 * 256.0;
        while (lad < kZoloLadder - 1 && Mi < Mb) {
            Mi *= 4.0;
            ++lad;
        }
        if (Mi < Mb && tid == 0) a.status[env] = 1;
    }
    const double* zt = a.zolo + (size_t)lad * 2 * NPOLE;  // t_j
    const double* zw = zt + NPOLE;                        // w_j

    // pivots of T_s + t_q, q = 0..15 (q = 16: unshifted, for log det): product recurrences, rescaled
    double* num = reinterpret_cast<double*>(As);                      // [2][17][n]
    double* den = num + 2 * (NPOLE + 1) * n;                          // [2][17][n]
    float* gd = reinterpret_cast<float*>(den + 2 * (NPOLE + 1) * n);  // [16][n]    diag of (T_s+t_q)^-1
    float* lcl = gd + NPOLE * n;                                      // [16][n+1]  c_q[l]
    // one warp-scan per (direction, pole): 34 recurrences over the 32 warps
    for (int task = warp; task < 2 * (NPOLE + 1); task += TT / 32) {
        const bool fwd = task < NPOLE + 1;
        const int q = fwd ? task : task - (NPOLE + 1);
        const double tq = (q < NPOLE) ? zt[q] : 0.0;
        double* nm = num + ((fwd ? 0 : 1) * (NPOLE + 1) + q) * n;
        double* dn = den + ((fwd ? 0 : 1) * (NPOLE + 1) + q) * n;
        const double sh = shift0 + tq;
        warp_recurrence(
            n,
            [&](int s2, double& al, double& be) {
                const int i = fwd ? s2 : n - 1 - s2;
                al = dd[i] + sh;
                be = (s2 > 0) ? e2s[fwd ? i - 1 : i] : 0.0;
            },
            [&](int s2, double pn, double pp) {
                const int i = fwd ? s2 : n - 1 - s2;
                nm[i] = pn;
                dn[i] = pp;
            });
    }
    __syncthreads();
    COVO_STAMP(a, 12);
    // log det T_s = sum log dp_i (unshifted forward pivots)
    {
        double l = 0.0;
        if (tid < n) l = log(num[(0 * (NPOLE + 1) + NPOLE) * n + tid] / den[(0 * (NPOLE + 1) + NPOLE) * n + tid]);
        l = wsumd(l);
        if (lane == 0) rdbuf[warp] = l;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < TT / 32; ++w) s += rdbuf[w];
            sc[4] = s;
        }
    }
    // generators of the semiseparable inverses: (T_s+t_q)^-1[i][l] = g_q[i] * prod_{s=i}^{l-1} c_q[s], l >= i,
    //   g_q[i] = 1 / (dp_i + dm_i - a_i),   c_q[s] = -b_s / dm_{s+1}   (c = 0 where the tridiagonal splits).
    float* cq = lcl;  // [16][n+1]: c_q[l] as float (re-uses the slot of the lo log-prefix, which is not needed)
    for (int idx = tid; idx < NPOLE * n; idx += TT) {
        const int q = idx / n, i = idx - q * n;
        const double tq = zt[q];
        const double dp = num[(0 * (NPOLE + 1) + q) * n + i] / den[(0 * (NPOLE + 1) + q) * n + i];
        const double dm = num[(1 * (NPOLE + 1) + q) * n + i] / den[(1 * (NPOLE + 1) + q) * n + i];
        const double ai = dd[i] + shift0 + tq;
        gd[idx] = (float)(1.0 / (dp + dm - ai));
        float c = 0.f;
        if (i < n - 1) {
            const double dm1 = num[(1 * (NPOLE + 1) + q) * n + i + 1] / den[(1 * (NPOLE + 1) + q) * n + i + 1];
            c = (float)(-ee[i] / dm1);
        }
        cq[q * (n + 1) + i] = c;
    }
    __syncthreads();
    // P32_q[l] = prod_{s=l}^{l+31} c_q[s]: the factor that advances an entry 32 columns along a row.
    // It overwrites the (now dead) pivot scratch at the start of the region.
    float* p32 = reinterpret_cast<float*>(num);  // [16][n]
    for (int idx = tid; idx < NPOLE * n; idx += TT) {
        const int q = idx / n, l = idx - q * n;
        float pr = 0.f;
        if (l + 32 <= n - 1) {
            pr = 1.f;
#pragma unroll 8
            for (int s2 = 0; s2 < 32; ++s2) pr *= cq[q * (n + 1) + l + s2];
        }
        p32[idx] = pr;
    }
    __syncthreads();
    COVO_STAMP(a, 13);
    // F = exp(log_const / 2) * sum_q w_q (T_s + t_q)^-1, upper triangle only (row i, columns l >= i): one warp
    // per row, lanes over the columns (coalesced stores).  Lane j starts at column i + j with the product of
    // its first j factors (a 5-step warp scan per pole), then hops 32 columns at a time with P32.
    {
        const double logdet = sc[4];
        // controllers/covo.py:124-128: log_const = (n * 2 log(sigma) * 2 + sum log o) / n
        const double log_const = (4.0 * n * log((double)a.sample_sigma) + logdet) / (double)n;
        const float cscale = (float)exp(0.5 * log_const);
        float* Fg = a.F + (long long)env * n * n;
        for (int i = part + nparts * warp; i < n; i += nparts * (TT / 32)) {  // rows interleaved over the CTAs
            float val[NPOLE];
#pragma unroll
            for (int q = 0; q < NPOLE; ++q) {
                // exclusive prefix product of c_q[i .. i+31] across the lanes
                float c = (i + lane < n) ? cq[q * (n + 1) + i + lane] : 0.f;
                float pr = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float up = __shfl_up_sync(0xffffffffu, pr, o);
                    if (lane >= o) pr *= up;
                }
                float ex = __shfl_up_sync(0xffffffffu, pr, 1);
                if (lane == 0) ex = 1.f;
                val[q] = gd[q * n + i] * ex * ((float)zw[q] * cscale);
            }
            for (int l = i + lane; l < n; l += 32) {
                float f = 0.f;
#pragma unroll
                for (int q = 0; q < NPOLE; ++q) f += val[q];
                Fg[(long long)i * n + l] = f;
                Fg[(long long)l * n + i] = f;  // full symmetric storage (the sandwich kernel reads F by columns)
                if (l + 32 < n) {
#pragma unroll
                    for (int q = 0; q < NPOLE; ++q) val[q] *= p32[q * n + l];
                }
            }
        }
        if (tid == 0 && part == 0) {
            double* dg = a.diag + (long long)env * 4 * n;
            dg[2 * n + 0] = lam_min;
            dg[2 * n + 1] = sc[2];
            dg[2 * n + 2] = sc[3];
            dg[2 * n + 3] = logdet;
            dg[2 * n + 4] = (double)lad;
        }
    }
    __syncthreads();
    COVO_STAMP(a, 14);
}

// ---------------------------------------------------------------------------------------------
// E1b: Q^T = H_{n-3} ... H_1 H_0 from the stored reflectors.  P <- H_m P = P - tau_m v_m (v_m^T P) acts on every
// COLUMN of P independently: one warp per column pair (rows over the lanes, float2 = two columns, as in E1), no
// communication between warps at all.  The kernel depends only on E1, so the host runs it on a side stream next
// to E2 and joins before the sandwich kernel: an explicit Q for free.
// ---------------------------------------------------------------------------------------------
constexpr int kQaccWarpsSingle = 4, kQaccWarpsBatch = 8, kQaccChunk = 32;  // batches: 16 columns per CTA share one streamed copy of the reflectors

template <int kQaccWarps>
__global__ void __launch_bounds__(kQaccWarps * 32) qacc_kernel(const SigmaArgs a) {
    extern __shared__ __align__(16) float qsmf[];  // [2][kQaccChunk][n] reflector chunks (cp.async), then tau[256]
    const int n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, env = blockIdx.y;
    const float* Vg = a.Vh + (long long)env * n * n;
    const float* taug = a.tau + (long long)env * n;
    float* Qg = a.Qt + (long long)env * n * n;
    float* taus = qsmf + 2 * kQaccChunk * n;
    const int j0 = 2 * (blockIdx.x * kQaccWarps + warp);  // this warp's column pair
    const int nref = n - 2, nch = (nref + kQaccChunk - 1) / kQaccChunk, nv4 = n >> 2;
    auto prefetch = [&](int c) {
        const int m0 = c * kQaccChunk, rows = min(kQaccChunk, nref - m0);
        float* dst = qsmf + (c & 1) * kQaccChunk * n;
        for (int idx = tid; idx < rows * nv4; idx += kQaccWarps * 32) {
            unsigned d = (unsigned)__cvta_generic_to_shared(dst + 4 * idx);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(Vg + (long long)m0 * n + 4 * idx) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(0);
    for (int i = tid; i < n; i += kQaccWarps * 32) taus[i] = taug[i];
    float2 P2[RS];
#pragma unroll
    for (int r = 0; r < RS; ++r) {
        const int i = lane + 32 * r;
        P2[r] = make_float2(i == j0 ? 1.f : 0.f, i == j0 + 1 ? 1.f : 0.f);
    }
    for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) {
            prefetch(c + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int m0 = c * kQaccChunk, rows = min(kQaccChunk, nref - m0);
        const float* buf = qsmf + (c & 1) * kQaccChunk * n;
        if (j0 < n) {
            for (int mm = 0; mm < rows; ++mm) {
                const float tau = taus[m0 + mm];
                float v[RS];
#pragma unroll
                for (int r = 0; r < RS; ++r) v[r] = (lane + 32 * r < n) ? buf[mm * n + lane + 32 * r] : 0.f;
                float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
                for (int r = 0; r < RS; r += 2) sa = __ffma2_rn(P2[r], make_float2(v[r], v[r]), sa);
#pragma unroll
                for (int r = 1; r < RS; r += 2) sb = __ffma2_rn(P2[r], make_float2(v[r], v[r]), sb);
                float2 z = __fadd2_rn(sa, sb);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    z.x += __shfl_xor_sync(0xffffffffu, z.x, o);
                    z.y += __shfl_xor_sync(0xffffffffu, z.y, o);
                }
#pragma unroll
                for (int r = 0; r < RS; ++r) {
                    const float tv = -tau * v[r];
                    P2[r] = __ffma2_rn(make_float2(tv, tv), z, P2[r]);
                }
            }
        }
        __syncthreads();  // the buffer is refilled two chunks later
    }
    if (j0 < n) {
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int i = lane + 32 * r;
            if (i < n) *reinterpret_cast<float2*>(Qg + (long long)i * n + j0) = P2[r];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// E3: Sigma = Q F Q^T = P^T F P with P = Q^T from E1b and the symmetric F from E2, one launch.  CTA J owns JW
// columns of the result: Z = F P[:, J] (thread k = row k of Z, F streamed through shared memory row by row --
// F is symmetric, so "row l, element k" is read with consecutive k), then Sigma[i, J] = sum_k P[k][i] Z[k][:]
// (thread i, P streamed the same way), for the rows i on or below the diagonal; the tile is stored together with
// its mirror image, so Sigma is exactly symmetric by construction (controllers/covo.py:132 symmetrises its product
// the same way up to rounding) and the Cholesky kernel can skip its symmetrisation pass.
// JW = 8 (25 CTAs per matrix) for latency, 16 for batches (half the L2 -> SM streaming per matrix).
// ---------------------------------------------------------------------------------------------
constexpr int kSwThreads = 256, kSwChunk = 32;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__host__ __device__ inline size_t sandwich_smem_bytes(int n, int jw) {
    return (size_t)n * (2 * jw + 2 * kSwChunk) * sizeof(float);
}

template <int JW>
__global__ void __launch_bounds__(kSwThreads) sandwich_kernel(const float* __restrict__ Qt, const float* __restrict__ F,
                                                             float* __restrict__ cov, int n) {
    extern __shared__ __align__(16) float ssm[];
    float* Pj = ssm;                 // [n][JW]  P[:, J]
    float* Zs = Pj + n * JW;         // [n][JW]  Z = F P[:, J]
    float* Rs = Zs + n * JW;         // [2][32][n]  row chunks of F, then of P, double-buffered (cp.async)
    const int env = blockIdx.y, tid = threadIdx.x;
    const int J0 = blockIdx.x * JW;
    Qt += (long long)env * n * n;
    F += (long long)env * n * n;
    cov += (long long)env * n * n;
    const int nv4 = n >> 2;
    const int nch = (n + kSwChunk - 1) / kSwChunk;
    auto prefetch = [&](const float* M, int c, int slot) {
        const int l0 = c * kSwChunk, rows = min(kSwChunk, n - l0);
        float* dst = Rs + slot * kSwChunk * n;
        for (int idx = tid; idx < rows * nv4; idx += kSwThreads) cp_async16(dst + 4 * idx, M + (long long)l0 * n + 4 * idx);
        cp_async_commit();
    };
    // P[:, J] (16-byte pieces; pieces beyond column n are zero-filled)
    for (int idx = tid; idx < n * (JW / 4); idx += kSwThreads) {
        const int l = idx / (JW / 4), q4 = idx - l * (JW / 4);
        if (J0 + 4 * q4 < n) cp_async16(Pj + l * JW + 4 * q4, Qt + (long long)l * n + J0 + 4 * q4);
        else *reinterpret_cast<float4*>(Pj + l * JW + 4 * q4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    prefetch(F, 0, 0);
    float acc[JW];
    // ---- pass 1: Z[k][:] = sum_l F[l][k] P[l][J], thread k; pass 2: Sigma[i][J] = sum_k P[k][i] Z[k][:], thread i
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const float* M = pass ? Qt : F;
        const float* B = pass ? Zs : Pj;
#pragma unroll
        for (int jj = 0; jj < JW; ++jj) acc[jj] = 0.f;
        for (int c = 0; c < nch; ++c) {
            // chunk c sits in slot (pass * nch + c) & 1; the next chunk (of this pass or the first of the next) is prefetched
            const int g = pass * nch + c;
            if (c + 1 < nch) {
                prefetch(M, c + 1, (g + 1) & 1);
                cp_async_wait<1>();
            } else if (pass == 0) {
                prefetch(Qt, 0, (g + 1) & 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            if (tid < n) {
                const int l0 = c * kSwChunk, rows = min(kSwChunk, n - l0);
                const float* fb = Rs + (g & 1) * kSwChunk * n + tid;
#pragma unroll 4
                for (int l = 0; l < rows; ++l) {
                    const float f = fb[l * n];
                    const float4* pr = reinterpret_cast<const float4*>(B + (l0 + l) * JW);
#pragma unroll
                    for (int q4 = 0; q4 < JW / 4; ++q4) {
                        const float4 pv = pr[q4];
                        acc[4 * q4 + 0] = fmaf(f, pv.x, acc[4 * q4 + 0]);
                        acc[4 * q4 + 1] = fmaf(f, pv.y, acc[4 * q4 + 1]);
                        acc[4 * q4 + 2] = fmaf(f, pv.z, acc[4 * q4 + 2]);
                        acc[4 * q4 + 3] = fmaf(f, pv.w, acc[4 * q4 + 3]);
                    }
                }
            }
            __syncthreads();  // the slot is refilled two chunks later
        }
        if (pass == 0) {
            if (tid < n) {
#pragma unroll
                for (int q4 = 0; q4 < JW / 4; ++q4)
                    reinterpret_cast<float4*>(Zs + tid * JW)[q4] = make_float4(acc[4 * q4], acc[4 * q4 + 1], acc[4 * q4 + 2], acc[4 * q4 + 3]);
            }
            __syncthreads();
        }
    }
    // lower triangle + mirror image
    if (tid < n) {
        const int i = tid;
#pragma unroll
        for (int jj = 0; jj < JW; ++jj) {
            const int j = J0 + jj;
            if (j <= i && j < n) {
                cov[(long long)i * n + j] = acc[jj];
                if (j < i) cov[(long long)j * n + i] = acc[jj];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// E4: Cholesky (blocked right-looking, NB = 8), one CTA per matrix.
//   per panel: (a) one warp factors the 8x8 diagonal block in registers (shuffles), (b) one thread per row
//   solves its 8 panel entries and also stores them transposed (Lp[c][row]) so that (c) the rank-8 trailing
//   update reads both operands as conflict-free float4 and runs on all lower-triangle 4x4 tiles at once.
// ---------------------------------------------------------------------------------------------
constexpr int TC = 1024;
__global__ void __launch_bounds__(TC, 1) cholesky_kernel(const SigmaArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int n = a.n, n_pad = a.n_pad, tid = threadIdx.x, env = blockIdx.x, lane = tid & 31, warp = tid >> 5;
    float* As = reinterpret_cast<float*>(smraw);  // [n][n]
    float* Lp = As + n * n;                       // panel buffers, see below
    float* covg = a.cov + (long long)env * n * n;
    // pipeline mode: let the dependent grid (the rollout kernel, launched with programmatic stream serialisation) start now;
    // it synchronises with this kernel through a.progress only
    if (a.progress) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    COVO_STAMP(a, 23);
    // (a_cov + a_cov.T)/2, controllers/covo.py:132, in ONE pass: the row part of a float4 is read coalesced, its
    // four transposed partners straight from L2; all loads of a batch are in flight together.
    if (a.cov_symmetric) {
        // written by the sandwich kernel: exactly symmetric by construction -> plain coalesced load
        const int nv4 = (n * n) >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(covg);
        constexpr int kB = 10;
        for (int base = 0; base < nv4; base += kB * TC) {
            float4 rv[kB];
#pragma unroll
            for (int k = 0; k < kB; ++k) {
                const int idx = base + k * TC + tid;
                if (idx < nv4) rv[k] = g4[idx];
            }
#pragma unroll
            for (int k = 0; k < kB; ++k) {
                const int idx = base + k * TC + tid;
                if (idx < nv4) reinterpret_cast<float4*>(As)[idx] = rv[k];
            }
        }
    } else {
        const int nv4 = (n * n) >> 2, nq = n >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(covg);
        constexpr int kBatch = 5;
        for (int base = 0; base < nv4; base += kBatch * TC) {
            float4 rv[kBatch], cv[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const int idx = base + k * TC + tid;
                if (idx < nv4) {
                    const int i = idx / nq, j = 4 * (idx - i * nq);
                    rv[k] = g4[idx];
                    cv[k].x = covg[(long long)j * n + i];
                    cv[k].y = covg[(long long)(j + 1) * n + i];
                    cv[k].z = covg[(long long)(j + 2) * n + i];
                    cv[k].w = covg[(long long)(j + 3) * n + i];
                }
            }
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const int idx = base + k * TC + tid;
                if (idx < nv4) {
                    const float4 sv = make_float4(0.5f * (rv[k].x + cv[k].x), 0.5f * (rv[k].y + cv[k].y),
                                                  0.5f * (rv[k].z + cv[k].z), 0.5f * (rv[k].w + cv[k].w));
                    reinterpret_cast<float4*>(As)[idx] = sv;
                }
            }
        }
        __syncthreads();  // every transposed partner has been read from HBM / L2 before anything is overwritten
        for (int idx = tid; idx < nv4; idx += TC) reinterpret_cast<float4*>(covg)[idx] = reinterpret_cast<const float4*>(As)[idx];
    }
    __syncthreads();
    COVO_STAMP(a, 24);
    // Right-looking blocked factorisation (NB = 8) with LOOK-AHEAD: while the other warps apply panel p to the
    // trailing matrix, the eight "panel warps" first update the 8 columns of panel p+1, then factor its diagonal
    // block (privately, in registers: no shuffles, no broadcast) and solve its rows, so the serial part of a step
    // hides behind the rank-8 update.  The panel chain is bound by its instruction count and by the shared-memory
    // hand-overs between its stages (NB = 4 was measured slower: twice the hand-overs); it owns the highest warp
    // ids because the warp scheduler favours them.
    constexpr int kPanelThreads = 256;
    float* LpA = Lp;                 // [8][n_pad] panel, transposed (double-buffered)
    float* LpB = Lp + 8 * n_pad;
    float* Lp8 = LpB + 8 * n_pad;    // [8][8] factored diagonal block, parked until its rows are no longer being read
    const int pt = tid - (TC - kPanelThreads);
    auto factor_panel = [&](int jb, int nb, float* LpOut) {
        const int nrows = n - jb - nb;
        if (pt < max(nrows, 1)) {
            float d[8][8], linv[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 p0 = (r < nb) ? *reinterpret_cast<const float4*>(As + (jb + r) * n + jb) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 p1 = (r < nb && nb == 8) ? *reinterpret_cast<const float4*>(As + (jb + r) * n + jb + 4)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                d[r][0] = p0.x; d[r][1] = p0.y; d[r][2] = p0.z; d[r][3] = p0.w;
                d[r][4] = p1.x; d[r][5] = p1.y; d[r][6] = p1.z; d[r][7] = p1.w;
                if (r >= nb) d[r][r] = 1.f;
            }
            bool bad = false;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float dcc = d[c][c];
                if (!(dcc > 0.f)) {
                    bad = true;
                    dcc = 1e-30f;
                }
                const float rinv = rsqrt_newton(dcc);
                linv[c] = rinv;
                d[c][c] = dcc * rinv;
#pragma unroll
                for (int r = c + 1; r < 8; ++r) d[r][c] *= rinv;
#pragma unroll
                for (int c2 = c + 1; c2 < 8; ++c2)
#pragma unroll
                    for (int r = c2; r < 8; ++r) d[r][c2] = fmaf(-d[r][c], d[c2][c], d[r][c2]);
            }
            if (pt < nrows) {
                const int i = jb + nb + pt;
                float x[8];
                const float4 p0 = *reinterpret_cast<const float4*>(As + i * n + jb);
                x[0] = p0.x; x[1] = p0.y; x[2] = p0.z; x[3] = p0.w;
                if (nb == 8) {
                    const float4 p1 = *reinterpret_cast<const float4*>(As + i * n + jb + 4);
                    x[4] = p1.x; x[5] = p1.y; x[6] = p1.z; x[7] = p1.w;
                } else {
                    x[4] = x[5] = x[6] = x[7] = 0.f;
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float sx = x[c];
#pragma unroll
                    for (int c2 = 0; c2 < c; ++c2) sx = fmaf(-x[c2], d[c][c2], sx);
                    x[c] = sx * linv[c];
                }
                *reinterpret_cast<float4*>(As + i * n + jb) = make_float4(x[0], x[1], x[2], x[3]);
                if (nb == 8) *reinterpret_cast<float4*>(As + i * n + jb + 4) = make_float4(x[4], x[5], x[6], x[7]);
#pragma unroll
                for (int c = 0; c < 8; ++c) LpOut[c * n_pad + i] = x[c];
            }
            if (pt == 0) {
                if (bad) a.status[env] = 2;
                // the block rows are still being read by the other row threads: park the factor, write back later
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) Lp8[r * 8 + c] = (c <= r) ? d[r][c] : 0.f;
            }
        }
    };
    // prologue: panel 0
    if (pt >= 0) factor_panel(0, min(8, n), LpA);
    __syncthreads();
    for (int jb = 0, it = 0; jb < n; jb += 8, ++it) {
        const int nb = min(8, n - jb);
        float* LpCur = (it & 1) ? LpB : LpA;
        float* LpNext = (it & 1) ? LpA : LpB;
        if (jb == 8) COVO_STAMP(a, 26);
        // the factored diagonal block of panel jb (parked by panel thread 0) goes back into As
        if (pt >= 0 && pt < 64) {
            const int r = pt >> 3, c = pt & 7;
            if (r < nb && c < nb) As[(jb + r) * n + jb + c] = Lp8[pt];
            if (a.progress && c < nb) {  // rows jb .. jb+7 of the packed columns jb + c (zeros above the diagonal)
                a.Lt[(long long)env * a.lt_stride + lt_col_offset(jb + c, n_pad) + r] = (r < nb) ? Lp8[pt] : 0.f;
            }
        }
        if (a.progress && pt < 0) {
            // columns jb .. jb+nb-1 are final: rows below the block come from the transposed panel, rows >= n are padding
            float* Ltg = a.Lt + (long long)env * a.lt_stride;
            const int len = n_pad - jb - 8;
            for (int q = tid; q < nb * len; q += TC - kPanelThreads) {
                const int c = q / len, i = jb + 8 + (q - c * len);
                Ltg[lt_col_offset(jb + c, n_pad) + (i - jb)] = (i < n) ? LpCur[c * n_pad + i] : 0.f;
            }
        }  // (published after the barrier at the end of the iteration)
        const int r0 = jb + nb;           // first row / column of the trailing matrix
        if (r0 >= n) break;
        const int nbn = min(8, n - r0);   // width of the next panel
        if (pt >= 0) {
            // (1) rank-8 update of the strip: rows i >= r0, columns r0 .. r0 + nbn - 1
            const int i = r0 + pt;
            if (i < n) {
                float x[8];
                const float4 p0 = *reinterpret_cast<const float4*>(As + i * n + r0);
                x[0] = p0.x; x[1] = p0.y; x[2] = p0.z; x[3] = p0.w;
                if (nbn == 8) {
                    const float4 p1 = *reinterpret_cast<const float4*>(As + i * n + r0 + 4);
                    x[4] = p1.x; x[5] = p1.y; x[6] = p1.z; x[7] = p1.w;
                } else {
                    x[4] = x[5] = x[6] = x[7] = 0.f;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float li = LpCur[k * n_pad + i];
                    const float4 l0 = *reinterpret_cast<const float4*>(LpCur + k * n_pad + r0);
                    x[0] = fmaf(-li, l0.x, x[0]); x[1] = fmaf(-li, l0.y, x[1]);
                    x[2] = fmaf(-li, l0.z, x[2]); x[3] = fmaf(-li, l0.w, x[3]);
                    if (nbn == 8) {
                        const float4 l1 = *reinterpret_cast<const float4*>(LpCur + k * n_pad + r0 + 4);
                        x[4] = fmaf(-li, l1.x, x[4]); x[5] = fmaf(-li, l1.y, x[5]);
                        x[6] = fmaf(-li, l1.z, x[6]); x[7] = fmaf(-li, l1.w, x[7]);
                    }
                }
                *reinterpret_cast<float4*>(As + i * n + r0) = make_float4(x[0], x[1], x[2], x[3]);
                if (nbn == 8) *reinterpret_cast<float4*>(As + i * n + r0 + 4) = make_float4(x[4], x[5], x[6], x[7]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");  // panel warps only
            // (2) factor panel p+1 while the other warps are busy with the trailing update
            factor_panel(r0, nbn, LpNext);
        } else {
            // (c) trailing update on the lower-triangle 4x4 tiles right of the strip: 64 packed FFMA2 per tile
            const int c0 = r0 + nbn;
            const int T = (n - c0) >> 2;  // n, r0, nbn are multiples of 4
            const int ntiles = T * (T + 1) / 2;
            for (int q = tid; q < ntiles; q += TC - kPanelThreads) {
                int ti = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
                while (ti * (ti + 1) / 2 > q) --ti;
                while ((ti + 1) * (ti + 2) / 2 <= q) ++ti;
                const int tk = q - ti * (ti + 1) / 2;
                const int i = c0 + 4 * ti, kk = c0 + 4 * tk;
                float2 o[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 av = *reinterpret_cast<const float4*>(As + (i + r) * n + kk);
                    o[r][0] = make_float2(av.x, av.y);
                    o[r][1] = make_float2(av.z, av.w);
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 li = *reinterpret_cast<const float4*>(LpCur + c * n_pad + i);
                    const float4 lk = *reinterpret_cast<const float4*>(LpCur + c * n_pad + kk);
                    const float lir[4] = {-li.x, -li.y, -li.z, -li.w};
                    const float2 lk0 = make_float2(lk.x, lk.y), lk1 = make_float2(lk.z, lk.w);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float2 l2 = make_float2(lir[r], lir[r]);
                        o[r][0] = __ffma2_rn(l2, lk0, o[r][0]);
                        o[r][1] = __ffma2_rn(l2, lk1, o[r][1]);
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    *reinterpret_cast<float4*>(As + (i + r) * n + kk) = make_float4(o[r][0].x, o[r][0].y, o[r][1].x, o[r][1].y);
            }
        }
        __syncthreads();
        // one release by one thread publishes the block: the CTA barrier orders every thread's stores before it and a
        // release is cumulative (PTX memory model).  The publisher is the last trailing-update thread, idle in late panels.
        if (a.progress && tid == TC - kPanelThreads - 32)
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.progress + env), "r"((int)((unsigned)a.epoch + (unsigned)(it + 1))) : "memory");
    }
    __syncthreads();
    if (a.progress && tid == 0)  // the last block (its iteration left the loop before the barrier)
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.progress + env), "r"((int)((unsigned)a.epoch + (unsigned)((n + 7) >> 3))) : "memory");
    COVO_STAMP(a, 25);
    // outputs: row-major L (upper part zeroed) and the packed k-major factor
    if (a.L) {
        float* Lg = a.L + (long long)env * n * n;
        for (int i = warp; i < n; i += TC / 32)
            for (int j = lane; j < n; j += 32) Lg[i * n + j] = (j <= i) ? As[i * n + j] : 0.f;
    }
    if (a.Lt && !a.progress) {
        float* Ltg = a.Lt + (long long)env * a.lt_stride;
        for (int k = warp; k < n; k += TC / 32) {
            const int rs = k & ~7;
            const int off = lt_col_offset(k, n_pad);
            for (int r = rs + lane; r < n_pad; r += 32) Ltg[off + (r - rs)] = (r >= k && r < n) ? As[r * n + k] : 0.f;
        }
    }
    __syncthreads();
    COVO_STAMP(a, 27);
}

// ---------------------------------------------------------------------------------------------
static size_t trifunc_smem(int n) {
    return (size_t)trifunc_region_floats(n) * 4 + 256 * 8 * 3 + 16 * 8 + 64 * 8 + 64 * 4;
}
static size_t chol_smem(int n) { return (size_t)n * n * 4 + (size_t)(2 * 8 * round_up8(n) + 64) * 4; }

template <int C2, int NC>
static cudaError_t launch_e1(const SigmaArgs& a, int n_env, cudaStream_t st) {
    static size_t conf[32] = {};
    const size_t smem = e1_smem_bytes(a.n, C2);
    cudaError_t e = ensure_smem_attr(tridiag_reg_kernel<C2, NC>, smem, conf);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_env * NC);
    cfg.blockDim = dim3(ET);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tridiag_reg_kernel<C2, NC>, a);
}

// Cluster width of E1 (CTAs per matrix).  Few matrices: spread each one over 4 SMs (latency); many matrices
// (batched environments, the offline schedule): 2 SMs each, which still fills the machine.
static int e1_cluster_override() {  // COVO_E1_CLUSTER = 1 | 2 | 4 | 8 pins the cluster width (tuning, tests)
    const char* e = getenv("COVO_E1_CLUSTER");
    const int v = e ? atoi(e) : 0;
    return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 0;
}

cudaError_t launch_tridiag(const SigmaArgs& a, int n_env, cudaStream_t st) {
    if (a.n > kSigmaMaxN || (a.n & 3)) return cudaErrorInvalidValue;
    cudaError_t e;
    int nc = e1_cluster_override() ? e1_cluster_override() : (n_env <= 18 ? 8 : (n_env <= 37 ? 4 : 2));
    if (a.n <= 64) nc = min(nc, 2);
    else if (a.n <= 128) nc = min(nc, 4);
    // column-pair slots per warp: n <= 16 NC C2
    if (nc == 1) {  // one SM per matrix (only through COVO_E1_CLUSTER=1: register spills make it slower than 2 CTAs)
        switch ((a.n + 15) / 16) {
            case 1: case 2: case 3: case 4: e = launch_e1<4, 1>(a, n_env, st); break;
            case 5: case 6: case 7: case 8: e = launch_e1<8, 1>(a, n_env, st); break;
            case 9: case 10: e = launch_e1<10, 1>(a, n_env, st); break;
            case 11: case 12: case 13: e = launch_e1<13, 1>(a, n_env, st); break;
            default: e = launch_e1<14, 1>(a, n_env, st); break;
        }
    } else if (nc >= 8) {
        e = (a.n <= 128) ? launch_e1<1, 8>(a, n_env, st) : launch_e1<2, 8>(a, n_env, st);
    } else if (nc >= 4) {
        switch ((a.n + 63) / 64) {
            case 1: e = launch_e1<1, 4>(a, n_env, st); break;
            case 2: e = launch_e1<2, 4>(a, n_env, st); break;
            case 3: e = launch_e1<3, 4>(a, n_env, st); break;
            default: e = launch_e1<4, 4>(a, n_env, st); break;
        }
    } else {  // 2 CTAs per matrix
        switch ((a.n + 31) / 32) {
            case 1: e = launch_e1<1, 2>(a, n_env, st); break;
            case 2: e = launch_e1<2, 2>(a, n_env, st); break;
            case 3: case 4: e = launch_e1<4, 2>(a, n_env, st); break;
            case 5: case 6: e = launch_e1<6, 2>(a, n_env, st); break;
            default: e = launch_e1<7, 2>(a, n_env, st); break;
        }
    }
    return e;
}

cudaError_t launch_qacc(const SigmaArgs& a, int n_env, cudaStream_t st) {
    if (a.n > kSigmaMaxN || (a.n & 3)) return cudaErrorInvalidValue;
    const size_t smem = ((size_t)2 * kQaccChunk * a.n + 256) * sizeof(float);
    static size_t conf[32] = {}, confb[32] = {};
    cudaError_t e;
    if (n_env > 18) {
        e = ensure_smem_attr(qacc_kernel<kQaccWarpsBatch>, smem, confb);
        if (e != cudaSuccess) return e;
        qacc_kernel<kQaccWarpsBatch><<<dim3((a.n / 2 + kQaccWarpsBatch - 1) / kQaccWarpsBatch, n_env), kQaccWarpsBatch * 32, smem, st>>>(a);
    } else {
        e = ensure_smem_attr(qacc_kernel<kQaccWarpsSingle>, smem, conf);
        if (e != cudaSuccess) return e;
        qacc_kernel<kQaccWarpsSingle><<<dim3((a.n / 2 + kQaccWarpsSingle - 1) / kQaccWarpsSingle, n_env), kQaccWarpsSingle * 32, smem, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_trifunc(const SigmaArgs& a, int n_env, cudaStream_t st) {
    if (a.n > kSigmaMaxN || (a.n & 3)) return cudaErrorInvalidValue;
    cudaError_t e;
    static size_t conf[32] = {};
    size_t smem = trifunc_smem(a.n);
    e = ensure_smem_attr(sigma_trifunc_kernel, smem, conf);
    if (e != cudaSuccess) return e;
    // CTAs per matrix: the scalar stages are repeated by each of them, the rows of F are shared out
    const int nb = (n_env <= 18) ? 8 : (n_env <= 74 ? 2 : 1);
    SigmaArgs a2 = a;
    {
        const char* ev = getenv("COVO_E2_POINTS");  // tuning: trial shifts per multisection round
        const int v = ev ? atoi(ev) : 0;
        a2.e2_points = (v == 8 || v == 16 || v == 32) ? v : 16;
    }
    sigma_trifunc_kernel<<<dim3(nb, n_env), TT, smem, st>>>(a2);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// E3 for BATCHES on the tensor cores: Sigma[:, J] = P^T (F P[:, J]) for a 64-column tile J per CTA, both products as
// mma.sync m16n8k8 TF32 with the operands split in two TF32 words each (x = hi + lo; hi*hi + hi*lo + lo*hi: the 3xTF32 scheme,
// float32-grade products, float32 accumulation).  A CTA streams F (pass 1, 32-column chunks) and P (pass 2, 32-row chunks) through
// shared memory once -- four CTAs per matrix instead of thirteen, 16 K MACs per streamed element instead of 16 -- and keeps P[:, J]
// and Z = F P[:, J] resident.  Only the tiles on or below the diagonal are formed in pass 2; every entry is stored together with its
// mirror image, so Sigma stays exactly symmetric as with the SIMT kernel.  512 matrices of order 200: 1.78 ms -> see DESIGN.md.
// ---------------------------------------------------------------------------------------------
constexpr int kTcThreads = 256, kTcJW = 64, kTcNP = 224, kTcBS = 72 /* row pitch of Pj / Zs */, kTcAS1 = 36 /* F chunk pitch */, kTcAS2 = 216 /* P chunk pitch */;
constexpr int kTcChunkFloats = kTcNP * kTcAS1;  // 8064 >= 32 * 216
constexpr size_t kTcSmemBytes = (size_t)(2 * kTcNP * kTcBS + 2 * kTcChunkFloats) * sizeof(float);

__device__ __forceinline__ void tf32_split(float x, unsigned& hi, unsigned& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(kTcThreads, 1) sandwich_tc_kernel(const float* __restrict__ Qt, const float* __restrict__ F,
                                                                     float* __restrict__ cov, int n) {
    extern __shared__ __align__(16) float ssm[];
    float* Pj = ssm;                      // [224][72]  P[:, J], zero beyond n
    float* Zs = Pj + kTcNP * kTcBS;       // [224][72]  Z = F P[:, J]
    float* Ch = Zs + kTcNP * kTcBS;       // [2][8064]  chunk buffers
    const int env = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int J0 = blockIdx.x * kTcJW;
    Qt += (long long)env * n * n;
    F += (long long)env * n * n;
    cov += (long long)env * n * n;
    const int MT = (n + 15) >> 4, nch = (n + 31) >> 5;
    // P[:, J]: rows l < n, columns J0 .. J0 + 63 (n is a multiple of 4: 16-byte pieces are all-in or all-out); zero elsewhere
    for (int idx = tid; idx < kTcNP * (kTcJW / 4); idx += kTcThreads) {
        const int l = idx >> 4, q4 = idx & 15;
        float* dst = Pj + l * kTcBS + 4 * q4;
        if (l < n && J0 + 4 * q4 < n) cp_async16(dst, Qt + (long long)l * n + J0 + 4 * q4);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // pass 1 chunk c: Ch[m][0..31] = F[m][32 c ..], m < 224 (zero outside the matrix); pass 2 chunk c: Ch[r][i] = P[32 c + r][i], i < 216
    auto prefetch1 = [&](int c, int slot) {
        float* dst = Ch + slot * kTcChunkFloats;
        const int l0 = 32 * c;
        for (int idx = tid; idx < kTcNP * 8; idx += kTcThreads) {
            const int m = idx >> 3, q4 = idx & 7;
            float* d = dst + m * kTcAS1 + 4 * q4;
            if (m < n && l0 + 4 * q4 < n) cp_async16(d, F + (long long)m * n + l0 + 4 * q4);
            else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
    };
    auto prefetch2 = [&](int c, int slot) {
        float* dst = Ch + slot * kTcChunkFloats;
        const int k0 = 32 * c;
        for (int idx = tid; idx < 32 * (kTcAS2 / 4); idx += kTcThreads) {
            const int r = idx / (kTcAS2 / 4), q4 = idx - r * (kTcAS2 / 4);
            float* d = dst + r * kTcAS2 + 4 * q4;
            if (k0 + r < n && 4 * q4 < n) cp_async16(d, Qt + (long long)(k0 + r) * n + 4 * q4);
            else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
    };
    prefetch1(0, 0);
    float acc[2][8][4];
    // ---------------- pass 1: Z = F P[:, J]; warp w owns the m-tiles w and w + 8 ----------------
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[a][nt][e] = 0.f;
    for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) prefetch1(c + 1, (c + 1) & 1);
        else prefetch2(0, (c + 1) & 1);
        cp_async_wait<1>();
        __syncthreads();
        const float* A = Ch + (c & 1) * kTcChunkFloats;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            unsigned ah[2][4], al[2][4];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int m0 = 16 * (warp + 8 * a);  // (m-tiles beyond MT read zero rows or rows of the padding: their results are never stored)
                const float* ap = A + (m0 + g) * kTcAS1 + 8 * ks + t;
                tf32_split(m0 < 16 * MT ? ap[0] : 0.f, ah[a][0], al[a][0]);
                tf32_split(m0 < 16 * MT ? ap[8 * kTcAS1] : 0.f, ah[a][1], al[a][1]);
                tf32_split(m0 < 16 * MT ? ap[4] : 0.f, ah[a][2], al[a][2]);
                tf32_split(m0 < 16 * MT ? ap[8 * kTcAS1 + 4] : 0.f, ah[a][3], al[a][3]);
            }
            const float* bp = Pj + (32 * c + 8 * ks + t) * kTcBS + g;
            unsigned bh[8][2], bl[8][2];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                tf32_split(bp[8 * nt], bh[nt][0], bl[nt][0]);
                tf32_split(bp[4 * kTcBS + 8 * nt], bh[nt][1], bl[nt][1]);
            }
            // the three terms of a tile go to the same accumulator: issued term by term over all sixteen tiles, so that consecutive
            // mma instructions are independent (tile by tile, three dependent mma in a row, the kernel ran at a tenth of this rate)
#pragma unroll
            for (int term = 0; term < 3; ++term)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        if (16 * (warp + 8 * a) >= 16 * MT) continue;  // warp-uniform
                        mma_tf32(acc[a][nt], term == 0 ? al[a] : ah[a], term == 1 ? bl[nt] : bh[nt]);
                    }
        }
        __syncthreads();  // the slot is refilled two chunks later
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int m0 = 16 * (warp + 8 * a);
        if (m0 >= kTcNP) continue;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const bool live = m0 < 16 * MT;
            *reinterpret_cast<float2*>(Zs + (m0 + g) * kTcBS + 8 * nt + 2 * t) = live ? make_float2(acc[a][nt][0], acc[a][nt][1]) : make_float2(0.f, 0.f);
            *reinterpret_cast<float2*>(Zs + (m0 + g + 8) * kTcBS + 8 * nt + 2 * t) = live ? make_float2(acc[a][nt][2], acc[a][nt][3]) : make_float2(0.f, 0.f);
        }
    }
    // ---------------- pass 2: Sigma[i][J] = sum_k P[k][i] Z[k][J] for the m-tiles on or below the diagonal block ----------------
    const int mt0 = J0 >> 4;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[a][nt][e] = 0.f;
    for (int c = 0; c < nch; ++c) {
        const int gidx = nch + c;
        if (c + 1 < nch) {
            prefetch2(c + 1, (gidx + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();  // (also orders the Zs stores of pass 1 before their first use)
        const float* A = Ch + (gidx & 1) * kTcChunkFloats;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            unsigned ah[2][4], al[2][4];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int mt = mt0 + warp + 8 * a;
                const bool live = mt < MT;
                const float* ap = A + (8 * ks + t) * kTcAS2 + 16 * mt + g;  // A[m][kk] = P[k0 + kk][m]
                tf32_split(live ? ap[0] : 0.f, ah[a][0], al[a][0]);
                tf32_split(live ? ap[8] : 0.f, ah[a][1], al[a][1]);
                tf32_split(live ? ap[4 * kTcAS2] : 0.f, ah[a][2], al[a][2]);
                tf32_split(live ? ap[4 * kTcAS2 + 8] : 0.f, ah[a][3], al[a][3]);
            }
            const float* bp = Zs + (32 * c + 8 * ks + t) * kTcBS + g;
            unsigned bh[8][2], bl[8][2];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                tf32_split(bp[8 * nt], bh[nt][0], bl[nt][0]);
                tf32_split(bp[4 * kTcBS + 8 * nt], bh[nt][1], bl[nt][1]);
            }
#pragma unroll
            for (int term = 0; term < 3; ++term)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        if (mt0 + warp + 8 * a >= MT) continue;  // warp-uniform
                        mma_tf32(acc[a][nt], term == 0 ? al[a] : ah[a], term == 1 ? bl[nt] : bh[nt]);
                    }
        }
        __syncthreads();
    }
    // lower triangle + mirror image
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int mt = mt0 + warp + 8 * a;
        if (mt >= MT) continue;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 16 * mt + g + ((e & 2) ? 8 : 0), j = J0 + 8 * nt + 2 * t + (e & 1);
                if (i < n && j <= i) {
                    cov[(long long)i * n + j] = acc[a][nt][e];
                    if (j < i) cov[(long long)j * n + i] = acc[a][nt][e];
                }
            }
    }
}

cudaError_t launch_sandwich(const SigmaArgs& a, int n_env, cudaStream_t st) {
    if (a.n > kSigmaMaxN || (a.n & 3)) return cudaErrorInvalidValue;
    cudaError_t e;
    static size_t conf8[32] = {}, conf16[32] = {};
    if (n_env <= 18) {
        const size_t bytes = sandwich_smem_bytes(a.n, 8);
        e = ensure_smem_attr(sandwich_kernel<8>, bytes, conf8);
        if (e != cudaSuccess) return e;
        sandwich_kernel<8><<<dim3((a.n + 7) / 8, n_env), kSwThreads, bytes, st>>>(a.Qt, a.F, a.cov, a.n);
    } else if (!(getenv("COVO_SANDWICH") && strncmp(getenv("COVO_SANDWICH"), "simt", 4) == 0)) {
        // batches: the two products on the tensor cores (3xTF32), four CTAs per matrix
        static size_t conftc[32] = {};
        e = ensure_smem_attr(sandwich_tc_kernel, kTcSmemBytes, conftc);
        if (e != cudaSuccess) return e;
        sandwich_tc_kernel<<<dim3((a.n + kTcJW - 1) / kTcJW, n_env), kTcThreads, kTcSmemBytes, st>>>(a.Qt, a.F, a.cov, a.n);
    } else {
        const size_t bytes = sandwich_smem_bytes(a.n, 16);
        e = ensure_smem_attr(sandwich_kernel<16>, bytes, conf16);
        if (e != cudaSuccess) return e;
        sandwich_kernel<16><<<dim3((a.n + 15) / 16, n_env), kSwThreads, bytes, st>>>(a.Qt, a.F, a.cov, a.n);
    }
    return cudaGetLastError();
}

cudaError_t launch_sigma(const SigmaArgs& a, int n_env, cudaStream_t st) {
    cudaError_t e = launch_tridiag(a, n_env, st);
    if (e == cudaSuccess) e = launch_qacc(a, n_env, st);
    if (e == cudaSuccess) e = launch_trifunc(a, n_env, st);
    if (e == cudaSuccess) e = launch_sandwich(a, n_env, st);
    return e;
}

cudaError_t launch_cholesky(const SigmaArgs& a, int n_env, cudaStream_t st) {
    if (a.n > kSigmaMaxN || (a.n & 3)) return cudaErrorInvalidValue;
    static size_t conf[32] = {};
    size_t smem = chol_smem(a.n);
    cudaError_t e = ensure_smem_attr(cholesky_kernel, smem, conf);
    if (e != cudaSuccess) return e;
    cholesky_kernel<<<n_env, TC, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace covo
