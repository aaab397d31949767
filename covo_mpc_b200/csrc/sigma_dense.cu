// EXPERIMENTAL (selected with COVO_SIGMA=dense; NOT the default, NOT yet run on hardware: written at the end of round 1
// after the GPU budget was spent -- see DESIGN.md section 9 and tools/study_dense_sigma.py for the numerical study behind it).
//
// optimize_sigma (controllers/covo.py:116-132) WITHOUT an eigen-decomposition.  The reference computes
//     Sigma = U diag(s) U^T,  s_k = exp(c/2) / sqrt(o_k),  o_k = lambda_k - lambda_min + 1e-2,  c = (4 n log sigma + sum log o_k) / n
// which is  Sigma = exp(c/2) * A^(-1/2)  with  A = (R + R^T)/2 - lambda_min I + 1e-2 I  and  sum log o_k = log det A.
// So only lambda_min, log det A and the matrix function A^(-1/2) are needed:
//   D1  lanczos_kernel          lambda_min / lambda_max of R by k <= 32 Lanczos steps (fp64 arithmetic on the fp32 matrix; the
//                               lowest eigenvalue is well separated, the Ritz value is exact to 1e-14 after ~24 steps) and the
//                               extreme eigenvalues of the k x k Lanczos matrix by 32-way multisection (Sturm counts)
//   D2  shifted_inverse_kernel  grid (16 + 1, E): CTA j factors A + t_j I (Cholesky in shared memory, the look-ahead scheme of
//                               E4), inverts the factor in place (one warp per column, upper triangle holds X^T) and forms
//                               w_j (A + t_j I)^-1 = w_j X^T X; the extra CTA factors A itself for log det A.
//                               x^(-1/2) ~ sum_j w_j / (x + t_j): Zolotarev's partial fractions on [1e-2, M], the ladder of E2
//   D3  combine_kernel          Sigma = exp(c/2) * sum_j w_j (A + t_j I)^-1, written symmetric
// The 198 dependent Householder steps of E1 (214 us) become ~32 dependent matrix-vector products and 17 independent
// factorisations.  Study (CPU emulation, fp32 solves): relative Frobenius error vs the float64 eigen-decomposition 1e-7 .. 4e-6 on
// the tracking / zigzag path for H = 8 .. 50; up to 1.6e-4 when cond(A) ~ 1e5 (hover, t = 0, H = 50), where fp32 LAPACK is no
// better.
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
constexpr int kLanczosMax = 32;
constexpr int kLanczosSteps = 24;  // steps of the cluster kernel (see lanczos_cluster_kernel)
constexpr int TL = 256;   // lanczos_kernel threads
constexpr int TD = 1024;  // shifted_inverse_kernel threads

__device__ __forceinline__ float rsqrt_newton_d(float x) {
    float r;
#if defined(COVO_CPU_EMU)
    r = 1.0f / sqrtf(x);
#else
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
#endif
    return r * fmaf(-0.5f * x * r, r, 1.5f);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the CTA (TL threads); red: [TL / 32] doubles.  Two barriers: the result may be consumed and red reused at once.
__device__ __forceinline__ double block_sum_d(double v, double* red) {
    v = warp_sum_d(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < TL / 32; ++w) s += red[w];
    __syncthreads();
    return s;
}

// number of eigenvalues of the symmetric tridiagonal (al[0..k), be[0..k-1)) below x  (Sturm sequence of the LDL^T pivots)
__device__ __forceinline__ int sturm_count(const double* al, const double* be, int k, double x) {
    int cnt = 0;
    double q = al[0] - x;
    if (q < 0.0) ++cnt;
    for (int i = 1; i < k; ++i) {
        if (fabs(q) < 1e-300) q = -1e-300;
        q = (al[i] - x) - be[i - 1] * be[i - 1] / q;
        if (q < 0.0) ++cnt;
    }
    return cnt;
}

// smallest x in [lo, hi] with sturm_count(x) >= target, by 32-way multisection (one warp, all lanes return the result)
__device__ __forceinline__ double warp_multisect(const double* al, const double* be, int k, double lo, double hi, int target) {
    const int lane = threadIdx.x & 31;
    for (int round = 0; round < 14; ++round) {
        const double step = (hi - lo) / 33.0;
        const double x = lo + step * (double)(lane + 1);
        const bool ge = sturm_count(al, be, k, x) >= target;
        const unsigned m = __ballot_sync(0xffffffffu, ge);
        if (m == 0u) {
            lo = lo + step * 32.0;  // the crossing is in the last sub-interval
        } else {
            const int first = __ffs(m) - 1;
            hi = lo + step * (double)(first + 1);
            lo = lo + step * (double)first;
        }
    }
    return 0.5 * (lo + hi);
}

}  // namespace

struct DenseArgs {
    int n, n_pad;
    float sample_sigma;
    const float* R;      // [E][n][n]
    double* scal;        // [E][4]: lambda_min, lambda_max (Ritz), log det A, unused
    float* Xbuf;         // [E][kZoloPoles][n][n]  lower triangles of w_j (A + t_j I)^-1
    float* cov;          // [E][n][n]
    const double* zolo;  // the ladder of sigma.cu: [kZoloLadder][2][kZoloPoles]
    int* status;         // [E]
    float* Asym = nullptr;       // optional [E][n][n]: (R + R^T)/2 written by the Lanczos kernel for the factorisation kernels
    long long* prof = nullptr;   // optional clock64() stamps (slots 48..)
};

// ---------------------------------------------------------------------------------------------------------------------------
// D1
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TL, 1) lanczos_kernel(const DenseArgs a) {
    COVO_DYN_SMEM(smraw);
    const int n = a.n, tid = threadIdx.x, env = blockIdx.x, ld = n + 1;
    double* v = reinterpret_cast<double*>(smraw);  // [n]
    double* vp = v + n;                            // [n]
    double* red = vp + n;                          // [8]
    double* al = red + 8;                          // [32]
    double* be = al + kLanczosMax;                 // [32]
    float* Rs = reinterpret_cast<float*>(be + kLanczosMax);  // [n][n + 1]: odd stride, a thread reads its own row conflict-free
    const float* Rg = a.R + (long long)env * n * n;
    for (int idx = tid; idx < n * n; idx += TL) {
        const int i = idx / n, j = idx - i * n;
        Rs[i * ld + j] = 0.5f * (Rg[idx] + Rg[(long long)j * n + i]);  // (R + R^T)/2, controllers/covo.py:117
    }
    // fixed start vector with a component along every eigenvector in practice (no symmetry of the problem is aligned with it)
    double w = 0.0, x0 = 0.0;
    if (tid < n) x0 = cos(0.37 * (double)tid + 0.1) + 0.01 * (double)tid / (double)n;
    const double nrm0 = sqrt(block_sum_d(x0 * x0, red));  // (also orders the Rs stores before the first product)
    if (tid < n) {
        v[tid] = x0 / nrm0;
        vp[tid] = 0.0;
    }
    __syncthreads();
    const int k_max = min(kLanczosMax, n);
    int k = 0;
    double beta = 0.0;
    for (int it = 0; it < k_max; ++it) {
        w = 0.0;
        if (tid < n) {
            const float* row = Rs + tid * ld;
            double acc0 = 0.0, acc1 = 0.0;
            int j = 0;
            for (; j + 1 < n; j += 2) {
                acc0 = fma((double)row[j], v[j], acc0);
                acc1 = fma((double)row[j + 1], v[j + 1], acc1);
            }
            if (j < n) acc0 = fma((double)row[j], v[j], acc0);
            w = (acc0 + acc1) - beta * vp[tid];
        }
        const double alpha = block_sum_d(tid < n ? w * v[tid] : 0.0, red);
        if (tid < n) w -= alpha * v[tid];
        const double b2 = block_sum_d(w * w, red);
        beta = sqrt(b2);
        if (tid == 0) {
            al[it] = alpha;
            be[it] = beta;
        }
        k = it + 1;
        if (!(beta > 1e-200)) break;  // invariant subspace found (uniform across the CTA)
        if (tid < n) {
            vp[tid] = v[tid];
            v[tid] = w / beta;
        }
        __syncthreads();
    }
    __syncthreads();
    // extreme eigenvalues of the k x k Lanczos matrix: warp 0 the smallest, warp 1 the largest
    if (tid < 64) {
        double gl = 1e300, gu = -1e300;
        for (int i = 0; i < k; ++i) {
            const double r = ((i > 0) ? fabs(be[i - 1]) : 0.0) + ((i < k - 1) ? fabs(be[i]) : 0.0);
            gl = fmin(gl, al[i] - r);
            gu = fmax(gu, al[i] + r);
        }
        const double pad = 1e-12 * fmax(fabs(gl), fabs(gu)) + 1e-300;
        const bool low = tid < 32;
        const double ev = warp_multisect(al, be, k, gl - pad, gu + pad, low ? 1 : k);
        if ((tid & 31) == 0) a.scal[(long long)env * 4 + (low ? 0 : 1)] = ev;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// D1' lanczos2_kernel -- the same recurrence on 1024 threads with the matrix held ONCE, as float64, in shared memory
// (packed lower triangle, 161 KB at n = 200): no float -> double conversion inside the loop (F2F.F64.F32 runs at a quarter of the
// DFMA rate and made D1 a 150 us kernel), and every element is read once per product and used twice (y_i += a_ij v_j and
// y_j += a_ij v_i).  One warp owns ~7 rows (paired short + long); lane l owns the columns l, l + 32, ...: the row sums are
// reduced across the warp with an 8-value butterfly (9 shuffles instead of 40), the column sums stay lane-private and are added
// across warps through a [32][n] buffer in a fixed order (bit-reproducible).
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int TL2 = 1024;

// Eight values per lane -> their warp totals: value q = 4 b4 + 2 b3 + b2 (bits of the lane index) ends up in every lane of that group
__device__ __forceinline__ double multi_reduce8(double (&x)[8], int lane) {
    bool hi = (lane & 16) != 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double send = hi ? x[q] : x[q + 4], keep = hi ? x[q + 4] : x[q];
        x[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = (lane & 8) != 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const double send = hi ? x[q] : x[q + 2], keep = hi ? x[q + 2] : x[q];
        x[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = (lane & 4) != 0;
    {
        const double send = hi ? x[0] : x[1], keep = hi ? x[1] : x[0];
        x[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    x[0] += __shfl_xor_sync(0xffffffffu, x[0], 2);
    x[0] += __shfl_xor_sync(0xffffffffu, x[0], 1);
    return x[0];
}

struct Lanczos2Smem {
    int nrow_doubles, nslot;
    size_t bytes;
    bool fits;
};
// Row i (group g = i / 32) holds its STRICTLY lower elements zero-padded to 32 (g + 1) doubles, so that the product loop has
// compile-time trip counts and no predicates; the diagonal lives apart.
__host__ __device__ inline int lz_row_offset(int i) {
    const int g = i >> 5;
    return 512 * g * (g + 1) + 32 * (i & 31) * (g + 1);
}
// column-sum slots: as many of the 32 warps as fit next to the matrix (16 at n = 200: two hand-over passes)
__host__ __device__ inline Lanczos2Smem lanczos2_layout(int n) {
    Lanczos2Smem L;
    L.nrow_doubles = lz_row_offset(n);
    const size_t fixed = ((size_t)L.nrow_doubles + 3 * 224 + 4 * 32 + 2 * kLanczosMax) * sizeof(double) + 64;
    L.nslot = 32;
    while (L.nslot > 1 && fixed + (size_t)L.nslot * n * sizeof(double) > (size_t)227 * 1024) L.nslot >>= 1;
    L.bytes = fixed + (size_t)L.nslot * n * sizeof(double);
    L.fits = L.bytes <= (size_t)227 * 1024 && L.nslot >= 8;
    return L;
}

__device__ __forceinline__ int lz_row_of(int warp, int k) { return (k & 1) ? 32 * k + 31 - warp : 32 * k + warp; }

// number of eigenvalues of the k x k Lanczos matrix below x: sign changes of the leading principal minors p_i(x), the
// division-free form of the Sturm sequence (al, be pre-scaled so that nothing over- or underflows in k <= 32 steps)
__device__ __forceinline__ int sturm_count_poly(const double* al, const double* b2, int k, double x) {
    double pm = 1.0, p = al[0] - x;
    int sg_prev = 1, cnt = 0;
    {
        const int sg = (p > 0.0) ? 1 : ((p < 0.0) ? -1 : -sg_prev);
        cnt += (sg != sg_prev);
        sg_prev = sg;
    }
    for (int i = 1; i < k; ++i) {
        const double pn = fma(al[i] - x, p, -b2[i - 1] * pm);
        pm = p;
        p = pn;
        const int sg = (p > 0.0) ? 1 : ((p < 0.0) ? -1 : -sg_prev);
        cnt += (sg != sg_prev);
        sg_prev = sg;
    }
    return cnt;
}

#define DENSE_STAMP(slot)                                                        \
    do {                                                                         \
        if (a.prof && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) a.prof[slot] = clock64(); \
    } while (0)

// one row of the product: KK + 1 chunks of 32 columns (compile-time), no predicates: the padding is zero
template <int KK>
__device__ __forceinline__ double lz_row_product(const double* __restrict__ Ai, const double (&vr)[7], double vi, double (&cacc)[7], int lane) {
    double r = 0.0;
#pragma unroll
    for (int c = 0; c <= KK; ++c) {
        const double aij = Ai[lane + 32 * c];
        r = fma(aij, vr[c], r);
        cacc[c] = fma(aij, vi, cacc[c]);
    }
    return r;
}

__global__ void __launch_bounds__(TL2, 1) lanczos2_kernel(const DenseArgs a) {
    COVO_DYN_SMEM(smraw);
    const int n = a.n, tid = threadIdx.x, env = blockIdx.x, lane = tid & 31, warp = tid >> 5;
    const Lanczos2Smem L = lanczos2_layout(n);
    double* Ad = reinterpret_cast<double*>(smraw);  // padded strictly-lower rows, row i at lz_row_offset(i)
    double* dg = Ad + L.nrow_doubles;               // [224] diagonal
    double* v = dg + 224;                           // [224] current vector, zero beyond n
    double* rowy = v + 224;                         // [224] row-part of the product
    double* red = rowy + 224;                       // [2][32] alpha partials (by iteration parity), [2][32] beta partials
    double* al = red + 4 * 32;                      // [kLanczosMax]
    double* be = al + kLanczosMax;                  // [kLanczosMax]
    double* cpart = be + kLanczosMax;               // [nslot][n]
    const float* Rg = a.R + (long long)env * n * n;
    float* Asym = a.Asym ? a.Asym + (long long)env * n * n : nullptr;
    DENSE_STAMP(48);
    // (R + R^T)/2 in float32 as controllers/covo.py:117 forms it, then widened.  Both passes read global memory along rows
    // (7 independent loads in flight per thread): pass 1 the lower triangle with the diagonal (row i, columns j <= i), pass 2 the
    // upper one (row j, columns i > j) into the same slots.
    for (int i = warp; i < n; i += 32) {
        double* Ai = Ad + lz_row_offset(i);
        const int len = 32 * ((i >> 5) + 1);
        float x[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const int j = lane + 32 * c;
            x[c] = (j <= i) ? Rg[(long long)i * n + j] : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const int j = lane + 32 * c;
            if (j < len) Ai[j] = (j < i) ? (double)x[c] : 0.0;
            if (j == i) {
                dg[i] = (double)x[c];
                if (Asym) Asym[(long long)i * n + i] = x[c];
            }
        }
    }
    if (tid >= n && tid < 224) dg[tid] = 0.0;
    __syncthreads();
    for (int j = warp; j < n; j += 32) {
        float x[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const int i = j + 1 + lane + 32 * c;
            x[c] = (i < n) ? Rg[(long long)j * n + i] : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const int i = j + 1 + lane + 32 * c;
            if (i < n) {
                double* p = Ad + lz_row_offset(i) + j;
                const float sym = 0.5f * ((float)*p + x[c]);
                *p = (double)sym;
                if (Asym) {  // the symmetrised matrix for the factorisation kernels (they then read rows only)
                    Asym[(long long)j * n + i] = sym;
                    Asym[(long long)i * n + j] = sym;
                }
            }
        }
    }
    double vj = 0.0, vprev = 0.0, dj = 0.0;  // thread j < n keeps its components in registers
    if (tid < 224) {
        double x0 = 0.0;
        if (tid < n) x0 = cos(0.37 * (double)tid + 0.1) + 0.01 * (double)tid / (double)n;
        vj = x0;
    }
    {  // normalise the start vector
        const double p2 = warp_sum_d(vj * vj);
        if (lane == 0) red[warp] = p2;
        __syncthreads();  // (also orders the matrix stores before the first product)
        double s2 = 0.0;
#pragma unroll
        for (int q = 0; q < 7; ++q) s2 += red[q];
        vj /= sqrt(s2);
    }
    if (tid < 224) {
        v[tid] = vj;
        dj = dg[tid];
    }
    // the rows of this warp (paired short + long) and where they start
    int roff[7];
#pragma unroll
    for (int kk = 0; kk < 7; ++kk) {
        const int i = lz_row_of(warp, kk);
        roff[kk] = (i < n) ? lz_row_offset(i) : -1;
    }
    const int passes = 32 / L.nslot, my_pass = warp / L.nslot;
    double* cp = cpart + (warp % L.nslot) * n;
    __syncthreads();
    DENSE_STAMP(49);
    const int k_max = min(kLanczosMax, n);
    int k = 0;
    double beta = 0.0;
    for (int it = 0; it < k_max; ++it) {
        // ---- y = A v: strictly lower part, used twice (row sums and column sums) ----------------------------------------------
        double vr[7], cacc[7], rs[8];
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            vr[c] = v[lane + 32 * c];
            cacc[c] = 0.0;
        }
        rs[0] = (roff[0] >= 0) ? lz_row_product<0>(Ad + roff[0], vr, v[lz_row_of(warp, 0)], cacc, lane) : 0.0;
        rs[1] = (roff[1] >= 0) ? lz_row_product<1>(Ad + roff[1], vr, v[lz_row_of(warp, 1)], cacc, lane) : 0.0;
        rs[2] = (roff[2] >= 0) ? lz_row_product<2>(Ad + roff[2], vr, v[lz_row_of(warp, 2)], cacc, lane) : 0.0;
        rs[3] = (roff[3] >= 0) ? lz_row_product<3>(Ad + roff[3], vr, v[lz_row_of(warp, 3)], cacc, lane) : 0.0;
        rs[4] = (roff[4] >= 0) ? lz_row_product<4>(Ad + roff[4], vr, v[lz_row_of(warp, 4)], cacc, lane) : 0.0;
        rs[5] = (roff[5] >= 0) ? lz_row_product<5>(Ad + roff[5], vr, v[lz_row_of(warp, 5)], cacc, lane) : 0.0;
        rs[6] = (roff[6] >= 0) ? lz_row_product<6>(Ad + roff[6], vr, v[lz_row_of(warp, 6)], cacc, lane) : 0.0;
        rs[7] = 0.0;
        const double tot = multi_reduce8(rs, lane);
        {
            const int q = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            const int i = lz_row_of(warp, q);
            if ((lane & 3) == 0 && q < 7 && i < n) rowy[i] = tot;
        }
        for (int ps = 0; ps < passes; ++ps) {
            if (my_pass == ps) {
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    const int j = lane + 32 * c;
                    if (j < n) cp[j] = (ps == 0) ? cacc[c] : cp[j] + cacc[c];
                }
            }
            __syncthreads();
        }
        // ---- three-term recurrence, thread j --------------------------------------------------------------------------
        double w = 0.0;
        if (tid < n) {
            double s0 = rowy[tid], s1 = dj * vj, s2 = 0.0, s3 = 0.0;
            for (int q = 0; q < L.nslot; q += 4) {
                s0 += cpart[q * n + tid];
                s1 += cpart[(q + 1) * n + tid];
                s2 += cpart[(q + 2) * n + tid];
                s3 += cpart[(q + 3) * n + tid];
            }
            w = ((s0 + s1) + (s2 + s3)) - beta * vprev;
        }
        double* ra = red + (it & 1) * 32;
        double* rb = red + 64 + (it & 1) * 32;
        if (warp < 7) {
            const double pa = warp_sum_d(w * vj);
            if (lane == 0) ra[warp] = pa;
        }
        __syncthreads();
        double alpha = 0.0;
#pragma unroll
        for (int q = 0; q < 7; ++q) alpha += ra[q];
        w -= alpha * vj;
        if (warp < 7) {
            const double pb = warp_sum_d(w * w);
            if (lane == 0) rb[warp] = pb;
        }
        __syncthreads();
        double b2 = 0.0;
#pragma unroll
        for (int q = 0; q < 7; ++q) b2 += rb[q];
        beta = sqrt(b2);
        if (tid == 0) {
            al[it] = alpha;
            be[it] = beta;
        }
        k = it + 1;
        if (!(beta > 1e-200)) break;  // invariant subspace found (uniform across the CTA)
        if (tid < n) {
            vprev = vj;
            vj = w / beta;
            v[tid] = vj;
        }
        __syncthreads();
    }
    __syncthreads();
    DENSE_STAMP(50);
    // Extreme eigenvalues of the k x k Lanczos matrix by multisection of the (division-free) Sturm count.  Threads 0..127 look for
    // the smallest with FOUR trial shifts each (513-way, 5 rounds: 513^5 = 3.5e13 of the Gershgorin interval; the four independent
    // recurrences hide the DFMA latency); threads 128..255 bracket the largest to 129^-2 (it only selects the approximation interval).
    if (tid < 256) {
        double gl = 1e300, gu = -1e300;
        for (int i = 0; i < k; ++i) {
            const double r = ((i > 0) ? fabs(be[i - 1]) : 0.0) + ((i < k - 1) ? fabs(be[i]) : 0.0);
            gl = fmin(gl, al[i] - r);
            gu = fmax(gu, al[i] + r);
        }
        const double pad = 1e-12 * fmax(fabs(gl), fabs(gu)) + 1e-300;
        gl -= pad;
        gu += pad;
        // scaled copy: (T - gl) / (gu - gl) has its spectrum in [0, 1]
        double* sal = cpart;             // [32]  (the column buffers are dead)
        double* sb2 = cpart + 32;        // [32]
        int* first = reinterpret_cast<int*>(cpart + 64);  // [2][8]
        const double isc = 1.0 / (gu - gl);
        if (tid < k) {
            sal[tid] = (al[tid] - gl) * isc;
            const double b = be[tid] * isc;
            sb2[tid] = b * b;
        }
        COVO_NAMED_BARRIER(2, 256);
        const bool low = tid < 128;
        const int t = tid & 127, target = low ? 1 : k;
        double lo = 0.0, hi = 1.0;
        for (int round = 0; round < 5; ++round) {
            int f_mine = 1 << 20;  // first trial index (of this thread's) whose count reaches the target
            double step;
            if (low) {
                step = (hi - lo) / 513.0;
#pragma unroll
                for (int u = 3; u >= 0; --u) {
                    const int idx = 4 * t + u;  // trial shifts in increasing order across (t, u)
                    if (sturm_count_poly(sal, sb2, k, lo + step * (double)(idx + 1)) >= target) f_mine = idx;
                }
            } else {
                step = (hi - lo) / 129.0;
                if (round < 2 && sturm_count_poly(sal, sb2, k, lo + step * (double)(t + 1)) >= target) f_mine = t;
            }
            // first index over the 4 warps of this search (monotone predicate: the minimum over threads)
            int fm = f_mine;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) fm = min(fm, __shfl_xor_sync(0xffffffffu, fm, o));
            int* fr = first + (round & 1) * 8;
            if (lane == 0) fr[warp] = fm;
            COVO_NAMED_BARRIER(2, 256);
            const int w0 = low ? 0 : 4;
            const int f = min(min(fr[w0], fr[w0 + 1]), min(fr[w0 + 2], fr[w0 + 3]));
            const int nsub = low ? 512 : 128;
            if (low || round < 2) {
                if (f >= nsub) {
                    lo = lo + step * (double)nsub;  // the crossing is in the last sub-interval
                } else {
                    hi = lo + step * (double)(f + 1);
                    lo = lo + step * (double)f;
                }
            }
        }
        if (t == 0) a.scal[(long long)env * 4 + (low ? 0 : 1)] = gl + (low ? 0.5 * (lo + hi) : hi) * (gu - gl);
    }
    DENSE_STAMP(51);
}

// ---------------------------------------------------------------------------------------------------------------------------
// D2
// ---------------------------------------------------------------------------------------------------------------------------
// Blocked right-looking Cholesky (NB = 8, look-ahead) of the n x n matrix in As (row-major, stride n), in place: the scheme of
// cholesky_kernel (sigma.cu), restated here as a device function.  Lp: 2 x [8][n_pad] panel buffers + [8][8].  Returns with the
// lower triangle of As holding L; `bad` is set when a pivot is not positive.
__device__ __forceinline__ void chol_factor_smem(float* As, float* Lp, int n, int n_pad, int* bad_out) {
    const int tid = threadIdx.x;
    constexpr int kPanelThreads = 256;
    float* LpA = Lp;
    float* LpB = Lp + 8 * n_pad;
    float* Lp8 = LpB + 8 * n_pad;
    const int pt = tid - (TD - kPanelThreads);
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
                const float rinv = rsqrt_newton_d(dcc);
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
                if (bad) *bad_out = 1;
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) Lp8[r * 8 + c] = (c <= r) ? d[r][c] : 0.f;
            }
        }
    };
    if (pt >= 0) factor_panel(0, min(8, n), LpA);
    __syncthreads();
    for (int jb = 0, it = 0; jb < n; jb += 8, ++it) {
        const int nb = min(8, n - jb);
        float* LpCur = (it & 1) ? LpB : LpA;
        float* LpNext = (it & 1) ? LpA : LpB;
        if (pt >= 0 && pt < 64) {
            const int r = pt >> 3, c = pt & 7;
            if (r < nb && c < nb) As[(jb + r) * n + jb + c] = Lp8[pt];
        }
        const int r0 = jb + nb;
        if (r0 >= n) break;
        const int nbn = min(8, n - r0);
        if (pt >= 0) {
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
            COVO_NAMED_BARRIER(1, 256);
            factor_panel(r0, nbn, LpNext);
        } else {
            const int c0 = r0 + nbn;
            const int T = (n - c0) >> 2;
            const int ntiles = T * (T + 1) / 2;
            for (int q = tid; q < ntiles; q += TD - kPanelThreads) {
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
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TD, 1) shifted_inverse_kernel(const DenseArgs a) {
    COVO_DYN_SMEM(smraw);
    const int n = a.n, n_pad = a.n_pad, tid = threadIdx.x, pole = blockIdx.x, env = blockIdx.y, lane = tid & 31, warp = tid >> 5;
    float* As = reinterpret_cast<float*>(smraw);  // [n][n]
    float* Lp = As + n * n;                       // Cholesky panel buffers
    float* dinv = Lp + 2 * 8 * n_pad + 64;        // [n] 1 / L_ii  (= X_ii)
    int* bad = reinterpret_cast<int*>(dinv + n_pad);
    const double lam_min = a.scal[(long long)env * 4 + 0], lam_max = a.scal[(long long)env * 4 + 1];
    // ladder entry covering [1e-2, M]; the Ritz value can only underestimate lambda_max: 2 % of the width as margin
    int lad = 0;
    {
        const double Mb = 1.02 * (lam_max - lam_min) + kOffset;
        double Mi = kOffset * (1.0 - 1e-7) * 256.0;
        while (lad < kZoloLadder - 1 && Mi < Mb) {
            Mi *= 4.0;
            ++lad;
        }
        if (Mi < Mb && tid == 0) a.status[env] = 1;
    }
    const double* zt = a.zolo + (size_t)lad * 2 * kZoloPoles;
    const bool want_logdet = pole == kZoloPoles;
    const double shift = (kOffset - lam_min) + (want_logdet ? 0.0 : zt[pole]);
    const float wj = want_logdet ? 0.f : (float)zt[kZoloPoles + pole];
    if (tid == 0) *bad = 0;
    // A + t_j I = (R + R^T)/2 + shift I  (the shift is added in double and rounded once)
    const float* Rg = a.R + (long long)env * n * n;
    for (int idx = tid; idx < n * n; idx += TD) {
        const int i = idx / n, j = idx - i * n;
        float val = 0.5f * (Rg[idx] + Rg[(long long)j * n + i]);
        if (i == j) val = (float)((double)val + shift);
        As[idx] = val;
    }
    __syncthreads();
    chol_factor_smem(As, Lp, n, n_pad, bad);
    if (*bad && tid == 0) a.status[env] = 2;
    if (want_logdet) {
        if (warp == 0) {
            double s = 0.0;
            for (int i = lane; i < n; i += 32) s += log((double)As[i * n + i]);
            s = warp_sum_d(s);
            if (lane == 0) a.scal[(long long)env * 4 + 2] = 2.0 * s;
        }
        return;
    }
    // X = L^-1 (lower triangular).  X_ii = 1 / L_ii lives in dinv; X_ic (i > c) is stored TRANSPOSED in the strict upper triangle,
    // As[c][i], so the factor (strict lower triangle + diagonal) is never overwritten.  One warp per column c, lanes over k:
    //     X_ic = -dinv[i] * sum_{k = c}^{i-1} L_ik X_kc
    for (int i = tid; i < n; i += TD) dinv[i] = 1.0f / As[i * n + i];
    __syncthreads();
    for (int c = warp; c < n; c += TD / 32) {
        float* Xc = As + c * n;  // row c of the upper triangle: X_kc at Xc[k], k > c
        const float xcc = dinv[c];
        for (int i = c + 1; i < n; ++i) {
            const float* Li = As + i * n;
            float s = 0.f;
            for (int k = c + 1 + lane; k < i; k += 32) s = fmaf(Li[k], Xc[k], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) Xc[i] = -(s + Li[c] * xcc) * dinv[i];
            __syncwarp();
        }
    }
    __syncthreads();
    // w_j (A + t_j I)^-1 = w_j X^T X:  P_ab = sum_{r >= a} X_ra X_rb  (a >= b).  One warp per (a, b), lanes over r.
    float* Xg = a.Xbuf + ((long long)env * kZoloPoles + pole) * n * n;
    const int npairs = n * (n + 1) / 2;
    for (int q = warp; q < npairs; q += TD / 32) {
        int ia = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
        while (ia * (ia + 1) / 2 > q) --ia;
        while ((ia + 1) * (ia + 2) / 2 <= q) ++ia;
        const int ib = q - ia * (ia + 1) / 2;  // ib <= ia
        const float* Xa = As + ia * n;
        const float* Xb = As + ib * n;
        float s = 0.f;
        for (int r = ia + 1 + lane; r < n; r += 32) s = fmaf(Xa[r], Xb[r], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            const float xab = (ia == ib) ? dinv[ia] : Xb[ia];  // X_ab, the r = a term (X_aa = dinv[a])
            Xg[ia * n + ib] = wj * (s + dinv[ia] * xab);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// D2' (alternative to D2, COVO_SIGMA=dense-gj): w_j (A + t_j I)^-1 by in-place Gauss-Jordan elimination without pivoting (A + t_j I
// is SPD: every pivot is a positive Schur complement), the matrix RESIDENT IN REGISTERS.  512 threads as a 16 x 32 grid; thread
// (ty, tx) owns the elements (ty + 16 a, tx + 32 b), a < RI, b < 7, rows packed in pairs for FFMA2.  Step k:
//     p = a_kk;  a_ij -= a_ik a_kj / p  (i, j != k);  row k <- r / p;  column k <- -c / p;  a_kk <- 1 / p.
// With this sign convention the matrix stays symmetric on the not yet eliminated index set and ANTI-symmetric between eliminated
// and remaining indices (a_im = -a_mi for m < k <= i), so column k is row k with the sign of (i < k): only the ROW is published
// (by the one warp that owns it; double-buffered, one barrier per step), never the column.  The loop over k is unrolled over the
// 16-row blocks, so every register-tile index is a compile-time constant and the row / column fix-ups touch 7 and 14 registers
// instead of the whole tile.  log det A = sum log p_k comes for free (the CTA without a pole).  n^3 FMA per matrix instead of the
// three n^3 / 3 sweeps of D2, but no serial panel chain.  Numerics (CPU study, fp32): Sigma to 1e-6 on the tracking / zigzag path.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int TG = 512;

template <int RI>
__global__ void __launch_bounds__(TG, 1) gj_inverse_kernel(const DenseArgs a) {
    COVO_DYN_SMEM(smraw);
    constexpr int CJ = 7, RP = (RI + 1) / 2;
    const int n = a.n, tid = threadIdx.x, pole = blockIdx.x, env = blockIdx.y, tx = tid & 31, ty = tid >> 5;
    float* rbuf = reinterpret_cast<float*>(smraw);  // [2][512]: row k and the column multipliers, double-buffered
    const double lam_min = a.scal[(long long)env * 4 + 0], lam_max = a.scal[(long long)env * 4 + 1];
    int lad = 0;
    {
        const double Mb = 1.02 * (lam_max - lam_min) + kOffset;
        double Mi = kOffset * (1.0 - 1e-7) * 256.0;
        while (lad < kZoloLadder - 1 && Mi < Mb) {
            Mi *= 4.0;
            ++lad;
        }
        if (Mi < Mb && tid == 0) a.status[env] = 1;
    }
    const double* zt = a.zolo + (size_t)lad * 2 * kZoloPoles;
    const bool want_logdet = pole == kZoloPoles;
    const double shift = (kOffset - lam_min) + (want_logdet ? 0.0 : zt[pole]);
    const float wj = want_logdet ? 0.f : (float)zt[kZoloPoles + pole];
    const float* Rg = a.R + (long long)env * n * n;
    // tile load: (R + R^T)/2 + shift I; elements outside the matrix form an identity block (never a pivot, no coupling)
    float2 acc[RP][CJ];  // [row pair][col]: .x = tile row 2q, .y = tile row 2q + 1
#pragma unroll
    for (int q = 0; q < RP; ++q)
#pragma unroll
        for (int b = 0; b < CJ; ++b) {
            float v2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = ty + 16 * (2 * q + h), j = tx + 32 * b;
                float v = (i == j) ? 1.f : 0.f;
                if (i < n && j < n) {
                    v = 0.5f * (Rg[(long long)i * n + j] + Rg[(long long)j * n + i]);
                    if (i == j) v = (float)((double)v + shift);
                }
                v2[h] = v;
            }
            acc[q][b] = make_float2(v2[0], v2[1]);
        }
    float my_pivot = 1.f;  // thread k keeps pivot k: the logarithms are taken once, after the elimination
    bool bad = false;
#pragma unroll
    for (int ak = 0; ak < 2 * RP; ++ak) {  // 16-row block of the pivot: tile row ak, tile column ak / 2 -- compile-time
        const int qk = ak >> 1, bk = ak >> 1;  // pivot row lives in acc[qk][.].(ak & 1 ? y : x); pivot column is tile column bk
        if (16 * ak < n) {
            for (int kk = 0; kk < 16; ++kk) {
                const int k = 16 * ak + kk;
                if (k >= n) break;
                float* rb = rbuf + (k & 1) * 512;  // [256] row k, then [256] the column multipliers -a_ik / p
                float* cb = rb + 256;
                float piv = 0.f;
                if (ty == kk) {
                    // The warp that owns row k publishes it, and with it the multipliers of the column: by (anti)symmetry
                    // a_ik = (i < k ? -1 : 1) a_ki, so -a_ik / p is the row again, signed and scaled -- the other 15 warps
                    // never see the column, only these two vectors.
                    const float own = (ak & 1) ? acc[qk][bk].y : acc[qk][bk].x;  // lane k mod 32 holds the pivot
                    const float p = __shfl_sync(0xffffffffu, own, k & 31);
                    if (!(p > 0.f)) bad = true;
                    piv = 1.0f / p;
#pragma unroll
                    for (int b = 0; b < CJ; ++b) {
                        const int j = tx + 32 * b;
                        const float r = (ak & 1) ? acc[qk][b].y : acc[qk][b].x;
                        rb[j] = r;
                        cb[j] = (j < k ? r : -r) * piv;
                    }
                }
                __syncthreads();
                if (tid == k) my_pivot = rb[k];
                float rk[CJ];
#pragma unroll
                for (int b = 0; b < CJ; ++b) rk[b] = rb[tx + 32 * b];
                float2 cn[RP];
#pragma unroll
                for (int q = 0; q < RP; ++q) cn[q] = make_float2(cb[ty + 32 * q], cb[ty + 32 * q + 16]);
#pragma unroll
                for (int q = 0; q < RP; ++q)
#pragma unroll
                    for (int b = 0; b < CJ; ++b) acc[q][b] = __ffma2_rn(cn[q], make_float2(rk[b], rk[b]), acc[q][b]);
                if (tx == (k & 31)) {  // column k (tile column bk): -a_ik / p
#pragma unroll
                    for (int q = 0; q < RP; ++q) acc[q][bk] = cn[q];
                }
                if (ty == kk) {  // row k: a_kj / p, pivot 1 / p
#pragma unroll
                    for (int b = 0; b < CJ; ++b) {
                        const float v = (tx + 32 * b == k) ? piv : rk[b] * piv;
                        if (ak & 1) acc[qk][b].y = v;
                        else acc[qk][b].x = v;
                    }
                }
                // no second barrier: step k + 1 publishes into the other buffer, and step k + 2 writes this one only after the
                // barrier of step k + 1, which every thread reaches after it has finished reading here
            }
        }
    }
    if (bad && tid == 0) a.status[env] = 2;
    if (want_logdet) {  // log det A = sum_k log p_k  (n <= TG: one pivot per thread)
        double lp = (tid < n) ? log((double)my_pivot) : 0.0;
        lp = warp_sum_d(lp);
        double* red = reinterpret_cast<double*>(rbuf);
        __syncthreads();  // the row buffers are no longer read
        if (tx == 0) red[ty] = lp;
        __syncthreads();
        if (tid == 0) {
            double sum = 0.0;
            for (int w = 0; w < TG / 32; ++w) sum += red[w];
            a.scal[(long long)env * 4 + 2] = sum;
        }
        return;
    }
    float* Xg = a.Xbuf + ((long long)env * kZoloPoles + pole) * n * n;
#pragma unroll
    for (int q = 0; q < RP; ++q)
#pragma unroll
        for (int b = 0; b < CJ; ++b) {
            const int j = tx + 32 * b;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = ty + 16 * (2 * q + h);
                if (i < n && j <= i) Xg[i * n + j] = wj * (h ? acc[q][b].y : acc[q][b].x);  // lower triangle, as D2
            }
        }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Thread-block-cluster primitives shared by D1'' and D2'': distributed shared memory stores that carry their own completion
// signal (st.async ... mbarrier::complete_tx), mbarrier waits, the cluster barrier.  tests/emu provides CPU stand-ins.
// ---------------------------------------------------------------------------------------------------------------------------
#if defined(COVO_CPU_EMU)
__device__ __forceinline__ unsigned gjb_rank() { return emu_cluster_rank(); }
__device__ __forceinline__ void gjb_cluster_sync() { emu_cluster_barrier(); }
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
__device__ __forceinline__ float gjb_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * fmaf(-x, r, 2.0f);
}
#endif

// ---------------------------------------------------------------------------------------------------------------------------
// D1'' lanczos_cluster_kernel -- the Lanczos recurrence on an 8-CTA cluster.  Measured on B200: scalar float64 instructions issue at
// ~16 lanes / clock / SM, so a 200 x 200 product is >= 2500 cycles on one SM whatever the layout (D1 and D1' both sit at ~3.5 us per
// iteration); the only way down is more SMs.  CTA c keeps rows [c R, (c + 1) R) of (R + R^T)/2 in REGISTERS as float64, four
// threads per row (columns 4 k + p), and per iteration
//     y = A w / beta  (w: last iteration's unnormalised vector, complete in every CTA's shared memory),
//     w' = y - beta v_prev,  warp partials of w'.v and w'.w',
// then ONE exchange: every row leader sends its w' and every warp its two partials to all 8 CTAs with st.async (the stores count
// themselves into the receiver's mbarrier: no fence, no cluster barrier), everybody waits for its own mbarrier and finishes
// alpha, beta^2 = w'.w' - alpha^2 and w = w' - alpha v redundantly.  The all-to-all makes every iteration an implicit barrier,
// so two buffers (iteration parity) are enough.  CTA 0 then finds the smallest Ritz value: five rounds of 128-way multisection
// (division-free Sturm counts: the bracket is 2.8e-11 of the Gershgorin interval wide) and a short Newton polish from its left end
// (monotone for a real-rooted polynomial).  Required accuracy: |error| << 1e-2 * 2e-5 = 2e-7 ABSOLUTE (the offset 1e-2 of
// controllers/covo.py:121 sets the scale), not relative to |R|.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int LC_CL = 8;      // CTAs
constexpr int LC_T = 128;     // threads per CTA: 32 rows x 4 column phases
constexpr int LC_KMAX = 56;   // columns per thread (n <= 224)

struct LcSmem {
    double v[2][224];                 // normalised Lanczos vector v_k (by iteration parity), all n entries, zero beyond n
    double u[2][224];                 // u = A v_k - beta_{k-1} v_{k-1}, gathered from all CTAs (by iteration parity)
    double part[2][LC_CL * 4][2];     // (u.v_k, u.u) partial of every warp of the cluster
    double al[kLanczosMax], be[kLanczosMax];
    double sal[kLanczosMax], sb2[kLanczosMax];
    unsigned long long bar[2];
    int first[2][4];
    float salf[kLanczosMax], sb2f[kLanczosMax];
};

// float32 twin of sturm_count_poly for the first, coarse multisection rounds (the bracket is widened afterwards by far more than
// its rounding error can move a crossing)
__device__ __forceinline__ int sturm_count_poly_f32(const float* al, const float* b2, int k, float x) {
    float pm = 1.0f, p = al[0] - x;
    int sg_prev = 1, cnt = 0;
    {
        const int sg = (p > 0.f) ? 1 : ((p < 0.f) ? -1 : -sg_prev);
        cnt += (sg != sg_prev);
        sg_prev = sg;
    }
    for (int i = 1; i < k; ++i) {
        const float pn = fmaf(al[i] - x, p, -b2[i - 1] * pm);
        pm = p;
        p = pn;
        const int sg = (p > 0.f) ? 1 : ((p < 0.f) ? -1 : -sg_prev);
        cnt += (sg != sg_prev);
        sg_prev = sg;
    }
    return cnt;
}

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
    double* dst = (what == 0) ? r.u + parity * 224 : r.part + parity * (LC_CL * 4 * 2) + (what - 1);
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
    const unsigned dst = (what == 0) ? r.u + parity * 224 * 8 : r.part + parity * (LC_CL * 4 * 2 * 8) + (what - 1) * 8;
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst), "l"(__double_as_longlong(v)),
                 "r"(r.bar + parity * 8)
                 : "memory");
}
#endif

// characteristic polynomial of the scaled k x k Lanczos matrix and its derivative at x (three-term recurrences)
__device__ __forceinline__ void lc_poly_newton(const double* al, const double* b2, int k, double x, double& p_out, double& dp_out) {
    double pm = 1.0, p = al[0] - x, dm = 0.0, d = -1.0;
    for (int i = 1; i < k; ++i) {
        const double t = al[i] - x;
        const double pn = fma(t, p, -b2[i - 1] * pm);
        const double dn = fma(t, d, -b2[i - 1] * dm) - p;
        pm = p;
        p = pn;
        dm = d;
        d = dn;
    }
    p_out = p;
    dp_out = d;
}

__global__ void __launch_bounds__(LC_T, 1) lanczos_cluster_kernel(const DenseArgs a) {
    COVO_DYN_SMEM(smraw);
    LcSmem& sm = *reinterpret_cast<LcSmem*>(smraw);
    const int n = a.n, tid = threadIdx.x, env = blockIdx.y, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)gjb_rank();
    const int R = (n + LC_CL - 1) / LC_CL;      // rows per CTA (25 at n = 200)
    const int rl = tid >> 2, ph = tid & 3;      // local row, column phase
    const int row = rank * R + rl;
    const bool has_row = rl < R && row < n;
    const bool leader = has_row && ph == 0;
    const int kc = (n + 3) >> 2;                // columns per thread
    const float* Rg = a.R + (long long)env * n * n;
    float* Asym = a.Asym ? a.Asym + (long long)env * n * n : nullptr;
    DENSE_STAMP(48);
    if (tid == 0) {
        gjb_mbar_init(&sm.bar[0], 1);
        gjb_mbar_init(&sm.bar[1], 1);
#if !defined(COVO_CPU_EMU)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    // rows of (R + R^T)/2 (float32, controllers/covo.py:117) widened into registers
    double ar[LC_KMAX];
    {
        float xa[LC_KMAX], xb[LC_KMAX];  // all loads in flight before the first store (the compiler must assume Asym aliases R)
#pragma unroll
        for (int k = 0; k < LC_KMAX; ++k) {
            const int j = 4 * k + ph;
            const bool ok = has_row && j < n;
            xa[k] = ok ? __ldg(Rg + (long long)row * n + j) : 0.f;
            xb[k] = ok ? __ldg(Rg + (long long)j * n + row) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < LC_KMAX; ++k) {
            const int j = 4 * k + ph;
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
    __syncthreads();
    double vj = has_row ? sm.v[0][row] : 0.0, vprev = 0.0;  // row leaders keep v_k[row], v_{k-1}[row]
    const int n_warps_total = LC_CL * 4;
    const int tx_bytes = n * 8 + n_warps_total * 16;
    gjb_cluster_sync();  // barriers initialised everywhere before the first send
    DENSE_STAMP(49);
    // 24 steps: over 72 scenarios (three tasks, six seeds, four flight phases, H = 50) the smallest Ritz value is within 2e-8 of
    // lambda_min after 20 steps and 5e-12 after 24 (needed: << 2e-7 absolute); every step is a dependent ~1 us exchange
    const int k_max = min(kLanczosSteps, n);
    int kdone = 0;
    double beta_prev = 0.0;
    LcRemote rem[LC_CL];  // where this thread's row entry / this warp's partial slot / the barrier live in every CTA of the cluster
#pragma unroll
    for (int r = 0; r < LC_CL; ++r)
        rem[r] = lc_remote(&sm.u[0][has_row ? row : 0], &sm.part[0][rank * 4 + warp][0], &sm.bar[0], (unsigned)r);
    for (int it = 0; it < k_max; ++it) {
        const int pc = it & 1;  // v_k is in v[pc]; this iteration's exchange uses u[pc], part[pc], bar[pc]
        // y_row = (A v_k)_row: four threads per row, four independent chains each
        // (dependent float64 operations are ~40 cycles apart on this pipe: eight chains of seven, not four of thirteen)
        double ac[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const double* vv = sm.v[pc];
#pragma unroll
        for (int k = 0; k < LC_KMAX; k += 8) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (k + e < kc) ac[e] = fma(ar[k + e], vv[4 * (k + e) + ph], ac[e]);
        }
        double y = ((ac[0] + ac[1]) + (ac[2] + ac[3])) + ((ac[4] + ac[5]) + (ac[6] + ac[7]));
        y += __shfl_xor_sync(0xffffffffu, y, 1);
        y += __shfl_xor_sync(0xffffffffu, y, 2);
        // u = y - beta_{k-1} v_{k-1} (row leaders); partials of u.v_k and u.u over the 8 rows of the warp
        const double u = leader ? y - beta_prev * vprev : 0.0;
        double pa = u * vj, pu = u * u;
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            pa += __shfl_xor_sync(0xffffffffu, pa, o);
            pu += __shfl_xor_sync(0xffffffffu, pu, o);
        }
        if (tid == 0) gjb_mbar_expect(&sm.bar[pc], tx_bytes);
        long long tq0 = 0;
        if (a.prof && tid == 0 && rank == 0 && blockIdx.y == 0) tq0 = clock64();
        // the exchange: u of this row and the warp's partials to every CTA of the cluster (this one included)
#pragma unroll
        for (int r = 0; r < LC_CL; ++r) {
            if (leader) lc_send(rem[r], 0, pc, u);
            if (lane == 0) {
                lc_send(rem[r], 1, pc, pa);
                lc_send(rem[r], 2, pc, pu);
            }
        }
        gjb_mbar_wait(&sm.bar[pc], (unsigned)((it >> 1) & 1));
        if (a.prof && tid == 0 && rank == 0 && blockIdx.y == 0) a.prof[52] = (it == 0 ? 0 : a.prof[52]) + (clock64() - tq0);  // send + wait
        // alpha = u.v_k, beta_k^2 = |u - alpha v_k|^2 = u.u - alpha^2: every warp sums the 32 partial pairs with the same butterfly
        double qa = (lane < n_warps_total) ? sm.part[pc][lane][0] : 0.0;
        double qu = (lane < n_warps_total) ? sm.part[pc][lane][1] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {  // the two sums interleaved: five levels of (shuffle + add) each
            qa += __shfl_xor_sync(0xffffffffu, qa, o);
            qu += __shfl_xor_sync(0xffffffffu, qu, o);
        }
        const double alpha = qa;
        double b2 = qu - alpha * alpha;
        if (b2 < 1e-3 * qu) {
            // |u|^2 - alpha^2 cancels when beta << |alpha| (Krylov space nearly exhausted, small n): form |u - alpha v_k|^2 directly.
            // Every CTA holds all of u and v_k, so this needs no exchange; the branch is uniform across the cluster.
            double ps = 0.0;
            for (int j = tid; j < n; j += LC_T) {
                const double wj = sm.u[pc][j] - alpha * vv[j];
                ps = fma(wj, wj, ps);
            }
            ps = warp_sum_d(ps);
            double* red = &sm.sal[0];  // free until the Ritz stage
            if (lane == 0) red[warp] = ps;
            __syncthreads();
            b2 = (red[0] + red[1]) + (red[2] + red[3]);
            __syncthreads();
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
        if (tid == 0) {
            sm.al[it] = alpha;
            sm.be[it] = beta_new;
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
        __syncthreads();
    }
    __syncthreads();
    DENSE_STAMP(50);
    // ---- smallest Ritz value (CTA 0) ---------------------------------------------------------------------------------------
    if (rank == 0) {
        const int k = kdone;
        double gl = 1e300, gu = -1e300;
        for (int i = 0; i < k; ++i) {
            const double r = ((i > 0) ? fabs(sm.be[i - 1]) : 0.0) + ((i < k - 1) ? fabs(sm.be[i]) : 0.0);
            gl = fmin(gl, sm.al[i] - r);
            gu = fmax(gu, sm.al[i] + r);
        }
        const double pad = 1e-12 * fmax(fabs(gl), fabs(gu)) + 1e-300;
        gl -= pad;
        gu += pad;
        const double isc = 1.0 / (gu - gl);
        if (tid < k) {
            sm.sal[tid] = (sm.al[tid] - gl) * isc;
            const double b = sm.be[tid] * isc;
            sm.sb2[tid] = b * b;
            sm.salf[tid] = (float)sm.sal[tid];
            sm.sb2f[tid] = (float)sm.sb2[tid];
        }
        __syncthreads();
        double lo = 0.0, hi = 1.0;
        // rounds 0, 1 in float32 (the float64 pipe issues 16 lanes / clock: a float64 round costs ~0.65 us), then the bracket
        // [lo, hi] (6e-5 wide) is widened by 2e-5 on both sides -- 10x what float32 rounding of the scaled recurrence can move a
        // crossing -- and three float64 rounds bring it to 1e-4 / 129^3 = 4.7e-11 of the Gershgorin interval
        for (int round = 0; round < 5; ++round) {
            if (round == 2) {
                lo = fmax(lo - 2e-5, 0.0);
                hi = fmin(hi + 2e-5, 1.0);
            }
            const double step = (hi - lo) / 129.0;
            const double xs = lo + step * (double)(tid + 1);
            const bool ge = (round < 2 ? sturm_count_poly_f32(sm.salf, sm.sb2f, k, (float)xs) : sturm_count_poly(sm.sal, sm.sb2, k, xs)) >= 1;
            const unsigned m = __ballot_sync(0xffffffffu, ge);
            if (lane == 0) sm.first[round & 1][warp] = m ? 32 * warp + __ffs(m) - 1 : 128;
            __syncthreads();
            const int* fr = sm.first[round & 1];
            const int f = min(min(fr[0], fr[1]), min(fr[2], fr[3]));
            if (f >= 128) {
                lo = lo + step * 128.0;
            } else {
                hi = lo + step * (double)(f + 1);
                lo = lo + step * (double)f;
            }
        }
        // Newton from the left end: lo is below every root, the iteration increases monotonically to the smallest one
        if (tid == 0) {
            double x = lo;
            for (int itn = 0; itn < 4; ++itn) {  // polish: quadratic for a simple root, harmless (stays inside the bracket) otherwise
                double p, dp;
                lc_poly_newton(sm.sal, sm.sb2, k, x, p, dp);
                if (!(dp != 0.0)) break;
                const double xn = x - p / dp;
                if (!(xn > x) || xn > hi) break;  // converged to rounding (or left the bracket: keep the last safe iterate)
                x = xn;
            }
            a.scal[(long long)env * 4 + 0] = gl + x * (gu - gl);
            a.scal[(long long)env * 4 + 1] = gu;  // upper bound of the spectrum of T (selects the approximation interval only)
        }
    }
    DENSE_STAMP(51);
    gjb_cluster_sync();  // nobody leaves while a peer could still be sending to it
}

// ---------------------------------------------------------------------------------------------------------------------------
// D2'' gjb_inverse_kernel (COVO_SIGMA=dense-gjb): the same in-place Gauss-Jordan sweep, BLOCKED (8 pivots per step) and spread over
// a 2-CTA cluster per pole, the matrix resident in registers.  Sweeping the index block K with P = A_KK:
//     A_IJ -= A_IK P^-1 A_KJ,   A_KJ <- P^-1 A_KJ =: G,   A_IK <- -A_IK P^-1,   A_KK <- P^-1.
// With the sign convention of D2' the matrix is symmetric on the unswept index set and ANTI-symmetric between swept and unswept
// indices, so the column panel is the row panel again: A_iK = sigma(i) (A_Ki)^T, sigma = -1 for swept i.  Everything a step needs
// therefore follows from the 8 raw pivot rows (8 x n) alone:
//     A_ij -= sigma(i) sum_s raw[s][i] G[s][j],    row K_s <- G[s][:],   column K_s <- -sigma(i) G[s][i],   block KK <- P^-1.
// Roles (640 threads per CTA):
//   * 16 UPDATE warps hold the matrix: thread (ty = warp, tx = lane) owns rows ty + 16 (rank + 2 a'), a' < 8 (row pairs packed for
//     FFMA2) and columns tx + 32 b, b < 7.  Per step: 224 FFMA2 per thread against 28 + 16 vector loads.
//   * 4 SOLVER warps run one block ahead: they wait for the raw rows of block m + 1 (published through distributed shared memory with
//     st.async, completion counted by an mbarrier -- no fences, no cluster barrier), invert the 8 x 8 pivot block on one warp
//     (lane = two entries, 8 shuffle-driven pivots), build G = P^-1 raw and the signed multiplier table, while the update warps
//     are still applying block m.
//   * look-ahead: after the barrier that opens step m, the warps that own the rows of block m + 1 (one row per warp) apply step m to
//     that row first (56 FMAs, same operation order as the full update, so the values are bit-identical) and publish it.
// Flow control: the raw panels live in a ring of 4 slots.  Blocks are owned in pairs (tile = 16 rows = 2 blocks, tiles alternate
// between the CTAs), so a CTA can run at most 3 blocks ahead of its peer before it needs a panel from it: 4 slots never collide.
// Cost model at n = 200: 25 steps x max(update 1800 cycles, solver chain ~1600) instead of 200 steps x 1100.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int GB_CL = 2;     // CTAs per matrix
constexpr int GB_UT = 512;   // update threads
constexpr int GB_ST = 128;   // solver threads
constexpr int GB_T = GB_UT + GB_ST;
constexpr int GB_NP = 224;   // padded order: 14 row tiles of 16, 7 column slots of 32
constexpr int GB_RS = 256;   // row stride of a raw panel (the unused 8th row slot addresses rows up to 255)
constexpr int GB_SLOTS = 4;

template <int V>
struct IntC {
    static constexpr int value = V;
};

struct GjbSmem {
    float raw[GB_SLOTS][8][GB_RS];  // pivot-row panels [s][j]
    float stage[2][8][GB_NP];       // a pivot row on its way out (by block parity): source of the bulk copies
    float G[2][4][GB_NP][2];        // [parity][s / 2][j][s & 1]
    float2 Mneg[2][16][4][8];       // [parity][ty][row pair q][s]: (-sigma(i0) raw[s][i0], -sigma(i1) raw[s][i1]); 0 for pivot rows
    float Pinv[2][64];
    float piv[GB_NP];
    unsigned long long rawbar[GB_SLOTS];
    int bad;
};


__global__ void __launch_bounds__(GB_T, 1) gjb_inverse_kernel(const DenseArgs a) {
    COVO_DYN_SMEM(smraw);
    GjbSmem& sm = *reinterpret_cast<GjbSmem*>(smraw);
    const int n = a.n, tid = threadIdx.x, env = blockIdx.y, pole = blockIdx.x / GB_CL;
    const int rank = (int)gjb_rank();
    const int nblk = (n + 7) >> 3;
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
    const double* zt = a.zolo + (size_t)lad * 2 * kZoloPoles;
    const bool want_logdet = pole == kZoloPoles;
    const double shift = (kOffset - lam_min) + (want_logdet ? 0.0 : zt[pole]);
    const float wj = want_logdet ? 0.f : (float)zt[kZoloPoles + pole];
    if (tid == 0) {
        sm.bad = 0;
        for (int q = 0; q < GB_SLOTS; ++q) gjb_mbar_init(&sm.rawbar[q], 1);
#if !defined(COVO_CPU_EMU)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    for (int i = tid; i < GB_SLOTS * 8 * GB_RS; i += GB_T) (&sm.raw[0][0][0])[i] = 0.f;
    gjb_cluster_sync();  // barriers initialised and panels zeroed everywhere before anybody publishes

    if (tid >= GB_UT) {
        // ================================================ solver warps ================================================
        const int sidx = tid - GB_UT, lane = sidx & 31, swarp = sidx >> 5;
        for (int m = 0; m < nblk; ++m) {
            const int slot = m & (GB_SLOTS - 1), par = m & 1, K0 = 8 * m;
            if (sidx == 0) gjb_mbar_expect(&sm.rawbar[slot], 8 * GB_NP * 4);
            const bool pf = a.prof && sidx == 0 && blockIdx.x == 0 && blockIdx.y == 0;
            long long t0 = pf ? clock64() : 0;
            gjb_mbar_wait(&sm.rawbar[slot], (unsigned)((m / GB_SLOTS) & 1));
            if (pf) {
                const long long t1 = clock64();
                a.prof[54] = (m == 0 ? 0 : a.prof[54]) + (t1 - t0);
                t0 = t1;
            }
            const float(*rw)[GB_RS] = sm.raw[slot];
            if (swarp == 0) {
                // P^-1 by an in-place Gauss-Jordan sweep of the 8 x 8 block: lane = (row r, columns c0, c0 + 1)
                const int r = lane >> 2, c0 = (lane & 3) * 2;
                float x0 = rw[r][K0 + c0], x1 = rw[r][K0 + c0 + 1];
#pragma unroll
                for (int sp = 0; sp < 8; ++sp) {
                    const float mine = (sp & 1) ? x1 : x0;
                    const float p = __shfl_sync(0xffffffffu, mine, (sp << 2) | (sp >> 1));
                    const float prs = __shfl_sync(0xffffffffu, mine, (r << 2) | (sp >> 1));   // P[r][sp]
                    const float ps0 = __shfl_sync(0xffffffffu, x0, (sp << 2) | (lane & 3));   // P[sp][c0]
                    const float ps1 = __shfl_sync(0xffffffffu, x1, (sp << 2) | (lane & 3));   // P[sp][c0 + 1]
                    if (lane == 0) {
                        sm.piv[K0 + sp] = p;
                        if (!(p > 0.f)) sm.bad = 1;
                    }
                    const float rinv = gjb_rcp(p);
                    if (r == sp) {
                        x0 = (c0 == sp) ? rinv : ps0 * rinv;
                        x1 = (c0 + 1 == sp) ? rinv : ps1 * rinv;
                    } else {
                        const float f = prs * rinv;
                        x0 = (c0 == sp) ? -f : fmaf(-f, ps0, x0);
                        x1 = (c0 + 1 == sp) ? -f : fmaf(-f, ps1, x1);
                    }
                }
                sm.Pinv[par][r * 8 + c0] = x0;
                sm.Pinv[par][r * 8 + c0 + 1] = x1;
            } else {
                // signed, negated, pair-packed multipliers of this CTA's rows
                for (int e = sidx - 32; e < 16 * 4 * 8; e += GB_ST - 32) {
                    const int ty = e >> 5, q = (e >> 3) & 3, sp = e & 7;
                    float2 val;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int tile = rank + GB_CL * (2 * q + h), i = ty + 16 * tile;
                        float x = 0.f;
                        if (tile < GB_NP / 16 && !(i >= K0 && i < K0 + 8)) x = (i < K0) ? rw[sp][i] : -rw[sp][i];
                        if (h) val.y = x;
                        else val.x = x;
                    }
                    sm.Mneg[par][ty][q][sp] = val;
                }
            }
            COVO_NAMED_BARRIER(1, GB_ST);
            if (pf) {
                const long long t1 = clock64();
                a.prof[55] = (m == 0 ? 0 : a.prof[55]) + (t1 - t0);
                t0 = t1;
            }
            {
                // G = P^-1 raw: thread = columns (sidx, sidx + 128) packed for FFMA2; P^-1 rows come as broadcast float4 loads
                const int j0 = sidx, j1 = min(sidx + GB_ST, GB_RS - 1);  // (the second column of sidx >= 96 is padding: never stored)
                float2 col[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) col[t] = make_float2(rw[t][j0], rw[t][j1]);
                const float4* pv = reinterpret_cast<const float4*>(sm.Pinv[par]);
#pragma unroll
                for (int sp = 0; sp < 8; sp += 2) {
                    float2 g[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float4 pa = pv[(sp + u) * 2], pb = pv[(sp + u) * 2 + 1];
                        float2 acc2 = make_float2(0.f, 0.f);
                        acc2 = __ffma2_rn(make_float2(pa.x, pa.x), col[0], acc2);
                        acc2 = __ffma2_rn(make_float2(pa.y, pa.y), col[1], acc2);
                        acc2 = __ffma2_rn(make_float2(pa.z, pa.z), col[2], acc2);
                        acc2 = __ffma2_rn(make_float2(pa.w, pa.w), col[3], acc2);
                        acc2 = __ffma2_rn(make_float2(pb.x, pb.x), col[4], acc2);
                        acc2 = __ffma2_rn(make_float2(pb.y, pb.y), col[5], acc2);
                        acc2 = __ffma2_rn(make_float2(pb.z, pb.z), col[6], acc2);
                        acc2 = __ffma2_rn(make_float2(pb.w, pb.w), col[7], acc2);
                        g[u] = acc2;
                    }
                    // [s / 2][j][s & 1]: the two s of this pair are adjacent
                    *reinterpret_cast<float2*>(&sm.G[par][sp >> 1][j0][0]) = make_float2(g[0].x, g[1].x);
                    if (sidx + GB_ST < GB_NP) *reinterpret_cast<float2*>(&sm.G[par][sp >> 1][sidx + GB_ST][0]) = make_float2(g[0].y, g[1].y);
                }
            }
            if (pf) {
                const long long t1 = clock64();
                a.prof[56] = (m == 0 ? 0 : a.prof[56]) + (t1 - t0);
                t0 = t1;
            }
            __syncthreads();  // opens step m for the update warps
            if (pf) a.prof[57] = (m == 0 ? 0 : a.prof[57]) + (clock64() - t0);
        }
    } else {
        // ================================================ update warps ================================================
        const int tx = tid & 31, ty = tid >> 5;
        const float* Ag = a.Asym ? a.Asym + (long long)env * n * n : nullptr;
        const float* Rg = a.R + (long long)env * n * n;
        float2 acc[4][7];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int b = 0; b < 7; ++b) {
                float v2[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int tile = rank + GB_CL * (2 * q + h), i = ty + 16 * tile, j = tx + 32 * b;
                    float v = (i == j && tile < GB_NP / 16) ? 1.f : 0.f;  // identity padding: never coupled, pivots 1
                    if (i < n && j < n) {
                        v = Ag ? Ag[(long long)i * n + j] : 0.5f * (Rg[(long long)i * n + j] + Rg[(long long)j * n + i]);
                        if (i == j) v = (float)((double)v + shift);
                    }
                    v2[h] = v;
                }
                acc[q][b] = make_float2(v2[0], v2[1]);
            }
        // block 0 lives in tile 0 = CTA 0, row slot 0 (.x of pair 0): warps 0..7 publish their row
        // a row leaves as ONE bulk copy per destination CTA (this one included): staged in shared memory, fenced for the async proxy
        auto publish_row = [&](int blk, int s_row, const float (&vals)[7]) {
            float* st = sm.stage[blk & 1][s_row];
#pragma unroll
            for (int b = 0; b < 7; ++b) st[tx + 32 * b] = vals[b];
            gjb_fence_async_proxy();
            __syncwarp();
            if (tx == 0) {
                const int slot = blk & (GB_SLOTS - 1);
#pragma unroll
                for (int r = 0; r < GB_CL; ++r) gjb_bulk_send(&sm.raw[slot][s_row][0], st, GB_NP * 4, (unsigned)r, &sm.rawbar[slot]);
            }
        };
        if (rank == 0 && ty < 8) {
            float vals[7];
#pragma unroll
            for (int b = 0; b < 7; ++b) vals[b] = acc[0][b].x;
            publish_row(0, ty, vals);
        }
        const bool pfu = a.prof && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
        long long tu0 = pfu ? clock64() : 0;
        for (int m = 0; m < nblk; ++m) {
            __syncthreads();  // G, multipliers and P^-1 of block m are in place; everybody is done with step m - 1
            if (pfu) {
                const long long t1 = clock64();
                a.prof[58] = (m == 0 ? 0 : a.prof[58]) + (t1 - tu0);
                tu0 = t1;
            }
            const int par = m & 1, K0 = 8 * m;
            const float2* mrow = &sm.Mneg[par][ty][0][0];
            // One row pair: all eight s of the rank-8 update (G is re-read per pair here; the bulk below re-uses it across pairs)
            auto update_pair = [&](auto QC) {
                constexpr int q = decltype(QC)::value;
#pragma unroll
                for (int qt = 0; qt < 4; ++qt) {
                    const float4 m4 = *reinterpret_cast<const float4*>(&mrow[q * 8 + 2 * qt]);
                    const float2 ma = make_float2(m4.x, m4.y), mb = make_float2(m4.z, m4.w);
#pragma unroll
                    for (int b = 0; b < 7; ++b) {
                        const float2 g2 = *reinterpret_cast<const float2*>(&sm.G[par][qt][tx + 32 * b][0]);
                        acc[q][b] = __ffma2_rn(ma, make_float2(g2.x, g2.x), acc[q][b]);
                        acc[q][b] = __ffma2_rn(mb, make_float2(g2.y, g2.y), acc[q][b]);
                    }
                }
            };
            // ---- look-ahead: the warps that own the rows of block m + 1 update THAT pair first and publish their row -------------
            int q_done = -1;
            if (m + 1 < nblk) {
                const int tile1 = (m + 1) >> 1;
                if (rank == tile1 % GB_CL && (ty >> 3) == ((m + 1) & 1)) {
                    const int a1 = tile1 / GB_CL, q1 = a1 >> 1, h1 = a1 & 1, i = ty + 16 * tile1;
                    q_done = q1;
                    if (q1 == 0) update_pair(IntC<0>());
                    else if (q1 == 1) update_pair(IntC<1>());
                    else if (q1 == 2) update_pair(IntC<2>());
                    else update_pair(IntC<3>());
                    float tmp[7];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q == q1) {
#pragma unroll
                            for (int b = 0; b < 7; ++b) tmp[b] = h1 ? acc[q][b].y : acc[q][b].x;
                        }
                    if ((tx & ~7) == (K0 & 31)) {  // its entries in the pivot columns of step m: -G[s][i]  (i is unswept)
                        const int sc = tx - (K0 & 31), b0 = K0 >> 5;
                        const float v = -sm.G[par][sc >> 1][i][sc & 1];
#pragma unroll
                        for (int b = 0; b < 7; ++b)
                            if (b == b0) tmp[b] = v;
                    }
                    publish_row(m + 1, ty & 7, tmp);
                }
            }
            // ---- the rank-8 update of everything else this thread owns ---------------------------------------------------
#pragma unroll
            for (int qt = 0; qt < 4; ++qt) {  // two of the eight s at a time
                float2 g2[7];
#pragma unroll
                for (int b = 0; b < 7; ++b) g2[b] = *reinterpret_cast<const float2*>(&sm.G[par][qt][tx + 32 * b][0]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q == q_done) continue;  // warp-uniform
                    const float4 m4 = *reinterpret_cast<const float4*>(&mrow[q * 8 + 2 * qt]);
                    const float2 ma = make_float2(m4.x, m4.y), mb = make_float2(m4.z, m4.w);
#pragma unroll
                    for (int b = 0; b < 7; ++b) {
                        acc[q][b] = __ffma2_rn(ma, make_float2(g2[b].x, g2[b].x), acc[q][b]);
                        acc[q][b] = __ffma2_rn(mb, make_float2(g2[b].y, g2[b].y), acc[q][b]);
                    }
                }
            }
            // ---- fix-ups: pivot rows <- G (P^-1 inside the block), pivot columns <- -sigma(i) G[s][i] ---------------------
            const int tile0 = m >> 1;
            const bool pivot_warp = (rank == tile0 % GB_CL) && ((ty >> 3) == (m & 1));
            const int q0 = (tile0 / GB_CL) >> 1, h0 = (tile0 / GB_CL) & 1;
            if (pivot_warp) {
                const int sr = ty & 7;
#pragma unroll
                for (int b = 0; b < 7; ++b) {
                    const int j = tx + 32 * b;
                    const float v = (j >= K0 && j < K0 + 8) ? sm.Pinv[par][sr * 8 + (j - K0)] : sm.G[par][sr >> 1][j][sr & 1];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q == q0) {
                            if (h0) acc[q][b].y = v;
                            else acc[q][b].x = v;
                        }
                }
            }
            if ((tx & ~7) == (K0 & 31)) {
                const int sc = tx - (K0 & 31), b0 = K0 >> 5;
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int tile = rank + GB_CL * (2 * q + h), i = ty + 16 * tile;
                        if (tile < GB_NP / 16 && !(i >= K0 && i < K0 + 8)) {
                            const float g = sm.G[par][sc >> 1][i][sc & 1];
                            const float v = (i < K0) ? g : -g;
#pragma unroll
                            for (int b = 0; b < 7; ++b)
                                if (b == b0) {
                                    if (h) acc[q][b].y = v;
                                    else acc[q][b].x = v;
                                }
                        }
                    }
            }
            if (pfu) {
                const long long t1 = clock64();
                a.prof[59] = (m == 0 ? 0 : a.prof[59]) + (t1 - tu0);
                tu0 = t1;
            }
        }
        // ---- results ---------------------------------------------------------------------------------------------------------
        if (!want_logdet) {
            float* Xg = a.Xbuf + ((long long)env * kZoloPoles + pole) * n * n;
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int b = 0; b < 7; ++b) {
                    const int j = tx + 32 * b;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int tile = rank + GB_CL * (2 * q + h), i = ty + 16 * tile;
                        if (tile < GB_NP / 16 && i < n && j <= i) Xg[(long long)i * n + j] = wj * (h ? acc[q][b].y : acc[q][b].x);
                    }
                }
        } else if (rank == 0 && ty < 7) {  // log det A = sum of the logarithms of the scalar pivots (both CTAs hold all of them)
            double lp = 0.0;
            const int i = tx + 32 * ty;
            if (i < n) lp = log((double)sm.piv[i]);
            lp = warp_sum_d(lp);
            double* red = reinterpret_cast<double*>(&sm.G[0][0][0][0]);  // dead: every step is over for these warps' inputs
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
    gjb_cluster_sync();  // nobody leaves while a peer could still be sending to it
}

// ---------------------------------------------------------------------------------------------------------------------------
// D3
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) combine_kernel(const DenseArgs a) {
    const int n = a.n, env = blockIdx.y;
    const double logdet = a.scal[(long long)env * 4 + 2];
    // controllers/covo.py:123-127: log_const = (2 * n * 2 log(sigma) + sum log o) / n;  Sigma = exp(log_const / 2) A^(-1/2)
    const double log_const = (4.0 * (double)n * log((double)a.sample_sigma) + logdet) / (double)n;
    const float scale = (float)exp(0.5 * log_const);
    const float* Xg = a.Xbuf + (long long)env * kZoloPoles * n * n;
    float* cov = a.cov + (long long)env * n * n;
    const int npairs = n * (n + 1) / 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < npairs; q += gridDim.x * blockDim.x) {
        int ia = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
        while (ia * (ia + 1) / 2 > q) --ia;
        while ((ia + 1) * (ia + 2) / 2 <= q) ++ia;
        const int ib = q - ia * (ia + 1) / 2;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kZoloPoles; ++j) s += Xg[(long long)j * n * n + ia * n + ib];
        s *= scale;
        cov[ia * n + ib] = s;
        cov[ib * n + ia] = s;  // (a_cov + a_cov.T)/2 (:132) holds by construction
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
size_t sigma_dense_scratch_floats(int n) { return (size_t)kZoloPoles * n * n; }

#if !defined(COVO_CPU_EMU)
cudaError_t launch_sigma_dense(const SigmaArgs& s, double* scal, float* Xbuf, int n_env, cudaStream_t st, int variant) {
    if (s.n > kSigmaMaxN || (s.n & 3)) return cudaErrorInvalidValue;
    DenseArgs a;
    a.n = s.n;
    a.n_pad = s.n_pad;
    a.sample_sigma = s.sample_sigma;
    a.R = s.R;
    a.scal = scal;
    a.Xbuf = Xbuf;
    a.cov = s.cov;
    a.zolo = s.zolo;
    a.status = s.status;
    a.prof = s.prof;
    static size_t conf1[32] = {}, conf2[32] = {}, conf3[32] = {};
    cudaError_t e;
    // Lanczos kernel: the 8-CTA cluster version by default; COVO_LANCZOS=v1 | v2 select the single-CTA kernels (development)
    const char* lz = getenv("COVO_LANCZOS");
    const int lz_kind = (lz && !strcmp(lz, "v1")) ? 1 : (lz && !strcmp(lz, "v2") && lanczos2_layout(a.n).fits) ? 2 : 3;
    a.Asym = (lz_kind == 1) ? nullptr : s.F;  // the symmetrised matrix, written by the Lanczos kernel (F is unused on this path)
    if (lz_kind == 1) {
        const size_t smem1 = (size_t)(2 * a.n + 8 + 2 * kLanczosMax) * sizeof(double) + (size_t)a.n * (a.n + 1) * sizeof(float);
        e = ensure_smem_attr(lanczos_kernel, smem1, conf1);
        if (e != cudaSuccess) return e;
        lanczos_kernel<<<n_env, TL, smem1, st>>>(a);
    } else if (lz_kind == 2) {
        const size_t smem3 = lanczos2_layout(a.n).bytes;
        e = ensure_smem_attr(lanczos2_kernel, smem3, conf3);
        if (e != cudaSuccess) return e;
        lanczos2_kernel<<<n_env, TL2, smem3, st>>>(a);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(LC_CL, n_env);
        cfg.blockDim = dim3(LC_T);
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
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (variant == 3) {  // blocked Gauss-Jordan on a 2-CTA cluster per pole
        static size_t conf4[32] = {};
        e = ensure_smem_attr(gjb_inverse_kernel, sizeof(GjbSmem), conf4);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(GB_CL * (kZoloPoles + 1), n_env);
        cfg.blockDim = dim3(GB_T);
        cfg.dynamicSmemBytes = sizeof(GjbSmem);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = GB_CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, gjb_inverse_kernel, a);
        if (e != cudaSuccess) return e;
    } else if (variant == 2) {  // register-resident Gauss-Jordan
        gj_inverse_kernel<14><<<dim3(kZoloPoles + 1, n_env), TG, 1024 * sizeof(float), st>>>(a);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    } else {
        const size_t smem2 = ((size_t)a.n * a.n + 2 * 8 * a.n_pad + 64 + a.n_pad) * sizeof(float) + 16;
        e = ensure_smem_attr(shifted_inverse_kernel, smem2, conf2);
        if (e != cudaSuccess) return e;
        shifted_inverse_kernel<<<dim3(kZoloPoles + 1, n_env), TD, smem2, st>>>(a);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const int npairs = a.n * (a.n + 1) / 2;
    combine_kernel<<<dim3((npairs + 255) / 256, n_env), 256, 0, st>>>(a);
    return cudaGetLastError();
}

#endif

}  // namespace covo
