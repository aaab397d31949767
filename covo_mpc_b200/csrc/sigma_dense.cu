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

#include "common.cuh"
#include "sigma.cuh"

namespace covo {

namespace {

constexpr double kOffset = 1e-2;  // controllers/covo.py:121
constexpr int kLanczosMax = 32;
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
    static size_t conf1[32] = {}, conf2[32] = {};
    const size_t smem1 = (size_t)(2 * a.n + 8 + 2 * kLanczosMax) * sizeof(double) + (size_t)a.n * (a.n + 1) * sizeof(float);
    cudaError_t e = ensure_smem_attr(lanczos_kernel, smem1, conf1);
    if (e != cudaSuccess) return e;
    lanczos_kernel<<<n_env, TL, smem1, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (variant == 2) {  // register-resident Gauss-Jordan
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
