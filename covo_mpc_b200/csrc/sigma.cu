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

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int TT = 1024;
// shared region holding the matrix during E1 and the fp64 pivot scratch + (gd, cu) during E2
__host__ __device__ inline int tridiag_region_floats(int n) {
    int a = n * n, b = (2 * 2 * (kZoloPoles + 1) * n) * 2 + 2 * kZoloPoles * n;
    return ((a > b ? a : b) + 3) & ~3;
}
constexpr int NPOLE = kZoloPoles;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
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

// number of eigenvalues of tridiag(d, e) strictly below x: sign changes of the Sturm sequence,
// evaluated division-free with rescaling (a zero term takes the sign opposite to its predecessor).
__device__ int sturm_count(const double* d, const double* e, int n, double x) {
    double pm = 1.0, p = d[0] - x;
    if (p == 0.0) p = -1e-300;
    int cnt = p < 0.0;
    for (int i = 1; i < n; ++i) {
        double e2 = e[i - 1] * e[i - 1];
        double pn = (d[i] - x) * p - e2 * pm;
        if (pn == 0.0) pn = (p > 0.0) ? -1e-300 : 1e-300;
        cnt += ((pn < 0.0) != (p < 0.0));
        double ap = fabs(pn);
        if (ap > 1e200) {
            pn *= 1e-200;
            p *= 1e-200;
        } else if (ap < 1e-200) {
            pn *= 1e200;
            p *= 1e200;
        }
        pm = p;
        p = pn;
    }
    return cnt;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// E1 + E2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TT, 1) sigma_tridiag_kernel(const SigmaArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int n = a.n, tid = threadIdx.x, env = blockIdx.x, lane = tid & 31, warp = tid >> 5;
    float* As = reinterpret_cast<float*>(smraw);  // [n][n]
    float* part = As + tridiag_region_floats(n);  // [4096]
    float* vs = part + 4096;                      // [256] zero-extended Householder vector
    float* ws = vs + 256;                         // [256]
    double* dd = reinterpret_cast<double*>(ws + 256);  // [256]
    double* ee = dd + 256;                             // [256]
    double* sc = ee + 256;                             // [16] scalars
    float* red = reinterpret_cast<float*>(sc + 16);    // [64]
    int* ired = reinterpret_cast<int*>(red + 64);      // [4]

    const float* Rg = a.R + (long long)env * n * n;
    float* Vg = a.Vh + (long long)env * n * n;
    float* taug = a.tau + (long long)env * n;

    for (int idx = tid; idx < n * n; idx += TT) {
        int i = idx / n, j = idx - i * n;
        As[idx] = 0.5f * (Rg[idx] + Rg[j * n + i]);  // R <- (R + R^T)/2, controllers/covo.py:117
    }
    if (tid < 256) {
        vs[tid] = 0.f;
        ws[tid] = 0.f;
    }
    __syncthreads();

    // ---- E1: Householder tridiagonalisation (LAPACK ssytd2 recurrences) -----------------------
    for (int k = 0; k < n - 2; ++k) {
        if (warp == 0) {
            const float* x = As + k * n;
            float s = 0.f;
            for (int j = k + 2 + lane; j < n; j += 32) s = fmaf(x[j], x[j], s);
            s = wsum(s);
            float x0 = x[k + 1];
            float beta, tau, scale;
            if (s == 0.f) {
                beta = x0;
                tau = 0.f;
                scale = 0.f;
            } else {
                beta = -copysignf(sqrtf(fmaf(x0, x0, s)), x0);
                tau = (beta - x0) / beta;
                scale = 1.0f / (x0 - beta);
            }
            if (lane == 0) {
                red[0] = beta;
                red[1] = tau;
                red[2] = scale;
            }
        }
        __syncthreads();
        const float beta = red[0], tau = red[1], scale = red[2];
        if (tid < n) {
            float v = (tid <= k) ? 0.f : ((tid == k + 1) ? 1.f : As[k * n + tid] * scale);
            if (tau == 0.f && tid > k + 1) v = 0.f;
            vs[tid] = v;
            Vg[k * n + tid] = v;
        }
        if (tid == 0) {
            dd[k] = (double)As[k * n + k];
            ee[k] = (double)beta;
            taug[k] = tau;
        }
        __syncthreads();
        if (tau != 0.f) {  // block-uniform
            const int c0 = (k + 1) & ~3;
            const int ncg = (n - c0) >> 2;
            const int m = n - (k + 1);
            const int chunks = min(TT / ncg, m);
            const int rows_per = (m + chunks - 1) / chunks;
            const int cg = tid % ncg, ch = tid / ncg;
            const int i0 = k + 1 + ch * rows_per, i1 = min(i0 + rows_per, n);
            // p = tau * A v   (column partials over row chunks; A symmetric)
            if (ch < chunks) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int i = i0; i < i1; ++i) {
                    const float4 av = *reinterpret_cast<const float4*>(As + i * n + c0 + 4 * cg);
                    const float vi = vs[i];
                    acc.x = fmaf(av.x, vi, acc.x);
                    acc.y = fmaf(av.y, vi, acc.y);
                    acc.z = fmaf(av.z, vi, acc.z);
                    acc.w = fmaf(av.w, vi, acc.w);
                }
                *reinterpret_cast<float4*>(part + (ch * ncg + cg) * 4) = acc;
            }
            __syncthreads();
            float pj = 0.f, contrib = 0.f;
            if (tid < n && tid > k) {
                const int g = (tid - c0) >> 2, comp = (tid - c0) & 3;
                float s = 0.f;
                for (int c = 0; c < chunks; ++c) s += part[(c * ncg + g) * 4 + comp];
                pj = tau * s;
                contrib = pj * vs[tid];
            }
            contrib = wsum(contrib);
            if (lane == 0) red[8 + warp] = contrib;
            __syncthreads();
            float alpha = 0.f;
#pragma unroll
            for (int w = 0; w < TT / 32; ++w) alpha += red[8 + w];
            if (tid < n) ws[tid] = (tid > k) ? pj - 0.5f * tau * alpha * vs[tid] : 0.f;
            __syncthreads();
            // A <- A - v w^T - w v^T
            if (ch < chunks) {
                const float4 wj = *reinterpret_cast<const float4*>(ws + c0 + 4 * cg);
                const float4 vj = *reinterpret_cast<const float4*>(vs + c0 + 4 * cg);
                for (int i = i0; i < i1; ++i) {
                    const float vi = vs[i], wi = ws[i];
                    float4* p = reinterpret_cast<float4*>(As + i * n + c0 + 4 * cg);
                    float4 av = *p;
                    av.x -= fmaf(vi, wj.x, wi * vj.x);
                    av.y -= fmaf(vi, wj.y, wi * vj.y);
                    av.z -= fmaf(vi, wj.z, wi * vj.z);
                    av.w -= fmaf(vi, wj.w, wi * vj.w);
                    *p = av;
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        dd[n - 2] = (double)As[(n - 2) * n + (n - 2)];
        dd[n - 1] = (double)As[(n - 1) * n + (n - 1)];
        ee[n - 2] = (double)As[(n - 1) * n + (n - 2)];
        ee[n - 1] = 0.0;
        taug[n - 2] = 0.f;
        taug[n - 1] = 0.f;
    }
    if (tid < n) {
        Vg[(n - 2) * n + tid] = 0.f;
        Vg[(n - 1) * n + tid] = 0.f;
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
        lo = wmind(lo);
        hi = wmaxd(hi);
        double* rd = reinterpret_cast<double*>(part);
        if (lane == 0) {
            rd[warp] = lo;
            rd[32 + warp] = hi;
        }
        __syncthreads();
        if (tid == 0) {
            double l = rd[0], h = rd[32];
            for (int w = 1; w < TT / 32; ++w) {
                l = fmin(l, rd[w]);
                h = fmax(h, rd[32 + w]);
            }
            double pad = 1e-9 * fmax(1.0, fmax(fabs(l), fabs(h)));
            sc[0] = l - pad;
            sc[1] = h + pad;
            sc[2] = l;
            sc[3] = h;
        }
        __syncthreads();
    }
    // lam_min by multisection: 5 rounds x 1025-fold shrink
    for (int round = 0; round < 5; ++round) {
        const double lo = sc[0], hi = sc[1];
        if (tid == 0) ired[0] = TT;
        __syncthreads();
        const double x = lo + (hi - lo) * ((double)(tid + 1) / (double)(TT + 1));
        if (sturm_count(dd, ee, n, x) >= 1) atomicMin(ired, tid);
        __syncthreads();
        if (tid == 0) {
            const int ts = ired[0];
            const double step = (hi - lo) / (double)(TT + 1);
            sc[0] = (ts == 0) ? lo : lo + step * ts;
            sc[1] = (ts == TT) ? hi : lo + step * (ts + 1);
        }
        __syncthreads();
    }
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
    double* num = reinterpret_cast<double*>(As);        // [2][17][n]
    double* den = num + 2 * (NPOLE + 1) * n;            // [2][17][n]
    float* gd = reinterpret_cast<float*>(den + 2 * (NPOLE + 1) * n);  // [16][n]  diag of (T_s+t_q)^-1
    float* cu = gd + NPOLE * n;                                        // [16][n]  -b_l / dm_{l+1}
    if (tid < NPOLE + 1 || (tid >= 32 && tid < 32 + NPOLE + 1)) {
        const bool fwd = tid < 32;
        const int q = fwd ? tid : tid - 32;
        const double tq = (q < NPOLE) ? zt[q] : 0.0;
        double* nm = num + ((fwd ? 0 : 1) * (NPOLE + 1) + q) * n;
        double* dn = den + ((fwd ? 0 : 1) * (NPOLE + 1) + q) * n;
        double pm = 0.0, p = 1.0;
        for (int s = 0; s < n; ++s) {
            const int i = fwd ? s : n - 1 - s;
            const double ai = dd[i] + shift0 + tq;
            double b2 = 0.0;
            if (s > 0) {
                const double b = fwd ? ee[i - 1] : ee[i];
                b2 = b * b;
            }
            double pn = ai * p - b2 * pm;
            nm[i] = pn;
            dn[i] = p;
            const double ap = fabs(pn);
            if (ap > 1e200) {
                pn *= 1e-200;
                p *= 1e-200;
            } else if (ap < 1e-200) {
                pn *= 1e200;
                p *= 1e200;
            }
            pm = p;
            p = pn;
        }
    }
    __syncthreads();
    // log det T_s = sum log dp_i (unshifted forward pivots)
    {
        double l = 0.0;
        if (tid < n) l = log(num[(0 * (NPOLE + 1) + NPOLE) * n + tid] / den[(0 * (NPOLE + 1) + NPOLE) * n + tid]);
        l = wsumd(l);
        double* rd = reinterpret_cast<double*>(part);
        if (lane == 0) rd[warp] = l;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < TT / 32; ++w) s += rd[w];
            sc[4] = s;
        }
    }
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
        cu[idx] = c;
    }
    __syncthreads();
    // F = exp(log_const / 2) * sum_q w_q (T_s + t_q)^-1, written straight to HBM (rows i and n-1-i per thread)
    {
        const double logdet = sc[4];
        // controllers/covo.py:124-128: log_const = (n * 2 log(sigma) * 2 + sum log o) / n
        const double log_const = (4.0 * n * log((double)a.sample_sigma) + logdet) / (double)n;
        const float cscale = (float)exp(0.5 * log_const);
        float* Fg = a.F + (long long)env * n * n;
        float wq[NPOLE];
#pragma unroll
        for (int q = 0; q < NPOLE; ++q) wq[q] = (float)zw[q] * cscale;
        if (tid < n) {
            // thread t < n/2 takes row t, thread t >= n/2 takes row n-1-(t-n/2): balances the chain lengths per warp
            const int i = (tid < n / 2) ? tid : (n - 1 - (tid - n / 2));
            float prod[NPOLE];
#pragma unroll
            for (int q = 0; q < NPOLE; ++q) prod[q] = gd[q * n + i];
            for (int l = i; l < n; ++l) {
                float f = 0.f;
#pragma unroll
                for (int q = 0; q < NPOLE; ++q) f = fmaf(wq[q], prod[q], f);
                Fg[(long long)i * n + l] = f;
                Fg[(long long)l * n + i] = f;
#pragma unroll
                for (int q = 0; q < NPOLE; ++q) prod[q] *= cu[q * n + l];
            }
        }
        if (tid == 0) {
            double* dg = a.diag + (long long)env * 4 * n;
            dg[2 * n + 0] = lam_min;
            dg[2 * n + 1] = sc[2];
            dg[2 * n + 2] = sc[3];
            dg[2 * n + 3] = logdet;
            dg[2 * n + 4] = (double)lad;
        }
        if (tid < n) {
            double* dg = a.diag + (long long)env * 4 * n;
            dg[tid] = dd[tid];
            dg[n + tid] = ee[tid];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// E3: out(:, c) = Q in(c, :)^T, one warp per vector; Q = H_0 H_1 ... H_{n-3}
// ---------------------------------------------------------------------------------------------
constexpr int kApplyWarps = 8;
constexpr int kMaxLi = kSigmaMaxN / 32;  // 7

template <bool STORE_STRIDED>
__global__ void __launch_bounds__(kApplyWarps * 32) applyq_kernel(const float* __restrict__ Vh,
                                                                 const float* __restrict__ tau,
                                                                 const float* __restrict__ in, float* __restrict__ out,
                                                                 int n) {
    const int env = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kApplyWarps + warp;
    if (c >= n) return;
    Vh += (long long)env * n * n;
    tau += (long long)env * n;
    in += (long long)env * n * n;
    out += (long long)env * n * n;
    float x[kMaxLi], v[kMaxLi], vn[kMaxLi];
#pragma unroll
    for (int li = 0; li < kMaxLi; ++li) {
        int i = lane + 32 * li;
        x[li] = (i < n) ? in[(long long)c * n + i] : 0.f;
    }
    int k = n - 3;
#pragma unroll
    for (int li = 0; li < kMaxLi; ++li) {
        int i = lane + 32 * li;
        v[li] = (i < n && k >= 0) ? __ldg(Vh + (long long)k * n + i) : 0.f;
    }
    for (; k >= 0; --k) {
        const float tk = __ldg(tau + k);
        if (k > 0) {
#pragma unroll
            for (int li = 0; li < kMaxLi; ++li) {
                int i = lane + 32 * li;
                vn[li] = (i < n) ? __ldg(Vh + (long long)(k - 1) * n + i) : 0.f;
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int li = 0; li < kMaxLi; ++li) dot = fmaf(v[li], x[li], dot);
        dot = wsum(dot) * tk;
#pragma unroll
        for (int li = 0; li < kMaxLi; ++li) x[li] = fmaf(-dot, v[li], x[li]);
#pragma unroll
        for (int li = 0; li < kMaxLi; ++li) v[li] = vn[li];
    }
#pragma unroll
    for (int li = 0; li < kMaxLi; ++li) {
        int i = lane + 32 * li;
        if (i < n) {
            if (STORE_STRIDED) out[(long long)i * n + c] = x[li];
            else out[(long long)c * n + i] = x[li];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// E4: Cholesky (blocked right-looking, NB = 8), one CTA per matrix
// ---------------------------------------------------------------------------------------------
constexpr int TC = 512;
__global__ void __launch_bounds__(TC, 1) cholesky_kernel(const SigmaArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int n = a.n, n_pad = a.n_pad, tid = threadIdx.x, env = blockIdx.x, lane = tid & 31, warp = tid >> 5;
    float* As = reinterpret_cast<float*>(smraw);  // [n][n]
    float* L11 = As + n * n;                      // [8][8]
    float* covg = a.cov + (long long)env * n * n;
    for (int idx = tid; idx < n * n; idx += TC) {
        int i = idx / n, j = idx - i * n;
        As[idx] = 0.5f * (covg[idx] + covg[j * n + i]);  // (a_cov + a_cov.T)/2, controllers/covo.py:132
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += TC) covg[idx] = As[idx];

    for (int jb = 0; jb < n; jb += 8) {
        const int nb = min(8, n - jb);
        // (a) factor the nb x nb diagonal block with one warp; lane r holds row r
        if (warp == 0) {
            float r8[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) r8[c] = (lane < nb && c < nb) ? As[(jb + lane) * n + jb + c] : 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float dcc = __shfl_sync(0xffffffffu, r8[c], c);
                if (c < nb) {
                    if (!(dcc > 0.f)) {
                        if (lane == 0) a.status[env] = 2;
                        dcc = 1e-30f;
                    }
                    float d = sqrtf(dcc);
                    if (lane == c) r8[c] = d;
                    else if (lane > c) r8[c] = r8[c] / d;
                }
#pragma unroll
                for (int c2 = c + 1; c2 < 8; ++c2) {
                    float l2 = __shfl_sync(0xffffffffu, r8[c], c2);
                    if (c2 < nb && lane >= c2) r8[c2] = fmaf(-r8[c], l2, r8[c2]);
                }
            }
            if (lane < nb) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c < nb) {
                        float v = (c <= lane) ? r8[c] : 0.f;
                        As[(jb + lane) * n + jb + c] = v;
                        L11[lane * 8 + c] = v;
                    }
                }
            }
        }
        __syncthreads();
        // (b) panel: L21 = A21 L11^-T, one thread per row
        for (int i = jb + nb + tid; i < n; i += TC) {
            float x[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = (c < nb) ? As[i * n + jb + c] : 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c < nb) {
                    float s = x[c];
#pragma unroll
                    for (int c2 = 0; c2 < c; ++c2) s = fmaf(-x[c2], L11[c * 8 + c2], s);
                    x[c] = s / L11[c * 8 + c];
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < nb) As[i * n + jb + c] = x[c];
        }
        __syncthreads();
        // (c) trailing update of the lower triangle in 4x4 tiles
        const int r0 = jb + nb;
        const int T = (n - r0) >> 2;
        for (int ti = warp; ti < T; ti += TC / 32) {
            for (int tk = lane; tk <= ti; tk += 32) {
                const int i = r0 + 4 * ti, kk = r0 + 4 * tk;
                float li[4][8], lk[4][8];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 p0 = *reinterpret_cast<const float4*>(As + (i + r) * n + jb);
                    const float4 q0 = *reinterpret_cast<const float4*>(As + (kk + r) * n + jb);
                    li[r][0] = p0.x; li[r][1] = p0.y; li[r][2] = p0.z; li[r][3] = p0.w;
                    lk[r][0] = q0.x; lk[r][1] = q0.y; lk[r][2] = q0.z; lk[r][3] = q0.w;
                    if (nb == 8) {
                        const float4 p1 = *reinterpret_cast<const float4*>(As + (i + r) * n + jb + 4);
                        const float4 q1 = *reinterpret_cast<const float4*>(As + (kk + r) * n + jb + 4);
                        li[r][4] = p1.x; li[r][5] = p1.y; li[r][6] = p1.z; li[r][7] = p1.w;
                        lk[r][4] = q1.x; lk[r][5] = q1.y; lk[r][6] = q1.z; lk[r][7] = q1.w;
                    } else {
#pragma unroll
                        for (int c = 4; c < 8; ++c) li[r][c] = lk[r][c] = 0.f;
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float4* pa = reinterpret_cast<float4*>(As + (i + r) * n + kk);
                    float4 av = *pa;
                    float o[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                        for (int c = 0; c < 8; ++c) o[cc] = fmaf(-li[r][c], lk[cc][c], o[cc]);
                    *pa = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        __syncthreads();
    }
    // outputs: row-major L (upper part zeroed) and the packed k-major factor
    if (a.L) {
        float* Lg = a.L + (long long)env * n * n;
        for (int idx = tid; idx < n * n; idx += TC) {
            int i = idx / n, j = idx - i * n;
            Lg[idx] = (j <= i) ? As[idx] : 0.f;
        }
    }
    if (a.Lt) {
        float* Ltg = a.Lt + (long long)env * a.lt_stride;
        for (int k = warp; k < n; k += TC / 32) {
            const int rs = k & ~7;
            const int off = lt_col_offset(k, n_pad);
            for (int r = rs + lane; r < n_pad; r += 32) Ltg[off + (r - rs)] = (r >= k && r < n) ? As[r * n + k] : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
static size_t tridiag_smem(int n) {
    return (size_t)tridiag_region_floats(n) * 4 + 4096 * 4 + 256 * 4 * 2 + 256 * 8 * 2 + 16 * 8 + 64 * 4 + 16;
}
static size_t chol_smem(int n) { return (size_t)n * n * 4 + 64 * 4; }

cudaError_t launch_sigma(const SigmaArgs& a, int n_env, cudaStream_t st) {
    if (a.n > kSigmaMaxN || (a.n & 3)) return cudaErrorInvalidValue;
    static size_t conf = 0;
    size_t smem = tridiag_smem(a.n);
    if (smem > conf) {
        cudaError_t e = cudaFuncSetAttribute(sigma_tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        conf = smem;
    }
    sigma_tridiag_kernel<<<n_env, TT, smem, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 g((a.n + kApplyWarps - 1) / kApplyWarps, n_env);
    applyq_kernel<true><<<g, kApplyWarps * 32, 0, st>>>(a.Vh, a.tau, a.F, a.Z, a.n);
    applyq_kernel<false><<<g, kApplyWarps * 32, 0, st>>>(a.Vh, a.tau, a.Z, a.cov, a.n);
    return cudaGetLastError();
}

cudaError_t launch_cholesky(const SigmaArgs& a, int n_env, cudaStream_t st) {
    if (a.n > kSigmaMaxN || (a.n & 3)) return cudaErrorInvalidValue;
    static size_t conf = 0;
    size_t smem = chol_smem(a.n);
    if (smem > conf) {
        cudaError_t e = cudaFuncSetAttribute(cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        conf = smem;
    }
    cholesky_kernel<<<n_env, TC, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace covo
