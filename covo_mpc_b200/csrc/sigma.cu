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
__device__ int sturm_count(const double* d, const double* e2, int n, double x) {
    // e2[i] = e[i]^2 (coupling between i and i+1).  The loads run one iteration ahead of the fp64 chain.
    double pm = 1.0, p = d[0] - x;
    if (p == 0.0) p = -1e-300;
    int cnt = (__double2hiint(p) >> 31) & 1;
    double dn = d[1], en = e2[0];
    for (int i0 = 1; i0 < n; i0 += 8) {
        const int i1 = min(i0 + 8, n);
        for (int i = i0; i < i1; ++i) {
            const double di = dn, ei = en;
            if (i + 1 < n) {
                dn = d[i + 1];
                en = e2[i];
            }
            double pn = fma(di - x, p, -(ei * pm));
            // sign tests on the high word (integer pipe): only the recurrence itself runs on the fp64 pipe
            int hn = __double2hiint(pn);
            const int hp = __double2hiint(p);
            if (((hn & 0x7fffffff) | __double2loint(pn)) == 0) {
                pn = (hp < 0) ? 1e-300 : -1e-300;
                hn = __double2hiint(pn);
            }
            cnt += ((hn ^ hp) >> 31) & 1;
            pm = p;
            p = pn;
        }
        // the terms grow by at most ~(|d - x| + |e|) per step, so a range check every 8 steps is enough
        const int ex = (__double2hiint(p) >> 20) & 0x7ff;
        if (ex > 1023 + 400) {
            p *= 1e-120;
            pm *= 1e-120;
        } else if (ex < 1023 - 400) {
            p *= 1e120;
            pm *= 1e120;
        }
    }
    return cnt;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// E1 + E2
// ---------------------------------------------------------------------------------------------
// Row pitch of the shared-memory matrix: n + 4 or n + 8 so that pitch == 4 (mod 8): eight consecutive rows
// of one float4 column group then fall into eight different bank quadruples (conflict-free LDS.128).
__host__ __device__ inline int tridiag_pitch(int n) { return (n & 7) ? n + 8 : n + 4; }
__host__ __device__ inline int tridiag_region_floats(int n) {
    int a = n * tridiag_pitch(n);
    // E2 scratch: fp64 pivot numerators/denominators [2][2][17][n], then float gd[16][n], and the
    // (hi, lo, sign/zero) prefix arrays of the semiseparable generators [16][n+1] x 3
    int b = (2 * 2 * (kZoloPoles + 1) * n) * 2 + kZoloPoles * n + 3 * kZoloPoles * (n + 1);
    return ((a > b ? a : b) + 3) & ~3;
}

__global__ void __launch_bounds__(TT, 1) sigma_tridiag_kernel(const SigmaArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int n = a.n, tid = threadIdx.x, env = blockIdx.x, lane = tid & 31, warp = tid >> 5;
    const int LD = tridiag_pitch(n);
    float* As = reinterpret_cast<float*>(smraw);          // [n][LD]
    float* vs = As + tridiag_region_floats(n);            // [2048] v/q double buffers + w (see E1)
    float* ws = vs + 256;
    float* rw4 = vs + 2048;                               // [256] float4: per-row (-v, -w, vnext, 0) of the pass
    float* taus = vs + 3072;                              // [256] tau_k
    double* dd = reinterpret_cast<double*>(vs + 3328);    // [256]
    double* ee = dd + 256;                                // [256]
    double* e2s = ee + 256;                               // [256] e^2
    double* sc = e2s + 256;                               // [16] scalars
    double* rdbuf = sc + 16;                              // [64]
    float* red = reinterpret_cast<float*>(rdbuf + 64);    // [64]
    int* ired = reinterpret_cast<int*>(red + 64);         // [4]

    const float* Rg = a.R + (long long)env * n * n;
    float* Vg = a.Vh + (long long)env * n * n;
    float* taug = a.tau + (long long)env * n;

    // R <- (R + R^T)/2 (controllers/covo.py:117): coalesced load, then symmetrise in shared memory
    for (int i = warp; i < n; i += TT / 32)
        for (int j = lane; j < n; j += 32) As[i * LD + j] = Rg[i * n + j];
    for (int i = tid; i < 3328; i += TT) vs[i] = 0.f;
    __syncthreads();
    for (int i = warp; i < n; i += TT / 32)
        for (int j = lane; j < i; j += 32) {
            const float s2 = 0.5f * (As[i * LD + j] + As[j * LD + i]);
            As[i * LD + j] = s2;
            As[j * LD + i] = s2;
        }
    __syncthreads();
    COVO_STAMP(a, 8);

    // ---- E1: Householder tridiagonalisation (LAPACK ssytd2 recurrences), one pass over A per step -----
    // thread = (column group cgi of 4 columns, row class ch of 16): rows k+1+ch, k+1+ch+16, ...
    // State entering step k: A updated through step k-1; v_k (vs[cur], zero-extended, v[k+1] = 1), tau_k and
    // the raw matvec q_k = A v_k (qs[cur]).  Step k:
    //   A) s = q.v (block reduction)                      -> w = tau q - (tau^2 s / 2) v
    //   B) warp 0 forms the UPDATED row k+1 from (A, v, w), i.e. the next Householder vector v_{k+1}
    //      (look-ahead) while the other threads publish w;
    //   C) one pass: A <- A - v w^T - w v^T fused with q_{k+1} = A_new v_{k+1}.
    const int cgi = tid >> 4, ch = tid & 15;
    float* vbuf[2] = {vs, vs + 512};  // vs, ws are reused as [cur/next] v and q buffers (4 x 256 floats total)
    float* qbuf[2] = {ws, ws + 512};
    // layout: vs[0..255] v0 | ws[0..255] q0 | vs+512 v1 | ws+512 q1   (allocated below: 1024 floats)
    int cur = 0;
    // prologue: v_0, tau_0 from row 0; q_0 = A v_0
    if (warp == 0) {
        const float* x = As;
        float sig = 0.f;
        for (int j = 2 + lane; j < n; j += 32) sig = fmaf(x[j], x[j], sig);
        sig = wsum(sig);
        const float x0 = x[1];
        float beta, tau, scale;
        if (sig == 0.f) {
            beta = x0; tau = 0.f; scale = 0.f;
        } else {
            beta = -copysignf(sqrtf(fmaf(x0, x0, sig)), x0);
            tau = (beta - x0) / beta;
            scale = 1.0f / (x0 - beta);
        }
        for (int j = lane; j < 256; j += 32) {
            float v = 0.f;
            if (j == 1) v = 1.f;
            else if (j >= 2 && j < n && tau != 0.f) v = x[j] * scale;
            vbuf[0][j] = v;
        }
        if (lane == 0) {
            red[0] = beta;
            red[1] = tau;
        }
    }
    __syncthreads();
    {
        const int c0 = 0, col = c0 + 4 * cgi;
        const bool active = cgi < (n >> 2);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            for (int i = 1 + ch; i < n; i += 16) {
                const float4 av = *reinterpret_cast<const float4*>(As + i * LD + col);
                const float vi = vbuf[0][i];
                acc.x = fmaf(av.x, vi, acc.x);
                acc.y = fmaf(av.y, vi, acc.y);
                acc.z = fmaf(av.z, vi, acc.z);
                acc.w = fmaf(av.w, vi, acc.w);
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        float part0 = 0.f;
        if (active && ch == 0) {
            *reinterpret_cast<float4*>(qbuf[0] + col) = acc;
            const float4 vv = *reinterpret_cast<const float4*>(vbuf[0] + col);
            part0 = acc.x * vv.x + acc.y * vv.y + acc.z * vv.z + acc.w * vv.w;  // v is zero for j <= 0
        }
        part0 += __shfl_xor_sync(0xffffffffu, part0, 16);
        if (lane == 0) red[8 + warp] = part0;
    }
    __syncthreads();
    long long pa[6] = {0, 0, 0, 0, 0, 0}, pt0 = 0;
#define PH(i) do { if (a.prof && tid == 0) { long long t_ = clock64(); pa[i] += t_ - pt0; pt0 = t_; } } while (0)
    if (a.prof && tid == 0) pt0 = clock64();
    for (int k = 0; k < n - 2; ++k) {
        const float* v = vbuf[cur];
        const float* q = qbuf[cur];
        float* vn = vbuf[cur ^ 1];
        float* qn = qbuf[cur ^ 1];
        const float beta = red[0], tau = red[1];
        // A) bookkeeping of step k.  No global stores inside the loop: a barrier after a global store waits for the
        //    L2 round trip.  v_k is parked LAPACK-style in the dead part of row k of A (entries j >= k+2; v[k+1] = 1
        //    is implicit) and everything is written to HBM once after the loop.
        if (tid >= 512 && tid < 512 + 256) {
            const int j = tid - 512;
            if (j >= k + 2 && j < n) As[k * LD + j] = v[j];
        }
        if (tid == 800) {
            dd[k] = (double)As[k * LD + k];
            ee[k] = (double)beta;
            taus[k] = tau;
        }
        // B) look-ahead Householder vector of step k+1 from the UPDATED row k+1, all threads:
        //      s = q.v (partials left in red[8..39] by the previous pass), c2 = tau^2 s / 2,
        //      w_j = tau q_j - c2 v_j,   updated row  r_j = A[k+1][j] - w_j - w_{k+1} v_j   (v_{k+1} = 1)
        //    B1: thread j forms w_j, r_j and its share of sigma = sum_{j >= k+3} r_j^2  -> barrier
        //    B2: every warp finishes sigma, the scalar chain (beta', tau', scale'), thread j publishes v'_j
        // Only warps 0..7 (one thread per entry of the row) run B; the other 24 warps go straight to the barriers,
        // so the eight working warps see near single-warp latencies on their shuffle / MUFU chains.
        float wj = 0.f, rj = 0.f;
        const int j = tid;
        if (warp < 8) {
            const float sdot = wsum(red[8 + lane]);
            const float c2 = 0.5f * tau * tau * sdot;
            const int r1 = k + 1;
            const float wr1 = fmaf(tau, q[r1], -c2 * v[r1]);
            const float vj = v[j];
            float sq = 0.f;
            if (j > k && j < n) wj = fmaf(tau, q[j], -c2 * vj);
            if (j >= k + 2 && j < n) rj = As[r1 * LD + j] - wj - wr1 * vj;
            if (j >= k + 3) sq = rj * rj;
            ws[1024 + j] = wj;
            if (j == k + 2) red[4] = rj;  // x0
            sq = wsum(sq);
            if (lane == 0) red[40 + warp] = sq;
        }
        __syncthreads();
        if (warp < 8) {
            float sig = red[40 + (lane & 7)];
            sig += __shfl_xor_sync(0xffffffffu, sig, 4);
            sig += __shfl_xor_sync(0xffffffffu, sig, 2);
            sig += __shfl_xor_sync(0xffffffffu, sig, 1);
            const float x0 = red[4];
            float nbeta, ntau, nscale;
            if (sig == 0.f) {
                nbeta = x0; ntau = 0.f; nscale = 0.f;
            } else {
                // MUFU rsq / rcp + one Newton step (~1 ulp), instead of the IEEE sqrt / div sequences
                const float nn2 = fmaf(x0, x0, sig);
                float rs = rsqrtf(nn2);
                rs = rs * (1.5f - 0.5f * nn2 * rs * rs);
                nbeta = -copysignf(nn2 * rs, x0);
                float rb = __frcp_rn(nbeta);
                ntau = (nbeta - x0) * rb;
                nscale = __frcp_rn(x0 - nbeta);
            }
            float vv = 0.f;
            if (j == k + 2 && j < n) vv = 1.f;
            else if (j >= k + 3 && j < n && ntau != 0.f) vv = rj * nscale;
            vn[j] = vv;
            reinterpret_cast<float4*>(rw4)[j] = make_float4(-v[j], -wj, vv, 0.f);
            if (tid == 0) {
                red[2] = nbeta;
                red[3] = ntau;
            }
        }
        __syncthreads();
        PH(1);
        // C) fused pass
        {
            const float* wv = ws + 1024;
            const int c0 = (k + 1) & ~3;
            const int ncg = (n - c0) >> 2;
            // thread map: 64 column groups x 16 row classes while more than 32 column groups are alive, then
            // 32 x 32 (halves the rows per thread once half of the column-group slots would idle)
            const bool wide = false;  // measured: the 32 x 32 map is slower (fixed per-step costs dominate the tail)
            const int cg2 = wide ? (tid >> 5) : cgi;
            const int ch2 = wide ? lane : ch;
            const int rstep = wide ? 32 : 16;
            const bool active = cg2 < ncg;
            const int col = c0 + 4 * cg2;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) {
                const float4 wj4 = *reinterpret_cast<const float4*>(wv + col);
                const float4 vj4 = *reinterpret_cast<const float4*>(v + col);
                const float4* rowp = reinterpret_cast<const float4*>(rw4);
#pragma unroll 4
                for (int i = k + 1 + ch2; i < n; i += rstep) {
                    const float4 r4 = rowp[i];  // (-v_i, -w_i, vnext_i, 0): one LDS.128 per row
                    float4* p = reinterpret_cast<float4*>(As + i * LD + col);
                    float4 av = *p;
                    av.x = fmaf(r4.x, wj4.x, fmaf(r4.y, vj4.x, av.x));
                    av.y = fmaf(r4.x, wj4.y, fmaf(r4.y, vj4.y, av.y));
                    av.z = fmaf(r4.x, wj4.z, fmaf(r4.y, vj4.z, av.z));
                    av.w = fmaf(r4.x, wj4.w, fmaf(r4.y, vj4.w, av.w));
                    *p = av;
                    acc.x = fmaf(av.x, r4.z, acc.x);
                    acc.y = fmaf(av.y, r4.z, acc.y);
                    acc.z = fmaf(av.z, r4.z, acc.z);
                    acc.w = fmaf(av.w, r4.z, acc.w);
                }
            }
            if (wide) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
                acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16);
                acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
                acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
            }
            // leaders publish q_{k+1} and their share of s_{k+1} = q_{k+1} . v_{k+1}
            float partn = 0.f;
            if (active && ch2 == 0) {
                *reinterpret_cast<float4*>(qn + col) = acc;
                const float4 vv = *reinterpret_cast<const float4*>(vn + col);  // zero for j <= k+1
                partn = acc.x * vv.x + acc.y * vv.y + acc.z * vv.z + acc.w * vv.w;
            }
            if (!wide) partn += __shfl_xor_sync(0xffffffffu, partn, 16);
            else partn = __shfl_sync(0xffffffffu, partn, 0);
            // red[8..39] of this step were consumed before sync B, so they can be overwritten here
            if (lane == 0) red[8 + warp] = partn;
            if (tid == 0) {
                red[0] = red[2];
                red[1] = red[3];
            }
        }
        __syncthreads();
        PH(2);
        cur ^= 1;
    }
    if (a.prof && tid == 0) for (int i = 0; i < 5; ++i) a.prof[40 + i] = pa[i];
    if (tid == 0) {
        dd[n - 2] = (double)As[(n - 2) * LD + (n - 2)];
        dd[n - 1] = (double)As[(n - 1) * LD + (n - 1)];
        ee[n - 2] = (double)As[(n - 1) * LD + (n - 2)];
        ee[n - 1] = 0.0;
        taus[n - 2] = 0.f;
        taus[n - 1] = 0.f;
    }
    __syncthreads();
    // reflectors -> HBM, row k = v_k zero-extended (coalesced), and tau
    for (int k2 = warp; k2 < n; k2 += TT / 32) {
        const bool live = (k2 < n - 2) && (taus[k2] != 0.f);
        for (int j = lane; j < n; j += 32) {
            float vv = 0.f;
            if (k2 < n - 2) {
                if (j == k2 + 1) vv = 1.f;
                else if (j >= k2 + 2 && live) vv = As[k2 * LD + j];
            }
            Vg[(long long)k2 * n + j] = vv;
        }
    }
    if (tid < n) taug[tid] = taus[tid];
    __syncthreads();
    // Compact-WY factors for apply-Q: block m holds reflectors k = k_hi-7 .. k_hi, k_hi = n-3-8m (ascending local
    // index r <-> k = k_hi-7+r; missing ones have tau = 0).  H_{k_lo} ... H_{k_hi} = I - V T V^T, T upper triangular
    // (LAPACK slarft, forward / columnwise): T[i][i] = tau_i, T[0:i, i] = -tau_i T[0:i,0:i] (V[:,0:i]^T v_i).
    {
        const int nblk = (n - 2 + kWyBlock - 1) / kWyBlock;
        float* Twg = a.Tw + (long long)env * (n / kWyBlock + 1) * 64;
        float* gsm = reinterpret_cast<float*>(rdbuf);  // per-warp scratch is not needed: one warp per block, Gram in regs
        (void)gsm;
        for (int m = warp; m < nblk; m += TT / 32) {
            const int khi = n - 3 - kWyBlock * m;
            // Gram entries g[r][c] = v_r . v_c for r < c, all lanes end with the sums
            float g[8][8];
            float vr[8][7];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int k = khi - 7 + r;
#pragma unroll
                for (int li = 0; li < 7; ++li) {
                    const int i = lane + 32 * li;
                    float vv = 0.f;
                    if (k >= 0 && i < n) {
                        if (i == k + 1) vv = 1.f;
                        else if (i >= k + 2 && taus[k] != 0.f) vv = As[k * LD + i];
                    }
                    vr[r][li] = vv;
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = r + 1; c < 8; ++c) {
                    float d = 0.f;
#pragma unroll
                    for (int li = 0; li < 7; ++li) d = fmaf(vr[r][li], vr[c][li], d);
                    g[r][c] = d;
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = r + 1; c < 8; ++c) g[r][c] += __shfl_xor_sync(0xffffffffu, g[r][c], o);
            float tauk[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int k = khi - 7 + r;
                tauk[r] = (k >= 0) ? taus[k] : 0.f;
            }
            float T[8][8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) T[r][c] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                T[i][i] = tauk[i];
#pragma unroll
                for (int r = 0; r < i; ++r) {
                    float acc = 0.f;
#pragma unroll
                    for (int c = r; c < i; ++c) acc = fmaf(T[r][c], g[c][i], acc);
                    T[r][i] = -tauk[i] * acc;
                }
            }
            if (lane < 8) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (lane == r) Twg[m * 64 + r * 8 + c] = T[r][c];
            }
        }
    }
    __syncthreads();
    COVO_STAMP(a, 9);

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
    // lam_min by multisection.  fp64 issues at ~1/9 of the fp32 rate on this part and the Sturm recurrence is
    // a serial chain, so the cost is (evaluation points) x (rounds): 128 points (one warp per scheduler) x 6
    // rounds shrink the Gershgorin bracket by 129^6 ~ 4.6e12, i.e. to ~1e-11 absolute.
    constexpr int MS = 128;
    for (int round = 0; round < 6; ++round) {
        const double lo = sc[0], hi = sc[1];
        if (tid == 0) ired[0] = MS;
        __syncthreads();
        // spread the four evaluation warps over the four schedulers: warps 0..3
        if (tid < MS) {
            const double x = lo + (hi - lo) * ((double)(tid + 1) / (double)(MS + 1));
            if (sturm_count(dd, e2s, n, x) >= 1) atomicMin(ired, tid);
        }
        __syncthreads();
        if (tid == 0) {
            const int ts = ired[0];
            const double step = (hi - lo) / (double)(MS + 1);
            sc[0] = (ts == 0) ? lo : lo + step * ts;
            sc[1] = (ts == MS) ? hi : lo + step * (ts + 1);
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
    if (tid < NPOLE + 1 || (tid >= 32 && tid < 32 + NPOLE + 1)) {
        const bool fwd = tid < 32;
        const int q = fwd ? tid : tid - 32;
        const double tq = (q < NPOLE) ? zt[q] : 0.0;
        double* nm = num + ((fwd ? 0 : 1) * (NPOLE + 1) + q) * n;
        double* dn = den + ((fwd ? 0 : 1) * (NPOLE + 1) + q) * n;
        double pm = 0.0, p = 1.0;
        const double sh = shift0 + tq;
        double an = dd[fwd ? 0 : n - 1] + sh, bn = 0.0;
        for (int s = 0; s < n; ++s) {
            const int i = fwd ? s : n - 1 - s;
            const double ai = an, b2 = bn;
            if (s + 1 < n) {  // operands of the next step, off the dependent chain
                const int i2 = fwd ? s + 1 : n - 2 - s;
                an = dd[i2] + sh;
                bn = e2s[fwd ? i2 - 1 : i2];
            }
            double pn = fma(ai, p, -(b2 * pm));
            nm[i] = pn;
            dn[i] = p;
            // range control on the exponent field (integer test, keeps the fp64 pipe for the recurrence)
            const int ex = (__double2hiint(pn) >> 20) & 0x7ff;
            if (ex > 1023 + 400) {
                pn *= 1e-120;
                p *= 1e-120;
            } else if (ex < 1023 - 400) {
                pn *= 1e120;
                p *= 1e120;
            }
            pm = p;
            p = pn;
        }
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
        for (int i = warp; i < n; i += TT / 32) {
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
                if (l + 32 < n) {
#pragma unroll
                    for (int q = 0; q < NPOLE; ++q) val[q] *= p32[q * n + l];
                }
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
    __syncthreads();
    COVO_STAMP(a, 14);
}

// ---------------------------------------------------------------------------------------------
// E3: out(:, c) = Q in(c, :)^T;  Q = H_0 H_1 ... H_{n-3}.  One warp owns two vectors (their elements spread
// over the lanes, 7 registers each); the reflectors stream through shared memory in chunks of 32 rows with a
// two-stage cp.async pipeline, so the serial chain per reflector is dot -> 5 shuffles -> axpy and nothing waits
// on global memory.
// ---------------------------------------------------------------------------------------------
constexpr int kApplyWarps = 4;
constexpr int kApplyCols = 1;                 // vectors per warp (the instruction stream per warp is the critical path)
constexpr int kApplyChunk = 32;               // reflectors per pipeline stage
constexpr int kMaxLi = kSigmaMaxN / 32;       // 7

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <bool STORE_STRIDED, bool UPPER_ONLY>
__global__ void __launch_bounds__(kApplyWarps * 32) applyq_kernel(const float* __restrict__ Vh,
                                                                 const float* __restrict__ Tw,
                                                                 const float* __restrict__ in, float* __restrict__ out,
                                                                 int n) {
    extern __shared__ __align__(16) float sv[];  // [2][kApplyChunk][n] reflectors, then [2][4][64] T factors
    const int env = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = (blockIdx.x * kApplyWarps + warp) * kApplyCols;
    Vh += (long long)env * n * n;
    Tw += (long long)env * (n / kWyBlock + 1) * 64;
    in += (long long)env * n * n;
    out += (long long)env * n * n;
    float* sT = sv + 2 * kApplyChunk * n;
    float x[kApplyCols][kMaxLi];
#pragma unroll
    for (int j = 0; j < kApplyCols; ++j)
#pragma unroll
        for (int li = 0; li < kMaxLi; ++li) {
            int i = lane + 32 * li;
            float xv = 0.f;
            if (i < n && c0 + j < n) {
                const int cc = c0 + j;
                // `in` is symmetric; when only its upper triangle is stored (F), read (min, max)
                xv = UPPER_ONLY ? in[(long long)min(cc, i) * n + max(cc, i)] : in[(long long)cc * n + i];
            }
            x[j][li] = xv;
        }
    const int nref = n - 2;                                   // reflectors k = 0 .. n-3, applied in descending k
    const int nchunks = (nref + kApplyChunk - 1) / kApplyChunk;
    const int vec_per_row = n >> 2;
    auto prefetch = [&](int j) {
        // chunk j holds k = kh, kh-1, ..., kl  (kh = n-3 - j*chunk); slot s <-> k = kh - s.  Its 4 WY blocks are
        // the blocks m = 4j .. 4j+3 of the T table.
        const int kh = n - 3 - j * kApplyChunk;
        const int cnt = min(kApplyChunk, kh + 1);
        float* dst = sv + (j & 1) * kApplyChunk * n;
        for (int idx = tid; idx < cnt * vec_per_row; idx += kApplyWarps * 32) {
            int s = idx / vec_per_row, v4 = idx - s * vec_per_row;
            cp_async16(dst + s * n + 4 * v4, Vh + (long long)(kh - s) * n + 4 * v4);
        }
        const int nblk = (cnt + kWyBlock - 1) / kWyBlock;
        for (int idx = tid; idx < nblk * 16; idx += kApplyWarps * 32)
            cp_async16(sT + (j & 1) * 256 + idx * 4, Tw + (long long)(4 * j) * 64 + idx * 4);
        cp_async_commit();
    };
    prefetch(0);
    for (int j = 0; j < nchunks; ++j) {
        if (j + 1 < nchunks) {
            prefetch(j + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int kh = n - 3 - j * kApplyChunk;
        const int cnt = min(kApplyChunk, kh + 1);
        const float* buf = sv + (j & 1) * kApplyChunk * n;
        const int nblk = (cnt + kWyBlock - 1) / kWyBlock;
        for (int q = 0; q < nblk; ++q) {
            // block q: slots 8q .. 8q+7 (descending k); ascending local index r <-> slot 8q + 7 - r
            const float* Tb = sT + (j & 1) * 256 + q * 64;
            // y = V^T x  (slots past the end of the chunk hold stale data but their T rows/cols are zero: tau = 0)
            float y[kApplyCols][8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int s = 8 * q + 7 - r;
                float v[kMaxLi];
#pragma unroll
                for (int li = 0; li < kMaxLi; ++li) {
                    int i = lane + 32 * li;
                    v[li] = (i < n && s < cnt) ? buf[s * n + i] : 0.f;
                }
#pragma unroll
                for (int jj = 0; jj < kApplyCols; ++jj) {
                    float d = 0.f;
#pragma unroll
                    for (int li = 0; li < kMaxLi; ++li) d = fmaf(v[li], x[jj][li], d);
                    y[jj][r] = d;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int jj = 0; jj < kApplyCols; ++jj)
#pragma unroll
                    for (int r = 0; r < 8; ++r) y[jj][r] += __shfl_xor_sync(0xffffffffu, y[jj][r], o);
            // z = T y (upper triangular), then x -= V z
            float z[kApplyCols][8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int jj = 0; jj < kApplyCols; ++jj) z[jj][r] = 0.f;
#pragma unroll
                for (int c = r; c < 8; ++c) {
                    const float t = Tb[r * 8 + c];
#pragma unroll
                    for (int jj = 0; jj < kApplyCols; ++jj) z[jj][r] = fmaf(t, y[jj][c], z[jj][r]);
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int s = 8 * q + 7 - r;
#pragma unroll
                for (int li = 0; li < kMaxLi; ++li) {
                    int i = lane + 32 * li;
                    const float v = (i < n && s < cnt) ? buf[s * n + i] : 0.f;
#pragma unroll
                    for (int jj = 0; jj < kApplyCols; ++jj) x[jj][li] = fmaf(-z[jj][r], v, x[jj][li]);
                }
            }
        }
        __syncthreads();  // the buffer is recycled two chunks later
    }
#pragma unroll
    for (int jj = 0; jj < kApplyCols; ++jj)
#pragma unroll
        for (int li = 0; li < kMaxLi; ++li) {
            int i = lane + 32 * li;
            if (i < n && c0 + jj < n) {
                if (STORE_STRIDED) out[(long long)i * n + (c0 + jj)] = x[jj][li];
                else out[(long long)(c0 + jj) * n + i] = x[jj][li];
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
    float* Lp = As + n * n;                       // [8][n_pad]  panel, transposed
    float* L11 = Lp + 8 * n_pad;                  // [8][8]
    float* covg = a.cov + (long long)env * n * n;
    COVO_STAMP(a, 23);
    // (a_cov + a_cov.T)/2, controllers/covo.py:132: coalesced load, symmetrise in shared memory, write back
    for (int i = warp; i < n; i += TC / 32)
        for (int j = lane; j < n; j += 32) As[i * n + j] = covg[i * n + j];
    __syncthreads();
    for (int i = warp; i < n; i += TC / 32)
        for (int j = lane; j < i; j += 32) {
            const float s2 = 0.5f * (As[i * n + j] + As[j * n + i]);
            As[i * n + j] = s2;
            As[j * n + i] = s2;
        }
    __syncthreads();
    for (int i = warp; i < n; i += TC / 32)
        for (int j = lane; j < n; j += 32) covg[i * n + j] = As[i * n + j];
    COVO_STAMP(a, 24);
    for (int jb = 0; jb < n; jb += 8) {
        const int nb = min(8, n - jb);
        if (jb == 8) COVO_STAMP(a, 26);
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
                    float rinv = rsqrtf(dcc);
                    rinv = rinv * (1.5f - 0.5f * dcc * rinv * rinv);
                    if (lane == c) r8[c] = dcc * rinv;  // sqrt
                    else if (lane > c) r8[c] = r8[c] * rinv;
                }
#pragma unroll
                for (int c2 = c + 1; c2 < 8; ++c2) {
                    float l2 = __shfl_sync(0xffffffffu, r8[c], c2);
                    if (c2 < nb && lane >= c2) r8[c2] = fmaf(-r8[c], l2, r8[c2]);
                }
            }
            if (lane < 8) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float v = (lane < nb && c < nb && c <= lane) ? r8[c] : 0.f;
                    if (lane < nb && c < nb) As[(jb + lane) * n + jb + c] = v;
                    // L11 holds the factor with the RECIPROCAL diagonal (the panel solve multiplies)
                    L11[lane * 8 + c] = (c == lane) ? ((lane < nb) ? 1.0f / r8[c] : 1.0f) : v;
                }
            }
        }
        __syncthreads();
        // (b) panel: L21 = A21 L11^-T
        for (int i = jb + nb + tid; i < n; i += TC) {
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
                float s = x[c];
#pragma unroll
                for (int c2 = 0; c2 < c; ++c2) s = fmaf(-x[c2], L11[c * 8 + c2], s);
                x[c] = s * L11[c * 8 + c];
            }
            *reinterpret_cast<float4*>(As + i * n + jb) = make_float4(x[0], x[1], x[2], x[3]);
            if (nb == 8) *reinterpret_cast<float4*>(As + i * n + jb + 4) = make_float4(x[4], x[5], x[6], x[7]);
#pragma unroll
            for (int c = 0; c < 8; ++c) Lp[c * n_pad + i] = x[c];
        }
        __syncthreads();
        // (c) trailing update on the lower-triangle 4x4 tiles
        const int r0 = jb + nb;
        const int T = (n - r0) >> 2;
        const int ntiles = T * (T + 1) / 2;
        for (int q = tid; q < ntiles; q += TC) {
            int ti = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
            while (ti * (ti + 1) / 2 > q) --ti;
            while ((ti + 1) * (ti + 2) / 2 <= q) ++ti;
            const int tk = q - ti * (ti + 1) / 2;
            const int i = r0 + 4 * ti, kk = r0 + 4 * tk;
            float o[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 av = *reinterpret_cast<const float4*>(As + (i + r) * n + kk);
                o[r][0] = av.x; o[r][1] = av.y; o[r][2] = av.z; o[r][3] = av.w;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 li = *reinterpret_cast<const float4*>(Lp + c * n_pad + i);
                const float4 lk = *reinterpret_cast<const float4*>(Lp + c * n_pad + kk);
                const float lir[4] = {li.x, li.y, li.z, li.w};
                const float lkr[4] = {lk.x, lk.y, lk.z, lk.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) o[r][cc] = fmaf(-lir[r], lkr[cc], o[r][cc]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                *reinterpret_cast<float4*>(As + (i + r) * n + kk) = make_float4(o[r][0], o[r][1], o[r][2], o[r][3]);
        }
        __syncthreads();
    }
    COVO_STAMP(a, 25);
    // outputs: row-major L (upper part zeroed) and the packed k-major factor
    if (a.L) {
        float* Lg = a.L + (long long)env * n * n;
        for (int i = warp; i < n; i += TC / 32)
            for (int j = lane; j < n; j += 32) Lg[i * n + j] = (j <= i) ? As[i * n + j] : 0.f;
    }
    if (a.Lt) {
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
static size_t tridiag_smem(int n) {
    return (size_t)tridiag_region_floats(n) * 4 + 3328 * 4 + 256 * 8 * 3 + 16 * 8 + 64 * 8 + 64 * 4 + 16;
}
static size_t chol_smem(int n) { return (size_t)n * n * 4 + (size_t)8 * round_up8(n) * 4 + 64 * 4; }

cudaError_t launch_sigma(const SigmaArgs& a, int n_env, cudaStream_t st) {
    if (a.n > kSigmaMaxN || (a.n & 3)) return cudaErrorInvalidValue;
    static size_t conf[32] = {};
    size_t smem = tridiag_smem(a.n);
    cudaError_t e0 = ensure_smem_attr(sigma_tridiag_kernel, smem, conf);
    if (e0 != cudaSuccess) return e0;
    sigma_tridiag_kernel<<<n_env, TT, smem, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int cols_per_cta = kApplyWarps * kApplyCols;
    dim3 g((a.n + cols_per_cta - 1) / cols_per_cta, n_env);
    const size_t asm_bytes = (size_t)(2 * kApplyChunk * a.n + 2 * 256) * sizeof(float);
    static size_t conf_q1[32] = {}, conf_q2[32] = {};
    e = ensure_smem_attr(applyq_kernel<true, true>, asm_bytes, conf_q1);
    if (e == cudaSuccess) e = ensure_smem_attr(applyq_kernel<false, false>, asm_bytes, conf_q2);
    if (e != cudaSuccess) return e;
    applyq_kernel<true, true><<<g, kApplyWarps * 32, asm_bytes, st>>>(a.Vh, a.Tw, a.F, a.Z, a.n);
    applyq_kernel<false, false><<<g, kApplyWarps * 32, asm_bytes, st>>>(a.Vh, a.Tw, a.Z, a.cov, a.n);
    return cudaGetLastError();
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
