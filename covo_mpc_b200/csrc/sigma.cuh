// Internal declarations for the CoVO covariance kernels (sigma.cu).
#pragma once
#include "common.cuh"

namespace covo {

constexpr int kZoloPoles = 16;    // poles of the rational approximation of x^(-1/2)
constexpr int kDensePoles = 13;   // poles of the dense path (float64 inverses: 13 poles approximate x^(-1/2) to 2e-7 on the ladder's intervals)
constexpr int kZoloLadder = 10;   // spectral-ratio ladder: M/m = 4^(4+i), i = 0..9
constexpr int kSigmaMaxN = 224;   // n = 4H limit of the shared-memory resident kernels (H <= 56)
constexpr double kCovoOffset = 1e-2;  // "offset = -min_eign + 1e-2", controllers/covo.py:120-121

struct SigmaArgs {
    long long* prof = nullptr;  // optional: clock64() stamps at phase boundaries (debug)
    int n, n_pad;
    float sample_sigma;
    const float* R;      // [E][n][n]
    float* Vh;           // [E][n][n]  row k = Householder vector v_k (zeros for index <= k, v[k+1] = 1)
    float* tau;          // [E][n]
    float* Qt;           // [E][n][n]  Q^T = H_{n-3} ... H_1 H_0 (accumulated from the reflectors by qacc_kernel)
    float* F;            // [E][n][n]  exp(log_const/2) * (T - lam_min + offset)^(-1/2), full symmetric
    float* cov;          // [E][n][n]  Sigma = Q F Q^T, exactly symmetric          -> a_cov
    float* L;            // [E][n][n]  optional: lower Cholesky factor, row-major
    float* Lt;           // [E][lt_size]  packed k-major factor for the sampler
    double* diag;        // [E][4][n]: d, e, then (lam_min, gersh_lo, gersh_hi, logdet, ladder idx) for tests
    const double* zolo;  // [kZoloLadder][2][kZoloPoles]  shifts t_j, weights w_j (host-computed)
    int* status;         // [E]  0 ok, 1 spectral ratio beyond ladder, 2 Cholesky breakdown
    long long lt_stride;
    int e2_points = 16;     // trial shifts (= warps) per multisection round in E2 (measured: 16 beats 32 and 8)
    int cov_symmetric = 0;  // cov is exactly symmetric already (written by the sandwich kernel): Cholesky skips (C + C^T)/2
    // Cholesky -> rollout pipeline: when set, every finished 8-column block of the packed factor is written to Lt at once and
    // progress[env] is released to epoch + (blocks done), so that a concurrently running rollout kernel can start sampling
    int* progress = nullptr;
    int epoch = 0;
};

// host: Zolotarev/Hale-Higham-Trefethen nodes for x^(-1/2) on [m, M]
void zolotarev_nodes(double m, double M, int N, double* t, double* w);
void zolotarev_table(double* table /* [kZoloLadder][2][kZoloPoles] */);
void zolotarev_table_dense(double* table /* [kZoloLadder][2][kDensePoles] */);
constexpr size_t kZoloTableDoubles = (size_t)kZoloLadder * 2 * (kZoloPoles + kDensePoles);  // both ladders, E2's first
double zolotarev_ladder_M(int i);

cudaError_t launch_tridiag(const SigmaArgs& a, int n_env, cudaStream_t st);   // E1: R -> (d, e), reflectors
cudaError_t launch_qacc(const SigmaArgs& a, int n_env, cudaStream_t st);      // E1b: reflectors -> Q^T (independent of E2)
cudaError_t launch_trifunc(const SigmaArgs& a, int n_env, cudaStream_t st);   // E2: (d, e) -> F
cudaError_t launch_sandwich(const SigmaArgs& a, int n_env, cudaStream_t st);  // E3: Q F Q^T -> cov
cudaError_t launch_sigma(const SigmaArgs& a, int n_env, cudaStream_t st);     // E1 + E2 + E3: R -> cov
cudaError_t launch_cholesky(const SigmaArgs& a, int n_env, cudaStream_t st);  // cov -> L, Lt (and symmetrise cov)

// Tridiagonalisation-free optimize_sigma (sigma_dense.cu), the opt-in FAST path: lambda_min by adaptive Lanczos, A^(-1/2) as a
// 16-pole rational function with one float32 Gauss-Jordan inverse per pole, log det A from one more.  Writes a.cov (symmetric).
// Every pole is a cluster of CTAs (8 + 17 * 4 CTAs per matrix); less accurate than E1-E3 (see covo_b200.h: covo_get_sigma_path).
// scal: [n_env][4] doubles, Xbuf: [n_env][sigma_dense_scratch_floats(n)] floats.
size_t sigma_dense_scratch_floats(int n);
// ev_mid1 / ev_mid2 (optional): recorded after the Lanczos kernel and after the inverses (per-kernel timing).
cudaError_t launch_sigma_dense(const SigmaArgs& a, double* scal, float* Xbuf, int n_env, cudaStream_t st, cudaEvent_t ev_mid1 = nullptr,
                               cudaEvent_t ev_mid2 = nullptr);

}  // namespace covo
