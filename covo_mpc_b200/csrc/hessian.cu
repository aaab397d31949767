// K3 + K4: exact control-space Hessian of the nominal-rollout cost.
//
// Replaces controllers/covo.py:134-185 (get_hessian = jacfwd(jacfwd(get_cumulated_cost)) over the
// Python-unrolled H-step rollout with deterministic=True, no termination freeze, no discount).
//
// The reference pushes (4H)^2 second-order tangent lanes through all H steps.  Here the same exact
// Hessian is assembled from per-transition derivatives (verified against the forward-over-forward
// oracle to round-off, tests/test_hessian_gpu.py):
//   z_t = (x_t, u_t),  x_{t+1} = F(z_t),  c_t(x_t) = -r(x_t),  A_t = dF/dx, B_t = dF/du
//   adjoint      lam_H = 0,  lam_t = grad c_t + A_t^T lam_{t+1}
//   Lagrangian   W_t = blkdiag(hess c_t, 0) + sum_k lam_{t+1,k} hess F_k(z_t)        (17 x 17)
//   backward     P_H = 0,  X = P_{t+1} [A_t B_t],
//                S_t = W_t^{xu} + A_t^T X_B,  D_t = W_t^{uu} + B_t^T X_B,  P_t = W_t^{xx} + A_t^T X_A
//   forward      Phi = B_I;  for J > I:  R[I,J] = Phi^T S_J,  Phi <- A_J Phi;   R[I,I] = D_I
// Kernel 1 (grid = H x envs): CTA t re-runs the nominal rollout to x_t, then 153 threads evaluate the
//   transition and the cost once each in hyper-dual arithmetic (quad_model.cuh) -- one (a <= b) pair of
//   the 17 local inputs per thread -- giving A_t, B_t, grad/hess c_t and all 13 hess F_k.
// Kernel 2 (grid = envs): adjoint, contraction and the 13x13 backward recursion.
// Kernel 3 (grid = 4H/2 x envs): the 4H forward chains, one half-warp each, spread over the SMs.
// Sub-gradient conventions (clip ties, |.|, sqrt at 0) are those of quad_model.cuh / the oracle.
#include <cuda_runtime.h>

#include "common.cuh"
#include "hessian.cuh"
#include "hessian_local.cuh"

namespace covo {

size_t hessian_workspace_floats(int H) { return (size_t)H * (14 * NPAIR + 14 * NZ); }

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160) hess_local_kernel(const HessianArgs a) {
    const int t = blockIdx.x, env = blockIdx.y, tid = threadIdx.x;
    const int H = a.H;
    __shared__ float sx[16], su[4], sfd[4], spt[4], svt[4];
    if (tid == 0) {
        if (t == 0 && a.status) a.status[env] = 0;  // first kernel of the covariance step: the status is per step, not sticky
        if (t == 0 && a.progress) a.progress[env] = 0;  // the previous step's rollout (its only reader) is complete: stream order
        const float* st_g = a.state24 + (long long)env * kStateFloats;
        QState<float> s;
        float fd[3], pt[3], vt[3];
        load_state24(st_g, s, fd, pt, vt);
        const int t0 = a.time[env];
        const float* mu = a.a_mean + (long long)env * 4 * H;
        for (int h = 0; h < t; ++h) {
            int hs = a.shift ? min(h + 1, H - 1) : h;
            float u[4] = {mu[hs * 4 + 0], mu[hs * 4 + 1], mu[hs * 4 + 2], mu[hs * 4 + 3]};
            quad_step(s, u, fd, a.env);
            // deterministic=True: the disturbance is zero after the first step (envs/quadrotor.py:234-235)
            fd[0] = fd[1] = fd[2] = 0.f;
        }
        if (t > 0) {
            int row = min(t0 + t, a.traj_len - 1);  // clamped gather, dynamics/free.py:153-155
            for (int k = 0; k < 3; ++k) {
                pt[k] = a.pos_traj[((long long)env * a.traj_stride + (long long)row * 3) + k];
                vt[k] = a.vel_traj[((long long)env * a.traj_stride + (long long)row * 3) + k];
            }
        }
        for (int k = 0; k < 3; ++k) sx[k] = s.p[k];
        for (int k = 0; k < 4; ++k) sx[3 + k] = s.q[k];
        for (int k = 0; k < 3; ++k) sx[7 + k] = s.v[k];
        for (int k = 0; k < 3; ++k) sx[10 + k] = s.w[k];
        int hs = a.shift ? min(t + 1, H - 1) : t;
        for (int k = 0; k < 4; ++k) su[k] = mu[hs * 4 + k];
        for (int k = 0; k < 3; ++k) { sfd[k] = fd[k]; spt[k] = pt[k]; svt[k] = vt[k]; }
    }
    __syncthreads();
    if (a.prof && tid == 0 && t == H - 1 && env == 0) a.prof[6] = clock64();
    if (tid >= NPAIR) return;
    int pa, pb;
    pair_from_index(tid, pa, pb);
    float x[13], u[4], fd[3], pt[3], vt[3], Fab[14], Fa[14];
    for (int k = 0; k < 13; ++k) x[k] = sx[k];
    for (int k = 0; k < 4; ++k) u[k] = su[k];
    for (int k = 0; k < 3; ++k) { fd[k] = sfd[k]; pt[k] = spt[k]; vt[k] = svt[k]; }
    hess_local_task(x, u, fd, pt, vt, a.env, pa, pb, Fab, Fa);
    float* ws = a.workspace + ((long long)env * H + t) * (14 * NPAIR + 14 * NZ);
    for (int k = 0; k < 14; ++k) ws[k * NPAIR + tid] = Fab[k];
    if (pa == pb) {
        float* g = ws + 14 * NPAIR;
        for (int k = 0; k < 14; ++k) g[k * NZ + pa] = Fa[k];
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int kAsmThreads = 1024;
constexpr int NZP = 20;  // row pitch of [A_t | B_t] in shared memory: 13 + 4 entries padded to 5 float4
constexpr int kFwRec = NX * NZP + NX * 4 + 16;  // 328 floats per step handed to the forward-chain kernel (float4 multiple)

__global__ void __launch_bounds__(kAsmThreads) hess_assemble_kernel(const HessianArgs a) {
    extern __shared__ __align__(16) float smf[];
    const int env = blockIdx.x, tid = threadIdx.x;
    const int H = a.H;
    // shared layout (everything read as float4 first)
    float* S = smf;                    // [H][13][4]
    float* D = S + H * NX * 4;         // [H][4][4]
    float* G = D + H * 16;             // [H][13][NZP]  (A_t | B_t | pad)
    float* cg = G + H * NX * NZP;      // [H][13]
    float* lam = cg + H * NX;          // [H+1][13]
    float* W = lam + (H + 1) * NX;     // [H][153]
    float* P = W + H * NPAIR;          // [13][13]
    float* X = P + NX * NX;            // [13][NZP]
    const float* wsb = a.workspace + (long long)env * H * (14 * NPAIR + 14 * NZ);
    const int rec = 14 * NPAIR + 14 * NZ;
    COVO_STAMP(a, 0);

    // [A_t | B_t | grad c_t] rows of the workspace: 14 x 17 contiguous floats per t, several loads in flight
    {
        const int per_t = 14 * NZ;  // 238
        const int total = H * per_t;
#pragma unroll 4
        for (int i = tid; i < total; i += kAsmThreads) {
            const int t = i / per_t, rr = i - t * per_t;
            const int r = rr / NZ, c = rr - r * NZ;
            const float val = __ldg(wsb + (long long)t * rec + 14 * NPAIR + rr);
            if (r < NX) G[(t * NX + r) * NZP + c] = val;
            else if (c < NX) cg[t * NX + c] = (t >= 1) ? val : 0.f;  // c_0 is constant in U
        }
        for (int i = tid; i < H * NX; i += kAsmThreads) {  // zero the three pad columns
            G[i * NZP + 17] = 0.f;
            G[i * NZP + 18] = 0.f;
            G[i * NZP + 19] = 0.f;
        }
    }
    for (int i = tid; i < NX; i += blockDim.x) lam[H * NX + i] = 0.f;
    __syncthreads();
    COVO_STAMP(a, 1);

    // adjoint, serial in t, 13 lanes of warp 0
    if (tid < 32) {
        for (int t = H - 1; t >= 1; --t) {
            if (tid < NX) {
                float acc = cg[t * NX + tid];
                const float* At = G + t * NX * NZP;
                const float* ln = lam + (t + 1) * NX;
#pragma unroll
                for (int j = 0; j < NX; ++j) acc = fmaf(At[j * NZP + tid], ln[j], acc);
                lam[t * NX + tid] = acc;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    COVO_STAMP(a, 2);

    // Lagrangian Hessians W_t (packed upper, 153 entries); the 14 loads of one output are in flight together
    for (int i = tid; i < H * NPAIR; i += blockDim.x) {
        int t = i / NPAIR, pi = i - t * NPAIR;
        const float* T = wsb + (long long)t * rec;
        float tv[14];
#pragma unroll
        for (int k = 0; k < 14; ++k) tv[k] = __ldg(T + k * NPAIR + pi);
        float acc = (t >= 1) ? tv[13] : 0.f;
        const float* ln = lam + (t + 1) * NX;
#pragma unroll
        for (int k = 0; k < NX; ++k) acc = fmaf(ln[k], tv[k], acc);
        W[i] = acc;
    }
    for (int i = tid; i < NX * NX; i += blockDim.x) P[i] = 0.f;
    __syncthreads();
    COVO_STAMP(a, 3);

    // backward recursion for P (13x13), S_t (13x4), D_t (4x4): only the first ten warps take part (289 threads
    // of work), on their own named barrier -- two 10-warp barriers per step instead of two 32-warp ones
    if (tid < 320) {
        for (int t = H - 1; t >= 0; --t) {
            const float* Gt = G + t * NX * NZP;
            if (tid < NX * NZ) {  // X = P [A B]
                int r = tid / NZ, c = tid - r * NZ;
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < NX; ++k) acc = fmaf(P[r * NX + k], Gt[k * NZP + c], acc);
                X[r * NZP + c] = acc;
            }
            asm volatile("bar.sync 1, 320;" ::: "memory");
            if (tid < NZ * NZ) {  // W + [A B]^T X
                int r = tid / NZ, c = tid - r * NZ;
                if (!(r >= NX && c < NX)) {  // lower-left block is the transpose of S, not needed
                    int lo = min(r, c), hi = max(r, c);
                    float acc = W[t * NPAIR + pair_index(lo, hi)];
#pragma unroll
                    for (int k = 0; k < NX; ++k) acc = fmaf(Gt[k * NZP + r], X[k * NZP + c], acc);
                    if (r < NX && c < NX) P[r * NX + c] = acc;
                    else if (r < NX) S[(t * NX + r) * 4 + (c - NX)] = acc;
                    else D[t * 16 + (r - NX) * 4 + (c - NX)] = acc;
                }
            }
            asm volatile("bar.sync 1, 320;" ::: "memory");
        }
    }
    __syncthreads();
    COVO_STAMP(a, 4);

    // export [A_t | B_t] (13 x 20), S_t (13 x 4) and D_t (4 x 4) for the forward-chain kernel: they overwrite the
    // head of each per-step record of the workspace, which this CTA has fully consumed above
    {
        float* wso = a.workspace + (long long)env * H * rec;
        for (int i = tid; i < H * kFwRec; i += kAsmThreads) {
            const int t = i / kFwRec, rr = i - t * kFwRec;
            float v;
            if (rr < NX * NZP) v = G[t * NX * NZP + rr];
            else if (rr < NX * NZP + NX * 4) v = S[t * NX * 4 + (rr - NX * NZP)];
            else v = D[t * 16 + (rr - NX * NZP - NX * 4)];
            wso[(long long)t * rec + rr] = v;
        }
    }
    COVO_STAMP(a, 5);
}

// ---------------------------------------------------------------------------------------------
// Forward chains R[I, J] = Phi^T S_J, Phi <- A_J Phi (Phi = d x_J / d u_{I,c}, starts as column c of B_I), J > I, and
// the diagonal blocks R[I, I] = D_I.  One HALF-WARP per chain (I, c): lane r < 13 forms row r of A_J Phi, lanes 0..3
// also the output column q; the half-warp re-assembles Phi with 13 shuffles.  A lone warp issues roughly one
// dependent instruction every four cycles, so the step time is its instruction count: ~60 here (a quad-per-chain
// version needed ~120, a thread-per-chain version 250).  Two chains per warp, one warp per CTA, CTAs spread
// over the SMs.
// ---------------------------------------------------------------------------------------------
constexpr int kFwThreads = 128;
constexpr int kFwChains = 2;  // per CTA (single environments: every chain at single-warp latency on an SM of its own)
// Batches of environments: 40 chains (20 warps) per CTA share one staged copy of the records -- 5 CTAs per environment instead of 100,
// a twentieth of the staging traffic (3.3 MB per environment with 2 chains per CTA) and 60 resident warps per SM instead of 3.
constexpr int kFwChainsBatch = 40, kFwThreadsBatch = 640;

template <int kFwChains, int kFwThreads>
__global__ void __launch_bounds__(kFwThreads) hess_forward_kernel_t(const HessianArgs a) {
    extern __shared__ __align__(16) float fsm[];  // [H][kFwRec]
    const int env = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int H = a.H, n = 4 * H;
    const int rec = 14 * NPAIR + 14 * NZ;
    const float* wsb = a.workspace + (long long)env * H * rec;
    const int id0 = blockIdx.x * kFwChains;  // first chain of this CTA
    const int Imin = id0 >> 2;
    // stage the records J >= Imin (16-byte pieces, all threads)
    {
        const int v4 = kFwRec / 4;
        for (int i = tid; i < (H - Imin) * v4; i += kFwThreads) {
            const int t = Imin + i / v4, q4 = i - (t - Imin) * v4;
            unsigned d = (unsigned)__cvta_generic_to_shared(fsm + t * kFwRec + 4 * q4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(wsb + (long long)t * rec + 4 * q4) : "memory");
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (tid >= 16 * kFwChains) return;
    const int id = id0 + (tid >> 4), r = lane & 15;  // chain, row of Phi owned by this lane (13..15: idle rows)
    if (id >= n) return;  // whole half-warps drop out together (n is even)
    const int I = id >> 2, c = id & 3;
    const unsigned hmask = 0xFFFFu << (lane & 16);
    const int base = lane & 16;
    float* Rg = a.R + (long long)env * n * n;
    const int rr = min(r, NX - 1), q = r & 3;
    float phi[NX];
    {
        const float* GI = fsm + I * kFwRec;
#pragma unroll
        for (int k = 0; k < NX; ++k) phi[k] = GI[k * NZP + NX + c];
        // D is symmetric up to round-off; symmetrise so R is exactly symmetric
        const float* DI = fsm + I * kFwRec + NX * NZP + NX * 4;
        if (r < 4) Rg[(long long)id * n + 4 * I + q] = 0.5f * (DI[c * 4 + q] + DI[q * 4 + c]);
    }
    for (int J = I + 1; J < H; ++J) {
        const float* GJ = fsm + J * kFwRec;
        const float* SJ = GJ + NX * NZP;
        // output column q (lanes 0..3 store it)
        float r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int k = 0; k < NX; k += 2) {
            r0 = fmaf(phi[k], SJ[k * 4 + q], r0);
            if (k + 1 < NX) r1 = fmaf(phi[k + 1], SJ[(k + 1) * 4 + q], r1);
        }
        const float rq = r0 + r1;
        if (r < 4) {
            Rg[(long long)id * n + 4 * J + q] = rq;
            Rg[(long long)(4 * J + q) * n + id] = rq;
        }
        if (J == H - 1) break;
        // row rr of A_J Phi
        const float4 a0 = *reinterpret_cast<const float4*>(GJ + rr * NZP);
        const float4 a1 = *reinterpret_cast<const float4*>(GJ + rr * NZP + 4);
        const float4 a2 = *reinterpret_cast<const float4*>(GJ + rr * NZP + 8);
        const float a12 = GJ[rr * NZP + 12];
        float e0 = a0.x * phi[0], e1 = a0.y * phi[1];
        e0 = fmaf(a0.z, phi[2], e0);
        e1 = fmaf(a0.w, phi[3], e1);
        e0 = fmaf(a1.x, phi[4], e0);
        e1 = fmaf(a1.y, phi[5], e1);
        e0 = fmaf(a1.z, phi[6], e0);
        e1 = fmaf(a1.w, phi[7], e1);
        e0 = fmaf(a2.x, phi[8], e0);
        e1 = fmaf(a2.y, phi[9], e1);
        e0 = fmaf(a2.z, phi[10], e0);
        e1 = fmaf(a2.w, phi[11], e1);
        e0 = fmaf(a12, phi[12], e0);
        const float np = e0 + e1;
#pragma unroll
        for (int k = 0; k < NX; ++k) phi[k] = __shfl_sync(hmask, np, base + k);
    }
}

size_t hessian_assemble_smem(int H) {
    size_t f = (size_t)H * NX * NZP + H * NX + (H + 1) * NX + (size_t)H * NPAIR + H * NX * 4 + H * 16 + NX * NX + NX * NZP;
    return f * sizeof(float);
}

cudaError_t launch_hessian(const HessianArgs& a, int n_env, cudaStream_t st) {
    dim3 g1(a.H, n_env);
    hess_local_kernel<<<g1, 160, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    size_t smem = hessian_assemble_smem(a.H);
    static size_t configured[32] = {};
    e = ensure_smem_attr(hess_assemble_kernel, smem, configured);
    if (e != cudaSuccess) return e;
    hess_assemble_kernel<<<n_env, kAsmThreads, smem, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t fsmem = (size_t)a.H * kFwRec * sizeof(float);
    static size_t configured_fw[32] = {}, configured_fwb[32] = {};
    if (n_env > 18) {
        e = ensure_smem_attr(hess_forward_kernel_t<kFwChainsBatch, kFwThreadsBatch>, fsmem, configured_fwb);
        if (e != cudaSuccess) return e;
        hess_forward_kernel_t<kFwChainsBatch, kFwThreadsBatch>
            <<<dim3((4 * a.H + kFwChainsBatch - 1) / kFwChainsBatch, n_env), kFwThreadsBatch, fsmem, st>>>(a);
    } else {
        e = ensure_smem_attr(hess_forward_kernel_t<kFwChains, kFwThreads>, fsmem, configured_fw);
        if (e != cudaSuccess) return e;
        hess_forward_kernel_t<kFwChains, kFwThreads><<<dim3((4 * a.H + kFwChains - 1) / kFwChains, n_env), kFwThreads, fsmem, st>>>(a);
    }
    return cudaGetLastError();
}

const void* hess_local_kernel_address() { return reinterpret_cast<const void*>(&hess_local_kernel); }

}  // namespace covo
