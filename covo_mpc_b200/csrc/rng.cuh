// Counter-based Gaussian field for "production mode": eps[i][c] is a pure function of
// (seed, stream, GLOBAL sample index i, column c), so an N-sharded run on G GPUs draws exactly the
// samples a 1-GPU run draws.  Philox-4x32-10 (Salmon et al., SC'11) + Box-Muller.
//
// This replaces jax.random.split / multivariate_normal's normal draw (controllers/covo.py:212-217,
// mppi.py:53-60).  It is NOT JAX's Threefry stream (un-pinned third-party arithmetic);
// seed-identical parity with the reference goes through the explicit-eps entry points instead.
// The oracle restates this generator in oracle/oracle_np.py: philox_normals.
#pragma once
#include <stdint.h>

namespace covo {

__host__ __device__ __forceinline__ void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// four standard normals for (global sample i, column block b): columns 4b .. 4b+3
__device__ __forceinline__ void philox_normal4(unsigned long long seed, uint32_t stream, uint32_t i, uint32_t b,
                                               float z[4]) {
    uint32_t c[4] = {i, b, stream, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float s = 2.3283064365386963e-10f;  // 2^-32
    float u0 = (__uint2float_rn(c[0]) + 0.5f) * s, u1 = (__uint2float_rn(c[1]) + 0.5f) * s;
    float u2 = (__uint2float_rn(c[2]) + 0.5f) * s, u3 = (__uint2float_rn(c[3]) + 0.5f) * s;
    // fast intrinsics: MUFU lg2 / rsq / sin / cos; the field only has to be Gaussian, and the oracle is fed the
    // device's own draws when bit-for-bit agreement matters (covo_debug_eps)
    const float l0 = -2.0f * __logf(u0), l1 = -2.0f * __logf(u2);
    const float r0 = l0 * rsqrtf(fmaxf(l0, 1e-30f)), r1 = l1 * rsqrtf(fmaxf(l1, 1e-30f));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u1 - 3.141592653589793f, &s0, &c0);
    __sincosf(6.283185307179586f * u3 - 3.141592653589793f, &s1, &c1);
    z[0] = r0 * c0;
    z[1] = r0 * s0;
    z[2] = r1 * c1;
    z[3] = r1 * s1;
}

}  // namespace covo
