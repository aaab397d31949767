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
// `env`: environment index of a batched handle (counter word 3), so that the environments of a batch draw independent fields;
// environment 0 is the stream of a single-environment handle
__device__ __forceinline__ void philox_normal4(unsigned long long seed, uint32_t stream, uint32_t i, uint32_t b,
                                               float z[4], uint32_t env = 0u) {
    uint32_t c[4] = {i, b, stream, env};
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

// ---------------------------------------------------------------------------------------------------------------------
// JAX-compatible stream: Threefry-2x32-20 with jax.random's legacy counter layout, so that a
// caller holding the reference's PRNGKey gets the draws jax.random.split / normal would have produced
// (controllers/covo.py:212-217, mppi.py:53-60).  The integers are pinned by Random123 / JAX-documentation known answers
// (tests/test_jaxrng.py, host twin covo_mpc_b200/jaxrng.py); the float stage may differ from XLA's in the last ulp.
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ __forceinline__ void threefry2x32_20(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
    const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    x0 += ks[0];
    x1 += ks[1];
#pragma unroll
    for (int g = 0; g < 5; ++g) {
        if (g & 1) {
            x0 += x1; x1 = rotl32(x1, 17) ^ x0;
            x0 += x1; x1 = rotl32(x1, 29) ^ x0;
            x0 += x1; x1 = rotl32(x1, 16) ^ x0;
            x0 += x1; x1 = rotl32(x1, 24) ^ x0;
        } else {
            x0 += x1; x1 = rotl32(x1, 13) ^ x0;
            x0 += x1; x1 = rotl32(x1, 15) ^ x0;
            x0 += x1; x1 = rotl32(x1, 26) ^ x0;
            x0 += x1; x1 = rotl32(x1, 6) ^ x0;
        }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
}

// element m of jax.random.random_bits(key, total) for EVEN total (legacy layout: counter j paired with j + total/2)
__host__ __device__ __forceinline__ uint32_t jax_bits_elem(uint32_t k0, uint32_t k1, uint32_t m, uint32_t total) {
    const uint32_t half = total >> 1;
    uint32_t x0 = m < half ? m : m - half, x1 = x0 + half;
    threefry2x32_20(k0, k1, x0, x1);
    return m < half ? x0 : x1;
}

// key i of jax.random.split(key, num)
__host__ __device__ __forceinline__ void jax_split_elem(uint32_t k0, uint32_t k1, uint32_t i, uint32_t num, uint32_t& o0,
                                                         uint32_t& o1) {
    o0 = jax_bits_elem(k0, k1, 2u * i, 2u * num);
    o1 = jax_bits_elem(k0, k1, 2u * i + 1u, 2u * num);
}

// jax.random.normal from 32 random bits: sqrt(2) * erfinv(u), u uniform on [nextafter(-1, 0), 1); erfinv = the
// single-precision polynomial XLA expands it to.  Explicit _rn arithmetic: no FMA contraction, like the host twin.
__device__ __forceinline__ float jax_normal_from_bits(uint32_t bits) {
    const float lo = -0.99999994f;
    float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
    float u = fmaxf(lo, __fadd_rn(__fmul_rn(f, 2.0f), lo));
    float w = -logf(__fmul_rn(1.0f - u, 1.0f + u));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = __fadd_rn(3.43273939e-07f, __fmul_rn(p, w));
        p = __fadd_rn(-3.5233877e-06f, __fmul_rn(p, w));
        p = __fadd_rn(-4.39150654e-06f, __fmul_rn(p, w));
        p = __fadd_rn(0.00021858087f, __fmul_rn(p, w));
        p = __fadd_rn(-0.00125372503f, __fmul_rn(p, w));
        p = __fadd_rn(-0.00417768164f, __fmul_rn(p, w));
        p = __fadd_rn(0.246640727f, __fmul_rn(p, w));
        p = __fadd_rn(1.50140941f, __fmul_rn(p, w));
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f;
        p = __fadd_rn(0.000100950558f, __fmul_rn(p, w));
        p = __fadd_rn(0.00134934322f, __fmul_rn(p, w));
        p = __fadd_rn(-0.00367342844f, __fmul_rn(p, w));
        p = __fadd_rn(0.00573950773f, __fmul_rn(p, w));
        p = __fadd_rn(-0.0076224613f, __fmul_rn(p, w));
        p = __fadd_rn(0.00943887047f, __fmul_rn(p, w));
        p = __fadd_rn(1.00167406f, __fmul_rn(p, w));
        p = __fadd_rn(2.83297682f, __fmul_rn(p, w));
    }
    return __fmul_rn(1.41421354f, __fmul_rn(p, u));
}

// Fills the standard normals of ONE sample as the reference's samplers draw them:
//   dense (CoVO, covo.py:213-221): act_keys = split(act_key, N); z[0:n] = normal(act_keys[i], (n,))
//   per-step (MPPI, mppi.py:53-61): keys = split(act_keys[i], H); z[4h:4h+4] = normal(keys[h], (4,))
// `emit(column, value)` receives every column once.  n = 4H is even by construction.
template <class Emit>
__device__ __forceinline__ void jax_sample_normals(uint32_t ak0, uint32_t ak1, uint32_t i_global, uint32_t n_total, int n, int H,
                                                   bool per_step, int part, int parts, Emit emit) {
    uint32_t s0, s1;
    jax_split_elem(ak0, ak1, i_global, n_total, s0, s1);
    if (!per_step) {
        const int half = n >> 1;
        for (int c = part; c < half; c += parts) {
            uint32_t x0 = (uint32_t)c, x1 = (uint32_t)(c + half);
            threefry2x32_20(s0, s1, x0, x1);
            emit(c, jax_normal_from_bits(x0));
            emit(c + half, jax_normal_from_bits(x1));
        }
    } else {
        for (int h = part; h < H; h += parts) {
            uint32_t h0, h1;
            jax_split_elem(s0, s1, (uint32_t)h, (uint32_t)H, h0, h1);
            uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;
            threefry2x32_20(h0, h1, a0, a1);
            threefry2x32_20(h0, h1, b0, b1);
            emit(4 * h + 0, jax_normal_from_bits(a0));
            emit(4 * h + 1, jax_normal_from_bits(b0));
            emit(4 * h + 2, jax_normal_from_bits(a1));
            emit(4 * h + 3, jax_normal_from_bits(b1));
        }
    }
}

}  // namespace covo
