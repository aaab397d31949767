"""JAX-compatible counter PRNG on the host.

The reference threads ``jax.random`` keys through everything that draws (trajectory generators
dynamics/utils.py:87-130, 183-251; reset / noisy state envs/quadrotor.py:265-312, 314-351; the samplers
controllers/covo.py:212-221, controllers/mppi.py:53-61).  This module restates the part of ``jax.random`` those call
sites use, for JAX's default ``threefry2x32`` implementation in its legacy (non-partitionable) mode -- the default of
every JAX release that still has ``jax.tree_map``, which the reference calls (controllers/covo.py:240):

  PRNGKey(seed)            -> uint32[2] = [seed >> 32, seed & 0xffffffff]
  random_bits(key, shape)  -> Threefry-2x32-20 over counters arange(size), first half paired with second half,
                              odd sizes padded with one zero counter
  split(key, num)          -> random_bits(key, (num, 2))
  uniform(key, shape, lo, hi) -> bitcast((bits >> 9) | 0x3f800000) - 1, scaled, clamped to >= lo
  normal(key, shape)       -> sqrt(2) * erfinv(uniform(lo = nextafter(-1, 0), hi = 1)), float32 erfinv by the
                              polynomial XLA expands it to (Giles' single-precision approximation)

JAX is not installable in the build container, so the integers are pinned by the Random123 known-answer vectors of
Threefry-2x32-20 and by values printed in JAX's own documentation (tests/test_jaxrng.py); the float stage can differ
from XLA's in the last ulp of log / sqrt.  The device generator (csrc/rng.cuh, threefry path) produces the same stream.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
F = np.float32

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def is_key(x) -> bool:
    """True for what the reference would call a PRNGKey: two uint32 words."""
    return isinstance(x, np.ndarray) and x.dtype == np.uint32 and x.shape == (2,)


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al., SC'11); x0, x1 uint32 arrays of equal shape."""
    with np.errstate(over="ignore"):
        k0, k1 = U32(key[0]), U32(key[1])
        ks = (k0, k1, U32(k0 ^ k1 ^ U32(0x1BD11BDA)))
        x0 = (np.asarray(x0, U32) + ks[0]).astype(U32)
        x1 = (np.asarray(x1, U32) + ks[1]).astype(U32)
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = (x0 + x1).astype(U32)
                x1 = _rotl(x1, r) ^ x0
            x0 = (x0 + ks[(g + 1) % 3]).astype(U32)
            x1 = (x1 + ks[(g + 2) % 3] + U32(g + 1)).astype(U32)
    return x0, x1


def random_bits(key, shape=()) -> np.ndarray:
    n = int(np.prod(shape, dtype=np.int64))
    c = np.arange(n, dtype=U32)
    if n % 2:
        c = np.concatenate([c, np.zeros(1, U32)])
    m = c.size // 2
    y0, y1 = threefry2x32(key, c[:m], c[m:])
    return np.concatenate([y0, y1])[:n].reshape(shape)


def split(key, num: int = 2) -> np.ndarray:
    return random_bits(key, (num, 2))


def _unit_floats(bits):
    return ((bits >> U32(9)) | U32(0x3F800000)).view(F) - F(1.0)


def uniform(key, shape=(), minval=0.0, maxval=1.0) -> np.ndarray:
    lo, hi = F(minval), F(maxval)
    f = _unit_floats(random_bits(key, shape))
    return np.maximum(lo, f * (hi - lo) + lo).astype(F)


_ERFINV_CENTRAL = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
                   -0.00417768164, 0.246640727, 1.50140941)
_ERFINV_TAIL = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
                0.00943887047, 1.00167406, 2.83297682)


def erfinv32(x) -> np.ndarray:
    x = np.asarray(x, F)
    w = (-np.log((F(1.0) - x) * (F(1.0) + x))).astype(F)

    def horner(t, cs):
        p = np.full_like(t, F(cs[0]))
        for c in cs[1:]:
            p = F(c) + p * t
        return p

    with np.errstate(invalid="ignore"):
        p = np.where(w < F(5.0), horner(w - F(2.5), _ERFINV_CENTRAL), horner(np.sqrt(w) - F(3.0), _ERFINV_TAIL))
    return (p * x).astype(F)


def normal(key, shape=()) -> np.ndarray:
    lo = np.nextafter(F(-1.0), F(0.0))
    return (F(np.sqrt(2.0)) * erfinv32(uniform(key, shape, lo, 1.0))).astype(F)


def sample_keys(act_key, n_samples: int) -> np.ndarray:
    """act_keys = jax.random.split(act_key, N) (controllers/covo.py:213, mppi.py:54)."""
    return split(act_key, n_samples)


def covo_normals(act_key, n_samples: int, n: int) -> np.ndarray:
    """The standard normals behind ``vmap(multivariate_normal)(split(act_key, N))`` of controllers/covo.py:213-221:
    row i = normal(act_keys[i], (n,))."""
    keys = split(act_key, n_samples)
    return np.stack([normal(k, (n,)) for k in keys])


def mppi_normals(act_key, n_samples: int, horizon: int, u_dim: int = 4) -> np.ndarray:
    """controllers/mppi.py:53-61: per sample i, keys = split(act_keys[i], H); element (i, h) = normal(keys[h], (u,))."""
    out = np.empty((n_samples, horizon, u_dim), F)
    for i, k in enumerate(split(act_key, n_samples)):
        for h, kh in enumerate(split(k, horizon)):
            out[i, h] = normal(kh, (u_dim,))
    return out
