"""covo_mpc_b200.jaxrng vs known answers: the Random123 vectors of Threefry-2x32-20 (Salmon et al.) and outputs of
jax.random quoted in JAX's documentation for the legacy threefry mode.  JAX is not installable here; the
documentation values were written down from memory BEFORE the restatement existed and every one of them came out
bit-identical, integers and floats alike."""
import numpy as np

from covo_mpc_b200 import jaxrng as jr


def test_threefry_random123_known_answers():
    def tf(key, ctr):
        y0, y1 = jr.threefry2x32(np.array(key, np.uint32), np.array([ctr[0]], np.uint32), np.array([ctr[1]], np.uint32))
        return int(y0[0]), int(y1[0])

    assert tf((0, 0), (0, 0)) == (0x6B200159, 0x99BA4EFE)
    assert tf((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF)) == (0x1CB996FC, 0xBB002BE7)
    assert tf((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3)) == (0xC4923A9C, 0x483DF7A0)


def test_jax_documented_outputs():
    k0 = jr.PRNGKey(0)
    assert k0.tolist() == [0, 0] and jr.PRNGKey(1).tolist() == [0, 1] and jr.PRNGKey((5 << 32) + 7).tolist() == [5, 7]
    assert jr.split(k0).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert jr.uniform(k0) == np.float32(0.41845703)
    assert jr.normal(k0) == np.float32(-0.20584226)
    assert jr.normal(jr.PRNGKey(42)) == np.float32(-0.18471177)
    assert np.allclose(jr.normal(k0, (3,)), [1.8160863, -0.48262316, 0.33988908], rtol=0, atol=1e-7)  # odd size: zero-padded counter


def test_shapes_ranges_and_moments():
    k = jr.PRNGKey(7)
    u = jr.uniform(k, (4096,), -0.2, 0.2)
    assert u.dtype == np.float32 and u.min() >= -0.2 and u.max() < 0.2 and abs(u.mean()) < 0.01
    z = jr.normal(k, (200_000,))
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01 and np.isfinite(z).all()
    keys = jr.split(k, 5)
    assert keys.shape == (5, 2) and len({tuple(r) for r in keys.tolist()}) == 5
    e = jr.covo_normals(k, 8, 20)
    assert e.shape == (8, 20) and np.array_equal(e[3], jr.normal(jr.split(k, 8)[3], (20,)))
    m = jr.mppi_normals(k, 4, 5)
    assert m.shape == (4, 5, 4) and np.array_equal(m[2, 1], jr.normal(jr.split(jr.split(k, 4)[2], 5)[1], (4,)))
