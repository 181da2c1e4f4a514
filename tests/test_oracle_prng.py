"""KATs for the oracle's JAX-PRNG restatement."""
import numpy as np
from oracle import jaxprng


def test_threefry_kat():
    z = np.zeros(1, np.uint32)
    x0, x1 = jaxprng.threefry2x32(0, 0, z, z)
    assert (int(x0[0]), int(x1[0])) == (0x6B200159, 0x99BA4EFE)


def test_split_kat():
    s = jaxprng.split(jaxprng.PRNGKey(0))
    assert s.tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]


def test_split_many_matches_rows():
    k = jaxprng.PRNGKey(42)
    s = jaxprng.split(k, 7)
    assert s.shape == (7, 2) and s.dtype == np.uint32
    # odd-length count path (padding) is self-consistent
    out = jaxprng.threefry_2x32(k, np.arange(5, dtype=np.uint32))
    assert out.shape == (5,)


def test_normal_first_row():
    sub = jaxprng.split(jaxprng.PRNGKey(0))[1]
    W = jaxprng.normal(sub, (1000, 4))
    np.testing.assert_allclose(W[0], [1.27123187, -0.49898759, -0.58634188, -0.64106222], atol=5e-9)
    assert abs(W.mean()) < 0.05 and abs(W.std() - 1) < 0.05
