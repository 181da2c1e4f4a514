"""Oracle (test infrastructure): JAX 0.4.28 threefry2x32 PRNG restated in NumPy.

Third-party dependency restated: jax==0.4.28 (`jax/_src/prng.py`,
`jax/_src/random.py`), used by the reference at REG:47 (random.split),
REG:99 (random.normal), MAIN:275,311 (PRNGKey / split).  Default
(non-partitionable) threefry implementation, 64-bit mode (x64 enabled by the
reference's __init__.py:18-25).

Known-answer tests (tests/test_oracle_prng.py):
  threefry2x32(key=(0,0), ctr=(0,0))           -> (0x6b200159, 0x99ba4efe)
  split(PRNGKey(0))                             -> [[4146024105, 967050713],[2718843009,1272950319]]
"""
import numpy as np
from scipy.special import erfinv

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def _rotl(x, r):
    return ((x << _U32(r)) | (x >> _U32(32 - r))).astype(_U32)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds. x0/x1 are uint32 arrays (lanes)."""
    with np.errstate(over="ignore"):
        k0 = _U32(k0)
        k1 = _U32(k1)
        ks = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
        x0 = (x0.astype(_U32) + ks[0]).astype(_U32)
        x1 = (x1.astype(_U32) + ks[1]).astype(_U32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(_U32)
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(_U32)
            x1 = (x1 + ks[(i + 2) % 3] + _U32(i + 1)).astype(_U32)
    return x0, x1


def threefry_2x32(key, counts):
    """jax `threefry_2x32(keypair, count)`: odd lengths are zero padded, first
    half goes to lane 0, second half to lane 1, output is concat(out0, out1)."""
    counts = np.asarray(counts, dtype=_U32).ravel()
    odd = counts.size % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, _U32)])
    half = counts.size // 2
    o0, o1 = threefry2x32(key[0], key[1], counts[:half], counts[half:])
    out = np.concatenate([o0, o1])
    return out[:-1] if odd else out


def split(key, num: int = 2) -> np.ndarray:
    """random.split (REG:47, MAIN:311): threefry over arange(2*num) -> (num, 2)."""
    return threefry_2x32(key, np.arange(2 * num, dtype=_U32)).reshape(num, 2)


def random_bits64(key, size: int) -> np.ndarray:
    b = threefry_2x32(key, np.arange(2 * size, dtype=_U32))
    return (b[:size].astype(np.uint64) << np.uint64(32)) | b[size:].astype(np.uint64)


def uniform64(key, shape, lo, hi):
    size = int(np.prod(shape))
    bits = random_bits64(key, size)
    fl = ((bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)).view(np.float64) - 1.0
    return np.maximum(lo, fl * (hi - lo) + lo).reshape(shape)


def normal(key, shape):
    """random.normal float64 (REG:99): sqrt(2) * erfinv(uniform(nextafter(-1,0), 1))."""
    lo = np.nextafter(np.float64(-1.0), np.float64(0.0))
    u = uniform64(key, shape, lo, np.float64(1.0))
    return np.sqrt(2.0) * erfinv(u)
