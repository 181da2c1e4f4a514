"""Host-side JAX-compatible PRNG keys (threefry2x32), enough for `fit(key=...)`:
PRNGKey(seed) and split(key, num) as jax.random does them (MAIN:275, 311).  The per-regression
split / normal draws run on the device (csrc/prng.cu)."""
import numpy as np

_U32 = np.uint32


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def _threefry2x32(key, x0, x1):
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    with np.errstate(over="ignore"):
        k0, k1 = _U32(key[0]), _U32(key[1])
        ks = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in rot[i % 2]:
                x0 = x0 + x1
                x1 = (x1 << _U32(r)) | (x1 >> _U32(32 - r))
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def split(key, num=2):
    key = np.asarray(key, dtype=_U32)
    counts = np.arange(2 * num, dtype=_U32)
    o0, o1 = _threefry2x32(key, counts[:num].copy(), counts[num:].copy())
    return np.concatenate([o0, o1]).reshape(num, 2)
