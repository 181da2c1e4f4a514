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


def as_key(key):
    """`fit(key=...)` accepts what the reference accepts (MAIN:275, 311): a legacy uint32[2] JAX key (any array-like,
    including a jax.Array), a new-style typed key (`jax.random.key`, unwrapped through `jax.random.key_data`), or
    None (PRNGKey(0)).  An integer is taken as a seed."""
    if key is None:
        return PRNGKey(0)
    if isinstance(key, (int, np.integer)):
        return PRNGKey(int(key))
    try:
        arr = np.asarray(key)
    except Exception:  # noqa: BLE001  typed keys refuse __array__
        arr = None
    if arr is None or arr.dtype == object or arr.shape != (2,):
        data = None
        for getter in ("_base_array", "key_data"):
            data = getattr(key, getter, None)
            if data is not None:
                break
        if data is None:
            try:
                import jax
                data = jax.random.key_data(key)
            except Exception as e:  # noqa: BLE001
                raise TypeError(f"cannot interpret {type(key).__name__} as a threefry key (uint32[2])") from e
        arr = np.asarray(data() if callable(data) else data)
    if arr.shape != (2,):
        raise TypeError(f"a threefry key has shape (2,), got {arr.shape}")
    return (arr.astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(_U32)
