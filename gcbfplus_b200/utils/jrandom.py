"""Host-side restatement of the `jax.random` calls the reference's reset path makes, so that a
seed produces the same obstacles / start / goal positions as the reference without JAX:

    jr.PRNGKey(seed), jr.split(key, n), jr.uniform(key, shape, minval=, maxval=)     (float32, x64 off)

Algorithm (public, jax/_src/prng.py of the 0.4.x line the reference pins; `jax_threefry_partitionable` off,
its default there): keys are uint32[2]; every draw is Threefry-2x32 (20 rounds, Salmon et al. SC'11) over a
counter array `iota(n)` that is split in two halves (x0 = first half, x1 = second half, zero padded when odd);
split(key, n) = threefry(key, iota(2n)).reshape(n, 2); uniform takes 32 random bits per element, keeps the top
23 as the mantissa of a float in [1, 2), subtracts 1 and scales: max(minval, f * (maxval - minval) + minval).

Pinned by known answers (tests/test_oracle.py): the Random123 Threefry-2x32-20 vectors, and the values
jax prints for split(PRNGKey(0)) and uniform(PRNGKey(0)).
"""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_U32 = np.uint32


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0: int, k1: int, x0: np.ndarray, x1: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Threefry-2x32, 20 rounds.  k0, k1 scalars; x0, x1 uint32 arrays of equal shape."""
    with np.errstate(over="ignore"):
        k0, k1 = np.asarray(k0, dtype=_U32), np.asarray(k1, dtype=_U32)
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = x0.astype(_U32) + ks[0]
        x1 = x1.astype(_U32) + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + _U32(g + 1)
    return x0, x1


def _threefry_2x32(key: np.ndarray, count: np.ndarray) -> np.ndarray:
    """key uint32 [2] or [E, 2] (a batch of independent keys, the reference's vmap over keys);
    count uint32 [n] -> [n] or [E, n]."""
    flat = count.ravel().astype(_U32)
    n = flat.size
    odd = n % 2
    if odd:
        flat = np.concatenate([flat, np.zeros(1, _U32)])
    h = flat.size // 2
    key = np.asarray(key, dtype=_U32)
    if key.ndim == 1:
        y0, y1 = threefry2x32(key[0], key[1], flat[:h], flat[h:])
        return np.concatenate([y0, y1])[:n].reshape(count.shape)
    y0, y1 = threefry2x32(key[:, 0:1], key[:, 1:2], flat[None, :h], flat[None, h:])
    return np.concatenate([y0, y1], axis=1)[:, :n]


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def split(key: np.ndarray, num: int = 2) -> np.ndarray:
    """key [2] -> uint32 [num, 2];  keys [E, 2] -> [E, num, 2]."""
    key = np.asarray(key, dtype=_U32)
    out = _threefry_2x32(key, np.arange(2 * num, dtype=_U32))
    return out.reshape(num, 2) if key.ndim == 1 else out.reshape(key.shape[0], num, 2)


def random_bits(key: np.ndarray, shape: Sequence[int]) -> np.ndarray:
    """key [2] -> uint32 `shape`;  keys [E, 2] -> [E, *shape]."""
    n = int(np.prod(shape)) if len(shape) else 1
    key = np.asarray(key, dtype=_U32)
    out = _threefry_2x32(key, np.arange(n, dtype=_U32))
    return out.reshape(tuple(shape)) if key.ndim == 1 else out.reshape(key.shape[0], *shape)


def uniform(key: np.ndarray, shape: Sequence[int] = (), minval: float = 0.0, maxval: float = 1.0) -> np.ndarray:
    """float32 samples in [minval, maxval)."""
    f = np.float32
    bits = random_bits(key, shape)
    fl = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - f(1.0)
    lo, hi = f(minval), f(maxval)
    return np.maximum(lo, fl * (hi - lo) + lo).astype(np.float32)


def is_key(x) -> bool:
    return isinstance(x, np.ndarray) and x.dtype == _U32 and x.ndim in (1, 2) and x.shape[-1] == 2


def as_key(x: Union[int, np.ndarray]) -> np.ndarray:
    return x if is_key(x) else PRNGKey(int(x))
