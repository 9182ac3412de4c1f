"""Host-side restatement of the `jax.random` calls the reference's reset path makes, so that a
seed produces the same obstacles / start / goal positions as the reference without JAX:

    jr.PRNGKey(seed), jr.split(key, n), jr.uniform(key, shape, minval=, maxval=)     (float32, x64 off)

Algorithm (public, jax/_src/prng.py): keys are uint32[2]; every draw is Threefry-2x32 (20 rounds, Salmon et al.
SC'11).  Two stream layouts exist and the reference pins neither (requirements.txt: `jax>=0.4.14`):

* LEGACY (`jax_threefry_partitionable=False`, the default of the 0.4.x line the reference was written on, and
  this module's default): the counter array `iota(n)` is split in two halves (x0 = first half, x1 = second half,
  zero padded when odd); split(key, n) = threefry(key, iota(2n)).reshape(n, 2); random bits = the two output halves
  concatenated.
* PARTITIONABLE (`jax_threefry_partitionable=True`, the default from JAX 0.5.0 on): element i of a draw of any
  shape uses the 64-bit row-major index i as the counter pair (hi(i), lo(i)); split(key, n)[i] is the output PAIR
  of threefry(key, (hi, lo)) and 32-bit random bits are y0 ^ y1 of that pair.  Select it with
  set_partitionable(True) or GCBF_THREEFRY_PARTITIONABLE=1 (host sampler, oracle and device reset kernel follow).

uniform takes 32 random bits per element, keeps the top 23 as the mantissa of a float in [1, 2), subtracts 1 and
scales: max(minval, f * (maxval - minval) + minval).

Pinned by known answers (tests/test_oracle.py): the Random123 Threefry-2x32-20 vectors for the block function (both
modes use it), and for the LEGACY layout the values jax prints for split(PRNGKey(0)), split(PRNGKey(42)) and
uniform(PRNGKey(0)).  The PARTITIONABLE layout is restated from the published algorithm and cross-checked between
three independent implementations here (this module, oracle/reset.py, the device kernel) but no jax-printed value
for it is available offline: it is unpinned until tests/golden/ref_io_*.npz (which record the mode) exist.
"""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import numpy as np

import os

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_U32 = np.uint32
PARTITIONABLE = os.environ.get("GCBF_THREEFRY_PARTITIONABLE", "0") == "1"


def set_partitionable(flag: bool) -> bool:
    """Select the threefry stream layout (see the module docstring); returns the previous setting."""
    global PARTITIONABLE
    old, PARTITIONABLE = PARTITIONABLE, bool(flag)
    return old


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


def _threefry_pairs(key: np.ndarray, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Partitionable layout: (y0, y1) of threefry(key, (hi(i), lo(i))) for i < n (n < 2**32 here, so hi = 0).
    key [2] -> two [n] arrays;  keys [E, 2] -> two [E, n] arrays."""
    lo = np.arange(n, dtype=_U32)
    hi = np.zeros(n, dtype=_U32)
    key = np.asarray(key, dtype=_U32)
    if key.ndim == 1:
        return threefry2x32(key[0], key[1], hi, lo)
    return threefry2x32(key[:, 0:1], key[:, 1:2], hi[None], lo[None])


def split(key: np.ndarray, num: int = 2) -> np.ndarray:
    """key [2] -> uint32 [num, 2];  keys [E, 2] -> [E, num, 2]."""
    key = np.asarray(key, dtype=_U32)
    if PARTITIONABLE:
        y0, y1 = _threefry_pairs(key, num)
        return np.stack([y0, y1], axis=-1)
    out = _threefry_2x32(key, np.arange(2 * num, dtype=_U32))
    return out.reshape(num, 2) if key.ndim == 1 else out.reshape(key.shape[0], num, 2)


def random_bits(key: np.ndarray, shape: Sequence[int]) -> np.ndarray:
    """key [2] -> uint32 `shape`;  keys [E, 2] -> [E, *shape]."""
    n = int(np.prod(shape)) if len(shape) else 1
    key = np.asarray(key, dtype=_U32)
    if PARTITIONABLE:
        y0, y1 = _threefry_pairs(key, n)
        out = y0 ^ y1
    else:
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
