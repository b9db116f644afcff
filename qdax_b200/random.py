"""Host-facing mirror of the slice of `jax.random` that QDax's MAP-Elites path uses
(/root/reference: README.md:78-87, qdax/core/map_elites.py:81,133,177,181,214,241).

Keys are two uint32 words (the `jax.random.key_data` layout) held in a NumPy array on the host: key
derivation is control-plane glue (a handful of Threefry blocks per generation), while every *stream* of random
numbers (`bits`, `uniform`, `normal`) is produced on the GPU by libqdx.so.  Threefry-2x32-20 with JAX's
partitionable ("fold-like") derivation rules: split(key, n)[i] = threefry(key, counter=(hi32(i), lo32(i))).
"""

from __future__ import annotations

from typing import Sequence, Union

import ctypes as C

import numpy as np
import torch

from qdax_b200 import _native

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = 0xFFFFFFFF


def _threefry2x32(k0: int, k1: int, c0: int, c1: int):
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0, x1 = (c0 + ks[0]) & _M32, (c1 + ks[1]) & _M32
    for g in range(5):
        for r in _ROT[g % 2]:
            x0 = (x0 + x1) & _M32
            x1 = ((x1 << r) | (x1 >> (32 - r))) & _M32
            x1 ^= x0
        x0 = (x0 + ks[(g + 1) % 3]) & _M32
        x1 = (x1 + ks[(g + 2) % 3] + g + 1) & _M32
    return x0, x1


def key(seed: int) -> np.ndarray:
    """jax.random.key(seed): key data (hi32(seed), lo32(seed))."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & _M32], dtype=np.uint32)


PRNGKey = key


def key_data(k) -> np.ndarray:
    return np.asarray(k, dtype=np.uint32)


def split(k, num: int = 2) -> np.ndarray:
    """jax.random.split(key, num) -> (num, 2) uint32; rows unpack like `key, subkey = split(key)`.  Host-side
    (libqdx.so's qdx_host_split; `_threefry2x32` above is the same block in Python, kept for num >= 2^31)."""
    k0, k1 = _native.key_words(k)
    out = np.empty((num, 2), dtype=np.uint32)
    if num < (1 << 31):
        _native.call("qdx_host_split", C.c_uint32(k0), C.c_uint32(k1), C.c_int32(num), C.c_void_p(out.ctypes.data))
        return out
    for i in range(num):
        out[i] = _threefry2x32(k0, k1, (i >> 32) & _M32, i & _M32)
    return out


def _device(device) -> torch.device:
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("qdax_b200.random streams are generated on the GPU (no CPU fallback) and no CUDA device is visible")
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def bits(k, shape: Sequence[int] = (), device=None) -> torch.Tensor:
    n = int(np.prod(shape)) if len(tuple(shape)) else 1
    return _native.random_stream(k, n, 0, _device(device)).reshape(tuple(shape))


def uniform(k, shape: Sequence[int] = (), dtype=torch.float32, minval: float = 0.0, maxval: float = 1.0, device=None) -> torch.Tensor:
    """jax.random.uniform(key, shape, float32, minval, maxval)."""
    if dtype != torch.float32:
        raise TypeError("float32 only")
    n = int(np.prod(shape)) if len(tuple(shape)) else 1
    return _native.random_stream(k, n, 1, _device(device), float(minval), float(maxval)).reshape(tuple(shape))


def normal(k, shape: Sequence[int] = (), dtype=torch.float32, device=None) -> torch.Tensor:
    """jax.random.normal(key, shape, float32)."""
    if dtype != torch.float32:
        raise TypeError("float32 only")
    n = int(np.prod(shape)) if len(tuple(shape)) else 1
    return _native.random_stream(k, n, 2, _device(device)).reshape(tuple(shape))
