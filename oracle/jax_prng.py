"""ORACLE (test infrastructure only -- never imported by the product path).

NumPy restatement of the parts of ``jax.random`` (jax==0.8.0, pinned in
/root/reference/uv.lock:1128-1129) that the MAP-Elites hot path of QDax 0.5.1
calls.  JAX is a third-party dependency that is NOT vendored in /root/reference
and is not installable in this image, so this file restates the *published*
algorithm (Threefry-2x32, 20 rounds, Salmon et al. SC'11; JAX's
``jax_threefry_partitionable=True`` key-derivation rules) and is anchored on

  * the Random123 known-answer vectors for Threefry-2x32-20 (tests/test_oracle_prng.py),
  * the reference call sites listed next to each function below.

PARITY UNPINNED at the jaxlib boundary: no reference test fixes a value of the
PRNG stream, and no jaxlib exists here to generate one.  tools/dump_jax_golden.py
regenerates the golden vectors under a real JAX install when one is available.

Call sites in the reference (file:line under /root/reference):
  split    : qdax/core/map_elites.py:81,133,177,181,214,241
             qdax/core/emitters/standard_emitters.py:55,65
             qdax/core/emitters/repertoire_selectors/uniform_selector.py:48
             qdax/core/emitters/mutation_operators.py:41,53,107,205,220
  uniform  : qdax/core/emitters/mutation_operators.py:54-60, README.md:82-87
  normal   : qdax/core/emitters/mutation_operators.py:207,210
  choice   : qdax/core/emitters/repertoire_selectors/uniform_selector.py:49-55
             qdax/core/emitters/mutation_operators.py:42-44,134
"""

from __future__ import annotations

import math

import numpy as np

U32 = np.uint32
F32 = np.float32

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = U32(0x1BD11BDA)


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, c0, c1):
    """Threefry-2x32, 20 rounds.  All arguments broadcastable uint32 arrays.

    Key schedule ks = (k0, k1, k0^k1^0x1BD11BDA); after round group g (4 rounds)
    inject (ks[(g+1)%3], ks[(g+2)%3] + g + 1).
    """
    with np.errstate(over="ignore"):
        k0 = np.atleast_1d(np.asarray(k0, dtype=U32))
        k1 = np.atleast_1d(np.asarray(k1, dtype=U32))
        x0 = np.atleast_1d(np.asarray(c0, dtype=U32)).copy()
        x1 = np.atleast_1d(np.asarray(c1, dtype=U32)).copy()
        ks = (k0, k1, k0 ^ k1 ^ _PARITY)
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x0 ^ x1
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + U32(g + 1)
    return x0, x1


def key(seed: int) -> np.ndarray:
    """jax.random.key(seed) -> key data words (hi32(seed), lo32(seed))."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=U32)


def _iota_2x32(n: int):
    i = np.arange(n, dtype=np.uint64)
    return (i >> np.uint64(32)).astype(U32), (i & np.uint64(0xFFFFFFFF)).astype(U32)


def split(k: np.ndarray, num: int = 2) -> np.ndarray:
    """jax.random.split in partitionable ("fold-like") mode:
    split(key, n)[i] = threefry(key, ctr=(hi32(i), lo32(i))) taken as a 2-word key."""
    hi, lo = _iota_2x32(num)
    o0, o1 = threefry2x32(k[0], k[1], hi, lo)
    return np.stack([o0, o1], axis=-1)


def random_bits(k: np.ndarray, shape) -> np.ndarray:
    """32-bit random bits, partitionable mode: bits[flat i] = o0 ^ o1 of
    threefry(key, ctr=(hi32(i), lo32(i)))."""
    shape = tuple(int(s) for s in np.atleast_1d(shape)) if not isinstance(shape, tuple) else shape
    n = int(np.prod(shape)) if len(shape) else 1
    hi, lo = _iota_2x32(n)
    o0, o1 = threefry2x32(k[0], k[1], hi, lo)
    return (o0 ^ o1).reshape(shape)


def bits_to_unit_float(bits: np.ndarray) -> np.ndarray:
    """f = bitcast((bits >> 9) | 0x3F800000) - 1.0  in [0, 1)."""
    fb = (bits >> U32(9)) | U32(0x3F800000)
    return fb.view(F32) - F32(1.0)


def uniform(k: np.ndarray, shape, minval=0.0, maxval=1.0) -> np.ndarray:
    """jax.random.uniform(key, shape, float32, minval, maxval)."""
    f = bits_to_unit_float(random_bits(k, tuple(shape)))
    lo = F32(minval)
    scale = F32(F32(maxval) - lo)
    return np.maximum(lo, (f * scale + lo).astype(F32))


# Giles' single-precision erfinv as used by XLA's ErfInv32 expansion.
_ERFINV_LT5 = (
    2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
    0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941,
)
_ERFINV_GE5 = (
    -0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
    0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682,
)


def erfinv_f32(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p((-(x * x)).astype(F32)).astype(F32)
        lt = w < F32(5.0)
        wl = (w - F32(2.5)).astype(F32)
        wg = (np.sqrt(w).astype(F32) - F32(3.0)).astype(F32)
        ww = np.where(lt, wl, wg).astype(F32)
        p = np.where(lt, F32(_ERFINV_LT5[0]), F32(_ERFINV_GE5[0])).astype(F32)
        for i in range(1, 9):
            c = np.where(lt, F32(_ERFINV_LT5[i]), F32(_ERFINV_GE5[i])).astype(F32)
            p = (c + (p * ww).astype(F32)).astype(F32)
        res = (p * x).astype(F32)
        edge = np.abs(x) == F32(1.0)
        res = np.where(edge, x * np.finfo(F32).max, res).astype(F32)
    return res


_NORMAL_LO = np.nextafter(F32(-1.0), F32(0.0), dtype=F32)
_SQRT2 = F32(np.sqrt(2))


def normal(k: np.ndarray, shape) -> np.ndarray:
    """jax.random.normal(key, shape, float32) = sqrt(2) * erfinv(uniform(-1+eps, 1))."""
    u = uniform(k, shape, _NORMAL_LO, 1.0)
    return (_SQRT2 * erfinv_f32(u)).astype(F32)


def choice_p_replace(k: np.ndarray, p: np.ndarray, num: int) -> np.ndarray:
    """jax.random.choice(key, arange(n), (num,), replace=True, p=p):
         cum = cumsum(p); r = cum[-1] * (1 - uniform(key, (num,))); searchsorted(cum, r, 'left').
    The cumsum is the sequential float32 running sum (np.cumsum semantics); XLA's
    summation order is unverifiable here -- this is the canonical rule."""
    cum = np.cumsum(np.asarray(p, dtype=F32), dtype=F32)
    u = uniform(k, (num,))
    r = (cum[-1] * (F32(1.0) - u).astype(F32)).astype(F32)
    return np.searchsorted(cum, r, side="left").astype(np.int32)


def permutation_rounds(n: int) -> int:
    """jax.random.permutation's number of sort rounds: ceil(3 ln n / ln(2^32 - 1))."""
    if n <= 1:
        return 1
    return int(np.ceil(3 * math.log(max(1, n)) / math.log(np.iinfo(np.uint32).max)))


def permutation_indices(k: np.ndarray, n: int) -> np.ndarray:
    """jax.random.permutation(key, n): `rounds` passes of
    key, sub = split(key); stable sort of the array by random_bits(sub, (n,))."""
    x = np.arange(n, dtype=np.int32)
    for _ in range(permutation_rounds(n)):
        ks = split(k)
        k, sub = ks[0], ks[1]
        sort_keys = random_bits(sub, (n,))
        order = np.argsort(sort_keys, kind="stable")
        x = x[order]
    return x


def choice_no_replace(k: np.ndarray, n: int, m: int) -> np.ndarray:
    """jax.random.choice(key, arange(n), (m,), replace=False) = permutation(key, n)[:m]."""
    return permutation_indices(k, n)[:m]


def choice_uniform_replace(k: np.ndarray, n: int, m: int) -> np.ndarray:
    """jax.random.choice(key, arange(n), (m,)) (replace=True, p=None) = randint(key, (m,), 0, n).

    jax.random.randint draws two 32-bit words per element from split(key) halves and
    combines them modulo the span (jax/_src/random.py `_randint`)."""
    ks = split(k)
    hi_bits = random_bits(ks[0], (m,)).astype(np.uint64)
    lo_bits = random_bits(ks[1], (m,)).astype(np.uint64)
    span = np.uint64(n)
    # multiplier = ((2^16 mod span)^2) mod span  == 2^32 mod span
    mult = np.uint64(((1 << 16) % n) ** 2 % n)
    with np.errstate(over="ignore"):
        off = ((hi_bits % span) * mult + (lo_bits % span)) % span
    return off.astype(np.int32)
