"""Pins the oracle's PRNG: Random123 KATs for Threefry-2x32-20, and values recalled from the JAX
documentation for jax.random.split / normal under jax_threefry_partitionable=True (the default of the
jax 0.8.0 pinned by the reference, uv.lock:1128-1129)."""
import numpy as np

from oracle import jax_prng as jr

KATS = [  # Random123 kat_vectors, threefry2x32 20 rounds: (ctr, key) -> out
    ((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6B200159, 0x99BA4EFE)),
    ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
    ((0x243F6A88, 0x85A308D3), (0x13198A2E, 0x03707344), (0xC4923A9C, 0x483DF7A0)),
]


def test_threefry_kats_numpy_and_c(co):
    for ctr, key, out in KATS:
        o = jr.threefry2x32(key[0], key[1], ctr[0], ctr[1])
        assert (int(o[0][0]), int(o[1][0])) == out
        assert co.threefry2x32(key[0], key[1], ctr[0], ctr[1]) == out


def test_key_layout():
    assert jr.key(42).tolist() == [0, 42]
    assert jr.key((7 << 32) | 9).tolist() == [7, 9]


def test_jax_doc_values():
    # jax docs ("Pseudorandom numbers" tutorial / jax.random module docs), jax >= 0.5 defaults
    assert jr.split(jr.key(0)).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]
    assert np.float32(jr.normal(jr.key(0), (1,))[0]) == np.float32(1.6226422)
    assert np.float32(jr.normal(jr.key(42), (1,))[0]) == np.float32(-0.028304616)
    # SURVEY.md 8c derived value
    assert jr.split(jr.key(42)).tolist() == [[1832780943, 270669613], [64467757, 2916123636]]


def test_c_oracle_matches_numpy_stream(co):
    k = jr.key(2026)
    assert (co.split(k, 7) == jr.split(k, 7)).all()
    assert (co.random_bits(k, 4097) == jr.random_bits(k, (4097,))).all()
    assert (co.uniform(k, 4097) == jr.uniform(k, (4097,))).all()
    assert (co.uniform(k, 257, -2.0, 3.0) == jr.uniform(k, (257,), -2.0, 3.0)).all()
    a, b = co.normal(k, 1 << 16), jr.normal(k, (1 << 16,))
    assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3)) < 1e-6
    assert abs(float(a.mean())) < 0.02 and abs(float(a.std()) - 1) < 0.02


def test_uniform_range_and_normal_edges():
    bits = np.array([0, 0xFFFFFFFF, 0x000001FF, 0x00000200], dtype=np.uint32)
    f = jr.bits_to_unit_float(bits)
    assert f[0] == 0.0 and f[1] == np.float32(1 - 2.0**-23) and f[2] == 0.0 and f[3] == np.float32(2.0**-23)
    assert np.isfinite(jr.erfinv_f32(np.array([-0.99999994, 0.99999994], dtype=np.float32))).all()


def test_spec_math_accuracy(co):
    rng = np.random.default_rng(1)
    x = -rng.random(200000).astype(np.float32)
    x = x[x > -1]
    ref = np.log1p(x.astype(np.float64))
    assert np.max(np.abs(co.math_probe(0, x) - ref) / np.abs(ref)) < 3e-7
    xs = -np.logspace(-12, -1, 20000).astype(np.float32)
    ref = np.log1p(xs.astype(np.float64))
    assert np.max(np.abs(co.math_probe(0, xs) - ref) / np.abs(ref)) < 3e-7
    th = ((rng.random(200000) - 0.5) * 2000).astype(np.float32)
    assert np.abs(co.math_probe(2, th) - np.sin(th.astype(np.float64))).max() < 2e-7
    assert np.abs(co.math_probe(3, th) - np.cos(th.astype(np.float64))).max() < 2e-7
    u = (rng.random(200000) * 2 - 1).astype(np.float32)
    assert np.max(np.abs(co.math_probe(1, u) - jr.erfinv_f32(u)) / np.maximum(np.abs(jr.erfinv_f32(u)), 1e-6)) < 2e-6


def test_permutation_and_choices():
    p = jr.permutation_indices(jr.key(3), 100)
    assert sorted(p.tolist()) == list(range(100))
    c = jr.choice_no_replace(jr.key(3), 100, 10)
    assert len(set(c.tolist())) == 10 and (c == p[:10]).all()
    r = jr.choice_uniform_replace(jr.key(4), 100, 1000)
    assert r.min() >= 0 and r.max() < 100
