"""CPU tests of the host-side logic: the C-ABI library loads and exports every symbol of include/qdx.h, the
selection-table closed form equals np.cumsum, the host key chain equals the oracle's, grid detection, the
collective helpers under gloo with world_size 2, and the product path refuses to run without CUDA."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import jax_prng as jr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from qdax_b200 import _lib

    header = open(os.path.join(ROOT, "include", "qdx.h")).read()
    declared = set(re.findall(r"^int (qdx_\w+)\(", header, flags=re.M))
    assert len(declared) >= 20
    h = _lib.lib()
    for name in declared:
        assert hasattr(h, name), f"{name} declared in include/qdx.h but not exported by libqdx.so"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == declared
    assert h.qdx_version() >= 100


def test_select_table_closed_form_equals_cumsum():
    from qdax_b200 import _native

    Ms = list(range(1, 3000)) + [4096, 9999, 10000, 10001, 50000, 65536, 100000, 1 << 20, 3000017]
    for M in Ms:
        T, nseg = _native.host_select_table(M)
        ref = np.cumsum(np.full(M, np.float32(1.0) / np.float32(M), dtype=np.float32), dtype=np.float32)
        assert np.array_equal(T, ref), M
        assert nseg <= 128
    rng = np.random.default_rng(0)
    for M in (1, 2, 3, 100, 7821, 10000, 50000):
        ref = np.cumsum(np.full(M, np.float32(1.0) / np.float32(M), dtype=np.float32), dtype=np.float32)
        u = rng.integers(0, 1 << 23, 20000).astype(np.float32) * np.float32(2.0**-23)
        r = (ref[-1] * (np.float32(1) - u)).astype(np.float32)
        assert np.array_equal(_native.host_select_rank(M, r), np.searchsorted(ref, r, side="left") + 1)
        assert np.array_equal(_native.host_select_rank(M, ref), np.arange(1, M + 1))


def test_host_key_chain_matches_oracle():
    from qdax_b200 import random as qr

    assert qr.key(42).tolist() == [0, 42]
    for seed in (0, 42, 2**40 + 7):
        k = qr.key(seed)
        assert (k == jr.key(seed)).all()
        for n in (1, 2, 3, 8):
            assert (qr.split(k, n) == jr.split(k, n)).all()
    assert qr.split(qr.key(0)).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]   # JAX docs value
    key, sub = qr.split(qr.key(42))
    assert key.tolist() == [1832780943, 270669613] and sub.tolist() == [64467757, 2916123636]


def test_host_generation_keys_follow_the_reference_split_chain():
    """qdx_host_generation_keys == the jax.random.split chain spelled out in SURVEY.md Appendix B (map_elites.py:177,214,241;
    distributed_map_elites.py:124; standard_emitters.py:55; uniform_selector.py:48; mutation_operators.py:205,220)."""
    from qdax_b200 import _native

    def chain(emit):
        k1, k2, kv = jr.split(emit, 3)
        kv2, kl = jr.split(kv)
        return np.concatenate([jr.split(k1)[1], jr.split(k2)[1], kl, jr.split(kv2, 1)[0]]).astype(np.uint32)

    for seed in (0, 42, 123456789):
        key = jr.key(seed)
        upd = chain(jr.split(jr.split(key)[1])[1])                      # update: split -> ask: split -> emit
        assert np.array_equal(np.array(list(_native.host_generation_keys(_native.KEYMODE_UPDATE, key))), upd)
        dst = chain(jr.split(key)[1])                                   # distributed update: split -> emit
        assert np.array_equal(np.array(list(_native.host_generation_keys(_native.KEYMODE_DIST_UPDATE, key))), dst)
        assert np.array_equal(np.array(list(_native.host_generation_keys(_native.KEYMODE_EMIT, key))), chain(key))
        carry = np.array(key, dtype=np.uint32)                          # scan_update: key, subkey = split(key); update(subkey)
        ref_carry = np.array(key, dtype=np.uint32)
        for _ in range(3):
            got = np.array(list(_native.host_generation_keys(_native.KEYMODE_SCAN, None, carry)))
            nxt, sub = jr.split(ref_carry)
            assert np.array_equal(got, chain(jr.split(jr.split(sub)[1])[1]))
            ref_carry = np.array(nxt, dtype=np.uint32)
            assert np.array_equal(carry, ref_carry)


def test_no_cpu_fallback():
    import torch

    from qdax_b200 import _native
    from qdax_b200.tasks.arm import arm_scoring_function

    with pytest.raises(RuntimeError, match="CUDA only"):
        arm_scoring_function(torch.zeros(4, 8))
    with pytest.raises(RuntimeError, match="CUDA only"):
        _native.require_cuda(torch.zeros(3), "x")


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "qdax_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_grid_detection_host_side():
    import torch

    from qdax_b200 import _native
    from qdax_b200.core.containers.mapelites_repertoire import compute_euclidean_centroids
    from oracle import qdax_numpy as qn

    for shape in [(2, 2), (100, 100), (3, 5), (4, 5, 6), (7,), (1, 4)]:
        c = compute_euclidean_centroids(shape, 0.0, 1.0, device="cpu")
        assert np.array_equal(c.numpy(), qn.compute_euclidean_centroids(shape, 0.0, 1.0))
        g = _native.detect_grid(c)
        assert g is not None and sorted(g.n) == sorted(shape)
        k = np.arange(c.shape[0])
        rec = np.stack([g.axes.numpy()[sum(g.n[:d]):sum(g.n[:d + 1])][(k // max(g.stride[d], 1)) % g.n[d]] if g.n[d] > 1
                        else np.full(len(k), g.axes.numpy()[sum(g.n[:d])]) for d in range(len(shape))], axis=-1)
        assert np.array_equal(rec, c.numpy())
    c = compute_euclidean_centroids((0.0, 1.0) and (10, 10), [0.0, -1.0], [2.0, 1.0], device="cpu")   # per-dim ranges
    assert _native.detect_grid(c) is not None
    assert _native.detect_grid(torch.rand(100, 2)) is None                       # CVT-like
    shuffled = compute_euclidean_centroids((5, 5), 0.0, 1.0, device="cpu")[torch.randperm(25)]
    assert _native.detect_grid(shuffled) is None                                 # a grid in the wrong order is not a grid


GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["QDX_ROOT"])
from qdax_b200 import parallel
from oracle import c_oracle as co, jax_prng as jr, qdax_numpy as qn
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
# all_gather_rows: rank r's rows at [r*B, (r+1)*B)
x = torch.full((3, 2), float(rank)) + torch.arange(3.0)[:, None]
g = parallel.all_gather_rows(x)
assert g.shape == (3 * world, 2) and all(float(g[r * 3 + i, 0]) == r + i for r in range(world) for i in range(3))
# unsigned 64-bit max through a signed collective
rng = np.random.default_rng(rank)
keys = rng.integers(0, 2**64, 1000, dtype=np.uint64)
keys[:10] = 0
t = torch.from_numpy(keys.view(np.int64).copy())
parallel.all_reduce_max_u64_(t)
allk = np.stack([np.random.default_rng(r).integers(0, 2**64, 1000, dtype=np.uint64) for r in range(world)])
allk[:, :10] = 0
assert np.array_equal(t.numpy().view(np.uint64), allk.max(axis=0))
# disjoint-row merge keeps bit patterns (-0.0, NaN payloads)
st = torch.zeros(8, 4)
st[rank::world] = torch.tensor([-0.0, 1.5, float("nan"), -3.0])
parallel.all_reduce_disjoint_rows_(st)
assert np.array_equal(st.numpy().view(np.uint32)[:world], np.tile(np.array([-0.0, 1.5, np.nan, -3.0], np.float32).view(np.uint32), (world, 1)))
assert parallel.all_equal(st) and not parallel.all_equal(torch.full((4,), float(rank)))
# winners-only exchange == all-gather exchange == single-process oracle, simulated with oracle compute per rank:
# each rank offers its shard (global indices) into a packed-key table; max-reduce; the elected winners are the oracle's.
K, D, B = 64, 8, 200
cent = qn.compute_euclidean_centroids((8, 8), 0.0, 1.0)
rs = np.random.default_rng(100)
rep_f = np.where(rs.random(K) < 0.5, rs.standard_normal(K), -np.inf).astype(np.float32)
Fall = np.round(rs.standard_normal(world * B), 1).astype(np.float32)
Dall = rs.random((world * B, 2)).astype(np.float32)
Gall = rs.random((world * B, D)).astype(np.float32)
cells = co.cells(Dall, cent)
def okey(v):
    u = np.where(v == 0, np.float32(0), v).view(np.uint32).astype(np.uint64)
    return np.where(u & 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000)
tab = np.zeros(K, np.uint64)
for i in range(rank * B, (rank + 1) * B):
    if Fall[i] > rep_f[cells[i]]:
        tab[cells[i]] = max(tab[cells[i]], (okey(Fall[i:i+1])[0] << np.uint64(31)) | np.uint64((~np.uint32(i)) & 0x7FFFFFFF))
assert (tab < np.uint64(1) << np.uint64(63)).all()          # 63-bit keys: signed max == unsigned max
t = torch.from_numpy(tab.view(np.int64).copy())
parallel.all_reduce_max_i64_(t)
win = t.numpy().view(np.uint64)
_, f2, _, sidx = co.add(np.zeros((K, D)), rep_f, np.zeros((K, 2)), Gall, Fall, Dall, cells, "first")
exp = np.full(K, -1)
for i in range(world * B - 1, -1, -1):
    if sidx[i] < K: exp[sidx[i]] = i
got = np.where(win != 0, ((~(win & np.uint64(0x7FFFFFFF)).astype(np.uint32)) & np.uint32(0x7FFFFFFF)).astype(np.int64), -1)
assert np.array_equal(got, exp)
dist.barrier()
if rank == 0: print("GLOO_OK")
dist.destroy_process_group()
'''


def test_collective_helpers_gloo_world2(tmp_path):
    script = tmp_path / "gloo_worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, QDX_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "GLOO_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_jax_ffi_layer_is_gated_not_faked():
    """The jax.ffi shim cannot run here (no JAX in the image): the C++ side must compile to an empty translation unit without
    the XLA headers, and the Python side must fail loudly on import instead of pretending."""
    import importlib
    import subprocess

    src = os.path.join(ROOT, "qdax_b200", "csrc", "qdx_xla_ffi.cc")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    for handler in ("QdxCells", "QdxScore", "QdxAdd", "QdxIsolineVariation", "QdxScanUpdate"):
        assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({handler}," in text
    try:
        import jax  # noqa: F401
        have_jax = True
    except ImportError:
        have_jax = False
    if not have_jax:
        with pytest.raises(ImportError, match="jax"):
            importlib.import_module("qdax_b200.jax_ffi")


def test_multi_sample_scoring_and_mels_driver_wiring():
    """qdax/utils/sampling.py:111-152 and qdax/core/mels.py:33-60: sample s is scored with split(key, num_samples)[s], results are
    stacked on axis 1, and MELS is MAPElites with that wrapper on a MELSRepertoire (CPU tensors: no kernel is involved here)."""
    import functools

    import torch

    from qdax_b200 import random as qr
    from qdax_b200.core.containers.mels_repertoire import MELSRepertoire
    from qdax_b200.core.mels import MELS
    from qdax_b200.utils.sampling import multi_sample_scoring_function

    seen = []

    def scoring(x, key):
        seen.append(tuple(int(w) for w in key))
        noise = float(int(key[1]) % 7)
        return x.sum(dim=1) + noise, torch.stack([x[:, 0] + noise, x[:, 1]], dim=1), {"aux": x[:, :1] * noise, "tag": "t"}

    x = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    key = qr.key(5)
    f, d, extra = multi_sample_scoring_function(x, key, scoring, 3)
    assert f.shape == (4, 3) and d.shape == (4, 3, 2) and extra["aux"].shape == (4, 3, 1) and extra["tag"] == ["t"] * 3
    assert seen == [tuple(int(w) for w in k) for k in jr.split(jr.key(5), 3)]
    for s in range(3):
        fs, ds, _ = scoring(x, jr.split(jr.key(5), 3)[s])
        assert torch.equal(f[:, s], fs) and torch.equal(d[:, s], ds)
    me = MELS(scoring, emitter=None, metrics_function=lambda r: {}, num_samples=3)
    assert isinstance(me._scoring_function, functools.partial) and me._scoring_function.keywords["num_samples"] == 3
    assert me._repertoire_init.__func__ is MELSRepertoire.init.__func__ and me._num_samples == 3
    f2, d2, _ = me._scoring_function(x, key)
    assert torch.equal(f2, f) and torch.equal(d2, d)


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/qdx.h must compile as C99 and as C++ on its own (no torch / CUDA types)."""
    hdr = os.path.join(ROOT, "include", "qdx.h")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr], ["g++", "-std=c++11", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    includes = re.findall(r"^\s*#\s*include\s*[<\"]([^>\"]+)[>\"]", open(hdr).read(), flags=re.M)
    assert includes == ["stdint.h"], includes           # nothing but the C standard integer types
