"""JAX side of the thin jax.ffi layer over libqdx.so (csrc/qdx_xla_ffi.cc): registration of the XLA FFI handlers and the
drop-in functions a QDax checkout would call instead of its jnp code.

NOT EXECUTED in this repository's environment: jax / jaxlib are absent from the image and the GPU box (no wheel, no
network), so importing this module raises ImportError here.  With a JAX install:

    make -C qdax_b200/csrc ffi XLA_FFI_INCLUDE=$(python -c "import jax; print(jax.ffi.include_dir())")
    import qdax_b200.jax_ffi as qffi
    qdax.core.containers.mapelites_repertoire.get_cells_indices = qffi.get_cells_indices      # mapelites_repertoire.py:111
    (rep_arrays, ws, key), metrics = jax.lax.scan(qffi.make_scan_update(...), carry, (), length=n)   # map_elites.py:197-225

Everything below only shapes ffi_call signatures; the computation is the C ABI of include/qdx.h, which the tests and
benches of this repository drive through ctypes (qdax_b200/_native.py)."""

from __future__ import annotations

import ctypes
import os

import numpy as np

try:
    import jax
    import jax.numpy as jnp
except ImportError as e:  # pragma: no cover - the only path reachable in this image
    raise ImportError("qdax_b200.jax_ffi needs jax >= 0.4.38 (jax.ffi); this environment has no JAX. The same kernels are "
                      "reachable through qdax_b200's torch-tensor API or plain ctypes (INTEGRATION.md section 3).") from e

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.environ.get("QDX_FFI_LIB_PATH", os.path.join(_HERE, "libqdx_xla_ffi.so"))
TASK_IDS = {"arm": 0, "rastrigin": 1, "sphere": 2}

if not os.path.exists(_SHIM):
    raise ImportError(f"{_SHIM} not found: build it with `make -C qdax_b200/csrc ffi XLA_FFI_INCLUDE=$(python -c "
                      "\"import jax; print(jax.ffi.include_dir())\")`")
_lib = ctypes.cdll.LoadLibrary(_SHIM)
_core = ctypes.cdll.LoadLibrary(os.path.join(_HERE, "libqdx.so"))
for _name, _sym in (("qdx_cells", "QdxCells"), ("qdx_score", "QdxScore"), ("qdx_add", "QdxAdd"),
                    ("qdx_isoline_variation", "QdxIsolineVariation"), ("qdx_scan_update", "QdxScanUpdate")):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_lib, _sym)), platform="CUDA")


def workspace_bytes(num_centroids: int) -> int:
    n = ctypes.c_int64(0)
    rc = _core.qdx_workspace_bytes(ctypes.c_int64(num_centroids), ctypes.byref(n))
    if rc != 0:
        raise RuntimeError(f"qdx_workspace_bytes rc={rc}")
    return int(n.value)


def new_workspace(num_centroids: int) -> "jax.Array":
    """Zero-initialised per-repertoire workspace, carried next to the repertoire (all zeros is the initial state)."""
    return jnp.zeros((workspace_bytes(num_centroids),), dtype=jnp.uint8)


def get_cells_indices(batch_of_descriptors, centroids):
    """Drop-in for qdax/core/containers/mapelites_repertoire.py:111-137."""
    out = jax.ShapeDtypeStruct((batch_of_descriptors.shape[0],), jnp.int32)
    return jax.ffi.ffi_call("qdx_cells", out)(batch_of_descriptors.astype(jnp.float32), centroids.astype(jnp.float32))


def scoring_function(task: str, desc_dim: int = 2):
    """arm / rastrigin / sphere `*_scoring_function(params, key)` -- qdax/tasks/arm.py:41-50, standard_functions.py:27-48."""
    tid = np.int32(TASK_IDS[task])

    def fn(params, key=None):
        B = params.shape[0]
        outs = (jax.ShapeDtypeStruct((B,), jnp.float32), jax.ShapeDtypeStruct((B, desc_dim), jnp.float32))
        f, d = jax.ffi.ffi_call("qdx_score", outs)(params.astype(jnp.float32), task=tid)
        return f, d, {}

    return fn


def isoline_variation(x1, x2, key, iso_sigma, line_sigma, minval=None, maxval=None):
    """Drop-in for qdax/core/emitters/mutation_operators.py:175-226 (single-array genotype)."""
    out = jax.ShapeDtypeStruct(x1.shape, jnp.float32)
    return jax.ffi.ffi_call("qdx_isoline_variation", out)(
        x1, x2, jax.random.key_data(key).astype(jnp.uint32), iso_sigma=np.float32(iso_sigma), line_sigma=np.float32(line_sigma),
        has_min=np.int32(minval is not None), minval=np.float32(minval or 0.0), has_max=np.int32(maxval is not None),
        maxval=np.float32(maxval or 0.0))


def add(ws, genotypes, fitnesses, descriptors, centroids, batch_of_genotypes, batch_of_descriptors, batch_of_fitnesses,
        first_wins: bool = True, qd_offset: float = 0.0):
    """MapElitesRepertoire.add on raw arrays (mapelites_repertoire.py:173-266): returns (ws, genotypes, fitnesses (K,),
    descriptors, cells (B,), metrics (4,)); the first four alias their inputs (in place under jit with donation)."""
    K, B = centroids.shape[0], batch_of_genotypes.shape[0]
    outs = (jax.ShapeDtypeStruct(ws.shape, jnp.uint8), jax.ShapeDtypeStruct(genotypes.shape, jnp.float32),
            jax.ShapeDtypeStruct((K,), jnp.float32), jax.ShapeDtypeStruct(descriptors.shape, jnp.float32),
            jax.ShapeDtypeStruct((B,), jnp.int32), jax.ShapeDtypeStruct((4,), jnp.float32))
    return jax.ffi.ffi_call("qdx_add", outs, input_output_aliases={0: 0, 1: 1, 2: 2, 3: 3})(
        ws, genotypes, fitnesses.reshape(K), descriptors, centroids, batch_of_genotypes, batch_of_fitnesses.reshape(B),
        batch_of_descriptors, first_wins=np.int32(first_wins), qd_offset=np.float32(qd_offset))


def make_scan_update(task: str, batch_size: int, grid_shape, axes, stride, lo, hi, iso_sigma: float, line_sigma: float,
                     minval=None, maxval=None, first_wins: bool = True, qd_offset: float = 0.0):
    """`scan_update(carry, _)` for jax.lax.scan with carry = ((genotypes, fitnesses (K,), descriptors), centroids, ws, key data):
    one whole generation of MAPElites.scan_update (map_elites.py:197-225) per call, fused emitter + task + grid."""
    attrs = dict(batch=np.int64(batch_size), task=np.int32(TASK_IDS[task]), iso_sigma=np.float32(iso_sigma), line_sigma=np.float32(line_sigma),
                 has_min=np.int32(minval is not None), minval=np.float32(minval or 0.0), has_max=np.int32(maxval is not None),
                 maxval=np.float32(maxval or 0.0), first_wins=np.int32(first_wins), qd_offset=np.float32(qd_offset),
                 n0=np.int32(grid_shape[0]), n1=np.int32(grid_shape[1]), stride0=np.int32(stride[0]), stride1=np.int32(stride[1]),
                 lo0=np.float32(lo[0]), lo1=np.float32(lo[1]), hi0=np.float32(hi[0]), hi1=np.float32(hi[1]))

    def scan_update(carry, _=None):
        (g, f, d), centroids, ws, key = carry
        K, D = g.shape
        outs = (jax.ShapeDtypeStruct(ws.shape, jnp.uint8), jax.ShapeDtypeStruct((2,), jnp.uint32), jax.ShapeDtypeStruct(g.shape, jnp.float32),
                jax.ShapeDtypeStruct((K,), jnp.float32), jax.ShapeDtypeStruct(d.shape, jnp.float32),
                jax.ShapeDtypeStruct((batch_size, D), jnp.float32), jax.ShapeDtypeStruct((batch_size,), jnp.float32),
                jax.ShapeDtypeStruct((batch_size, 2), jnp.float32), jax.ShapeDtypeStruct((4,), jnp.float32))
        ws, key, g, f, d, _og, _of, _od, m = jax.ffi.ffi_call("qdx_scan_update", outs, input_output_aliases={0: 0, 2: 2, 3: 3, 4: 4})(
            ws, key, g, f, d, centroids, axes, **attrs)
        return ((g, f, d), centroids, ws, key), {"qd_score": m[0], "max_fitness": m[1], "coverage": m[2]}

    return scan_update
