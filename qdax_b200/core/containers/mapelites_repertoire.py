"""MAP-Elites repertoire -- mirrors qdax/core/containers/mapelites_repertoire.py of the reference
(compute_cvt_centroids :30-72, compute_euclidean_centroids :75-108, get_cells_indices :111-137,
MapElitesRepertoire :140-388) with the insertion rule executed by libqdx.so:

  cells   = nearest centroid, first-index argmin  (grid fast path with exact re-rank, or brute force)
  offer   = per-cell best offspring as a packed 64-bit (fitness-key, index) atomicMax  == segment_max + tie-break
  commit  = scatter of the winners' rows into the HBM-resident repertoire

The repertoire keeps value semantics: `add` returns a new object and leaves `self` untouched (arrays are
copied device-side, 4 MB at K=10^4, D=100) unless `_donate=True` (what jax buffer donation would do), which
MAPElites.update / scan use for the carried repertoire.
"""

from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from qdax_b200 import _native
from qdax_b200 import random as qrandom
from qdax_b200 import tree_util
from qdax_b200.core.emitters.repertoire_selectors.selector import Selector
from qdax_b200.core.containers.ga_repertoire import GARepertoire

TIE_BREAKS = ("first", "last")


def lloyd_cvt_centroids(x: torch.Tensor, num_centroids: int, num_iterations: int = 100) -> Tuple[torch.Tensor, int]:
    """CVT of the samples x (N, Dd) in [0, 1) by Lloyd iterations on the GPU: initial centroids = the
    first `num_centroids` samples (i.i.d. uniform, so a random subset), assignment by the cell-assignment kernels (tcgen05
    pass + exact re-rank for 8 <= Dd <= 32, FP32 brute force otherwise; first index on ties), order-free integer-atomic
    update (qdx_kmeans_accumulate / qdx_kmeans_update), stop when no sample changes cell.  Returns (centroids, iterations
    run).  Bit-reproducible: tests/test_kmeans.py checks it against the NumPy restatement of the same rule."""
    x = _native.require_cuda(x, "samples")
    N, Dd = x.shape
    K = int(num_centroids)
    if K > N:
        raise ValueError("more centroids than samples")
    dev = x.device
    cent = x[:K].clone()
    acc = torch.empty(K * Dd, dtype=torch.int64, device=dev)
    count = torch.empty(K, dtype=torch.int32, device=dev)
    changed = torch.empty(1, dtype=torch.int32, device=dev)
    prev = None
    it = 0
    for it in range(1, num_iterations + 1):
        cells = _native.cells(x, cent, None, allow_index=False)
        _native.call("qdx_kmeans_accumulate", _native._ptr(x), _native._ptr(cells), _native._ptr(prev), C.c_int64(N), C.c_int32(Dd),
                     C.c_int64(K), _native._ptr(acc), _native._ptr(count), _native._ptr(changed), _native._stream())
        if prev is not None and int(changed.item()) == 0:
            break                                   # fixed point: the centroids are the means of their own cells
        new_cent = torch.empty_like(cent)           # a fresh tensor: the per-tessellation caches are keyed on the object
        _native.call("qdx_kmeans_update", _native._ptr(acc), _native._ptr(count), _native._ptr(cent), C.c_int64(K), C.c_int32(Dd),
                     _native._ptr(new_cent), _native._stream())
        cent, prev = new_cent, cells
    return cent, it


def compute_cvt_centroids(num_descriptors: int, num_init_cvt_samples: int, num_centroids: int,
                          minval: Union[float, List[float]], maxval: Union[float, List[float]], key, device=None,
                          backend: str = "sklearn", num_iterations: int = 100) -> torch.Tensor:
    """CVT centroids (reference :30-72): uniform samples in the unit cube -> k-means -> rescale to [minval, maxval].
    backend="sklearn" (default) is the reference's own call, scikit-learn KMeans(k-means++, n_init=1) on the host;
    backend="gpu" runs Lloyd iterations on the device (lloyd_cvt_centroids) -- the only practical way to K = 50 000 centroids
    in 32-D -- and gives a different (equally valid) CVT: scikit-learn's initialisation and stopping rule are not reproduced.
    Set-up code, not part of the generation step."""
    ks = qrandom.split(key)
    key, subkey = ks[0], ks[1]
    x = qrandom.uniform(subkey, (num_init_cvt_samples, num_descriptors), device=device)
    if backend == "gpu":
        cent, _ = lloyd_cvt_centroids(x, num_centroids, num_iterations)
        lo_t = torch.as_tensor(minval, dtype=torch.float32, device=cent.device)
        hi_t = torch.as_tensor(maxval, dtype=torch.float32, device=cent.device)
        return (cent * (hi_t - lo_t) + lo_t).contiguous()
    if backend != "sklearn":
        raise ValueError("backend must be 'sklearn' or 'gpu'")
    from numpy.random import RandomState
    from sklearn.cluster import KMeans

    k_means = KMeans(init="k-means++", n_clusters=num_centroids, n_init=1, random_state=RandomState(qrandom.key_data(key)))
    k_means.fit(x.cpu().numpy())
    lo = np.asarray(minval, dtype=np.float32)
    hi = np.asarray(maxval, dtype=np.float32)
    cent = k_means.cluster_centers_.astype(np.float32) * (hi - lo) + lo
    return torch.from_numpy(cent.astype(np.float32)).to(x.device)


def _linspace_f32(start: float, stop: float, num: int) -> torch.Tensor:
    # jnp.linspace in float32: start*(1-step) + stop*step, step = iota(div)/div, endpoint appended
    s, e = torch.tensor(start, dtype=torch.float32), torch.tensor(stop, dtype=torch.float32)
    if num == 1:
        return s.reshape(1)
    div = num - 1
    step = torch.arange(div, dtype=torch.float32) / torch.tensor(float(div), dtype=torch.float32)
    out = s * (1.0 - step) + e * step
    return torch.cat([out, e.reshape(1)])


def compute_euclidean_centroids(grid_shape: Tuple[int, ...], minval: Union[float, List[float]],
                                maxval: Union[float, List[float]], device=None) -> torch.Tensor:
    """Centroids of a regular grid (reference :75-108; meshgrid indexing 'xy', float32).  Host-side set-up;
    the result is uploaded to the current CUDA device."""
    lin = []
    for n in grid_shape:
        offset = 1 / (2 * n)
        lin.append(_linspace_f32(offset, 1.0 - offset, n))
    meshes = torch.meshgrid(*lin, indexing="xy")
    cent = torch.stack([m.reshape(-1) for m in meshes], dim=-1)
    lo = torch.as_tensor(minval, dtype=torch.float32)
    hi = torch.as_tensor(maxval, dtype=torch.float32)
    cent = cent * (hi - lo) + lo
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    return cent.contiguous().to(device)


def get_cells_indices(batch_of_descriptors: torch.Tensor, centroids: torch.Tensor) -> torch.Tensor:
    """Nearest-centroid cell of each descriptor, first minimum (reference :111-137).  int32 (batch,)."""
    d = _native.require_cuda(batch_of_descriptors, "batch_of_descriptors")
    c = _native.require_cuda(centroids, "centroids")
    return _native.cells(d, c, _native.grid_of(centroids))


class MapElitesRepertoire(GARepertoire):
    """Repertoire of MAP-Elites (reference :140-160): genotypes (K, D), fitnesses (K, 1) with -inf for empty
    cells, descriptors (K, Dd), centroids (K, Dd), extra_scores, static keys_extra_scores."""

    def __init__(self, genotypes, fitnesses, descriptors, centroids, extra_scores: Optional[Dict[str, Any]] = None,
                 keys_extra_scores: Tuple[str, ...] = (), tie_break: str = "first"):
        super().__init__(genotypes, fitnesses, extra_scores, keys_extra_scores)
        self.descriptors = descriptors
        self.centroids = centroids
        if tie_break not in TIE_BREAKS:
            raise ValueError(f"tie_break must be one of {TIE_BREAKS}")
        self.tie_break = tie_break

    # ---- plumbing ------------------------------------------------------------------------------------
    def _grid(self) -> Optional[_native.Grid]:
        return _native.grid_of(self.centroids)

    def _clone_state(self) -> "MapElitesRepertoire":
        if tree_util.is_tree(self.genotypes):
            flat, spec = self._packed_genotypes()
            genotypes = tree_util.unpack(flat.clone(), spec)
        else:
            genotypes = self.genotypes.clone()
        new = self.replace(genotypes=genotypes, fitnesses=self.fitnesses.clone(), descriptors=self.descriptors.clone(),
                           extra_scores={k: v.clone() for k, v in self.extra_scores.items()})
        return new

    # ---- reference surface ---------------------------------------------------------------------------
    def select(self, key, num_samples: int, selector: Optional[Selector] = None) -> "MapElitesRepertoire":
        """reference :162-171."""
        return super().select(key, num_samples, selector)

    def sample(self, key, num_samples: int) -> torch.Tensor:
        """Alias named by the north star: genotypes of `select(key, num_samples)`."""
        return self.select(key, num_samples).genotypes

    def add(self, batch_of_genotypes, batch_of_descriptors, batch_of_fitnesses, batch_of_extra_scores=None, *,
            _donate: bool = False, _metrics_out: Optional[torch.Tensor] = None, _qd_offset: float = 0.0,
            _cells: Optional[torch.Tensor] = None) -> "MapElitesRepertoire":
        """Add a batch to the repertoire (reference :173-266): cell assignment, per-cell best offspring
        (segment_max, ties to the first / last offspring index per `tie_break`), strict improvement over the
        current occupant, scatter of genotype / fitness / descriptor / filtered extra scores."""
        if batch_of_extra_scores is None:
            batch_of_extra_scores = {}
        self._raise_if_error()
        extras = self.filter_extra_scores(batch_of_extra_scores)
        new = self if _donate else self._clone_state()
        rep_g, spec = new._packed_genotypes()
        if _donate and spec is not None and not tree_util.is_packed_view(new.genotypes, rep_g):
            # the leaves were not views of one packed buffer (a repertoire built through the constructor / replace()):
            # pack() copied them, so the rows committed below would land in a temporary -- rebind the leaves to it
            new.genotypes = tree_util.unpack(rep_g, spec)
        if spec is not None:                      # pytree genotype: one packed row per individual (reference :234-240)
            g2, _ = tree_util.pack(batch_of_genotypes, spec)
            g = g2
        else:
            g = _native.require_cuda(batch_of_genotypes, "batch_of_genotypes")
            g2 = g.reshape(g.shape[0], -1)
        B = g2.shape[0]
        d = _native.require_cuda(batch_of_descriptors, "batch_of_descriptors")
        f = _native.require_cuda(batch_of_fitnesses, "batch_of_fitnesses").reshape(-1)
        if f.numel() != B or d.shape[0] != B:
            raise ValueError("batch size mismatch between genotypes, descriptors and fitnesses")
        K = new.centroids.shape[0]
        rep_f = new.fitnesses.reshape(-1)
        if rep_g.shape[1] != g2.shape[1]:
            raise ValueError("genotype dimension mismatch")
        ws = new._workspace()
        first = new.tie_break == "first"
        if _cells is None:
            _native.cells(d, new.centroids, new._grid(), ws, rep_f, f, offer=True, first_wins=first)
        else:
            _native.offer_cells(_native.require_cuda(_cells, "cells", torch.int32), f, ws, rep_f, first_wins=first)
        added = None
        if extras:
            added = torch.full((K,), -1, dtype=torch.int32, device=g.device)
        _native.commit(ws, g2, f, d, rep_g, rep_f, new.descriptors, first_wins=first, qd_offset=_qd_offset,
                       metrics_out=_metrics_out, added_cells=added)
        if extras:  # reference :250-257; generic (non-hot) path
            cells_changed = torch.nonzero(added >= 0).reshape(-1)
            src = added[cells_changed].long()
            new.extra_scores = {k: _scatter_rows(new.extra_scores[k], cells_changed, v, src) for k, v in extras.items()}
        if _native.DEBUG_SYNC:
            ws.check()
        return new

    @classmethod
    def init(cls, genotypes, fitnesses, descriptors, centroids, *args, extra_scores=None,
             keys_extra_scores: Tuple[str, ...] = (), tie_break: str = "first", **kwargs) -> "MapElitesRepertoire":
        """reference :268-326."""
        if extra_scores is None and len(args) > 0 and isinstance(args[0], dict):
            extra_scores = args[0]  # MAPElites passes extra_scores positionally (map_elites.py:124-130)
        if extra_scores is None:
            extra_scores = {}
        extra_scores = {k: v for k, v in extra_scores.items() if k in keys_extra_scores}
        first_extra = {k: v[0] for k, v in extra_scores.items()}
        first_genotype = tree_util.tree_map(lambda x: x[0], genotypes) if tree_util.is_tree(genotypes) else genotypes[0]   # :304
        rep = cls.init_default(genotype=first_genotype, centroids=centroids, one_extra_score=first_extra,
                               keys_extra_scores=keys_extra_scores, tie_break=tie_break)
        return rep.add(genotypes, descriptors, fitnesses, extra_scores, _donate=True)

    @classmethod
    def init_default(cls, genotype, centroids, one_extra_score=None, keys_extra_scores: Tuple[str, ...] = (),
                     tie_break: str = "first") -> "MapElitesRepertoire":
        """reference :328-388: fitness -inf, genotypes 0, descriptors 0."""
        centroids = _native.require_cuda(centroids, "centroids")
        if one_extra_score is None:
            one_extra_score = {}
        one_extra_score = {k: v for k, v in one_extra_score.items() if k in keys_extra_scores}
        K = centroids.shape[0]
        dev = centroids.device
        if tree_util.is_tree(genotype):           # :342-347: zeros_like every leaf with a leading K; here one packed buffer
            spec = tree_util.spec_of(genotype, batched=False)
            default_genotypes = tree_util.unpack(torch.zeros((K, spec.total), dtype=torch.float32, device=dev), spec)
        else:
            default_genotypes = torch.zeros((K,) + tuple(genotype.shape), dtype=torch.float32, device=dev)
        return cls(
            genotypes=default_genotypes,
            fitnesses=torch.full((K, 1), float("-inf"), dtype=torch.float32, device=dev),
            descriptors=torch.zeros_like(centroids),
            centroids=centroids,
            extra_scores={k: torch.zeros((K,) + tuple(v.shape), dtype=v.dtype, device=dev) for k, v in one_extra_score.items()},
            keys_extra_scores=keys_extra_scores,
            tie_break=tie_break,
        )


def _scatter_rows(dst: torch.Tensor, rows: torch.Tensor, src: torch.Tensor, src_rows: torch.Tensor) -> torch.Tensor:
    out = dst.clone()
    out[rows] = src[src_rows].to(out.dtype)
    return out
