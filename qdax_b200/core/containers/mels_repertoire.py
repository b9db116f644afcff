"""MAP-Elites Low-Spread repertoire -- mirrors qdax/core/containers/mels_repertoire.py:60-297 of the reference.

Every individual is evaluated `num_samples` times; `add` takes descriptors (batch, num_samples, Dd) and fitnesses
(batch, num_samples), files the individual under its most frequent cell, stores the mean fitness, the centroid of that
cell as descriptor and the spread (mean pairwise descriptor distance), and replaces an occupant only when the fitness is
higher AND the spread is not larger.  Native path: qdx_cells over all batch * num_samples descriptors, qdx_mels_offer
(mode / spread / mean + election), the streaming qdx_commit, qdx_scatter_rows_by_source for the spreads."""

from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Tuple

import torch

from qdax_b200 import _native
from qdax_b200 import tree_util
from qdax_b200.core.containers.mapelites_repertoire import MapElitesRepertoire, _scatter_rows


class MELSRepertoire(MapElitesRepertoire):
    """reference :60-87: MapElitesRepertoire + `spreads` (K,), inf for empty cells."""

    def __init__(self, genotypes, fitnesses, descriptors, centroids, spreads, extra_scores: Optional[Dict[str, Any]] = None,
                 keys_extra_scores: Tuple[str, ...] = (), tie_break: str = "first"):
        super().__init__(genotypes, fitnesses, descriptors, centroids, extra_scores, keys_extra_scores, tie_break)
        self.spreads = spreads

    def _clone_state(self) -> "MELSRepertoire":
        new = super()._clone_state()
        new.spreads = self.spreads.clone()
        return new

    def add(self, batch_of_genotypes, batch_of_descriptors, batch_of_fitnesses, batch_of_extra_scores=None, *,
            _donate: bool = False, **_unused) -> "MELSRepertoire":
        """reference :89-230."""
        if batch_of_extra_scores is None:
            batch_of_extra_scores = {}
        extras = self.filter_extra_scores(batch_of_extra_scores)
        f_all = _native.require_cuda(batch_of_fitnesses, "batch_of_fitnesses")
        if f_all.dim() != 2:
            raise ValueError("MELSRepertoire.add expects fitnesses of shape (batch_size, num_samples)")
        B, S = f_all.shape
        d_all = _native.require_cuda(batch_of_descriptors, "batch_of_descriptors").reshape(B * S, -1)
        self._raise_if_error()
        new = self if _donate else self._clone_state()
        rep_g, spec = new._packed_genotypes()
        if _donate and spec is not None and not tree_util.is_packed_view(new.genotypes, rep_g):
            new.genotypes = tree_util.unpack(rep_g, spec)        # foreign leaves were copied by pack(): rebind to the packed buffer
        if spec is not None:
            g2, _ = tree_util.pack(batch_of_genotypes, spec)
        else:
            g2 = _native.require_cuda(batch_of_genotypes, "batch_of_genotypes").reshape(B, -1)
        if g2.shape[0] != B or rep_g.shape[1] != g2.shape[1]:
            raise ValueError("batch size / genotype dimension mismatch")
        K, Dd = new.centroids.shape
        dev = f_all.device
        rep_f = new.fitnesses.reshape(-1)
        ws = new._workspace()
        first = new.tie_break == "first"
        cells_all = _native.cells(d_all, new.centroids, new._grid())                                    # :143-145
        cell = torch.empty(B, dtype=torch.int32, device=dev)
        f = torch.empty(B, dtype=torch.float32, device=dev)
        spread = torch.empty(B, dtype=torch.float32, device=dev)
        desc = torch.empty((B, Dd), dtype=torch.float32, device=dev)
        _native.call("qdx_mels_offer", _native._ptr(cells_all), _native._ptr(d_all), _native._ptr(f_all), C.c_int64(B), C.c_int32(S),
                     C.c_int32(Dd), _native._ptr(new.centroids), C.c_int64(K), ws.ptr, _native._ptr(rep_f), _native._ptr(new.spreads),
                     C.c_int32(first), _native._ptr(cell), _native._ptr(f), _native._ptr(spread), _native._ptr(desc), _native._stream())
        added = torch.full((K,), -1, dtype=torch.int32, device=dev)
        _native.commit(ws, g2, f, desc, rep_g, rep_f, new.descriptors, first_wins=first, added_cells=added)          # :193-208
        _native.call("qdx_scatter_rows_by_source", _native._ptr(added), _native._ptr(spread), C.c_int64(K), C.c_int64(1),
                     _native._ptr(new.spreads), _native._stream())                                                    # :209-211
        if extras:                                                                                                    # :214-220
            cells_changed = torch.nonzero(added >= 0).reshape(-1)
            src = added[cells_changed].long()
            new.extra_scores = {k: _scatter_rows(new.extra_scores[k], cells_changed, v, src) for k, v in extras.items()}
        if _native.DEBUG_SYNC:
            ws.check()
        return new

    @classmethod
    def init_default(cls, genotype, centroids, one_extra_score=None, keys_extra_scores: Tuple[str, ...] = (),
                     tie_break: str = "first") -> "MELSRepertoire":
        """reference :232-297: fitness -inf, genotypes 0, descriptors 0, spreads +inf (any spread is smaller)."""
        base = MapElitesRepertoire.init_default(genotype, centroids, one_extra_score, keys_extra_scores, tie_break)
        K = base.centroids.shape[0]
        return cls(base.genotypes, base.fitnesses, base.descriptors, base.centroids,
                   torch.full((K,), float("inf"), dtype=torch.float32, device=base.centroids.device),
                   base.extra_scores, keys_extra_scores, tie_break)
