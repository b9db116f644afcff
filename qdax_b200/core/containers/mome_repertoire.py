"""Multi-Objective MAP-Elites repertoire -- mirrors qdax/core/containers/mome_repertoire.py:35-433 of the reference for the
insertion rule: genotypes (K, L, D), fitnesses (K, L, C) with -inf rows for the empty slots of a
Pareto front, descriptors (K, L, Dd), centroids (K, Dd); `add` scans the batch in index order, every offspring updating the
front of its cell (_update_masked_pareto_front :72-209).  Executed by qdx_cells + qdx_mome_add (one warp per cell; offspring
of one cell sequentially, cells in parallel).

PARITY UNPINNED (no reference test fixes a value; jax is not installable here): oracle/qdax_containers_numpy.py is the literal
restatement, and its header explains the `float * bool` products of the source and the two readings of them; `ieee_literal`
selects the reading (default False = what the compiled lax.scan body computes)."""

from __future__ import annotations

import ctypes as C
import warnings
from typing import Any, Dict, Optional, Tuple

import torch

from qdax_b200 import _native
from qdax_b200.core.containers.mapelites_repertoire import MapElitesRepertoire


class MOMERepertoire(MapElitesRepertoire):
    def __init__(self, genotypes, fitnesses, descriptors, centroids, extra_scores: Optional[Dict[str, Any]] = None,
                 keys_extra_scores: Tuple[str, ...] = (), ieee_literal: bool = False):
        super().__init__(genotypes, fitnesses, descriptors, centroids, extra_scores, keys_extra_scores)
        self.ieee_literal = bool(ieee_literal)

    @property
    def repertoire_capacity(self) -> int:
        """reference :52-63."""
        return int(self.genotypes.shape[0] * self.genotypes.shape[1])

    def _clone_state(self) -> "MOMERepertoire":
        return self.replace(genotypes=self.genotypes.clone(), fitnesses=self.fitnesses.clone(), descriptors=self.descriptors.clone())

    def select(self, key, num_samples: int, selector=None):
        raise NotImplementedError("MOMEUniformSelector (reference mome_uniform_selector.py) is outside the accelerated hot path; "
                                  "only the insertion rule MOMERepertoire.add is (SURVEY.md 8f rank 3)")

    def add(self, batch_of_genotypes, batch_of_descriptors, batch_of_fitnesses, batch_of_extra_scores=None, *,
            _donate: bool = False) -> "MOMERepertoire":
        """reference :211-322."""
        if self.keys_extra_scores:
            raise NotImplementedError("extra scores in the MOME repertoire are outside the accelerated path (the reference warns that it does not store them)")
        g = _native.require_cuda(batch_of_genotypes, "batch_of_genotypes")
        d = _native.require_cuda(batch_of_descriptors, "batch_of_descriptors")
        f = _native.require_cuda(batch_of_fitnesses, "batch_of_fitnesses")
        B = g.shape[0]
        K, L, Cn = self.fitnesses.shape
        if f.dim() != 2 or f.shape != (B, Cn) or d.shape[0] != B:
            raise ValueError("MOMERepertoire.add expects fitnesses of shape (batch_size, num_criteria)")
        new = self if _donate else self._clone_state()
        g2 = g.reshape(B, -1)
        D = g2.shape[1]
        if new.genotypes.reshape(K, L, -1).shape[2] != D:
            raise ValueError("genotype dimension mismatch")
        cells = _native.cells(d, new.centroids, new._grid())                                   # :239-240
        _native.call("qdx_mome_add", _native._ptr(new.fitnesses), _native._ptr(new.genotypes), _native._ptr(new.descriptors), C.c_int64(K),
                     C.c_int32(L), C.c_int32(Cn), C.c_int64(D), C.c_int32(d.shape[1]), _native._ptr(cells), _native._ptr(f), _native._ptr(g2),
                     _native._ptr(d), C.c_int64(B), C.c_int32(int(new.ieee_literal)), _native._stream())
        return new

    @classmethod
    def init(cls, genotypes, fitnesses, descriptors, centroids, pareto_front_max_length: int, *args, extra_scores=None,
             keys_extra_scores: Tuple[str, ...] = (), ieee_literal: bool = False, **kwargs) -> "MOMERepertoire":
        """reference :324-416: fitness -inf, genotypes 0, descriptors 0, then add the first batch."""
        warnings.warn("This type of repertoire does not store the extra scores computed by the scoring function", stacklevel=2)
        centroids = _native.require_cuda(centroids, "centroids")
        K, L = centroids.shape[0], int(pareto_front_max_length)
        dev = centroids.device
        rep = cls(
            genotypes=torch.zeros((K, L) + tuple(genotypes.shape[1:]), dtype=torch.float32, device=dev),
            fitnesses=torch.full((K, L, fitnesses.shape[1]), float("-inf"), dtype=torch.float32, device=dev),
            descriptors=torch.zeros((K, L, descriptors.shape[1]), dtype=torch.float32, device=dev),
            centroids=centroids, extra_scores={}, keys_extra_scores=(), ieee_literal=ieee_literal,
        )
        return rep.add(genotypes, descriptors, fitnesses, extra_scores, _donate=True)
