"""GARepertoire -- mirrors qdax/core/containers/ga_repertoire.py:16-180 (fields, size, select,
filter_extra_scores).  Its own `add` (sort-and-truncate) is not on the MAP-Elites path; it is provided with
torch library ops for completeness of the base class only."""

from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch

from qdax_b200 import _native
from qdax_b200 import tree_util
from qdax_b200.core.emitters.repertoire_selectors.selector import Selector
from qdax_b200.core.containers.repertoire import Repertoire


class GARepertoire(Repertoire):
    def __init__(self, genotypes, fitnesses, extra_scores: Optional[Dict[str, Any]] = None,
                 keys_extra_scores: Tuple[str, ...] = ()):
        self.genotypes = genotypes
        self.fitnesses = fitnesses
        self.extra_scores = {} if extra_scores is None else extra_scores
        self.keys_extra_scores = tuple(keys_extra_scores)
        self._ws: Optional[_native.Workspace] = None

    # ---- native plumbing shared by subclasses -------------------------------------------------------
    def _workspace(self) -> _native.Workspace:
        if self._ws is None or self._ws.K != self.fitnesses.shape[0]:
            self._ws = _native.Workspace(int(self.fitnesses.shape[0]), self.fitnesses.device)
        return self._ws

    def _raise_if_error(self) -> None:
        """A device-side error reported by an earlier kernel of this repertoire's workspace (QDX_ERR_*): raised here, at
        the next API boundary, from the host mirror of the flag -- never blocks (QDX_DEBUG_SYNC=1: synchronises and checks)."""
        if self._ws is not None:
            self._ws.raise_if_error()

    def _tensor_fields(self):
        return {k: v for k, v in vars(self).items() if isinstance(v, torch.Tensor) and not k.startswith("_")}

    def _packed_genotypes(self):
        """(genotypes as (N, D_total) rows, TreeSpec or None).  A pytree genotype lives in one packed buffer whose
        leaves `self.genotypes` exposes as views (qdax_b200/tree_util.py), so this is free."""
        g = self.genotypes
        if tree_util.is_tree(g):
            return tree_util.pack(g)
        return g.reshape(g.shape[0], -1), None

    def _gather(self, idx: torch.Tensor):
        """x[idx] for every array field (reference uniform_selector.py:57-60)."""
        updates = {name: _native.gather_rows(_native.require_cuda(t, name), idx) for name, t in self._tensor_fields().items()}
        if tree_util.is_tree(self.genotypes):
            flat, spec = self._packed_genotypes()
            updates["genotypes"] = tree_util.unpack(_native.gather_rows(flat, idx), spec)
        updates["extra_scores"] = {k: v[idx.long()] for k, v in self.extra_scores.items()}
        new = self.replace(**updates)
        new._ws = None
        return new

    # ---- reference surface ----------------------------------------------------------------------------
    @property
    def size(self) -> int:
        return int(self.fitnesses.shape[0])

    def select(self, key, num_samples: int, selector: Optional[Selector] = None):
        if selector is None:
            from ..emitters.repertoire_selectors.uniform_selector import UniformSelector

            selector = UniformSelector(select_with_replacement=True)
        return selector.select(self, key, num_samples)

    def filter_extra_scores(self, extra_scores: Dict[str, Any]) -> Dict[str, Any]:
        return {k: v for k, v in extra_scores.items() if k in self.keys_extra_scores}

    def add(self, batch_of_genotypes, batch_of_fitnesses, batch_of_extra_scores=None):
        """ga_repertoire.py:66-117: keep the `size` fittest of parents + offspring (not the MAP-Elites path)."""
        cand = torch.cat([self.genotypes, batch_of_genotypes], dim=0)
        cf = torch.cat([self.fitnesses, batch_of_fitnesses.reshape(batch_of_fitnesses.shape[0], -1)], dim=0)
        order = torch.flip(torch.argsort(cf.sum(dim=1), stable=True), dims=[0])[: self.size]
        return self.replace(genotypes=cand[order], fitnesses=cf[order], extra_scores={})

    @classmethod
    def init(cls, genotypes, fitnesses, population_size: int, *args, extra_scores=None, keys_extra_scores=(), **kwargs):
        f = fitnesses.reshape(fitnesses.shape[0], -1)
        rep = cls(
            genotypes=torch.zeros((population_size,) + tuple(genotypes.shape[1:]), dtype=genotypes.dtype, device=genotypes.device),
            fitnesses=torch.full((population_size, f.shape[1]), float("-inf"), dtype=torch.float32, device=genotypes.device),
            extra_scores={}, keys_extra_scores=keys_extra_scores,
        )
        return rep.add(genotypes, f, extra_scores)
