"""UniformSelector -- mirrors qdax/core/emitters/repertoire_selectors/uniform_selector.py:14-62.

select(): p = occupied / sum(occupied); key, subkey = split(key); choice(subkey, arange(K), (n,), p=p); gather
every field (select_with_replacement=False: the Gumbel top-k trick of jax.random.choice, qdx_select_indices_without_replacement).
On the GPU: an occupancy scan builds the occupied-cell list and the float32 running-sum segments
(qdx_select_prepare), qdx_select_indices draws the index stream, qdx_gather_rows gathers the rows."""

from __future__ import annotations

import torch

from qdax_b200 import _native
from qdax_b200.core.emitters.repertoire_selectors.selector import Selector, unfold_repertoire


class UniformSelector(Selector):
    def __init__(self, select_with_replacement: bool = True):
        self.select_with_replacement = select_with_replacement

    def select_indices(self, repertoire, key, num_samples: int) -> torch.Tensor:
        """The index stream of `select` (int32, device)."""
        rep = unfold_repertoire(repertoire)
        f = _native.require_cuda(rep.fitnesses, "fitnesses")
        if f.dim() == 2 and f.shape[1] != 1:
            raise NotImplementedError("multi-objective fitnesses are outside the MAP-Elites hot path")
        ws = rep._workspace()
        ws.raise_if_error()
        _native.ensure_selection(f.reshape(-1), ws)
        if not self.select_with_replacement:     # jax.random.choice(replace=False): Gumbel top-k (reference :19-20, :54)
            return _native.select_indices_without_replacement(f.reshape(-1), ws, key, num_samples)
        return _native.select_indices(ws, key, num_samples, f.device)

    def select(self, repertoire, key, num_samples: int):
        rep = unfold_repertoire(repertoire)
        idx = self.select_indices(rep, key, num_samples)
        return rep._gather(idx)
