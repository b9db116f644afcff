"""Dominated Novelty Search repertoire -- mirrors qdax/core/containers/dns_repertoire.py:79-273 of the reference.
`add` = concat population + batch, dominated novelty (k nearest fitter neighbours), survivors =
argsort(meta_fitness)[::-1][:size]; executed by qdx_dns_add without materialising any (N, N) array."""

from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch

from qdax_b200 import _native
from qdax_b200.core.containers.ga_repertoire import GARepertoire


class DominatedNoveltyRepertoire(GARepertoire):
    def __init__(self, genotypes, fitnesses, descriptors, k: int, extra_scores: Optional[Dict[str, Any]] = None,
                 keys_extra_scores: Tuple[str, ...] = ()):
        super().__init__(genotypes, fitnesses, extra_scores, keys_extra_scores)
        self.descriptors = descriptors
        self.k = int(k)

    def add(self, batch_of_genotypes, batch_of_descriptors, batch_of_fitnesses, batch_of_extra_scores=None) -> "DominatedNoveltyRepertoire":
        """reference :94-165."""
        if self.keys_extra_scores:
            raise NotImplementedError("extra scores in the DNS repertoire are outside the accelerated path")
        g = _native.require_cuda(batch_of_genotypes, "batch_of_genotypes")
        d = _native.require_cuda(batch_of_descriptors, "batch_of_descriptors")
        f = _native.require_cuda(batch_of_fitnesses, "batch_of_fitnesses").reshape(-1)
        P = self.genotypes.shape[0]
        og, of, od, meta, surv = _native.dns_add(self.genotypes.reshape(P, -1), self.fitnesses.reshape(-1), self.descriptors,
                                                 g.reshape(g.shape[0], -1), f, d, self.k)
        new = self.replace(genotypes=og.reshape(self.genotypes.shape), fitnesses=of.reshape(P, 1), descriptors=od)
        new._last_meta_fitness, new._last_survivors = meta, surv
        return new

    @classmethod
    def init(cls, genotypes, fitnesses, descriptors, population_size: int, k: int, *args, extra_scores=None,
             keys_extra_scores: Tuple[str, ...] = (), **kwargs) -> "DominatedNoveltyRepertoire":
        """reference :167-212."""
        rep = cls.init_default(genotype=genotypes[0], descriptor_dim=descriptors.shape[-1], population_size=population_size,
                               keys_extra_scores=keys_extra_scores, k=k, device=genotypes.device)
        return rep.add(genotypes, descriptors, fitnesses, extra_scores)

    @classmethod
    def init_default(cls, genotype, descriptor_dim: int, population_size: int, one_extra_score=None,
                     keys_extra_scores: Tuple[str, ...] = (), k: int = 15, device=None) -> "DominatedNoveltyRepertoire":
        """reference :214-273: fitness -inf, genotypes 0, descriptors NaN."""
        dev = genotype.device if device is None else device
        return cls(
            genotypes=torch.zeros((population_size,) + tuple(genotype.shape), dtype=torch.float32, device=dev),
            fitnesses=torch.full((population_size, 1), float("-inf"), dtype=torch.float32, device=dev),
            descriptors=torch.full((population_size, descriptor_dim), float("nan"), dtype=torch.float32, device=dev),
            k=k, extra_scores={}, keys_extra_scores=keys_extra_scores,
        )
