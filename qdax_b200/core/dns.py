"""Dominated Novelty Search driver -- mirrors qdax/core/dns.py:21-253 of the reference (same ask / score / tell
skeleton and key chain as MAPElites, flat population + DominatedNoveltyRepertoire)."""

from __future__ import annotations

from typing import Any, Callable, Optional

from qdax_b200 import random as qrandom
from qdax_b200.core.containers.dns_repertoire import DominatedNoveltyRepertoire
from qdax_b200.core.emitters.emitter import Emitter, EmitterState


class DominatedNoveltySearch:
    def __init__(self, scoring_function: Optional[Callable], emitter: Emitter, metrics_function: Callable, population_size: int,
                 k: int, repertoire_init: Callable = DominatedNoveltyRepertoire.init) -> None:
        self._scoring_function = scoring_function
        self._emitter = emitter
        self._metrics_function = metrics_function
        self._population_size = population_size
        self._k = k
        self._repertoire_init = lambda g, f, d, _p, _k, extra=None: repertoire_init(g, f, d, population_size, k, extra)

    def init(self, genotypes, key):
        """reference :84-117."""
        if self._scoring_function is None:
            raise ValueError("Scoring function is not set.")
        ks = qrandom.split(key)
        key, subkey = ks[0], ks[1]
        fitnesses, descriptors, extra_scores = self._scoring_function(genotypes, subkey)
        return self.init_ask_tell(genotypes=genotypes, fitnesses=fitnesses, descriptors=descriptors, key=key, extra_scores=extra_scores)

    def init_ask_tell(self, genotypes, fitnesses, descriptors, key, extra_scores=None):
        """reference :119-149."""
        if extra_scores is None:
            extra_scores = {}
        repertoire = self._repertoire_init(genotypes, fitnesses, descriptors, self._population_size, self._k, extra_scores)
        ks = qrandom.split(key)
        key, subkey = ks[0], ks[1]
        emitter_state = self._emitter.init(key=subkey, repertoire=repertoire, genotypes=genotypes, fitnesses=fitnesses,
                                           descriptors=descriptors, extra_scores=extra_scores)
        return repertoire, emitter_state, self._metrics_function(repertoire)

    def update(self, repertoire: DominatedNoveltyRepertoire, emitter_state: Optional[EmitterState], key):
        """reference :151-186."""
        if self._scoring_function is None:
            raise ValueError("Scoring function is not set.")
        ks = qrandom.split(key)
        key, subkey = ks[0], ks[1]
        genotypes, extra_info = self.ask(repertoire, emitter_state, subkey)
        ks = qrandom.split(key)
        key, subkey = ks[0], ks[1]
        fitnesses, descriptors, extra_scores = self._scoring_function(genotypes, subkey)
        return self.tell(genotypes=genotypes, fitnesses=fitnesses, descriptors=descriptors, repertoire=repertoire,
                         emitter_state=emitter_state, extra_scores=extra_scores, extra_info=extra_info)

    def scan_update(self, carry, _: Any = None):
        """reference :188-208."""
        repertoire, emitter_state, key = carry
        ks = qrandom.split(key)
        key, subkey = ks[0], ks[1]
        repertoire, emitter_state, metrics = self.update(repertoire, emitter_state, subkey)
        return (repertoire, emitter_state, key), metrics

    def ask(self, repertoire, emitter_state, key):
        """reference :210-219."""
        ks = qrandom.split(key)
        key, subkey = ks[0], ks[1]
        return self._emitter.emit(repertoire, emitter_state, subkey)

    def tell(self, genotypes, fitnesses, descriptors, repertoire, emitter_state, extra_scores=None, extra_info=None):
        """reference :221-253."""
        if extra_scores is None:
            extra_scores = {}
        if extra_info is None:
            extra_info = {}
        repertoire = repertoire.add(genotypes, descriptors, fitnesses, extra_scores)
        emitter_state = self._emitter.state_update(emitter_state=emitter_state, repertoire=repertoire, genotypes=genotypes,
                                                   fitnesses=fitnesses, descriptors=descriptors,
                                                   extra_scores={**extra_scores, **extra_info})
        return repertoire, emitter_state, self._metrics_function(repertoire)
