"""MAP-Elites Low-Spread driver -- mirrors qdax/core/mels.py:23-60 of the reference: MAPElites whose scoring function is wrapped
so that every solution is evaluated `num_samples` times, on a MELSRepertoire (qdax_b200/core/containers/mels_repertoire.py).
All methods are inherited; a MELSRepertoire takes the generic path of MAPElites (emit / score / add called one by one, each of
them native)."""

from __future__ import annotations

from functools import partial
from typing import Callable

from qdax_b200.core.containers.mels_repertoire import MELSRepertoire
from qdax_b200.core.emitters.emitter import Emitter
from qdax_b200.core.map_elites import MAPElites
from qdax_b200.utils.sampling import multi_sample_scoring_function


class MELS(MAPElites):
    def __init__(self, scoring_function: Callable, emitter: Emitter, metrics_function: Callable, num_samples: int,
                 repertoire_init: Callable = MELSRepertoire.init) -> None:
        """reference :33-60."""
        super().__init__(partial(multi_sample_scoring_function, scoring_fn=scoring_function, num_samples=num_samples), emitter,
                         metrics_function, repertoire_init)
        self._num_samples = num_samples
