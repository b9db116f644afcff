"""Emitter ABC and EmitterState -- mirrors qdax/core/emitters/emitter.py:9-137 (the plugin interface
MAPElites drives: init / emit / state_update / batch_size / use_all_data)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Optional, Tuple

from qdax_b200.core.containers.repertoire import PyTreeNode, Repertoire


class EmitterState(PyTreeNode):
    """State carried by stateful emitters (reference emitter.py:9-26)."""


class Emitter(ABC):
    def init(self, key, repertoire: Repertoire, genotypes, fitnesses, descriptors, extra_scores) -> Optional[EmitterState]:
        """reference emitter.py:30-50: stateless emitters return None."""
        return None

    @abstractmethod
    def emit(self, repertoire: Optional[Repertoire], emitter_state: Optional[EmitterState], key) -> Tuple[Any, dict]:
        """reference emitter.py:52-72."""

    def state_update(self, emitter_state: Optional[EmitterState], repertoire: Optional[Repertoire] = None, genotypes=None,
                     fitnesses=None, descriptors=None, extra_scores=None) -> Optional[EmitterState]:
        """reference emitter.py:74-107: identity by default."""
        return emitter_state

    @property
    @abstractmethod
    def batch_size(self) -> int:
        """reference emitter.py:109-116."""

    @property
    def use_all_data(self) -> bool:
        """reference emitter.py:118-137."""
        return False
