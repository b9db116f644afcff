"""Abstract repertoire -- mirrors qdax/core/containers/repertoire.py:16-57 of the reference.

The reference derives from flax.struct.PyTreeNode (immutable dataclass with `.replace`).  `PyTreeNode` below
gives the same value semantics on plain Python objects holding CUDA tensors."""

from __future__ import annotations

import copy
from abc import ABC, abstractmethod
from typing import Any, Optional


class PyTreeNode:
    """Minimal stand-in for flax.struct.PyTreeNode: `replace(**updates)` returns a shallow copy."""

    def replace(self, **updates: Any):
        new = copy.copy(self)
        for k, v in updates.items():
            if not hasattr(new, k):
                raise AttributeError(f"{type(self).__name__} has no field {k!r}")
            object.__setattr__(new, k, v)
        if "fitnesses" in updates and hasattr(new, "_ws"):
            object.__setattr__(new, "_ws", None)     # the device workspace describes one fitness array
        return new


class Repertoire(PyTreeNode, ABC):
    """Abstract class for any repertoire of genotypes (reference repertoire.py:16-57)."""

    @classmethod
    @abstractmethod
    def init(cls) -> "Repertoire":
        """Create a repertoire."""

    @abstractmethod
    def select(self, key, num_samples: int, selector: Optional[Any] = None) -> "Repertoire":
        """Selects individuals from the repertoire."""

    @abstractmethod
    def add(self) -> "Repertoire":
        """Implements the rule to add new genotypes to a repertoire."""
