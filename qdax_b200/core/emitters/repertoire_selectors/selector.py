"""Selector ABC and unfold_repertoire -- mirrors qdax/core/emitters/repertoire_selectors/selector.py:21-57."""

from __future__ import annotations

import abc
from typing import Generic, TypeVar

import numpy as np
import torch

GARepertoireT = TypeVar("GARepertoireT")
MapElitesRepertoireT = TypeVar("MapElitesRepertoireT")


def unfold_repertoire(repertoire):
    """Flatten the base dimensions of a repertoire (reference selector.py:21-34).  Repertoires on the
    MAP-Elites / DNS path already have a single base dimension, in which case this is the identity."""
    base_shape = tuple(repertoire.fitnesses.shape[:-1])
    if len(base_shape) == 1:
        return repertoire
    size = int(np.prod(base_shape))
    updates = {}
    for name, value in vars(repertoire).items():
        if isinstance(value, torch.Tensor) and tuple(value.shape[: len(base_shape)]) == base_shape:
            updates[name] = value.reshape((size,) + tuple(value.shape[len(base_shape):]))
    return repertoire.replace(**updates)


class Selector(abc.ABC, Generic[GARepertoireT]):
    """A selector is an object that selects the individuals from a population."""

    @abc.abstractmethod
    def select(self, repertoire, key, num_samples: int):
        """Selects individuals from the repertoire (reference selector.py:40-57)."""
