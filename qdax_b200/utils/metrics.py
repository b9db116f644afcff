"""QD metrics -- mirrors qdax/utils/metrics.py:19-49 (CSVLogger) and :74-98 (default_qd_metrics)."""

from __future__ import annotations

import csv
from typing import Dict, List

import torch

from qdax_b200 import _native


class CSVLogger:
    """reference metrics.py:19-49."""

    def __init__(self, filename: str, header: List) -> None:
        self._filename = filename
        self._header = header
        with open(self._filename, "w") as file:
            csv.DictWriter(file, fieldnames=self._header).writeheader()

    def log(self, metrics: Dict[str, float]) -> None:
        with open(self._filename, "a") as file:
            csv.DictWriter(file, fieldnames=self._header).writerow(metrics)


def default_qd_metrics(repertoire, qd_offset: float) -> Dict[str, torch.Tensor]:
    """reference metrics.py:74-98: qd_score = sum of non-empty fitnesses + qd_offset * #filled,
    coverage = 100 * filled fraction, max_fitness.  One small reduction kernel; 0-d device tensors."""
    m = _native.metrics(_native.require_cuda(repertoire.fitnesses, "fitnesses").reshape(-1), float(qd_offset))
    return {"qd_score": m[0], "max_fitness": m[1], "coverage": m[2]}


def default_ga_metrics(repertoire) -> Dict[str, torch.Tensor]:
    """reference metrics.py:52-71."""
    return {"max_fitness": torch.max(repertoire.fitnesses, dim=0).values}
