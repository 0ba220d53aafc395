"""Distance metrics of the embeddings index, evaluated on the GPU.

Mirror of aidial_rag/retrievers/embeddings_metrics.py: the same ``Metric`` enum and
an ``ENUM_TO_METRIC`` table of ``f(query, docs) -> float64 distances`` callables
("smaller is better").  Each callable uploads ``docs`` and runs ``drag_distances``;
``EmbeddingsIndex`` does not go through these (it keeps the matrix resident and
uses the fused ``drag_topk``), they exist so that callers of the reference's table
(e.g. its tests/test_embeddings_metrics.py) keep working.
"""

from __future__ import annotations

from enum import StrEnum
from typing import Callable, Dict

import numpy as np


class Metric(StrEnum):  # embeddings_metrics.py:7-11
    COSINE_SIM = "cosine_sim"
    EUCLIDEAN_DIST = "euclidean_dist"
    SQEUCLIDEAN_DIST = "sqeuclidean_dist"
    INNER_PRODUCT = "inner_product"


def _metric(metric: Metric) -> Callable[[np.ndarray, np.ndarray], np.ndarray]:
    def distances(query: np.ndarray, docs: np.ndarray) -> np.ndarray:
        from dial_rag_b200.device_index import DeviceMatrix

        docs = np.asarray(docs)
        if docs.ndim != 2:
            raise ValueError(f"docs must be 2-d, got shape {docs.shape}")
        return DeviceMatrix(docs).distances(np.asarray(query), metric)

    distances.__name__ = f"_metric_for_{metric.value}"
    return distances


ENUM_TO_METRIC: Dict[Metric, Callable[[np.ndarray, np.ndarray], np.ndarray]] = {
    m: _metric(m) for m in Metric
}

assert len(ENUM_TO_METRIC) == len(Metric)
