"""Shared test helpers: golden loaders and the mapping oracle <-> package types."""

from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from tests.synth import synth_index

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_json(name: str):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def synth_cases():
    """Yield (entry, data) for every seeded case in search_synth.json, checking that the
    generator still reproduces the arrays the reference was run on."""
    for entry in load_json("search_synth.json"):
        data = synth_index(**entry["spec"])
        dim = entry["spec"]["dim"]
        flat = np.concatenate([e.reshape(-1, dim) for _, e in data["docs"]])
        assert sha(flat) == entry["sha_matrix"], "synthetic generator drifted from the golden inputs"
        assert sha(data["queries"]) == entry["sha_queries"]
        yield entry, data


def unit_docs(order: str):
    doc1 = (np.array([0, 1], dtype=np.int64), np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dtype=np.float32))
    doc2 = (np.array([0], dtype=np.int64), np.array([[1.0, 0.0, 0.0]], dtype=np.float32))
    doc3 = (np.array([], dtype=np.int64), np.array([], dtype=np.float32))
    return {"123": [doc1, doc2, doc3], "321": [doc3, doc2, doc1], "empty": [], "3": [doc3]}[order]
