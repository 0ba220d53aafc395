"""The C-ABI library builds, loads and exports every symbol include/drag_b200.h declares.
No compute calls here (no GPU in the authoring container)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from dial_rag_b200 import _native

    if not os.path.exists(_native.LIB_PATH):
        import importlib.util

        spec = importlib.util.spec_from_file_location(
            "drag_build", os.path.join(ROOT, "ai-dial-rag_b200", "csrc", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return _native.LIB_PATH


def declared_functions():
    text = open(os.path.join(ROOT, "include", "drag_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drag_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in drag_b200.h but not exported"


def test_python_binding_covers_header(lib_path):
    from dial_rag_b200 import _native

    assert set(declared_functions()) == set(_native.EXPORTED_SYMBOLS)
    lib = _native.load()
    assert lib.drag_abi_version() == 2


def test_errors_surface_as_exceptions(lib_path):
    from dial_rag_b200 import _native

    lib = _native.load()
    # argument validation happens before any CUDA call
    rc = lib.drag_row_sqnorm(None, 7, 10, 384, None, None)
    assert rc == 1
    with pytest.raises(_native.DragError, match="bad dtype"):
        _native.check(rc)
    rc = lib.drag_topk(0, None, 0, 10, 384, None, None, 1, 5000, 3, 0, None, None, None, None, 0, None)
    assert rc == 1 and b"k" in lib.drag_last_error()


def test_library_is_sm100a_only(lib_path):
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np

    from dial_rag_b200._native import DragError
    from dial_rag_b200.retrievers.embeddings_index import DocIndex, EmbeddingsIndex
    from dial_rag_b200.records import RetrievalType

    idx = EmbeddingsIndex(RetrievalType.TEXT, [DocIndex(np.array([0]), np.ones((1, 3), dtype=np.float32))])
    with pytest.raises(DragError, match="no CPU path"):
        idx.find(np.ones(3))
