"""Host-side handle of the CUDA encoder (``drag_encoder`` in include/drag_b200.h).

Stands where ``HuggingFaceBgeEmbeddings(...).client`` (a SentenceTransformer) stands in
the reference (aidial_rag/embeddings/embeddings.py:52-66): it owns the model weights on
the device and turns token ids into L2-normalised CLS embeddings.  Inputs are PACKED
(sequences back to back + ``cu_seqlens``), so there is no pad-to-longest waste and no
need for sentence-transformers' sort-by-length trick.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np

from dial_rag_b200 import _native
from dial_rag_b200._native import BertShape, DragError


@dataclass(frozen=True)
class EncoderShape:
    """bge-small-en ``config.json`` values."""

    vocab: int = 30522
    hidden: int = 384
    layers: int = 12
    heads: int = 12
    inter: int = 1536
    max_pos: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12


BGE_SMALL = EncoderShape()


def weight_order(shape: EncoderShape = BGE_SMALL) -> List[str]:
    """HF ``BertModel`` state-dict keys in the order drag_encoder_create expects."""
    names = [
        "embeddings.word_embeddings.weight",
        "embeddings.position_embeddings.weight",
        "embeddings.token_type_embeddings.weight",
        "embeddings.LayerNorm.weight",
        "embeddings.LayerNorm.bias",
    ]
    for i in range(shape.layers):
        p = f"encoder.layer.{i}."
        for lin in ("attention.self.query", "attention.self.key", "attention.self.value",
                    "attention.output.dense", "intermediate.dense", "output.dense"):
            names += [p + lin + ".weight", p + lin + ".bias"]
        names += [p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias",
                  p + "output.LayerNorm.weight", p + "output.LayerNorm.bias"]
    return names


def _as_f32(value) -> np.ndarray:
    if hasattr(value, "detach"):  # torch tensor
        value = value.detach().to("cpu").float().numpy()
    return np.ascontiguousarray(value, dtype=np.float32)


def load_model_dir(path: str) -> Dict[str, np.ndarray]:
    """Read ``model.safetensors`` (or ``pytorch_model.bin``) of an HF BERT checkpoint directory,
    as ``BGE_EMBEDDINGS_MODEL_PATH`` points to in the reference (embeddings.py:28-32)."""
    st = os.path.join(path, "model.safetensors")
    if os.path.exists(st):
        from safetensors.numpy import load_file

        raw = load_file(st)
    else:
        import torch

        raw = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
    out = {}
    for k, v in raw.items():
        k = k[5:] if k.startswith("bert.") else k
        out[k] = _as_f32(v)
    return out


class B200Encoder:
    def __init__(self, weights: Mapping[str, object], shape: EncoderShape = BGE_SMALL,
                 device: int = 0, max_tokens: int = 65536):
        self.lib = _native.load()
        self.shape = shape
        self.device = int(device)
        self.max_tokens = int(max_tokens)
        names = weight_order(shape)
        missing = [n for n in names if n not in weights]
        if missing:
            raise KeyError(f"weights are missing {len(missing)} tensors, e.g. {missing[:3]}")
        arrays = [_as_f32(weights[n]) for n in names]
        expect = {
            names[0]: (shape.vocab, shape.hidden),
            names[1]: (shape.max_pos, shape.hidden),
        }
        for n, a in zip(names, arrays):
            if n in expect and tuple(a.shape) != expect[n]:
                raise ValueError(f"{n}: expected shape {expect[n]}, got {a.shape}")
        ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        cshape = BertShape(shape.vocab, shape.hidden, shape.layers, shape.heads, shape.inter,
                           shape.max_pos, shape.type_vocab, shape.ln_eps)
        handle = C.c_void_p()
        _native.check(self.lib.drag_encoder_create(C.byref(cshape), ptrs, len(arrays), self.device,
                                                   self.max_tokens, C.byref(handle)))
        self._handle = handle

    def close(self) -> None:
        h, self._handle = self._handle, None
        if h:
            self.lib.drag_encoder_destroy(h)

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ------------------------------------------------------------------ packing
    def pack(self, token_lists: Sequence[Sequence[int]]):
        """Truncate to ``max_pos`` (keeping the final [SEP], like the HF tokenizer's
        truncation) and pack to ``(ids int32[T], cu_seqlens int32[n+1])``."""
        mp = self.shape.max_pos
        lens = np.fromiter((min(len(t), mp) for t in token_lists), dtype=np.int64, count=len(token_lists))
        if len(lens) and lens.min() < 1:
            raise ValueError("every sequence needs at least one token ([CLS])")
        cu = np.zeros(len(lens) + 1, dtype=np.int32)
        np.cumsum(lens, out=cu[1:])
        ids = np.empty(int(cu[-1]), dtype=np.int32)
        for i, t in enumerate(token_lists):
            if len(t) > mp:
                ids[cu[i]:cu[i + 1] - 1] = t[: mp - 1]
                ids[cu[i + 1] - 1] = t[-1]
            else:
                ids[cu[i]:cu[i + 1]] = t
        return ids, cu

    # ------------------------------------------------------------------ host API
    def embed_packed(self, ids: np.ndarray, cu_seqlens: np.ndarray) -> np.ndarray:
        """Host ids in, host float32 ``[n, hidden]`` out (drag_encoder_embed_host); splits the
        batch when it exceeds ``max_tokens``."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        n = len(cu) - 1
        out = np.empty((n, self.shape.hidden), dtype=np.float32)
        start = 0
        while start < n:
            # largest end with cu[end] - cu[start] <= max_tokens
            end = int(np.searchsorted(cu, cu[start] + self.max_tokens, side="right")) - 1
            end = min(max(end, start + 1), n)
            sub_cu = np.ascontiguousarray(cu[start:end + 1] - cu[start], dtype=np.int32)
            sub_ids = ids[cu[start]:cu[end]]
            _native.check(self.lib.drag_encoder_embed_host(
                self._handle, sub_ids.ctypes.data, sub_cu.ctypes.data, end - start, out[start:end].ctypes.data))
            start = end
        return out

    def embed_token_lists(self, token_lists: Sequence[Sequence[int]]) -> np.ndarray:
        if len(token_lists) == 0:
            return np.zeros((0, self.shape.hidden), dtype=np.float32)
        return self.embed_packed(*self.pack(token_lists))

    # ------------------------------------------------------------------ device API
    def forward_device(self, d_ids, d_cu, h_cu: np.ndarray, d_out, stream: Optional[int] = None) -> None:
        """Device tensors in/out (torch cuda tensors), asynchronous on ``stream``."""
        import torch

        h_cu = np.ascontiguousarray(h_cu, dtype=np.int32)
        if stream is None:
            stream = torch.cuda.current_stream(d_ids.device).cuda_stream
        _native.check(self.lib.drag_encoder_forward(self._handle, d_ids.data_ptr(), d_cu.data_ptr(),
                                                    h_cu.ctypes.data, len(h_cu) - 1, d_out.data_ptr(), stream))

    def debug_hidden(self, ids: np.ndarray, cu_seqlens: np.ndarray, layer: int) -> np.ndarray:
        """fp32 copy of the hidden states after ``layer`` (0 = embedding LayerNorm)."""
        import torch

        dev = torch.device("cuda", self.device)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        d_ids = torch.from_numpy(ids).to(dev)
        d_cu = torch.from_numpy(cu).to(dev)
        hidden = torch.empty((len(ids), self.shape.hidden), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _native.check(self.lib.drag_encoder_forward_debug(self._handle, d_ids.data_ptr(), d_cu.data_ptr(),
                                                          cu.ctypes.data, len(cu) - 1, int(layer),
                                                          hidden.data_ptr(), stream))
        torch.cuda.synchronize(dev)
        return hidden.cpu().numpy()

    # ------------------------------------------------------------------ profiling (bench.py)
    KERNEL_CLASSES = ("embed_ln", "gemm_qkv", "attention", "gemm_out_ln", "gemm_up_gelu", "gemm_down_ln", "pool_normalize",
                      "cls_tail", "gemm_mlp_fused")

    def profile_begin(self, max_launches: int = 8192) -> None:
        _native.check(self.lib.drag_encoder_profile_begin(self._handle, int(max_launches)))

    def profile_end(self) -> Dict[str, Dict[str, float]]:
        ms = (C.c_double * len(self.KERNEL_CLASSES))()
        n = (C.c_int * len(self.KERNEL_CLASSES))()
        _native.check(self.lib.drag_encoder_profile_end(self._handle, ms, n))
        return {k: {"ms": ms[i], "launches": n[i]} for i, k in enumerate(self.KERNEL_CLASSES)}
