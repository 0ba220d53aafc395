"""Headline benchmark: chunks/s embedded (bge-small-en shape, 256-token chunks) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.

Workload (BASELINE.json configs[1]): synthetic 256-token chunks (ids uniform in [1000, 30522),
[CLS] first, [SEP] last, no padding), seeded random-init weights of the bge-small-en
architecture.  A step = one packed forward of `--chunks` chunks per GPU (default 1024).
`value` times the device-resident call (token ids already in HBM); `e2e` times the C-ABI
host-buffer call (drag_encoder_embed_host: pinned staging, H2D ids, forward, D2H embeddings).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

SEQ_LEN = 256
HIDDEN = 384
FLOP_PER_CHUNK = 12 * (2 * 384 * (3 * 384 + 384 + 2 * 1536)) * SEQ_LEN + 12 * 4 * SEQ_LEN * SEQ_LEN * 384  # 12.080e9
METRIC = "chunks/s embedded (bge-small-en, 256 tok); queries/s top-100 on 10M×384 index"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"],
                "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 100 ms while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self._nvml = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self._nvml.nvmlDeviceGetClockInfo(self._h, self._nvml.NVML_CLOCK_SM))
                mask = self._nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms are meant to use every core the process may
    run on (the N>=2 reference arm ran single-threaded)."""
    import torch

    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def synth_batches(n_batches: int, chunks: int, rank: int):
    from tests.synth import synth_token_batch

    out = []
    for b in range(n_batches):
        ids, cu = synth_token_batch(seed=1000 + 97 * rank + b, n_seq=chunks, seq_len=SEQ_LEN)
        out.append(ids)
    return out, cu


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU implementation of the path (fp32 torch BERT, oracle port)
# ------------------------------------------------------------------------------------------
def cpu_encoder(weights):
    """Returns (embed(token_lists)->np.ndarray, description)."""
    import torch

    from oracle import encoder as oenc

    try:
        model = oenc.hf_bert_model(weights)

        @torch.no_grad()
        def embed(token_lists):
            # sentence-transformers batching: minibatch 32 (all sequences have equal length here)
            out = []
            for s in range(0, len(token_lists), 32):
                ids = torch.tensor(token_lists[s:s + 32], dtype=torch.int64)
                hidden = model(input_ids=ids, attention_mask=torch.ones_like(ids)).last_hidden_state
                out.append(oenc.pool_and_normalize(hidden).numpy())
            return np.concatenate(out)

        return embed, "transformers.BertModel fp32 (reference backend='torch' graph), minibatch 32"
    except Exception:  # noqa: BLE001
        return (lambda tl: oenc.encode_token_lists(weights, tl)), "oracle/encoder.py fp32 restatement, minibatch 32"


def time_cpu_baseline(weights, budget_s: float = 15.0, first: int = 32):
    import torch

    from oracle import encoder as oenc
    from tests.synth import synth_token_batch

    use_all_host_threads()
    embed, what = cpu_encoder(weights)
    ids, cu = synth_token_batch(seed=77, n_seq=first, seq_len=SEQ_LEN)
    lists = oenc.packed_to_lists(ids, cu)
    embed(lists[:8])  # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    embed(lists)
    dt = time.perf_counter() - t0
    done, total = first, dt
    reps = int(max(0, min(16, (budget_s - dt) // max(dt, 1e-3))))
    for _ in range(reps):
        t0 = time.perf_counter()
        embed(lists)
        total += time.perf_counter() - t0
        done += first
    return {"value": done / total, "unit": "chunks/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{done} chunks x {SEQ_LEN} tokens in {total:.1f} s; {what}"}


def run_reference(args, rank: int):
    if rank != 0:
        return
    import torch

    from oracle import encoder as oenc
    from tests.synth import synth_token_batch

    use_all_host_threads()
    weights = oenc.synth_weights(seed=0, style="hf_init")
    embed, what = cpu_encoder(weights)
    sample = args.ref_chunks
    ids, cu = synth_token_batch(seed=1000, n_seq=sample, seq_len=SEQ_LEN)
    lists = oenc.packed_to_lists(ids, cu)
    for _ in range(max(args.warmup, 0)):
        embed(lists[: min(8, sample)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        embed(lists)
    dt = time.perf_counter() - t0
    value = args.steps * sample / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: synthetic {SEQ_LEN}-token chunks, bge-small-en architecture, seeded random weights",
                   "chunks_per_step": sample, "seq_len": SEQ_LEN,
                   "note": "each step is a bounded sample of the workload on the host cores"},
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {sample} chunks x {SEQ_LEN} tokens; {what}"},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# secondary metric: exact top-k search (single GPU legs)
# ------------------------------------------------------------------------------------------
def search_roofline(peaks, n_queries: int, rows: int, ms: float, elem_bytes: int, batched: bool):
    """(roofline dict, path description, matrix passes) of one search batch on one GPU."""
    if batched:
        # tensor-core candidate pass + float64 re-rank: one read of the bf16 scoring copy per group of
        # query tiles; bound = max(bytes / HBM peak, 2*Q*N*D flop / bf16 peak)
        n_qt, sms = -(-n_queries // 128), 148
        groups = next(g for g in range(1, n_qt + 1)   # mirrors pick_group() in drag_topk.cu
                      if (sms // -(-n_qt // g)) * -(-n_qt // g) * 100 >= sms * 95 or -(-n_qt // g) == 1)
        flops = 2.0 * n_queries * rows * HIDDEN
        t_mma = flops / (peaks["tflops_burst"] * 1e12)
        t_hbm = rows * HIDDEN * 2 / (peaks["hbm_gbs"] * 1e9)
        bound = "tensor" if t_mma >= t_hbm else "hbm"
        roof = {"bound": bound,
                "achieved": flops / (ms / 1e3) / 1e12 if bound == "tensor" else rows * HIDDEN * 2 / (ms / 1e3) / 1e9,
                "peak": peaks["tflops_burst"] if bound == "tensor" else peaks["hbm_gbs"],
                "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                "frac": max(t_mma, t_hbm) / (ms / 1e3),
                "t_hbm_ms": t_hbm * 1e3, "t_mma_ms": t_mma * 1e3,
                "note": "algorithmic work = ONE pass over the matrix per query batch: 2*Q*N*D flop on the bf16 scoring "
                        "copy (N*D*2 bytes); the binding one of the two is reported; peak = measured bf16 burst / HBM copy"}
        return roof, "tcgen05 bf16 candidate scores under a certified error bound + float64 re-rank (drag_topk_batch)", groups
    passes = -(-n_queries // 4) if n_queries >= 4 else 1
    bytes_per_pass = rows * HIDDEN * elem_bytes
    roof = {"bound": "hbm", "achieved": passes * bytes_per_pass / (ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": passes * bytes_per_pass / (ms / 1e3) / 1e9 / peaks["hbm_gbs"],
            # ncu --set full of scan_ring_kernel: dram__bytes_read.sum = 1.0016 x N*D*4 (profiles/r01_f_scan_ring_ncu.txt)
            "traffic": 1.0016 * passes * bytes_per_pass,
            "note": "algorithmic bytes = one read of the matrix per pass of <=4 queries; the time includes the query "
                    "preparation and list-merge kernels of the call"}
    return roof, "float64 scan (drag_topk)", passes


def bench_search(torch, device, peaks, rows: int, n_queries: int, k: int, steps: int, warmup: int, seed: int,
                 library_baseline: bool = False):
    from dial_rag_b200.device_index import DeviceMatrix

    g = torch.Generator(device=device).manual_seed(seed)
    mat = torch.empty((rows, HIDDEN), dtype=torch.float32, device=device)
    step_rows = 1 << 20
    for s in range(0, rows, step_rows):
        blk = torch.randn((min(step_rows, rows - s), HIDDEN), generator=g, device=device)
        mat[s:s + step_rows] = blk / blk.norm(dim=1, keepdim=True)
    dm = DeviceMatrix(mat)
    q = torch.randn((n_queries, HIDDEN), generator=g, device=device, dtype=torch.float32)
    q = (q / q.norm(dim=1, keepdim=True)).double()
    q_host = q.cpu().numpy()
    for _ in range(warmup):
        dm.topk_device(q, k, "inner_product")
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        dm.topk_device(q, k, "inner_product")
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / steps
    # end to end through the host API (query H2D, ids/distances D2H)
    t0 = time.perf_counter()
    for _ in range(steps):
        dm.topk(q_host, k, "inner_product")
    e2e_ms = 1e3 * (time.perf_counter() - t0) / steps
    batched = dm._use_batch(n_queries, k, 3)
    roof, path, passes = search_roofline(peaks, n_queries, rows, ms, 4, batched)
    lib = None
    if library_baseline:
        try:
            lib, lib_idx = gpu_library_search(torch, device, mat, q, k)
            ours = dm.topk_device(q, k, "inner_product")[1]
            # how often the library's fp32 ranking reproduces the exact (float64, lowest-id ties) one
            lib["ids_equal_to_exact_frac"] = float((lib_idx == ours).all(dim=1).float().mean().item())
            lib["recall_at_k_vs_exact"] = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / k
                                                        for a, b in zip(lib_idx[:64].cpu().numpy(), ours[:64].cpu().numpy())]))
        except Exception as exc:  # noqa: BLE001
            lib = {"error": repr(exc)}
    out = {
        "workload": f"exact top-{k} inner product, {rows}x{HIDDEN} fp32 index resident in HBM, batch {n_queries} queries; {path}",
        "queries_per_s": n_queries / (ms / 1e3), "ms_per_batch": ms,
        "e2e_queries_per_s": n_queries / (e2e_ms / 1e3),
        "matrix_passes_per_batch": passes,
        "fallback_queries": int(dm.last_batch_fallbacks),
        "roofline": roof,
    }
    if lib is not None:
        out["gpu_library_baseline"] = lib
    del dm, mat
    torch.cuda.empty_cache()
    return out


def bench_query_path(torch, device, weights, rows: int = 1_000_000, q_len: int = 512, k: int = 20):
    """configs[4]: 512-token query embedding + exact top-20 over a 1M-chunk fp32 index.
    Latency at batch 1 (p50 / p90 of host wall time per query, call to result on the host, and the
    device time by CUDA events) and throughput at batch 256."""
    from dial_rag_b200.device_index import DeviceMatrix
    from dial_rag_b200.embeddings.encoder import B200Encoder

    g = torch.Generator(device=device).manual_seed(5)
    mat = torch.randn((rows, HIDDEN), generator=g, device=device)
    mat /= mat.norm(dim=1, keepdim=True)
    dm = DeviceMatrix(mat)
    enc = B200Encoder(weights, device=device.index, max_tokens=256 * q_len)
    rng = np.random.default_rng(5)
    out = {"workload": f"configs[4]: {q_len}-token query -> embedding -> exact top-{k} (squared euclidean, the product default) "
                       f"over a {rows}x{HIDDEN} fp32 index resident in HBM"}
    for batch, iters in ((1, 200), (256, 10)):
        ids = rng.integers(1000, 30522, size=(8, batch, q_len), dtype=np.int32)
        ids[:, :, 0], ids[:, :, -1] = 101, 102
        cu = np.arange(0, batch * q_len + 1, q_len, dtype=np.int32)
        d_ids = [torch.from_numpy(ids[i].reshape(-1)).to(device) for i in range(8)]
        d_cu = torch.from_numpy(cu).to(device)
        d_emb = torch.empty((batch, HIDDEN), dtype=torch.float32, device=device)

        def device_step(i):
            enc.forward_device(d_ids[i % 8], d_cu, cu, d_emb)
            return dm.topk_device(d_emb.double(), k, "sqeuclidean_dist")

        def host_step(i):
            emb = enc.embed_packed(ids[i % 8].reshape(-1), cu)          # host ids in, host embeddings out
            return dm.topk(emb.astype(np.float64), k, "sqeuclidean_dist")  # host query in, host (rows, distances) out

        for i in range(3):
            device_step(i)
            host_step(i)
        torch.cuda.synchronize(device)
        dev_ms, host_ms = [], []
        for i in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            device_step(i)
            e1.record()
            torch.cuda.synchronize(device)
            dev_ms.append(e0.elapsed_time(e1))
        for i in range(iters):
            t0 = time.perf_counter()
            host_step(i)
            host_ms.append(1e3 * (time.perf_counter() - t0))
        dev_ms.sort()
        host_ms.sort()
        key = "batch1" if batch == 1 else f"batch{batch}"
        out[key] = {
            "device_ms_p50": dev_ms[len(dev_ms) // 2], "device_ms_p90": dev_ms[int(len(dev_ms) * 0.9)],
            "host_api_ms_p50": host_ms[len(host_ms) // 2], "host_api_ms_p90": host_ms[int(len(host_ms) * 0.9)],
            "queries_per_s_device": batch / (dev_ms[len(dev_ms) // 2] / 1e3),
            "queries_per_s_host_api": batch / (host_ms[len(host_ms) // 2] / 1e3),
            "iterations": iters,
        }
    # ---- batch-1 query latency while an indexing thread keeps the bulk workspace busy (the reference's two
    # embedding pools, cpu_pools.py:50-59): the query runs in the encoder's query workspace on its high-priority stream
    import threading

    load_ids = rng.integers(1000, 30522, size=(256, 256), dtype=np.int32)
    load_ids[:, 0], load_ids[:, -1] = 101, 102
    load_cu = np.arange(0, 256 * 256 + 1, 256, dtype=np.int32)
    stop = threading.Event()
    forwards = [0]

    def indexing_load():
        while not stop.is_set():
            enc.embed_packed(load_ids.reshape(-1), load_cu)
            forwards[0] += 1

    q_ids = rng.integers(1000, 30522, size=(8, q_len), dtype=np.int32)
    q_ids[:, 0], q_ids[:, -1] = 101, 102
    q_cu = np.array([0, q_len], dtype=np.int32)
    worker = threading.Thread(target=indexing_load, daemon=True)
    worker.start()
    time.sleep(0.2)
    lat = []
    for i in range(300):
        t0 = time.perf_counter()
        emb = enc.embed_packed(q_ids[i % 8], q_cu)
        dm.topk(emb.astype(np.float64), k, "sqeuclidean_dist")
        lat.append(1e3 * (time.perf_counter() - t0))
    stop.set()
    worker.join(timeout=30)
    lat.sort()
    out["batch1_under_indexing_load"] = {
        "host_api_ms_p50": lat[len(lat) // 2], "host_api_ms_p90": lat[int(len(lat) * 0.9)], "host_api_ms_p99": lat[int(len(lat) * 0.99)],
        "indexing_forwards_during_measurement": forwards[0],
        "load": "a second host thread embeds 256 x 256-token chunks back to back through the same encoder (bulk workspace)",
    }
    enc.close()
    del dm, mat, enc
    torch.cuda.empty_cache()
    return out


def bench_text_path(torch, device, weights, n_texts: int = 1024, steps: int = 5):
    """raw text -> WordPiece (host, C++ fast path of the C-ABI library) -> packed ids -> CUDA encoder ->
    host float32 embeddings, through the langchain-shaped ``embed_documents_numpy`` the indexing path calls.
    Synthetic ASCII text (~256 tokens per chunk) over a synthetic 20k-word vocabulary."""
    import tempfile

    from dial_rag_b200.embeddings.embeddings import B200BgeEmbeddings
    from dial_rag_b200.embeddings.tokenizer import WordPieceTokenizer

    rng = np.random.default_rng(11)
    syll = ["ka", "to", "mi", "ra", "ne", "su", "lo", "vi", "an", "er", "st", "on", "qu", "ph", "gl", "ice", "alp", "ber"]
    words = sorted({"".join(rng.choice(syll, size=rng.integers(1, 4))) for _ in range(40000)})[:20000]
    vocab = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    vocab += list("abcdefghijklmnopqrstuvwxyz") + ["##" + c for c in "abcdefghijklmnopqrstuvwxyz"] + list(".,;:!?'-") + words
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "vocab.txt")
        with open(path, "w") as f:
            f.write("\n".join(dict.fromkeys(vocab)) + "\n")
        tok = WordPieceTokenizer.from_vocab_file(path)
        texts = []
        for _ in range(n_texts):
            ws = rng.choice(words, size=215)
            texts.append(" ".join(w.capitalize() + ("," if i % 11 == 10 else "." if i % 17 == 16 else "") for i, w in enumerate(ws)))
        ids, cu = tok.encode_packed(texts)
        n_tokens = int(cu[-1])
        t0 = time.perf_counter()
        for _ in range(steps):
            tok.encode_packed(texts)
        fast_s = (time.perf_counter() - t0) / steps
        t0 = time.perf_counter()
        tok.encode_batch(texts)
        ref_s = time.perf_counter() - t0
        impl = B200BgeEmbeddings(weights, tok, device=device.index, max_tokens=n_texts * 512)
        for _ in range(2):
            impl.embed_documents_numpy(texts)
        t0 = time.perf_counter()
        for _ in range(steps):
            out = impl.embed_documents_numpy(texts)
        e2e_s = (time.perf_counter() - t0) / steps
        # ... and through build_embeddings (the indexing entry point: batches of EMBEDDINGS_BATCH_SIZE, batch i+1
        # tokenised while batch i is on the GPU)
        import asyncio
        import io

        from dial_rag_b200.embeddings import embeddings as emb

        emb.configure(impl)
        many = texts * 8
        asyncio.run(emb.build_embeddings(texts, io.StringIO()))
        t0 = time.perf_counter()
        rows = list(asyncio.run(emb.build_embeddings(many, io.StringIO())))
        build_s = time.perf_counter() - t0
        assert len(rows) == len(many)
        emb.configure(None)
        impl.client.close()
    return {
        "workload": f"{n_texts} synthetic ASCII chunks, {n_tokens / n_texts:.0f} tokens each, text -> tokenizer -> encoder -> host embeddings",
        "tokens_per_s_wordpiece_native": n_tokens / fast_s, "tokens_per_s_wordpiece_reference": n_tokens / ref_s,
        "host_threads": os.cpu_count(),
        "chunks_per_s_text_to_embedding": n_texts / e2e_s,
        "chunks_per_s_build_embeddings": len(many) / build_s,
        "embedding_checksum": float(np.asarray(out, dtype=np.float64).sum()),
    }


def bench_index_build(torch, device, rows: int = 1_000_000, generic_rows: int = 100_000):
    """embeddings of `rows` chunks -> persisted MultiEmbeddings -> DocIndex -> device-resident index
    (upload, bit-exact row norms, bf16 scoring copy).  The matrix path is what build_embeddings produces
    (pack_embedding_matrix: zero-copy views, O(1) flatten); the generic path is the reference's per-item format
    (pack_simple_embeddings + the flatten loop, as for records deserialised from storage)."""
    from dial_rag_b200.records import RetrievalType
    from dial_rag_b200.retrievers.embeddings_index import (EmbeddingsIndex, create_index_by_chunk, pack_embedding_matrix,
                                                           pack_simple_embeddings)

    rng = np.random.default_rng(11)
    emb = rng.standard_normal((rows, HIDDEN), dtype=np.float32)
    t0 = time.perf_counter()
    multi = pack_embedding_matrix(emb)
    t1 = time.perf_counter()
    doc = create_index_by_chunk(multi)
    t2 = time.perf_counter()
    index = EmbeddingsIndex(RetrievalType.TEXT, [doc], limit=20)
    index.find(emb[0].astype(np.float64))          # first use uploads the matrix
    torch.cuda.synchronize(device)
    t3 = time.perf_counter()
    g0 = time.perf_counter()
    gdoc = create_index_by_chunk(pack_simple_embeddings(list(emb[:generic_rows])))
    g1 = time.perf_counter()
    assert np.array_equal(gdoc.embeddings, emb[:generic_rows])
    out = {
        "workload": f"{rows} x {HIDDEN} float32 chunk embeddings -> MultiEmbeddings -> DocIndex -> resident device index + first query",
        "pack_s": t1 - t0, "flatten_s": t2 - t1, "upload_and_first_query_s": t3 - t2,
        "chunks_per_s": rows / (t3 - t0),
        "generic_item_path_chunks_per_s": generic_rows / (g1 - g0),
        "note": "pack = one ItemEmbeddings view per chunk (host Python); flatten is O(1) for the matrix path",
    }
    del index, doc, multi, emb
    torch.cuda.empty_cache()
    return out


def bench_wide_dims(torch, device, peaks, rows: int = 250_000):
    """the other consumers of the index (description / multimodal retrievers) use D = 1024 / 1408,
    beyond the tensor-core batch path (D <= 512): they run on the float64 scan."""
    from dial_rag_b200.device_index import DeviceMatrix

    out = {}
    for dim in (1024, 1408):
        g = torch.Generator(device=device).manual_seed(dim)
        mat = torch.randn((rows, dim), generator=g, device=device)
        dm = DeviceMatrix(mat)
        res = {}
        for nq in (1, 16):
            q = torch.randn((nq, dim), generator=g, device=device, dtype=torch.float32).double()
            for _ in range(3):
                dm.topk_device(q, 20, "cosine_sim")
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                dm.topk_device(q, 20, "cosine_sim")
            e1.record()
            torch.cuda.synchronize(device)
            ms = e0.elapsed_time(e1) / 20
            passes = -(-nq // 4)
            res[f"q{nq}"] = {"ms": ms, "queries_per_s": nq / (ms / 1e3), "matrix_passes": passes,
                             "hbm_frac": passes * rows * dim * 4 / (ms / 1e3) / 1e9 / peaks["hbm_gbs"]}
        out[f"d{dim}"] = res
        del dm, mat
        torch.cuda.empty_cache()
    out["workload"] = f"exact top-20 cosine, {rows} rows fp32, float64 scan (drag_topk): 1 query and 16 queries per call"
    return out


def bench_search_sharded(torch, dist, device, rank, world, peaks, rows_per_gpu: int, n_queries: int, k: int, steps: int, warmup: int):
    """configs[3] shape: bf16 index row-sharded over the GPUs, replicated queries, one NCCL all-gather of
    the per-shard top-k candidates + drag_topk_merge on every rank.  Parity inside the run: the first
    ``4 * world`` queries are PLANTED -- query i is written (bf16) into row (7919 i) mod rows of shard i mod world, so
    it must come back first with exactly that global row id, from every shard in turn."""
    from dial_rag_b200.sharded import ShardedIndex

    g = torch.Generator(device=device).manual_seed(400 + rank)
    mat = torch.empty((rows_per_gpu, HIDDEN), dtype=torch.bfloat16, device=device)
    step_rows = 1 << 20
    for s in range(0, rows_per_gpu, step_rows):
        blk = torch.randn((min(step_rows, rows_per_gpu - s), HIDDEN), generator=g, device=device)
        mat[s:s + step_rows] = (blk / blk.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    gq = torch.Generator().manual_seed(4)
    q = torch.randn((n_queries, HIDDEN), generator=gq)
    q = (q / q.norm(dim=1, keepdim=True)).double()
    planted = min(4 * world, n_queries)
    expect = []
    for i in range(planted):
        shard, local = i % world, (7919 * i) % rows_per_gpu
        expect.append(shard * rows_per_gpu + local)
        if shard == rank:
            mat[local] = q[i].to(device).to(torch.bfloat16)
    q = q.numpy()
    idx = ShardedIndex(mat, row_start=rank * rows_per_gpu, storage="bf16", device=device.index)
    del mat
    for _ in range(warmup):
        out = idx.topk(q, k, "inner_product")
    planted_ok = bool(np.array_equal(out[1][:planted, 0], np.array(expect, dtype=np.int64)))
    sorted_ok = bool((np.diff(out[0], axis=1) >= 0).all())
    dist.barrier()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = idx.topk(q, k, "inner_product")
    torch.cuda.synchronize(device)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sec = float(dt[0]) / steps
    # second pass: where the time goes (CUDA events between the stages of the call; not part of the timed region)
    stages = {}
    for _ in range(3):
        idx.topk(q, k, "inner_product", timers=stages)
    stages = {name: ms / 3 for name, ms in stages.items()}
    ok = torch.tensor([int(planted_ok and sorted_ok)], dtype=torch.int32, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    roof, path, passes = search_roofline(peaks, n_queries, rows_per_gpu, sec * 1e3, 2,
                                         idx._matrix._use_batch(n_queries, k, 3))
    roof["unit"] += " per GPU"
    return {
        "workload": f"exact top-{k} inner product, {world * rows_per_gpu}x{HIDDEN} bf16 index row-sharded over {world} GPUs, "
                    f"batch {n_queries} replicated queries, NCCL all-gather candidate merge (host queries in, host results out); {path}",
        "queries_per_s": n_queries / sec, "ms_per_batch": sec * 1e3, "steps": steps,
        "allgather_bytes_per_rank": n_queries * (2 * k + 1) * 8,
        "matrix_passes_per_batch": passes,
        "stage_ms_rank0": stages,
        "roofline": roof,
        "parity": {"planted_rows_first_on_every_rank": bool(int(ok[0])), "planted_queries": planted,
                   "distances_sorted": sorted_ok},
        "fallback_queries": int(idx._matrix.last_batch_fallbacks),
    }


def cpu_search_baseline(rows: int = 1_000_000, n_queries: int = 4, k: int = 100):
    import torch

    from oracle import search as osearch
    from tests.synth import synth_matrix, synth_queries

    use_all_host_threads()
    m = synth_matrix(seed=2, rows=rows, dim=HIDDEN)
    q = synth_queries(seed=3, n=n_queries, dim=HIDDEN)
    want = []
    t0 = time.perf_counter()
    for i in range(n_queries):
        want.append(osearch.topk_rows("inner_product", k, q[i], m)[0])
    dt = time.perf_counter() - t0
    # parity at bench scale: the same queries through the CUDA batch path (tensor-core candidates + float64 re-rank)
    # and through the float64 scan, on the same rows: ids must be identical to the reference path's
    parity = {}
    if torch.cuda.is_available():
        from dial_rag_b200.device_index import DeviceMatrix

        dm = DeviceMatrix(m)
        _, got_batch, _ = dm.topk(q, k, "inner_product")
        dq = torch.from_numpy(q).to(dm.matrix.device)
        got_scan = dm.topk_device(dq, k, "inner_product", allow_batch=False)[1].cpu().numpy()
        parity = {"search_ids_equal_batch_path": bool(all(np.array_equal(got_batch[i], want[i]) for i in range(n_queries))),
                  "search_ids_equal_scan_path": bool(all(np.array_equal(got_scan[i], want[i]) for i in range(n_queries))),
                  "search_parity_case": f"{n_queries} queries, top-{k}, {rows}x{HIDDEN} fp32, ids vs the reference numpy path"}
        del dm
        torch.cuda.empty_cache()
    return {"parity": parity,
            "value": n_queries / dt, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_queries} queries, top-{k}, {rows}x{HIDDEN} fp32 (reference numpy path: float64 query, full stable argsort); "
                      f"linear-in-N extrapolation to 10M rows: {n_queries / dt / (10_000_000 / rows):.4f} q/s"}



# ------------------------------------------------------------------------------------------
# the generic GPU-library bar on the same B200: what the reference's own CUDA path would run
# ------------------------------------------------------------------------------------------
def gpu_library_encoder(torch, device, weights, host_ids, cu, chunks: int):
    """HF ``BertModel`` in fp16 with SDPA attention on the GPU -- the reference's CUDA configuration
    (aidial_rag/embeddings/embeddings.py:43-48: torch_dtype float16, attn_implementation sdpa) -- on the same
    256-token chunks, minibatch 32 (sentence-transformers' default) and 256."""
    from transformers import BertConfig, BertModel

    cfg = BertConfig(vocab_size=30522, hidden_size=384, num_hidden_layers=12, num_attention_heads=12, intermediate_size=1536,
                     max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12, hidden_act="gelu")
    try:
        cfg._attn_implementation = "sdpa"
        model = BertModel(cfg, add_pooling_layer=False)
        attn = "sdpa"
    except Exception:  # noqa: BLE001
        cfg._attn_implementation = "eager"
        model = BertModel(cfg, add_pooling_layer=False)
        attn = "eager"
    model.load_state_dict({k: v for k, v in weights.items()}, strict=False)
    model = model.half().to(device).eval()
    ids = torch.from_numpy(host_ids[0].astype(np.int64)).view(chunks, SEQ_LEN).to(device)
    out = {"what": f"transformers.BertModel fp16 + {attn} attention on this GPU (torch {torch.__version__}), CLS pooling + 2x normalize"}

    @torch.no_grad()
    def run(bs):
        embs = []
        for s0 in range(0, chunks, bs):
            x = ids[s0:s0 + bs]
            h = model(input_ids=x, attention_mask=torch.ones_like(x)).last_hidden_state[:, 0].float()
            h = torch.nn.functional.normalize(torch.nn.functional.normalize(h, dim=1), dim=1)
            embs.append(h)
        return torch.cat(embs)

    for bs in (32, 256):
        for _ in range(2):
            run(bs)
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            emb = run(bs)
        e1.record()
        torch.cuda.synchronize(device)
        out[f"chunks_per_s_bs{bs}"] = 3 * chunks / (e0.elapsed_time(e1) / 1e3)
    out["checksum"] = float(emb.double().sum().item())
    del model
    torch.cuda.empty_cache()
    return out, emb.cpu().numpy()


def gpu_library_search(torch, device, mat, q64, k: int):
    """``torch.matmul`` + ``torch.topk`` on the same matrix: fp32 scores in row blocks of 1M (a [Q, 10M] score
    matrix does not fit), running top-k merge.  Not exact in the reference's sense (fp32 accumulate, no tie rule)."""
    q = q64.float()
    rows = mat.shape[0]
    blk = 1 << 20

    def run():
        best_s = best_i = None
        for s0 in range(0, rows, blk):
            sc = q @ mat[s0:s0 + blk].T
            v, i = torch.topk(sc, min(k, sc.shape[1]), dim=1)
            i = i + s0
            if best_s is None:
                best_s, best_i = v, i
            else:
                cs, ci = torch.cat((best_s, v), 1), torch.cat((best_i, i), 1)
                best_s, sel = torch.topk(cs, k, dim=1)
                best_i = torch.gather(ci, 1, sel)
        return best_s, best_i

    run()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        _, idx = run()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / 2
    return {"queries_per_s": q.shape[0] / (ms / 1e3), "ms_per_batch": ms,
            "what": "torch.matmul (fp32) + torch.topk in 1M-row blocks with a running merge, same matrix and queries"}, idx


# ------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunks", type=int, default=1024, help="chunks per GPU per step")
    ap.add_argument("--ref-chunks", type=int, default=64, help="chunks per step of the CPU reference arm")
    ap.add_argument("--no-search", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the HF fp16+SDPA / torch.matmul+topk legs")
    ap.add_argument("--search-rows", type=int, default=10_000_000)
    ap.add_argument("--shard-rows", type=int, default=12_500_000, help="rows per GPU of the sharded bf16 index (N>1)")
    ap.add_argument("--shard-queries", type=int, default=4096)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from dial_rag_b200.embeddings.encoder import B200Encoder
    import synth_weights as synth  # neutral seeded-weights generator (the CPU arms draw the same tensors through the oracle)

    peaks = load_peaks()
    weights = synth.synth_weights(seed=0, style="hf_init")
    chunks = args.chunks
    tokens = chunks * SEQ_LEN
    enc = B200Encoder(weights, device=local_rank, max_tokens=tokens)

    n_batches = min(args.warmup + args.steps, 8)
    host_ids, cu = synth_batches(n_batches, chunks, rank)
    d_ids = [torch.from_numpy(x).to(device) for x in host_ids]
    d_cu = torch.from_numpy(cu).to(device)
    d_out = torch.empty((chunks, HIDDEN), dtype=torch.float32, device=device)
    stream = torch.cuda.current_stream(device)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---------------- device-resident timing (value): nothing but the forwards inside the events ----------------
    for i in range(args.warmup):
        enc.forward_device(d_ids[i % n_batches], d_cu, cu, d_out)
    torch.cuda.synchronize(device)
    barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for i in range(args.steps):
            enc.forward_device(d_ids[(args.warmup + i) % n_batches], d_cu, cu, d_out)
        e1.record(stream)
        torch.cuda.synchronize(device)
    barrier()
    ms_total = e0.elapsed_time(e1)
    checksum = float(d_out.double().sum().item())
    last_batch = (args.warmup + args.steps - 1) % n_batches
    first_rows = d_out[:8].cpu().numpy()          # for the parity check below

    # ---------------- second pass, instrumented: CUDA events around every launch (per-kernel breakdown, roofline) ---------
    enc.profile_begin(args.steps * 64 + 64)
    for i in range(args.steps):
        enc.forward_device(d_ids[(args.warmup + i) % n_batches], d_cu, cu, d_out)
    torch.cuda.synchronize(device)
    prof = enc.profile_end()

    # ---------------- end to end through the host-buffer C-ABI call ----------------
    for i in range(min(args.warmup, 2)):
        enc.embed_packed(host_ids[i % n_batches], cu)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out_host = enc.embed_packed(host_ids[(args.warmup + i) % n_batches], cu)
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---------------- ragged variant: lengths uniform 16..256, packed -- no padding work ----------------
    ragged = None
    if rank == 0:
        from tests.synth import synth_token_batch
        r_ids, r_cu = synth_token_batch(seed=77, n_seq=chunks, seq_len=SEQ_LEN, ragged=True)
        for _ in range(2):
            enc.embed_packed(r_ids, r_cu)
        t0 = time.perf_counter()
        for _ in range(max(3, args.steps // 2)):
            enc.embed_packed(r_ids, r_cu)
        r_s = (time.perf_counter() - t0) / max(3, args.steps // 2)
        ragged = {"workload": f"{chunks} chunks, lengths uniform in [16, {SEQ_LEN}], packed (host ids in, host embeddings out)",
                  "tokens": int(r_cu[-1]), "chunks_per_s": chunks / r_s, "tokens_per_s": float(r_cu[-1]) / r_s,
                  "padded_tokens_a_pad_to_longest_batch_would_process": chunks * SEQ_LEN}
    barrier()

    if world > 1:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0]), float(t[1])

    ms_per_step = ms_total / args.steps
    value = world * chunks * args.steps / (ms_total / 1e3)
    e2e_value = world * chunks * args.steps / e2e_s

    # ---------------- roofline of the dominant kernel ----------------
    flops = {
        "gemm_qkv": 2.0 * tokens * 384 * 1152, "gemm_out_ln": 2.0 * tokens * 384 * 384,
        "gemm_up_gelu": 2.0 * tokens * 384 * 1536, "gemm_down_ln": 2.0 * tokens * 1536 * 384,
        "gemm_mlp_fused": 4.0 * tokens * 384 * 1536,   # FFN-up + GELU + FFN-down + residual in one launch (drag_mlp.cuh)
        "attention": 4.0 * chunks * SEQ_LEN * SEQ_LEN * 384,
    }
    kernel_ms = sum(v["ms"] for v in prof.values())
    breakdown = {}
    for name, v in prof.items():
        if v["launches"] == 0:
            continue
        avg_ms = v["ms"] / v["launches"]
        item = {"launches_per_step": v["launches"] / args.steps, "avg_ms": avg_ms, "share_of_kernel_time": v["ms"] / kernel_ms}
        if name in flops:
            item["tflops"] = flops[name] / (avg_ms / 1e3) / 1e12
        breakdown[name] = item
    # The dominant kernels are the tcgen05 GEMMs (gemm_kernel<...> instantiations + the fused feed-forward mlp_kernel:
    # one pipeline design, together the largest share of the step); attention is reported next to them.
    gemm_names = [n for n in breakdown if n.startswith("gemm_")]
    gemm_ms = sum(prof[n]["ms"] for n in gemm_names)
    gemm_launches = sum(prof[n]["launches"] for n in gemm_names)
    gemm_flop = sum(flops[n] * prof[n]["launches"] for n in gemm_names)   # the CLS-only last layer launches fewer full-size GEMMs
    achieved = gemm_flop / (gemm_ms / 1e3) / 1e12
    clk_hz = 1e6 * (clocks.summary().get("sm_mhz") or 1965.0)
    exp_per_launch = chunks * 12.0 * SEQ_LEN * SEQ_LEN
    att = breakdown["attention"]
    att["exp_per_s"] = exp_per_launch / (att["avg_ms"] / 1e3)
    att["mufu_peak_exp_per_s"] = 16.0 * 148 * clk_hz   # measured: 16 ex2 / clock / SM (scripts/ubench/pipes.cu)
    att["frac_of_mufu_peak"] = att["exp_per_s"] / att["mufu_peak_exp_per_s"]
    att["frac_of_tensor_peak"] = att["tflops"] / peaks["tflops_sustained"]
    att["note"] = ("head_dim 32: 1 exponential per 128 flop, so the MUFU pipe (16 ex2/clk/SM, measured) bounds this op "
                   "(0.19 ms per 1024 chunks at 1.9 GHz) long before the tensor pipe does; frac_of_mufu_peak is its roofline "
                   "fraction; see DESIGN.md section 4")
    # algorithmic FLOPs of the step as executed (last layer: Q/attention/out-proj/FFN for the [CLS] rows only) next to
    # the MFU convention's 12.080 GFLOP/chunk (dense, all rows of all layers)
    roofline = {
        "kernel": "tcgen05 GEMM kernels: gemm_kernel template (QKV, out-proj; FFN-up / FFN-down for small batches) + mlp_kernel "
                  "(FFN-up + GELU + FFN-down + residual fused, the 1536-wide intermediate stays in tensor memory)",
        "bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
        "frac": achieved / peaks["tflops_sustained"],
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, launch-weighted mean over the step's GEMM-class launches, from
        # the ncu --set full captures in profiles/r02_k_ncu_full_summary.md: mlp_kernel 375 MB (measured at 262 144 tokens),
        # QKV 146 MB and out-projection 123.5 MB at 65 536 tokens (x4: an upper bound on what L2 retention hides there)
        "traffic": 487.0e6 * tokens / 262144.0,
        "traffic_source": "profiles/r02_k_ncu_full_summary.md (ncu --set full; scaled linearly in tokens); algorithmic "
                          "operand+output bytes per launch average 611 MB at 262 144 tokens",
        "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step)",
        "flop_per_launch": gemm_flop / gemm_launches, "avg_launch_ms": gemm_ms / gemm_launches,
        "share_of_step": gemm_ms / kernel_ms,
        "by_instantiation": {n: {"tflops": breakdown[n]["tflops"], "frac": breakdown[n]["tflops"] / peaks["tflops_sustained"],
                                 "avg_ms": breakdown[n]["avg_ms"]} for n in gemm_names},
        "attention_avg_ms": att["avg_ms"], "attention_frac_of_mufu_peak": att["frac_of_mufu_peak"],
        "attention_share_of_step": att["share_of_kernel_time"],
        "whole_step_tflops_per_gpu": value / world * FLOP_PER_CHUNK / 1e12,
        "whole_step_frac_of_sustained_peak": value / world * FLOP_PER_CHUNK / 1e12 / peaks["tflops_sustained"],
        "whole_step": {"tflops_per_gpu": value / world * FLOP_PER_CHUNK / 1e12,
                       "frac_of_sustained_peak": value / world * FLOP_PER_CHUNK / 1e12 / peaks["tflops_sustained"],
                       "flop_per_chunk": FLOP_PER_CHUNK,
                       "note": "MFU convention: dense 12.080 GFLOP per 256-token chunk (all rows of all 12 layers); the "
                               "last layer actually runs Q/attention/out-proj/FFN on the pooled [CLS] rows only"},
    }

    line = {
        "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": f"configs[1]: synthetic {SEQ_LEN}-token chunks (packed, no padding), bge-small-en architecture "
                        "(12 layers, 384-d, 12 heads, FFN 1536), seeded random-init weights; data-parallel over GPUs",
            "chunks_per_gpu_per_step": chunks, "global_chunks_per_step": world * chunks, "seq_len": SEQ_LEN,
            "parallelism": f"dp{world}",
            "l2": "per-step activation working set (~2 GB at 1024 chunks) is far larger than the 126 MB L2 and "
                  "token-id batches rotate between steps; weights (42 MB bf16) are L2-resident by design",
            "output_checksum": checksum,
            "timing": "value: CUDA events around the K forwards only; the per-kernel breakdown comes from a second, instrumented pass",
        },
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": int(tokens * 4 + (chunks + 1) * 4),
                "d2h_bytes_per_step": int(chunks * HIDDEN * 4),
                "api": "drag_encoder_embed_host via B200Encoder.embed_packed (host int32 ids in, host float32 embeddings out)"},
        "gpu_launches": int(sum(v["launches"] for v in prof.values())),
        "roofline": roofline,
        "parity": {},
        "extra": {"kernels": breakdown},
    }
    if ragged is not None:
        line["extra"]["ragged_16_256"] = ragged
        line["e2e"]["ragged_16_256_chunks_per_s"] = ragged["chunks_per_s"]
        line["e2e"]["ragged_16_256_tokens_per_s"] = ragged["tokens_per_s"]
    parity = line["parity"]

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = time_cpu_baseline(weights)
        # parity at bench scale: the first 8 chunks of the timed 1024-chunk forward against the fp32 CPU reference path
        try:
            from oracle import encoder as oenc

            embed, _ = cpu_encoder(weights)
            want = embed(oenc.packed_to_lists(host_ids[last_batch][: 8 * SEQ_LEN], cu[:9]))
            cos = (first_rows.astype(np.float64) * want).sum(1) / (np.linalg.norm(first_rows, axis=1) * np.linalg.norm(want, axis=1))
            parity["encoder_min_cosine_8_of_1024_chunks"] = float(cos.min())
            parity["encoder_cosine_ok"] = bool(cos.min() >= 0.9995)
            parity["e2e_equals_device_path"] = bool(np.array_equal(out_host[:8], first_rows))   # same batch, both entry points
        except Exception as exc:  # noqa: BLE001
            parity["encoder_error"] = repr(exc)
    elif rank == 0:
        line["cpu_baseline"] = None

    lib_rows = None
    if rank == 0 and world == 1 and not args.no_library_baseline:
        try:
            line["gpu_library_baseline"], lib_rows = gpu_library_encoder(torch, device, weights, host_ids, cu, chunks)
            glb = line["gpu_library_baseline"]
            glb["unit"] = "chunks/s"
            glb["value"] = max(glb["chunks_per_s_bs32"], glb["chunks_per_s_bs256"])
            glb["speedup_over_library"] = value / glb["value"]
            ours0 = enc.embed_packed(host_ids[0], cu)
            glb["min_cosine_ours_vs_library_fp16"] = float(((ours0 * lib_rows).sum(1) / (np.linalg.norm(ours0, axis=1) * np.linalg.norm(lib_rows, axis=1))).min())
        except Exception as exc:  # noqa: BLE001
            line["gpu_library_baseline"] = {"error": repr(exc)}

    enc.close()
    del enc, d_ids
    torch.cuda.empty_cache()

    if world == 1 and not args.no_search:
        try:
            search = {"batch": bench_search(torch, device, peaks, args.search_rows, 1000, 100, steps=5, warmup=2, seed=2,
                                            library_baseline=not args.no_library_baseline),
                      "batch256_top20_1m": bench_search(torch, device, peaks, 1_000_000, 256, 20, steps=20, warmup=3, seed=5),
                      "single_query": bench_search(torch, device, peaks, 1_000_000, 1, 20, steps=50, warmup=5, seed=5)}
            line["extra"]["search"] = search
            # the second half of the headline metric, as scalars the driver keeps
            sb, s1, s256 = search["batch"], search["single_query"], search["batch256_top20_1m"]
            roofline.update({
                "search_10m_qps": sb["queries_per_s"], "search_10m_ms_per_batch": sb["ms_per_batch"],
                "search_10m_frac": sb["roofline"]["frac"], "search_bound": sb["roofline"]["bound"],
                "search_10m_achieved": sb["roofline"]["achieved"], "search_10m_peak": sb["roofline"]["peak"],
                "search_10m_unit": sb["roofline"]["unit"], "search_10m_fallback_queries": sb["fallback_queries"],
                "search_single_1m_ms": s1["ms_per_batch"], "search_single_1m_hbm_frac": s1["roofline"]["frac"],
                "search_batch256_1m_ms": s256["ms_per_batch"], "search_batch256_1m_frac": s256["roofline"]["frac"],
            })
            line["e2e"]["search_qps"] = sb["e2e_queries_per_s"]
            line["e2e"]["search_note"] = "host float64 queries in, host (distance, row id) out: drag_topk_batch through DeviceMatrix.topk"
            if "gpu_library_baseline" in sb and "queries_per_s" in sb["gpu_library_baseline"] and isinstance(line.get("gpu_library_baseline"), dict):
                line["gpu_library_baseline"]["search_qps"] = sb["gpu_library_baseline"]["queries_per_s"]
                line["gpu_library_baseline"]["search_speedup_over_library"] = sb["queries_per_s"] / sb["gpu_library_baseline"]["queries_per_s"]
                line["gpu_library_baseline"]["search_library_ids_equal_to_exact_frac"] = sb["gpu_library_baseline"].get("ids_equal_to_exact_frac")
            if not args.no_cpu_baseline:
                cpu_s = cpu_search_baseline()
                parity.update(cpu_s.pop("parity"))
                search["cpu_baseline"] = cpu_s
                if isinstance(line.get("cpu_baseline"), dict):
                    line["cpu_baseline"]["search_qps"] = cpu_s["value"]
                    line["cpu_baseline"]["search_sample"] = cpu_s["sample"]
            qp = bench_query_path(torch, device, weights)
            line["extra"]["query_path"] = qp
            line["e2e"]["query_batch1_host_api_ms_p50"] = qp["batch1"]["host_api_ms_p50"]
            line["e2e"]["query_batch256_qps_host_api"] = qp["batch256"]["queries_per_s_host_api"]
            line["e2e"]["query_batch1_host_api_ms_p50_under_indexing_load"] = qp["batch1_under_indexing_load"]["host_api_ms_p50"]
            line["e2e"]["query_batch1_host_api_ms_p99_under_indexing_load"] = qp["batch1_under_indexing_load"]["host_api_ms_p99"]
            roofline["query_batch1_device_ms_p50"] = qp["batch1"]["device_ms_p50"]
            roofline["query_batch256_qps_device"] = qp["batch256"]["queries_per_s_device"]
            line["extra"]["text_path"] = bench_text_path(torch, device, weights)
            line["e2e"]["text_to_embedding_chunks_per_s"] = line["extra"]["text_path"]["chunks_per_s_build_embeddings"]
            wide = bench_wide_dims(torch, device, peaks)
            line["extra"]["search_wide_dims"] = wide
            for dname in ("d1024", "d1408"):
                roofline[f"search_{dname}_single_250k_ms"] = wide[dname]["q1"]["ms"]
                roofline[f"search_{dname}_single_250k_hbm_frac"] = wide[dname]["q1"]["hbm_frac"]
                roofline[f"search_{dname}_q16_250k_ms"] = wide[dname]["q16"]["ms"]
            ib = bench_index_build(torch, device)
            line["extra"]["index_build"] = ib
            line["e2e"]["index_build_chunks_per_s_1m"] = ib["chunks_per_s"]
            line["e2e"]["index_build_generic_item_path_chunks_per_s"] = ib["generic_item_path_chunks_per_s"]
        except Exception as exc:  # noqa: BLE001 - the secondary metric must not lose the headline line
            line["extra"]["search_error"] = repr(exc)

    if world > 1 and not args.no_search:
        try:
            sh = bench_search_sharded(torch, dist, device, rank, world, peaks, args.shard_rows, args.shard_queries, 100,
                                      steps=max(10, min(args.steps, 20)), warmup=2)
            line["extra"]["search_sharded"] = sh
            roofline.update({"sharded_qps": sh["queries_per_s"], "sharded_ms_per_batch": sh["ms_per_batch"],
                             "sharded_frac": sh["roofline"]["frac"], "sharded_bound": sh["roofline"]["bound"],
                             "sharded_rows": world * args.shard_rows, "sharded_queries": args.shard_queries,
                             "sharded_steps": sh["steps"]})
            for name, ms in sh["stage_ms_rank0"].items():
                roofline[f"sharded_stage_ms_{name}"] = ms
            line["e2e"]["sharded_search_qps"] = sh["queries_per_s"]
            parity.update({"sharded_" + k2: v2 for k2, v2 in sh["parity"].items()})
        except Exception as exc:  # noqa: BLE001
            line["extra"]["search_sharded"] = {"error": repr(exc)}

    parity["all_ok"] = bool(all(v for k2, v in parity.items() if isinstance(v, bool)))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
