"""GPU parity of the batched search path (drag_topk_batch: tcgen05 candidate scores under a
certified error bound + float64 re-rank) against the float64 scan (drag_topk) and the oracle."""

import ctypes as C

import numpy as np
import pytest

from oracle import search as osearch
from tests.synth import synth_matrix, synth_queries

pytestmark = pytest.mark.gpu

C1 = 2.0**-8 * 1.01 + 2.0**-11  # the bound drag_topk_batch certifies with (drag_topk.cu run_batch)


def _bf16_round(x: np.ndarray) -> np.ndarray:
    import torch

    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def _tc_keys(m: np.ndarray, q: np.ndarray, metric: str) -> np.ndarray:
    """Approximate keys of the score kernel for every (query, row) of a small matrix."""
    import torch

    from dial_rag_b200 import _native
    from dial_rag_b200.device_index import DeviceMatrix, _metric_code, _ptr

    dm = DeviceMatrix(m)
    state = dm._batch_prepare()
    lib = _native.load()
    dev = dm.matrix.device
    d_q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float64)).to(dev)
    out = torch.empty((len(q), dm.n_rows), dtype=torch.float32, device=dev)
    need = C.c_size_t(0)
    _native.check(lib.drag_topk_batch_workspace_bytes(0, len(q), 1, dm.dim, C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    colvec = {"inner_product": None, "cosine_sim": state["inv"]}.get(metric, dm.row_sq)
    _native.check(lib.drag_debug_tc_keys(0, _ptr(state["shadow"]), dm.n_rows, dm.dim, _ptr(colvec), _metric_code(metric),
                                         _ptr(d_q), len(q), _ptr(out), _ptr(ws), need.value,
                                         torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("normalise", [True, False])
def test_tc_keys_within_certified_bound(normalise):
    """|approximate key - exact key| stays inside E = C1*|q|*|d| (+ key-arithmetic rounding), and the
    tensor-core fp32 accumulation alone (vs the exact sum of the bf16-rounded products) inside 2^-11."""
    m = synth_matrix(seed=21, rows=3000, dim=384, normalise=normalise)
    q = synth_queries(seed=22, n=150, dim=384, normalise=normalise)
    if not normalise:
        m *= np.random.Generator(np.random.PCG64(5)).uniform(0.05, 20.0, size=(len(m), 1)).astype(np.float32)
    m64, qn = m.astype(np.float64), np.linalg.norm(q, axis=1)
    dn = np.linalg.norm(m64, axis=1)
    exact = q @ m64.T
    scale = qn[:, None] * dn[None, :]
    got = _tc_keys(m, q, "inner_product").astype(np.float64)
    assert np.max(np.abs(got - exact) / scale) <= C1
    rounded = _bf16_round(q.astype(np.float32)).astype(np.float64) @ _bf16_round(m).astype(np.float64).T
    acc_err = np.max(np.abs(got - rounded) / scale)
    assert acc_err <= 2.0**-11, acc_err
    print(f"max |s~-s|/(|q||d|) = {np.max(np.abs(got - exact) / scale):.3e} (bound {C1:.3e}); accumulation {acc_err:.3e}")
    # (sq)euclidean keys: s - |d|^2/2 with the float32 numpy-order |d|^2
    row_sq = np.sum(m**2, axis=1).astype(np.float64)
    got = _tc_keys(m, q, "sqeuclidean_dist").astype(np.float64)
    dmax = dn.max()
    bound = (C1 + 2.0**-22) * dmax * qn[:, None] + 2.0**-22 * dmax * dmax
    assert np.all(np.abs(got - (exact - 0.5 * row_sq[None, :])) <= bound)
    # cosine keys: s / max(|d|, 1e-8)
    got = _tc_keys(m, q, "cosine_sim").astype(np.float64)
    assert np.all(np.abs(got - exact / np.maximum(dn, 1e-8)[None, :]) <= (C1 + 2.0**-19) * qn[:, None])


def _planted(rows, seed=2, normalise=True):
    m = synth_matrix(seed=seed, rows=rows, dim=384, normalise=normalise)
    rng = np.random.Generator(np.random.PCG64(9))
    dst = rng.integers(0, len(m), size=3000)
    m[dst] = m[rng.integers(0, len(m), size=3000)]
    return m, dst


@pytest.mark.parametrize("metric", osearch.ALL_METRICS)
def test_batch_equals_scan_and_oracle_200k(metric):
    """200k x 384 with planted duplicates, 300 queries: the batched path returns exactly what the
    float64 scan returns (rows and distances bit for bit), and the oracle's rows."""
    import torch

    from dial_rag_b200.device_index import DeviceMatrix

    m, dst = _planted(200_003)
    q = synth_queries(seed=3, n=300, dim=384)
    q[0] = m[dst[0]].astype(np.float64)   # exact hit with duplicates
    q[1] = 3.7 * q[1]                      # un-normalised query
    dm = DeviceMatrix(m)
    d_q = torch.from_numpy(q).cuda()
    for k in (1, 20, 100, 256):
        bd, br, bc = dm.topk_device(d_q, k, metric)
        assert dm.last_batch_fallbacks <= (3 if metric == "euclidean_dist" else 0), dm.last_batch_fallbacks
        sd, sr, sc = dm.topk_device(d_q[:24], k, metric, allow_batch=False)
        assert torch.equal(br[:24], sr) and torch.equal(bc[:24], sc), (metric, k)
        assert torch.equal(bd[:24].view(torch.int64), sd.view(torch.int64)), (metric, k)
        rows = br.cpu().numpy()
        for i in (0, 1, 2, 150, 299):
            want_rows, _ = osearch.topk_rows(metric, k, q[i], m)
            assert np.array_equal(rows[i], want_rows), (metric, k, i)


def test_batch_unnormalised_rows_and_ragged_query_count():
    """Rows with norms spread over 0.1..30, 131 queries (not a multiple of the 128-query tile),
    product-default metric and inner product, rows not a multiple of the 256-row tile."""
    import torch

    from dial_rag_b200.device_index import DeviceMatrix

    m = synth_matrix(seed=31, rows=70_001, dim=384, normalise=False)
    m *= np.random.Generator(np.random.PCG64(6)).uniform(0.1, 30.0, size=(len(m), 1)).astype(np.float32) / 19.6
    q = synth_queries(seed=32, n=131, dim=384, normalise=False)
    dm = DeviceMatrix(m)
    d_q = torch.from_numpy(q).cuda()
    for metric in ("sqeuclidean_dist", "inner_product", "cosine_sim"):
        bd, br, _ = dm.topk_device(d_q, 50, metric)
        sd, sr, _ = dm.topk_device(d_q, 50, metric, allow_batch=False)
        assert torch.equal(br, sr), metric
        assert torch.equal(bd.view(torch.int64), sd.view(torch.int64)), metric
        for i in (0, 130):
            want_rows, _ = osearch.topk_rows(metric, 50, q[i], m)
            assert np.array_equal(br[i].cpu().numpy(), want_rows), (metric, i)


def test_batch_adversarial_ties_fall_back_to_scan():
    """Every row identical: all keys tie, the candidate lists overflow, status=1 sends the queries to
    the float64 scan and the answer is still the first k row ids."""
    from dial_rag_b200.device_index import DeviceMatrix

    m = np.tile(synth_matrix(seed=8, rows=1, dim=384), (50_000, 1))
    q = synth_queries(seed=9, n=16, dim=384)
    dm = DeviceMatrix(m)
    _, rows, count = dm.topk(q, 100, "inner_product")
    assert dm.last_batch_fallbacks == 16
    assert np.array_equal(rows, np.tile(np.arange(100), (16, 1))) and count.tolist() == [100] * 16
    # clustered data that does NOT overflow: 40 exact copies of every one of 1000 base rows
    base = synth_matrix(seed=10, rows=1000, dim=384)
    m = np.repeat(base, 40, axis=0)
    dm = DeviceMatrix(m)
    q = synth_queries(seed=11, n=64, dim=384)
    q[5] = base[17].astype(np.float64)
    _, rows, _ = dm.topk(q, 100, "sqeuclidean_dist")
    for i in (0, 5, 63):
        want_rows, _ = osearch.topk_rows("sqeuclidean_dist", 100, q[i], m)
        assert np.array_equal(rows[i], want_rows), i


def test_batch_bf16_storage():
    """bf16 index (BASELINE configs[3] storage): exact top-k of the stored (rounded) values."""
    import torch

    from dial_rag_b200.device_index import DeviceMatrix

    m = synth_matrix(seed=4, rows=120_000, dim=384)
    q = synth_queries(seed=14, n=260, dim=384)
    rounded = _bf16_round(m)
    dm = DeviceMatrix(m, storage="bf16", row_id_base=1_000_000)
    dist, rows, _ = dm.topk(q, 100, "inner_product")
    assert dm.last_batch_fallbacks == 0
    sd, sr, _ = dm.topk_device(torch.from_numpy(q[:8]).cuda(), 100, "inner_product", allow_batch=False)
    assert np.array_equal(rows[:8], sr.cpu().numpy()) and np.array_equal(dist[:8], sd.cpu().numpy())
    for i in (0, 100, 259):
        want_rows, want_d = osearch.topk_rows("inner_product", 100, q[i], rounded)
        assert np.array_equal(rows[i] - 1_000_000, want_rows)
        np.testing.assert_allclose(dist[i], want_d, rtol=1e-11, atol=1e-12)
        full = -(m.astype(np.float64) @ q[i])
        assert np.max(np.abs(full[want_rows] - dist[i])) <= 2.0**-8   # BASELINE's stated score tolerance


def test_batch_nan_and_inf_inputs_use_the_scan():
    from dial_rag_b200.device_index import DeviceMatrix

    m = synth_matrix(seed=41, rows=20_000, dim=384)
    q = synth_queries(seed=42, n=10, dim=384)
    q[3, 7] = np.nan
    dm = DeviceMatrix(m)
    _, rows, _ = dm.topk(q, 10, "inner_product")
    assert dm.last_batch_fallbacks >= 1
    assert np.array_equal(rows[3], np.arange(10))  # every distance NaN: stable order = row order
    want_rows, _ = osearch.topk_rows("inner_product", 10, q[4], m)
    assert np.array_equal(rows[4], want_rows)
    m[123, 5] = np.inf
    dm = DeviceMatrix(m)
    assert not dm._use_batch(10, 10, 3)  # non-finite rows: the certificate does not apply


def test_batch_full_size_property_2m_rows_1000_queries():
    """At BASELINE scale (on-device synthetic rows): the batched answer equals the float64 scan's on a
    sample of the queries, distances ascend and ids are unique."""
    import torch

    from dial_rag_b200.device_index import DeviceMatrix

    g = torch.Generator(device="cuda").manual_seed(2)
    mat = torch.randn((2_000_000, 384), generator=g, device="cuda")
    mat /= mat.norm(dim=1, keepdim=True)
    mat[1_500_000:1_500_100] = mat[:100]  # duplicates far apart
    q = torch.randn((1000, 384), generator=g, device="cuda")
    q = (q / q.norm(dim=1, keepdim=True)).double()
    q[0] = mat[5].double()
    dm = DeviceMatrix(mat)
    bd, br, bc = dm.topk_device(q, 100, "inner_product")
    assert dm.last_batch_fallbacks == 0
    assert bool((bd[:, 1:] >= bd[:, :-1]).all()) and bc.tolist() == [100] * 1000
    assert all(len(set(r)) == 100 for r in br[::97].cpu().tolist())
    pick = torch.tensor([0, 1, 2, 3, 500, 501, 998, 999], device="cuda")
    sd, sr, _ = dm.topk_device(q[pick].contiguous(), 100, "inner_product", allow_batch=False)
    assert torch.equal(br[pick], sr) and torch.equal(bd[pick].view(torch.int64), sd.view(torch.int64))
    assert br[0, 0].item() == 5 and br[0, 1].item() == 1_500_005
