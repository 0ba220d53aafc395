"""GPU parity of the embedding half: each fused kernel in isolation (drag_debug_*), the
layer taps, and the end-to-end embeddings against the fp32 oracle / HF BertModel goldens.
Tolerance: BASELINE.md section 5 -- cosine >= 0.9995 against the fp32 reference."""

import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

from oracle import encoder as oenc
from tests.synth import synth_token_batch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COS_BAR = 0.9995


def _lib():
    from dial_rag_b200 import _native

    return _native, _native.load()


def _bf16(x):
    return x.to(torch.bfloat16)


def _row_stats(x, parts=3):
    """[M][3] float2 partial (sum, sum^2) of the rows of x, split unevenly over the 3 slots."""
    xf = x.float()
    M, W = xf.shape
    cuts = [0, W // 3, W // 3 + 7, W]
    out = torch.zeros(M, parts, 2, device=x.device)
    for i in range(parts):
        blk = xf[:, cuts[i]:cuts[i + 1]]
        out[:, i, 0] = blk.sum(1)
        out[:, i, 1] = (blk * blk).sum(1)
    return out.contiguous()


@pytest.mark.parametrize("variant,N,K", [(0, 1152, 384), (1, 1536, 384), (2, 384, 384), (2, 384, 1536), (0, 192, 64),
                                         (10, 1152, 384), (11, 1536, 384), (12, 384, 384), (12, 384, 1536), (10, 192, 64),
                                         (22, 384, 1536), (32, 384, 1536)])
@pytest.mark.parametrize("M", [128, 77, 1000, 20000])
def test_tcgen05_gemm_vs_torch(variant, N, K, M):
    """Each fused GEMM epilogue (LayerNorm folded, see drag_gemm.cuh) against plain torch fp32;
    variants 10-12 are the CTA-pair (cta_group::2) forms of 0-2, +20 = fp16 operands (the FFN-down
    GEMM reads the fp16 GELU output).  The GELU epilogue (variant 1) writes fp16."""
    native, lib = _lib()
    code, variant = variant, variant % 10
    f16_in = code >= 20
    pair_form = (code % 20) >= 10
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + variant)
    a = torch.randn(M, K, device="cuda", generator=g) + 0.3
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    a, w = (a.half(), w.half()) if f16_in else (_bf16(a), _bf16(w))
    colc = torch.randn(N, device="cuda", generator=g) * 0.2
    cold = torch.randn(N, device="cuda", generator=g) * 0.1
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5
    res = _bf16(torch.randn(M, N, device="cuda", generator=g) * 2 + 0.5)
    eps = 1e-12
    normed = res if variant == 2 else a  # rows the folded LayerNorm statistics describe
    stats = _row_stats(normed)
    width = normed.shape[1]
    mu = normed.float().mean(1, keepdim=True)
    rstd = torch.rsqrt(normed.float().var(1, unbiased=False, keepdim=True) + eps)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16 if variant == 1 else torch.bfloat16)
    out_stats = torch.full((M, 3, 2), float("nan"), device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    native.check(lib.drag_debug_gemm(0, code, a.data_ptr(), w.data_ptr(), colc.data_ptr(), cold.data_ptr(),
                                     gamma.data_ptr(), stats.data_ptr(), res.data_ptr(), out.data_ptr(),
                                     out_stats.data_ptr(), M, N, K, 1.0 / width, eps, stream))
    torch.cuda.synchronize()
    acc = a.float() @ w.float().T
    if variant == 2:
        ref = acc + cold + (res.float() - mu) * rstd * gamma
    else:
        ref = rstd * (acc - mu * colc) + cold
        if variant == 1:
            ref = torch.nn.functional.gelu(ref)
    got = out.float()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 0.01 * scale + 0.02, (code, M, N, K, err, scale)
    assert (got - ref).abs().mean().item() <= 0.004 * max(ref.abs().mean().item(), 1e-3) + 1e-3
    if variant == 2:
        tile, parts = (192, 2) if pair_form else (128, 3)   # the pair kernels leave the third slot untouched
        want = torch.stack([torch.stack((ref[:, i * tile:(i + 1) * tile].sum(1), (ref[:, i * tile:(i + 1) * tile] ** 2).sum(1)), 1)
                            for i in range(parts)], 1)
        assert torch.allclose(out_stats[:, :parts], want, rtol=2e-4, atol=2e-2), (out_stats[:, :parts] - want).abs().max()


@pytest.mark.parametrize("M", [256, 77, 1000, 20000, 40001])
def test_fused_mlp_vs_torch(M):
    """The fused feed-forward kernel (drag_mlp.cuh: FFN-up + GELU + FFN-down + residual LayerNorm in one launch, the
    1536-wide intermediate stays in tensor memory) against plain torch fp32, and against the two separate GEMM
    kernels it replaces."""
    native, lib = _lib()
    H, F = 384, 1536
    g = torch.Generator(device="cuda").manual_seed(M + 5)
    x = torch.randn(M, H, device="cuda", generator=g) * 1.5 + 0.3
    x[:, 7] *= 12.0                                       # an outlier channel, as trained BERTs have
    x = _bf16(x)
    w1g = _bf16(torch.randn(F, H, device="cuda", generator=g) * 0.05)
    up_c = w1g.float().sum(1).contiguous()                # the folded LayerNorm term c_n = sum_k W1g_nk
    up_d = (torch.randn(F, device="cuda", generator=g) * 0.3).contiguous()
    w2 = (torch.randn(H, F, device="cuda", generator=g) * 0.05).half()
    cold = (torch.randn(H, device="cuda", generator=g) * 0.1).contiguous()
    gamma = (torch.rand(H, device="cuda", generator=g) + 0.5).contiguous()
    eps = 1e-12
    stats = _row_stats(x)
    mu = x.float().mean(1, keepdim=True)
    rstd = torch.rsqrt(x.float().var(1, unbiased=False, keepdim=True) + eps)
    out = torch.full((M, H), float("nan"), device="cuda", dtype=torch.bfloat16)
    out_stats = torch.full((M, 3, 2), float("nan"), device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    native.check(lib.drag_debug_mlp(0, x.data_ptr(), stats.data_ptr(), w1g.data_ptr(), up_c.data_ptr(), up_d.data_ptr(),
                                    w2.data_ptr(), cold.data_ptr(), gamma.data_ptr(), out.data_ptr(), out_stats.data_ptr(),
                                    M, eps, stream))
    torch.cuda.synchronize()
    h = torch.nn.functional.gelu(rstd * (x.float() @ w1g.float().T - mu * up_c) + up_d)
    ref = h @ w2.float().T + cold + (x.float() - mu) * rstd * gamma
    got = out.float()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 0.01 * scale + 0.03, (M, err, scale)
    assert (got - ref).abs().mean().item() <= 0.004 * max(ref.abs().mean().item(), 1e-3) + 2e-3
    want = torch.stack([torch.stack((ref[:, i * 192:(i + 1) * 192].sum(1), (ref[:, i * 192:(i + 1) * 192] ** 2).sum(1)), 1)
                        for i in range(2)], 1)
    # (the statistics are taken from the kernel's own fp32 values: tanh-form GELU, fp16 intermediate)
    assert torch.allclose(out_stats[:, :2], want, rtol=1e-3, atol=0.1), (out_stats[:, :2] - want).abs().max()
    assert (out_stats[:, 2] == 0).all()
    # ... and the unfused pair kernels on the same data (same rounding points: fp16 intermediate, bf16 output)
    hbuf = torch.empty(M, F, device="cuda", dtype=torch.float16)
    out2 = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    st2 = torch.zeros(M, 3, 2, device="cuda")
    native.check(lib.drag_debug_gemm(0, 11, x.data_ptr(), w1g.data_ptr(), up_c.data_ptr(), up_d.data_ptr(), None,
                                     stats.data_ptr(), None, hbuf.data_ptr(), None, M, F, H, 1.0 / H, eps, stream))
    native.check(lib.drag_debug_gemm(0, 32, hbuf.data_ptr(), w2.data_ptr(), None, cold.data_ptr(), gamma.data_ptr(),
                                     stats.data_ptr(), x.data_ptr(), out2.data_ptr(), st2.data_ptr(), M, H, F, 1.0 / H, eps, stream))
    torch.cuda.synchronize()
    assert (got - out2.float()).abs().max().item() <= 0.01 * scale + 0.03


@pytest.mark.parametrize("lens", [[1, 2, 17, 64, 65, 128, 256, 300, 512, 33], [512], [300, 77], [130] * 4],
                         ids=["ragged10", "one512", "two", "four130"])
@pytest.mark.parametrize("variant", [3, 0])
def test_attention_vs_torch(variant, lens):
    """Every attention kernel against torch fp32 softmax(QK^T/sqrt(32))V per packed sequence; the short
    batches exercise the 128- and 64-query tiles the mma.sync kernel picks when few sequences are in flight."""
    native, lib = _lib()
    heads, hd = 12, 32
    cu = np.zeros(len(lens) + 1, dtype=np.int32)
    cu[1:] = np.cumsum(lens)
    T = int(cu[-1])
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = _bf16(torch.randn(T, 3 * heads * hd, device="cuda", generator=g) * 1.5)
    ctx = torch.full((T, heads * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    d_cu = torch.from_numpy(cu).cuda()
    native.check(lib.drag_debug_attention(0, variant, qkv.data_ptr(), ctx.data_ptr(), d_cu.data_ptr(), len(lens), T,
                                          max(lens), heads, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = ctx.float()
    assert torch.isfinite(got).all()
    for i, n in enumerate(lens):
        blk = qkv[cu[i]:cu[i + 1]].float().view(n, 3, heads, hd)
        q, k, v = (blk[:, j].permute(1, 0, 2) for j in range(3))
        p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(hd), dim=-1)
        ref = (p @ v).permute(1, 0, 2).reshape(n, heads * hd)
        err = (got[cu[i]:cu[i + 1]] - ref).abs().max().item()
        assert err <= 0.03, (variant, n, err)


@pytest.mark.parametrize("variant", [3, 0])
def test_attention_peaked_garbage_tail_and_batch_invariance(variant):
    """What the encoder does to an attention kernel and random inputs do not: peaked logits (the reference of a row is
    raised between key blocks), NaN in the never-written rows behind the last sequence (0 x NaN must not reach anybody's
    output), and bit-identical rows whether a sequence is processed alone or inside a batch."""
    native, lib = _lib()
    heads, hd = 12, 32
    lens = [5, 130, 101, 2, 257, 64, 400]
    cu = np.zeros(len(lens) + 1, dtype=np.int32)
    cu[1:] = np.cumsum(lens)
    T = int(cu[-1])
    tail = 200
    g = torch.Generator(device="cuda").manual_seed(9)
    qkv = torch.randn(T + tail, 3 * heads * hd, device="cuda", generator=g) * 1.5
    qkv[:, : 2 * heads * hd] *= 3.0                       # logits of +-100 before the 1/sqrt(32): sharply peaked rows
    qkv = _bf16(qkv)
    qkv[T:] = float("nan")
    st = torch.cuda.current_stream().cuda_stream

    def run(q, cu_np, n_tokens):
        ctx = torch.full((n_tokens, heads * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
        d_cu = torch.from_numpy(cu_np).cuda()
        lens_ = np.diff(cu_np)
        native.check(lib.drag_debug_attention(0, variant, q.data_ptr(), ctx.data_ptr(), d_cu.data_ptr(), len(lens_), n_tokens,
                                              int(lens_.max()), heads, st))
        torch.cuda.synchronize()
        return ctx

    whole = run(qkv, cu, T + tail)
    got = whole[:T].float()
    assert torch.isfinite(got).all()
    for i, n in enumerate(lens):
        blk = qkv[cu[i]:cu[i + 1]].float().view(n, 3, heads, hd)
        q, k, v = (blk[:, j].permute(1, 0, 2) for j in range(3))
        p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(hd), dim=-1)
        ref = (p @ v).permute(1, 0, 2).reshape(n, heads * hd)
        err = (got[cu[i]:cu[i + 1]] - ref).abs().max().item()
        assert err <= 0.06, (variant, n, err)
        # the same sequence alone, again followed by NaN rows
        alone = torch.cat((qkv[cu[i]:cu[i + 1]], qkv[T:T + 130]))
        one = run(alone, np.array([0, n], dtype=np.int32), n + 130)
        assert torch.equal(one[:n], whole[cu[i]:cu[i + 1]]), (variant, n)


@pytest.fixture(scope="module", params=["hf_init", "stress", "outlier"])
def model(request):
    from dial_rag_b200.embeddings.encoder import B200Encoder

    style = request.param
    seed = {"hf_init": 0, "stress": 7, "outlier": 11}[style]
    w = oenc.synth_weights(seed=seed, style=style)
    enc = B200Encoder(w, device=0, max_tokens=16384)
    yield style, w, enc
    enc.close()


def _golden(style):
    z = np.load(os.path.join(GOLDEN, "encoder_bge.npz"))
    lens = z[f"{style}:lens"]
    toks = z[f"{style}:tokens"]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    return toks.astype(np.int32), cu, z[f"{style}:embeddings"]


def _cos(a, b):
    return (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1))


def test_layer_taps_vs_oracle(model):
    style, w, enc = model
    ids, cu = synth_token_batch(seed=31, n_seq=5, seq_len=200, ragged=True, min_len=5)
    lists = oenc.packed_to_lists(ids, cu)
    for layer in (0, 1, 2, 12):
        got = enc.debug_hidden(ids, cu, layer)
        shape = oenc.BertShape(layers=layer) if layer else None
        for i, t in enumerate(lists):
            tt = torch.tensor([t])
            if layer == 0:
                ref = oenc.bert_hidden(w, tt, torch.ones_like(tt), oenc.BertShape(layers=0))[0].numpy()
            else:
                ref = oenc.bert_hidden(w, tt, torch.ones_like(tt), shape)[0].numpy()
            g = got[cu[i]:cu[i + 1]]
            c = _cos(g, ref)
            # per-TOKEN cosine of a hidden state.  The outlier weights put five channels at +-100, where one bf16 ulp of the
            # raw residual stream is 0.5, and they contain ill-conditioned tokens: token 113 of sequence 4 drifts to cosine
            # 0.90 - 0.97 in the middle layers and heals to 0.996 - 0.999 at the output with EITHER attention kernel (each
            # rounds differently: profiles/r02_outlier_token_sensitivity.txt), so for that style the per-token floor is
            # loose and the mean carries the check; the pooled embeddings meet 0.9995 (test_embeddings_vs_hf_golden)
            if style == "outlier" and layer > 0:
                assert c.min() >= 0.99 and c.mean() >= 0.9995, (style, layer, i, float(c.min()), float(c.mean()))
            else:
                assert c.min() >= (0.99999 if layer == 0 else 0.999), (style, layer, i, float(c.min()))
            # bf16 activations through `layer` layers; the stress weights (6x larger Q/K, random LN affine) put the
            # noise of the deepest tap right at 0.15, so that one gets headroom -- the contract metric is the cosine
            # (the outlier weights carry five channels of +-100, where one bf16 ulp is 0.5 and every layer rounds the raw
            # residual stream twice: their bound is relative to the CHANNEL's scale -- measured drift 0.3 % at the
            # embedding tap, 1 % after layer 1, 2.5 % after layer 2 (scripts/outlier_debug.py); the consumers discount
            # these channels, the pooled embedding stays at cosine 0.99998)
            slack = 0.05 * np.abs(ref).max(0, keepdims=True) if style == "outlier" else 0.0
            assert (np.abs(g - ref) <= {0: 0.03, 12: 0.25}.get(layer, 0.15) + slack).all(), (style, layer, i)


def test_embeddings_vs_hf_golden(model):
    """cosine >= 0.9995 against HF BertModel fp32 outputs (tests/golden/encoder_bge.npz)."""
    style, w, enc = model
    ids, cu, want = _golden(style)
    got = enc.embed_packed(ids, cu)
    assert got.shape == want.shape and got.dtype == np.float32
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    c = _cos(got, want)
    assert c.min() >= COS_BAR, (style, c)


def test_batch_composition_invariance(model):
    """An embedding does not depend on what else is in the packed batch."""
    style, w, enc = model
    ids, cu, _ = _golden(style)
    whole = enc.embed_packed(ids, cu)
    for i in (0, 2, 7):
        one = enc.embed_packed(ids[cu[i]:cu[i + 1]], np.array([0, cu[i + 1] - cu[i]], dtype=np.int32))
        assert np.array_equal(one[0], whole[i]), (style, i)
    order = np.arange(len(cu) - 1)[::-1]
    lists = oenc.packed_to_lists(ids, cu)
    rev = enc.embed_token_lists([lists[i] for i in order])
    assert np.array_equal(rev[::-1], whole)


def test_more_sequences_than_one_grid_dimension():
    """70 000 two-token sequences in ONE forward: the attention launch is cut at gridDim.y = 65 535 and the
    halves must meet seamlessly (compared with the same sequences embedded in two batches)."""
    import synth_weights
    from dial_rag_b200.embeddings.encoder import B200Encoder

    n = 70_000
    rng = np.random.default_rng(3)
    ids = rng.integers(1000, 30522, size=(n, 2), dtype=np.int32)
    ids[:, 0] = 101
    cu = np.arange(0, 2 * n + 1, 2, dtype=np.int32)
    enc = B200Encoder(synth_weights.synth_weights(seed=0, style="hf_init"), device=0, max_tokens=2 * n)
    try:
        whole = enc.embed_packed(ids.reshape(-1), cu)
        half = n // 2
        a = enc.embed_packed(ids[:half].reshape(-1), cu[: half + 1])
        b = enc.embed_packed(ids[half:].reshape(-1), cu[: n - half + 1])
    finally:
        enc.close()
    assert np.isfinite(whole).all()
    assert np.array_equal(whole[:half], a) and np.array_equal(whole[half:], b)
    # identical inputs give identical rows wherever they sit in the batch
    same = np.flatnonzero((ids[:, 1] == ids[0, 1]))
    assert all(np.array_equal(whole[i], whole[0]) for i in same)


def test_config2_shape_vs_oracle(model):
    """BASELINE config 2 shape (256-token chunks, all real tokens): oracle on a 64-chunk prefix."""
    style, w, enc = model
    ids, cu = synth_token_batch(seed=1, n_seq=64, seq_len=256)
    got = enc.embed_packed(ids, cu)
    want = oenc.encode_token_lists(w, oenc.packed_to_lists(ids, cu))
    assert _cos(got, want).min() >= COS_BAR


def test_splits_large_batches_and_truncates(model):
    style, w, enc = model
    ids, cu = synth_token_batch(seed=2, n_seq=100, seq_len=256)  # 25600 tokens > max_tokens=16384
    got = enc.embed_packed(ids, cu)
    # bitwise equal whatever the batch is split into, as long as the feed-forward block runs in the same kernel class
    # (batches of >= 12288 tokens: the fused kernel; below: the two GEMM kernels) ...
    first = enc.embed_packed(ids[: cu[50]], cu[:51])     # 12 800 tokens; `got` went through as 16 384 + 9 216 tokens
    assert np.array_equal(got[:50], first)
    # ... and equal to rounding noise across the two classes
    small = enc.embed_packed(ids[: cu[10]], cu[:11])
    assert _cos(small, got[:10]).min() >= 0.9999     # two bf16 pipelines with independent rounding: 1e-5 .. 5e-5 apart
    assert np.abs(small - got[:10]).max() <= 1e-2    # the stress weights leave one component near 0.9: a bf16 ulp there is 4e-3
    long = [101] + list(range(1000, 1700)) + [102]  # 702 tokens -> truncated to 512 keeping [SEP]
    emb = enc.embed_token_lists([long])
    want = oenc.encode_token_lists(w, [long])
    assert _cos(emb, want).min() >= COS_BAR


def test_invalid_inputs_raise(model):
    from dial_rag_b200._native import DragError

    style, w, enc = model
    with pytest.raises(DragError):
        enc.embed_packed(np.array([101, 102], dtype=np.int32), np.array([0, 0, 2], dtype=np.int32))  # empty sequence
    with pytest.raises(ValueError):
        enc.embed_token_lists([[]])
