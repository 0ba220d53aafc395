// Fused multi-head self-attention for the packed (padding-free) batch, head_dim = 32
// (SURVEY.md 8a row a5:  softmax(Q K^T / sqrt(32) + key mask) V ).
//
// One CTA per (sequence, head, 256-query tile).  K and V of the whole sequence (<= 512 x 32 bf16 each) are
// staged once in shared memory with cp.async (rows padded to 80 bytes: conflict-free fragment
// reads); each warp then owns 32 query rows at a time (two m16 tiles sharing every K/V fragment
// load) and runs an online-softmax loop over 32-key blocks with m16n8k16 bf16 tensor-core MMAs,
// fp32 scores/accumulators.  The row reference is raised lazily (only when a score exceeds it by
// 2^8), the row sum is a ninth MMA column of ones, so the per-score instruction stream is
// fma + ex2 + half a pack + half a max.  Keys beyond the sequence length never exist in the
// packed layout, so the "attention mask" is the loop bound.
//
// Measured pipe rates on B200 (scripts/ubench/pipes.cu): MUFU ex2 16/clk/SM, mma.sync m16n8k16
// 2048 flop/clk/SM: at head_dim 32 the exponentials and the legacy tensor pipe both bound this op at
// ~0.18 ms per 1024 x 256 tokens; see DESIGN.md for the tcgen05 variants.
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

namespace drag {
namespace attn {

constexpr int HEAD_DIM = 32;
constexpr int KV_STRIDE = 40;  // bf16 elements per padded smem row (80 bytes)
constexpr int WARPS = 4;
constexpr int ROWS_PER_CTA = 256;  // query rows per CTA: up to two 32-row passes per warp

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, bool valid) {
  uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(sz) : "memory");
}

// 2^x for x <= 0 (softmax numerators): one MUFU.EX2, flush-to-zero, no range fix-up code
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

constexpr uint32_t ONES_BF16X2 = 0x3f803f80u;
// The running reference of a row is only raised when some score of the block exceeds it by more than
// this many powers of two (the softmax is invariant to the reference; numerators then stay <= 2^8,
// exact in bf16's exponent range and in the fp32 accumulators).
constexpr float LAZY_LOG2 = 8.f;

// S = Q K^T of one 32-key block for the MT x 16 query rows a warp owns (MT = 2: every K fragment is
// loaded once for both row tiles)
template <int MT>
__device__ __forceinline__ void score_block(const __nv_bfloat16* ks, int key0, const uint32_t (&qa)[MT][2][4], int lane,
                                            float (&s)[MT][4][4]) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    // K fragments of 8 keys x 32 dims: four 8x8 blocks (d 0-7, 8-15, 16-23, 24-31) in one ldmatrix
    uint32_t kf[4];
    ldmatrix_x4(kf, ks + (size_t)(key0 + nt * 8 + (lane & 7)) * KV_STRIDE + (lane >> 3) * 8);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
      mma_bf16_16816(s[mt][nt], qa[mt][0], kf[0], kf[1]);
      mma_bf16_16816(s[mt][nt], qa[mt][1], kf[2], kf[3]);
    }
  }
}

// Softmax numerators of one scored block and O += P V, L += P 1.  o[mt][0..3] accumulate P V,
// o[mt][4] accumulates P 1 (the softmax denominator comes out of the tensor core over exactly the
// bf16 weights of the numerator).
template <int MT, bool MASKED>
__device__ __forceinline__ void attend_block(const __nv_bfloat16* vs, int key0, int S, float (&s)[MT][4][4], float scale_log2,
                                             float lazy_raw, int lane, float (&m)[MT][2], float (&o)[MT][5][4]) {
  const int t = lane & 3;
  // keys past the end of the sequence exist only in the last block
  if (MASKED) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int key = key0 + nt * 8 + 2 * t;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (key >= S) { s[mt][nt][0] = -INFINITY; s[mt][nt][2] = -INFINITY; }
        if (key + 1 >= S) { s[mt][nt][1] = -INFINITY; s[mt][nt][3] = -INFINITY; }
      }
    }
  }
  float lm[MT][2];
  bool raise = false;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    lm[mt][0] = max3(fmaxf(s[mt][0][0], s[mt][0][1]), s[mt][1][0], s[mt][1][1]);
    lm[mt][0] = max3(max3(lm[mt][0], s[mt][2][0], s[mt][2][1]), s[mt][3][0], s[mt][3][1]);
    lm[mt][1] = max3(fmaxf(s[mt][0][2], s[mt][0][3]), s[mt][1][2], s[mt][1][3]);
    lm[mt][1] = max3(max3(lm[mt][1], s[mt][2][2], s[mt][2][3]), s[mt][3][2], s[mt][3][3]);
    raise = raise || lm[mt][0] > m[mt][0] + lazy_raw || lm[mt][1] > m[mt][1] + lazy_raw;
  }
  if (__any_sync(0xffffffffu, raise)) {
    // (key 0 of every sequence is valid, so the reference is finite from the first block on)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v = lm[mt][h];
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
        // the decision is per ROW (v is the row's block maximum in all four lanes of its quad), so a row's
        // sequence of references -- and with it every bit of its result -- depends on that row's scores only,
        // not on which rows share its warp (tests/test_encoder_gpu.py::test_batch_composition_invariance)
        const float mn = v > m[mt][h] + lazy_raw ? v : m[mt][h];
        const float corr = fast_exp2((m[mt][h] - mn) * scale_log2);
        m[mt][h] = mn;
#pragma unroll
        for (int dt = 0; dt < 5; ++dt) { o[mt][dt][2 * h] *= corr; o[mt][dt][2 * h + 1] *= corr; }
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const float off_lo = m[mt][0] * scale_log2, off_hi = m[mt][1] * scale_log2;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      s[mt][nt][0] = fast_exp2(fmaf(s[mt][nt][0], scale_log2, -off_lo));
      s[mt][nt][1] = fast_exp2(fmaf(s[mt][nt][1], scale_log2, -off_lo));
      s[mt][nt][2] = fast_exp2(fmaf(s[mt][nt][2], scale_log2, -off_hi));
      s[mt][nt][3] = fast_exp2(fmaf(s[mt][nt][3], scale_log2, -off_hi));
    }
  }
  // P fragments come straight from the score accumulators
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    const int mtx = lane >> 3, r = lane & 7;
    const __nv_bfloat16* vrow = vs + (size_t)(key0 + kk * 16 + (mtx & 1) * 8 + r) * KV_STRIDE + (mtx >> 1) * 8;
    uint32_t vb[2][4];
    ldmatrix_x4_trans(vb[0], vrow);
    ldmatrix_x4_trans(vb[1], vrow + 16);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      uint32_t pa[4];
      pa[0] = pack2(s[mt][2 * kk][0], s[mt][2 * kk][1]);
      pa[1] = pack2(s[mt][2 * kk][2], s[mt][2 * kk][3]);
      pa[2] = pack2(s[mt][2 * kk + 1][0], s[mt][2 * kk + 1][1]);
      pa[3] = pack2(s[mt][2 * kk + 1][2], s[mt][2 * kk + 1][3]);
      mma_bf16_16816(o[mt][0], pa, vb[0][0], vb[0][1]);
      mma_bf16_16816(o[mt][1], pa, vb[0][2], vb[0][3]);
      mma_bf16_16816(o[mt][2], pa, vb[1][0], vb[1][1]);
      mma_bf16_16816(o[mt][3], pa, vb[1][2], vb[1][3]);
      mma_bf16_16816(o[mt][4], pa, ONES_BF16X2, ONES_BF16X2);
    }
  }
}

// all key blocks for MT x 16 query rows starting at q0, then O / L -> ctx.  Software pipelined: the
// Q K^T MMAs of block i+1 are issued before the exponentials of block i, so the tensor pipe and the
// MUFU pipe work at the same time inside one warp.
template <int MT>
__device__ __forceinline__ void attend_rows(const __nv_bfloat16* ks, const __nv_bfloat16* vs, const __nv_bfloat16* qs,
                                            __nv_bfloat16* ctx_base, int row0, int q0, int S, int hidden, float scale_log2,
                                            int lane) {
  const int g = lane >> 2, t = lane & 3;
  // Q fragments (A operand), 2 k-steps of 16 over head_dim, from the staged Q tile (rows >= S are zero)
  uint32_t qa[MT][2][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const __nv_bfloat16* qrow = qs + (size_t)(q0 - row0 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * KV_STRIDE + (lane >> 4) * 8;
    ldmatrix_x4(qa[mt][0], qrow);
    ldmatrix_x4(qa[mt][1], qrow + 16);
  }
  float m[MT][2], o[MT][5][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    m[mt][0] = m[mt][1] = -INFINITY;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[mt][i][j] = 0.f;
  }
  const float lazy_raw = LAZY_LOG2 / scale_log2;
  const int n_blocks = (S + 31) >> 5;
  float sa[MT][4][4], sb[MT][4][4];
  score_block<MT>(ks, 0, qa, lane, sa);
  int kb = 0;
  // two blocks per trip (the score buffers ping-pong without register moves); every block in here is full
#pragma unroll 1
  for (; kb + 2 < n_blocks; kb += 2) {
    score_block<MT>(ks, (kb + 1) * 32, qa, lane, sb);
    attend_block<MT, false>(vs, kb * 32, S, sa, scale_log2, lazy_raw, lane, m, o);
    score_block<MT>(ks, (kb + 2) * 32, qa, lane, sa);
    attend_block<MT, false>(vs, (kb + 1) * 32, S, sb, scale_log2, lazy_raw, lane, m, o);
  }
  // one or two blocks remain; only the very last one can be partial
  if (kb + 2 == n_blocks) {
    score_block<MT>(ks, (kb + 1) * 32, qa, lane, sb);
    attend_block<MT, false>(vs, kb * 32, S, sa, scale_log2, lazy_raw, lane, m, o);
    if (S & 31) attend_block<MT, true>(vs, (kb + 1) * 32, S, sb, scale_log2, lazy_raw, lane, m, o);
    else attend_block<MT, false>(vs, (kb + 1) * 32, S, sb, scale_log2, lazy_raw, lane, m, o);
  } else {
    if (S & 31) attend_block<MT, true>(vs, kb * 32, S, sa, scale_log2, lazy_raw, lane, m, o);
    else attend_block<MT, false>(vs, kb * 32, S, sa, scale_log2, lazy_raw, lane, m, o);
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r_lo = q0 + mt * 16 + g, r_hi = r_lo + 8;
    // every column of the ones tile holds the row sum
    const float inv_lo = 1.f / o[mt][4][0], inv_hi = 1.f / o[mt][4][2];
    __nv_bfloat16* out_lo = ctx_base + (size_t)r_lo * hidden + 2 * t;
    __nv_bfloat16* out_hi = ctx_base + (size_t)r_hi * hidden + 2 * t;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      if (r_lo < S) *reinterpret_cast<uint32_t*>(out_lo + dt * 8) = pack2(o[mt][dt][0] * inv_lo, o[mt][dt][1] * inv_lo);
      if (r_hi < S) *reinterpret_cast<uint32_t*>(out_hi + dt * 8) = pack2(o[mt][dt][2] * inv_hi, o[mt][dt][3] * inv_hi);
    }
  }
}

// qkv : [T, 3*hidden] bf16, per token [Q(hidden) | K(hidden) | V(hidden)], head h at columns h*32
// ctx : [T, hidden] bf16
// grid = (heads * ceil(max_len / rows_per_cta), n_seq): blockIdx.x = q_tile * heads + head; block = WARPS*32;
// rows_per_cta = 256 (throughput: K/V staged once per 256 queries) or 128 / 64 (few sequences: more CTAs);
// dynamic smem = smem_bytes(max_len)
__host__ __device__ inline size_t smem_bytes(int max_len) {
  const int s_pad = (max_len + 31) & ~31;
  const int q_pad = s_pad < ROWS_PER_CTA ? s_pad : ROWS_PER_CTA;
  return (size_t)(2 * s_pad + q_pad) * KV_STRIDE * 2;
}

__global__ void __launch_bounds__(WARPS * 32, 3)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ ctx,
                 const int* __restrict__ cu_seqlens, int heads, float scale_log2, int rows_per_cta, int len_lo, int len_hi) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int head = blockIdx.x % heads;
  const int q_tile = blockIdx.x / heads;
  const int seq = blockIdx.y;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // programmatic dependent launch: the QKV GEMM's outputs are visible from here on
  const int tok0 = cu_seqlens[seq];
  const int S = cu_seqlens[seq + 1] - tok0;
  const int row0 = q_tile * rows_per_cta;
  if (row0 >= S || S <= len_lo || S > len_hi) return;   // (len_lo, len_hi]: the length class this launch handles
  const int rows = min(S - row0, rows_per_cta);
  const int hidden = heads * HEAD_DIM;
  const int s_pad = (S + 31) & ~31;
  const int q_pad = (rows + 31) & ~31;
  __nv_bfloat16* ks = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* vs = ks + (size_t)s_pad * KV_STRIDE;
  __nv_bfloat16* qs = vs + (size_t)s_pad * KV_STRIDE;
  const int ld = 3 * hidden;
  const __nv_bfloat16* q_base = qkv + (size_t)tok0 * ld + head * HEAD_DIM;

  // stage K, V (whole sequence) and this CTA's Q rows: 4 x 16-byte chunks per row; rows >= S are zero filled
  for (int i = threadIdx.x; i < s_pad * 8; i += WARPS * 32) {
    const int row = i >> 3, part = i & 7;
    const bool valid = row < S;
    const __nv_bfloat16* src = q_base + (part >= 4 ? 2 * hidden : hidden) + (size_t)(valid ? row : 0) * ld + (part & 3) * 8;
    __nv_bfloat16* dst = (part >= 4 ? vs : ks) + (size_t)row * KV_STRIDE + (part & 3) * 8;
    cp_async_16(dst, src, valid);
  }
  for (int i = threadIdx.x; i < q_pad * 4; i += WARPS * 32) {
    const int row = i >> 2, chunk = i & 3;
    const bool valid = row0 + row < S;
    cp_async_16(qs + (size_t)row * KV_STRIDE + chunk * 8, q_base + (size_t)(valid ? row0 + row : 0) * ld + chunk * 8, valid);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* ctx_base = ctx + (size_t)tok0 * hidden + head * HEAD_DIM;
  // the CTA's 16-row tiles are dealt to the warps in contiguous runs; a warp takes its run two tiles
  // at a time (shared K/V fragments), then a single one if the run is odd
  const int n16 = (rows + 15) >> 4;
  const int per_warp = (n16 + WARPS - 1) / WARPS;
  int tile = warp * per_warp;
  const int tile_end = min(tile + per_warp, n16);
  for (; tile + 2 <= tile_end; tile += 2)
    attend_rows<2>(ks, vs, qs, ctx_base, row0, row0 + tile * 16, S, hidden, scale_log2, lane);
  if (tile < tile_end) attend_rows<1>(ks, vs, qs, ctx_base, row0, row0 + tile * 16, S, hidden, scale_log2, lane);
}

}  // namespace attn
}  // namespace drag
