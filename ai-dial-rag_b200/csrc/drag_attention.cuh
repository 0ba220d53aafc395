// Fused multi-head self-attention for the packed (padding-free) batch, head_dim = 32
// (SURVEY.md 8a row a5:  softmax(Q K^T / sqrt(32) + key mask) V ).
//
// One CTA per (sequence, head).  K and V of the whole sequence (<= 512 x 32 bf16 each) are
// staged once in shared memory with cp.async (rows padded to 80 bytes: conflict-free fragment
// reads); each warp then owns 16 query rows at a time and runs an online-softmax loop over
// 64-key blocks with m16n8k16 bf16 tensor-core MMAs, fp32 scores/accumulators.  Keys beyond the
// sequence length never exist in the packed layout, so the "attention mask" is the loop bound.
//
// NOTE: this kernel uses warp-level mma.sync; moving QK^T / PV to tcgen05 with the
// score tile in TMEM is the next step for this kernel.
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

namespace drag {
namespace attn {

constexpr int HEAD_DIM = 32;
constexpr int KV_STRIDE = 40;  // bf16 elements per padded smem row (80 bytes)
constexpr int WARPS = 8;

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

// One 64-key block of the online softmax for the 16 query rows a warp owns.
template <bool MASKED>
__device__ __forceinline__ void attend_block(const __nv_bfloat16* ks, const __nv_bfloat16* vs, int key0, int S,
                                             const uint32_t (&qa)[2][4], float scale_log2, int lane, float& m_lo,
                                             float& m_hi, float& l_lo, float& l_hi, float (&o)[4][4]) {
  const int g = lane >> 2, t = lane & 3;
      float s[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    // K fragments of 8 keys x 32 dims: four 8x8 blocks (d 0-7, 8-15, 16-23, 24-31) in one ldmatrix
    uint32_t kf[4];
    ldmatrix_x4(kf, ks + (size_t)(key0 + nt * 8 + (lane & 7)) * KV_STRIDE + (lane >> 3) * 8);
    mma_bf16_16816(s[nt], qa[0], kf[0], kf[1]);
    mma_bf16_16816(s[nt], qa[1], kf[2], kf[3]);
  }
  // keys past the end of the sequence exist only in the (peeled) last block
  if (MASKED) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = key0 + nt * 8 + 2 * t;
      if (key >= S) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (key + 1 >= S) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
    }
  }
  float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    mx_lo = fmaxf(mx_lo, fmaxf(s[nt][0], s[nt][1]));
    mx_hi = fmaxf(mx_hi, fmaxf(s[nt][2], s[nt][3]));
  }
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
  // key 0 of every sequence is valid, so the running max is finite from the first block on
  const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
  const float corr_lo = fast_exp2((m_lo - mn_lo) * scale_log2), corr_hi = fast_exp2((m_hi - mn_hi) * scale_log2);
  m_lo = mn_lo;
  m_hi = mn_hi;
  const float off_lo = mn_lo * scale_log2, off_hi = mn_hi * scale_log2;
  float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    s[nt][0] = fast_exp2(fmaf(s[nt][0], scale_log2, -off_lo));
    s[nt][1] = fast_exp2(fmaf(s[nt][1], scale_log2, -off_lo));
    s[nt][2] = fast_exp2(fmaf(s[nt][2], scale_log2, -off_hi));
    s[nt][3] = fast_exp2(fmaf(s[nt][3], scale_log2, -off_hi));
    sum_lo += s[nt][0] + s[nt][1];
    sum_hi += s[nt][2] + s[nt][3];
  }
  l_lo = l_lo * corr_lo + sum_lo;
  l_hi = l_hi * corr_hi + sum_hi;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    o[dt][0] *= corr_lo; o[dt][1] *= corr_lo;
    o[dt][2] *= corr_hi; o[dt][3] *= corr_hi;
  }
  // O += P V : P fragments come straight from the score accumulators
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t pa[4];
    pa[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    const int mtx = lane >> 3, r = lane & 7;
    const __nv_bfloat16* vrow = vs + (size_t)(key0 + kk * 16 + (mtx & 1) * 8 + r) * KV_STRIDE + (mtx >> 1) * 8;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh) {
      uint32_t vb[4];
      ldmatrix_x4_trans(vb, vrow + dh * 16);
      mma_bf16_16816(o[dh * 2], pa, vb[0], vb[1]);
      mma_bf16_16816(o[dh * 2 + 1], pa, vb[2], vb[3]);
    }
  }
}

// qkv : [T, 3*hidden] bf16, per token [Q(hidden) | K(hidden) | V(hidden)], head h at columns h*32
// ctx : [T, hidden] bf16
// grid = (heads, n_seq), block = WARPS*32, dynamic smem = 2 * round_up(max_len, 64) * 80 bytes
__global__ void __launch_bounds__(WARPS * 32, 3)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ ctx,
                 const int* __restrict__ cu_seqlens, int hidden, float scale_log2) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int head = blockIdx.x;
  const int seq = blockIdx.y;
  const int tok0 = cu_seqlens[seq];
  const int S = cu_seqlens[seq + 1] - tok0;
  const int s_pad = (S + 63) & ~63;
  __nv_bfloat16* ks = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* vs = ks + (size_t)s_pad * KV_STRIDE;
  const int ld = 3 * hidden;
  const __nv_bfloat16* q_base = qkv + (size_t)tok0 * ld + head * HEAD_DIM;
  const __nv_bfloat16* k_base = q_base + hidden;
  const __nv_bfloat16* v_base = q_base + 2 * hidden;

  // stage K and V: 4 x 16-byte chunks per row each; rows >= S are zero filled
  for (int i = threadIdx.x; i < s_pad * 8; i += WARPS * 32) {
    const int row = i >> 3, part = i & 7;
    const bool is_v = part >= 4;
    const int chunk = part & 3;
    const bool valid = row < S;
    const __nv_bfloat16* src = (is_v ? v_base : k_base) + (size_t)(valid ? row : 0) * ld + chunk * 8;
    __nv_bfloat16* dst = (is_v ? vs : ks) + (size_t)row * KV_STRIDE + chunk * 8;
    cp_async_16(dst, src, valid);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_kblocks = s_pad >> 6;

  for (int q0 = warp * 16; q0 < S; q0 += WARPS * 16) {
    // Q fragments (A operand), 2 k-steps of 16 over head_dim
    uint32_t qa[2][4];
    const int r_lo = q0 + g, r_hi = q0 + g + 8;
#pragma unroll
    for (int ks2 = 0; ks2 < 2; ++ks2) {
      const int c = ks2 * 16 + 2 * t;
      qa[ks2][0] = r_lo < S ? *reinterpret_cast<const uint32_t*>(q_base + (size_t)r_lo * ld + c) : 0u;
      qa[ks2][1] = r_hi < S ? *reinterpret_cast<const uint32_t*>(q_base + (size_t)r_hi * ld + c) : 0u;
      qa[ks2][2] = r_lo < S ? *reinterpret_cast<const uint32_t*>(q_base + (size_t)r_lo * ld + c + 8) : 0u;
      qa[ks2][3] = r_hi < S ? *reinterpret_cast<const uint32_t*>(q_base + (size_t)r_hi * ld + c + 8) : 0u;
    }
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = 0.f;

    const int n_full = S >> 6;  // blocks with 64 valid keys
    for (int kb = 0; kb < n_full; ++kb)
      attend_block<false>(ks, vs, kb * 64, S, qa, scale_log2, lane, m_lo, m_hi, l_lo, l_hi, o);
    if (n_full < n_kblocks)
      attend_block<true>(ks, vs, n_full * 64, S, qa, scale_log2, lane, m_lo, m_hi, l_lo, l_hi, o);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const float inv_lo = 1.f / l_lo, inv_hi = 1.f / l_hi;
    __nv_bfloat16* out_lo = ctx + (size_t)(tok0 + r_lo) * hidden + head * HEAD_DIM + 2 * t;
    __nv_bfloat16* out_hi = ctx + (size_t)(tok0 + r_hi) * hidden + head * HEAD_DIM + 2 * t;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      if (r_lo < S) *reinterpret_cast<uint32_t*>(out_lo + dt * 8) = pack2(o[dt][0] * inv_lo, o[dt][1] * inv_lo);
      if (r_hi < S) *reinterpret_cast<uint32_t*>(out_hi + dt * 8) = pack2(o[dt][2] * inv_hi, o[dt][3] * inv_hi);
    }
  }
}

}  // namespace attn
}  // namespace drag
