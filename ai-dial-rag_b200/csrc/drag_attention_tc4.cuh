// Multi-head self-attention on tcgen05 for packed sequences of up to 256 tokens, head_dim = 32.
//
// What the persistent kernel (drag_attention_tc3.cuh) measured: at head_dim 32 the op is bound by the exponentials, a warp
// gets a MUFU.EX2 issued only every ~18 clocks, and one / two / four warps per scheduler reach 6.9 / 10.7-12.2 / 14.4-15.0 of the
// pipe's 16 exponentials per clock and SM.  A thread = row design with 128 score registers fits two warps per scheduler.
// This kernel gets to FOUR by making the CTA small instead of clever: one CTA = one (sequence, head, 128-query tile) job,
// 64-key blocks (64 score registers), 128 tensor-memory columns, 40 KB of shared memory, 64 registers per thread on
// average -- four CTAs per SM, each at a different point of its job, so the fixed latencies of one (barrier waits,
// tensor-memory loads, the job's prologue and epilogue) run under the exponentials of the other three, and the hardware
// block scheduler does the load balancing the persistent kernel needed a descriptor ring for.
//
// CTA = 8 warps:  warps 0-3 softmax (thread = query row; 104 registers) | warp 4 TMA producer | warp 5 MMA issuer |
//                 warps 6-7 idle (a warpgroup has four warps; 24 registers)
// TMEM (128 columns): S [0,64) scores of the next block | P [64,96) bf16 weights, two per column | O [96,128)
// Per 64-key block n:  S(n) = Q K^T (M=128, N=64, K=16 x2)  ->  softmax thread: one tcgen05.ld pass, row maximum, lazily
// raised reference, p = 2^(s c - m c), bf16 pairs -> tensor memory  ->  O (+)= P(n) V (M=128, N=32, K=16 x4, A from tensor
// memory, V as stored: MN-major B).  S(n+1) is issued before P(n) V, as in the persistent kernel.
#pragma once

#include <cuda_bf16.h>

#include "drag_attention_tc3.cuh"

namespace drag {
namespace attn4 {

using namespace attn3;   // descriptors, packed-pair arithmetic, ex2 / max3 / round_bf16x2, umma_bf16_ts

constexpr int KEYS = 64;                                 // keys per block
constexpr int MAX_LEN = 256;
constexpr int MAX_TILES = MAX_LEN / TILE;                // K / V tiles of 128 tokens
constexpr int THREADS4 = 256;
constexpr int REGS_SOFTMAX = 104, REGS_OTHER = 24;       // 128 x 104 + 128 x 24 = 256 x 64
constexpr int TMEM_COLS4 = 128;
constexpr int S_COL = 0, P_COL4 = KEYS, O_COL4 = KEYS + KEYS / 2;
constexpr int BAR_BYTES4 = 64;

__host__ __device__ inline size_t smem_bytes4(int max_len) {
  const int tiles = (max_len + TILE - 1) / TILE;
  return (size_t)(1 + 2 * tiles) * QKV_TILE_BYTES + BAR_BYTES4 + 1024 /*alignment*/;
}

// qkv : [T, 3*hidden] bf16 (tensor map: box 32 columns x 128 rows, 64-byte swizzle);  ctx : [T, hidden] bf16
// grid = (heads * ceil(max_len / 128), n_seq), block = 256, dynamic smem = smem_bytes4(longest sequence <= 256);
// only sequences with len_lo < length <= len_hi are processed
__global__ void __launch_bounds__(THREADS4, 4)
attention_tc4_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ ctx,
                     const int* __restrict__ cu_seqlens, int heads, float scale_log2, int len_lo, int len_hi, int stagger_sms, unsigned stagger_ns) {
  extern __shared__ uint8_t smem_raw[];
  const int head = blockIdx.x % heads;
  const int qt = blockIdx.x / heads;
  const int seq = blockIdx.y;
  const int tok0 = __ldg(cu_seqlens + seq);          // (an input of the forward, not written by any of its kernels)
  const int S = __ldg(cu_seqlens + seq + 1) - tok0;
  if (qt * TILE >= S || S <= len_lo || S > len_hi) return;
  // The four CTAs of an SM start together and, all jobs being alike, would stay in lockstep for the whole launch -- all in
  // their prologue, then all in their exponentials (measured: 45 % MUFU, the same as one CTA per SM).  The first wave is
  // staggered once; every slot then finishes and refills at its own time.
  {
    const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x, sms = (unsigned)stagger_sms;
    if (lin < 4u * sms && lin >= sms) __nanosleep((lin / sms) * stagger_ns);
  }
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int hidden = heads * HEAD_DIM;
  const int n_tiles = (S + TILE - 1) / TILE;          // K / V tiles
  const int n_blocks = (S + KEYS - 1) / KEYS;
  uint8_t* q_s = smem;                                 // [128][32] bf16
  uint8_t* k_s = smem + QKV_TILE_BYTES;                // n_tiles x [128][32]
  uint8_t* v_s = k_s + (size_t)n_tiles * QKV_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + (size_t)n_tiles * QKV_TILE_BYTES);
  uint64_t* kv_full = bars;       // TMA -> MMA
  uint64_t* s_full = bars + 1;    // MMA -> softmax: phase n = S(n) complete (and every MMA issued before it)
  uint64_t* s_free = bars + 2;    // softmax -> MMA: phase n = S(n) is in registers
  uint64_t* p_full = bars + 3;    // softmax -> MMA: phase n = P(n) is in tensor memory
  uint64_t* o_full = bars + 4;    // MMA -> softmax: phase n = P(n) V has retired
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 4 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_qkv);
    tc::mbar_init(kv_full, 1);
    tc::mbar_init(s_full, 1);
    tc::mbar_init(s_free, 4);
    tc::mbar_init(p_full, 4);
    tc::mbar_init(o_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 5) {
    tc::tmem_alloc(tmem_ptr_smem, TMEM_COLS4);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  tc::pdl_launch_dependents();
  tc::pdl_wait();   // programmatic dependent launch: the QKV GEMM's outputs are visible from here on

  if (warp >= 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));
    if (warp == 4) {
      // ===================== TMA producer: Q tile, K tiles, V tiles of this (sequence, head) =====================
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(kv_full, (uint32_t)((1 + 2 * n_tiles) * QKV_TILE_BYTES));
        tc::tma_load_2d(&tmap_qkv, kv_full, q_s, head * HEAD_DIM, tok0 + qt * TILE);
        for (int t = 0; t < n_tiles; ++t)
          tc::tma_load_2d(&tmap_qkv, kv_full, k_s + (size_t)t * QKV_TILE_BYTES, hidden + head * HEAD_DIM, tok0 + t * TILE);
        for (int t = 0; t < n_tiles; ++t)
          tc::tma_load_2d(&tmap_qkv, kv_full, v_s + (size_t)t * QKV_TILE_BYTES, 2 * hidden + head * HEAD_DIM, tok0 + t * TILE);
      }
      __syncwarp();
    } else if (warp == 5) {
      // ===================== MMA issuer:  S(0) | { S(n+1), P(n) V } ... =====================
      constexpr uint32_t idesc_s = idesc_bf16(TILE, KEYS, false);
      constexpr uint32_t idesc_o = idesc_bf16(TILE, HEAD_DIM, true);
      const uint64_t q_desc = desc_k_sw64(tc::smem_u32(q_s));
      const uint32_t k_base = tc::smem_u32(k_s), v_base = tc::smem_u32(v_s);
      auto issue_scores = [&](int n) {
        if (n > 0) tc::mbar_wait(s_free, (uint32_t)(n - 1) & 1);
        tc::tc_fence_after();
        const uint64_t k_desc = desc_k_sw64(k_base + (uint32_t)(n * KEYS * 64));   // 64 keys = 4096 bytes down the K rows
        __syncwarp();
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < HEAD_DIM / 16; ++k)
            tc::umma_bf16(tmem_base + S_COL, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
          tc::umma_commit(s_full);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int n) {
        tc::mbar_wait(p_full, (uint32_t)n & 1);
        tc::tc_fence_after();
        const uint32_t v_blk = v_base + (uint32_t)(n * KEYS * 64);
        // Keys beyond the sequence: their weights are exact zeros, but the V rows behind them belong to other sequences or to
        // the never-written tail of the buffer and may hold anything, NaN included (0 x NaN = NaN): only the 16-key steps that
        // contain a key of the sequence are issued, and the rows of the last step beyond the sequence are zeroed first.
        const int valid = S - n * KEYS;
        const int n_steps = valid >= KEYS ? KEYS / 16 : (valid + 15) >> 4;
        if (valid < n_steps * 16) {
          const int chunks = (n_steps * 16 - valid) * 4;   // 16-byte chunks (the 64-byte swizzle permutes chunks inside a row only)
          for (int c = lane; c < chunks; c += 32)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(v_blk + (uint32_t)(valid * 64 + c * 16)), "r"(0u) : "memory");
          tc::fence_proxy_async();
        }
        const uint64_t b_desc = desc_mn_sw64(v_blk);
        __syncwarp();
        if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < KEYS / 16; ++kk)
            if (kk < n_steps)
              umma_bf16_ts(tmem_base + O_COL4, tmem_base + (uint32_t)(P_COL4 + kk * 8), b_desc + (uint64_t)(kk * (1024 >> 4)), idesc_o, (n | kk) != 0 ? 1u : 0u);
          tc::umma_commit(o_full);
        }
        __syncwarp();
      };
      tc::mbar_wait(kv_full, 0);
      issue_scores(0);
      for (int n = 0; n < n_blocks; ++n) {
        if (n + 1 < n_blocks) issue_scores(n + 1);
        issue_pv(n);
      }
    }
  } else {
    // ===================== softmax: thread = one query row =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOFTMAX));
    const int row = warp * 32 + lane;
    const uint32_t lane_tmem = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float lazy_raw = LAZY_LOG2 / scale_log2;
    const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
    float m = -INFINITY, l = 0.f;
    for (int b = 0; b < n_blocks; ++b) {
      const int valid = S - b * KEYS;              // keys of this block inside the sequence (>= 1)
      uint32_t r[KEYS];
      if (b == 0) {                                // (the scores of later blocks were waited for one block ahead)
        tc::mbar_wait(s_full, 0);
        tc::tc_fence_after();
      }
      tc::tmem_ld32(lane_tmem + S_COL, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
      tc::tmem_ld32(lane_tmem + S_COL + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(s_free);      // the next block's scores may overwrite the buffer
      if (valid < KEYS) {
#pragma unroll
        for (int i = 0; i < KEYS; ++i)
          if (i >= valid) r[i] = 0xff800000u;      // keys beyond the sequence
      }
      // row maximum: eight chains of 3-input maxima (8 x 3 + 8 x 2 x 2 + 8 = 64)
      float mx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = max3(__uint_as_float(r[3 * c]), __uint_as_float(r[3 * c + 1]), __uint_as_float(r[3 * c + 2]));
#pragma unroll
      for (int i = 24; i + 15 < KEYS; i += 16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) mx[c] = max3(mx[c], __uint_as_float(r[i + 2 * c]), __uint_as_float(r[i + 2 * c + 1]));
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) mx[c] = max3(mx[c], __uint_as_float(r[56 + 2 * c]), __uint_as_float(r[57 + 2 * c]));
      const float mb = max3(max3(mx[0], mx[1], mx[2]), max3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7]));
      // the reference: block 0 sets it (key 0 is always valid: finite); later blocks raise it lazily
      float corr = 1.f;
      bool raise = false;
      if (b == 0) {
        m = mb;
      } else if (mb > m + lazy_raw) {
        corr = ex2((m - mb) * scale_log2);
        m = mb;
        raise = true;
      }
      const float off = m * scale_log2;
      const uint64_t noff2 = pack_f32x2(-off, -off);
      // the bf16 pairs overwrite the scores they were computed from (pair i/2 <= i: already consumed): no second array
      uint32_t(&pk)[KEYS / 2] = *reinterpret_cast<uint32_t(*)[KEYS / 2]>(&r[0]);
      uint64_t l2a = pack_f32x2(0.f, 0.f), l2b = l2a;
#pragma unroll
      for (int c = 0; c < KEYS / 32; ++c) {
        if (c * 32 < valid) {   // (a partial block skips the exponentials of a 32-key chunk that holds no key of the sequence)
#pragma unroll
          for (int i = c * 32; i < c * 32 + 32; i += 4) {
            float p0, p1, p2, p3;
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), scale2, noff2), p0, p1);
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), scale2, noff2), p2, p3);
            p0 = ex2(p0); p1 = ex2(p1); p2 = ex2(p2); p3 = ex2(p3);
            l2a = add_f32x2(l2a, pack_f32x2(p0, p1));
            l2b = add_f32x2(l2b, pack_f32x2(p2, p3));
            r[i >> 1] = round_bf16x2(p0, p1);
            r[(i >> 1) + 1] = round_bf16x2(p2, p3);
          }
        } else {
#pragma unroll
          for (int i = c * 16; i < c * 16 + 16; ++i) r[i] = 0u;
        }
      }
      float la, lb, lc, ld;
      unpack_f32x2(l2a, la, lb);
      unpack_f32x2(l2b, lc, ld);
      const float lsum = (la + lb) + (lc + ld);
      // One wait before P is stored.  Inside the job: the next block's scores S(b+1) -- issued after P(b-1) V by the same
      // thread, so P(b-1) V has retired as well; at the last block: P(b-1) V itself.  Either way the P columns are free and
      // O is complete up to block b-1.
      if (b + 1 < n_blocks) tc::mbar_wait(s_full, (uint32_t)(b + 1) & 1);
      else if (b > 0) tc::mbar_wait(o_full, (uint32_t)(b - 1) & 1);
      tc::tc_fence_after();
      if (b > 0 && __any_sync(0xffffffffu, raise)) {
        // rare: a raised reference rescales the running O (complete: P(b-1) V has retired) and l
        uint32_t o[HEAD_DIM];
        tc::tmem_ld32(lane_tmem + O_COL4, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < HEAD_DIM; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
        tc::tmem_st32(lane_tmem + O_COL4, o);
        l *= corr;
      }
      l += lsum;
      tc::tmem_st32(lane_tmem + P_COL4, pk);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_full);
    }
    // O / l -> ctx (rows inside the sequence only)
    tc::mbar_wait(o_full, (uint32_t)(n_blocks - 1) & 1);
    tc::tc_fence_after();
    uint32_t o[HEAD_DIM];
    tc::tmem_ld32(lane_tmem + O_COL4, o);
    tc::tmem_ld_wait();
    const int q_row = qt * TILE + row;
    if (q_row < S) {
      const float inv = 1.f / l;
      uint4* dst = reinterpret_cast<uint4*>(ctx + (size_t)(tok0 + q_row) * hidden + head * HEAD_DIM);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dst[j] = make_uint4(pack2(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                            pack2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                            pack2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                            pack2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv));
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, TMEM_COLS4);
  }
}

}  // namespace attn4
}  // namespace drag
