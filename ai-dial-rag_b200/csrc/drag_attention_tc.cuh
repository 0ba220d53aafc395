// Multi-head self-attention on tcgen05 for the packed (padding-free) batch, head_dim = 32
// (SURVEY.md 8a row a5:  softmax(Q K^T / sqrt(32) + key mask) V ).
//
// Persistent kernel, one CTA per SM, looping over (sequence, head) items:
//   warp 8     TMA producer: Q, K and V of the next item (64-byte swizzle: a row of 32 bf16 is one
//              swizzle row) straight out of the packed [T, 3*hidden] QKV buffer, double buffered
//   warp 9     MMA issuer (one thread)
//   warps 0-3  softmax group 0 -- query tiles 0, 2, ... of the item   (thread = one query row)
//   warps 4-7  softmax group 1 -- query tiles 1, 3, ...
// Per 128-query tile and 128-key block (online softmax across blocks):
//   S[128 x 128] = Q K^T            tcgen05.mma M=128 N=128 K=16 x2 -> TMEM (one S buffer per group)
//   softmax group: tcgen05.ld the S row into registers, row max, p = 2^(s*c - m*c), row sum,
//                  P -> bf16 -> shared memory in the K-major 128B-swizzle layout (A operand)
//   PV[128 x 32] = P V              tcgen05.mma M=128 N=32 K=16 x8 -> TMEM; V is consumed as stored,
//                  [key][32 dims], i.e. an MN-major B operand
//   the group folds PV into its fp32 O registers (rescaled by 2^((m_old-m_new)c)); O / l -> ctx.
// While one group runs its exponentials the tensor core works for the other one.  The kernel is
// bound by the exponentials (MUFU: 16 per clock per SM).  Keys beyond the sequence do not exist in
// the packed layout: the tail of the last key block is masked before the row max.
#pragma once

#include <cuda_bf16.h>

#include "drag_tc.cuh"

namespace drag {
namespace attn_tc {

constexpr int HEAD_DIM = 32;
constexpr int TILE = 128;                    // queries per tile = keys per block
constexpr int QKV_TILE_BYTES = TILE * HEAD_DIM * 2;   // 8 KB: [128][32] bf16, 64-byte swizzle
constexpr int P_BYTES = 2 * TILE * 64 * 2;   // 32 KB per softmax group: two [128][64] bf16 k-blocks
constexpr int GROUPS = 2;
constexpr int THREADS = 32 * (4 * GROUPS + 2);
constexpr int TMEM_COLS = 512;               // S0 [0,128)  S1 [128,256)  group g: PV [256+64g, +32)  L [+32, +48)
constexpr int PV_COL = 256;
constexpr int PV_STRIDE = 64;
constexpr int L_OFF = 32;                    // L = P . 1: the softmax denominator comes out of the tensor core
constexpr int ONES_BYTES = 2048;             // [16 rows][64] bf16 ones: the B operand of L (any swizzle: all equal)

__host__ __device__ inline int item_stages(int max_len) {
  const int tiles = (max_len + TILE - 1) / TILE;
  return tiles <= 3 ? 2 : 1;
}
__host__ __device__ inline size_t smem_bytes(int max_len) {
  const int tiles = (max_len + TILE - 1) / TILE;
  return (size_t)item_stages(max_len) * 3 * tiles * QKV_TILE_BYTES + (size_t)GROUPS * P_BYTES + ONES_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;
}

// K-major operand, rows of 64 bytes (32 bf16) under the 64-byte swizzle: 8-row groups 512 B apart
__device__ __forceinline__ uint64_t desc_k_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO: unused for swizzled K-major
  d |= (uint64_t)(512 >> 4) << 32;        // SBO
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
  return d;
}
// MN-major operand ([k][32 contiguous n], 64-byte rows, 64-byte swizzle): one atom along N,
// groups of 8 k-rows 512 B apart (SBO); LBO (stride between N atoms) is not exercised at N = 32
__device__ __forceinline__ uint64_t desc_mn_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
  d |= (uint64_t)(512 >> 4) << 16;        // LBO
  d |= (uint64_t)(512 >> 4) << 32;        // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// query tiles of an item that belong to softmax group g
__device__ __forceinline__ int group_tiles(int n_tiles, int g) { return (n_tiles - g + 1) >> 1; }

// qkv : [T, 3*hidden] bf16 (tensor map: box 32 columns x 128 rows, 64-byte swizzle)
// ctx : [T, hidden] bf16
// grid = min(#SMs, heads * n_seq), block = THREADS, dynamic smem = smem_bytes(longest sequence)
__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ ctx,
                    const int* __restrict__ cu_seqlens, int n_seq, int heads, int max_tiles, int n_stages,
                    float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int hidden = heads * HEAD_DIM;
  const size_t stage_bytes = (size_t)3 * max_tiles * QKV_TILE_BYTES;  // [Q tiles | K tiles | V tiles]
  uint8_t* p_smem = smem + (size_t)n_stages * stage_bytes;            // multiple of 8 KB: 1024-aligned
  uint8_t* ones_smem = p_smem + (size_t)GROUPS * P_BYTES;            // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones_smem + ONES_BYTES);
  uint64_t* kv_full = bars;          // [2]  TMA -> MMA
  uint64_t* kv_empty = bars + 2;     // [2]  MMA -> TMA (all MMAs of the item retired)
  uint64_t* s_full = bars + 4;       // [GROUPS] MMA -> softmax: S complete
  uint64_t* p_full = bars + 6;       // [GROUPS] softmax -> MMA: P in shared memory, S and PV consumed
  uint64_t* o_full = bars + 8;       // [GROUPS] MMA -> softmax: PV complete (P reusable)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_items = n_seq * heads;

  if (warp == 4 * GROUPS) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmap_qkv);
      for (int s = 0; s < 2; ++s) {
        tc::mbar_init(&kv_full[s], 1);
        tc::mbar_init(&kv_empty[s], 1);
      }
      for (int g = 0; g < GROUPS; ++g) {
        tc::mbar_init(&s_full[g], 1);
        tc::mbar_init(&p_full[g], 4);
        tc::mbar_init(&o_full[g], 1);
      }
      tc::fence_barrier_init();
    }
    __syncwarp();
  }
  if (warp == 4 * GROUPS + 1) {
    tc::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tc::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < ONES_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(ones_smem)[i] = 0x3f803f80u;
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 4 * GROUPS) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int seq = item / heads, head = item % heads;
        const int tok0 = __ldg(cu_seqlens + seq);
        const int S = __ldg(cu_seqlens + seq + 1) - tok0;
        const int n_tiles = (S + TILE - 1) / TILE;
        tc::mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* base = smem + (size_t)stage * stage_bytes;
        tc::mbar_arrive_expect_tx(&kv_full[stage], (uint32_t)(3 * n_tiles * QKV_TILE_BYTES));
        for (int t = 0; t < n_tiles; ++t) {
          const int row = tok0 + t * TILE;
          tc::tma_load_2d(&tmap_qkv, &kv_full[stage], base + (size_t)t * QKV_TILE_BYTES, head * HEAD_DIM, row);
          tc::tma_load_2d(&tmap_qkv, &kv_full[stage], base + (size_t)(max_tiles + t) * QKV_TILE_BYTES, hidden + head * HEAD_DIM, row);
          tc::tma_load_2d(&tmap_qkv, &kv_full[stage], base + (size_t)(2 * max_tiles + t) * QKV_TILE_BYTES, 2 * hidden + head * HEAD_DIM, row);
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 4 * GROUPS + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(TILE, TILE, false);
      constexpr uint32_t idesc_o = idesc_bf16(TILE, HEAD_DIM, true);
      constexpr uint32_t idesc_l = idesc_bf16(TILE, 16, false);
      const uint64_t ones_desc = tc::umma_desc_sw128(tc::smem_u32(ones_smem));
      int stage = 0;
      uint32_t phase = 0;
      uint32_t p_waits[GROUPS] = {0, 0};  // completed waits on p_full[g]
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int seq = item / heads;
        const int S = __ldg(cu_seqlens + seq + 1) - __ldg(cu_seqlens + seq);
        const int n_tiles = (S + TILE - 1) / TILE;
        tc::mbar_wait(&kv_full[stage], phase);
        tc::tc_fence_after();
        const uint32_t base = tc::smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t k_base = base + (uint32_t)(max_tiles * QKV_TILE_BYTES);
        const uint32_t v_base = base + (uint32_t)(2 * max_tiles * QKV_TILE_BYTES);
        // every group walks (its query tiles) x (all key blocks); the steps of the groups alternate
        // group g walks (its query tiles) x (all key blocks); at step s it gets the P V of its iteration
        // s-1 (once the group has published P) and the Q K^T of its iteration s: the groups alternate
        const int iters0 = group_tiles(n_tiles, 0) * n_tiles;
        for (int step = 0; step <= iters0; ++step) {
#pragma unroll
          for (int g = 0; g < GROUPS; ++g) {
            const int iters = group_tiles(n_tiles, g) * n_tiles;
            if (step >= 1 && step <= iters) {
              const int kb = (step - 1) % n_tiles;
              tc::mbar_wait(&p_full[g], p_waits[g] & 1);
              ++p_waits[g];
              tc::tc_fence_after();
              const uint32_t p_addr = tc::smem_u32(p_smem + (size_t)g * P_BYTES);
              const uint32_t v_addr = v_base + (uint32_t)(kb * QKV_TILE_BYTES);
#pragma unroll
              for (int j = 0; j < TILE / 16; ++j) {
                const uint64_t a_desc = tc::umma_desc_sw128(p_addr + (uint32_t)((j >> 2) * (TILE * 128))) + (uint64_t)((j & 3) * 2);
                const uint64_t b_desc = desc_mn_sw64(v_addr + (uint32_t)(j * 16 * 64));
                tc::umma_bf16(tmem_base + PV_COL + g * PV_STRIDE, a_desc, b_desc, idesc_o, j != 0 ? 1u : 0u);
                tc::umma_bf16(tmem_base + PV_COL + g * PV_STRIDE + L_OFF, a_desc, ones_desc, idesc_l, j != 0 ? 1u : 0u);
              }
              tc::umma_commit(&o_full[g]);
            }
            if (step < iters) {
              // S is free: the group arrived on p_full (S already in its registers) before this point
              const int qt = g + GROUPS * (step / n_tiles), kb = step % n_tiles;
              const uint64_t q_desc = desc_k_sw64(base + (uint32_t)(qt * QKV_TILE_BYTES));
              const uint64_t k_desc = desc_k_sw64(k_base + (uint32_t)(kb * QKV_TILE_BYTES));
#pragma unroll
              for (int k = 0; k < HEAD_DIM / 16; ++k)
                tc::umma_bf16(tmem_base + g * TILE, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
              tc::umma_commit(&s_full[g]);
            }
          }
        }
        tc::umma_commit(&kv_empty[stage]);  // every MMA of the item has retired: the stage is reusable
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax groups: thread = query row =====================
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t s_tmem = lane_base + g * TILE;
    const uint32_t pv_tmem = lane_base + PV_COL + g * PV_STRIDE;
    const uint32_t p_row = tc::smem_u32(p_smem + (size_t)g * P_BYTES) + (uint32_t)row * 128u;
    uint32_t n_s = 0, n_o = 0;  // completed waits on s_full[g] / o_full[g]
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int seq = item / heads, head = item % heads;
      const int tok0 = __ldg(cu_seqlens + seq);
      const int S = __ldg(cu_seqlens + seq + 1) - tok0;
      const int n_tiles = (S + TILE - 1) / TILE;
      for (int qt = g; qt < n_tiles; qt += GROUPS) {
        float m = -INFINITY, l = 0.f;
        float o[HEAD_DIM];
#pragma unroll
        for (int i = 0; i < HEAD_DIM; ++i) o[i] = 0.f;
        for (int kb = 0; kb < n_tiles; ++kb) {
          const int n_keys = S - kb * TILE;  // valid keys in this block (>= 1)
          tc::mbar_wait(&s_full[g], n_s & 1);
          ++n_s;
          tc::tc_fence_after();
          // pass 1: row maximum.  S is read from TMEM twice (32 columns at a time, the next load in
          // flight during the current chunk) instead of holding 128 scores in registers.
          float mx = -INFINITY;
          {
            uint32_t ra[32], rb[32];
            tc::tmem_ld32(s_tmem, ra);
#pragma unroll
            for (int c = 0; c < TILE / 32; ++c) {
              tc::tmem_ld_wait();
              uint32_t (&cur)[32] = (c & 1) ? rb : ra;
              uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
              if (c + 1 < TILE / 32) tc::tmem_ld32(s_tmem + (c + 1) * 32, nxt);
              if (n_keys < (c + 1) * 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (c * 32 + i >= n_keys) cur[i] = 0xff800000u;
              }
#pragma unroll
              for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(cur[i]), __uint_as_float(cur[i + 1]));
            }
          }
          const float m_new = fmaxf(m, mx);
          const float alpha = ex2((m - m_new) * scale_log2);  // 0 for the first block (m = -inf)
          m = m_new;
          const float off = m_new * scale_log2;
          if (kb > 0) {
            // fold the previous block's P V into O (it was computed with the old maximum), then rescale
            tc::mbar_wait(&o_full[g], n_o & 1);
            ++n_o;
            tc::tc_fence_after();
            uint32_t pv[HEAD_DIM];
            tc::tmem_ld32(pv_tmem, pv);
            const uint32_t lv = tc::tmem_ld1(pv_tmem + L_OFF);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < HEAD_DIM; ++i) o[i] = (o[i] + __uint_as_float(pv[i])) * alpha;
            l = (l + __uint_as_float(lv)) * alpha;
          }
          // exponentials; P -> bf16 -> smem.  The row sum is NOT taken here: L = P . 1 comes out of the
          // tensor core next to P V, over exactly the bf16 weights the numerator uses.
          {
            uint32_t ra[32], rb[32];
            tc::tmem_ld32(s_tmem, ra);
#pragma unroll
            for (int c = 0; c < TILE / 32; ++c) {
              const int c0 = c * 32;
              tc::tmem_ld_wait();
              uint32_t (&cur)[32] = (c & 1) ? rb : ra;
              uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
              if (c + 1 < TILE / 32) tc::tmem_ld32(s_tmem + (c + 1) * 32, nxt);
              uint32_t pk[16];
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float p0 = ex2(fmaf(__uint_as_float(cur[i]), scale_log2, -off));
                float p1 = ex2(fmaf(__uint_as_float(cur[i + 1]), scale_log2, -off));
                if (c0 + i >= n_keys) p0 = 0.f;       // keys beyond the sequence (last block only)
                if (c0 + i + 1 >= n_keys) p1 = 0.f;
                pk[i >> 1] = pack2(p0, p1);
              }
              // k-block (c0 / 64) of P; 16-byte chunk j of row r lives at chunk (j ^ (r & 7))
              const uint32_t blk = p_row + (uint32_t)((c0 >> 6) * (TILE * 128));
              const int chunk0 = (c0 & 32) >> 3;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t chunk = (uint32_t)((chunk0 + j) ^ (row & 7));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + chunk * 16u), "r"(pk[4 * j]),
                             "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
              }
            }
          }
          tc::fence_proxy_async();   // P (generic-proxy stores) -> visible to the tensor core
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&p_full[g]);
        }
        // last block's P V, then O / l -> ctx
        tc::mbar_wait(&o_full[g], n_o & 1);
        ++n_o;
        tc::tc_fence_after();
        uint32_t pv[HEAD_DIM];
        tc::tmem_ld32(pv_tmem, pv);
        const uint32_t lv = tc::tmem_ld1(pv_tmem + L_OFF);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        const int qrow = qt * TILE + row;
        if (qrow < S) {
          const float inv = 1.f / (l + __uint_as_float(lv));
          uint4* dst = reinterpret_cast<uint4*>(ctx + (size_t)(tok0 + qrow) * hidden + head * HEAD_DIM);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = (o[8 * j + e] + __uint_as_float(pv[8 * j + e])) * inv;
            dst[j] = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
          }
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4 * GROUPS + 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace attn_tc
}  // namespace drag
