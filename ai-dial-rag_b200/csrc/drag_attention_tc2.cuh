// Multi-head self-attention on tcgen05 for the packed (padding-free) batch, head_dim = 32
// (SURVEY.md 8a row a5:  softmax(Q K^T / sqrt(32) + key mask) V ).  Second design: at head_dim 32 the op
// is bound by the exponentials (MUFU: 16 per clock per SM, measured), so everything else is arranged to
// keep four softmax warps per scheduler busy with  load S -> fma -> ex2 -> pack -> store P  and to take
// every tensor-core round trip off their critical path.
//
// Persistent kernel, one CTA per SM, 18 warps.  The CTA walks a stream of "tiles": a tile is one
// 128-query tile of one (sequence, head) item against one 256-key super-block.
//   warp 16      TMA producer: Q, K and V of the next items (64-byte swizzle: a row of 32 bf16 is one
//                swizzle row) straight out of the packed [T, 3*hidden] QKV buffer, 1-4 stages
//   warp 17      MMA issuer (one thread):  S(t+1) = Q K^T is issued BEFORE P(t) V, into the other half of
//                TMEM, so the scores of the next tile are ready when the softmax warps get there
//   warps 0-15   softmax: FOUR threads per query row -- warp w owns TMEM lanes 32 (w % 4) .. +31 and the
//                64-key unit w / 4 of the row
// Per tile:
//   S[128 x 256] = Q K^T        tcgen05.mma M=128 N=256 K=16 x2 -> 256 TMEM columns (two buffers)
//   unit max m_u                tcgen05.ld + 3-input max over the thread's own 64 scores: every (row, unit)
//                               keeps its OWN softmax reference, so nothing waits for the other threads
//   P_u = 2^(s*c - m_u*c)       bf16 -> shared memory, one [128 x 64] K-major 128B-swizzled buffer per
//                               unit (the A operand of P V); unit sums l_u in fp32 registers
//   O_u[128 x 32] = P_u V_u     tcgen05.mma M=128 N=32 K=16 x4 per unit into the first (already consumed)
//                               32 score columns of that unit; V is consumed as stored ([key][32]: an
//                               MN-major B operand)
//   combine (one tile LATER)    O = sum_u w_u O_u / sum_u w_u l_u with w_u = 2^((m_u - max_u m_u) c): exactly
//                               the softmax; m_u, l_u travel through shared memory, each thread finishes 8
//                               of the 32 output columns.  Sequences of 257..512 tokens are two tiles per
//                               query tile, folded into the same running (m, O, l).
// Keys beyond the sequence do not exist in the packed layout: the tail unit is masked.
#pragma once

#include <cuda_bf16.h>

#include "drag_attention_tc.cuh"   // descriptors, ex2, max3, pack2
#include "drag_tc.cuh"

namespace drag {
namespace attn_tc2 {

using attn_tc::HEAD_DIM;
using attn_tc::QKV_TILE_BYTES;
using attn_tc::TILE;

constexpr int SOFTMAX_WARPS = 16;
constexpr int TMA_WARP = SOFTMAX_WARPS;
constexpr int MMA_WARP = TMA_WARP + 1;
constexpr int THREADS = 32 * (MMA_WARP + 1);   // 576
constexpr int TMEM_COLS = 512;                 // score buffer b: columns [256 b, 256 b + 256)
constexpr int SB_KEYS = 256;                   // keys per super-block
constexpr int UNIT = 64;                       // keys per unit (one softmax thread per row and unit)
constexpr int UNITS = SB_KEYS / UNIT;
constexpr int P_UNIT_BYTES = TILE * UNIT * 2;  // 16 KB: [128 rows][64 keys] bf16, 128-byte swizzle
constexpr int P_BYTES = UNITS * P_UNIT_BYTES;
constexpr int MX_SLOTS = 3;                    // tiles whose (m_u, l_u) are live at once: written in tile t, read in tile t+1
constexpr int MX_BYTES = MX_SLOTS * UNITS * TILE * 4;   // [slot][unit][row] unit maxima (same again for unit sums)
constexpr int MAX_STAGES = 4;
constexpr int SMEM_BUDGET = 227 * 1024;

__host__ __device__ inline int item_stages(int max_len) {
  const int tiles = (max_len + TILE - 1) / TILE;
  const int fit = (SMEM_BUDGET - P_BYTES - 2 * MX_BYTES - 1024 - 256) / (3 * tiles * QKV_TILE_BYTES);
  return fit < MAX_STAGES ? fit : MAX_STAGES;   // 4, 3, 2, 1 stages for 1..4 tiles
}
__host__ __device__ inline size_t smem_bytes(int max_len) {
  const int tiles = (max_len + TILE - 1) / TILE;
  return (size_t)item_stages(max_len) * 3 * tiles * QKV_TILE_BYTES + P_BYTES + 2 * MX_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// The four threads of a query row sit in the four warps that share a TMEM lane quarter (w, w+4, w+8, w+12):
// those 128 threads synchronise among themselves (named barrier 1 + quarter), not the whole CTA.
__device__ __forceinline__ void softmax_barrier(int quarter) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(SOFTMAX_WARPS * 32 / 4) : "memory");
}

// Walks the CTA's tiles in a fixed order: items blockIdx.x + i * gridDim.x, their query tiles, and per
// query tile one tile per 256-key super-block.  Every role runs the same walker.
struct TileWalker {
  const int* cu;
  int n_items, heads;
  int it;        // CTA-local item ordinal
  int item;      // global item index
  int S, tok0;
  int n_tiles;   // 128-row tiles of the item (query tiles = key tiles)
  int n_sb;      // key super-blocks: 1 or 2
  int qt;        // query tile
  int sb;        // key super-block of the current tile
  bool job_first, job_last;    // first / last tile of the query tile
  bool item_first, item_last;
  __device__ void load_item() {
    if (item < n_items) {
      const int seq = item / heads;
      tok0 = __ldg(cu + seq);
      S = __ldg(cu + seq + 1) - tok0;
      n_tiles = (S + TILE - 1) / TILE;
      n_sb = n_tiles > 2 ? 2 : 1;
    }
  }
  __device__ void init(const int* cu_seqlens, int items, int n_heads) {
    cu = cu_seqlens; n_items = items; heads = n_heads;
    it = 0; item = blockIdx.x; qt = 0; sb = -1; n_tiles = 0; n_sb = 1; S = 0; tok0 = 0;
    load_item();
  }
  __device__ bool next() {
    if (item >= n_items) return false;
    if (++sb >= n_sb) {
      sb = 0;
      if (++qt >= n_tiles) {
        qt = 0; ++it; item += gridDim.x;
        load_item();
        if (item >= n_items) return false;
      }
    }
    job_first = sb == 0;
    job_last = sb == n_sb - 1;
    item_first = qt == 0 && sb == 0;
    item_last = qt == n_tiles - 1 && job_last;
    return true;
  }
};

// 32 scores of a row (one tcgen05.ld) into the running row maximum
template <bool MASKED>
__device__ __forceinline__ float max_of_32(uint32_t taddr, int valid, float mx) {
  uint32_t r[32];
  tc::tmem_ld32(taddr, r);
  tc::tmem_ld_wait();
  if (MASKED) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i >= valid) r[i] = 0xff800000u;   // keys beyond the sequence
  }
#pragma unroll
  for (int i = 0; i < 32; i += 2) mx = attn_tc::max3(mx, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
  return mx;
}

// 32 scores of a row -> 2^(s c - m c) as bf16 into the row's 64 bytes of the P buffer (four 16-byte chunks
// starting at chunk0 of the 128-byte swizzled row); returns the sum of the 32 weights
template <bool MASKED>
__device__ __forceinline__ float exp_of_32(uint32_t taddr, int valid, float scale_log2, float off, uint32_t p_row, int chunk0, int row) {
  uint32_t r[32];
  tc::tmem_ld32(taddr, r);
  tc::tmem_ld_wait();
  uint32_t pk[16];
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float p0 = attn_tc::ex2(fmaf(__uint_as_float(r[i]), scale_log2, -off));
    float p1 = attn_tc::ex2(fmaf(__uint_as_float(r[i + 1]), scale_log2, -off));
    if (MASKED) {
      if (i >= valid) p0 = 0.f;
      if (i + 1 >= valid) p1 = 0.f;
    }
    l0 += p0;
    l1 += p1;
    pk[i >> 1] = attn_tc::pack2(p0, p1);
  }
  // 16-byte chunk j of row r lives at chunk (j ^ (r & 7)) of the row's 128 bytes
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t chunk = (uint32_t)((chunk0 + j) ^ (row & 7));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + chunk * 16u), "r"(pk[4 * j]),
                 "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
  }
  return l0 + l1;
}

// qkv : [T, 3*hidden] bf16 (tensor map: box 32 columns x 128 rows, 64-byte swizzle)
// ctx : [T, hidden] bf16
// grid = min(#SMs, heads * n_seq), block = THREADS, dynamic smem = smem_bytes(longest sequence)
__global__ void __launch_bounds__(THREADS, 1)   // 18 warps: 96 registers per thread (112 no longer fits the register file's allocation granularity)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ ctx,
                     const int* __restrict__ cu_seqlens, int n_seq, int heads, int max_tiles, int n_stages,
                     float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int hidden = heads * HEAD_DIM;
  const size_t stage_bytes = (size_t)3 * max_tiles * QKV_TILE_BYTES;  // [Q tiles | K tiles | V tiles]
  uint8_t* p_smem = smem + (size_t)n_stages * stage_bytes;            // multiple of 8 KB: 1024-aligned
  float* mx_smem = reinterpret_cast<float*>(p_smem + P_BYTES);
  float* lx_smem = mx_smem + MX_BYTES / 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + P_BYTES + 2 * MX_BYTES);
  uint64_t* kv_full = bars;                    // [MAX_STAGES] TMA -> MMA
  uint64_t* kv_empty = bars + MAX_STAGES;      // [MAX_STAGES] MMA -> TMA (all MMAs of the item retired)
  uint64_t* s_full = bars + 2 * MAX_STAGES;    // [2] MMA -> softmax: the scores of the tile in buffer b are complete
  uint64_t* s_free = s_full + 2;               // [2] softmax -> MMA: buffer b (scores and O) may be overwritten
  uint64_t* p_full = s_free + 2;               // softmax -> MMA: P of the tile is in shared memory
  uint64_t* o_full = p_full + 1;               // MMA -> softmax: O of the tile is complete (and P is reusable)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_items = n_seq * heads;

  if (warp == TMA_WARP) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmap_qkv);
      for (int s = 0; s < MAX_STAGES; ++s) {
        tc::mbar_init(&kv_full[s], 1);
        tc::mbar_init(&kv_empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        tc::mbar_init(&s_full[b], 1);
        tc::mbar_init(&s_free[b], SOFTMAX_WARPS);
      }
      tc::mbar_init(p_full, SOFTMAX_WARPS);
      tc::mbar_init(o_full, 1);
      tc::fence_barrier_init();
    }
    __syncwarp();
  }
  if (warp == MMA_WARP) {
    tc::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == TMA_WARP) {
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
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer:  S(0) | S(t+1), P(t) V | ... =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s256 = attn_tc::idesc_bf16(TILE, 256, false);
      constexpr uint32_t idesc_s128 = attn_tc::idesc_bf16(TILE, 128, false);
      constexpr uint32_t idesc_o = attn_tc::idesc_bf16(TILE, HEAD_DIM, true);
      TileWalker w;
      w.init(cu_seqlens, n_items, heads);
      uint32_t n_pv = 0;
      // scores of the walker's current tile (ordinal t) into buffer t & 1
      auto issue_scores = [&](uint32_t t) {
        const int stage = w.it % n_stages;
        if (w.item_first) tc::mbar_wait(&kv_full[stage], (uint32_t)(w.it / n_stages) & 1);
        tc::mbar_wait(&s_free[t & 1], ((t >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t base = tc::smem_u32(smem + (size_t)stage * stage_bytes);
        const int key_tiles = w.n_tiles - 2 * w.sb >= 2 ? 2 : 1;
        const uint64_t q_desc = attn_tc::desc_k_sw64(base + (uint32_t)(w.qt * QKV_TILE_BYTES));
        const uint64_t k_desc = attn_tc::desc_k_sw64(base + (uint32_t)((max_tiles + 2 * w.sb) * QKV_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < HEAD_DIM / 16; ++k)
          tc::umma_bf16(tmem_base + (t & 1) * SB_KEYS, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2),
                        key_tiles == 2 ? idesc_s256 : idesc_s128, k != 0 ? 1u : 0u);
        tc::umma_commit(&s_full[t & 1]);
      };
      bool have = w.next();
      if (have) issue_scores(0);
      for (uint32_t t = 0; have; ++t) {
        // what P(t) V needs to know about tile t, before the walker moves on
        const bool item_last = w.item_last;
        const int stage = w.it % n_stages;
        const int n_units = (min(w.S - w.sb * SB_KEYS, SB_KEYS) + UNIT - 1) / UNIT;
        const uint32_t v_base = tc::smem_u32(smem + (size_t)stage * stage_bytes) + (uint32_t)(2 * max_tiles * QKV_TILE_BYTES) +
                                (uint32_t)(w.sb * SB_KEYS * 64);
        have = w.next();
        // with a single stage the next item cannot land before this item's last MMAs have been issued
        const bool scores_first = have && !(n_stages == 1 && w.item_first);
        if (scores_first) issue_scores(t + 1);
        tc::mbar_wait(p_full, t & 1);
        tc::tc_fence_after();
        for (int u = 0; u < n_units; ++u) {
          const uint64_t a_desc = tc::umma_desc_sw128(tc::smem_u32(p_smem + (size_t)u * P_UNIT_BYTES));
#pragma unroll
          for (int jj = 0; jj < UNIT / 16; ++jj) {
            const uint64_t b_desc = attn_tc::desc_mn_sw64(v_base + (uint32_t)((u * UNIT + jj * 16) * 64));
            tc::umma_bf16(tmem_base + (t & 1) * SB_KEYS + u * UNIT, a_desc + (uint64_t)(jj * 2), b_desc, idesc_o, jj != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(o_full);
        if (item_last) tc::umma_commit(&kv_empty[stage]);   // every MMA that reads the item's stage has been issued
        if (have && !scores_first) issue_scores(t + 1);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax: four threads per query row =====================
    const int u = warp >> 2;                        // the 64-key unit of the row this thread owns
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_tmem = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t p_row = tc::smem_u32(p_smem + (size_t)u * P_UNIT_BYTES) + (uint32_t)row * 128u;
    // running softmax state of this thread's 8 output columns over the tiles of a query tile
    float m_run = -INFINITY, l_run = 0.f;
    float o_run[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o_run[i] = 0.f;
    // the previous tile: its O_u are read (and combined) one tile later
    bool pend = false, pend_first = false, pend_last = false;
    uint32_t pend_t = 0;
    int pend_units = 0, pend_qrow = 0, pend_S = 0, pend_tok0 = 0, pend_head = 0;
    auto drain = [&]() {
      // O_u of the pending tile: complete once its MMAs retire (they also release the P buffers)
      tc::mbar_wait(o_full, pend_t & 1);
      tc::tc_fence_after();
      const uint32_t o_tmem = lane_tmem + (pend_t & 1) * SB_KEYS + 8 * u;
      uint32_t pv[UNITS][8];
#pragma unroll
      for (int v = 0; v < UNITS; ++v)
        if (v < pend_units) tmem_ld8(o_tmem + v * UNIT, pv[v]);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_free[pend_t & 1]);
      const float* mxs = mx_smem + (pend_t % MX_SLOTS) * (UNITS * TILE) + row;
      const float* lxs = lx_smem + (pend_t % MX_SLOTS) * (UNITS * TILE) + row;
      if (pend_first) {
        m_run = -INFINITY;
        l_run = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) o_run[i] = 0.f;
      }
      float m_new = m_run;
#pragma unroll
      for (int v = 0; v < UNITS; ++v)
        if (v < pend_units) m_new = fmaxf(m_new, mxs[v * TILE]);
      // unit 0 of the first tile holds key 0, which is always valid: m_new is finite
      const float keep = attn_tc::ex2((m_run - m_new) * scale_log2);   // 0 for the first tile (m_run = -inf)
      l_run *= keep;
#pragma unroll
      for (int i = 0; i < 8; ++i) o_run[i] *= keep;
#pragma unroll
      for (int v = 0; v < UNITS; ++v) {
        if (v < pend_units) {
          const float wv = attn_tc::ex2((mxs[v * TILE] - m_new) * scale_log2);
          l_run = fmaf(wv, lxs[v * TILE], l_run);
#pragma unroll
          for (int i = 0; i < 8; ++i) o_run[i] = fmaf(wv, __uint_as_float(pv[v][i]), o_run[i]);
        }
      }
      m_run = m_new;
      if (pend_last && pend_qrow < pend_S) {
        const float inv = 1.f / l_run;
        *reinterpret_cast<uint4*>(ctx + (size_t)(pend_tok0 + pend_qrow) * hidden + pend_head * HEAD_DIM + 8 * u) =
            make_uint4(attn_tc::pack2(o_run[0] * inv, o_run[1] * inv), attn_tc::pack2(o_run[2] * inv, o_run[3] * inv),
                       attn_tc::pack2(o_run[4] * inv, o_run[5] * inv), attn_tc::pack2(o_run[6] * inv, o_run[7] * inv));
      }
      pend = false;
    };
    TileWalker w;
    w.init(cu_seqlens, n_items, heads);
    for (uint32_t t = 0; w.next(); ++t) {
      const uint32_t buf = t & 1;
      const uint32_t s_tmem = lane_tmem + buf * SB_KEYS + (uint32_t)(u * UNIT);
      const int n_keys = min(w.S - w.sb * SB_KEYS, SB_KEYS);
      const int valid = n_keys - u * UNIT;           // keys of this thread's unit inside the sequence
      tc::mbar_wait(&s_full[buf], (t >> 1) & 1);
      tc::tc_fence_after();
      // ---------------- unit maximum, P_u = 2^(s c - m_u c) -> shared memory ----------------
      // The previous tile is drained after this tile's unit maximum: late enough for its MMAs to have
      // (nearly) retired, BEFORE the first P store (the P buffers are single: P(t-1) must have been read),
      // before this warp's arrival on p_full (so o_full can never run a whole phase ahead of a waiting
      // warp), and early enough for S(t+1), which needs the freed buffer, to be computed behind this
      // tile's exponentials.
      float mu = -INFINITY, lu = 0.f;
      if (valid >= UNIT) {
        mu = max_of_32<false>(s_tmem, 32, mu);
        mu = max_of_32<false>(s_tmem + 32, 32, mu);
      } else if (valid > 0) {
        mu = max_of_32<true>(s_tmem, valid, mu);
        if (valid > 32) mu = max_of_32<true>(s_tmem + 32, valid - 32, mu);
      }
      if (pend) drain();
      const float off = mu * scale_log2;
      if (valid >= UNIT) {
        lu = exp_of_32<false>(s_tmem, 32, scale_log2, off, p_row, 0, row);
        lu += exp_of_32<false>(s_tmem + 32, 32, scale_log2, off, p_row, 4, row);
      } else if (valid > 0) {
        // tail unit: keys beyond the sequence get weight 0 (the MMA reads the whole unit)
        lu = exp_of_32<true>(s_tmem, valid, scale_log2, off, p_row, 0, row);
        lu += exp_of_32<true>(s_tmem + 32, valid - 32, scale_log2, off, p_row, 4, row);
      }
      mx_smem[(t % MX_SLOTS) * (UNITS * TILE) + u * TILE + row] = mu;
      lx_smem[(t % MX_SLOTS) * (UNITS * TILE) + u * TILE + row] = lu;
      tc::fence_proxy_async();   // P (generic-proxy stores) -> visible to the tensor core
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_full);
      softmax_barrier(warp & 3); // one per tile: publishes this tile's (m_u, l_u) to the row's other three threads
      pend = true; pend_t = t; pend_first = w.job_first; pend_last = w.job_last;
      pend_units = (n_keys + UNIT - 1) / UNIT;
      pend_qrow = w.qt * TILE + row; pend_S = w.S; pend_tok0 = w.tok0; pend_head = w.item % heads;
    }
    if (pend) drain();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace attn_tc2
}  // namespace drag
