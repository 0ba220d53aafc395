// Persistent warp-specialised tcgen05 GEMM with fused epilogues for the BERT encoder
// (SURVEY.md 8a row a5):   OUT[M, N] = epilogue( A[M, K] . W[N, K]^T )
//
//   A  : bf16 activations, row-major [M, K]          (K-major operand)
//   W  : bf16 weights as HF stores them, [N, K] row-major (K-major operand)
//   acc: fp32 in TMEM, double buffered (2 x BLOCK_N <= 512 columns)
//
// LayerNorm is FOLDED into the GEMMs instead of being a pass of its own.  Activations are kept
// "raw" (pre-LayerNorm, bf16) together with per-row partial sums (sum v, sum v^2); for a consumer
// GEMM whose input is LN(v) = (v - mu) * rstd * gamma + beta we use
//     LN(v) . W^T = rstd * ( v . (gamma (.) W)^T  -  mu * c ) + d,
//     c_n = sum_k (gamma (.) W)_nk,   d_n = sum_k beta_k W_nk + bias_n,
// i.e. the tensor cores multiply the raw rows by gamma-scaled weights and the epilogue applies
// the per-row (mu, rstd) and per-column (c, d) terms.  Epilogues:
//     EPI_LNIN       out = rstd*(acc - mu*c) + d                               (QKV projection)
//     EPI_LNIN_GELU  out = gelu(rstd*(acc - mu*c) + d)   stored as FP16            (FFN up)
//     EPI_RES        out = acc + cold + LN(residual_raw)  (raw, pre-LN) and the row's partial
//                    (sum, sum^2) over this tile's columns -> out_stats   (attention out / FFN down)
//
// CG = 2 runs the same pipeline on CTA PAIRS (tcgen05 cta_group::2): a pair owns a 256 x BLOCK_N tile,
// each CTA stages its own 128 rows of A and HALF of the W tile (BLOCK_N/2 rows), the leader issues
// M=256 MMAs that read both halves, and each CTA drains its own 128 accumulator rows.  Operand bytes
// per flop drop by a third or more, which is what bounds the K=384 GEMMs (L2 -> SM traffic).
//
// WS = true ("weights stationary", the K = 384 GEMMs): the L2 -> SM fabric (~6300 B/clk chip-wide) is what
// bounds these GEMMs -- a 256 x 192 pair tile moves 336 KB of operands and 96 KB of output for 2304 MMA
// clocks.  A worker therefore keeps ONE column block of W (all of K) resident in shared memory for the
// whole kernel and walks down the rows: only A streams through the ring (-33..37% fabric bytes).
//
// CTA = 2 + EPI_WARPS warps, one CTA per SM, looping over 128 x BLOCK_N output tiles:
//   warp 0   TMA producer (cp.async.bulk.tensor, 128B swizzle) into a STAGES-deep smem ring
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) per k-step;
//            tcgen05.commit releases ring slots / publishes the accumulator
//   warps 2+ epilogue: tcgen05.ld (thread = one row, 32 columns at a time) -> math -> bf16 ->
//            128B-swizzled staging tile in smem -> TMA store (coalesced, no LSU scatter)
#pragma once

#include <cuda_bf16.h>

#include "drag_tc.cuh"

namespace drag {
namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int STATS_PARTS = 3;          // partial (sum, sum^2) slots per row
constexpr int STORE_COLS = 64;          // columns per TMA store box (128 bytes of bf16)
constexpr int STORE_ROWS = 32;          // rows per TMA store box (one epilogue warp)
constexpr int STAGING_BYTES = STORE_ROWS * STORE_COLS * 2;  // 4 KB per epilogue warp

enum { EPI_LNIN = 0, EPI_LNIN_GELU = 1, EPI_RES = 2 };

struct GemmParams {
  int M, N, K;                     // M = valid rows (tokens)
  const float* colc;               // [N] LNIN: c_n
  const float* cold;               // [N] LNIN: d_n ; RES: bias_n + beta_n (beta of the residual's LN)
  const float* gamma;              // [N] RES: gamma of the residual's LN
  const float2* in_stats;          // [M][STATS_PARTS] partial sums of the rows the LN applies to
                                   //   (LNIN: the A rows, RES: the residual rows)
  float inv_width;                 // 1 / (feature width the LN statistics run over)
  float ln_eps;
  const __nv_bfloat16* residual;   // [M, N] raw residual rows (RES)
  float2* out_stats;               // [M][STATS_PARTS] (RES): slot n_blk of every row
  int f16_operands;                // A and W hold fp16 (the FFN-down GEMM: GELU writes fp16), else bf16
  int dbg;                         // ablation switches for the probes (DRAG_GEMM_DBG): 1 no output stores, 2 no epilogue work at all, 4 all stores land on the first row block (L2-resident)
};

template <int BLOCK_N>
struct Cfg {
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "invalid UMMA N");
  static_assert(2 * BLOCK_N <= 512, "accumulator must double-buffer in 512 TMEM columns");
  static_assert(STAGE_BYTES % 1024 == 0 && (B_STAGE_BYTES / 2) % 1024 == 0, "stage must keep 1024-byte alignment");
};

// staging buffers per epilogue warp: a warp with several store groups per tile double buffers; the residual
// epilogue always does (the next tile's residual box is TMA-loaded into the other buffer a tile ahead)
template <int BLOCK_N, int EPI_WARPS, int EPI = EPI_LNIN>
constexpr int stage_bufs() { return (EPI == EPI_RES || (BLOCK_N / (EPI_WARPS / 4)) > STORE_COLS) ? 2 : 1; }

constexpr int WS_K_BLOCKS = 6;   // weights-stationary kernels hold K = 384 (6 k-blocks) of their W column block

template <int BLOCK_N, int STAGES, int EPI_WARPS, int CG = 1, bool WS = false, int EPI = EPI_LNIN>
constexpr size_t smem_bytes() {
  // ring (A + W, or A only next to the resident W block) + per-warp store staging + [barriers, tmem pointer,
  // statistics exchange | 4 KB] + [column constants: 2 tiles x 2 vectors x BLOCK_N floats <= 4 KB] + slack for
  // manual 1024-byte alignment
  return (WS ? (size_t)STAGES * A_STAGE_BYTES + (size_t)WS_K_BLOCKS * (Cfg<BLOCK_N>::B_STAGE_BYTES / CG)
             : (size_t)STAGES * (A_STAGE_BYTES + Cfg<BLOCK_N>::B_STAGE_BYTES / CG)) +
         (size_t)EPI_WARPS * stage_bufs<BLOCK_N, EPI_WARPS, EPI>() * STAGING_BYTES + 8192 + 1024;
}

// Packed fp32 pairs (FFMA2: one instruction per two lanes of the epilogue arithmetic)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// GELU(x) = x * Phi(x) with Phi(x) ~ 0.5 * (1 + tanh(x * (c0 + c1 x^2 + c2 x^4))): minimax fit on
// [-8, 8] (max abs error 2.5e-5 vs the erf form).  Evaluated on PAIRS in fp16 (HFMA2 / MUFU.TANH.F16):
// half the instructions of the fp32 form; the result is stored as fp16 (11-bit significand, finer than
// the bf16 the other activations use), which the FFN-down GEMM reads as its A operand.
__device__ __forceinline__ uint32_t gelu_f16x2(float x_lo, float x_hi) {
  uint32_t x, x2, p, t, hx, o;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x) : "f"(x_hi), "f"(x_lo));
  asm("mul.rn.f16x2 %0, %1, %1;" : "=r"(x2) : "r"(x));
  asm("min.f16x2 %0, %1, %2;" : "=r"(x2) : "r"(x2), "r"(0x54005400u));                      // min(x^2, 64)
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(0x8dc28dc2u), "r"(x2), "r"(0x28bd28bdu));  // -0.00035152 x2 + 0.0370057
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(x2), "r"(0x3a613a61u));          // ... x2 + 0.7975079
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(p) : "r"(x), "r"(p));
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(p));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(hx) : "r"(x), "r"(0x38003800u));                       // 0.5 x
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(o) : "r"(hx), "r"(t), "r"(hx));
  return o;
}

// 16-byte shared-memory load (a pointer derived from the dynamic shared-memory base through casts is a GENERIC pointer
// to the compiler: plain dereferences become LD.E instead of LDS)
__device__ __forceinline__ float4 lds_f4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(tc::smem_u32(p)));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// per-row LayerNorm statistics from the partial sums
__device__ __forceinline__ void row_stats(const float2* stats, int row, bool ok, float inv_width, float eps, float& mu, float& rstd) {
  float s = 0.f, ss = 0.f;
  if (ok) {
#pragma unroll
    for (int i = 0; i < STATS_PARTS; ++i) {
      float2 p = __ldg(stats + (size_t)row * STATS_PARTS + i);
      s += p.x;
      ss += p.y;
    }
  }
  mu = s * inv_width;
  rstd = rsqrtf(fmaxf(ss * inv_width - mu * mu, 0.f) + eps);
}

template <int BLOCK_N, int EPI, int EPI_WARPS, int STAGES, int CG, bool WS>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
            const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res, GemmParams p) {
  using C = Cfg<BLOCK_N>;
  static_assert(CG == 1 || CG == 2, "single CTAs or CTA pairs");
  constexpr int B_BYTES = C::B_STAGE_BYTES / CG;          // this CTA's part of the W tile (one k-block)
  constexpr int STAGE_BYTES = WS ? A_STAGE_BYTES : A_STAGE_BYTES + B_BYTES;
  constexpr int W_RESIDENT_BYTES = WS ? WS_K_BLOCKS * B_BYTES : 0;
  constexpr int TILE_M = BLOCK_M * CG;
  static_assert(EPI_WARPS % 4 == 0 && EPI_WARPS >= 4 && EPI_WARPS <= 16, "epilogue warps come in groups of 4 (one per TMEM lane quarter)");
  static_assert(EPI != EPI_RES || EPI_WARPS <= 12, "the statistics exchange of EPI_RES handles at most three column groups");
  constexpr int COL_GROUPS = EPI_WARPS / 4;
  constexpr int COLS_PER_THREAD = BLOCK_N / COL_GROUPS;
  constexpr int STAGE_BUFS = stage_bufs<BLOCK_N, EPI_WARPS, EPI>();
  static_assert(EPI != EPI_RES || COLS_PER_THREAD == STORE_COLS, "the residual epilogue handles one 64-column group per warp and tile");
  static_assert(COLS_PER_THREAD % STORE_COLS == 0, "epilogue stores 64-column groups");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* w_res = smem;                                   // WS: [k_blocks][this CTA's W rows][64] resident
  uint8_t* ring = smem + W_RESIDENT_BYTES;
  uint8_t* staging = ring + (size_t)STAGES * STAGE_BYTES;  // 1024-aligned, 4 KB per epilogue warp
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + (size_t)EPI_WARPS * STAGE_BUFS * STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* w_bar = tmem_empty_bar + 2;                    // WS: the resident W block has landed
  uint64_t* res_bar = w_bar + 1;                           // RES: [EPI_WARPS][2] the warp's residual box has landed in staging buffer b
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + 32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = p.N / BLOCK_N;
  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = p.K / BLOCK_K;
  const uint32_t cta_rank = CG == 2 ? tc::cluster_ctarank() : 0u;   // 0 = pair leader
  const int worker = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // The worker's tile walk.  Default: tiles worker, worker + n_workers, ... in row-major tile order.
  // WS: workers are dealt to column blocks (worker % n_tiles) and walk down the rows with the
  // other workers of their column block; workers beyond a whole number of groups stay idle.
  const int ws_groups = n_workers / n_tiles;
  const int ws_n_blk = worker % n_tiles, ws_g = worker / n_tiles;
  const int my_tiles = WS ? ((ws_g < ws_groups && ws_g < m_tiles) ? (m_tiles - ws_g + ws_groups - 1) / ws_groups : 0)
                          : (worker < num_tiles ? (num_tiles - worker + n_workers - 1) / n_workers : 0);
  auto tile_m = [&](int it) { return WS ? ws_g + it * ws_groups : (worker + it * n_workers) / n_tiles; };
  auto tile_n = [&](int it) { return WS ? ws_n_blk : (worker + it * n_workers) % n_tiles; };

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_a);
    tc::tma_prefetch_desc(&tmap_w);
    tc::tma_prefetch_desc(&tmap_out);
    if (EPI == EPI_RES) tc::tma_prefetch_desc(&tmap_res);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&tmem_full_bar[s], 1);
      tc::mbar_init(&tmem_empty_bar[s], CG * EPI_WARPS);   // the leader's collects both CTAs' epilogue warps
    }
    tc::mbar_init(w_bar, 1);
    for (int s = 0; s < 2 * EPI_WARPS; ++s) tc::mbar_init(&res_bar[s], 1);
    tc::fence_barrier_init();
  }
  if (CG == 2) tc::cluster_sync_all();   // the peer's barriers must exist before anything signals them
  if (warp == 1) {
    if (CG == 2) {
      tc::tmem_alloc_pair(tmem_ptr_smem, C::TMEM_COLS);
      tc::tmem_relinquish_pair();
    } else {
      tc::tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
      tc::tmem_relinquish();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  tc::pdl_launch_dependents();
  tc::pdl_wait();   // the set-up above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (WS && my_tiles > 0) {
        // the worker's column block of W, all of K, once
        if (CG == 2) {
          if (cta_rank == 0) tc::mbar_arrive_expect_tx(w_bar, (uint32_t)(2 * k_blocks * B_BYTES));
          const uint32_t bar = tc::mapa_shared(tc::smem_u32(w_bar), 0);
          for (int kb = 0; kb < k_blocks; ++kb)
            tc::tma_load_2d_pair(&tmap_w, bar, w_res + (size_t)kb * B_BYTES, kb * BLOCK_K, ws_n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / 2));
        } else {
          tc::mbar_arrive_expect_tx(w_bar, (uint32_t)(k_blocks * B_BYTES));
          for (int kb = 0; kb < k_blocks; ++kb)
            tc::tma_load_2d(&tmap_w, w_bar, w_res + (size_t)kb * B_BYTES, kb * BLOCK_K, ws_n_blk * BLOCK_N);
        }
      }
      for (int it = 0; it < my_tiles; ++it) {
        const int m_blk = tile_m(it), n_blk = tile_n(it);
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = ring + (size_t)stage * STAGE_BYTES;
          if (CG == 2) {
            // both CTAs load their halves; all bytes are credited to the LEADER's full barrier
            if (cta_rank == 0) tc::mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            const uint32_t bar = tc::mapa_shared(tc::smem_u32(&full_bar[stage]), 0);
            tc::tma_load_2d_pair(&tmap_a, bar, a_dst, kb * BLOCK_K, m_blk * TILE_M + (int)cta_rank * BLOCK_M);
            if (!WS) tc::tma_load_2d_pair(&tmap_w, bar, a_dst + A_STAGE_BYTES, kb * BLOCK_K, n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / 2));
          } else {
            tc::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            tc::tma_load_2d(&tmap_a, &full_bar[stage], a_dst, kb * BLOCK_K, m_blk * BLOCK_M);
            if (!WS) tc::tma_load_2d(&tmap_w, &full_bar[stage], a_dst + A_STAGE_BYTES, kb * BLOCK_K, n_blk * BLOCK_N);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (cta_rank == 0 && tc::elect_one()) {
      const uint32_t idesc = p.f16_operands ? tc::umma_idesc_f16(TILE_M, BLOCK_N) : tc::umma_idesc_bf16(TILE_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (WS && my_tiles > 0) {
        tc::mbar_wait(w_bar, 0);
        tc::tc_fence_after();
      }
      for (int it = 0; it < my_tiles; ++it) {
        tc::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t a_addr = tc::smem_u32(ring + (size_t)stage * STAGE_BYTES);
          const uint64_t a_desc = tc::umma_desc_sw128(a_addr);
          const uint64_t b_desc = tc::umma_desc_sw128(WS ? tc::smem_u32(w_res + (size_t)kb * B_BYTES) : a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // +32 bytes per 16-element k-step inside the 128-byte swizzle row (encoded >> 4)
            if (CG == 2) tc::umma_bf16_pair(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            else tc::umma_bf16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // ring slot reusable (in both CTAs of a pair) once these MMAs retire
          if (CG == 2) tc::umma_commit_pair(&empty_bar[stage], 3);
          else tc::umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (CG == 2) tc::umma_commit_pair(&tmem_full_bar[acc], 3);
        else tc::umma_commit(&tmem_full_bar[acc]);
        if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    // Latency is hidden across tiles and chunks: the row statistics and the per-column constants of
    // the NEXT tile are fetched while the current one is processed (registers / cp.async into a
    // double-buffered shared-memory table), the TMEM load of the next 32 columns is in flight during
    // the math of the current 32, and store staging is double buffered.
    const int ew = warp - 2;
    const int et = ew * 32 + lane;          // index among the epilogue threads
    const int quarter = warp & 3;           // TMEM lanes this warp may touch: [32*quarter, +32)
    const int col_group = ew >> 2;
    const int row_in_tile = quarter * 32 + lane;
    uint8_t* stage_buf = staging + (size_t)ew * (STAGE_BUFS * STAGING_BYTES);
    const uint32_t stage_row = tc::smem_u32(stage_buf) + (uint32_t)lane * 128u;
    float* coltab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 4096);  // [2][2][BLOCK_N], second 4 KB of the region
    const float* col_v0 = EPI == EPI_RES ? p.gamma : p.colc;
    int acc = 0;
    uint32_t acc_phase = 0;
    int par = 0, sbuf = 0;
    const uint32_t leader_tmem_empty = CG == 2 ? tc::mapa_shared(tc::smem_u32(&tmem_empty_bar[0]), 0) : 0u;

    auto fetch_cols = [&](int it_, int par_) {
      // BLOCK_N/4 16-byte chunks of each of the two column vectors of the tile's column block
      const int n_blk_ = tile_n(it_);
      if (et < BLOCK_N / 2) {
        const bool second = et >= BLOCK_N / 4;
        const int chunk = second ? et - BLOCK_N / 4 : et;
        const float* src = (second ? p.cold : col_v0) + n_blk_ * BLOCK_N + chunk * 4;
        const uint32_t dst = tc::smem_u32(coltab + (par_ * 2 + (second ? 1 : 0)) * BLOCK_N + chunk * 4);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto fetch_stats = [&](int it_, float2 (&st)[STATS_PARTS]) {
      const int row_ = tile_m(it_) * TILE_M + (int)cta_rank * BLOCK_M + row_in_tile;
      // volatile asm: the loads must be ISSUED here (a tile ahead of their use), not sunk to the use
#pragma unroll
      for (int i = 0; i < STATS_PARTS; ++i) {
        st[i] = make_float2(0.f, 0.f);
        if (row_ < p.M)
          asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(st[i].x), "=f"(st[i].y) : "l"(p.in_stats + (size_t)row_ * STATS_PARTS + i));
      }
    };

    // RES: the warp's 32 x 64 box of raw residual rows arrives by TMA in the staging buffer the output
    // will be written to (same box, same swizzle: the epilogue works in place), one tile ahead
    auto fetch_residual = [&](int it_, int buf_) {
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(&res_bar[2 * ew + buf_], STAGING_BYTES);
        tc::tma_load_2d(&tmap_res, &res_bar[2 * ew + buf_], stage_buf + (size_t)buf_ * STAGING_BYTES, tile_n(it_) * BLOCK_N + col_group * COLS_PER_THREAD,
                        tile_m(it_) * TILE_M + (int)cta_rank * BLOCK_M + quarter * 32);
      }
    };
    float2 next_stats[STATS_PARTS];
    if (my_tiles > 0) {
      fetch_cols(0, 0);
      fetch_stats(0, next_stats);
      if (EPI == EPI_RES) fetch_residual(0, 0);
    }
    for (int it = 0; it < my_tiles; ++it) {
      const int m_blk = tile_m(it), n_blk = tile_n(it);
      const int row0 = m_blk * TILE_M + (int)cta_rank * BLOCK_M;   // first row of this CTA's 128-row slab
      const int row = row0 + row_in_tile;
      const bool row_ok = row < p.M;
      const int col0 = n_blk * BLOCK_N + col_group * COLS_PER_THREAD;
      float mu, rstd;
      {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int i = 0; i < STATS_PARTS; ++i) { s += next_stats[i].x; ss += next_stats[i].y; }
        mu = s * p.inv_width;
        rstd = rsqrtf(fmaxf(ss * p.inv_width - mu * mu, 0.f) + p.ln_eps);
      }
      // this tile's column constants are in coltab[par]; everybody is done with coltab[par ^ 1]
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      if (it + 1 < my_tiles) {
        fetch_cols(it + 1, par ^ 1);
        fetch_stats(it + 1, next_stats);
        if (EPI == EPI_RES) {
          // the other staging buffer was last read by the TMA store of tile it-1 (the newest committed group)
          if (lane == 0) tc::tma_store_wait_read();
          fetch_residual(it + 1, sbuf ^ 1);
        }
      }
      const float* tab0 = coltab + (par * 2) * BLOCK_N + col_group * COLS_PER_THREAD;
      const float* tab1 = tab0 + BLOCK_N;
      float s_sum = 0.f, s_sq = 0.f;

      tc::mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N + col_group * COLS_PER_THREAD);

      constexpr int N_CHUNKS = COLS_PER_THREAD / 32;
      if (p.dbg & 2) {
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) tc::mbar_arrive_cluster(leader_tmem_empty + (uint32_t)(acc * 8));
          else tc::mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        par ^= 1;
        continue;
      }
      uint32_t r[32], r_next[32];
      uint4 rv[4];
      tc::tmem_ld32(t_row, r);
      if (EPI == EPI_RES) tc::mbar_wait(&res_bar[2 * ew + sbuf], (uint32_t)((it >> 1) & 1));
#pragma unroll
      for (int ch = 0; ch < N_CHUNKS; ++ch) {
        const int c = ch * 32;
        const int half = ch & 1;
        tc::tmem_ld_wait();
        if (ch + 1 < N_CHUNKS) {
          // next 32 columns: in flight during this chunk's math
          tc::tmem_ld32(t_row + c + 32, r_next);
        }
        uint32_t o[16];
        if (EPI == EPI_RES) {
          // this chunk's 32 residual values: 16-byte chunk j of row r lives at chunk (j ^ (r & 7))
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t chunk = (uint32_t)((half * 4 + j) ^ (lane & 7));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rv[j].x), "=r"(rv[j].y), "=r"(rv[j].z), "=r"(rv[j].w)
                         : "r"(stage_row + (uint32_t)(sbuf * STAGING_BYTES) + chunk * 16u) : "memory");
            if (!row_ok) rv[j] = make_uint4(0, 0, 0, 0);   // rows past M hold stale data
          }
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(rv);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 ga = lds_f4(tab0 + c + 4 * i);
            const float4 cd = lds_f4(tab1 + c + 4 * i);
            const float g4[4] = {ga.x, ga.y, ga.z, ga.w};
            const float d4[4] = {cd.x, cd.y, cd.z, cd.w};
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = 4 * i + j;
              const uint32_t w = rw[e >> 1];
              const float res_raw = (e & 1) ? __uint_as_float(w & 0xffff0000u) : __uint_as_float(w << 16);
              const float a = g4[j] * rstd;
              v[j] = fmaf(res_raw - mu, a, __uint_as_float(r[e]) + d4[j]);
              s_sum += v[j];
              s_sq = fmaf(v[j], v[j], s_sq);
            }
            o[2 * i] = pack_bf16(v[0], v[1]);
            o[2 * i + 1] = pack_bf16(v[2], v[3]);
          }
        } else {
          const uint64_t nmu2 = pack_f32x2(-mu, -mu), rstd2 = pack_f32x2(rstd, rstd);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 cc = lds_f4(tab0 + c + 4 * i);
            const float4 cd = lds_f4(tab1 + c + 4 * i);
            // v = rstd * (acc - mu * c) + d on fp32 pairs
            const uint64_t v01 = fma_f32x2(rstd2, fma_f32x2(nmu2, pack_f32x2(cc.x, cc.y), pack_f32x2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]))),
                                           pack_f32x2(cd.x, cd.y));
            const uint64_t v23 = fma_f32x2(rstd2, fma_f32x2(nmu2, pack_f32x2(cc.z, cc.w), pack_f32x2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]))),
                                           pack_f32x2(cd.z, cd.w));
            float v0, v1, v2, v3;
            unpack_f32x2(v01, v0, v1);
            unpack_f32x2(v23, v2, v3);
            if (EPI == EPI_LNIN_GELU) {
              o[2 * i] = gelu_f16x2(v0, v1);
              o[2 * i + 1] = gelu_f16x2(v2, v3);
            } else {
              o[2 * i] = pack_bf16(v0, v1);
              o[2 * i + 1] = pack_bf16(v2, v3);
            }
          }
        }
        if (half == 0 && EPI != EPI_RES) {
          // the TMA store that last read this staging buffer must be done with it
          // (RES: already waited for before the residual box was loaded into it)
          if (lane == 0) {
            if (STAGE_BUFS == 2) tc::tma_store_wait_read_1();
            else tc::tma_store_wait_read();
          }
          __syncwarp();
        }
        // 128B-swizzled row: 16-byte chunk j of row r lives at chunk (j ^ (r & 7))
        const uint32_t srow = stage_row + (uint32_t)(sbuf * STAGING_BYTES);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t chunk = (uint32_t)((half * 4 + j) ^ (lane & 7));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + chunk * 16u), "r"(o[4 * j]),
                       "r"(o[4 * j + 1]), "r"(o[4 * j + 2]), "r"(o[4 * j + 3]) : "memory");
        }
        if (half == 1) {
          tc::fence_proxy_async();
          __syncwarp();
          if (!(p.dbg & 1) && tc::elect_one()) {
            tc::tma_store_2d(&tmap_out, stage_buf + (size_t)sbuf * STAGING_BYTES, col0 + c - 32, ((p.dbg & 4) ? (int)(blockIdx.x & 1) * BLOCK_M : row0) + quarter * 32);
            tc::tma_store_commit();
          }
          if (STAGE_BUFS == 2) sbuf ^= 1;
        }
        if (ch + 1 < N_CHUNKS) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = r_next[i];
        }
      }
      // all TMEM reads of this warp are complete -> hand the accumulator back to the MMA warp
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) tc::mbar_arrive_cluster(leader_tmem_empty + (uint32_t)(acc * 8));
        else tc::mbar_arrive(&tmem_empty_bar[acc]);
      }
      if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      par ^= 1;

      if (EPI == EPI_RES) {
        if (COL_GROUPS == 1) {
          if (row_ok) p.out_stats[(size_t)row * STATS_PARTS + n_blk] = make_float2(s_sum, s_sq);
        } else {
          // several warps share a row (column groups): combine through shared memory in a fixed order
          float2* part = reinterpret_cast<float2*>(tmem_ptr_smem + 4);  // [COL_GROUPS - 1][128]
          if (col_group > 0) part[(col_group - 1) * BLOCK_M + row_in_tile] = make_float2(s_sum, s_sq);
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
          if (col_group == 0 && row_ok) {
#pragma unroll
            for (int cg = 1; cg < COL_GROUPS; ++cg) {
              const float2 o2 = part[(cg - 1) * BLOCK_M + row_in_tile];
              s_sum += o2.x;
              s_sq += o2.y;
            }
            p.out_stats[(size_t)row * STATS_PARTS + n_blk] = make_float2(s_sum, s_sq);
          }
        }
      }
    }
    if (lane == 0) tc::tma_store_wait_all();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (CG == 2) tc::cluster_sync_all();   // the peer may still signal this CTA's barriers / read its operands
  if (warp == 1) {
    tc::tc_fence_after();
    if (CG == 2) tc::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace gemm
}  // namespace drag
