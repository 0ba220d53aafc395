// Fused feed-forward block of a BERT layer on tcgen05:
//
//     OUT[T, 384] = gelu( LN_1(X) . W1^T + b_1 ) . W2^T + b_2 + LN_1(X)        (raw, pre-LayerNorm, + row statistics)
//
// The two GEMM kernels it replaces (drag_gemm.cuh: EPI_LNIN_GELU then EPI_RES) send the 1536-wide intermediate through
// the L2 <-> SM fabric twice (0.8 GB written, 1.6 GB read back per layer at 262 144 tokens), and that fabric is what
// bounds them (profiles/r01_e_gemm_ablation.txt).  Here the intermediate never leaves the SM: a CTA PAIR (cta_group::2)
// owns 256 tokens, keeps their X rows resident in shared memory and walks the 1536 hidden units in 24 chunks of 64:
//
//   G1(c)   Hacc[c & 1] = X . W1g[chunk c]^T          M=256, N=64, K=384   (W1g = gamma_1 (.) W1: the LayerNorm is folded)
//   E1(c)   h = gelu(rstd * (Hacc - mu * colc) + cold) -> fp16 pairs, written back INTO the accumulator's own tensor
//           memory columns (32 fp32 columns read -> 16 packed columns written, per warp)
//   G2(c)   OUT += h . W2[:, chunk c]^T                M=256, N=2 x 192, K=64,  A operand FROM TENSOR MEMORY
//   E2      once per 256 tokens: OUT + cold2 + LN_1(X) -> bf16 + row statistics, TMA store
//
// Tensor memory (512 columns): OUT [0, 384) | Hacc[0] [384, 448) | Hacc[1] [448, 512).  The MMA issuer runs
// G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) | ...: E1(c) has the time of G2(c-1) + G1(c+1), and the tensor pipe executes
// the MMAs of one issuing thread in order, so G1(c+2) overwrites Hacc[c & 1] only after G2(c) has consumed the packed h
// in it.  Fabric bytes per pair and tile: 192 KB of X + 2.36 MB of weights + 192 KB out for 36 864 MMA clocks = 40 B per
// clock and SM, against 83 for the two separate kernels.
//
// Shared memory per CTA: X 96 KB (6 k-blocks [128 rows][64], 128-byte swizzle) | two weight slots of 48 KB (this CTA's
// half of a W1 chunk: 6 x [32 rows][64]; of a W2 chunk: 2 x [96 rows][64]) | 8 x 2 KB store staging | the per-column
// epilogue terms of both GEMMs (15 KB, loaded once).
// The residual rows LN_1(X) of E2 are the X rows themselves: every epilogue thread copies its row's 192 values from the
// X tile into registers before the tile is released (chunk 20), so the next tile's X streams in under the last chunks.
//
// CTA = 20 warps in five warpgroups with their own register budgets (setmaxnreg):
//   warps 0-3    warp 0 TMA producer, warp 1 MMA issuer (pair leader only), two idle                       64 registers
//   warps 4-11   epilogue class A: E1 on hidden units [0, 32) of every chunk (16 per warp) and all of E2   160 registers
//                (thread = token row; warps 4-7 / 8-11 take output columns [0, 192) / [192, 384))
//   warps 12-19  epilogue class B: E1 on hidden units [32, 64) of every chunk                               48 registers
// E1's instruction mix (per 32 elements: 112 half2 FMA-pipe, 32 MUFU.TANH, 32 FFMA2, 16 F2FP, measured) runs pipe after
// pipe in a single warp; four warps per scheduler interleave the pipes, two did not (1 560 clocks per chunk against a
// MUFU floor of 512: profiles/r02_mlp_development.txt).
#pragma once

#include <cuda_bf16.h>

#include "drag_gemm.cuh"

namespace drag {
namespace mlp {

constexpr int HIDDEN = 384;
constexpr int INTER = 1536;
constexpr int ROWS = 128;                        // token rows per CTA (a pair: 256)
constexpr int NC = 64;                           // hidden units per chunk
constexpr int CHUNKS = INTER / NC;               // 24
constexpr int KB = HIDDEN / 64;                  // k-blocks of X
constexpr int X_KB_BYTES = ROWS * 128;           // 16 KB
constexpr int X_BYTES = KB * X_KB_BYTES;         // 96 KB
constexpr int W1_KB_BYTES = (NC / 2) * 128;      // 4 KB: this CTA's 32 rows of a W1 chunk, one k-block
constexpr int W1_BYTES = KB * W1_KB_BYTES;       // 24 KB
constexpr int OUT_HALF = HIDDEN / 2;             // 192 output columns per G2 instruction
constexpr int W2_HALF_BYTES = (OUT_HALF / 2) * 128;   // 12 KB: this CTA's 96 rows of one output half, the chunk's 64 k
constexpr int W2_BYTES = 2 * W2_HALF_BYTES;      // 24 KB
constexpr int W_SLOT_BYTES = W1_BYTES + W2_BYTES;
constexpr int EPI_WARPS = 8;                     // class A (E1 + E2); class B adds another 8 for E1
constexpr int E1_WARPS = 16;
constexpr int THREADS = 128 + 32 * E1_WARPS;     // 640
constexpr int REGS_CONTROL = 64, REGS_A = 160, REGS_B = 48;   // 128 x 64 + 256 x 160 + 256 x 48 = 61 440 = 640 x 96: the CTA can only redistribute what it was launched with
constexpr int E1_COLS = NC / 4;                  // 16 hidden units per E1 warp and chunk
constexpr int STORE_COLS = 16;                   // E2 works in steps of 16 output columns
constexpr int STAGING_BYTES = 2048;              // per epilogue warp: two [32 rows][16 columns] bf16 boxes (32-byte swizzle)
constexpr int BAR_BYTES = 256;
constexpr int TAB1_BYTES = 2 * INTER * 4;        // colc[1536] | cold[1536]: the folded LN_1 terms of W1, resident
constexpr int TAB2_BYTES = 2 * HIDDEN * 4;       // gamma_1[384] | cold2[384]: the residual epilogue's column terms, resident
constexpr int TMEM_COLS = 512;
constexpr int OUT_COL = 0;
constexpr int HACC_COL = HIDDEN;                 // two buffers of 64 columns
constexpr int RES_CHUNK = 20;                    // the chunk at which the epilogue threads copy their residual values

constexpr size_t smem_bytes() {
  return (size_t)X_BYTES + 2 * W_SLOT_BYTES + EPI_WARPS * STAGING_BYTES + BAR_BYTES + TAB1_BYTES + TAB2_BYTES + 1024 /*alignment*/;
}
static_assert(smem_bytes() <= 227 * 1024, "fused MLP exceeds the shared memory of an SM");

struct MlpParams {
  int M;                       // valid token rows
  const float* up_c;           // [1536] c_n of the folded LN_1 (drag_gemm.cuh)
  const float* up_d;           // [1536] d_n
  const float* down_cold;      // [384] b_2 + beta_1
  const float* down_gamma;     // [384] gamma_1
  const float2* in_stats;      // [M][STATS_PARTS] partial (sum, sum^2) of the X rows
  float2* out_stats;           // [M][STATS_PARTS] of the OUT rows
  float inv_width, ln_eps;
  long long* trace;            // TRACE instantiation (DRAG_MLP_TRACE): clock sums of the waits of CTA 0's issuer [0..7] and first epilogue warp [8..19]
  int dbg;                     // probes only (DRAG_MLP_DBG): 1 = E1 without the LayerNorm fold / GELU arithmetic, 2 = E1 without the tensor-memory load as well, 4 = G1 strictly two chunks ahead across tile boundaries
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
               ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(taddr)
               : "memory");
}
// D[tmem of both CTAs] (+)= A[tmem: 16-bit pairs, one column per two k] . B[smem desc] over a CTA pair
__device__ __forceinline__ void umma_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// tmap_x  : X   [T, 384]   bf16, box 64 columns x 128 rows, 128-byte swizzle   (the A operand and the residual)
// tmap_w1 : W1g [1536, 384] bf16, box 64 x 32
// tmap_w2 : W2  [384, 1536] fp16 (moved as 16-bit words), box 64 x 96
// tmap_out: OUT [T, 384]   bf16, box 16 columns x 32 rows, 32-byte swizzle
// grid = 2 * min(#SMs / 2, ceil(T / 256)) as clusters of 2, block = THREADS, dynamic smem = smem_bytes()
template <bool TRACE = false>
__global__ void __launch_bounds__(THREADS, 1)
mlp_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
           const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_out, MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = smem;
  uint8_t* wslot = smem + X_BYTES;
  uint8_t* staging = wslot + 2 * W_SLOT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + EPI_WARPS * STAGING_BYTES);
  uint64_t* x_full = bars;            // leader: both CTAs' X rows have landed
  uint64_t* x_empty = bars + 1;       // G1 of the tile's last chunk has retired and the epilogue warps hold their residual values
  uint64_t* w1_full = bars + 2;       // [2] leader
  uint64_t* w1_empty = bars + 4;      // [2]
  uint64_t* w2_full = bars + 6;       // [2] leader
  uint64_t* w2_empty = bars + 8;      // [2]
  uint64_t* hacc_full = bars + 10;    // [2] G1(c) has retired
  uint64_t* hp_full = bars + 12;      // [2] leader: the 32 E1 warps of the pair have written h(c)
  uint64_t* out_full = bars + 14;     // G2 of the tile's last chunk has retired
  uint64_t* out_free = bars + 15;     // leader: the 16 epilogue warps have read OUT
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);
  float* tab1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);   // colc[1536] | cold[1536]
  float* tab2 = tab1 + 2 * INTER;                                                          // gamma[384] | cold2[384]
  static_assert(17 * 8 <= BAR_BYTES, "barrier region too small");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();   // 0 = pair leader
  const int pair = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
  const int m_tiles = (p.M + 2 * ROWS - 1) / (2 * ROWS);
  const int my_tiles = pair < m_tiles ? (m_tiles - pair + n_pairs - 1) / n_pairs : 0;
  const int total = my_tiles * CHUNKS;            // the pair's chunk sequence g = it * 24 + c

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_x);
    tc::tma_prefetch_desc(&tmap_w1);
    tc::tma_prefetch_desc(&tmap_w2);
    tc::tma_prefetch_desc(&tmap_out);
    tc::mbar_init(x_full, 1);
    tc::mbar_init(x_empty, 1 + EPI_WARPS);
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&w1_full[b], 1);
      tc::mbar_init(&w1_empty[b], 1);
      tc::mbar_init(&w2_full[b], 1);
      tc::mbar_init(&w2_empty[b], 1);
      tc::mbar_init(&hacc_full[b], 1);
      tc::mbar_init(&hp_full[b], 2 * E1_WARPS);
    }
    tc::mbar_init(out_full, 1);
    tc::mbar_init(out_free, 2 * EPI_WARPS);
    tc::fence_barrier_init();
  }
  // the per-column epilogue terms (weights-side constants: not written by any kernel of the forward), once
  for (int i = (int)threadIdx.x; i < INTER / 4; i += THREADS) {
    reinterpret_cast<float4*>(tab1)[i] = __ldg(reinterpret_cast<const float4*>(p.up_c) + i);
    reinterpret_cast<float4*>(tab1 + INTER)[i] = __ldg(reinterpret_cast<const float4*>(p.up_d) + i);
  }
  for (int i = (int)threadIdx.x; i < HIDDEN / 4; i += THREADS) {
    reinterpret_cast<float4*>(tab2)[i] = __ldg(reinterpret_cast<const float4*>(p.down_gamma) + i);
    reinterpret_cast<float4*>(tab2 + HIDDEN)[i] = __ldg(reinterpret_cast<const float4*>(p.down_cold) + i);
  }
  tc::cluster_sync_all();   // the peer's barriers must exist before anything signals them
  if (warp == 1) {
    tc::tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS);
    tc::tmem_relinquish_pair();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  tc::pdl_launch_dependents();
  tc::pdl_wait();   // everything above overlapped the previous kernel's tail

  // E1 of one warp: its 16 hidden units of chunk g.  h = gelu(rstd * (acc - mu * colc) + cold) as fp16 pairs, written over the
  // first 8 of the 16 accumulator columns just read: the A operand of G2's k-step `cg`
  auto e1_chunk = [&](int g, int cg, uint32_t lane_base, uint32_t leader_hp_full, float mu, float rstd) {
    const int c = g % CHUNKS, b = g & 1;
    const float* tab_c = tab1 + c * NC + cg * E1_COLS;
    const float* tab_d = tab_c + INTER;
    tc::mbar_wait(&hacc_full[b], (uint32_t)(g >> 1) & 1);
    tc::tc_fence_after();
    const uint32_t t_h = lane_base + (uint32_t)(HACC_COL + b * NC + cg * E1_COLS);
    uint32_t r[16];
    if (!(p.dbg & 2)) {
      tmem_ld16(t_h, r);
      tc::tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = 0x3c000000u + (uint32_t)i;
    }
    uint32_t o[8];
    if (p.dbg & 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (r[2 * i] >> 16) | (r[2 * i + 1] & 0xffff0000u);
    } else {
      const uint64_t nmu2 = gemm::pack_f32x2(-mu, -mu), rstd2 = gemm::pack_f32x2(rstd, rstd);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 cc = gemm::lds_f4(tab_c + 4 * i);
        const float4 cd = gemm::lds_f4(tab_d + 4 * i);
        const uint64_t v01 = gemm::fma_f32x2(rstd2, gemm::fma_f32x2(nmu2, gemm::pack_f32x2(cc.x, cc.y), gemm::pack_f32x2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]))),
                                             gemm::pack_f32x2(cd.x, cd.y));
        const uint64_t v23 = gemm::fma_f32x2(rstd2, gemm::fma_f32x2(nmu2, gemm::pack_f32x2(cc.z, cc.w), gemm::pack_f32x2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]))),
                                             gemm::pack_f32x2(cd.z, cd.w));
        float v0, v1, v2, v3;
        gemm::unpack_f32x2(v01, v0, v1);
        gemm::unpack_f32x2(v23, v2, v3);
        o[2 * i] = gemm::gelu_f16x2(v0, v1);
        o[2 * i + 1] = gemm::gelu_f16x2(v2, v3);
      }
    }
    tmem_st8(t_h, o);
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive_cluster(leader_hp_full + (uint32_t)(b * 8));
  };

  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONTROL));
   if (warp == 0) {
    // ===================== TMA producer: this CTA's rows of X, its halves of the weight chunks =====================
    if (tc::elect_one()) {
      const uint32_t bar_x = tc::mapa_shared(tc::smem_u32(x_full), 0);
      for (int it = 0; it < my_tiles; ++it) {
        const int row0 = (pair + it * n_pairs) * 2 * ROWS + (int)rank * ROWS;
        tc::mbar_wait(x_empty, (uint32_t)(it & 1) ^ 1);
        if (rank == 0) tc::mbar_arrive_expect_tx(x_full, 2 * X_BYTES);
        for (int kb = 0; kb < KB; ++kb) tc::tma_load_2d_pair(&tmap_x, bar_x, xs + (size_t)kb * X_KB_BYTES, kb * 64, row0);
        for (int c = 0; c < CHUNKS; ++c) {
          const int g = it * CHUNKS + c, b = g & 1;
          const uint32_t ph = (uint32_t)(g >> 1) & 1;
          uint8_t* slot = wslot + (size_t)b * W_SLOT_BYTES;
          tc::mbar_wait(&w1_empty[b], ph ^ 1);
          if (rank == 0) tc::mbar_arrive_expect_tx(&w1_full[b], 2 * W1_BYTES);
          const uint32_t bar1 = tc::mapa_shared(tc::smem_u32(&w1_full[b]), 0);
          for (int kb = 0; kb < KB; ++kb)
            tc::tma_load_2d_pair(&tmap_w1, bar1, slot + (size_t)kb * W1_KB_BYTES, kb * 64, c * NC + (int)rank * (NC / 2));
          tc::mbar_wait(&w2_empty[b], ph ^ 1);
          if (rank == 0) tc::mbar_arrive_expect_tx(&w2_full[b], 2 * W2_BYTES);
          const uint32_t bar2 = tc::mapa_shared(tc::smem_u32(&w2_full[b]), 0);
          for (int j = 0; j < 2; ++j)
            tc::tma_load_2d_pair(&tmap_w2, bar2, slot + W1_BYTES + (size_t)j * W2_HALF_BYTES, c * NC, j * OUT_HALF + (int)rank * (OUT_HALF / 2));
        }
      }
    }
    __syncwarp();
   } else if (warp == 1) {
    // ===================== MMA issuer (pair leader):  G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) | ... =====================
    if (rank == 0 && tc::elect_one()) {
      constexpr uint32_t idesc1 = tc::umma_idesc_bf16(2 * ROWS, NC);
      constexpr uint32_t idesc2 = tc::umma_idesc_f16(2 * ROWS, OUT_HALF);
      const uint32_t xs_addr = tc::smem_u32(xs), w_addr = tc::smem_u32(wslot);
      long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const bool tr = TRACE && p.trace != nullptr && blockIdx.x == 0;
      const long long t_begin = TRACE ? clock64() : 0;
      auto timed_wait = [&](uint64_t* bar, uint32_t parity, int slot) {
        if (TRACE && tr) {
          const long long t0 = clock64();
          tc::mbar_wait(bar, parity);
          tw[slot] += clock64() - t0;
        } else {
          tc::mbar_wait(bar, parity);
        }
      };
      auto g1 = [&](int g) {
        const int it = g / CHUNKS, c = g - it * CHUNKS, b = g & 1;
        if (c == 0) timed_wait(x_full, (uint32_t)it & 1, 0);
        timed_wait(&w1_full[b], (uint32_t)(g >> 1) & 1, 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(HACC_COL + b * NC);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t a_desc = tc::umma_desc_sw128(xs_addr + (uint32_t)(kb * X_KB_BYTES));
          const uint64_t b_desc = tc::umma_desc_sw128(w_addr + (uint32_t)(b * W_SLOT_BYTES + kb * W1_KB_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16_pair(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc1, (kb | k) != 0 ? 1u : 0u);
        }
        tc::umma_commit_pair(&w1_empty[b], 3);
        tc::umma_commit_pair(&hacc_full[b], 3);
        if (c == CHUNKS - 1) tc::umma_commit_pair(x_empty, 3);
      };
      auto g2 = [&](int g) {
        const int it = g / CHUNKS, c = g - it * CHUNKS, b = g & 1;
        const uint32_t ph = (uint32_t)(g >> 1) & 1;
        timed_wait(&w2_full[b], ph, 2);
        timed_wait(&hp_full[b], ph, 3);
        if (c == 0) timed_wait(out_free, ((uint32_t)it & 1) ^ 1, 4);
        tc::tc_fence_after();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint64_t b_desc = tc::umma_desc_sw128(w_addr + (uint32_t)(b * W_SLOT_BYTES + W1_BYTES + j * W2_HALF_BYTES));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            // h of hidden units [16 kk, 16 kk + 16): 8 packed columns at the start of the 16 accumulator columns they were
            // computed from (one E1 warp per lane quarter wrote them)
            const uint32_t a_tmem = tmem_base + (uint32_t)(HACC_COL + b * NC + kk * E1_COLS);
            umma_ts_pair(tmem_base + (uint32_t)(OUT_COL + j * OUT_HALF), a_tmem, b_desc + (uint64_t)(kk * 2), idesc2, (c | kk) != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit_pair(&w2_empty[b], 3);
        if (c == CHUNKS - 1) tc::umma_commit_pair(out_full, 3);
      };
      if (total > 0) {
        g1(0);
        g1(1);
        for (int g = 0; g < total; ++g) {
          g2(g);
          // G1 runs two chunks ahead; the first two G1 of the NEXT tile wait for its X rows, which can only be loaded once
          // this tile's last G1 has retired: they go out after this tile's last G2, so that E2 starts in the meantime
          const int c = g % CHUNKS;
          if (!(p.dbg & 4)) {
            if (c == CHUNKS - 2) continue;
            if (c == CHUNKS - 1 && g + 1 < total) g1(g + 1);
          }
          if (g + 2 < total) g1(g + 2);
        }
      }
      if (TRACE && tr) {
        tw[7] = clock64() - t_begin;
        for (int i = 0; i < 8; ++i) p.trace[i] = tw[i];
      }
    }
    __syncwarp();
   }
  } else if (warp >= 4 + EPI_WARPS) {
    // ===================== epilogue class B: E1 only =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_B));
    const int quarter = warp & 3;
    const int cg = 2 + ((warp - 4 - EPI_WARPS) >> 2);
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t leader_hp_full = tc::mapa_shared(tc::smem_u32(&hp_full[0]), 0);
    float mu = 0.f, rstd = 0.f;
    for (int g = 0; g < total; ++g) {
      const int it = g / CHUNKS;
      if (g - it * CHUNKS == 0) {
        const int row = (pair + it * n_pairs) * 2 * ROWS + (int)rank * ROWS + row_in_tile;
        gemm::row_stats(p.in_stats, row, row < p.M, p.inv_width, p.ln_eps, mu, rstd);
      }
      e1_chunk(g, cg, lane_base, leader_hp_full, mu, rstd);
    }
  } else {
    // ===================== epilogue class A: E1 per chunk, E2 per tile =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_A));
    const int ew = warp - 4;
    const int quarter = warp & 3;             // TMEM lanes this warp may touch: [32 * quarter, +32)
    const int half = ew >> 2;                 // E1: hidden units [16 half, +16) of every chunk; E2: output columns [192 half, +192)
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t leader_hp_full = tc::mapa_shared(tc::smem_u32(&hp_full[0]), 0);
    const uint32_t leader_out_free = tc::mapa_shared(tc::smem_u32(out_free), 0);
    uint8_t* stage_buf = staging + (size_t)ew * STAGING_BYTES;
    // 32-byte swizzle: 16-byte chunk j of row r lives at chunk j ^ ((r >> 2) & 1)
    const uint32_t stage_row = tc::smem_u32(stage_buf) + (uint32_t)lane * 32u;
    const uint32_t stage_swz = (uint32_t)((lane >> 2) & 1) * 16u;
    const uint32_t x_row = tc::smem_u32(xs) + (uint32_t)row_in_tile * 128u;

    auto tile_stats = [&](int it, float& mu, float& rstd) {
      const int row = (pair + it * n_pairs) * 2 * ROWS + (int)rank * ROWS + row_in_tile;
      gemm::row_stats(p.in_stats, row, row < p.M, p.inv_width, p.ln_eps, mu, rstd);
    };

    long long te[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const bool tr = TRACE && p.trace != nullptr && blockIdx.x == 0 && ew == 0;
    const long long t_begin = TRACE ? clock64() : 0;
    uint32_t res[OUT_HALF / 2];   // this row's raw X values of the thread's 192 output columns (bf16 pairs)
    float mu = 0.f, rstd = 0.f, mu_out = 0.f, rstd_out = 0.f;

    // E2 of tile `it`: OUT + cold2 + LN_1(X) -> bf16 + row statistics, in steps of 16 columns
    auto finish_tile = [&](int it) {
      const int row0 = (pair + it * n_pairs) * 2 * ROWS + (int)rank * ROWS;
      const int row = row0 + row_in_tile;
      const bool row_ok = row < p.M;
      long long t0 = (TRACE && tr) ? clock64() : 0;
      tc::mbar_wait(out_full, (uint32_t)it & 1);
      tc::tc_fence_after();
      if (TRACE && tr) { const long long t1 = clock64(); te[4] += t1 - t0; t0 = t1; }
      const uint32_t t_row = lane_base + (uint32_t)(OUT_COL + half * OUT_HALF);
      const uint64_t rstd2 = gemm::pack_f32x2(rstd_out, rstd_out);
      const uint64_t nmr2 = gemm::pack_f32x2(-mu_out * rstd_out, -mu_out * rstd_out);
      uint64_t sum2 = gemm::pack_f32x2(0.f, 0.f), sq2 = sum2;
      constexpr int STEPS = OUT_HALF / STORE_COLS;   // 12
      uint32_t r[16], r_next[16];
      tmem_ld16(t_row, r);
#pragma unroll
      for (int st = 0; st < STEPS; ++st) {
        tc::tmem_ld_wait();
        if (st + 1 < STEPS) {
          tmem_ld16(t_row + (st + 1) * STORE_COLS, r_next);   // in flight during this step's arithmetic
        } else {
          // all of this warp's OUT columns are in registers: the next tile's G2 may overwrite them
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster(leader_out_free);
        }
        if (TRACE && tr) { const long long t1 = clock64(); te[5] += t1 - t0; t0 = t1; }
        const int col = half * OUT_HALF + st * STORE_COLS;
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 ga = gemm::lds_f4(tab2 + col + 4 * i);
          const float4 cd = gemm::lds_f4(tab2 + HIDDEN + col + 4 * i);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            // two columns at a time on fp32 pairs: v = gamma * ((x - mu) * rstd) + (acc + cold2)
            const uint32_t w = res[st * 8 + 2 * i + h2];
            const uint64_t x2 = gemm::pack_f32x2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
            const uint64_t g2v = h2 == 0 ? gemm::pack_f32x2(ga.x, ga.y) : gemm::pack_f32x2(ga.z, ga.w);
            const uint64_t d2v = h2 == 0 ? gemm::pack_f32x2(cd.x, cd.y) : gemm::pack_f32x2(cd.z, cd.w);
            const uint64_t a2 = gemm::pack_f32x2(__uint_as_float(r[4 * i + 2 * h2]), __uint_as_float(r[4 * i + 2 * h2 + 1]));
            const uint64_t v2 = gemm::fma_f32x2(g2v, gemm::fma_f32x2(x2, rstd2, nmr2), add_f32x2(a2, d2v));
            sum2 = add_f32x2(sum2, v2);
            sq2 = gemm::fma_f32x2(v2, v2, sq2);
            float v0, v1;
            gemm::unpack_f32x2(v2, v0, v1);
            o[2 * i + h2] = gemm::pack_bf16(v0, v1);
          }
        }
        if (TRACE && tr) { const long long t1 = clock64(); te[6] += t1 - t0; t0 = t1; }
        // two staging boxes alternate: the store that read this one two steps ago must be done with it
        if (st >= 2) {
          if (lane == 0) tc::tma_store_wait_read_1();
          __syncwarp();
        }
        const uint32_t srow = stage_row + (uint32_t)((st & 1) * (STAGING_BYTES / 2));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + stage_swz), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (stage_swz ^ 16u)), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
        tc::fence_proxy_async();
        __syncwarp();
        if (tc::elect_one()) {
          tc::tma_store_2d(&tmap_out, stage_buf + (size_t)(st & 1) * (STAGING_BYTES / 2), col, row0 + quarter * 32);
          tc::tma_store_commit();
        }
        if (st + 1 < STEPS) {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = r_next[i];
        }
        if (TRACE && tr) { const long long t1 = clock64(); te[8] += t1 - t0; t0 = t1; }
      }
      if (row_ok) {
        float s0, s1, q0, q1;
        gemm::unpack_f32x2(sum2, s0, s1);
        gemm::unpack_f32x2(sq2, q0, q1);
        p.out_stats[(size_t)row * gemm::STATS_PARTS + half] = make_float2(s0 + s1, q0 + q1);
        if (half == 0) p.out_stats[(size_t)row * gemm::STATS_PARTS + 2] = make_float2(0.f, 0.f);
      }
      // the staging boxes are rewritten by the next tile's E2 only: drain the reads now
      if (lane == 0) tc::tma_store_wait_read();
      __syncwarp();
      if (TRACE && tr) { const long long t1 = clock64(); te[9] += t1 - t0; }
    };

    for (int g = 0; g < total; ++g) {
      const int it = g / CHUNKS, c = g - it * CHUNKS;
      if (c == 0) {
        if (it > 0) { mu_out = mu; rstd_out = rstd; }
        tile_stats(it, mu, rstd);
      }
      long long t0 = (TRACE && tr) ? clock64() : 0;
      if (c == 0 && it > 0) finish_tile(it - 1);   // G2 of the previous tile's last chunk is issued before this chunk's G1
      if (TRACE && tr) { const long long t1 = clock64(); te[3] += t1 - t0; t0 = t1; }
      e1_chunk(g, half, lane_base, leader_hp_full, mu, rstd);
      if (TRACE && tr) { const long long t1 = clock64(); te[2] += t1 - t0; t0 = t1; }

      if (c == RES_CHUNK) {
        // the row's residual values (columns [192 half, +192)) out of the X tile; the tile may then be refilled.
        // 16-byte chunk j of row r of a k-block lives at chunk j ^ (r & 7)
#pragma unroll
        for (int q = 0; q < OUT_HALF / 8; ++q) {
          const int kb = half * 3 + q / 8;
          const uint32_t chunk = (uint32_t)((q & 7) ^ (row_in_tile & 7));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(res[4 * q]), "=r"(res[4 * q + 1]), "=r"(res[4 * q + 2]), "=r"(res[4 * q + 3])
                       : "r"(x_row + (uint32_t)(kb * X_KB_BYTES) + chunk * 16u) : "memory");
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(x_empty);
      }
    }
    if (my_tiles > 0) {
      mu_out = mu;
      rstd_out = rstd;
      finish_tile(my_tiles - 1);
    }
    if (lane == 0) tc::tma_store_wait_all();
    if (TRACE && tr && lane == 0) {
      te[7] = clock64() - t_begin;
      for (int i = 0; i < 12; ++i) p.trace[8 + i] = te[i];
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();   // the peer may still signal this CTA's barriers / read its operands
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

}  // namespace mlp
}  // namespace drag
