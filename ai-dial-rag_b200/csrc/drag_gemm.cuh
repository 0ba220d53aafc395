// Persistent warp-specialised tcgen05 GEMM with fused epilogues for the BERT encoder
// (SURVEY.md 8a row a5):   OUT[M, N] = epilogue( A[M, K] . W[N, K]^T + bias )
//
//   A  : bf16 activations, row-major [M, K]          (K-major operand)
//   W  : bf16 weights exactly as HF stores them, [N, K] row-major (K-major operand)
//   acc: fp32 in TMEM
//
// CTA = 2 + EPI_WARPS warps, one CTA per SM, looping over 128 x BLOCK_N output tiles:
//   warp 0   TMA producer: cp.async.bulk.tensor loads of the A tile (128 x 64) and the W tile
//            (BLOCK_N x 64) into a STAGES-deep 128B-swizzled shared-memory ring
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M=128, N=BLOCK_N or 2 x BLOCK_N/2, K=16)
//            per 16-wide k-step; tcgen05.commit releases ring slots / publishes the accumulator
//   warps 2+ epilogue: tcgen05.ld the accumulator (thread = one row, 32 columns at a time) and
//            apply   EPI_BIAS          -> +bias                              (QKV projection)
//                    EPI_BIAS_GELU     -> +bias, erf-GELU                    (FFN up)
//                    EPI_BIAS_RES_LN   -> +bias +residual, LayerNorm over the full 384-wide row
//            then store bf16.  With 2*BLOCK_N <= 512 TMEM columns the accumulator is double
//            buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once

#include <cuda_bf16.h>

#include "drag_tc.cuh"

namespace drag {
namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;

enum { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RES_LN = 2 };

struct GemmParams {
  int M, N, K;                     // M = valid rows (tokens)
  const float* bias;               // [N]
  const float* gamma;              // [N]  (LN epilogue)
  const float* beta;               // [N]
  float ln_eps;
  const __nv_bfloat16* residual;   // [M, N] (LN epilogue)
  __nv_bfloat16* out;              // [M, N]
  float* out_f32;                  // optional fp32 copy of the output (debug taps), may be null
};

template <int BLOCK_N>
struct Cfg {
  static constexpr int UMMA_N = BLOCK_N > 256 ? BLOCK_N / 2 : BLOCK_N;
  static constexpr int N_SPLIT = BLOCK_N / UMMA_N;
  static constexpr int ACC_STAGES = (2 * BLOCK_N <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = (ACC_STAGES * BLOCK_N <= 128) ? 128 : (ACC_STAGES * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static_assert(UMMA_N % 16 == 0 && UMMA_N >= 16 && UMMA_N <= 256, "invalid UMMA N");
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024-byte alignment");
};

template <int BLOCK_N, int STAGES>
constexpr size_t smem_bytes() {
  // ring + barriers/tmem pointer/LN partials + slack for manual 1024-byte alignment
  return (size_t)STAGES * Cfg<BLOCK_N>::STAGE_BYTES + 4096 + 1024;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int BLOCK_N, int EPI, int EPI_WARPS, int STAGES>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, GemmParams p) {
  using C = Cfg<BLOCK_N>;
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "epilogue uses 4 or 8 warps");
  constexpr int COL_GROUPS = EPI_WARPS / 4;
  constexpr int COLS_PER_THREAD = BLOCK_N / COL_GROUPS;
  static_assert(COLS_PER_THREAD % 32 == 0, "epilogue works in 32-column chunks");
  static_assert(EPI != EPI_BIAS_RES_LN || C::ACC_STAGES == 1, "LN epilogue rewrites the accumulator in place");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float2* ln_part = reinterpret_cast<float2*>(tmem_ptr_smem + 4);  // [COL_GROUPS][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = p.N / BLOCK_N;
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_a);
    tc::tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&tmem_full_bar[s], 1);
      tc::mbar_init(&tmem_empty_bar[s], EPI_WARPS);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = ring + (size_t)stage * C::STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_STAGE_BYTES;
          tc::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          tc::tma_load_2d(&tmap_a, &full_bar[stage], a_dst, kb * BLOCK_K, m_blk * BLOCK_M);
#pragma unroll
          for (int h = 0; h < C::N_SPLIT; ++h)
            tc::tma_load_2d(&tmap_w, &full_bar[stage], b_dst + (size_t)h * C::UMMA_N * BLOCK_K * 2, kb * BLOCK_K,
                            n_blk * BLOCK_N + h * C::UMMA_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(BLOCK_M, C::UMMA_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        tc::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t a_addr = tc::smem_u32(ring + (size_t)stage * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
          const uint64_t a_desc = tc::umma_desc_sw128(a_addr);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
#pragma unroll
            for (int h = 0; h < C::N_SPLIT; ++h) {
              const uint64_t b_desc = tc::umma_desc_sw128(b_addr + (uint32_t)(h * C::UMMA_N * BLOCK_K * 2));
              // +32 bytes per 16-element k-step inside the 128-byte swizzle row (encoded >> 4)
              tc::umma_bf16(d_tmem + (uint32_t)(h * C::UMMA_N), a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2),
                            idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          tc::umma_commit(&empty_bar[stage]);  // ring slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;           // TMEM lanes this warp may touch: [32*quarter, +32)
    const int col_group = ew >> 2;
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int row = m_blk * BLOCK_M + row_in_tile;
      const bool row_ok = row < p.M;
      const int col0 = n_blk * BLOCK_N + col_group * COLS_PER_THREAD;
      tc::mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N + col_group * COLS_PER_THREAD);
      __nv_bfloat16* out_row = p.out + (size_t)row * p.N + col0;

      if (EPI == EPI_BIAS_RES_LN) {
        const __nv_bfloat16* res_row = p.residual + (size_t)row * p.N + col0;
        float s = 0.f, ss = 0.f;
#pragma unroll 1
        for (int c = 0; c < COLS_PER_THREAD; c += 32) {
          uint32_t r[32];
          tc::tmem_ld32(t_row + c, r);
          uint4 rv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            rv[i] = row_ok ? __ldg(reinterpret_cast<const uint4*>(res_row + c) + i) : make_uint4(0, 0, 0, 0);
          tc::tmem_ld_wait();
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(rv);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float r0 = __uint_as_float(rw[i] << 16), r1 = __uint_as_float(rw[i] & 0xffff0000u);
            float v0 = __uint_as_float(r[2 * i]) + __ldg(p.bias + col0 + c + 2 * i) + r0;
            float v1 = __uint_as_float(r[2 * i + 1]) + __ldg(p.bias + col0 + c + 2 * i + 1) + r1;
            s += v0 + v1;
            ss = fmaf(v0, v0, fmaf(v1, v1, ss));
            r[2 * i] = __float_as_uint(v0);
            r[2 * i + 1] = __float_as_uint(v1);
          }
          tc::tmem_st32(t_row + c, r);
        }
        tc::tmem_st_wait();
        if (COL_GROUPS > 1) {
          ln_part[col_group * 128 + row_in_tile] = make_float2(s, ss);
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
          float2 o = ln_part[(col_group ^ 1) * 128 + row_in_tile];
          s += o.x;
          ss += o.y;
        }
        const float inv_n = 1.0f / (float)BLOCK_N;
        const float mean = s * inv_n;
        const float var = fmaxf(ss * inv_n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.ln_eps);
#pragma unroll 1
        for (int c = 0; c < COLS_PER_THREAD; c += 32) {
          uint32_t r[32];
          tc::tmem_ld32(t_row + c, r);
          tc::tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int cc = col0 + c + 2 * i;
            float v0 = (__uint_as_float(r[2 * i]) - mean) * rstd * __ldg(p.gamma + cc) + __ldg(p.beta + cc);
            float v1 = (__uint_as_float(r[2 * i + 1]) - mean) * rstd * __ldg(p.gamma + cc + 1) + __ldg(p.beta + cc + 1);
            o[i] = pack_bf16(v0, v1);
            if (p.out_f32 && row_ok) {
              p.out_f32[(size_t)row * p.N + cc] = v0;
              p.out_f32[(size_t)row * p.N + cc + 1] = v1;
            }
          }
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<uint4*>(out_row + c)[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
        }
        if (COL_GROUPS > 1) asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");  // ln_part reusable
      } else {
#pragma unroll 1
        for (int c = 0; c < COLS_PER_THREAD; c += 32) {
          uint32_t r[32];
          tc::tmem_ld32(t_row + c, r);
          tc::tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int cc = col0 + c + 2 * i;
            float v0 = __uint_as_float(r[2 * i]) + __ldg(p.bias + cc);
            float v1 = __uint_as_float(r[2 * i + 1]) + __ldg(p.bias + cc + 1);
            if (EPI == EPI_BIAS_GELU) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
            o[i] = pack_bf16(v0, v1);
          }
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<uint4*>(out_row + c)[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
        }
      }
      // all TMEM reads of this warp are complete -> hand the accumulator back to the MMA warp
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace gemm
}  // namespace drag
