// Multi-head self-attention on tcgen05 for the packed (padding-free) batch, head_dim = 32
// (SURVEY.md 8a row a5:  softmax(Q K^T / sqrt(32) + key mask) V ).  Third design.
//
// At head_dim 32 the op is bound by the exponentials (one per 128 flop; MUFU: 16 per clock per SM, measured), so
// the kernel is built to keep the MUFU pipe fed: everything a softmax warp does besides exponentials was measured
// with an in-kernel timeline (scripts/attn_trace.py) and removed or shortened.
//
// Persistent kernel, one CTA per SM, 11 working warps (12 launched: register budgets are set per warpgroup):
//   warps 0-3   softmax group 0        thread = one query row of the group's current 128-query job
//   warps 4-7   softmax group 1        (jobs alternate between the groups)
//   warp  8     TMA producer: Q, K and V tiles (128 tokens x one head, 64-byte swizzle) of the next units
//               straight out of the packed [T, 3*hidden] QKV buffer into a 2..4-stage ring
//   warps 9,10  MMA issuers, one per softmax group.  The whole warp walks the job sequence (warp-uniform values:
//               descriptors live in uniform registers) and one lane picked by elect.sync issues -- `if (lane == 0)`
//               makes ptxas wrap every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA loop, and a P V step is only
//               16 clocks of tensor work, so the cost of ISSUING matters here.
// The producer publishes a small descriptor of every unit in shared memory; all walkers run on those.
// A unit = one (sequence, head) item, or 1/split of its query tiles when few items are in flight (the query
// path); a job = one 128-query tile of a unit against all keys of the sequence, in 128-key blocks.
//
// TMEM (256 columns per group):  S [0,128)  scores of the next block | P [128,192)  bf16 weights, two per column |
//                                O [192,224) and [224,256)  output accumulators of the group's even / odd jobs
// Per key block n of group g:
//   S(n) = Q K^T               tcgen05.mma M=128 N=128 K=16 x2, issued as soon as S(n-1) is in registers
//   softmax thread             ONE tcgen05.ld pass (128 scores -> registers, S handed back at once); row maximum
//                              (8 chains of 3-input maxima), lazily raised reference m (only when a score exceeds it by
//                              2^8: the softmax is invariant to the reference), p = 2^(s c - m c) (FFMA2 + MUFU.EX2),
//                              row sum in registers, bf16 pairs by add-half-ulp + PRMT (the F2FP conversion runs at a
//                              third of the ALU rate, measured); then one barrier wait -- inside a job for S(n+1), whose
//                              completion also says that P(n-1) V has retired (same issuing thread, issued earlier), at
//                              a job's last block for P(n-1) V itself: the P columns are free and O is complete up to
//                              block n-1 -- and P goes to TMEM with two tcgen05.st (no shared-memory round trip, no
//                              proxy fence)
//   O (+)= P(n) V              tcgen05.mma M=128 N=32 K=16 x8 with A FROM TMEM; V is consumed as stored ([key][32]:
//                              an MN-major B operand).  O stays in TMEM across the key blocks of a job; a raised
//                              reference rescales it in place (rare).
//   end of job                 O / l -> bf16 -> ctx (rows inside the sequence only), DEFERRED into the next block's
//                              barrier wait (the accumulators alternate between jobs), so nobody waits for a P V
// Keys beyond the sequence do not exist in the packed layout: the tail of the last block is masked before the
// row maximum.  Every row's result depends on that row's scores only (no cross-row decision), so it does not
// depend on the batch composition.
#pragma once

#include <cuda_bf16.h>

#include "drag_tc.cuh"

namespace drag {
namespace attn3 {

constexpr int HEAD_DIM = 32;
constexpr int TILE = 128;                               // queries per job = keys per block
constexpr int QKV_TILE_BYTES = TILE * HEAD_DIM * 2;     // 8 KB: [128][32] bf16, 64-byte swizzle
constexpr int GROUPS = 2;
constexpr int SOFTMAX_WARPS = 4 * GROUPS;
constexpr int TMA_WARP = SOFTMAX_WARPS;
constexpr int MMA_WARP = TMA_WARP + 1;                  // issuer of group 0; MMA_WARP + 1: group 1
constexpr int THREADS = 384;                            // 3 warpgroups: softmax 0, softmax 1, {TMA, 2 x MMA, one idle warp}
constexpr int SOFTMAX_REGS = 200;                       // setmaxnreg: 8 x 32 x 200 + 4 x 32 x 104 = 64 512 = 384 x 168
constexpr int OTHER_REGS = 104;
constexpr int TMEM_COLS = 512;
constexpr int GROUP_COLS = 256;                         // per group: S [0,128) | P [128,192) | O even jobs [192,224) | O odd jobs [224,256)
constexpr int P_COL = TILE;
constexpr int O_COL = TILE + TILE / 2;
constexpr int MAX_STAGES = 4;
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int BAR_BYTES = 512;
constexpr float LAZY_LOG2 = 8.f;                        // numerators stay <= 2^8

// ring stages (no P buffers in shared memory any more: even 512-token units get two)
__host__ __device__ inline int unit_stages(int max_len) {
  const int tiles = (max_len + TILE - 1) / TILE;
  const int fit = (SMEM_BUDGET - 1024 - BAR_BYTES) / (3 * tiles * QKV_TILE_BYTES);
  return fit < MAX_STAGES ? fit : MAX_STAGES;   // 4, 4, 3, 2 stages for 1..4 tiles
}
__host__ __device__ inline size_t smem_bytes(int max_len) {
  const int tiles = (max_len + TILE - 1) / TILE;
  return (size_t)unit_stages(max_len) * 3 * tiles * QKV_TILE_BYTES + 1024 /*alignment*/ + BAR_BYTES;
}
// few items in flight (the query path): an item's query tiles are dealt to `split` units so that the launch fills the GPU
__host__ __device__ inline int pick_split(int n_items, int max_len, int sms) {
  const int tiles = (max_len + TILE - 1) / TILE;
  int split = 1;
  while (split < tiles && n_items * split < sms) ++split;
  return split;
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
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (bf16 pairs, one 32-bit column per two k) comes from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
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
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// two non-negative fp32 -> bf16 pair, rounded half up: add half an ulp to the bits, keep the high halves (2 IADD + 1 PRMT
// on the ALU pipe instead of one F2FP at a third of its rate)
__device__ __forceinline__ uint32_t round_bf16x2(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}
// non-blocking probe (try_wait may suspend the thread)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(tc::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ int ld_acquire_shared(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(v) : "r"(tc::smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(p)), "r"(v) : "memory");
}

// ---- the softmax of one 128-key block of one row, scores in registers ----
__device__ __forceinline__ void load_scores(uint32_t s_tmem, uint32_t (&r)[TILE]) {
#pragma unroll
  for (int c = 0; c < TILE / 32; ++c) tc::tmem_ld32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
  tc::tmem_ld_wait();
}
__device__ __forceinline__ float row_max(const uint32_t (&r)[TILE]) {
  // eight independent chains of 3-input maxima
  float mx[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) mx[c] = max3(__uint_as_float(r[3 * c]), __uint_as_float(r[3 * c + 1]), __uint_as_float(r[3 * c + 2]));
#pragma unroll
  for (int i = 24; i + 15 < TILE; i += 16) {
#pragma unroll
    for (int c = 0; c < 8; ++c) mx[c] = max3(mx[c], __uint_as_float(r[i + 2 * c]), __uint_as_float(r[i + 2 * c + 1]));
  }
  // 24 + 16 * 6 = 120: eight scores left
#pragma unroll
  for (int c = 0; c < 4; ++c) mx[c] = max3(mx[c], __uint_as_float(r[120 + 2 * c]), __uint_as_float(r[121 + 2 * c]));
  return max3(max3(mx[0], mx[1], mx[2]), max3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7]));
}
// p = 2^(s c - m c) as bf16 pairs (two per P column); returns the row sum of the unrounded weights
// (the 32-key chunks of a sequence's last block that hold no key of the sequence are skipped: their P columns are zero)
__device__ __forceinline__ float exp_scores(const uint32_t (&r)[TILE], int valid, uint64_t scale2, uint64_t noff2, uint32_t (&pk)[TILE / 2]) {
  uint64_t l2a = pack_f32x2(0.f, 0.f), l2b = l2a;
#pragma unroll
  for (int c = 0; c < TILE / 32; ++c) {
    if (c * 32 < valid) {
#pragma unroll
      for (int i = c * 32; i < c * 32 + 32; i += 4) {
        float p0, p1, p2, p3;
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), scale2, noff2), p0, p1);
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), scale2, noff2), p2, p3);
        p0 = ex2(p0); p1 = ex2(p1); p2 = ex2(p2); p3 = ex2(p3);
        l2a = add_f32x2(l2a, pack_f32x2(p0, p1));
        l2b = add_f32x2(l2b, pack_f32x2(p2, p3));
        pk[i >> 1] = round_bf16x2(p0, p1);
        pk[(i >> 1) + 1] = round_bf16x2(p2, p3);
      }
    } else {
#pragma unroll
      for (int i = c * 16; i < c * 16 + 16; ++i) pk[i] = 0u;
    }
  }
  float la, lb, lc, ld;
  unpack_f32x2(l2a, la, lb);
  unpack_f32x2(l2b, lc, ld);
  return (la + lb) + (lc + ld);
}

// What the TMA producer publishes about every unit, in a ring of DESC_RING descriptors indexed by the unit's ordinal,
// BEFORE it waits for the unit's data; `published` counts them.  S = 0 marks the end of the CTA's stream.
// Jobs alternate between the groups, so a group's cursor passes over at most one unit without a job of its own before
// it is held up by its next job; the producer cannot run further ahead of a unit somebody still works in than the
// n_stages units the data ring holds.  A descriptor is therefore overwritten (DESC_RING = 2 * MAX_STAGES units later)
// only after every cursor has moved past it.
struct UnitDesc {
  int tok0;        // first token of the sequence
  int S;           // sequence length
  int head;
  int part;        // first query tile of the unit (then part + split, ...)
};
constexpr int DESC_RING = 2 * MAX_STAGES;

// The CTA's jobs in a fixed order: its units in ring order, per unit the query tiles part, part + split, ...; job ordinal
// n (CTA-wide) belongs to softmax group n & 1.  A cursor points at one job of one group.
struct Cursor {
  int it = 0;        // ordinal of the unit (-> ring stage it % n_stages, phase it / n_stages; descriptor it % DESC_RING)
  int job = -1;      // CTA-wide job ordinal of the current job
  int qt = 0, n_tiles = 0, S = 0, tok0 = 0, head = 0;
  bool have_unit = false, mine = false, done = false;
  // Moves to the next job of group g; `done` is set at the end of the stream.  BLOCKING: waits for the descriptors it
  // needs and returns true; otherwise returns false as soon as the next descriptor has not been published yet (the cursor
  // keeps its position: call again later).  on_skip(it) is called for every unit the cursor leaves without having found
  // a job of its group in it.  UNIFORM (whole-warp callers): decisions are taken for the warp as one and the descriptor
  // fields go through a warp broadcast, which makes them (and everything computed from them) warp-uniform for the compiler.
  template <bool UNIFORM, bool BLOCKING, typename OnSkip>
  __device__ __forceinline__ bool seek(int g, int split, const int* published, const UnitDesc* desc, OnSkip on_skip) {
    for (;;) {
      if (!have_unit) {
        if (BLOCKING) {
          const long long t0 = clock64();
          while (ld_acquire_shared(published) <= it)
            if (clock64() - t0 > 4000000000ll) { printf("drag_b200: attention unit descriptor wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x); __trap(); }
          if (UNIFORM) __syncwarp();
        } else {
          bool there = ld_acquire_shared(published) > it;
          if (UNIFORM) there = __all_sync(0xffffffffu, there);
          if (!there) return false;
        }
        const UnitDesc& d = desc[it % DESC_RING];
        S = d.S; tok0 = d.tok0; head = d.head;
        int part = d.part;
        if (UNIFORM) { S = __shfl_sync(0xffffffffu, S, 0); part = __shfl_sync(0xffffffffu, part, 0); }
        if (S == 0) { done = true; return true; }
        n_tiles = (S + TILE - 1) / TILE;
        qt = part - split;
        have_unit = true;
        mine = false;
      }
      qt += split;
      if (qt >= n_tiles) {
        if (!mine) on_skip(it);
        have_unit = false;
        ++it;
        continue;
      }
      ++job;
      if ((job & 1) == g) { mine = true; return true; }
    }
  }
  // does the group have another job in the current unit after this one?
  __device__ __forceinline__ bool last_of_group_in_unit(int split) const { return qt + 2 * split >= n_tiles; }
};

// Debug timeline (TRACE instantiation only, scripts/attn_trace.py): CTA 0 records clock64() at the phase boundaries of its
// first TRACE_BLOCKS key blocks -- per softmax warp 8 stamps per block, per issuer 4 stamps per block.
constexpr int TRACE_BLOCKS = 96;
constexpr int TRACE_WORDS = SOFTMAX_WARPS * TRACE_BLOCKS * 8 + GROUPS * TRACE_BLOCKS * 4;
__device__ long long* g_trace = nullptr;

// qkv : [T, 3*hidden] bf16 (tensor map: box 32 columns x 128 rows, 64-byte swizzle)
// ctx : [T, hidden] bf16
// grid = min(#SMs, units), block = THREADS, dynamic smem = smem_bytes(longest sequence); only sequences with
// len_lo < length <= len_hi are processed
template <bool TRACE = false>
__global__ void __launch_bounds__(THREADS, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ ctx,
                     const int* __restrict__ cu_seqlens, int n_seq, int heads, int max_tiles, int n_stages, int split,
                     float scale_log2, int len_lo, int len_hi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int hidden = heads * HEAD_DIM;
  const size_t stage_bytes = (size_t)3 * max_tiles * QKV_TILE_BYTES;  // [Q tiles | K tiles | V tiles]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)n_stages * stage_bytes);
  uint64_t* kv_full = bars;                    // [MAX_STAGES] TMA -> MMA: the unit has landed
  uint64_t* kv_empty = bars + MAX_STAGES;      // [MAX_STAGES] MMA -> TMA: one arrival per issuer (after its last P V in the unit, or on passing it)
  uint64_t* s_full = bars + 2 * MAX_STAGES;    // [GROUPS] MMA -> softmax: phase n = the scores of the group's block n are complete
                                               //          (and so is every MMA the group's issuer had issued before them)
  uint64_t* s_free = s_full + GROUPS;          // [GROUPS] softmax -> MMA: phase n = the scores of block n are in registers
  uint64_t* p_full = s_free + GROUPS;          // [GROUPS] softmax -> MMA: phase n = P(n) is in tensor memory (and O is rescaled / read out)
  uint64_t* o_full = p_full + GROUPS;          // [GROUPS] MMA -> softmax: phase n = P(n) V has retired (waited for at the end of the stream only)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + GROUPS);
  int* published = reinterpret_cast<int*>(tmem_ptr_smem + 1);        // descriptors written so far
  UnitDesc* desc = reinterpret_cast<UnitDesc*>(tmem_ptr_smem + 2);   // [DESC_RING]
  static_assert((2 * MAX_STAGES + 4 * GROUPS) * 8 + 8 + DESC_RING * sizeof(UnitDesc) <= BAR_BYTES, "barrier region too small");
  auto no_skip = [](int) {};

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_units = n_seq * heads * split;

  if (warp == TMA_WARP) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmap_qkv);
      for (int s = 0; s < MAX_STAGES; ++s) {
        tc::mbar_init(&kv_full[s], 1);
        tc::mbar_init(&kv_empty[s], GROUPS);
      }
      for (int g = 0; g < GROUPS; ++g) {
        tc::mbar_init(&s_full[g], 1);
        tc::mbar_init(&s_free[g], 4);
        tc::mbar_init(&p_full[g], 4);
        tc::mbar_init(&o_full[g], 1);
      }
      *published = 0;
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
  tc::pdl_launch_dependents();
  tc::pdl_wait();   // set-up overlapped the previous kernel's tail

  if (warp >= SOFTMAX_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(OTHER_REGS));
   if (warp == TMA_WARP) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int item = unit / split, part = unit - item * split;
        const int seq = item / heads, head = item - seq * heads;
        const int tok0 = __ldg(cu_seqlens + seq);
        const int S = __ldg(cu_seqlens + seq + 1) - tok0;
        const int n_tiles = (S + TILE - 1) / TILE;
        if (part >= n_tiles) continue;   // the unit has no jobs
        if (S <= len_lo || S > len_hi) continue;   // another kernel's length class (mixed batches)
        const int n_q = (n_tiles - part + split - 1) / split;
        tc::mbar_wait(&kv_empty[stage], phase ^ 1);
        UnitDesc& d = desc[it % DESC_RING];
        d.tok0 = tok0; d.S = S; d.head = head; d.part = part;
        st_release_shared(published, ++it);
        uint8_t* base = smem + (size_t)stage * stage_bytes;
        tc::mbar_arrive_expect_tx(&kv_full[stage], (uint32_t)((n_q + 2 * n_tiles) * QKV_TILE_BYTES));
        for (int t = 0; t < n_tiles; ++t) {
          const int row = tok0 + t * TILE;
          tc::tma_load_2d(&tmap_qkv, &kv_full[stage], base + (size_t)(max_tiles + t) * QKV_TILE_BYTES, hidden + head * HEAD_DIM, row);
          if (t >= part && (t - part) % split == 0)
            tc::tma_load_2d(&tmap_qkv, &kv_full[stage], base + (size_t)t * QKV_TILE_BYTES, head * HEAD_DIM, row);
        }
        for (int t = 0; t < n_tiles; ++t)
          tc::tma_load_2d(&tmap_qkv, &kv_full[stage], base + (size_t)(2 * max_tiles + t) * QKV_TILE_BYTES, 2 * hidden + head * HEAD_DIM, tok0 + t * TILE);
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
      // end of stream (the wait keeps the descriptor ring's distance argument intact)
      tc::mbar_wait(&kv_empty[stage], phase ^ 1);
      desc[it % DESC_RING].S = 0;
      st_release_shared(published, ++it);
    }
    __syncwarp();
   } else if (warp == MMA_WARP || warp == MMA_WARP + 1) {
    // ===================== MMA issuer of group g:  S(0) | { S(n+1), P(n) V } ... =====================
    // Per group the events alternate strictly -- s_free(n) (the scores of block n are in registers) precedes p_full(n) --
    // so blocking waits in that order never stall the other group, which has its own issuer.  INSIDE a job S(n+1) is
    // always issued before P(n) V and after P(n-1) V: the softmax threads rely on that order (s_full(n+1) tells them that
    // P(n-1) V is done).  Across jobs the next job's first scores go out early only if its unit has landed already -- with
    // a full ring it may be waiting for this very P V -- and the softmax threads wait for P(n-1) V itself.
    const int g = warp - MMA_WARP;
    constexpr uint32_t idesc_s = idesc_bf16(TILE, TILE, false);
    constexpr uint32_t idesc_o = idesc_bf16(TILE, HEAD_DIM, true);
    const uint32_t smem_base = tc::smem_u32(smem);
    const uint32_t tmem_g = tmem_base + (uint32_t)(g * GROUP_COLS);
    Cursor qk, pv;
    int qk_b = 0, pv_b = 0;
    uint32_t n_qk = 0, n_pv = 0, pv_job = 0;
    bool qk_ready = false;   // qk points at a block whose scores have not been issued yet
    auto landed = [&](int it) { return __all_sync(0xffffffffu, mbar_test(&kv_full[it % n_stages], (uint32_t)(it / n_stages) & 1)); };
    // a unit passed without a job of this group: its ring stage does not wait for this issuer
    auto skip_unit = [&](int it) { if (lane == 0) tc::mbar_arrive(&kv_empty[it % n_stages]); };
    auto issue_scores = [&]() {
      const int stage = qk.it % n_stages;
      tc::mbar_wait(&kv_full[stage], (uint32_t)(qk.it / n_stages) & 1);
      tc::mbar_wait(&s_free[g], (n_qk & 1) ^ 1);
      long long* tr = nullptr;
      if (TRACE && blockIdx.x == 0 && lane == 0 && n_qk < (uint32_t)TRACE_BLOCKS && g_trace) tr = g_trace + SOFTMAX_WARPS * TRACE_BLOCKS * 8 + ((size_t)g * TRACE_BLOCKS + n_qk) * 4;
      if (TRACE && tr) tr[2] = clock64();
      tc::tc_fence_after();
      const uint32_t base = smem_base + (uint32_t)stage * (uint32_t)stage_bytes;
      const uint64_t q_desc = desc_k_sw64(base + (uint32_t)(qk.qt * QKV_TILE_BYTES));
      const uint64_t k_desc = desc_k_sw64(base + (uint32_t)((max_tiles + qk_b) * QKV_TILE_BYTES));
      __syncwarp();
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < HEAD_DIM / 16; ++k)
          tc::umma_bf16(tmem_g, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
        tc::umma_commit(&s_full[g]);
      }
      __syncwarp();
      if (TRACE && tr) tr[3] = clock64();
      ++n_qk;
      if (++qk_b == qk.n_tiles) { qk_b = 0; qk_ready = false; }   // the next job is looked up when its turn comes
    };
    auto issue_pv = [&]() {
      const int stage = pv.it % n_stages;
      tc::mbar_wait(&p_full[g], n_pv & 1);
      long long* tr = nullptr;
      if (TRACE && blockIdx.x == 0 && lane == 0 && n_pv < (uint32_t)TRACE_BLOCKS && g_trace) tr = g_trace + SOFTMAX_WARPS * TRACE_BLOCKS * 8 + ((size_t)g * TRACE_BLOCKS + n_pv) * 4;
      if (TRACE && tr) tr[0] = clock64();
      tc::tc_fence_after();
      const uint32_t base = smem_base + (uint32_t)stage * (uint32_t)stage_bytes;
      // +16 keys: 8 columns of P (two bf16 per column), 1024 bytes down V's rows
      const uint64_t b_desc = desc_mn_sw64(base + (uint32_t)((2 * max_tiles + pv_b) * QKV_TILE_BYTES));
      const uint32_t o_tmem = tmem_g + (uint32_t)(O_COL + (pv_job & 1) * HEAD_DIM);
      const bool last_in_job = pv_b + 1 == pv.n_tiles;
      const bool release = last_in_job && pv.last_of_group_in_unit(split);
      // Keys beyond the sequence: their weights are exact zeros, but the V rows behind them belong to other sequences or
      // to the never-written tail of the buffer and may hold anything, NaN included (0 x NaN = NaN).  Only the 16-key steps
      // that contain a key of the sequence are issued, and the rows of the last step beyond the sequence are zeroed in
      // shared memory first (both groups' issuers may do that for the same tile: same zeros).
      const int valid = pv.S - pv_b * TILE;
      const int n_steps = valid >= TILE ? TILE / 16 : (valid + 15) >> 4;
      if (valid < n_steps * 16) {
        const uint32_t v_tile = base + (uint32_t)((2 * max_tiles + pv_b) * QKV_TILE_BYTES);
        const int chunks = (n_steps * 16 - valid) * 4;   // 16-byte chunks (the 64-byte swizzle permutes chunks inside a row only)
        for (int c = lane; c < chunks; c += 32)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(v_tile + (uint32_t)(valid * 64 + c * 16)), "r"(0u) : "memory");
        tc::fence_proxy_async();
      }
      __syncwarp();
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < TILE / 16; ++kk)
          if (kk < n_steps)
            umma_bf16_ts(o_tmem, tmem_g + (uint32_t)(P_COL + kk * 8), b_desc + (uint64_t)(kk * (1024 >> 4)), idesc_o, (pv_b | kk) != 0 ? 1u : 0u);
        tc::umma_commit(&o_full[g]);
        if (release) tc::umma_commit(&kv_empty[stage]);   // every MMA of this group that reads the unit's stage has been issued
      }
      __syncwarp();
      if (TRACE && tr) tr[1] = clock64();
      ++n_pv;
      if (last_in_job) {
        pv_b = 0;
        ++pv_job;
        // Blocking is safe here: everything this group owes the current unit has been issued, so the unit's stage
        // (which the next descriptor may be waiting for) is released by the other group's issuer alone.
        pv.template seek<true, true>(g, split, published, desc, skip_unit);
      } else {
        ++pv_b;
      }
    };
    pv.template seek<true, true>(g, split, published, desc, skip_unit);   // (units passed on the way are released at once)
    qk = pv;
    qk_ready = !qk.done;
    if (qk_ready) issue_scores();
    while (!pv.done) {
      bool scored = false;
      if (qk_ready) {
        // inside a job: the next block's scores first, always (same unit: landed)
        issue_scores();
        scored = true;
      } else if (!qk.done) {
        // the next job's first scores: early only if its descriptor is published and its unit has landed
        qk_ready = qk.template seek<true, false>(g, split, published, desc, no_skip) && !qk.done;
        if (qk_ready && landed(qk.it)) {
          issue_scores();
          scored = true;
        }
      }
      issue_pv();
      if (!scored && !qk.done) {
        if (!qk_ready) { qk.template seek<true, true>(g, split, published, desc, no_skip); qk_ready = !qk.done; }
        if (qk_ready) issue_scores();
      }
    }
    __syncwarp();
   }
  } else {
    // ===================== softmax groups: thread = one query row =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SOFTMAX_REGS));
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_tmem = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * GROUP_COLS);
    const uint32_t s_tmem = lane_tmem;
    const uint32_t p_tmem = lane_tmem + P_COL;
    const float lazy_raw = LAZY_LOG2 / scale_log2;
    const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
    uint32_t k = 0;      // key blocks this group has processed (-> barrier phases)
    uint32_t jobs = 0;   // jobs this group has finished (-> O accumulator parity)
    // a finished job whose output is still in its accumulator
    bool pend = false;
    float pend_l = 1.f;
    int pend_row = 0, pend_tok0 = 0, pend_head = 0, pend_acc = 0;
    bool pend_valid = false;
    auto flush = [&]() {   // O / l -> ctx of the pending job; its last P V has retired (the caller has seen to that)
      uint32_t o[HEAD_DIM];
      tc::tmem_ld32(lane_tmem + (uint32_t)(O_COL + pend_acc * HEAD_DIM), o);
      tc::tmem_ld_wait();
      if (pend_valid) {
        const float inv = 1.f / pend_l;
        uint4* dst = reinterpret_cast<uint4*>(ctx + (size_t)(pend_tok0 + pend_row) * hidden + pend_head * HEAD_DIM);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(pack2(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                              pack2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                              pack2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                              pack2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv));
      }
      pend = false;
    };
    Cursor w;
    while (w.template seek<false, true>(g, split, published, desc, no_skip), !w.done) {
      float m = -INFINITY, l = 0.f;
      for (int b = 0; b < w.n_tiles; ++b, ++k) {
        const int valid = w.S - b * TILE;            // keys of this block inside the sequence (>= 1)
        uint32_t r[TILE];
        long long* tr = nullptr;
        if (TRACE && blockIdx.x == 0 && lane == 0 && k < (uint32_t)TRACE_BLOCKS && g_trace) tr = g_trace + ((size_t)warp * TRACE_BLOCKS + k) * 8;
        if (TRACE && tr) tr[0] = clock64();
        if (b == 0) {   // (the scores of a job's later blocks were waited for one block ahead)
          tc::mbar_wait(&s_full[g], k & 1);
          tc::tc_fence_after();
        }
        if (TRACE && tr) tr[1] = clock64();
        const bool partial = valid < TILE;   // (the same for every row of the tile)
        load_scores(s_tmem, r);
        if (TRACE && tr) tr[2] = clock64();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&s_free[g]);   // the next block's scores may overwrite the buffer
        if (partial) {
#pragma unroll
          for (int i = 0; i < TILE; ++i)
            if (i >= valid) r[i] = 0xff800000u;       // keys beyond the sequence
        }
        const float mb = row_max(r);
        if (TRACE && tr) tr[3] = clock64() + (long long)(mb == 12345.f);   // (depends on the maximum: stamps after it)
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
        uint32_t pk[TILE / 2];
        // (a partial block skips the exponentials of the 32-key chunks that hold no key of the sequence)
        const float lsum = exp_scores(r, valid, scale2, noff2, pk);
        if (TRACE && tr) tr[4] = clock64() + (long long)(pk[TILE / 2 - 1] == 0x12345678u);   // (after the last exponential)
        // One wait before P is stored.  Inside a job: the next block's scores S(k+1) -- issued after P(k-1) V by the same
        // thread, so P(k-1) V has retired as well.  At a job's last block: P(k-1) V itself.  Either way the P columns are
        // free and O is complete up to block k-1.
        if (b + 1 < w.n_tiles) tc::mbar_wait(&s_full[g], (k + 1) & 1);
        else if (k > 0) tc::mbar_wait(&o_full[g], (k - 1) & 1);
        tc::tc_fence_after();
        if (TRACE && tr) tr[5] = clock64();
        if (pend) flush();   // the previous job's output (its accumulator is the other one)
        if (b > 0 && __any_sync(0xffffffffu, raise)) {
          // rare: a raised reference rescales the running O (complete: P(k-1) V has retired) and l
          const uint32_t o_tmem = lane_tmem + (uint32_t)(O_COL + (jobs & 1) * HEAD_DIM);
          uint32_t o[HEAD_DIM];
          tc::tmem_ld32(o_tmem, o);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < HEAD_DIM; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
          tc::tmem_st32(o_tmem, o);
          l *= corr;
        }
        l += lsum;
        tc::tmem_st32(p_tmem, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
        tc::tmem_st32(p_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
        tc::tmem_st_wait();
        if (TRACE && tr) tr[6] = clock64();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&p_full[g]);
        if (TRACE && tr) tr[7] = clock64();
      }
      // the job's output stays in its accumulator until the next block's wait (or the end of the stream)
      pend = true;
      pend_l = l;
      pend_row = w.qt * TILE + row;
      pend_valid = pend_row < w.S;
      pend_tok0 = w.tok0;
      pend_head = w.head;
      pend_acc = (int)(jobs & 1);
      ++jobs;
    }
    if (pend) {
      tc::mbar_wait(&o_full[g], (k - 1) & 1);
      tc::tc_fence_after();
      flush();
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace attn3
}  // namespace drag
