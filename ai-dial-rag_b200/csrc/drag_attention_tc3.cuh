// Multi-head self-attention on tcgen05 for the packed (padding-free) batch, head_dim = 32
// (SURVEY.md 8a row a5:  softmax(Q K^T / sqrt(32) + key mask) V ).  Third design.
//
// At head_dim 32 the op is bound by the exponentials (one per 128 flop; MUFU: 16 per clock per SM, measured),
// so the kernel is arranged around ONE rule: a softmax warp never waits for anything another softmax warp of
// the other group could be computing exponentials behind.
//
// Persistent kernel, one CTA per SM, 11 working warps (12 launched: register budgets are set per warpgroup):
//   warps 0-3   softmax group 0        thread = one query row of the group's current 128-query job
//   warps 4-7   softmax group 1        (jobs alternate between the groups: the two run out of phase)
//   warp  8     TMA producer: Q, K and V tiles (128 tokens x one head, 64-byte swizzle) of the next units
//               straight out of the packed [T, 3*hidden] QKV buffer into a 1..4-stage ring
//   warps 9,10  MMA issuers, one per softmax group, so the groups never wait for each other's tiles whatever the
//               sequence lengths are.  The whole warp walks the job sequence (warp-uniform values: descriptors live
//               in uniform registers) and one elected lane issues -- a P V step is only 16 clocks of tensor work, so
//               the cost of ISSUING a tcgen05.mma matters here
// The producer publishes a small descriptor of every unit next to its ring stage: the issuer and the softmax
// warps walk the job sequence on shared memory only (a walker on global loads stalled the issuer for
// thousands of cycles per job).
// A unit = one (sequence, head) item, or 1/split of its query tiles when few items are in flight (the query
// path); a job = one 128-query tile of a unit against all keys of the sequence, in 128-key blocks.
// Per job and key block, in group g:
//   S[128 x 128] = Q K^T        tcgen05.mma M=128 N=128 K=16 x2 -> TMEM columns [128 g, +128)
//   softmax thread              ONE tcgen05.ld pass: the row's 128 scores go to registers and the S buffer is
//                               handed back at once (the next block's scores are computed behind this block's
//                               exponentials); row maximum, lazily raised reference m (only when a score
//                               exceeds it by 2^8: the softmax is invariant to the reference), p = 2^(s c - m c)
//                               (FFMA2 + MUFU.EX2), bf16 pairs by TRUNCATION (one PRMT per pair: the F2FP
//                               conversion runs at a third of the ALU rate, measured) into the group's P buffer
//                               (K-major, 128-byte swizzle: the A operand of P V)
//   O[128 x 32] (+)= P V        tcgen05.mma M=128 N=32 K=16 x8 into TMEM columns [256 + 48 g, +32); V is
//   L[128 x 16] (+)= P 1        consumed as stored ([key][32]: an MN-major B operand); the row sums come out of the
//                               tensor core over exactly the truncated weights of the numerator (the truncation
//                               bias cancels in O / L, and the softmax loop carries no additions).  O and L stay
//                               in TMEM across the key blocks of a job; a raised reference rescales them (rare).
//   end of job                  O / L -> bf16 -> ctx (rows inside the sequence only)
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
constexpr int THREADS = 384;                            // 3 warpgroups: softmax 0, softmax 1, {TMA, MMA, two idle warps}
constexpr int SOFTMAX_REGS = 200;                       // setmaxnreg: 8 x 32 x 200 + 4 x 32 x 104 = 64 512 = 384 x 168
constexpr int OTHER_REGS = 104;
constexpr int TMEM_COLS = 512;
constexpr int O_COL = GROUPS * TILE;                    // group g: S at [128 g, +128), O at [256 + 48 g, +32), L at [256 + 48 g + 32, +16)
constexpr int OL_COLS = HEAD_DIM + 16;                  // O and the row sums L = P . 1 (a 16-column MMA against a tile of ones)
constexpr int ONES_BYTES = 1024;                        // [16][32] bf16 ones (all equal: layout and swizzle do not matter)
constexpr int P_ATOM_BYTES = TILE * 64 * 2;             // 16 KB: [128 rows][64 keys] bf16, 128-byte swizzle
constexpr int P_BYTES = 2 * P_ATOM_BYTES;               // one 128-key block of P
constexpr int MAX_STAGES = 4;
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int BAR_BYTES = 512;
constexpr float LAZY_LOG2 = 8.f;                        // numerators stay <= 2^8

__host__ __device__ inline int unit_stages(int max_len, int p_bufs) {
  const int tiles = (max_len + TILE - 1) / TILE;
  const int fit = (SMEM_BUDGET - GROUPS * p_bufs * P_BYTES - ONES_BYTES - 1024 - BAR_BYTES) / (3 * tiles * QKV_TILE_BYTES);
  return fit < 1 ? 1 : (fit < MAX_STAGES ? fit : MAX_STAGES);
}
__host__ __device__ inline size_t smem_bytes(int max_len, int p_bufs) {
  const int tiles = (max_len + TILE - 1) / TILE;
  return (size_t)unit_stages(max_len, p_bufs) * 3 * tiles * QKV_TILE_BYTES + (size_t)GROUPS * p_bufs * P_BYTES + ONES_BYTES + 1024 /*alignment*/ + BAR_BYTES;
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
// 2^x for a pair on the FMA / ALU pipes instead of the MUFU pipe (16 ex2 per clock per SM is what bounds this kernel):
// x = j + f with j = round(x) (magic-number add), 2^f for f in [-0.5, 0.5] by a degree-3 polynomial (relative error 1.0e-4,
// a twentieth of the bf16 rounding the weight gets next), then j goes straight into the exponent field.  Valid for
// x in [-126, 127): smaller arguments are clamped (2^-126 is as good as 0 for a softmax weight).
__device__ __forceinline__ void exp2_poly_pair(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  unpack_f32x2(x2, x0, x1);
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t xc = pack_f32x2(x0, x1);
  const uint64_t t2 = add_f32x2(xc, pack_f32x2(12582912.f, 12582912.f));          // 1.5 * 2^23: j in the low mantissa bits
  const uint64_t j2 = add_f32x2(t2, pack_f32x2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma_f32x2(j2, pack_f32x2(-1.f, -1.f), xc);
  uint64_t q2 = fma_f32x2(f2, pack_f32x2(0.05500893f, 0.05500893f), pack_f32x2(0.24221095f, 0.24221095f));
  q2 = fma_f32x2(q2, f2, pack_f32x2(0.6932829f, 0.6932829f));
  q2 = fma_f32x2(q2, f2, pack_f32x2(1.f, 1.f));
  float t0, t1, q0, q1;
  unpack_f32x2(t2, t0, t1);
  unpack_f32x2(q2, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
      : "memory");
}
// two non-negative fp32 -> bf16 pair by truncation (high halves): one PRMT
__device__ __forceinline__ uint32_t trunc_bf16x2(float lo, float hi) { return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632); }

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

// The CTA's jobs in a fixed order: its units in ring order, per unit the query tiles part, part + split, ...; job ordinal
// n (CTA-wide) belongs to softmax group n & 1.  A cursor points at one job of one group.
constexpr int DESC_RING = 2 * MAX_STAGES;

__device__ __forceinline__ int ld_acquire_shared(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(v) : "r"(tc::smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(p)), "r"(v) : "memory");
}

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

// qkv : [T, 3*hidden] bf16 (tensor map: box 32 columns x 128 rows, 64-byte swizzle)
// ctx : [T, hidden] bf16
// grid = min(#SMs, units), block = THREADS, dynamic smem = smem_bytes(longest sequence, P_BUFS)
// Debug timeline (TRACE instantiation only, scripts/attn_trace.py): CTA 0 records clock64() at the phase boundaries of its
// first TRACE_BLOCKS key blocks -- per softmax warp 8 stamps per block, per issuer 4 stamps per block.
constexpr int TRACE_BLOCKS = 96;
constexpr int TRACE_WORDS = SOFTMAX_WARPS * TRACE_BLOCKS * 8 + GROUPS * TRACE_BLOCKS * 4;
__device__ long long* g_trace = nullptr;

// POLY: every POLY-th pair of weights is computed by exp2_poly_pair instead of MUFU.EX2 (0 = none, 4 = 25 %, 2 = 50 %)
template <int P_BUFS, int POLY, bool TRACE = false>
__global__ void __launch_bounds__(THREADS, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ ctx,
                     const int* __restrict__ cu_seqlens, int n_seq, int heads, int max_tiles, int n_stages, int split,
                     float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int hidden = heads * HEAD_DIM;
  const size_t stage_bytes = (size_t)3 * max_tiles * QKV_TILE_BYTES;  // [Q tiles | K tiles | V tiles]
  uint8_t* p_smem = smem + (size_t)n_stages * stage_bytes;            // multiple of 8 KB: 1024-aligned
  uint8_t* ones_smem = p_smem + (size_t)GROUPS * P_BUFS * P_BYTES;   // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones_smem + ONES_BYTES);
  uint64_t* kv_full = bars;                    // [MAX_STAGES] TMA -> everybody: the unit (and its descriptor) has landed
  uint64_t* kv_empty = bars + MAX_STAGES;      // [MAX_STAGES] MMA -> TMA (every MMA that reads the unit has retired)
  uint64_t* s_full = bars + 2 * MAX_STAGES;    // [GROUPS] MMA -> softmax: the scores of the group's next block are complete
  uint64_t* s_free = s_full + GROUPS;          // [GROUPS] softmax -> MMA: the scores are in registers
  uint64_t* p_full = s_free + GROUPS;          // [GROUPS][2] softmax -> MMA: P of the group's block n is in shared memory (and O is
                                               //             rescaled): barrier n % P_BUFS, phase n / P_BUFS
  uint64_t* o_full = p_full + 2 * GROUPS;      // [GROUPS][2] MMA -> softmax: P V of block n has retired (P buffer n % P_BUFS reusable,
                                               //             O complete up to block n): barrier n % P_BUFS, phase n / P_BUFS
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 2 * GROUPS);
  int* published = reinterpret_cast<int*>(tmem_ptr_smem + 1);        // descriptors written so far
  UnitDesc* desc = reinterpret_cast<UnitDesc*>(tmem_ptr_smem + 2);   // [DESC_RING]
  static_assert((2 * MAX_STAGES + 6 * GROUPS) * 8 + 8 + DESC_RING * sizeof(UnitDesc) <= BAR_BYTES, "barrier region too small");
  auto no_skip = [](int) {};

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_units = n_seq * heads * split;

  if (warp == TMA_WARP) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmap_qkv);
      for (int s = 0; s < MAX_STAGES; ++s) {
        tc::mbar_init(&kv_full[s], 1);
        tc::mbar_init(&kv_empty[s], GROUPS);   // one arrival per issuer: after its last P V in the unit, or on passing a unit without a job of its group
      }
      for (int g = 0; g < GROUPS; ++g) {
        tc::mbar_init(&s_full[g], 1);
        tc::mbar_init(&s_free[g], 4);
        for (int b = 0; b < 2; ++b) {
          tc::mbar_init(&p_full[2 * g + b], 4);
          tc::mbar_init(&o_full[2 * g + b], 1);
        }
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
  for (int i = threadIdx.x; i < ONES_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(ones_smem)[i] = 0x3f803f80u;   // bf16 1.0 pairs
  tc::fence_proxy_async();   // generic-proxy stores -> visible to the tensor core
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

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
    // so blocking waits in that order never stall the other group, which has its own issuer.  The next block's scores
    // are issued BEFORE P(n) V unless their unit has not landed yet (its ring stage may be waiting for this very P V).
    const int g = warp - MMA_WARP;
    constexpr uint32_t idesc_s = idesc_bf16(TILE, TILE, false);
    constexpr uint32_t idesc_o = idesc_bf16(TILE, HEAD_DIM, true);
    constexpr uint32_t idesc_l = idesc_bf16(TILE, 16, false);
    const uint64_t ones_desc = desc_k_sw64(tc::smem_u32(ones_smem));   // 16 "columns" x 16 keys of ones, K-major: rows of 64 bytes
    const uint32_t smem_base = tc::smem_u32(smem);
    const uint32_t p_base0 = tc::smem_u32(p_smem) + (uint32_t)(g * P_BUFS * P_BYTES);
    // (all lanes run the control flow; one elected lane issues)
    Cursor qk, pv;
    int qk_b = 0, pv_b = 0;
    uint32_t n_qk = 0, n_pv = 0;
    bool qk_ready = false;   // qk points at a block whose scores have not been issued yet
    // a unit passed without a job of this group: its ring stage does not wait for this issuer
    auto skip_unit = [&](int it) { if (lane == 0) tc::mbar_arrive(&kv_empty[it % n_stages]); };
    auto landed = [&](int it) { return __all_sync(0xffffffffu, mbar_test(&kv_full[it % n_stages], (uint32_t)(it / n_stages) & 1)); };
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
          tc::umma_bf16(tmem_base + g * TILE, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
        tc::umma_commit(&s_full[g]);
      }
      __syncwarp();
      if (TRACE && tr) tr[3] = clock64();
      ++n_qk;
      if (++qk_b == qk.n_tiles) { qk_b = 0; qk_ready = false; }   // the next job is looked up when its turn comes
    };
    auto issue_pv = [&]() {
      const int stage = pv.it % n_stages;
      const int buf = (int)(n_pv % P_BUFS);
      tc::mbar_wait(&p_full[2 * g + buf], (n_pv / P_BUFS) & 1);
      long long* tr = nullptr;
      if (TRACE && blockIdx.x == 0 && lane == 0 && n_pv < (uint32_t)TRACE_BLOCKS && g_trace) tr = g_trace + SOFTMAX_WARPS * TRACE_BLOCKS * 8 + ((size_t)g * TRACE_BLOCKS + n_pv) * 4;
      if (TRACE && tr) tr[0] = clock64();
      tc::tc_fence_after();
      const uint32_t base = smem_base + (uint32_t)stage * (uint32_t)stage_bytes;
      // +16 keys: 32 bytes inside P's 128-byte swizzle row (4 steps per 64-key atom), 1024 bytes down V's rows
      const uint64_t a_desc = tc::umma_desc_sw128(p_base0 + (uint32_t)(buf * P_BYTES));
      const uint64_t b_desc = desc_mn_sw64(base + (uint32_t)((2 * max_tiles + pv_b) * QKV_TILE_BYTES));
      const bool last_in_job = pv_b + 1 == pv.n_tiles;
      const bool release = last_in_job && pv.last_of_group_in_unit(split);
      __syncwarp();
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < TILE / 16; ++kk)
        {
          const uint64_t a_kk = a_desc + (uint64_t)((kk >> 2) * (P_ATOM_BYTES >> 4) + (kk & 3) * 2);
          tc::umma_bf16(tmem_base + O_COL + g * OL_COLS, a_kk, b_desc + (uint64_t)(kk * (1024 >> 4)), idesc_o, (pv_b | kk) != 0 ? 1u : 0u);
          tc::umma_bf16(tmem_base + O_COL + g * OL_COLS + HEAD_DIM, a_kk, ones_desc, idesc_l, (pv_b | kk) != 0 ? 1u : 0u);
        }
        tc::umma_commit(&o_full[2 * g + buf]);
        if (release) tc::umma_commit(&kv_empty[stage]);   // every MMA of this group that reads the unit's stage has been issued
      }
      __syncwarp();
      if (TRACE && tr) tr[1] = clock64();
      ++n_pv;
      if (last_in_job) {
        pv_b = 0;
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
      // the scores of the next block go out BEFORE P(n) V -- unless their unit (or even its descriptor) is not there yet:
      // with a full ring it is waiting for this very P V
      if (!qk_ready && !qk.done) qk_ready = qk.template seek<true, false>(g, split, published, desc, no_skip) && !qk.done;
      bool scored = false;
      if (qk_ready && landed(qk.it)) {
        issue_scores();
        scored = true;
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
    const uint32_t lane_tmem = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t s_tmem = lane_tmem + (uint32_t)(g * TILE);
    const uint32_t o_tmem = lane_tmem + (uint32_t)(O_COL + g * OL_COLS);   // O, then L at +32
    const float lazy_raw = LAZY_LOG2 / scale_log2;
    const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
    (void)add_f32x2;
    uint32_t k = 0;   // key blocks this group has processed (-> barrier phases, P buffer)
    // waits until P V of the group's block n has retired (the MMAs retire in order: so have all earlier ones)
    auto wait_pv = [&](uint32_t n) { tc::mbar_wait(&o_full[2 * g + (int)(n % P_BUFS)], (n / P_BUFS) & 1); };
    Cursor w;
    while (w.template seek<false, true>(g, split, published, desc, no_skip), !w.done) {
      float m = -INFINITY;
      for (int b = 0; b < w.n_tiles; ++b, ++k) {
        const int valid = w.S - b * TILE;            // keys of this block inside the sequence (>= 1)
        uint32_t r[TILE];
        long long* tr = nullptr;
        if (TRACE && blockIdx.x == 0 && lane == 0 && k < (uint32_t)TRACE_BLOCKS && g_trace) tr = g_trace + ((size_t)warp * TRACE_BLOCKS + k) * 8;
        if (TRACE && tr) tr[0] = clock64();
        tc::mbar_wait(&s_full[g], k & 1);
        if (TRACE && tr) tr[1] = clock64();
        tc::tc_fence_after();
#pragma unroll
        for (int c = 0; c < TILE / 32; ++c) tc::tmem_ld32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
        tc::tmem_ld_wait();
        if (TRACE && tr) tr[2] = clock64();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&s_free[g]);   // the next block's scores may overwrite the buffer
        if (valid < TILE) {
#pragma unroll
          for (int i = 0; i < TILE; ++i)
            if (i >= valid) r[i] = 0xff800000u;       // keys beyond the sequence
        }
        // four independent chains of 3-input maxima
        float mx[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) mx[c] = max3(__uint_as_float(r[3 * c]), __uint_as_float(r[3 * c + 1]), __uint_as_float(r[3 * c + 2]));
#pragma unroll
        for (int i = 12; i + 7 < TILE; i += 8) {
#pragma unroll
          for (int c = 0; c < 4; ++c) mx[c] = max3(mx[c], __uint_as_float(r[i + 2 * c]), __uint_as_float(r[i + 2 * c + 1]));
        }
        // 12 + 8 * 14 = 124: four scores left
        mx[0] = max3(mx[0], __uint_as_float(r[TILE - 4]), __uint_as_float(r[TILE - 3]));
        mx[1] = max3(mx[1], __uint_as_float(r[TILE - 2]), __uint_as_float(r[TILE - 1]));
        const float mb = max3(fmaxf(mx[0], mx[1]), mx[2], mx[3]);
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
        // p = 2^(s c - m c), truncated to bf16 pairs (the row sums are taken by the tensor core over the same values)
        const float off = m * scale_log2;
        const uint64_t noff2 = pack_f32x2(-off, -off);
        uint32_t pk[TILE / 2];
#pragma unroll
        for (int i = 0; i < TILE; i += 4) {
          float p0, p1, p2, p3;
          const uint64_t xa = fma_f32x2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), scale2, noff2);
          const uint64_t xb = fma_f32x2(pack_f32x2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), scale2, noff2);
          if (POLY > 0 && ((i >> 1) % (POLY > 0 ? POLY : 1)) == 0) {
            exp2_poly_pair(xa, p0, p1);
          } else {
            unpack_f32x2(xa, p0, p1);
            p0 = ex2(p0); p1 = ex2(p1);
          }
          if (POLY > 0 && (((i >> 1) + 1) % (POLY > 0 ? POLY : 1)) == 0) {
            exp2_poly_pair(xb, p2, p3);
          } else {
            unpack_f32x2(xb, p2, p3);
            p2 = ex2(p2); p3 = ex2(p3);
          }
          pk[i >> 1] = trunc_bf16x2(p0, p1);
          pk[(i >> 1) + 1] = trunc_bf16x2(p2, p3);
        }
        if (TRACE && tr) tr[4] = clock64() + (long long)(pk[TILE / 2 - 1] == 0x12345678u);   // (after the last exponential)
        // the P V of P_BUFS blocks ago has retired: its P buffer is reusable
        if (k >= (uint32_t)P_BUFS) wait_pv(k - P_BUFS);
        if (TRACE && tr) tr[5] = clock64();
        if (b > 0 && __any_sync(0xffffffffu, raise)) {
          // rare: a raised reference rescales the running O and L (complete once the previous P V has retired)
          if (P_BUFS > 1) wait_pv(k - 1);
          tc::tc_fence_after();
          uint32_t o[HEAD_DIM], ls[16];
          tc::tmem_ld32(o_tmem, o);
          tmem_ld16(o_tmem + HEAD_DIM, ls);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < HEAD_DIM; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
#pragma unroll
          for (int i = 0; i < 16; ++i) ls[i] = __float_as_uint(__uint_as_float(ls[i]) * corr);
          tc::tmem_st32(o_tmem, o);
          tmem_st16(o_tmem + HEAD_DIM, ls);
          tc::tmem_st_wait();
        }
        // 16-byte chunk j of row r lives at chunk (j ^ (r & 7)) of the row's 128 bytes
        const uint32_t p_row = tc::smem_u32(p_smem + (size_t)(g * P_BUFS + (int)(k % P_BUFS)) * P_BYTES) + (uint32_t)row * 128u;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t chunk = (uint32_t)((j & 7) ^ (row & 7));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + (uint32_t)((j >> 3) * P_ATOM_BYTES) + chunk * 16u),
                       "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
        }
        if (TRACE && tr) tr[6] = clock64();
        tc::fence_proxy_async();   // P (generic-proxy stores) -> visible to the tensor core
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&p_full[2 * g + (int)(k % P_BUFS)]);
        if (TRACE && tr) tr[7] = clock64();
      }
      // ---------------- end of job: O / l -> ctx ----------------
      wait_pv(k - 1);
      tc::tc_fence_after();
      uint32_t o[HEAD_DIM];
      tc::tmem_ld32(o_tmem, o);
      const float l = __uint_as_float(tc::tmem_ld1(o_tmem + HEAD_DIM));   // every column of L holds the row sum
      tc::tmem_ld_wait();
      tc::tc_fence_before();   // ordered before this warp's next p_full arrival (the next job's first P V overwrites O and L)
      const int qrow = w.qt * TILE + row;
      if (qrow < w.S) {
        const float inv = 1.f / l;
        uint4* dst = reinterpret_cast<uint4*>(ctx + (size_t)(w.tok0 + qrow) * hidden + w.head * HEAD_DIM);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(pack2(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                              pack2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                              pack2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                              pack2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv));
      }
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
