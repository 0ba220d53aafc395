// Batched exact top-k: tensor-core candidate generation + exact float64 re-rank (sm_100a).
//
// The float64 scan of drag_topk.cu reads the matrix once per <= 4 queries; for a batch of
// hundreds of queries (BASELINE configs[2]/[3]) that is hundreds of passes.  This path answers the
// same question -- the reference's stable top-k under float64 scoring,
// embeddings_index.py:51-89 -- with ONE pass per group of query tiles:
//
//   1. score kernel (tcgen05): S~ = Q~ . D~^T with bf16 operands (D~ = bf16 copy of the rows,
//      Q~ = bf16(query)), fp32 accumulation in TMEM.  A CTA keeps one tile of 128 queries resident
//      in shared memory (A operand) and streams 256-row tiles of D~ through a TMA ring (B operand),
//      so TMEM lane = query, TMEM column = matrix row.  The epilogue turns a score into a
//      "bigger is better" fp32 key (IP: s; (sq)euclid: s - |d|^2/2; cosine: s / |d|) and appends
//      (key, row) to the query's candidate list iff key >= theta[q].
//   2. refine kernel: after every round of rows, theta[q] = (k-th best key so far) - 2*E[q] and the
//      candidate list is compacted.  E[q] is a certified bound on |key - exact key| (bf16 rounding
//      of both operands + accumulation), so every row of the exact top-k has key >= theta[q]
//      (proof in DESIGN.md: the k rows with the best approximate keys have exact keys >=
//      key_(k) - E, hence the exact k-th best is >= key_(k) - E and its approximate key >=
//      key_(k) - 2E).  Rounds grow geometrically (2k, 16k, 128k, ... rows) so the threshold is
//      already tight when most of the matrix streams by.
//   3. re-rank kernel: the surviving candidates (a few hundred per query) are re-scored with the
//      SAME float64 arithmetic as scan_vec_kernel and sorted on (float64 distance, row id).
//
// The result is therefore identical to the float64 scan's.  Anything the certificate cannot cover
// (candidate overflow on adversarial data, NaN keys, negative squared distances under sqrt) sets
// status[q] = 1 and the caller re-runs that query through drag_topk.
#pragma once

#include "drag_tc.cuh"

namespace drag {
namespace tcs {

constexpr int QT = 128;                      // queries per tile   (UMMA M)
constexpr int RT = 256;                      // matrix rows per tile (UMMA N)
constexpr int BK = 64;                       // bf16 elements per k-block (one 128-byte swizzle row)
constexpr int A_KB_BYTES = QT * BK * 2;      // 16 KB
constexpr int B_STAGE_BYTES = RT * BK * 2;   // 32 KB
constexpr int EPI_WARPS = 4;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int MAX_STAGES = 6;
constexpr int MAX_DIM = 512;
constexpr int MAX_BATCH_K = 256;
constexpr int RERANK_MAX = 2048;             // candidates the re-rank kernel sorts per query
constexpr int ROUND0_ROWS = 2048;            // first round collects every score of these rows
constexpr int ROUND_GROWTH = 8;
constexpr int POOL_CAP = 8192;               // new candidates per query and round (all regions together)
constexpr int SURV_CAP = RERANK_MAX;         // candidates kept per query between rounds
constexpr int MAX_SLOTS = 160;               // >= CTAs per query tile (<= number of SMs)

enum { MODE_IP = 0, MODE_L2 = 1, MODE_COS = 2 };

struct ScoreParams {
  int qt0;               // first query tile of this launch
  int n_qtiles;          // query tiles in this launch
  int ctas_per_qtile;
  int n_valid_q;         // queries beyond this index are padding
  long long row0, row1;  // rows of this round (row0 is a multiple of RT)
  int k_blocks;          // dim / 64
  int stages;
  const float* colvec;   // MODE_L2: |d|^2 (float32, numpy order)  MODE_COS: 1 / max(|d|, 1e-8)
  const float* theta;    // [Qpad]
  uint2* pool;           // [Qpad][POOL_CAP] (key bits, row): region `slot` of a query belongs to ONE CTA
  int* cnt;              // [Qpad][MAX_SLOTS] entries pushed into each region by this launch
  int2* meta;            // [Qpad] (regions used, region capacity) of this launch
  int region_cap;        // POOL_CAP / ctas_per_qtile
  int collect_all;       // round 0: every score is a candidate
};

__host__ __device__ inline size_t score_smem_bytes(int k_blocks, int stages) {
  return (size_t)k_blocks * A_KB_BYTES + (size_t)stages * B_STAGE_BYTES + 1024 /*alignment*/ + 512 /*barriers*/;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
score_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_m, ScoreParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_smem = smem;
  uint8_t* ring = smem + (size_t)p.k_blocks * A_KB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.stages * B_STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [MAX_STAGES]
  uint64_t* empty_bar = bars + MAX_STAGES;         // [MAX_STAGES]
  uint64_t* tmem_full_bar = bars + 2 * MAX_STAGES; // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint64_t* a_full_bar = tmem_empty_bar + 2;       // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(a_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = p.qt0 + (int)(blockIdx.x % p.n_qtiles);
  const int slot = blockIdx.x / p.n_qtiles;
  const int n_tiles = (int)((p.row1 - p.row0 + RT - 1) / RT);

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_q);
    tc::tma_prefetch_desc(&tmap_m);
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&tmem_full_bar[s], 1);
      tc::mbar_init(&tmem_empty_bar[s], EPI_WARPS);
    }
    tc::mbar_init(a_full_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_ptr_smem, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      // the query tile stays resident for the whole kernel
      tc::mbar_arrive_expect_tx(a_full_bar, (uint32_t)(p.k_blocks * A_KB_BYTES));
      for (int kb = 0; kb < p.k_blocks; ++kb)
        tc::tma_load_2d(&tmap_q, a_full_bar, a_smem + (size_t)kb * A_KB_BYTES, kb * BK, qt * QT);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = slot; t < n_tiles; t += p.ctas_per_qtile) {
        const int row = (int)(p.row0 + (long long)t * RT);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[stage], B_STAGE_BYTES);
          tc::tma_load_2d(&tmap_m, &full_bar[stage], ring + (size_t)stage * B_STAGE_BYTES, kb * BK, row);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(QT, RT);
      tc::mbar_wait(a_full_bar, 0);
      tc::tc_fence_after();
      const uint32_t a_base = tc::smem_u32(a_smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = slot; t < n_tiles; t += p.ctas_per_qtile) {
        tc::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * RT);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint64_t a_desc = tc::umma_desc_sw128(a_base + (uint32_t)(kb * A_KB_BYTES));
          const uint64_t b_desc = tc::umma_desc_sw128(tc::smem_u32(ring + (size_t)stage * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc::umma_bf16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          tc::umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: threshold filter =====================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) belong to this warp
    const int q = qt * QT + quarter * 32 + lane;
    const bool q_ok = q < p.n_valid_q;
    const bool all = p.collect_all != 0;
    float theta = INFINITY;
    if (q_ok) theta = all ? -INFINITY : __ldg(p.theta + q);
    // candidates go to a region only this thread writes: no atomics, stores never stall the warp
    uint2* region = p.pool + ((size_t)q * POOL_CAP + (size_t)slot * p.region_cap);
    int n_mine = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = slot; t < n_tiles; t += p.ctas_per_qtile) {
      const long long tile_row0 = p.row0 + (long long)t * RT;
      const long long left = p.row1 - tile_row0;
      const int n_valid = left < RT ? (int)left : RT;
      tc::mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * RT);
#pragma unroll 1
      for (int c0 = 0; c0 < RT; c0 += 32) {
        __syncwarp();  // the rare path below diverges; tcgen05.ld is warp-collective
        uint32_t r[32];
        tc::tmem_ld32(t_row + c0, r);
        float cv[32];
        if (MODE != MODE_IP) {
          const float* src = p.colvec + tile_row0 + c0;
          if (c0 + 32 <= n_valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
              cv[4 * i] = v.x; cv[4 * i + 1] = v.y; cv[4 * i + 2] = v.z; cv[4 * i + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) cv[i] = (c0 + i < n_valid) ? __ldg(src + i) : 0.f;
          }
        }
        tc::tmem_ld_wait();
        float key[32];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = __uint_as_float(r[i]);
          if (MODE == MODE_IP) key[i] = s;
          else if (MODE == MODE_L2) key[i] = fmaf(-0.5f, cv[i], s);
          else key[i] = s * cv[i];
          mx = fmaxf(mx, key[i]);
        }
        if ((all && q_ok) || mx >= theta) {
          // Rare once the threshold is tight (a handful of rows per query and round).  Branch-free:
          // a predicated 8-byte store per column, so a warp whose lanes hit in different columns
          // does not pay 32 divergent branches.
          const int room = p.region_cap;
          const uint32_t row_c0 = (uint32_t)(tile_row0 + c0);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool hit = (all || key[i] >= theta) && (c0 + i < n_valid);
            const uint32_t do_store = (hit && n_mine < room) ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %0, 0;\n\t"
                "@p st.global.v2.b32 [%1], {%2, %3};\n\t}"
                ::"r"(do_store), "l"(region + n_mine), "r"(__float_as_uint(key[i])), "r"(row_c0 + (uint32_t)i)
                : "memory");
            n_mine += hit ? 1 : 0;
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (q_ok) {
      p.cnt[(size_t)q * MAX_SLOTS + slot] = n_mine;  // > region_cap means the region overflowed
      if (slot == 0) p.meta[q] = make_int2(p.ctas_per_qtile, p.region_cap);
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------
// queries: float64 -> bf16 [Qpad, dim] (zero padded)
// ---------------------------------------------------------------------------------
__global__ void queries_to_bf16_kernel(const double* __restrict__ q, int n_queries, int q_pad, int dim,
                                       __nv_bfloat16* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)q_pad * dim;
  if (i >= n) return;
  const size_t row = i / dim;
  out[i] = row < (size_t)n_queries ? __float2bfloat16_rn((float)q[i]) : __float2bfloat16_rn(0.f);
}

__global__ void rows_to_bf16_kernel(const float4* __restrict__ in, size_t n4, uint2* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = __ldg(in + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    out[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// stats[0] = max |d|^2, stats[1] = min non-zero |d|^2, stats[2] = number of non-finite |d|^2;
// inv (optional) = 1 / max(|d|, 1e-8).  stats must be initialised to {0, +inf, 0}.
__global__ void row_norm_stats_kernel(const float* __restrict__ sq, long long n, float* __restrict__ inv,
                                      float* __restrict__ stats) {
  float mx = 0.f, mn = INFINITY, bad = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = sq[i];
    if (!(fabsf(v) <= 3.0e38f)) bad += 1.f;
    else {
      mx = fmaxf(mx, v);
      if (v > 0.f) mn = fminf(mn, v);
    }
    if (inv) inv[i] = 1.0f / fmaxf(sqrtf(v), 1e-8f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
    // all values are non-negative: the integer order of the bit patterns is the float order
    atomicMax(reinterpret_cast<int*>(stats), __float_as_int(mx));
    atomicMin(reinterpret_cast<int*>(stats + 1), __float_as_int(mn));
    if (bad > 0.f) atomicAdd(stats + 2, bad);
  }
}

// ---------------------------------------------------------------------------------
// refine: new threshold + compaction, one CTA per query
// ---------------------------------------------------------------------------------
struct RefineParams {
  const uint2* pool;      // new candidates of the round that just ran
  const int* cnt;
  const int2* meta;
  uint2* surv;            // [Qpad][SURV_CAP] candidates kept from earlier rounds
  int* surv_cnt;          // [Qpad]
  float* theta;
  int* overflow;
  const double* q_norm;   // max(|q|, 1e-8)
  double err_a, err_b;    // E[q] = err_a * |q| + err_b
  int k;
};

__device__ __forceinline__ uint32_t ordered_from_float_bits(uint32_t b) {
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0u;  // NaN: worst
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_ordered(uint32_t o) {
  const uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}

constexpr int REFINE_THREADS = 256;
constexpr int REFINE_MAX = POOL_CAP + SURV_CAP;   // entries a refine CTA can hold in shared memory

__host__ __device__ inline size_t refine_smem_bytes() { return (size_t)REFINE_MAX * 8; }

__global__ void __launch_bounds__(REFINE_THREADS) refine_kernel(RefineParams p) {
  extern __shared__ __align__(16) unsigned char smem_refine[];
  uint32_t* keys = reinterpret_cast<uint32_t*>(smem_refine);  // [REFINE_MAX] ordered keys
  uint32_t* rows = keys + REFINE_MAX;                           // [REFINE_MAX]
  __shared__ uint32_t hist[256];
  __shared__ int offs[MAX_SLOTS + 1];
  __shared__ uint32_t s_prefix, s_remaining, s_out;
  __shared__ int s_fail;
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  if (p.overflow[q]) return;
  const int2 meta = p.meta[q];
  const int n_slots = meta.x, region_cap = meta.y;
  const int n_old = p.surv_cnt[q];
  if (tid == 0) {
    int total = 0, fail = 0;
    for (int s = 0; s < n_slots; ++s) {
      const int c = p.cnt[(size_t)q * MAX_SLOTS + s];
      if (c > region_cap) fail = 1;
      offs[s] = total;
      total += c;
    }
    offs[n_slots] = total;
    if (n_old + total > REFINE_MAX) fail = 1;
    s_fail = fail;
    s_prefix = 0; s_remaining = (uint32_t)p.k; s_out = 0;
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) { p.overflow[q] = 1; p.theta[q] = INFINITY; p.surv_cnt[q] = 0; }
    return;
  }
  const int n_new = offs[n_slots];
  const int n = n_old + n_new;
  uint2* mine = p.surv + (size_t)q * SURV_CAP;
  for (int i = tid; i < n_old; i += REFINE_THREADS) {
    const uint2 e = mine[i];
    keys[i] = ordered_from_float_bits(e.x);
    rows[i] = e.y;
  }
  const uint2* pool = p.pool + (size_t)q * POOL_CAP;
  for (int j = tid; j < n_new; j += REFINE_THREADS) {
    int lo = 0, hi = n_slots;  // last slot with offs[slot] <= j
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (offs[mid] <= j) lo = mid; else hi = mid;
    }
    const uint2 e = pool[(size_t)lo * region_cap + (j - offs[lo])];
    keys[n_old + j] = ordered_from_float_bits(e.x);
    rows[n_old + j] = e.y;
  }
  __syncthreads();
  float theta_new = -INFINITY;
  if (n >= p.k) {
    // radix select of the k-th largest ordered key, 8 bits at a time from the top
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += REFINE_THREADS) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
      for (int i = tid; i < n; i += REFINE_THREADS) {
        const uint32_t v = keys[i];
        if ((v & mask) == prefix) atomicAdd(&hist[(v >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid < 32) {
        // find the bin where the count from the top reaches `remaining`: lane l owns bins [8l, 8l+8)
        uint32_t mine8[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { mine8[j] = hist[tid * 8 + j]; sum += mine8[j]; }
        uint32_t above = sum;  // inclusive suffix sum over lanes (lane 31 = top bins)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_down_sync(0xffffffffu, above, o);
          if (tid + o < 32) above += v;
        }
        const uint32_t remaining = s_remaining;
        const unsigned reach = __ballot_sync(0xffffffffu, above >= remaining);
        const int owner = reach ? 31 - __clz(reach) : 0;   // highest lane whose suffix reaches it
        if (tid == owner) {
          uint32_t left = remaining - (above - sum);         // still to find inside this lane's bins
          int b = 7;
          for (; b > 0; --b) {
            if (mine8[b] >= left) break;
            left -= mine8[b];
          }
          s_prefix = prefix | ((uint32_t)(tid * 8 + b) << shift);
          s_remaining = left;
        }
      }
      __syncthreads();
    }
    const float kth = float_from_ordered(s_prefix);
    const double margin = 2.0 * (p.err_a * p.q_norm[q] + p.err_b) * (1.0 + 1e-6) + 1e-37;
    theta_new = __double2float_rd((double)kth - margin);
    if (!(theta_new == theta_new)) theta_new = -INFINITY;
  }
  // survivors (order is irrelevant: the re-rank sorts on exact scores)
  const uint32_t theta_ord = ordered_from_float_bits(__float_as_uint(theta_new));
  for (int i0 = 0; i0 < n; i0 += REFINE_THREADS) {
    const int i = i0 + tid;
    const bool keep = i < n && (theta_new == -INFINITY || keys[i] >= theta_ord);
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    uint32_t base = 0;
    if ((tid & 31) == 0 && ballot) base = atomicAdd(&s_out, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
      const uint32_t pos = base + __popc(ballot & ((1u << (tid & 31)) - 1u));
      if (pos < (uint32_t)SURV_CAP) mine[pos] = make_uint2(__float_as_uint(float_from_ordered(keys[i])), rows[i]);
    }
  }
  __syncthreads();
  if (tid == 0) {
    if (s_out > (uint32_t)SURV_CAP) { p.overflow[q] = 1; p.theta[q] = INFINITY; p.surv_cnt[q] = 0; }
    else { p.surv_cnt[q] = (int)s_out; p.theta[q] = theta_new; }
  }
}

}  // namespace tcs
}  // namespace drag
