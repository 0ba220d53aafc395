// Exact top-k similarity scan for sm_100a.
//
// Replaces, for one row-major matrix shard,
//     distances = ENUM_TO_METRIC[metric](query_f64, docs)      embeddings_metrics.py:14-50
//     np.argsort(distances, kind="stable")[:limit]              embeddings_index.py:58,81
// with one streaming pass over the matrix:
//
//   scan kernel   : persistent CTAs, one warp per row; the row is read once with
//                   coalesced 16-byte loads, widened to float64 and dotted against up to
//                   QB queries held in registers (float64 accumulate, like the
//                   reference's numpy path); the metric turns the dot product into a
//                   "smaller is better" distance which is mapped to an order-preserving
//                   uint64 key.  Each warp keeps a sorted top-k list in shared memory and
//                   only touches it for rows that beat its current k-th key (threshold
//                   filter + 32-entry pending queue + warp-shuffle bitonic merge).
//                   Warp lists are tree-merged per CTA at the end.
//   merge kernel  : one CTA per query merges the per-CTA lists into the final top-k.
//
// Ordering is lexicographic on (key, row id): ascending distance, ties -> lowest row id,
// NaN last -- exactly what a stable argsort over the concatenated matrix yields.
#include "drag_common.cuh"
#include "drag_topk_tc.cuh"

namespace drag {
namespace topk {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint64_t KEY_SENTINEL = ~0ull;

template <typename RowT>
__device__ __forceinline__ RowT row_sentinel() { return (RowT)~(RowT)0; }

// order-preserving map double -> uint64 (NaN greatest, -0 == +0)
__device__ __forceinline__ uint64_t key_from_double(double v) {
  if (v != v) return KEY_SENTINEL;
  v = v + 0.0;
  uint64_t b = (uint64_t)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double double_from_key(uint64_t k) {
  if (k == KEY_SENTINEL) return __longlong_as_double(0x7ff8000000000000ll);
  uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

template <typename RowT>
__device__ __forceinline__ bool entry_less(uint64_t ka, RowT ra, uint64_t kb, RowT rb) {
  return ka < kb || (ka == kb && ra < rb);
}

template <typename RowT>
__device__ __forceinline__ RowT shfl_row(RowT v, int src);
template <>
__device__ __forceinline__ uint32_t shfl_row<uint32_t>(uint32_t v, int src) { return __shfl_sync(FULL, v, src); }
template <>
__device__ __forceinline__ uint64_t shfl_row<uint64_t>(uint64_t v, int src) {
  return (uint64_t)__shfl_sync(FULL, (unsigned long long)v, src);
}

// Merge 32 entries (one per lane, any order; unused lanes carry the sentinel) into a
// warp-private sorted list of the kpad smallest entries (shared memory).  Kept out of line:
// it runs only for the few rows that beat the current threshold.
template <typename RowT>
__device__ __noinline__ void merge_chunk_impl(uint64_t* keys, RowT* rows, int kpad, uint64_t pk, RowT pr) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      uint64_t ok = __shfl_xor_sync(FULL, (unsigned long long)pk, stride);
      RowT orow = shfl_row<RowT>(pr, lane ^ stride);
      bool up = (lane & size) == 0;
      bool lower = (lane & stride) == 0;
      bool take = (lower == up) ? entry_less<RowT>(ok, orow, pk, pr) : entry_less<RowT>(pk, pr, ok, orow);
      if (take) { pk = ok; pr = orow; }
    }
  }
  // 32 smallest of (list tail U chunk): elementwise min of ascending tail and reversed chunk
  uint64_t tk = keys[kpad - 32 + lane];
  RowT tr = rows[kpad - 32 + lane];
  uint64_t rk = __shfl_sync(FULL, (unsigned long long)pk, 31 - lane);
  RowT rr = shfl_row<RowT>(pr, 31 - lane);
  if (entry_less<RowT>(rk, rr, tk, tr)) { tk = rk; tr = rr; }
#pragma unroll
  for (int stride = 16; stride > 0; stride >>= 1) {
    uint64_t ok = __shfl_xor_sync(FULL, (unsigned long long)tk, stride);
    RowT orow = shfl_row<RowT>(tr, lane ^ stride);
    bool lower = (lane & stride) == 0;
    bool take = lower ? entry_less<RowT>(ok, orow, tk, tr) : entry_less<RowT>(tk, tr, ok, orow);
    if (take) { tk = ok; tr = orow; }
  }
  // ascending head + descending tail is bitonic -> one bitonic merge sorts the list
  __syncwarp();
  keys[kpad - 1 - lane] = tk;
  rows[kpad - 1 - lane] = tr;
  __syncwarp();
  for (int j = kpad >> 1; j >= 1; j >>= 1) {
    for (int t = lane; t < (kpad >> 1); t += 32) {
      int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
      uint64_t ka = keys[i], kb = keys[i + j];
      RowT ra = rows[i], rb = rows[i + j];
      if (entry_less<RowT>(kb, rb, ka, ra)) {
        keys[i] = kb; rows[i] = rb; keys[i + j] = ka; rows[i + j] = ra;
      }
    }
    __syncwarp();
  }
}

template <typename RowT>
__device__ __noinline__ void flush_impl(uint64_t* keys, RowT* rows, const uint64_t* pkeys, const RowT* prows,
                                        int kpad, int npend) {
  __syncwarp();
  const uint32_t lane = lane_id();
  uint64_t pk = (int)lane < npend ? pkeys[lane] : KEY_SENTINEL;
  RowT pr = (int)lane < npend ? prows[lane] : row_sentinel<RowT>();
  __syncwarp();
  merge_chunk_impl<RowT>(keys, rows, kpad, pk, pr);
}

// A warp-private sorted top-k list plus a 32-slot queue of entries that passed the
// threshold but are not merged yet.  thr_* / npend live in (warp-uniform) registers.
template <typename RowT>
struct WarpList {
  uint64_t* keys;   // [kpad] ascending
  RowT* rows;       // [kpad]
  uint64_t* pkeys;  // [32]
  RowT* prows;      // [32]
  int kpad, k;
  uint64_t thr_key;  // current k-th entry
  RowT thr_row;
  int npend;

  __device__ __forceinline__ void init(uint64_t* k_, RowT* r_, uint64_t* pk_, RowT* pr_, int kpad_, int kk) {
    keys = k_; rows = r_; pkeys = pk_; prows = pr_; kpad = kpad_; k = kk;
    for (int i = lane_id(); i < kpad; i += 32) { keys[i] = KEY_SENTINEL; rows[i] = row_sentinel<RowT>(); }
    thr_key = KEY_SENTINEL; thr_row = row_sentinel<RowT>(); npend = 0;
    __syncwarp();
  }
  __device__ __forceinline__ void reload_threshold() {
    thr_key = keys[k - 1];
    thr_row = rows[k - 1];
  }
  __device__ __forceinline__ void merge_chunk(uint64_t pk, RowT pr) {
    merge_chunk_impl<RowT>(keys, rows, kpad, pk, pr);
    reload_threshold();
  }
  __device__ __forceinline__ void flush() {
    if (npend == 0) return;
    flush_impl<RowT>(keys, rows, pkeys, prows, kpad, npend);
    reload_threshold();
    npend = 0;
  }
  // key/row are warp-uniform; control flow is warp-uniform
  __device__ __forceinline__ void push(uint64_t key, RowT row) {
    if (entry_less<RowT>(key, row, thr_key, thr_row)) {
      if (lane_id() == 0) { pkeys[npend] = key; prows[npend] = row; }
      if (++npend == 32) flush();
    }
  }
  // Merge another sorted list (ascending, sentinel padded, length n multiple of 32).
  template <typename KeyLoad, typename RowLoad>
  __device__ __forceinline__ void merge_sorted(int n, KeyLoad load_key, RowLoad load_row) {
    const uint32_t lane = lane_id();
    for (int c = 0; c < n; c += 32) {
      uint64_t pk = load_key(c + lane);
      RowT pr = load_row(c + lane);
      uint64_t fk = __shfl_sync(FULL, (unsigned long long)pk, 0);
      RowT fr = shfl_row<RowT>(pr, 0);
      if (!entry_less<RowT>(fk, fr, thr_key, thr_row)) break;  // rest of the list is worse
      merge_chunk(pk, pr);
    }
  }
};

__host__ __device__ inline size_t list_bytes(int kpad, int row_bytes) {
  return (size_t)(kpad + 32) * (8 + row_bytes);
}

// ---------------------------------------------------------------------------------
// row loaders
// ---------------------------------------------------------------------------------
template <typename T> struct RowVec;
template <> struct RowVec<float> {
  static constexpr int EPV = 4;
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
  }
  __device__ static __forceinline__ float scalar(const float* p) { return __ldg(p); }
};
template <> struct RowVec<__nv_bfloat16> {
  static constexpr int EPV = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint32_t w[4];
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ float scalar(const __nv_bfloat16* p) {
    return __uint_as_float(((uint32_t)*reinterpret_cast<const uint16_t*>(p)) << 16);
  }
};

struct ScanArgs {
  const void* mat;
  long long n_rows;
  int dim;
  const float* row_sq;     // [n_rows] or null
  const double* queries;   // first query of this group, [nq, dim]
  const double* q_sq;      // [nq]  sum(q*q) in numpy pairwise order
  const double* q_norm;    // [nq]  max(||q||, eps)
  int nq;                  // valid queries in this group (<= QB)
  int k, kpad;
  uint64_t* part_keys;     // [QB][gridDim.x][kpad]
  long long* part_rows;
  long long row_base;
};

template <int METRIC>
__device__ __forceinline__ double metric_value(double dot, double nn, float row_sq, double q_sq, double q_norm) {
  if (METRIC == DRAG_METRIC_INNER_PRODUCT) return -dot;
  if (METRIC == DRAG_METRIC_COSINE_SIM) {
    double dn = fmax(sqrt(nn), 1e-8);
    return -(dot / (dn * q_norm));
  }
  double sq = ((double)row_sq - 2.0 * dot) + q_sq;
  if (METRIC == DRAG_METRIC_EUCLIDEAN_DIST) return sqrt(sq);
  return sq;
}

// CTA-level epilogue shared by both scan kernels: flush, tree-merge warp lists, write out.
template <int QB, int NWARPS>
__device__ void block_finish(WarpList<uint32_t> (&lists)[QB], const ScanArgs& a, unsigned char* smem) {
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const size_t per_list = list_bytes(a.kpad, 4);
  if (warp < NWARPS) {
#pragma unroll
    for (int q = 0; q < QB; ++q) lists[q].flush();
  }
  for (int step = 1; step < NWARPS; step <<= 1) {
    __syncthreads();   // (a producer warp beyond NWARPS only takes part in the barriers)
    if ((warp % (2 * step)) == 0 && warp + step < NWARPS) {
#pragma unroll
      for (int q = 0; q < QB; ++q) {
        unsigned char* other = smem + ((size_t)(warp + step) * QB + q) * per_list;
        const uint64_t* ok = reinterpret_cast<const uint64_t*>(other);
        const uint32_t* orow = reinterpret_cast<const uint32_t*>(other + (size_t)(a.kpad + 32) * 8);
        lists[q].merge_sorted(a.kpad, [&](int i) { return ok[i]; }, [&](int i) { return orow[i]; });
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      if (q >= a.nq) break;
      size_t base = ((size_t)q * gridDim.x + blockIdx.x) * a.kpad;
      for (int i = lane; i < a.kpad; i += 32) {
        uint32_t r = lists[q].rows[i];
        a.part_keys[base + i] = lists[q].keys[i];
        a.part_rows[base + i] = (r == 0xffffffffu) ? -1ll : a.row_base + (long long)r;
      }
    }
  }
}

template <int QB>
__device__ __forceinline__ void init_lists(WarpList<uint32_t> (&lists)[QB], unsigned char* smem, int kpad, int k) {
  const int warp = threadIdx.x >> 5;
  const size_t per_list = list_bytes(kpad, 4);
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    unsigned char* base = smem + ((size_t)warp * QB + q) * per_list;
    uint64_t* keys = reinterpret_cast<uint64_t*>(base);
    uint64_t* pkeys = keys + kpad;
    uint32_t* rows = reinterpret_cast<uint32_t*>(base + (size_t)(kpad + 32) * 8);
    uint32_t* prows = rows + kpad;
    lists[q].init(keys, rows, pkeys, prows, kpad, k);
  }
}

// Vectorised scan: dim*sizeof(T) % 16 == 0, dim <= NCH*32*EPV.
template <typename T, int NCH, int QB, int R, int NWARPS, int METRIC>
__global__ void __launch_bounds__(NWARPS * 32, 1) scan_vec_kernel(ScanArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int EPV = RowVec<T>::EPV;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const T* mat = reinterpret_cast<const T*>(a.mat);

  WarpList<uint32_t> lists[QB];
  init_lists<QB>(lists, smem, a.kpad, a.k);

  double qreg[QB][NCH * EPV];
#pragma unroll
  for (int q = 0; q < QB; ++q)
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < EPV; ++j) {
        int e = c * 32 * EPV + lane * EPV + j;
        qreg[q][c * EPV + j] = (q < a.nq && e < a.dim) ? a.queries[(size_t)q * a.dim + e] : 0.0;
      }
  double qsq[QB], qnorm[QB];
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    qsq[q] = q < a.nq ? a.q_sq[q] : 0.0;
    qnorm[q] = q < a.nq ? a.q_norm[q] : 1.0;
  }

  const long long tile = (long long)NWARPS * R;
  for (long long base = (long long)blockIdx.x * tile; base < a.n_rows; base += (long long)gridDim.x * tile) {
    const long long row0 = base + (long long)warp * R;
    float v[R][NCH][EPV];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        int e = c * 32 * EPV + lane * EPV;
        if (row0 + r < a.n_rows && e < a.dim) {
          RowVec<T>::load(mat + (size_t)(row0 + r) * a.dim + e, v[r][c]);
        } else {
#pragma unroll
          for (int j = 0; j < EPV; ++j) v[r][c][j] = 0.f;
        }
      }
    float rsq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rsq[r] = 0.f;
      if ((METRIC == DRAG_METRIC_SQEUCLIDEAN_DIST || METRIC == DRAG_METRIC_EUCLIDEAN_DIST) && row0 + r < a.n_rows)
        rsq[r] = __ldg(a.row_sq + row0 + r);
    }
    double acc[R][QB], nn[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      nn[r] = 0.0;
#pragma unroll
      for (int q = 0; q < QB; ++q) acc[r][q] = 0.0;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < EPV; ++j) {
          double x = (double)v[r][c][j];
#pragma unroll
          for (int q = 0; q < QB; ++q) acc[r][q] = __fma_rn(x, qreg[q][c * EPV + j], acc[r][q]);
          if (METRIC == DRAG_METRIC_COSINE_SIM) nn[r] = __fma_rn(x, x, nn[r]);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int q = 0; q < QB; ++q) acc[r][q] += __shfl_xor_sync(FULL, acc[r][q], off);
        if (METRIC == DRAG_METRIC_COSINE_SIM) nn[r] += __shfl_xor_sync(FULL, nn[r], off);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r >= a.n_rows) break;
#pragma unroll
      for (int q = 0; q < QB; ++q) {
        if (q >= a.nq) break;
        double val = metric_value<METRIC>(acc[r][q], nn[r], rsq[r], qsq[q], qnorm[q]);
        lists[q].push(key_from_double(val), (uint32_t)(row0 + r));
      }
    }
  }
  block_finish<QB, NWARPS>(lists, a, smem);
}

// ---------------------------------------------------------------------------------
// Ring scan: the same arithmetic as scan_vec_kernel (same element order, same reduction: bit-identical
// scores), but the rows are STREAMED THROUGH SHARED MEMORY: a producer warp issues one bulk async copy
// (cp.async.bulk, mbarrier complete_tx) per tile of NWARPS*R consecutive rows -- a tile of a row-major matrix
// is one contiguous range -- into an n_stages-deep ring, so 100-190 KB of reads are in flight per SM
// whatever the compute warps are doing (the register-loading kernel leaves its loads exposed between
// tiles: long_scoreboard was its top stall at 63 % of the DRAM peak).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ring_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ring_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr_u32(bar)) : "memory");
}
__device__ __forceinline__ void ring_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ring_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_addr_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) {   // a protocol bug traps instead of hanging the GPU
      printf("drag_b200: scan ring wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void ring_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr_u32(bar)) : "memory");
}
template <typename T> struct RowVecShared;
template <> struct RowVecShared<float> {
  __device__ static __forceinline__ void load(const unsigned char* p, float (&v)[4]) {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(smem_addr_u32(p)));
  }
};
template <> struct RowVecShared<__nv_bfloat16> {
  __device__ static __forceinline__ void load(const unsigned char* p, float (&v)[8]) {
    uint32_t w[4];
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(smem_addr_u32(p)));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// block = (NWARPS + 1) warps: NWARPS compute warps (one warp per row, R rows per tile) + the producer.
// dynamic smem = lists (as scan_vec_kernel) | ring n_stages x (NWARPS*R rows) | barriers
template <typename T, int NCH, int QB, int R, int NWARPS, int METRIC>
__global__ void __launch_bounds__((NWARPS + 1) * 32, 1) scan_ring_kernel(ScanArgs a, int n_stages, unsigned lists_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int EPV = RowVec<T>::EPV;
  constexpr int TR = NWARPS * R;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const size_t row_bytes = (size_t)a.dim * sizeof(T);
  const size_t stage_bytes = (size_t)TR * row_bytes;
  unsigned char* ring = smem + lists_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)n_stages * stage_bytes);
  uint64_t* empty_bar = full_bar + n_stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      ring_mbar_init(&full_bar[s], 1);
      ring_mbar_init(&empty_bar[s], NWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  WarpList<uint32_t> lists[QB];
  if (warp < NWARPS) init_lists<QB>(lists, smem, a.kpad, a.k);
  __syncthreads();

  const long long tile_stride = (long long)gridDim.x * TR;
  if (warp == NWARPS) {
    // ===================== producer =====================
    if (lane == 0) {
      const unsigned char* mat = reinterpret_cast<const unsigned char*>(a.mat);
      int stage = 0;
      uint32_t phase = 0;
      for (long long base = (long long)blockIdx.x * TR; base < a.n_rows; base += tile_stride) {
        ring_mbar_wait(&empty_bar[stage], phase ^ 1);
        const long long rows_here = a.n_rows - base < TR ? a.n_rows - base : TR;
        const uint32_t bytes = (uint32_t)(rows_here * (long long)row_bytes);
        ring_mbar_expect_tx(&full_bar[stage], bytes);
        ring_bulk_load(ring + (size_t)stage * stage_bytes, mat + (size_t)base * row_bytes, bytes, &full_bar[stage]);
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== compute warps =====================
    double qreg[QB][NCH * EPV];
#pragma unroll
    for (int q = 0; q < QB; ++q)
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < EPV; ++j) {
          int e = c * 32 * EPV + lane * EPV + j;
          qreg[q][c * EPV + j] = (q < a.nq && e < a.dim) ? a.queries[(size_t)q * a.dim + e] : 0.0;
        }
    double qsq[QB], qnorm[QB];
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      qsq[q] = q < a.nq ? a.q_sq[q] : 0.0;
      qnorm[q] = q < a.nq ? a.q_norm[q] : 1.0;
    }
    int stage = 0;
    uint32_t phase = 0;
    for (long long base = (long long)blockIdx.x * TR; base < a.n_rows; base += tile_stride) {
      const long long row0 = base + (long long)warp * R;
      ring_mbar_wait(&full_bar[stage], phase);
      const unsigned char* tile = ring + (size_t)stage * stage_bytes + (size_t)(warp * R) * row_bytes;
      float v[R][NCH][EPV];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          int e = c * 32 * EPV + lane * EPV;
          if (row0 + r < a.n_rows && e < a.dim) {
            RowVecShared<T>::load(tile + (size_t)r * row_bytes + (size_t)e * sizeof(T), v[r][c]);
          } else {
#pragma unroll
            for (int j = 0; j < EPV; ++j) v[r][c][j] = 0.f;
          }
        }
      // the rows are in registers: hand the slot back to the producer
      __syncwarp();
      if (lane == 0) ring_mbar_arrive(&empty_bar[stage]);
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
      float rsq[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        rsq[r] = 0.f;
        if ((METRIC == DRAG_METRIC_SQEUCLIDEAN_DIST || METRIC == DRAG_METRIC_EUCLIDEAN_DIST) && row0 + r < a.n_rows)
          rsq[r] = __ldg(a.row_sq + row0 + r);
      }
      double acc[R][QB], nn[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        nn[r] = 0.0;
#pragma unroll
        for (int q = 0; q < QB; ++q) acc[r][q] = 0.0;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int j = 0; j < EPV; ++j) {
            double x = (double)v[r][c][j];
#pragma unroll
            for (int q = 0; q < QB; ++q) acc[r][q] = __fma_rn(x, qreg[q][c * EPV + j], acc[r][q]);
            if (METRIC == DRAG_METRIC_COSINE_SIM) nn[r] = __fma_rn(x, x, nn[r]);
          }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
          for (int q = 0; q < QB; ++q) acc[r][q] += __shfl_xor_sync(FULL, acc[r][q], off);
          if (METRIC == DRAG_METRIC_COSINE_SIM) nn[r] += __shfl_xor_sync(FULL, nn[r], off);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (row0 + r >= a.n_rows) break;
#pragma unroll
        for (int q = 0; q < QB; ++q) {
          if (q >= a.nq) break;
          double val = metric_value<METRIC>(acc[r][q], nn[r], rsq[r], qsq[q], qnorm[q]);
          lists[q].push(key_from_double(val), (uint32_t)(row0 + r));
        }
      }
    }
  }
  block_finish<QB, NWARPS>(lists, a, smem);
}

// Generic scan: any dim / alignment; lane l owns elements l, l+32, ...
template <typename T, int QB, int NWARPS, int METRIC>
__global__ void __launch_bounds__(NWARPS * 32, 1) scan_generic_kernel(ScanArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const T* mat = reinterpret_cast<const T*>(a.mat);
  WarpList<uint32_t> lists[QB];
  init_lists<QB>(lists, smem, a.kpad, a.k);
  double qsq[QB], qnorm[QB];
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    qsq[q] = q < a.nq ? a.q_sq[q] : 0.0;
    qnorm[q] = q < a.nq ? a.q_norm[q] : 1.0;
  }
  for (long long row = (long long)blockIdx.x * NWARPS + warp; row < a.n_rows; row += (long long)gridDim.x * NWARPS) {
    double acc[QB], nn = 0.0;
#pragma unroll
    for (int q = 0; q < QB; ++q) acc[q] = 0.0;
    const T* rp = mat + (size_t)row * a.dim;
    for (int e = lane; e < a.dim; e += 32) {
      double x = (double)RowVec<T>::scalar(rp + e);
#pragma unroll
      for (int q = 0; q < QB; ++q)
        if (q < a.nq) acc[q] = __fma_rn(x, __ldg(a.queries + (size_t)q * a.dim + e), acc[q]);
      if (METRIC == DRAG_METRIC_COSINE_SIM) nn = __fma_rn(x, x, nn);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int q = 0; q < QB; ++q) acc[q] += __shfl_xor_sync(FULL, acc[q], off);
      if (METRIC == DRAG_METRIC_COSINE_SIM) nn += __shfl_xor_sync(FULL, nn, off);
    }
    float rsq = 0.f;
    if (METRIC == DRAG_METRIC_SQEUCLIDEAN_DIST || METRIC == DRAG_METRIC_EUCLIDEAN_DIST) rsq = __ldg(a.row_sq + row);
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      if (q >= a.nq) break;
      double val = metric_value<METRIC>(acc[q], nn, rsq, qsq[q], qnorm[q]);
      lists[q].push(key_from_double(val), (uint32_t)row);
    }
  }
  block_finish<QB, NWARPS>(lists, a, smem);
}

// ---------------------------------------------------------------------------------
// final merge: one CTA per query, lists of `list_len` (multiple of 32) sorted entries
// ---------------------------------------------------------------------------------
struct MergeArgs {
  const uint64_t* in_keys;   // key form  [q][n_lists][list_len]         (FROM_DOUBLE = false)
  const long long* in_rows;
  const double* in_dist;     // double form [n_lists][nq][k] + counts     (FROM_DOUBLE = true)
  const int* in_count;       // [n_lists][nq]
  int n_lists, list_len, nq, k, kpad;
  double* out_dist;
  long long* out_row;
  int* out_count;
};

template <bool FROM_DOUBLE, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) merge_kernel(MergeArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int q = blockIdx.x;
  const size_t per_list = list_bytes(a.kpad, 8);
  unsigned char* base = smem + (size_t)warp * per_list;
  uint64_t* keys = reinterpret_cast<uint64_t*>(base);
  uint64_t* pkeys = keys + a.kpad;
  uint64_t* rows = pkeys + 32;
  uint64_t* prows = rows + a.kpad;
  WarpList<uint64_t> list;
  list.init(keys, rows, pkeys, prows, a.kpad, a.k);

  for (int l = warp; l < a.n_lists; l += NWARPS) {
    if (FROM_DOUBLE) {
      const size_t off = ((size_t)l * a.nq + q) * a.k;
      const int cnt = a.in_count[(size_t)l * a.nq + q];
      const int padded = (a.k + 31) & ~31;
      list.merge_sorted(
          padded,
          [&](int i) { return i < cnt ? key_from_double(a.in_dist[off + i]) : KEY_SENTINEL; },
          [&](int i) { return i < cnt ? (uint64_t)a.in_rows[off + i] : ~0ull; });
    } else {
      const size_t off = ((size_t)q * a.n_lists + l) * a.list_len;
      list.merge_sorted(
          a.list_len, [&](int i) { return a.in_keys[off + i]; },
          [&](int i) { return (uint64_t)a.in_rows[off + i]; });
    }
  }
  for (int step = 1; step < NWARPS; step <<= 1) {
    __syncthreads();
    if ((warp % (2 * step)) == 0 && warp + step < NWARPS) {
      unsigned char* other = smem + (size_t)(warp + step) * per_list;
      const uint64_t* ok = reinterpret_cast<const uint64_t*>(other);
      const uint64_t* orow = ok + a.kpad + 32;
      list.merge_sorted(a.kpad, [&](int i) { return ok[i]; }, [&](int i) { return orow[i]; });
    }
  }
  __syncthreads();
  if (warp == 0) {
    int cnt = 0;
    for (int i = lane; i < a.k; i += 32) {
      uint64_t r = list.rows[i];
      bool valid = r != ~0ull;
      a.out_dist[(size_t)q * a.k + i] = valid ? double_from_key(list.keys[i]) : __longlong_as_double(0x7ff8000000000000ll);
      a.out_row[(size_t)q * a.k + i] = valid ? (long long)r : -1ll;
      cnt += valid ? 1 : 0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(FULL, cnt, off);
    if (lane == 0) a.out_count[q] = cnt;
  }
}

// ---------------------------------------------------------------------------------
// numpy-order sums (embeddings_metrics.py:40-41)
// ---------------------------------------------------------------------------------
template <typename T>
__device__ float pairwise_sq_f32(const T* a, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) { float x = RowVec<T>::scalar(a + i); r = __fadd_rn(r, __fmul_rn(x, x)); }
    return r;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { float x = RowVec<T>::scalar(a + j); r[j] = __fmul_rn(x, x); }
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { float x = RowVec<T>::scalar(a + i + j); r[j] = __fadd_rn(r[j], __fmul_rn(x, x)); }
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) { float x = RowVec<T>::scalar(a + i); res = __fadd_rn(res, __fmul_rn(x, x)); }
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __fadd_rn(pairwise_sq_f32<T>(a, n2), pairwise_sq_f32<T>(a + n2, n - n2));
}

__device__ double pairwise_sq_f64(const double* a, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, __dmul_rn(a[i], a[i]));
    return r;
  }
  if (n <= 128) {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dmul_rn(a[j], a[j]);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], __dmul_rn(a[i + j], a[i + j]));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, __dmul_rn(a[i], a[i]));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(pairwise_sq_f64(a, n2), pairwise_sq_f64(a + n2, n - n2));
}

template <typename T>
__global__ void row_sqnorm_kernel(const T* mat, long long n_rows, int dim, float* out) {
  long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row < n_rows) out[row] = pairwise_sq_f32<T>(mat + (size_t)row * dim, dim);
}

__global__ void query_prep_kernel(const double* queries, int nq, int dim, double* q_sq, double* q_norm) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const double* p = queries + (size_t)q * dim;
  q_sq[q] = pairwise_sq_f64(p, dim);
  double s = 0.0;
  for (int i = 0; i < dim; ++i) s = __fma_rn(p[i], p[i], s);
  q_norm[q] = fmax(sqrt(s), 1e-8);
}

// One block per query: the query is staged in shared memory by the whole block (coalesced), then one thread
// forms the two sums in their fixed orders (numpy pairwise for q_sq) without a global-load latency per element.
__global__ void __launch_bounds__(128) query_prep_smem_kernel(const double* queries, int nq, int dim, double* q_sq, double* q_norm) {
  extern __shared__ __align__(16) double qs[];
  const int q = blockIdx.x;
  if (q >= nq) return;
  const double* p = queries + (size_t)q * dim;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) qs[i] = p[i];
  __syncthreads();
  // the two sums run side by side in two warps (each is one dependent float64 chain)
  if (threadIdx.x == 0) q_sq[q] = pairwise_sq_f64(qs, dim);
  if (threadIdx.x == 32) {
    double s = 0.0;
    for (int i = 0; i < dim; ++i) s = __fma_rn(qs[i], qs[i], s);
    q_norm[q] = fmax(sqrt(s), 1e-8);
  }
}

static int launch_query_prep(const double* d_queries, int nq, int dim, double* q_sq, double* q_norm, cudaStream_t st) {
  if ((size_t)dim * 8 <= 48 * 1024) query_prep_smem_kernel<<<nq, 128, (size_t)dim * 8, st>>>(d_queries, nq, dim, q_sq, q_norm);
  else query_prep_kernel<<<(nq + 63) / 64, 64, 0, st>>>(d_queries, nq, dim, q_sq, q_norm);
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

__global__ void rows_to_chunks_kernel(const long long* rows, long long n, const long long* doc_offsets, int n_docs,
                                      const long long* chunk_ids, long long* out_doc, long long* out_chunk) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long r = rows[i];
  if (r < 0) { out_doc[i] = -1; out_chunk[i] = -1; return; }
  int lo = 0, hi = n_docs;  // last d with doc_offsets[d] <= r
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (doc_offsets[mid] <= r) lo = mid; else hi = mid;
  }
  out_doc[i] = lo;
  out_chunk[i] = chunk_ids ? chunk_ids[r] : r - doc_offsets[lo];
}

// All distances of one query (the reference's ENUM_TO_METRIC[metric](query, docs)); one warp per row.
template <typename T, int METRIC>
__global__ void distances_kernel(const T* mat, long long n_rows, int dim, const float* row_sq, const double* query,
                                 const double* q_sq, const double* q_norm, double* out) {
  const uint32_t lane = lane_id();
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n_rows; row += warps) {
    const T* rp = mat + (size_t)row * dim;
    double acc = 0.0, nn = 0.0;
    for (int e = lane; e < dim; e += 32) {
      double x = (double)RowVec<T>::scalar(rp + e);
      acc = __fma_rn(x, __ldg(query + e), acc);
      if (METRIC == DRAG_METRIC_COSINE_SIM) nn = __fma_rn(x, x, nn);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      acc += __shfl_xor_sync(FULL, acc, off);
      if (METRIC == DRAG_METRIC_COSINE_SIM) nn += __shfl_xor_sync(FULL, nn, off);
    }
    float rsq = 0.f;
    if (METRIC == DRAG_METRIC_SQEUCLIDEAN_DIST || METRIC == DRAG_METRIC_EUCLIDEAN_DIST) rsq = __ldg(row_sq + row);
    if (lane == 0) out[row] = metric_value<METRIC>(acc, nn, rsq, q_sq[0], q_norm[0]);
  }
}

template <typename T>
static int launch_distances(int metric, const T* mat, long long n_rows, int dim, const float* row_sq,
                            const double* query, const double* q_sq, const double* q_norm, double* out, cudaStream_t st) {
  long long blocks = (n_rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  switch (metric) {
    case DRAG_METRIC_COSINE_SIM:
      distances_kernel<T, DRAG_METRIC_COSINE_SIM><<<(unsigned)blocks, 256, 0, st>>>(mat, n_rows, dim, row_sq, query, q_sq, q_norm, out); break;
    case DRAG_METRIC_EUCLIDEAN_DIST:
      distances_kernel<T, DRAG_METRIC_EUCLIDEAN_DIST><<<(unsigned)blocks, 256, 0, st>>>(mat, n_rows, dim, row_sq, query, q_sq, q_norm, out); break;
    case DRAG_METRIC_SQEUCLIDEAN_DIST:
      distances_kernel<T, DRAG_METRIC_SQEUCLIDEAN_DIST><<<(unsigned)blocks, 256, 0, st>>>(mat, n_rows, dim, row_sq, query, q_sq, q_norm, out); break;
    default:
      distances_kernel<T, DRAG_METRIC_INNER_PRODUCT><<<(unsigned)blocks, 256, 0, st>>>(mat, n_rows, dim, row_sq, query, q_sq, q_norm, out); break;
  }
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static int next_pow2_min32(int k) {
  int p = 32;
  while (p < k) p <<= 1;
  return p;
}

constexpr int MAX_K = 2048;
constexpr int MERGE_WARPS = 8;
constexpr int MERGE_WARPS_WIDE = 32;   // short lists: every warp walks fewer of the per-CTA lists (their loads are serial per warp)
constexpr size_t SMEM_BUDGET = 200 * 1024;

struct Plan {
  int kpad;
  int qb;       // queries per pass
  int nwarps;   // warps per CTA
  int grid;     // CTAs
  size_t smem;
};

static Plan make_plan(int device, int n_queries, int k) {
  Plan p;
  p.kpad = next_pow2_min32(k);
  p.qb = n_queries >= 4 ? 4 : (n_queries >= 2 ? 2 : 1);
  p.nwarps = (p.qb == 4) ? 8 : 16;
  while (p.nwarps > 1 && (size_t)p.nwarps * p.qb * list_bytes(p.kpad, 4) > SMEM_BUDGET) {
    if (p.qb > 1) p.qb >>= 1; else p.nwarps >>= 1;
  }
  // re-grow warps if shrinking qb left room (keeps the template set small: 16 or 8 or fewer)
  p.smem = (size_t)p.nwarps * p.qb * list_bytes(p.kpad, 4);
  p.grid = sm_count(device);
  return p;
}

template <typename T, int QB, int NWARPS, int METRIC>
static int launch_scan_t(const ScanArgs& a, const Plan& p, bool vec_ok, cudaStream_t st) {
  constexpr int EPV = RowVec<T>::EPV;
  constexpr int R = (QB == 1) ? 4 : 2;
  const int chunk = 32 * EPV;
  const int nch = (a.dim + chunk - 1) / chunk;
  dim3 grid(p.grid), block(NWARPS * 32);
#define DRAG_LAUNCH_VEC(NCH)                                                                       \
  {                                                                                                \
    auto kern = scan_vec_kernel<T, NCH, QB, R, NWARPS, METRIC>;                                    \
    DRAG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem)); \
    kern<<<grid, block, p.smem, st>>>(a);                                                          \
  }
  // ring kernel: rows streamed through shared memory (needs >= 2 stages next to the warp lists)
  // Single-query passes only: measured 2.67 vs 3.16 ms over 10M x 384 fp32 (5.75 TB/s, 88 % of the HBM copy peak); with 2-4
  // queries per pass the float64 work per row dominates and the extra producer warp costs registers (13.3 vs 7.7 ms).
  constexpr int RING_R = 2;
  constexpr int RING_WARPS = NWARPS;
  constexpr size_t RING_SMEM = 226 * 1024;
  const size_t stage_bytes = (size_t)RING_WARPS * RING_R * a.dim * sizeof(T);
  const size_t lists_bytes = (p.smem + 127) / 128 * 128;
  int n_stages = lists_bytes + 256 < RING_SMEM ? (int)((RING_SMEM - lists_bytes - 256) / stage_bytes) : 0;
  if (n_stages > 6) n_stages = 6;
  static const bool ring_off = getenv("DRAG_SCAN_RING") && atoi(getenv("DRAG_SCAN_RING")) == 0;
#define DRAG_LAUNCH_RING(NCH)                                                                      \
  {                                                                                                \
    auto kern = scan_ring_kernel<T, NCH, QB, RING_R, RING_WARPS, METRIC>;                          \
    const size_t smem = lists_bytes + (size_t)n_stages * stage_bytes + 256;                        \
    DRAG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RING_SMEM)); \
    kern<<<grid, dim3((RING_WARPS + 1) * 32), smem, st>>>(a, n_stages, (unsigned)lists_bytes);      \
  }
  bool ring_done = false;
  if constexpr (QB == 1) {   // (only the single-query ring kernels are instantiated)
    if (vec_ok && nch <= 4 && n_stages >= 2 && !ring_off && ((uintptr_t)a.mat & 15) == 0) {
      switch (nch) {
        case 1: DRAG_LAUNCH_RING(1) break;
        case 2: DRAG_LAUNCH_RING(2) break;
        case 3: DRAG_LAUNCH_RING(3) break;
        default: DRAG_LAUNCH_RING(4) break;
      }
      ring_done = true;
    }
  }
  if (ring_done) {
  } else if (vec_ok && nch <= 4) {
    switch (nch) {
      case 1: DRAG_LAUNCH_VEC(1) break;
      case 2: DRAG_LAUNCH_VEC(2) break;
      case 3: DRAG_LAUNCH_VEC(3) break;
      default: DRAG_LAUNCH_VEC(4) break;
    }
  } else {
    auto kern = scan_generic_kernel<T, QB, NWARPS, METRIC>;
    DRAG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    kern<<<grid, block, p.smem, st>>>(a);
  }
#undef DRAG_LAUNCH_VEC
#undef DRAG_LAUNCH_RING
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

template <typename T, int QB, int NWARPS>
static int launch_scan_m(int metric, const ScanArgs& a, const Plan& p, bool vec_ok, cudaStream_t st) {
  switch (metric) {
    case DRAG_METRIC_COSINE_SIM: return launch_scan_t<T, QB, NWARPS, DRAG_METRIC_COSINE_SIM>(a, p, vec_ok, st);
    case DRAG_METRIC_EUCLIDEAN_DIST: return launch_scan_t<T, QB, NWARPS, DRAG_METRIC_EUCLIDEAN_DIST>(a, p, vec_ok, st);
    case DRAG_METRIC_SQEUCLIDEAN_DIST: return launch_scan_t<T, QB, NWARPS, DRAG_METRIC_SQEUCLIDEAN_DIST>(a, p, vec_ok, st);
    default: return launch_scan_t<T, QB, NWARPS, DRAG_METRIC_INNER_PRODUCT>(a, p, vec_ok, st);
  }
}

template <typename T>
static int launch_scan(int metric, const ScanArgs& a, const Plan& p, bool vec_ok, cudaStream_t st) {
  // (qb, nwarps) combinations make_plan can produce
  if (p.qb == 4 && p.nwarps == 8) return launch_scan_m<T, 4, 8>(metric, a, p, vec_ok, st);
  if (p.qb == 2 && p.nwarps == 16) return launch_scan_m<T, 2, 16>(metric, a, p, vec_ok, st);
  if (p.qb == 2 && p.nwarps == 8) return launch_scan_m<T, 2, 8>(metric, a, p, vec_ok, st);
  if (p.qb == 1 && p.nwarps == 16) return launch_scan_m<T, 1, 16>(metric, a, p, vec_ok, st);
  if (p.qb == 1 && p.nwarps == 8) return launch_scan_m<T, 1, 8>(metric, a, p, vec_ok, st);
  if (p.qb == 1 && p.nwarps == 4) return launch_scan_m<T, 1, 4>(metric, a, p, vec_ok, st);
  return fail(DRAG_ERR_UNSUPPORTED, "no scan kernel for qb=%d nwarps=%d", p.qb, p.nwarps);
}

template <bool FROM_DOUBLE>
static int launch_merge(const MergeArgs& m, int n_blocks, cudaStream_t st) {
  const size_t wide = (size_t)MERGE_WARPS_WIDE * list_bytes(m.kpad, 8);
  if (wide <= SMEM_BUDGET) {
    DRAG_CUDA_OK(cudaFuncSetAttribute(merge_kernel<FROM_DOUBLE, MERGE_WARPS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide));
    merge_kernel<FROM_DOUBLE, MERGE_WARPS_WIDE><<<n_blocks, MERGE_WARPS_WIDE * 32, wide, st>>>(m);
  } else {
    const size_t smem = (size_t)MERGE_WARPS * list_bytes(m.kpad, 8);
    DRAG_CUDA_OK(cudaFuncSetAttribute(merge_kernel<FROM_DOUBLE, MERGE_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_kernel<FROM_DOUBLE, MERGE_WARPS><<<n_blocks, MERGE_WARPS * 32, smem, st>>>(m);
  }
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

struct Workspace {
  double* q_sq;
  double* q_norm;
  uint64_t* part_keys;
  long long* part_rows;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t carve(const Plan& p, int n_queries, void* base, Workspace* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* ptr = base ? (void*)((unsigned char*)base + off) : nullptr;
    off += align_up(bytes, 256);
    return ptr;
  };
  double* q_sq = (double*)take((size_t)n_queries * 8);
  double* q_norm = (double*)take((size_t)n_queries * 8);
  uint64_t* pk = (uint64_t*)take((size_t)p.qb * p.grid * p.kpad * 8);
  long long* pr = (long long*)take((size_t)p.qb * p.grid * p.kpad * 8);
  if (ws) { ws->q_sq = q_sq; ws->q_norm = q_norm; ws->part_keys = pk; ws->part_rows = pr; }
  return off;
}


// ---------------------------------------------------------------------------------
// batched path (drag_topk_tc.cuh): exact float64 re-rank of the surviving candidates.
// One CTA per query; a warp scores one candidate row with the arithmetic of scan_vec_kernel
// (same element order, same xor-shuffle reduction => bit-identical float64 scores), then the
// CTA sorts (key, row) in shared memory and writes the first k.
// ---------------------------------------------------------------------------------
struct RerankArgs {
  const void* mat;
  long long n_rows;
  int dim;
  const float* row_sq;
  const double* queries;
  const double* q_sq;
  const double* q_norm;
  const uint2* cand;   // [Qpad][SURV_CAP] survivors of the last refine
  const int* cnt;
  const int* overflow;
  int cap, k;
  long long row_base;
  double* out_dist;
  long long* out_row;
  int* out_count;
  int* out_status;
};

constexpr int RERANK_WARPS = 8;

template <typename T, int NCH, int METRIC>
__global__ void __launch_bounds__(RERANK_WARPS * 32) rerank_kernel(RerankArgs a) {
  constexpr int EPV = RowVec<T>::EPV;
  __shared__ uint64_t skeys[tcs::RERANK_MAX];
  __shared__ uint32_t srows[tcs::RERANK_MAX];
  __shared__ int s_bad;
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const T* mat = reinterpret_cast<const T*>(a.mat);
  const int m = a.cnt[q];
  const long long want_ll = a.n_rows < (long long)a.k ? a.n_rows : (long long)a.k;
  const int want = (int)want_ll;
  if (threadIdx.x == 0) s_bad = 0;
  if (a.overflow[q] || m > tcs::RERANK_MAX || m < want) {
    if (threadIdx.x == 0) { a.out_status[q] = 1; a.out_count[q] = 0; }
    return;
  }
  __syncthreads();
  double qreg[NCH * EPV];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < EPV; ++j) {
      const int e = c * 32 * EPV + lane * EPV + j;
      qreg[c * EPV + j] = e < a.dim ? a.queries[(size_t)q * a.dim + e] : 0.0;
    }
  const double qsq = a.q_sq[q], qnorm = a.q_norm[q];
  const uint2* mine = a.cand + (size_t)q * a.cap;
  for (int i = warp; i < m; i += RERANK_WARPS) {
    const uint32_t row = mine[i].y;
    double acc = 0.0, nn = 0.0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int e = c * 32 * EPV + lane * EPV;
      float v[EPV];
      if (e < a.dim) {
        RowVec<T>::load(mat + (size_t)row * a.dim + e, v);
      } else {
#pragma unroll
        for (int j = 0; j < EPV; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < EPV; ++j) {
        const double x = (double)v[j];
        acc = __fma_rn(x, qreg[c * EPV + j], acc);
        if (METRIC == DRAG_METRIC_COSINE_SIM) nn = __fma_rn(x, x, nn);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      acc += __shfl_xor_sync(FULL, acc, off);
      if (METRIC == DRAG_METRIC_COSINE_SIM) nn += __shfl_xor_sync(FULL, nn, off);
    }
    float rsq = 0.f;
    if (METRIC == DRAG_METRIC_SQEUCLIDEAN_DIST || METRIC == DRAG_METRIC_EUCLIDEAN_DIST) rsq = __ldg(a.row_sq + row);
    const double val = metric_value<METRIC>(acc, nn, rsq, qsq, qnorm);
    if (lane == 0) {
      skeys[i] = key_from_double(val);
      srows[i] = row;
      // a NaN distance (sqrt of a negative residue, NaN input) is ordered last by the reference; the
      // candidate filter cannot certify that case, so the caller re-runs the query through the scan
      if (val != val) s_bad = 1;
    }
  }
  int padded = 32;
  while (padded < m) padded <<= 1;
  for (int i = m + threadIdx.x; i < padded; i += RERANK_WARPS * 32) { skeys[i] = KEY_SENTINEL; srows[i] = 0xffffffffu; }
  __syncthreads();
  if (s_bad) {
    if (threadIdx.x == 0) { a.out_status[q] = 1; a.out_count[q] = 0; }
    return;
  }
  for (int size = 2; size <= padded; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (padded >> 1); t += RERANK_WARPS * 32) {
        const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int j = i + stride;
        const bool up = (i & size) == 0;
        const uint64_t ka = skeys[i], kb = skeys[j];
        const uint32_t ra = srows[i], rb = srows[j];
        const bool a_less = entry_less<uint32_t>(ka, ra, kb, rb);
        if (up ? !a_less : a_less) { skeys[i] = kb; srows[i] = rb; skeys[j] = ka; srows[j] = ra; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < a.k; i += RERANK_WARPS * 32) {
    const bool valid = i < want;
    a.out_dist[(size_t)q * a.k + i] = valid ? double_from_key(skeys[i]) : __longlong_as_double(0x7ff8000000000000ll);
    a.out_row[(size_t)q * a.k + i] = valid ? a.row_base + (long long)srows[i] : -1ll;
  }
  if (threadIdx.x == 0) { a.out_count[q] = want; a.out_status[q] = 0; }
}

template <typename T, int NCH>
static int launch_rerank_m(int metric, const RerankArgs& a, int n_queries, cudaStream_t st) {
  dim3 grid(n_queries), block(RERANK_WARPS * 32);
  switch (metric) {
    case DRAG_METRIC_COSINE_SIM: rerank_kernel<T, NCH, DRAG_METRIC_COSINE_SIM><<<grid, block, 0, st>>>(a); break;
    case DRAG_METRIC_EUCLIDEAN_DIST: rerank_kernel<T, NCH, DRAG_METRIC_EUCLIDEAN_DIST><<<grid, block, 0, st>>>(a); break;
    case DRAG_METRIC_SQEUCLIDEAN_DIST: rerank_kernel<T, NCH, DRAG_METRIC_SQEUCLIDEAN_DIST><<<grid, block, 0, st>>>(a); break;
    default: rerank_kernel<T, NCH, DRAG_METRIC_INNER_PRODUCT><<<grid, block, 0, st>>>(a); break;
  }
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

template <typename T>
static int launch_rerank(int metric, const RerankArgs& a, int n_queries, cudaStream_t st) {
  const int chunk = 32 * RowVec<T>::EPV;
  const int nch = (a.dim + chunk - 1) / chunk;
  switch (nch) {
    case 1: return launch_rerank_m<T, 1>(metric, a, n_queries, st);
    case 2: return launch_rerank_m<T, 2>(metric, a, n_queries, st);
    case 3: return launch_rerank_m<T, 3>(metric, a, n_queries, st);
    case 4: return launch_rerank_m<T, 4>(metric, a, n_queries, st);
    default: return fail(DRAG_ERR_UNSUPPORTED, "batched top-k: dim %d too large for the re-rank kernel", a.dim);
  }
}

// workspace of the batched path
struct BatchWorkspace {
  double* q_sq;
  double* q_norm;
  __nv_bfloat16* q_bf16;
  float* theta;
  int* cnt;        // [Qpad][MAX_SLOTS]
  int2* meta;      // [Qpad]
  int* surv_cnt;   // [Qpad]
  int* overflow;
  uint2* pool;     // [Qpad][POOL_CAP]
  uint2* surv;     // [Qpad][SURV_CAP]
};

static size_t carve_batch(int n_queries, int k, int dim, void* base, BatchWorkspace* ws) {
  const size_t q_pad = ((size_t)n_queries + tcs::QT - 1) / tcs::QT * tcs::QT;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* ptr = base ? (void*)((unsigned char*)base + off) : nullptr;
    off += align_up(bytes, 1024);
    return ptr;
  };
  double* q_sq = (double*)take(q_pad * 8);
  double* q_norm = (double*)take(q_pad * 8);
  __nv_bfloat16* qb = (__nv_bfloat16*)take(q_pad * dim * 2);
  float* theta = (float*)take(q_pad * 4);
  int* cnt = (int*)take(q_pad * (size_t)tcs::MAX_SLOTS * 4);
  int2* meta = (int2*)take(q_pad * 8);
  int* scnt = (int*)take(q_pad * 4);
  int* ovf = (int*)take(q_pad * 4);
  uint2* pool = (uint2*)take(q_pad * (size_t)tcs::POOL_CAP * 8);
  uint2* surv = (uint2*)take(q_pad * (size_t)tcs::SURV_CAP * 8);
  (void)k;
  if (ws) {
    ws->q_sq = q_sq; ws->q_norm = q_norm; ws->q_bf16 = qb; ws->theta = theta; ws->cnt = cnt; ws->meta = meta;
    ws->surv_cnt = scnt; ws->overflow = ovf; ws->pool = pool; ws->surv = surv;
  }
  return off;
}

template <int MODE>
static int launch_score(const CUtensorMap& tq, const CUtensorMap& tm, const tcs::ScoreParams& p, size_t smem, cudaStream_t st) {
  auto kern = tcs::score_kernel<MODE>;
  DRAG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<p.n_qtiles * p.ctas_per_qtile, tcs::THREADS, smem, st>>>(tq, tm, p);
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

}  // namespace topk
}  // namespace drag

using namespace drag;
using namespace drag::topk;

extern "C" int drag_row_sqnorm(const void* d_matrix, int dtype, int64_t n_rows, int dim, float* d_out, void* stream) {
  DRAG_REQUIRE(n_rows >= 0 && dim > 0, "drag_row_sqnorm: bad shape n_rows=%lld dim=%d", (long long)n_rows, dim);
  DRAG_REQUIRE(dtype == DRAG_F32 || dtype == DRAG_BF16, "drag_row_sqnorm: bad dtype %d", dtype);
  if (n_rows == 0) return DRAG_OK;
  DRAG_REQUIRE(d_matrix && d_out, "drag_row_sqnorm: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 128;
  const unsigned blocks = (unsigned)((n_rows + threads - 1) / threads);
  if (dtype == DRAG_F32)
    row_sqnorm_kernel<float><<<blocks, threads, 0, st>>>((const float*)d_matrix, n_rows, dim, d_out);
  else
    row_sqnorm_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>((const __nv_bfloat16*)d_matrix, n_rows, dim, d_out);
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

extern "C" int drag_distances(int device, const void* d_matrix, int dtype, int64_t n_rows, int dim,
                              const float* d_row_sqnorm, const double* d_query, int metric, double* d_out,
                              void* d_scratch16, void* stream) {
  DRAG_REQUIRE(dtype == DRAG_F32 || dtype == DRAG_BF16, "drag_distances: bad dtype %d", dtype);
  DRAG_REQUIRE(metric >= 0 && metric <= 3, "drag_distances: bad metric %d", metric);
  DRAG_REQUIRE(n_rows >= 0 && dim > 0, "drag_distances: bad shape");
  if (n_rows == 0) return DRAG_OK;
  DRAG_REQUIRE(d_matrix && d_query && d_out && d_scratch16, "drag_distances: null pointer");
  const bool needs_sq = metric == DRAG_METRIC_SQEUCLIDEAN_DIST || metric == DRAG_METRIC_EUCLIDEAN_DIST;
  DRAG_REQUIRE(!needs_sq || d_row_sqnorm, "drag_distances: (sq)euclidean metric needs d_row_sqnorm");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_distances: cannot select device %d", device);
  cudaStream_t st = (cudaStream_t)stream;
  double* q_sq = (double*)d_scratch16;
  double* q_norm = q_sq + 1;
  { int rc = launch_query_prep(d_query, 1, dim, q_sq, q_norm, st); if (rc) return rc; }
  if (dtype == DRAG_F32)
    return launch_distances<float>(metric, (const float*)d_matrix, n_rows, dim, d_row_sqnorm, d_query, q_sq, q_norm, d_out, st);
  return launch_distances<__nv_bfloat16>(metric, (const __nv_bfloat16*)d_matrix, n_rows, dim, d_row_sqnorm, d_query, q_sq, q_norm, d_out, st);
}

extern "C" int drag_topk_workspace_bytes(int device, int n_queries, int k, size_t* bytes) {
  DRAG_REQUIRE(bytes, "drag_topk_workspace_bytes: null out pointer");
  DRAG_REQUIRE(n_queries >= 0 && k >= 1 && k <= MAX_K, "drag_topk_workspace_bytes: need 1 <= k <= %d (k=%d)", MAX_K, k);
  if (sm_count(device) <= 0) return fail(DRAG_ERR_DEVICE, "drag_topk_workspace_bytes: cannot query device %d", device);
  Plan p = make_plan(device, n_queries > 0 ? n_queries : 1, k);
  *bytes = carve(p, n_queries > 0 ? n_queries : 1, nullptr, nullptr);
  return DRAG_OK;
}

extern "C" int drag_topk(int device, const void* d_matrix, int dtype, int64_t n_rows, int dim,
                         const float* d_row_sqnorm, const double* d_queries, int n_queries, int k, int metric,
                         int64_t row_id_base, double* d_out_dist, int64_t* d_out_row, int32_t* d_out_count,
                         void* d_workspace, size_t workspace_bytes, void* stream) {
  DRAG_REQUIRE(dtype == DRAG_F32 || dtype == DRAG_BF16, "drag_topk: bad dtype %d", dtype);
  DRAG_REQUIRE(metric >= 0 && metric <= 3, "drag_topk: bad metric %d", metric);
  DRAG_REQUIRE(k >= 1 && k <= MAX_K, "drag_topk: need 1 <= k <= %d (k=%d)", MAX_K, k);
  DRAG_REQUIRE(n_rows >= 0 && n_rows < 0xffffffffll, "drag_topk: n_rows=%lld out of range", (long long)n_rows);
  DRAG_REQUIRE(dim > 0 && n_queries >= 0, "drag_topk: bad dim/n_queries");
  if (n_queries == 0) return DRAG_OK;
  DRAG_REQUIRE(d_queries && d_out_dist && d_out_row && d_out_count && d_workspace, "drag_topk: null pointer");
  DRAG_REQUIRE(n_rows == 0 || d_matrix, "drag_topk: null matrix");
  const bool needs_sq = metric == DRAG_METRIC_SQEUCLIDEAN_DIST || metric == DRAG_METRIC_EUCLIDEAN_DIST;
  DRAG_REQUIRE(!needs_sq || n_rows == 0 || d_row_sqnorm, "drag_topk: (sq)euclidean metric needs d_row_sqnorm");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_topk: cannot select device %d", device);
  cudaStream_t st = (cudaStream_t)stream;

  Plan p = make_plan(device, n_queries, k);
  Workspace ws;
  size_t need = carve(p, n_queries, d_workspace, &ws);
  DRAG_REQUIRE(workspace_bytes >= need, "drag_topk: workspace too small (%zu < %zu)", workspace_bytes, need);

  { int rc = launch_query_prep(d_queries, n_queries, dim, ws.q_sq, ws.q_norm, st); if (rc) return rc; }

  const size_t esz = dtype == DRAG_F32 ? 4 : 2;
  const bool vec_ok = ((size_t)dim * esz) % 16 == 0 && ((uintptr_t)d_matrix % 16) == 0;
  // small matrices: do not launch more CTAs than there are row tiles
  const long long rows_per_cta = (long long)p.nwarps * 4;
  long long want = (n_rows + rows_per_cta - 1) / rows_per_cta;
  if (want < 1) want = 1;
  if (want < p.grid) p.grid = (int)want;

  for (int q0 = 0; q0 < n_queries; q0 += p.qb) {
    ScanArgs a;
    a.mat = d_matrix; a.n_rows = n_rows; a.dim = dim; a.row_sq = d_row_sqnorm;
    a.queries = d_queries + (size_t)q0 * dim;
    a.q_sq = ws.q_sq + q0; a.q_norm = ws.q_norm + q0;
    a.nq = (n_queries - q0) < p.qb ? (n_queries - q0) : p.qb;
    a.k = k; a.kpad = p.kpad;
    a.part_keys = ws.part_keys; a.part_rows = ws.part_rows; a.row_base = row_id_base;
    int rc = dtype == DRAG_F32 ? launch_scan<float>(metric, a, p, vec_ok, st)
                               : launch_scan<__nv_bfloat16>(metric, a, p, vec_ok, st);
    if (rc != DRAG_OK) return rc;
    MergeArgs m;
    m.in_keys = ws.part_keys; m.in_rows = ws.part_rows; m.in_dist = nullptr; m.in_count = nullptr;
    m.n_lists = p.grid; m.list_len = p.kpad; m.nq = a.nq; m.k = k; m.kpad = p.kpad;
    m.out_dist = d_out_dist + (size_t)q0 * k; m.out_row = (long long*)d_out_row + (size_t)q0 * k;
    m.out_count = d_out_count + q0;
    if ((rc = launch_merge<false>(m, a.nq, st)) != DRAG_OK) return rc;
  }
  return DRAG_OK;
}

extern "C" int drag_topk_merge(int device, const double* d_in_dist, const int64_t* d_in_row, const int32_t* d_in_count,
                               int n_shards, int n_queries, int k, double* d_out_dist, int64_t* d_out_row,
                               int32_t* d_out_count, void* stream) {
  DRAG_REQUIRE(k >= 1 && k <= MAX_K, "drag_topk_merge: need 1 <= k <= %d (k=%d)", MAX_K, k);
  DRAG_REQUIRE(n_shards >= 1 && n_queries >= 0, "drag_topk_merge: bad n_shards/n_queries");
  if (n_queries == 0) return DRAG_OK;
  DRAG_REQUIRE(d_in_dist && d_in_row && d_in_count && d_out_dist && d_out_row && d_out_count, "drag_topk_merge: null pointer");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_topk_merge: cannot select device %d", device);
  cudaStream_t st = (cudaStream_t)stream;
  MergeArgs m;
  m.in_keys = nullptr; m.in_rows = (const long long*)d_in_row; m.in_dist = d_in_dist; m.in_count = d_in_count;
  m.n_lists = n_shards; m.list_len = 0; m.nq = n_queries; m.k = k; m.kpad = next_pow2_min32(k);
  m.out_dist = d_out_dist; m.out_row = (long long*)d_out_row; m.out_count = d_out_count;
  { int rc = launch_merge<true>(m, n_queries, st); if (rc) return rc; }
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

extern "C" int drag_rows_to_chunks(const int64_t* d_rows, int64_t n, const int64_t* d_doc_offsets, int n_docs,
                                   const int64_t* d_row_chunk_ids, int64_t* d_out_doc, int64_t* d_out_chunk, void* stream) {
  DRAG_REQUIRE(n >= 0 && n_docs >= 1, "drag_rows_to_chunks: bad sizes");
  if (n == 0) return DRAG_OK;
  DRAG_REQUIRE(d_rows && d_doc_offsets && d_out_doc && d_out_chunk, "drag_rows_to_chunks: null pointer");
  rows_to_chunks_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      (const long long*)d_rows, n, (const long long*)d_doc_offsets, n_docs, (const long long*)d_row_chunk_ids,
      (long long*)d_out_doc, (long long*)d_out_chunk);
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

// ---------------------------------------------------------------------------------
// batched tensor-core path (drag_topk_tc.cuh)
// ---------------------------------------------------------------------------------
extern "C" int drag_rows_to_bf16(const float* d_in, int64_t n_elems, void* d_out, void* stream) {
  DRAG_REQUIRE(n_elems >= 0 && n_elems % 4 == 0, "drag_rows_to_bf16: element count must be a non-negative multiple of 4");
  if (n_elems == 0) return DRAG_OK;
  DRAG_REQUIRE(d_in && d_out, "drag_rows_to_bf16: null pointer");
  DRAG_REQUIRE(((uintptr_t)d_in % 16) == 0 && ((uintptr_t)d_out % 8) == 0, "drag_rows_to_bf16: misaligned buffers");
  const size_t n4 = (size_t)n_elems / 4;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  tcs::rows_to_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)d_in, n4, (uint2*)d_out);
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

extern "C" int drag_row_norm_stats(const float* d_row_sqnorm, int64_t n_rows, float* d_inv_norm, float* d_stats4, void* stream) {
  DRAG_REQUIRE(n_rows >= 0 && d_stats4, "drag_row_norm_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const float init[4] = {0.f, INFINITY, 0.f, 0.f};
  DRAG_CUDA_OK(cudaMemcpyAsync(d_stats4, init, sizeof(init), cudaMemcpyHostToDevice, st));
  if (n_rows == 0) return DRAG_OK;
  DRAG_REQUIRE(d_row_sqnorm, "drag_row_norm_stats: null pointer");
  long long blocks = (n_rows + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  tcs::row_norm_stats_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_row_sqnorm, n_rows, d_inv_norm, d_stats4);
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

static int batch_supported(int64_t n_rows, int dim, int k, const char** why) {
  if (dim % tcs::BK != 0 || dim > tcs::MAX_DIM) { *why = "dim must be a multiple of 64 and <= 512"; return 0; }
  if (k > tcs::MAX_BATCH_K) { *why = "k must be <= 256"; return 0; }
  if (n_rows >= (1ll << 31)) { *why = "n_rows must be < 2^31"; return 0; }
  return 1;
}

extern "C" int drag_topk_batch_workspace_bytes(int device, int n_queries, int k, int dim, size_t* bytes) {
  DRAG_REQUIRE(bytes, "drag_topk_batch_workspace_bytes: null out pointer");
  DRAG_REQUIRE(n_queries >= 0 && k >= 1 && dim >= 1, "drag_topk_batch_workspace_bytes: bad arguments");
  const char* why = "";
  if (!batch_supported(0, dim, k, &why)) return fail(DRAG_ERR_UNSUPPORTED, "drag_topk_batch: %s", why);
  (void)device;
  *bytes = carve_batch(n_queries > 0 ? n_queries : 1, k, dim, nullptr, nullptr);
  return DRAG_OK;
}

namespace {

// query tiles per launch: the smallest number of groups whose CTA grid fills >= 95% of the SMs
int pick_group(int n_qtiles, int sms) {
  for (int groups = 1; groups <= n_qtiles; ++groups) {
    const int g = (n_qtiles + groups - 1) / groups;
    if (g > sms) continue;
    const int used = (sms / g) * g;
    if (used * 100 >= sms * 95 || g == 1) return g;
  }
  return 1;
}

// collect-all round: every region entry carries (key, row); scatter the keys to out[q][row]
__global__ void debug_keys_kernel(const uint2* pool, const int* cnt, const int2* meta, int n_rows, int n_queries, float* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_queries * tcs::POOL_CAP) return;
  const size_t q = i / tcs::POOL_CAP;
  const int e = (int)(i % tcs::POOL_CAP);
  const int2 m = meta[q];
  const int slot = e / m.y, j = e % m.y;
  if (slot >= m.x || j >= cnt[q * tcs::MAX_SLOTS + slot]) return;
  const uint2 v = pool[i];
  if ((int)v.y < n_rows) out[q * n_rows + v.y] = __uint_as_float(v.x);
}

struct BatchPlan {
  int mode;
  double err_a, err_b;
  const float* colvec;
};

}  // namespace

// shared driver of drag_topk_batch and drag_debug_tc_scores
static int run_batch(int device, const void* d_matrix, int dtype, const void* d_shadow, int64_t n_rows, int dim,
                     const float* d_row_sqnorm, const float* d_row_inv_norm, float max_row_norm, const double* d_queries,
                     int n_queries, int k, int metric, int64_t row_id_base, double* d_out_dist, int64_t* d_out_row,
                     int32_t* d_out_count, int32_t* d_out_status, float* d_debug_keys, void* d_workspace,
                     size_t workspace_bytes, cudaStream_t st) {
  const int sms = sm_count(device);
  if (sms <= 0) return fail(DRAG_ERR_DEVICE, "drag_topk_batch: cannot query device %d", device);
  BatchWorkspace ws;
  const size_t need = carve_batch(n_queries, k, dim, d_workspace, &ws);
  DRAG_REQUIRE(workspace_bytes >= need, "drag_topk_batch: workspace too small (%zu < %zu)", workspace_bytes, need);
  const int q_pad = (n_queries + tcs::QT - 1) / tcs::QT * tcs::QT;
  const int n_qtiles = q_pad / tcs::QT;

  // certified bound E[q] = err_a * |q| + err_b on |approximate key - exact key|:
  //   both operands rounded to bf16 (relative 2^-9 each, Cauchy-Schwarz over the row) -> 2^-8 * 1.01,
  //   fp32 accumulation in the tensor core -> 2^-11 allowance (measured error is ~100x smaller),
  //   fp32 rounding of the key arithmetic (0.5*|d|^2 term / inverse norm)
  const double c1 = ldexp(1.0, -8) * 1.01 + ldexp(1.0, -11);
  const double dmax = (double)max_row_norm * (1.0 + 1e-6);
  BatchPlan plan;
  if (metric == DRAG_METRIC_INNER_PRODUCT) {
    plan.mode = tcs::MODE_IP; plan.err_a = c1 * dmax; plan.err_b = 0.0; plan.colvec = nullptr;
  } else if (metric == DRAG_METRIC_COSINE_SIM) {
    plan.mode = tcs::MODE_COS; plan.err_a = c1 + ldexp(1.0, -19); plan.err_b = 0.0; plan.colvec = d_row_inv_norm;
  } else {
    plan.mode = tcs::MODE_L2; plan.err_a = (c1 + ldexp(1.0, -22)) * dmax; plan.err_b = ldexp(1.0, -22) * dmax * dmax;
    plan.colvec = d_row_sqnorm;
  }

  { int rc = launch_query_prep(d_queries, n_queries, dim, ws.q_sq, ws.q_norm, st); if (rc) return rc; }
  {
    const size_t n = (size_t)q_pad * dim;
    tcs::queries_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_queries, n_queries, q_pad, dim, ws.q_bf16);
    DRAG_CUDA_OK(cudaGetLastError());
  }
  DRAG_CUDA_OK(cudaMemsetAsync(ws.overflow, 0, (size_t)q_pad * 4, st));
  DRAG_CUDA_OK(cudaMemsetAsync(ws.surv_cnt, 0, (size_t)q_pad * 4, st));

  CUtensorMap tq, tm;
  int rc = make_tmap_bf16(&tq, ws.q_bf16, (uint64_t)q_pad, (uint64_t)dim, tcs::QT);
  if (rc) return rc;
  if ((rc = make_tmap_bf16(&tm, d_shadow, (uint64_t)n_rows, (uint64_t)dim, tcs::RT))) return rc;

  const int k_blocks = dim / tcs::BK;
  int stages = tcs::MAX_STAGES;
  while (stages > 2 && tcs::score_smem_bytes(k_blocks, stages) > 227 * 1024) --stages;
  const size_t smem = tcs::score_smem_bytes(k_blocks, stages);
  DRAG_REQUIRE(smem <= 227 * 1024, "drag_topk_batch: dim %d does not fit the shared-memory plan", dim);
  const int group = pick_group(n_qtiles, sms);

  const size_t refine_smem = tcs::refine_smem_bytes();
  DRAG_CUDA_OK(cudaFuncSetAttribute(tcs::refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)refine_smem));

  long long done = 0, bound = tcs::ROUND0_ROWS;
  if (d_debug_keys) bound = n_rows;  // debug: one collect-all round over the whole (small) matrix
  while (done < n_rows) {
    const long long r1 = n_rows < bound ? n_rows : bound;
    const long long n_tiles = (r1 - done + tcs::RT - 1) / tcs::RT;
    for (int qt0 = 0; qt0 < n_qtiles; qt0 += group) {
      tcs::ScoreParams p;
      p.qt0 = qt0;
      p.n_qtiles = (n_qtiles - qt0) < group ? (n_qtiles - qt0) : group;
      long long per = sms / p.n_qtiles;
      if (per > n_tiles) per = n_tiles;
      if (per < 1) per = 1;
      p.ctas_per_qtile = (int)per;
      p.n_valid_q = n_queries;
      p.row0 = done; p.row1 = r1;
      p.k_blocks = k_blocks; p.stages = stages;
      p.colvec = plan.colvec; p.theta = ws.theta; p.pool = ws.pool; p.cnt = ws.cnt; p.meta = ws.meta;
      p.region_cap = tcs::POOL_CAP / p.ctas_per_qtile;
      p.collect_all = done == 0 ? 1 : 0;
      if (plan.mode == tcs::MODE_IP) rc = launch_score<tcs::MODE_IP>(tq, tm, p, smem, st);
      else if (plan.mode == tcs::MODE_L2) rc = launch_score<tcs::MODE_L2>(tq, tm, p, smem, st);
      else rc = launch_score<tcs::MODE_COS>(tq, tm, p, smem, st);
      if (rc) return rc;
    }
    if (d_debug_keys) {
      // keys of round 0 sit at index == row
      const size_t n = (size_t)n_queries * (size_t)tcs::POOL_CAP;
      debug_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws.pool, ws.cnt, ws.meta, (int)n_rows, n_queries, d_debug_keys);
      DRAG_CUDA_OK(cudaGetLastError());
      return DRAG_OK;
    }
    tcs::RefineParams rp;
    rp.pool = ws.pool; rp.cnt = ws.cnt; rp.meta = ws.meta; rp.surv = ws.surv; rp.surv_cnt = ws.surv_cnt;
    rp.theta = ws.theta; rp.overflow = ws.overflow; rp.q_norm = ws.q_norm;
    rp.err_a = plan.err_a; rp.err_b = plan.err_b; rp.k = k;
    tcs::refine_kernel<<<n_queries, tcs::REFINE_THREADS, refine_smem, st>>>(rp);
    DRAG_CUDA_OK(cudaGetLastError());
    done = r1;
    bound *= tcs::ROUND_GROWTH;
  }

  RerankArgs a;
  a.mat = d_matrix; a.n_rows = n_rows; a.dim = dim; a.row_sq = d_row_sqnorm; a.queries = d_queries;
  a.q_sq = ws.q_sq; a.q_norm = ws.q_norm; a.cand = ws.surv; a.cnt = ws.surv_cnt; a.overflow = ws.overflow;
  a.cap = tcs::SURV_CAP; a.k = k; a.row_base = row_id_base;
  a.out_dist = d_out_dist; a.out_row = (long long*)d_out_row; a.out_count = d_out_count; a.out_status = d_out_status;
  return dtype == DRAG_F32 ? launch_rerank<float>(metric, a, n_queries, st)
                           : launch_rerank<__nv_bfloat16>(metric, a, n_queries, st);
}

extern "C" int drag_topk_batch(int device, const void* d_matrix, int dtype, const void* d_shadow_bf16, int64_t n_rows,
                               int dim, const float* d_row_sqnorm, const float* d_row_inv_norm, float max_row_norm,
                               const double* d_queries, int n_queries, int k, int metric, int64_t row_id_base,
                               double* d_out_dist, int64_t* d_out_row, int32_t* d_out_count, int32_t* d_out_status,
                               void* d_workspace, size_t workspace_bytes, void* stream) {
  DRAG_REQUIRE(dtype == DRAG_F32 || dtype == DRAG_BF16, "drag_topk_batch: bad dtype %d", dtype);
  DRAG_REQUIRE(metric >= 0 && metric <= 3, "drag_topk_batch: bad metric %d", metric);
  DRAG_REQUIRE(k >= 1 && n_rows >= 1 && dim >= 1 && n_queries >= 0, "drag_topk_batch: bad sizes");
  const char* why = "";
  if (!batch_supported(n_rows, dim, k, &why)) return fail(DRAG_ERR_UNSUPPORTED, "drag_topk_batch: %s", why);
  if (n_queries == 0) return DRAG_OK;
  DRAG_REQUIRE(d_matrix && d_shadow_bf16 && d_queries && d_out_dist && d_out_row && d_out_count && d_out_status && d_workspace,
               "drag_topk_batch: null pointer");
  DRAG_REQUIRE(((uintptr_t)d_shadow_bf16 % 16) == 0 && ((uintptr_t)d_matrix % 16) == 0, "drag_topk_batch: matrix buffers must be 16-byte aligned");
  DRAG_REQUIRE(max_row_norm >= 0.f && max_row_norm <= 1e18f, "drag_topk_batch: max_row_norm %g is not usable", (double)max_row_norm);
  const bool needs_sq = metric == DRAG_METRIC_SQEUCLIDEAN_DIST || metric == DRAG_METRIC_EUCLIDEAN_DIST;
  DRAG_REQUIRE(!needs_sq || d_row_sqnorm, "drag_topk_batch: (sq)euclidean metric needs d_row_sqnorm");
  DRAG_REQUIRE(metric != DRAG_METRIC_COSINE_SIM || d_row_inv_norm, "drag_topk_batch: cosine metric needs d_row_inv_norm");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_topk_batch: cannot select device %d", device);
  return run_batch(device, d_matrix, dtype, d_shadow_bf16, n_rows, dim, d_row_sqnorm, d_row_inv_norm, max_row_norm,
                   d_queries, n_queries, k, metric, row_id_base, d_out_dist, d_out_row, d_out_count, d_out_status,
                   nullptr, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int drag_debug_tc_keys(int device, const void* d_shadow_bf16, int64_t n_rows, int dim, const float* d_colvec,
                                  int metric, const double* d_queries, int n_queries, float* d_out_keys,
                                  void* d_workspace, size_t workspace_bytes, void* stream) {
  DRAG_REQUIRE(d_shadow_bf16 && d_queries && d_out_keys && d_workspace, "drag_debug_tc_keys: null pointer");
  DRAG_REQUIRE(n_rows >= 1 && n_rows <= 4096 && n_queries >= 1, "drag_debug_tc_keys: 1 <= n_rows <= 4096");
  const char* why = "";
  if (!batch_supported(n_rows, dim, 1, &why)) return fail(DRAG_ERR_UNSUPPORTED, "drag_debug_tc_keys: %s", why);
  DRAG_REQUIRE(metric == DRAG_METRIC_INNER_PRODUCT || d_colvec, "drag_debug_tc_keys: this metric needs d_colvec");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_debug_tc_keys: cannot select device %d", device);
  // one collect-all round over the whole matrix; the keys come back as [n_queries, n_rows]
  const float* sq = (metric == DRAG_METRIC_SQEUCLIDEAN_DIST || metric == DRAG_METRIC_EUCLIDEAN_DIST) ? d_colvec : nullptr;
  const float* inv = metric == DRAG_METRIC_COSINE_SIM ? d_colvec : nullptr;
  return run_batch(device, d_shadow_bf16, DRAG_BF16, d_shadow_bf16, n_rows, dim, sq, inv, 1.0f, d_queries, n_queries, 1,
                   metric, 0, nullptr, nullptr, nullptr, nullptr, d_out_keys, d_workspace, workspace_bytes, (cudaStream_t)stream);
}
