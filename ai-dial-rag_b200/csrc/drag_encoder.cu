// bge-small-en encoder object behind the C ABI.
//
// Replaces HuggingFaceBgeEmbeddings -> SentenceTransformer.encode -> BertModel.forward ->
// Pooling(cls) -> Normalize -> F.normalize (aidial_rag/embeddings/embeddings.py:52-66, :79-96)
// on token ids.  The batch is PACKED (no padding): sequences sit back to back and
// cu_seqlens marks their boundaries, so padded positions cost nothing and the attention
// mask is implicit.
//
// Per forward:  embed (raw sum + row statistics)  ->  12 x { QKV GEMM | attention | out-proj GEMM +
//               residual | FFN-up GEMM + GELU | FFN-down GEMM + residual }  ->  CLS LayerNorm + 2x L2
//               normalise.  LayerNorms are folded into the GEMMs (see drag_gemm.cuh): hidden states
//               are stored raw (pre-LN, bf16) with per-row (sum, sum^2) partials.
// All GEMMs are the tcgen05/TMA kernel of drag_gemm.cuh with fused epilogues.
#include <stdlib.h>
#include <cuda_fp16.h>

#include <mutex>
#include <new>
#include <vector>

#include "drag_attention.cuh"
#include "drag_attention_tc3.cuh"
#include "drag_common.cuh"
#include "drag_gemm.cuh"
#include "drag_mlp.cuh"

namespace drag {
namespace enc {

using bf16 = __nv_bfloat16;

constexpr int HIDDEN = 384;
constexpr int HEAD_DIM = 32;
constexpr int QKV_BLOCK_N = 192;   // 1152 = 6 x 192
constexpr int FFN_BLOCK_N = 256;   // 1536 = 6 x 256
constexpr int FFN_BLOCK_N_PAIR = 256;  // CTA-pair FFN-up (a weights-stationary 192-column form measured slower inside the step: 0.389 vs 0.340 ms)
constexpr int RES_BLOCK_N = 128;   // 384 = 3 x 128 (one statistics slot per tile)
constexpr int RES_BLOCK_N_PAIR = 192;  // CTA-pair kernels: 384 = 2 x 192 (third statistics slot stays zero)
constexpr int PARTS = gemm::STATS_PARTS;

// ---------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x_raw[t] = word[ids[t]] + pos[t - start(seq(t))] + type[0] (bf16) and the row's (sum, sum^2);
// the embedding LayerNorm itself is folded into the first QKV GEMM / residual.  One warp per token,
// each lane owns 12 of the 384 features (3 x float4, coalesced).
__global__ void __launch_bounds__(256)
embed_kernel(const int* __restrict__ ids, const int* __restrict__ cu_seqlens, int n_seq, int total,
             const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type0,
             int vocab, bf16* __restrict__ out, float2* __restrict__ stats) {
  const int tok = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= total) return;
  int lo = 0, hi = n_seq;  // last s with cu[s] <= tok
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cu_seqlens[mid] <= tok) lo = mid; else hi = mid;
  }
  const int p = tok - cu_seqlens[lo];
  int id = ids[tok];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int e = c * 128 + lane * 4;
    float4 w = __ldg(reinterpret_cast<const float4*>(word + (size_t)id * HIDDEN + e));
    float4 pp = __ldg(reinterpret_cast<const float4*>(pos + (size_t)p * HIDDEN + e));
    float4 tt = __ldg(reinterpret_cast<const float4*>(type0 + e));
    const float v0 = w.x + pp.x + tt.x, v1 = w.y + pp.y + tt.y, v2 = w.z + pp.z + tt.z, v3 = w.w + pp.w + tt.w;
    s += (v0 + v1) + (v2 + v3);
    ss = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, ss))));
    *reinterpret_cast<uint2*>(out + (size_t)tok * HIDDEN + e) = make_uint2(gemm::pack_bf16(v0, v1), gemm::pack_bf16(v2, v3));
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (lane < PARTS) stats[(size_t)tok * PARTS + lane] = lane == 0 ? make_float2(s, ss) : make_float2(0.f, 0.f);
}

__device__ __forceinline__ void load_row12(const bf16* row, int lane, float (&v)[12]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint2 raw = *reinterpret_cast<const uint2*>(row + c * 128 + lane * 4);
    v[c * 4 + 0] = __uint_as_float(raw.x << 16); v[c * 4 + 1] = __uint_as_float(raw.x & 0xffff0000u);
    v[c * 4 + 2] = __uint_as_float(raw.y << 16); v[c * 4 + 3] = __uint_as_float(raw.y & 0xffff0000u);
  }
}

__device__ __forceinline__ void apply_ln12(float (&v)[12], int lane, const float2* stats, size_t row,
                                           const float* gamma, const float* beta, float eps) {
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < PARTS; ++i) { float2 p = stats[row * PARTS + i]; s += p.x; ss += p.y; }
  const float mu = s * (1.f / HIDDEN);
  const float rstd = rsqrtf(fmaxf(ss * (1.f / HIDDEN) - mu * mu, 0.f) + eps);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c * 128 + lane * 4));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c * 128 + lane * 4));
    v[c * 4 + 0] = (v[c * 4 + 0] - mu) * rstd * g.x + b.x; v[c * 4 + 1] = (v[c * 4 + 1] - mu) * rstd * g.y + b.y;
    v[c * 4 + 2] = (v[c * 4 + 2] - mu) * rstd * g.z + b.z; v[c * 4 + 3] = (v[c * 4 + 3] - mu) * rstd * g.w + b.w;
  }
}

// out[s] = normalize(normalize(LN(x_raw[cu[s]])))  -- final LayerNorm of the CLS row, CLS pooling and
// the two F.normalize(p=2, eps=1e-12)
// (cu_seqlens == nullptr: the rows are already compact, row = sequence -- the CLS-only last layer)
__global__ void __launch_bounds__(256)
pool_normalize_kernel(const bf16* __restrict__ x, const float2* __restrict__ stats, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, const int* __restrict__ cu_seqlens, int n_seq,
                      float* __restrict__ out) {
  const int seq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  tc::pdl_wait();   // programmatic dependent launch: the last layer's outputs are visible from here on
  if (seq >= n_seq) return;
  const size_t row = cu_seqlens ? (size_t)cu_seqlens[seq] : (size_t)seq;
  float v[12];
  load_row12(x + row * HIDDEN, lane, v);
  apply_ln12(v, lane, stats, row, gamma, beta, eps);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) ss = fmaf(v[i], v[i], ss);
  float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  ss = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) { v[i] *= inv; ss = fmaf(v[i], v[i], ss); }
  inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    *reinterpret_cast<float4*>(out + (size_t)seq * HIDDEN + c * 128 + lane * 4) =
        make_float4(v[c * 4] * inv, v[c * 4 + 1] * inv, v[c * 4 + 2] * inv, v[c * 4 + 3] * inv);
}

// Last layer, CLS-only: only the [CLS] row of every sequence is pooled, so after the last QKV
// projection nothing but those rows is needed.  One warp per (sequence, head): softmax(q_cls K^T / sqrt(32)) V for the
// CLS query alone (fp32 arithmetic on the bf16 Q/K/V, exact single-pass softmax) -> ctx_c[seq]; the block also gathers
// the raw CLS row of the residual stream and its statistics into compact rows (row = sequence), which the M = n_seq
// out-projection / FFN GEMMs of the last layer then work on.  Block = heads warps: together they read whole 768-byte
// K and V rows of the sequence's tokens.
__global__ void __launch_bounds__(384)
cls_attention_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ x, const float2* __restrict__ stats_x,
                     const int* __restrict__ cu_seqlens, int heads, float scale_log2, bf16* __restrict__ ctx_c,
                     bf16* __restrict__ res_c, float2* __restrict__ stats_c) {
  const int seq = blockIdx.x;
  const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hidden = heads * HEAD_DIM, ld = 3 * hidden;
  tc::pdl_launch_dependents();
  tc::pdl_wait();   // programmatic dependent launch: the QKV GEMM's outputs are visible from here on
  const int tok0 = cu_seqlens[seq];
  const int S = cu_seqlens[seq + 1] - tok0;
  // gather: raw CLS row + statistics
  res_c[(size_t)seq * hidden + head * HEAD_DIM + lane] = x[(size_t)tok0 * hidden + head * HEAD_DIM + lane];
  if (head == 0 && lane < PARTS) stats_c[(size_t)seq * PARTS + lane] = stats_x[(size_t)tok0 * PARTS + lane];
  auto unpack8 = [](const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  };
  float q[HEAD_DIM];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(qkv + (size_t)tok0 * ld + head * HEAD_DIM);
#pragma unroll
    for (int c = 0; c < 4; ++c) { float f[8]; unpack8(__ldg(qp + c), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[c * 8 + i] = f[i] * scale_log2; }
  }
  constexpr int MAX_PER_LANE = 16;   // max_pos 512 / 32 lanes
  float sc[MAX_PER_LANE];
  float m = -INFINITY;
#pragma unroll
  for (int t = 0; t < MAX_PER_LANE; ++t) {
    const int key = t * 32 + lane;
    sc[t] = -INFINITY;
    if (key < S) {
      const uint4* kp = reinterpret_cast<const uint4*>(qkv + (size_t)(tok0 + key) * ld + hidden + head * HEAD_DIM);
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) { float f[8]; unpack8(__ldg(kp + c), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) dot = fmaf(q[c * 8 + i], f[i], dot); }
      sc[t] = dot;
      m = fmaxf(m, dot);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float l = 0.f, acc[HEAD_DIM];
#pragma unroll
  for (int i = 0; i < HEAD_DIM; ++i) acc[i] = 0.f;
#pragma unroll
  for (int t = 0; t < MAX_PER_LANE; ++t) {
    const int key = t * 32 + lane;
    if (key < S) {
      const float p = exp2f(sc[t] - m);
      l += p;
      const uint4* vp = reinterpret_cast<const uint4*>(qkv + (size_t)(tok0 + key) * ld + 2 * hidden + head * HEAD_DIM);
#pragma unroll
      for (int c = 0; c < 4; ++c) { float f[8]; unpack8(__ldg(vp + c), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c * 8 + i] = fmaf(p, f[i], acc[c * 8 + i]); }
    }
  }
  l = warp_sum(l);
  float mine = 0.f;
#pragma unroll
  for (int i = 0; i < HEAD_DIM; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == i) mine = v;
  }
  ctx_c[(size_t)seq * hidden + head * HEAD_DIM + lane] = __float2bfloat16_rn(mine / l);
}

// debug tap: hidden[t] = LN(x_raw[t]) as fp32 (what HF's BertModel exposes as a layer output)
__global__ void __launch_bounds__(256)
ln_apply_kernel(const bf16* __restrict__ x, const float2* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int total, float* __restrict__ out) {
  const int tok = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= total) return;
  float v[12];
  load_row12(x + (size_t)tok * HIDDEN, lane, v);
  apply_ln12(v, lane, stats, (size_t)tok, gamma, beta, eps);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    *reinterpret_cast<float4*>(out + (size_t)tok * HIDDEN + c * 128 + lane * 4) =
        make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
}

__global__ void bf16_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

// ---------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------
static uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

static int make_tmap(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  static_assert(gemm::BLOCK_K == 64, "TMA boxes are 64 bf16 columns (one 128-byte swizzle row)");
  return make_tmap_bf16(map, base, rows, cols, box_rows);
}

struct Layer {
  bf16 *w_qkv, *w_o, *w_up, *w_down;   // [1152,384] (gamma_in folded) [384,384] [1536,384] (gamma_1 folded) [384,1536]
  float *qkv_c, *qkv_d;                // folded LN_in terms of the QKV projection   [1152]
  float *up_c, *up_d;                  // folded LN_1 terms of the FFN up projection [1536]
  float *o_cold, *o_gamma;             // out-proj residual: b_o + beta_in, gamma_in [384]
  float *down_cold, *down_gamma;       // FFN-down residual: b_down + beta_1, gamma_1 [384]
  float *ln2_g, *ln2_b;                // this layer's output LayerNorm (used by the taps / the pooler)
  CUtensorMap tm_qkv, tm_o, tm_up, tm_down;
  CUtensorMap tp_qkv, tp_o, tp_up, tp_down;   // CTA-pair kernels: each CTA loads half of a W tile
  CUtensorMap tf_w1, tf_w2;                   // fused feed-forward kernel: a CTA's half of a 64-unit chunk of W1 (32 rows) / W2 (96 rows)
};

}  // namespace enc
}  // namespace drag

using namespace drag;
using namespace drag::enc;

// Activation workspace + host-buffer staging of one forward in flight.  An encoder owns two: the BULK workspace
// (max_tokens: indexing batches) and a small QUERY workspace with its own high-priority stream, so that a query
// embedding never queues behind an indexing forward -- neither on the host lock nor in the stream (the reference keeps
// two thread pools for exactly that, aidial_rag/resources/cpu_pools.py:50-59).
struct Workspace {
  int64_t max_tokens = 0;  // padded to a multiple of 128
  // activations (bf16): x / y are RAW (pre-LayerNorm) hidden states with their row statistics
  bf16 *x = nullptr, *y = nullptr, *qkv = nullptr, *ctx = nullptr, *h = nullptr;
  float2 *stats_x = nullptr, *stats_y = nullptr;  // [T][PARTS] partial (sum, sum^2)
  CUtensorMap tm_x, tm_y, tm_ctx, tm_h;           // A-operand loads (128 x 64 boxes)
  CUtensorMap ts_x, ts_y, ts_qkv, ts_h;           // epilogue stores (32 x 64 boxes)
  CUtensorMap tm_qkv_heads;                       // attention loads: 128 tokens x one head (32 columns)
  CUtensorMap t32_x, t32_y;                       // fused feed-forward kernel stores (32 rows x 16 columns, 32-byte swizzle)
  // host-buffer path
  cudaStream_t stream = nullptr;
  int32_t *d_ids = nullptr, *d_cu = nullptr;
  float* d_out = nullptr;
  int32_t *p_ids = nullptr, *p_cu = nullptr;
  float* p_out = nullptr;
  int64_t out_cap = 0;
  std::mutex lock;            // host threads enqueue one forward at a time per workspace
  cudaEvent_t done = nullptr; // recorded behind every forward: the next one (on whatever stream) waits for it
  bool used = false;
};

constexpr int64_t QUERY_WORKSPACE_TOKENS = 8192;   // 16 queries of 512 tokens

struct drag_encoder {
  drag_bert_shape shape;
  int device = 0;
  int sms = 0;
  int64_t max_tokens = 0;  // padded to a multiple of 128
  std::vector<void*> allocs;
  float *word = nullptr, *pos = nullptr, *type0 = nullptr, *emb_g = nullptr, *emb_b = nullptr;
  std::vector<Layer> layers;
  Workspace ws[2];         // 0 = bulk, 1 = query
  // cta_group::2 (CTA-pair) GEMMs: bit 0 QKV, 1 out-proj, 2 FFN-up, 3 FFN-down; bit 4: weights-stationary FFN-up.  Measured on B200 (same run,
  // 262144 tokens): QKV 0.255 vs 0.267 ms, FFN-up 0.365 vs 0.375, FFN-down 0.339 vs 0.397 in favour of pairs;
  // the out-projection (N = K = 384, epilogue-bound) 0.178 vs 0.202 in favour of single CTAs.  DRAG_GEMM_PAIRS=<mask>.
  int gemm_pairs = 15;
  int attention_variant = -1;                     // -1 = by sequence length (launch_attention), 0 = mma.sync kernel, 3 = tcgen05 kernel (DRAG_ATTENTION=mma / tc3)
  // FFN-up + GELU + FFN-down + residual in one kernel (drag_mlp.cuh) for batches of at least FUSED_MLP_MIN_TOKENS tokens;
  // DRAG_FUSED_MLP=0: always the two GEMM kernels.  Measured (262 144 tokens): 0.52 ms against 0.33 + 0.33 ms.
  bool fused_mlp = true;
  // programmatic dependent launch of the forward's kernels: each one's set-up (barriers, tensor-memory allocation, descriptor
  // prefetch) overlaps its predecessor's tail (DRAG_PDL=0: plain stream order)
  bool pdl = true;
  bool cls_only = true;                           // last layer on the [CLS] rows only (DRAG_CLS_ONLY=0: all rows)
  // optional per-kernel-class timing (bench.py roofline): event pairs recorded around the launches of the bulk workspace
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;  // start/stop pairs
  std::vector<int> prof_class;
  size_t prof_used = 0;
};

constexpr int PDL_MAX_TOKENS = 16384;

enum KernelClass { KC_EMBED = 0, KC_GEMM_QKV, KC_ATTENTION, KC_GEMM_OUT_LN, KC_GEMM_UP_GELU, KC_GEMM_DOWN_LN, KC_POOL, KC_CLS_TAIL, KC_MLP, KC_COUNT };

// A pair of CTAs of the fused feed-forward kernel owns 256 tokens for ~20 us: small batches (the query path) spread
// better over the SMs as two GEMMs tiled over tokens AND columns.  Measured: 8 192 tokens 37 us fused against 31 us,
// 16 384 tokens 38 us against 52 us.
constexpr int FUSED_MLP_MIN_TOKENS = 12288;

struct ProfScope {
  drag_encoder* e;
  cudaStream_t st;
  cudaEvent_t stop = nullptr;
  ProfScope(drag_encoder* enc, const Workspace& ws, int klass, cudaStream_t s) : e(enc), st(s) {
    if (!e->profiling || &ws != &e->ws[0] || e->prof_used + 2 > e->prof_events.size()) return;
    cudaEventRecord(e->prof_events[e->prof_used], st);
    stop = e->prof_events[e->prof_used + 1];
    e->prof_class.push_back(klass);
    e->prof_used += 2;
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, st);
  }
};

namespace {

template <typename T>
int dev_alloc(drag_encoder* e, T** out, size_t count) {
  void* p = nullptr;
  cudaError_t err = cudaMalloc(&p, count * sizeof(T) + 256);
  if (err != cudaSuccess) {
    cudaGetLastError();
    return fail(DRAG_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(err));
  }
  e->allocs.push_back(p);
  *out = (T*)p;
  return DRAG_OK;
}

int upload_f32(drag_encoder* e, float** dst, const float* src, size_t n) {
  int rc = dev_alloc(e, dst, n);
  if (rc) return rc;
  DRAG_CUDA_OK(cudaMemcpy(*dst, src, n * 4, cudaMemcpyHostToDevice));
  return DRAG_OK;
}

// rows of several [r_i, cols] fp32 host matrices stacked into one bf16 device matrix
int upload_bf16(drag_encoder* e, bf16** dst, std::initializer_list<const float*> srcs, size_t rows_each, size_t cols) {
  const size_t n_each = rows_each * cols;
  std::vector<uint16_t> tmp(n_each * srcs.size());
  size_t off = 0;
  for (const float* s : srcs) {
    for (size_t i = 0; i < n_each; ++i) tmp[off + i] = f32_to_bf16_rne(s[i]);
    off += n_each;
  }
  int rc = dev_alloc(e, dst, tmp.size());
  if (rc) return rc;
  DRAG_CUDA_OK(cudaMemcpy(*dst, tmp.data(), tmp.size() * 2, cudaMemcpyHostToDevice));
  return DRAG_OK;
}

// [rows, cols] fp32 host matrix as an fp16 device matrix (the FFN-down weights: their A operand is the fp16 GELU output)
int upload_f16(drag_encoder* e, bf16** dst, const float* src, size_t rows, size_t cols) {
  std::vector<__half> tmp(rows * cols);
  for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = __float2half_rn(src[i]);
  int rc = dev_alloc(e, dst, tmp.size());
  if (rc) return rc;
  DRAG_CUDA_OK(cudaMemcpy(*dst, tmp.data(), tmp.size() * 2, cudaMemcpyHostToDevice));
  return DRAG_OK;
}

int upload_concat_f32(drag_encoder* e, float** dst, std::initializer_list<const float*> srcs, size_t n_each) {
  std::vector<float> tmp;
  for (const float* s : srcs) tmp.insert(tmp.end(), s, s + n_each);
  return upload_f32(e, dst, tmp.data(), tmp.size());
}

// One launch path for every kernel of the forward: optional CTA-pair clusters, optional programmatic dependent launch
// (the kernel must call tc::pdl_wait() before it touches anything an earlier kernel of the stream reads or writes)
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <int BLOCK_N, int EPI, int EPI_WARPS, int STAGES, int CG, bool WS = false>
int launch_gemm(const drag_encoder* e, const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tout,
                const CUtensorMap& tres, const gemm::GemmParams& p, cudaStream_t st, bool pdl = false) {
  // tres: the residual rows as 32 x 64 boxes (EPI_RES; the other epilogues ignore it)
  auto kern = gemm::gemm_kernel<BLOCK_N, EPI, EPI_WARPS, STAGES, CG, WS>;
  constexpr size_t smem = gemm::smem_bytes<BLOCK_N, STAGES, EPI_WARPS, CG, WS, EPI>();
  if (WS) {
    DRAG_REQUIRE(p.K <= gemm::WS_K_BLOCKS * gemm::BLOCK_K, "weights-stationary GEMM holds at most K=%d (got %d)", gemm::WS_K_BLOCKS * gemm::BLOCK_K, p.K);
    DRAG_REQUIRE(e->sms / CG >= p.N / BLOCK_N, "weights-stationary GEMM needs at least one worker per column block");
  }
  static_assert(smem <= 227 * 1024, "GEMM configuration exceeds the 227 KB shared memory of an SM");
  static std::once_flag once[16];
  static cudaError_t attr_err[16];
  const int dev_slot = e->device & 15;
  std::call_once(once[dev_slot], [&] {
    attr_err[dev_slot] = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (attr_err[dev_slot] != cudaSuccess)
    return fail(DRAG_ERR_CUDA, "cudaFuncSetAttribute(gemm smem=%zu) failed: %s", smem, cudaGetErrorString(attr_err[dev_slot]));
  const int m_tiles = (p.M + gemm::BLOCK_M * CG - 1) / (gemm::BLOCK_M * CG);
  const int tiles = m_tiles * (p.N / BLOCK_N);
  const int workers = e->sms / CG;   // CTAs, or CTA pairs (one pair per TPC)
  const int grid = CG * (tiles < workers ? tiles : workers);
  DRAG_CUDA_OK(launch_ex(kern, dim3(grid), dim3(64 + 32 * EPI_WARPS), smem, st, CG, pdl, ta, tw, tout, tres, p));
  return DRAG_OK;
}

// the fused feed-forward block (drag_mlp.cuh): out = gelu(LN_1(x) W1^T + b_1) W2^T + b_2 + LN_1(x), CTA pairs over 256-token tiles
int launch_mlp(const drag_encoder* e, const CUtensorMap& tx, const CUtensorMap& tw1, const CUtensorMap& tw2, const CUtensorMap& tout,
               const mlp::MlpParams& p, cudaStream_t st, bool pdl = false) {
  constexpr size_t smem = mlp::smem_bytes();
  static std::once_flag once[16];
  static cudaError_t attr_err[16];
  const int dev_slot = e->device & 15;
  std::call_once(once[dev_slot], [&] {
    attr_err[dev_slot] = cudaFuncSetAttribute(mlp::mlp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (attr_err[dev_slot] == cudaSuccess) attr_err[dev_slot] = cudaFuncSetAttribute(mlp::mlp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (attr_err[dev_slot] != cudaSuccess)
    return fail(DRAG_ERR_CUDA, "cudaFuncSetAttribute(mlp smem=%zu) failed: %s", smem, cudaGetErrorString(attr_err[dev_slot]));
  const int tiles = (p.M + 2 * mlp::ROWS - 1) / (2 * mlp::ROWS);
  const int pairs = e->sms / 2;
  const dim3 grid(2 * (tiles < pairs ? tiles : pairs));
  if (p.trace) DRAG_CUDA_OK(launch_ex(mlp::mlp_kernel<true>, grid, dim3(mlp::THREADS), smem, st, 2, false, tx, tw1, tw2, tout, p));
  else DRAG_CUDA_OK(launch_ex(mlp::mlp_kernel<false>, grid, dim3(mlp::THREADS), smem, st, 2, pdl, tx, tw1, tw2, tout, p));
  return DRAG_OK;
}

// Sequences longer than this go to the tcgen05 kernel, the others to the mma.sync kernel.  Measured on B200, 12 heads
// (profiles/r02_attention_*): 512-token sequences 0.71 ms per 512 x 512 on tcgen05 against 0.83 ms on mma.sync; 256 tokens
// 0.40-0.43 against 0.39; 128 tokens 0.29-0.32 against 0.28 -- a softmax warp gets a MUFU.EX2 issued only every ~18
// clocks, and the many small warps of the mma.sync kernel hide that better on short rows.
constexpr int ATTENTION_TC_ABOVE = 256;

// attention of the (sequence, head) items of the packed batch whose length is in (len_lo, len_hi]; max_len = the longest
// of those.  variant 0: mma.sync kernel, 3: tcgen05 kernel (7: with the debug timeline)
int launch_attention(int variant, const CUtensorMap& tm_qkv_heads, const bf16* qkv, bf16* ctx, const int32_t* d_cu,
                     int n_seq, int max_len, int heads, cudaStream_t st, int len_lo = 0, int len_hi = 1 << 30, bool pdl = false) {
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)HEAD_DIM);
  if (variant == 3 || variant == 7) {
    // tcgen05, two softmax groups, P in tensor memory (drag_attention_tc3.cuh); variant 7 = the same with the debug timeline
    int sms = 148;
    {
      int dev = 0;
      if (cudaGetDevice(&dev) == cudaSuccess && sm_count(dev) > 0) sms = sm_count(dev);
    }
    const int items = heads * n_seq;
    const int split = attn3::pick_split(items, max_len, sms);
    const int units = items * split;
    const size_t smem = attn3::smem_bytes(max_len);
    const int max_tiles = (max_len + attn3::TILE - 1) / attn3::TILE;
    const int grid = units < sms ? units : sms;
    const int stages = attn3::unit_stages(max_len);
    if (variant == 7)
      DRAG_CUDA_OK(launch_ex(attn3::attention_tc3_kernel<true>, dim3(grid), dim3(attn3::THREADS), smem, st, 1, false, tm_qkv_heads, ctx, d_cu, n_seq, heads, max_tiles, stages, split, scale_log2, len_lo, len_hi));
    else
      DRAG_CUDA_OK(launch_ex(attn3::attention_tc3_kernel<false>, dim3(grid), dim3(attn3::THREADS), smem, st, 1, pdl, tm_qkv_heads, ctx, d_cu, n_seq, heads, max_tiles, stages, split, scale_log2, len_lo, len_hi));
  } else {
    const size_t smem = attn::smem_bytes(max_len);
    // few sequences (the query path): smaller query tiles so that the launch still fills the GPU
    int rows_per_cta = attn::ROWS_PER_CTA;
    while (rows_per_cta > 64 && (long long)heads * n_seq * ((max_len + rows_per_cta - 1) / rows_per_cta) < 148) rows_per_cta >>= 1;
    const int q_tiles = (max_len + rows_per_cta - 1) / rows_per_cta;
    for (int s0 = 0; s0 < n_seq; s0 += 65535) {   // gridDim.y limit
      const int n = n_seq - s0 < 65535 ? n_seq - s0 : 65535;
      DRAG_CUDA_OK(launch_ex(attn::attention_kernel, dim3(heads * q_tiles, n), dim3(attn::WARPS * 32), smem, st, 1, pdl, qkv, ctx, d_cu + s0, heads, scale_log2, rows_per_cta, len_lo, len_hi));
    }
  }
  DRAG_CUDA_OK(cudaGetLastError());
  return DRAG_OK;
}

int attention_set_attributes() {
  // the shared-memory need is not monotonic in the sequence length (fewer stages for longer items): take the maximum
  size_t tc3_max = 0;
  for (int len = 128; len <= 512; len += 128) tc3_max = attn3::smem_bytes(len) > tc3_max ? attn3::smem_bytes(len) : tc3_max;
  DRAG_CUDA_OK(cudaFuncSetAttribute(attn3::attention_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc3_max));
  DRAG_CUDA_OK(cudaFuncSetAttribute(attn3::attention_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc3_max));
  DRAG_CUDA_OK(cudaFuncSetAttribute(attn::attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn::smem_bytes(512)));
  return DRAG_OK;
}

// the four GEMM configurations of a layer
#define DRAG_GEMM_QKV  launch_gemm<QKV_BLOCK_N, gemm::EPI_LNIN, 4, 4, 1>
#define DRAG_GEMM_UP   launch_gemm<FFN_BLOCK_N, gemm::EPI_LNIN_GELU, 8, 3, 1>
#define DRAG_GEMM_RES  launch_gemm<RES_BLOCK_N, gemm::EPI_RES, 8, 4, 1>
// ... and their CTA-pair (cta_group::2) forms: 256 x BLOCK_N tiles, half a W tile per CTA
// (the K = 384 ones keep their W column block resident in shared memory: "weights stationary")
#define DRAG_GEMM2_QKV launch_gemm<QKV_BLOCK_N, gemm::EPI_LNIN, 4, 6, 2, true>
#define DRAG_GEMM2_UP  launch_gemm<FFN_BLOCK_N_PAIR, gemm::EPI_LNIN_GELU, 8, 4, 2>
#define DRAG_GEMM2_UP_WS launch_gemm<FFN_BLOCK_N_PAIR, gemm::EPI_LNIN_GELU, 8, 3, 2, true>
#define DRAG_GEMM2_RES launch_gemm<RES_BLOCK_N_PAIR, gemm::EPI_RES, 12, 4, 2>
#define DRAG_GEMM2_RES_WS launch_gemm<RES_BLOCK_N_PAIR, gemm::EPI_RES, 12, 3, 2, true>

int forward_impl(drag_encoder* e, Workspace& ws, const int32_t* d_ids, const int32_t* d_cu, const int32_t* h_cu, int n_seq,
                 float* d_out, int stop_after_layer, float* d_hidden, cudaStream_t st) {
  DRAG_REQUIRE(e && d_ids && d_cu && h_cu, "drag_encoder_forward: null pointer");
  DRAG_REQUIRE(n_seq >= 0, "drag_encoder_forward: n_seq < 0");
  if (n_seq == 0) return DRAG_OK;
  DRAG_REQUIRE(h_cu[0] == 0, "drag_encoder_forward: cu_seqlens[0] must be 0");
  int max_len = 0, max_short = 0, n_long = 0;   // short / long: the two attention kernels' length classes
  for (int i = 0; i < n_seq; ++i) {
    const int len = h_cu[i + 1] - h_cu[i];
    DRAG_REQUIRE(len >= 1 && len <= e->shape.max_pos, "drag_encoder_forward: sequence %d has length %d (allowed 1..%d)", i, len, e->shape.max_pos);
    if (len > max_len) max_len = len;
    if (len > ATTENTION_TC_ABOVE) ++n_long;
    else if (len > max_short) max_short = len;
  }
  const int total = h_cu[n_seq];
  DRAG_REQUIRE((int64_t)total <= ws.max_tokens, "drag_encoder_forward: %d tokens exceed max_tokens=%lld", total, (long long)ws.max_tokens);
  const drag_bert_shape& sh = e->shape;

  const bool tap = d_hidden != nullptr;
  // Programmatic dependent launch pays where kernels are short (the query path: 62 launches of ~10 us, batch-1 forward
  // 0.69 -> 0.55 ms); on indexing batches, where every kernel fills the GPU for 0.1-0.5 ms, it measured 1 % slower.
  const bool pdl = e->pdl && total <= PDL_MAX_TOKENS;
  {
    ProfScope prof(e, ws, KC_EMBED, st);
    const int warps_per_block = 8;
    const int blocks = (total + warps_per_block - 1) / warps_per_block;
    embed_kernel<<<blocks, warps_per_block * 32, 0, st>>>(d_ids, d_cu, n_seq, total, e->word, e->pos, e->type0, sh.vocab,
                                                         ws.x, ws.stats_x);
    DRAG_CUDA_OK(cudaGetLastError());
  }
  const int n_layers = (tap && stop_after_layer < sh.layers) ? stop_after_layer : sh.layers;
  // the pooled output needs the [CLS] rows of the last layer only (the debug tap wants every row)
  const bool cls_tail = e->cls_only && !tap && d_out && n_layers == sh.layers && n_layers >= 1;
  for (int l = 0; l < n_layers; ++l) {
    const Layer& L = e->layers[l];
    const bool last_cls = cls_tail && l == n_layers - 1;
    gemm::GemmParams p{};
    p.M = total;
    p.ln_eps = sh.ln_eps;
    p.inv_width = 1.0f / HIDDEN;
    int rc;
    // QKV projection of LN_in(x): qkv = rstd*(x_raw . (gamma (.) Wqkv)^T - mu*c) + d
    p.N = 3 * HIDDEN; p.K = HIDDEN; p.colc = L.qkv_c; p.cold = L.qkv_d; p.in_stats = ws.stats_x;
    {
      ProfScope prof(e, ws, KC_GEMM_QKV, st);
      rc = (e->gemm_pairs & 1) ? DRAG_GEMM2_QKV(e, ws.tm_x, L.tp_qkv, ws.ts_qkv, ws.ts_qkv, p, st, pdl) : DRAG_GEMM_QKV(e, ws.tm_x, L.tm_qkv, ws.ts_qkv, ws.ts_qkv, p, st, pdl);
    }
    if (rc) return rc;
    // From here on the last layer works on COMPACT rows (row = sequence) when only the pooled output is wanted:
    // src / dst swap roles (the full-size buffers of the previous layer are dead).
    bf16 *res = ws.x, *mid = ws.y;
    float2 *res_stats = ws.stats_x, *mid_stats = ws.stats_y;
    const CUtensorMap *tm_res = &ws.tm_x, *ts_res = &ws.ts_x, *tm_mid = &ws.tm_y, *ts_mid = &ws.ts_y;
    if (last_cls) {
      ProfScope prof(e, ws, KC_CLS_TAIL, st);
      const float scale_log2 = 1.4426950408889634f / sqrtf((float)HEAD_DIM);
      DRAG_CUDA_OK(launch_ex(cls_attention_kernel, dim3(n_seq), dim3(sh.heads * 32), 0, st, 1, pdl, (const bf16*)ws.qkv, (const bf16*)ws.x, (const float2*)ws.stats_x,
                             d_cu, sh.heads, scale_log2, ws.ctx, ws.y, ws.stats_y));
      res = ws.y; mid = ws.x; res_stats = ws.stats_y; mid_stats = ws.stats_x;
      tm_res = &ws.tm_y; ts_res = &ws.ts_y; tm_mid = &ws.tm_x; ts_mid = &ws.ts_x;
      p.M = n_seq;
    } else {
      ProfScope prof(e, ws, KC_ATTENTION, st);
      if (e->attention_variant >= 0) {
        rc = launch_attention(e->attention_variant, ws.tm_qkv_heads, ws.qkv, ws.ctx, d_cu, n_seq, max_len, sh.heads, st, 0, 1 << 30, pdl);
      } else {
        // every sequence goes to the kernel of its length class whatever else is in the batch, so its embedding does not
        // depend on the batch composition; a mixed batch takes two launches
        rc = DRAG_OK;
        if (n_long < n_seq) rc = launch_attention(0, ws.tm_qkv_heads, ws.qkv, ws.ctx, d_cu, n_seq, max_short, sh.heads, st, 0, ATTENTION_TC_ABOVE, pdl);
        if (rc == DRAG_OK && n_long > 0)
          rc = launch_attention(3, ws.tm_qkv_heads, ws.qkv, ws.ctx, d_cu, n_seq, max_len, sh.heads, st, ATTENTION_TC_ABOVE, 1 << 30, pdl);
      }
      if (rc) return rc;
    }
    // mid_raw = ctx . Wo^T + b_o + LN_in(res_raw)   (+ row statistics of mid_raw)
    p.N = HIDDEN; p.K = HIDDEN; p.colc = nullptr; p.cold = L.o_cold; p.gamma = L.o_gamma; p.in_stats = res_stats;
    p.residual = res; p.out_stats = mid_stats;
    {
      ProfScope prof(e, ws, last_cls ? KC_CLS_TAIL : KC_GEMM_OUT_LN, st);
      rc = (e->gemm_pairs & 2) ? DRAG_GEMM2_RES_WS(e, ws.tm_ctx, L.tp_o, *ts_mid, *ts_res, p, st, pdl) : DRAG_GEMM_RES(e, ws.tm_ctx, L.tm_o, *ts_mid, *ts_res, p, st, pdl);
    }
    if (rc) return rc;
    if (e->fused_mlp && !last_cls && total >= FUSED_MLP_MIN_TOKENS) {
      // res_raw = gelu(LN_1(mid) . W1^T + b_1) . W2^T + b_2 + LN_1(mid_raw)  (+ row statistics) in one kernel
      mlp::MlpParams mp{};
      mp.M = total; mp.up_c = L.up_c; mp.up_d = L.up_d; mp.down_cold = L.down_cold; mp.down_gamma = L.down_gamma;
      mp.in_stats = mid_stats; mp.out_stats = res_stats; mp.inv_width = 1.0f / HIDDEN; mp.ln_eps = sh.ln_eps;
      ProfScope prof(e, ws, KC_MLP, st);
      rc = launch_mlp(e, *tm_mid, L.tf_w1, L.tf_w2, res == ws.x ? ws.t32_x : ws.t32_y, mp, st, pdl);
      if (rc) return rc;
      continue;
    }
    // h = gelu(LN_1(mid) . W1^T + b_1)
    p.N = sh.inter; p.K = HIDDEN; p.colc = L.up_c; p.cold = L.up_d; p.gamma = nullptr; p.in_stats = mid_stats;
    p.residual = nullptr; p.out_stats = nullptr;
    {
      ProfScope prof(e, ws, last_cls ? KC_CLS_TAIL : KC_GEMM_UP_GELU, st);
      rc = (e->gemm_pairs & 16) ? DRAG_GEMM2_UP_WS(e, *tm_mid, L.tp_up, ws.ts_h, ws.ts_h, p, st, pdl)
           : (e->gemm_pairs & 4) ? DRAG_GEMM2_UP(e, *tm_mid, L.tp_up, ws.ts_h, ws.ts_h, p, st, pdl) : DRAG_GEMM_UP(e, *tm_mid, L.tm_up, ws.ts_h, ws.ts_h, p, st, pdl);
    }
    if (rc) return rc;
    // res_raw = h . W2^T + b_2 + LN_1(mid_raw)   (+ row statistics of res_raw); LN_2 is applied by the consumers
    p.N = HIDDEN; p.K = sh.inter; p.colc = nullptr; p.cold = L.down_cold; p.gamma = L.down_gamma; p.in_stats = mid_stats;
    p.residual = mid; p.out_stats = res_stats; p.f16_operands = 1;   // h and W2 are fp16
    {
      ProfScope prof(e, ws, last_cls ? KC_CLS_TAIL : KC_GEMM_DOWN_LN, st);
      rc = (e->gemm_pairs & 8) ? DRAG_GEMM2_RES(e, ws.tm_h, L.tp_down, *ts_res, *ts_mid, p, st, pdl) : DRAG_GEMM_RES(e, ws.tm_h, L.tm_down, *ts_res, *ts_mid, p, st, pdl);
    }
    if (rc) return rc;
    (void)tm_res;
  }
  // LayerNorm that turns the current raw hidden state into the model's hidden state
  const float* fin_g = n_layers == 0 ? e->emb_g : e->layers[n_layers - 1].ln2_g;
  const float* fin_b = n_layers == 0 ? e->emb_b : e->layers[n_layers - 1].ln2_b;
  if (tap) {
    ln_apply_kernel<<<(total + 7) / 8, 256, 0, st>>>(ws.x, ws.stats_x, fin_g, fin_b, sh.ln_eps, total, d_hidden);
    DRAG_CUDA_OK(cudaGetLastError());
  }
  if (d_out) {
    ProfScope prof(e, ws, KC_POOL, st);
    const int blocks = (n_seq + 7) / 8;
    if (cls_tail)   // the last layer left compact rows (row = sequence) in y
      DRAG_CUDA_OK(launch_ex(pool_normalize_kernel, dim3(blocks), dim3(256), 0, st, 1, pdl, (const bf16*)ws.y, (const float2*)ws.stats_y, fin_g, fin_b, sh.ln_eps,
                             (const int*)nullptr, n_seq, d_out));
    else
      DRAG_CUDA_OK(launch_ex(pool_normalize_kernel, dim3(blocks), dim3(256), 0, st, 1, pdl, (const bf16*)ws.x, (const float2*)ws.stats_x, fin_g, fin_b, sh.ln_eps,
                             (const int*)d_cu, n_seq, d_out));
    DRAG_CUDA_OK(cudaGetLastError());
  }
  return DRAG_OK;
}

// W' = bf16(W (.) gamma) [N, K];  c_n = sum_k W'_nk;  d_n = sum_k beta_k W_nk + bias_n
int upload_folded(drag_encoder* e, bf16** w_dev, float** c_dev, float** d_dev, std::initializer_list<const float*> ws,
                  std::initializer_list<const float*> biases, size_t rows_each, size_t K, const float* gamma, const float* beta) {
  const size_t n_total = rows_each * ws.size();
  std::vector<uint16_t> wq(n_total * K);
  std::vector<float> c(n_total), d(n_total);
  size_t r0 = 0;
  auto bias_it = biases.begin();
  for (const float* w : ws) {
    const float* b = *bias_it++;
    for (size_t r = 0; r < rows_each; ++r) {
      double cs = 0.0, ds = 0.0;
      for (size_t k = 0; k < K; ++k) {
        const float wf = w[r * K + k];
        const uint16_t q = f32_to_bf16_rne(wf * gamma[k]);
        wq[(r0 + r) * K + k] = q;
        uint32_t u = (uint32_t)q << 16;
        float qf;
        memcpy(&qf, &u, 4);
        cs += (double)qf;
        ds += (double)beta[k] * (double)wf;
      }
      c[r0 + r] = (float)cs;
      d[r0 + r] = (float)(ds + (double)b[r]);
    }
    r0 += rows_each;
  }
  int rc = dev_alloc(e, w_dev, wq.size());
  if (rc) return rc;
  DRAG_CUDA_OK(cudaMemcpy(*w_dev, wq.data(), wq.size() * 2, cudaMemcpyHostToDevice));
  if ((rc = upload_f32(e, c_dev, c.data(), c.size()))) return rc;
  return upload_f32(e, d_dev, d.data(), d.size());
}

int upload_sum_f32(drag_encoder* e, float** dst, const float* a, const float* b, size_t n) {
  std::vector<float> tmp(n);
  for (size_t i = 0; i < n; ++i) tmp[i] = a[i] + b[i];
  return upload_f32(e, dst, tmp.data(), n);
}

int init_workspace(drag_encoder* e, Workspace& ws, int64_t tokens, bool high_priority) {
  int rc;
  const size_t H = HIDDEN, F = e->shape.inter;
  ws.max_tokens = (tokens + 127) / 128 * 128;
  const size_t T = (size_t)ws.max_tokens;
  if ((rc = dev_alloc(e, &ws.x, T * H))) return rc;
  if ((rc = dev_alloc(e, &ws.y, T * H))) return rc;
  if ((rc = dev_alloc(e, &ws.ctx, T * H))) return rc;
  if ((rc = dev_alloc(e, &ws.qkv, T * 3 * H))) return rc;
  if ((rc = dev_alloc(e, &ws.h, T * F))) return rc;
  if ((rc = dev_alloc(e, &ws.stats_x, T * PARTS))) return rc;
  if ((rc = dev_alloc(e, &ws.stats_y, T * PARTS))) return rc;
  // activations are zeroed once so that the tail rows of the last 128-row tile hold finite values
  DRAG_CUDA_OK(cudaMemset(ws.x, 0, T * H * 2));
  DRAG_CUDA_OK(cudaMemset(ws.y, 0, T * H * 2));
  DRAG_CUDA_OK(cudaMemset(ws.ctx, 0, T * H * 2));
  DRAG_CUDA_OK(cudaMemset(ws.qkv, 0, T * 3 * H * 2));
  DRAG_CUDA_OK(cudaMemset(ws.h, 0, T * F * 2));
  DRAG_CUDA_OK(cudaMemset(ws.stats_x, 0, T * PARTS * 8));
  DRAG_CUDA_OK(cudaMemset(ws.stats_y, 0, T * PARTS * 8));
  if ((rc = make_tmap(&ws.tm_x, ws.x, T, H, gemm::BLOCK_M))) return rc;
  if ((rc = make_tmap(&ws.tm_y, ws.y, T, H, gemm::BLOCK_M))) return rc;
  if ((rc = make_tmap(&ws.tm_ctx, ws.ctx, T, H, gemm::BLOCK_M))) return rc;
  if ((rc = make_tmap(&ws.tm_h, ws.h, T, F, gemm::BLOCK_M))) return rc;
  if ((rc = make_tmap(&ws.ts_x, ws.x, T, H, gemm::STORE_ROWS))) return rc;
  if ((rc = make_tmap(&ws.ts_y, ws.y, T, H, gemm::STORE_ROWS))) return rc;
  if ((rc = make_tmap(&ws.ts_qkv, ws.qkv, T, 3 * H, gemm::STORE_ROWS))) return rc;
  if ((rc = make_tmap(&ws.ts_h, ws.h, T, F, gemm::STORE_ROWS))) return rc;
  if ((rc = make_tmap_bf16_box(&ws.tm_qkv_heads, ws.qkv, T, 3 * H, attn3::TILE, attn3::HEAD_DIM))) return rc;
  if ((rc = make_tmap_bf16_box(&ws.t32_x, ws.x, T, H, 32, mlp::STORE_COLS))) return rc;
  if ((rc = make_tmap_bf16_box(&ws.t32_y, ws.y, T, H, 32, mlp::STORE_COLS))) return rc;
  // host-buffer path: own stream, pinned staging + device mirrors
  int lo = 0, hi = 0;
  DRAG_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = greatest priority (numerically lowest)
  DRAG_CUDA_OK(cudaStreamCreateWithPriority(&ws.stream, cudaStreamNonBlocking, high_priority ? hi : lo));
  DRAG_CUDA_OK(cudaEventCreateWithFlags(&ws.done, cudaEventDisableTiming));
  if ((rc = dev_alloc(e, &ws.d_ids, T))) return rc;
  if ((rc = dev_alloc(e, &ws.d_cu, T + 1))) return rc;
  if (cudaMallocHost((void**)&ws.p_ids, T * 4) != cudaSuccess || cudaMallocHost((void**)&ws.p_cu, (T + 1) * 4) != cudaSuccess) {
    cudaGetLastError();
    return fail(DRAG_ERR_NOMEM, "cudaMallocHost failed");
  }
  return DRAG_OK;
}

void free_workspace(Workspace& ws) {
  if (ws.stream) { cudaStreamSynchronize(ws.stream); cudaStreamDestroy(ws.stream); }
  if (ws.done) cudaEventDestroy(ws.done);
  if (ws.p_ids) cudaFreeHost(ws.p_ids);
  if (ws.p_cu) cudaFreeHost(ws.p_cu);
  if (ws.p_out) cudaFreeHost(ws.p_out);
  if (ws.d_out) cudaFree(ws.d_out);
}

// the workspace a forward over `total` tokens runs in: small batches (the query path) get the query workspace
Workspace& pick_workspace(drag_encoder* e, int64_t total) {
  return total <= e->ws[1].max_tokens ? e->ws[1] : e->ws[0];
}

// One forward enqueued on `st` inside workspace `ws` (caller holds ws.lock): the previous forward that used the
// workspace -- on whatever stream -- is waited for in-stream first, and this one leaves its own marker behind.
int forward_locked(drag_encoder* e, Workspace& ws, const int32_t* d_ids, const int32_t* d_cu, const int32_t* h_cu, int n_seq,
                   float* d_out, int stop_after_layer, float* d_hidden, cudaStream_t st) {
  if (ws.used) DRAG_CUDA_OK(cudaStreamWaitEvent(st, ws.done, 0));
  const int rc = forward_impl(e, ws, d_ids, d_cu, h_cu, n_seq, d_out, stop_after_layer, d_hidden, st);
  ws.used = true;
  DRAG_CUDA_OK(cudaEventRecord(ws.done, st));   // also after a failed enqueue: kernels launched so far still use the buffers
  return rc;
}

}  // namespace

extern "C" int drag_encoder_create(const drag_bert_shape* shape, const float* const* h_tensors, int n_tensors,
                                   int device, int64_t max_tokens, drag_encoder** out) {
  DRAG_REQUIRE(shape && h_tensors && out, "drag_encoder_create: null pointer");
  *out = nullptr;
  DRAG_REQUIRE(shape->hidden == HIDDEN, "drag_encoder_create: this build supports hidden=384 (got %d)", shape->hidden);
  DRAG_REQUIRE(shape->heads * HEAD_DIM == shape->hidden, "drag_encoder_create: head_dim must be 32");
  DRAG_REQUIRE(shape->inter % FFN_BLOCK_N == 0 && shape->inter % FFN_BLOCK_N_PAIR == 0 && shape->inter % gemm::BLOCK_K == 0, "drag_encoder_create: intermediate size must be a multiple of 256");
  static_assert(HIDDEN / RES_BLOCK_N == PARTS, "one statistics slot per residual-GEMM column tile");
  DRAG_REQUIRE(shape->layers >= 1 && shape->vocab >= 1 && shape->max_pos >= 1 && shape->max_pos <= 512, "drag_encoder_create: bad shape");
  DRAG_REQUIRE(n_tensors == 5 + 16 * shape->layers, "drag_encoder_create: expected %d tensors, got %d", 5 + 16 * shape->layers, n_tensors);
  DRAG_REQUIRE(max_tokens >= 1 && max_tokens <= (1ll << 30), "drag_encoder_create: bad max_tokens");
  for (int i = 0; i < n_tensors; ++i) DRAG_REQUIRE(h_tensors[i], "drag_encoder_create: tensor %d is null", i);
  int n_dev = 0, cc = 0, sms = 0;
  int rc = drag_device_info(device, &n_dev, &cc, &sms);
  if (rc) return rc;
  if (cc / 10 != 10) return fail(DRAG_ERR_DEVICE, "drag_encoder_create: device %d is sm_%d; this library is built for sm_100a only", device, cc);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_encoder_create: cannot select device %d", device);

  drag_encoder* e = new (std::nothrow) drag_encoder();
  if (!e) return fail(DRAG_ERR_NOMEM, "drag_encoder_create: out of host memory");
  e->shape = *shape;
  e->device = device;
  e->sms = sms;
  e->max_tokens = (max_tokens + 127) / 128 * 128;
  const size_t H = HIDDEN, F = shape->inter;
  auto bail = [&](int code) { drag_encoder_destroy(e); return code; };

  if ((rc = upload_f32(e, &e->word, h_tensors[0], (size_t)shape->vocab * H))) return bail(rc);
  if ((rc = upload_f32(e, &e->pos, h_tensors[1], (size_t)shape->max_pos * H))) return bail(rc);
  if ((rc = upload_f32(e, &e->type0, h_tensors[2], H))) return bail(rc);  // token_type 0 only
  if ((rc = upload_f32(e, &e->emb_g, h_tensors[3], H))) return bail(rc);
  if ((rc = upload_f32(e, &e->emb_b, h_tensors[4], H))) return bail(rc);
  e->layers.resize(shape->layers);
  for (int l = 0; l < shape->layers; ++l) {
    const float* const* t = h_tensors + 5 + 16 * l;
    // LayerNorm feeding this layer: the embedding LN for layer 0, else the previous layer's output LN
    const float* g_in = l == 0 ? h_tensors[3] : h_tensors[5 + 16 * (l - 1) + 14];
    const float* b_in = l == 0 ? h_tensors[4] : h_tensors[5 + 16 * (l - 1) + 15];
    const float* g_1 = t[12];  // attention.output.LayerNorm
    const float* b_1 = t[13];
    Layer& L = e->layers[l];
    if ((rc = upload_folded(e, &L.w_qkv, &L.qkv_c, &L.qkv_d, {t[0], t[2], t[4]}, {t[1], t[3], t[5]}, H, H, g_in, b_in))) return bail(rc);
    if ((rc = upload_bf16(e, &L.w_o, {t[6]}, H, H))) return bail(rc);
    if ((rc = upload_sum_f32(e, &L.o_cold, t[7], b_in, H))) return bail(rc);
    if ((rc = upload_f32(e, &L.o_gamma, g_in, H))) return bail(rc);
    if ((rc = upload_folded(e, &L.w_up, &L.up_c, &L.up_d, {t[8]}, {t[9]}, F, H, g_1, b_1))) return bail(rc);
    if ((rc = upload_f16(e, &L.w_down, t[10], H, F))) return bail(rc);
    if ((rc = upload_sum_f32(e, &L.down_cold, t[11], b_1, H))) return bail(rc);
    if ((rc = upload_f32(e, &L.down_gamma, g_1, H))) return bail(rc);
    if ((rc = upload_f32(e, &L.ln2_g, t[14], H))) return bail(rc);
    if ((rc = upload_f32(e, &L.ln2_b, t[15], H))) return bail(rc);
    if ((rc = make_tmap(&L.tm_qkv, L.w_qkv, 3 * H, H, QKV_BLOCK_N))) return bail(rc);
    if ((rc = make_tmap(&L.tm_o, L.w_o, H, H, RES_BLOCK_N))) return bail(rc);
    if ((rc = make_tmap(&L.tm_up, L.w_up, F, H, FFN_BLOCK_N))) return bail(rc);
    if ((rc = make_tmap(&L.tm_down, L.w_down, H, F, RES_BLOCK_N))) return bail(rc);
    if ((rc = make_tmap(&L.tp_qkv, L.w_qkv, 3 * H, H, QKV_BLOCK_N / 2))) return bail(rc);
    if ((rc = make_tmap(&L.tp_o, L.w_o, H, H, RES_BLOCK_N_PAIR / 2))) return bail(rc);
    if ((rc = make_tmap(&L.tp_up, L.w_up, F, H, FFN_BLOCK_N_PAIR / 2))) return bail(rc);
    if ((rc = make_tmap(&L.tp_down, L.w_down, H, F, RES_BLOCK_N_PAIR / 2))) return bail(rc);
    if ((rc = make_tmap(&L.tf_w1, L.w_up, F, H, mlp::NC / 2))) return bail(rc);
    if ((rc = make_tmap(&L.tf_w2, L.w_down, H, F, mlp::OUT_HALF / 2))) return bail(rc);
  }
  if ((rc = init_workspace(e, e->ws[0], e->max_tokens, false))) return bail(rc);
  if ((rc = init_workspace(e, e->ws[1], e->max_tokens < QUERY_WORKSPACE_TOKENS ? e->max_tokens : QUERY_WORKSPACE_TOKENS, true))) return bail(rc);
  {
    const char* v = getenv("DRAG_ATTENTION");
    if (v && strcmp(v, "tc3") == 0) e->attention_variant = 3;
    if (v && strcmp(v, "mma") == 0) e->attention_variant = 0;
    const char* gm = getenv("DRAG_GEMM_PAIRS");
    if (gm && gm[0] >= '0' && gm[0] <= '9') e->gemm_pairs = atoi(gm) & 31;
    const char* pd = getenv("DRAG_PDL");
    if (pd && pd[0] == '0') e->pdl = false;
    const char* fm = getenv("DRAG_FUSED_MLP");
    if (fm && fm[0] == '0') e->fused_mlp = false;
    const char* co = getenv("DRAG_CLS_ONLY");
    if (co && co[0] == '0') e->cls_only = false;
  }

  // the attention kernels need > 48 KB of dynamic shared memory
  if ((rc = attention_set_attributes())) return bail(rc);
  if (cudaDeviceSynchronize() != cudaSuccess) return bail(fail(DRAG_ERR_CUDA, "weight upload failed: %s", cudaGetErrorString(cudaGetLastError())));
  *out = e;
  return DRAG_OK;
}

extern "C" int drag_encoder_destroy(drag_encoder* e) {
  if (!e) return DRAG_OK;
  DeviceGuard guard(e->device);
  free_workspace(e->ws[0]);
  free_workspace(e->ws[1]);
  for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
  for (void* p : e->allocs) cudaFree(p);
  cudaGetLastError();
  delete e;
  return DRAG_OK;
}

static int64_t total_tokens_of(const int32_t* h_cu, int n_seq) { return (h_cu && n_seq > 0) ? (int64_t)h_cu[n_seq] : 0; }

extern "C" int drag_encoder_forward(drag_encoder* e, const int32_t* d_ids, const int32_t* d_cu_seqlens,
                                    const int32_t* h_cu_seqlens, int n_seq, float* d_out, void* stream) {
  DRAG_REQUIRE(e, "drag_encoder_forward: null encoder");
  DRAG_REQUIRE(d_out || n_seq == 0, "drag_encoder_forward: null output");
  Workspace& ws = pick_workspace(e, total_tokens_of(h_cu_seqlens, n_seq));
  std::lock_guard<std::mutex> hold(ws.lock);
  DeviceGuard guard(e->device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_encoder_forward: cannot select device %d", e->device);
  return forward_locked(e, ws, d_ids, d_cu_seqlens, h_cu_seqlens, n_seq, d_out, -1, nullptr, (cudaStream_t)stream);
}

extern "C" int drag_encoder_forward_debug(drag_encoder* e, const int32_t* d_ids, const int32_t* d_cu_seqlens,
                                          const int32_t* h_cu_seqlens, int n_seq, int stop_after_layer, float* d_hidden,
                                          void* stream) {
  DRAG_REQUIRE(e && d_hidden, "drag_encoder_forward_debug: null pointer");
  DRAG_REQUIRE(stop_after_layer >= 0 && stop_after_layer <= e->shape.layers, "drag_encoder_forward_debug: bad layer %d", stop_after_layer);
  Workspace& ws = pick_workspace(e, total_tokens_of(h_cu_seqlens, n_seq));
  std::lock_guard<std::mutex> hold(ws.lock);
  DeviceGuard guard(e->device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_encoder_forward_debug: cannot select device %d", e->device);
  return forward_locked(e, ws, d_ids, d_cu_seqlens, h_cu_seqlens, n_seq, nullptr, stop_after_layer, d_hidden, (cudaStream_t)stream);
}

extern "C" int drag_encoder_embed_host(drag_encoder* e, const int32_t* h_ids, const int32_t* h_cu_seqlens, int n_seq,
                                       float* h_out) {
  DRAG_REQUIRE(e, "drag_encoder_embed_host: null encoder");
  if (n_seq == 0) return DRAG_OK;
  DRAG_REQUIRE(h_ids && h_cu_seqlens && h_out && n_seq > 0, "drag_encoder_embed_host: null pointer");
  const int64_t total = h_cu_seqlens[n_seq];
  DRAG_REQUIRE(total >= n_seq && total <= e->max_tokens, "drag_encoder_embed_host: %lld tokens exceed max_tokens=%lld", (long long)total, (long long)e->max_tokens);
  Workspace& ws = pick_workspace(e, total);
  std::lock_guard<std::mutex> hold(ws.lock);
  DeviceGuard guard(e->device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_encoder_embed_host: cannot select device %d", e->device);
  // output staging grows on demand (power-of-two capacity)
  if (n_seq > ws.out_cap) {
    int64_t want = 1024;
    while (want < n_seq) want <<= 1;
    DRAG_CUDA_OK(cudaStreamSynchronize(ws.stream));
    if (ws.d_out) cudaFree(ws.d_out);
    if (ws.p_out) cudaFreeHost(ws.p_out);
    ws.d_out = nullptr; ws.p_out = nullptr; ws.out_cap = 0;
    if (cudaMalloc((void**)&ws.d_out, (size_t)want * HIDDEN * 4) != cudaSuccess ||
        cudaMallocHost((void**)&ws.p_out, (size_t)want * HIDDEN * 4) != cudaSuccess) {
      cudaGetLastError();
      return fail(DRAG_ERR_NOMEM, "drag_encoder_embed_host: cannot allocate output staging for %lld sequences", (long long)want);
    }
    ws.out_cap = want;
  }
  memcpy(ws.p_ids, h_ids, (size_t)total * 4);
  memcpy(ws.p_cu, h_cu_seqlens, (size_t)(n_seq + 1) * 4);
  DRAG_CUDA_OK(cudaMemcpyAsync(ws.d_ids, ws.p_ids, (size_t)total * 4, cudaMemcpyHostToDevice, ws.stream));
  DRAG_CUDA_OK(cudaMemcpyAsync(ws.d_cu, ws.p_cu, (size_t)(n_seq + 1) * 4, cudaMemcpyHostToDevice, ws.stream));
  int rc = forward_locked(e, ws, ws.d_ids, ws.d_cu, ws.p_cu, n_seq, ws.d_out, -1, nullptr, ws.stream);
  if (rc) { cudaStreamSynchronize(ws.stream); return rc; }
  DRAG_CUDA_OK(cudaMemcpyAsync(ws.p_out, ws.d_out, (size_t)n_seq * HIDDEN * 4, cudaMemcpyDeviceToHost, ws.stream));
  DRAG_CUDA_OK(cudaStreamSynchronize(ws.stream));
  memcpy(h_out, ws.p_out, (size_t)n_seq * HIDDEN * 4);
  return DRAG_OK;
}

// ---------------------------------------------------------------------------------
// kernel-level entry points used by the parity tests to check each fused kernel in isolation
// ---------------------------------------------------------------------------------
extern "C" int drag_debug_gemm(int device, int variant, const void* d_a, const void* d_w, const float* d_colc,
                               const float* d_cold, const float* d_gamma, const void* d_in_stats, const void* d_residual,
                               void* d_out, void* d_out_stats, int M, int N, int K, float inv_width, float ln_eps,
                               void* stream) {
  DRAG_REQUIRE(d_a && d_w && d_cold && d_in_stats && d_out && M >= 1, "drag_debug_gemm: null pointer / empty problem");
  DRAG_REQUIRE(K % gemm::BLOCK_K == 0 && K >= gemm::BLOCK_K, "drag_debug_gemm: K must be a multiple of 64");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_debug_gemm: cannot select device %d", device);
  drag_encoder fake;
  fake.device = device;
  fake.sms = sm_count(device);
  gemm::GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.colc = d_colc; p.cold = d_cold; p.gamma = d_gamma; p.in_stats = (const float2*)d_in_stats;
  p.inv_width = inv_width; p.ln_eps = ln_eps; p.residual = (const bf16*)d_residual; p.out_stats = (float2*)d_out_stats;
  CUtensorMap ta, tw, tout, tres;
  int rc = make_tmap(&ta, d_a, (uint64_t)M, (uint64_t)K, gemm::BLOCK_M);
  if (rc) return rc;
  if ((rc = make_tmap(&tout, d_out, (uint64_t)M, (uint64_t)N, gemm::STORE_ROWS))) return rc;
  tres = tout;
  if (d_residual && (rc = make_tmap(&tres, d_residual, (uint64_t)M, (uint64_t)N, gemm::STORE_ROWS))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (const char* dbg = getenv("DRAG_GEMM_DBG")) p.dbg = atoi(dbg);
  if (variant >= 20) { p.f16_operands = 1; variant -= 20; }   // A and W hold fp16
  switch (variant) {
    case 0:
      DRAG_REQUIRE(N % QKV_BLOCK_N == 0 && d_colc, "drag_debug_gemm: variant 0 needs N %% %d == 0 and colc", QKV_BLOCK_N);
      if ((rc = make_tmap(&tw, d_w, (uint64_t)N, (uint64_t)K, QKV_BLOCK_N))) return rc;
      return DRAG_GEMM_QKV(&fake, ta, tw, tout, tres, p, st);
    case 1:
      DRAG_REQUIRE(N % FFN_BLOCK_N == 0 && d_colc, "drag_debug_gemm: variant 1 needs N %% %d == 0 and colc", FFN_BLOCK_N);
      if ((rc = make_tmap(&tw, d_w, (uint64_t)N, (uint64_t)K, FFN_BLOCK_N))) return rc;
      return DRAG_GEMM_UP(&fake, ta, tw, tout, tres, p, st);
    case 2:
      DRAG_REQUIRE(N == HIDDEN && d_gamma && d_residual && d_out_stats, "drag_debug_gemm: variant 2 needs N=384, gamma, residual, out_stats");
      if ((rc = make_tmap(&tw, d_w, (uint64_t)N, (uint64_t)K, RES_BLOCK_N))) return rc;
      return DRAG_GEMM_RES(&fake, ta, tw, tout, tres, p, st);
    case 10:
      DRAG_REQUIRE(N % QKV_BLOCK_N == 0 && d_colc, "drag_debug_gemm: variant 10 needs N %% %d == 0 and colc", QKV_BLOCK_N);
      if ((rc = make_tmap(&tw, d_w, (uint64_t)N, (uint64_t)K, QKV_BLOCK_N / 2))) return rc;
      return DRAG_GEMM2_QKV(&fake, ta, tw, tout, tres, p, st);
    case 11:
      DRAG_REQUIRE(N % FFN_BLOCK_N_PAIR == 0 && d_colc, "drag_debug_gemm: variant 11 needs N %% %d == 0 and colc", FFN_BLOCK_N_PAIR);
      if ((rc = make_tmap(&tw, d_w, (uint64_t)N, (uint64_t)K, FFN_BLOCK_N_PAIR / 2))) return rc;
      return DRAG_GEMM2_UP(&fake, ta, tw, tout, tres, p, st);
    case 12:
      DRAG_REQUIRE(N == HIDDEN && d_gamma && d_residual && d_out_stats, "drag_debug_gemm: variant 12 needs N=384, gamma, residual, out_stats");
      if ((rc = make_tmap(&tw, d_w, (uint64_t)N, (uint64_t)K, RES_BLOCK_N_PAIR / 2))) return rc;
      return K <= gemm::WS_K_BLOCKS * gemm::BLOCK_K ? DRAG_GEMM2_RES_WS(&fake, ta, tw, tout, tres, p, st) : DRAG_GEMM2_RES(&fake, ta, tw, tout, tres, p, st);
    default:
      return fail(DRAG_ERR_INVALID, "drag_debug_gemm: unknown variant %d", variant);
  }
}

extern "C" int drag_debug_mlp(int device, const void* d_x, const void* d_in_stats, const void* d_w1g, const float* d_up_c,
                              const float* d_up_d, const void* d_w2, const float* d_down_cold, const float* d_down_gamma,
                              void* d_out, void* d_out_stats, int M, float ln_eps, void* stream) {
  DRAG_REQUIRE(d_x && d_in_stats && d_w1g && d_up_c && d_up_d && d_w2 && d_down_cold && d_down_gamma && d_out && d_out_stats && M >= 1,
               "drag_debug_mlp: bad arguments");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_debug_mlp: cannot select device %d", device);
  drag_encoder fake{};
  fake.device = device;
  fake.sms = sm_count(device);
  DRAG_REQUIRE(fake.sms >= 2, "drag_debug_mlp: cannot read the SM count of device %d", device);
  mlp::MlpParams p{};
  p.M = M; p.up_c = d_up_c; p.up_d = d_up_d; p.down_cold = d_down_cold; p.down_gamma = d_down_gamma;
  p.in_stats = (const float2*)d_in_stats; p.out_stats = (float2*)d_out_stats; p.inv_width = 1.0f / HIDDEN; p.ln_eps = ln_eps;
  if (const char* dbg = getenv("DRAG_MLP_DBG")) p.dbg = atoi(dbg);
  CUtensorMap tx, tw1, tw2, tout;
  int rc;
  if ((rc = make_tmap(&tx, d_x, (uint64_t)M, HIDDEN, mlp::ROWS))) return rc;
  if ((rc = make_tmap(&tw1, d_w1g, mlp::INTER, HIDDEN, mlp::NC / 2))) return rc;
  if ((rc = make_tmap(&tw2, d_w2, HIDDEN, mlp::INTER, mlp::OUT_HALF / 2))) return rc;
  if ((rc = make_tmap_bf16_box(&tout, d_out, (uint64_t)M, HIDDEN, 32, mlp::STORE_COLS))) return rc;
  if (getenv("DRAG_MLP_TRACE")) {
    // probes only: where CTA 0's MMA issuer and first epilogue warp spend their clocks
    long long* d_trace = nullptr;
    long long h[20];
    DRAG_CUDA_OK(cudaMalloc((void**)&d_trace, sizeof(h)));
    DRAG_CUDA_OK(cudaMemset(d_trace, 0, sizeof(h)));
    p.trace = d_trace;
    rc = launch_mlp(&fake, tx, tw1, tw2, tout, p, (cudaStream_t)stream);
    if (rc == DRAG_OK && cudaStreamSynchronize((cudaStream_t)stream) == cudaSuccess &&
        cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
      const int tiles = (M + 2 * mlp::ROWS - 1) / (2 * mlp::ROWS), pairs = fake.sms / 2;
      const int mine = (tiles + pairs - 1) / pairs;
      fprintf(stderr, "mlp trace (clocks per tile of CTA 0, %d tiles): issuer total %lld | waits x_full %lld w1_full %lld w2_full %lld hp_full %lld out_free %lld || "
              "epilogue warp total %lld | table+bar.sync %lld hacc_full wait %lld E1 body %lld residual copy + E2 %lld (E2: out_full wait %lld, "
              "tensor-memory loads %lld, arithmetic %lld, staging + TMA store %lld, final store drain %lld)\n", mine,
              h[7] / mine, h[0] / mine, h[1] / mine, h[2] / mine, h[3] / mine, h[4] / mine, h[15] / mine, h[8] / mine, h[9] / mine, h[10] / mine, h[11] / mine,
              h[12] / mine, h[13] / mine, h[14] / mine, h[16] / mine, h[17] / mine);
    }
    cudaFree(d_trace);
    return rc;
  }
  return launch_mlp(&fake, tx, tw1, tw2, tout, p, (cudaStream_t)stream);
}

extern "C" int drag_debug_attention(int device, int variant, const void* d_qkv, void* d_ctx, const int32_t* d_cu_seqlens,
                                    int n_seq, int n_tokens, int max_len, int heads, void* stream) {
  DRAG_REQUIRE(d_qkv && d_ctx && d_cu_seqlens && n_seq >= 1 && heads >= 1 && n_tokens >= 1, "drag_debug_attention: bad arguments");
  DRAG_REQUIRE(max_len >= 1 && max_len <= 512, "drag_debug_attention: max_len must be in 1..512");
  DRAG_REQUIRE(variant == 0 || variant == 3 || variant == 7, "drag_debug_attention: variant 0 (mma.sync), 3 (tcgen05, two softmax groups, P in tensor memory), 7 (3 with the debug timeline)");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_debug_attention: cannot select device %d", device);
  int rc = attention_set_attributes();
  if (rc) return rc;
  CUtensorMap tm;
  if (variant >= 1 &&
      (rc = make_tmap_bf16_box(&tm, d_qkv, (uint64_t)n_tokens, (uint64_t)3 * heads * HEAD_DIM, attn3::TILE, attn3::HEAD_DIM)))
    return rc;
  return launch_attention(variant, tm, (const bf16*)d_qkv, (bf16*)d_ctx, d_cu_seqlens, n_seq, max_len, heads, (cudaStream_t)stream);
}

// Phase timeline of the tcgen05 attention kernel (debug): points the kernel's trace pointer at `d_trace`
// (int64[drag_debug_attention_trace_words()], zeroed by the caller; nullptr switches tracing off); variant 7 of
// drag_debug_attention then records clock64() stamps of CTA 0.
extern "C" int drag_debug_attention_trace_words(void) { return attn3::TRACE_WORDS; }
extern "C" int drag_debug_set_attention_trace(int device, void* d_trace) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DRAG_ERR_DEVICE, "drag_debug_set_attention_trace: cannot select device %d", device);
  long long* p = (long long*)d_trace;
  DRAG_CUDA_OK(cudaMemcpyToSymbol(attn3::g_trace, &p, sizeof(p)));
  return DRAG_OK;
}

// ---------------------------------------------------------------------------------
// per-kernel-class device timing for bench.py's roofline (CUDA events on the launch stream)
// ---------------------------------------------------------------------------------
extern "C" int drag_encoder_profile_begin(drag_encoder* e, int max_launches) {
  DRAG_REQUIRE(e && max_launches >= 1 && max_launches <= (1 << 20), "drag_encoder_profile_begin: bad arguments");
  std::lock_guard<std::mutex> hold(e->ws[0].lock);
  DeviceGuard guard(e->device);
  while (e->prof_events.size() < (size_t)max_launches * 2) {
    cudaEvent_t ev;
    DRAG_CUDA_OK(cudaEventCreate(&ev));
    e->prof_events.push_back(ev);
  }
  e->prof_class.clear();
  e->prof_used = 0;
  e->profiling = true;
  return DRAG_OK;
}

// ms_by_class / launches_by_class: arrays of 8 (embed, gemm_qkv, attention, gemm_out_ln, gemm_up_gelu,
// gemm_down_ln, pool, cls_tail = the CLS-only kernels of the last layer).  The caller must have synchronised the stream(s) the forwards ran on.
extern "C" int drag_encoder_profile_end(drag_encoder* e, double* ms_by_class, int* launches_by_class) {
  DRAG_REQUIRE(e && ms_by_class && launches_by_class, "drag_encoder_profile_end: null pointer");
  std::lock_guard<std::mutex> hold(e->ws[0].lock);
  DeviceGuard guard(e->device);
  e->profiling = false;
  for (int i = 0; i < KC_COUNT; ++i) { ms_by_class[i] = 0.0; launches_by_class[i] = 0; }
  for (size_t i = 0; i < e->prof_class.size(); ++i) {
    float ms = 0.f;
    DRAG_CUDA_OK(cudaEventElapsedTime(&ms, e->prof_events[2 * i], e->prof_events[2 * i + 1]));
    ms_by_class[e->prof_class[i]] += ms;
    launches_by_class[e->prof_class[i]] += 1;
  }
  return DRAG_OK;
}
