/*
 * drag_b200.h -- C ABI of the B200-native semantic-retriever hot path for Dial RAG.
 *
 * This is the drop-in boundary: plain pointers and sizes, no
 * torch/pybind types.  The reference (epam/ai-dial-rag, pure Python) has no FFI
 * for this path today; each entry point below names the reference code it
 * replaces (paths relative to the reference repo) and INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, a DRAG_ERR_* code otherwise and never
 *     throws; drag_last_error() returns a thread-local message for the last
 *     failure on the calling thread;
 *   - "d_" pointers are DEVICE pointers owned by the caller (e.g. a torch tensor's
 *     data_ptr()) and must stay valid until the work queued on `stream` is done;
 *     "h_" pointers are HOST pointers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions are re-entrant.  An encoder object owns two activation workspaces,
 *     a bulk one (max_tokens) and a small one for query-sized batches (<= 8192
 *     tokens) with its own high-priority stream: an indexing call and a query
 *     call (the reference's two 1-worker pools, aidial_rag/resources/cpu_pools.py:
 *     50-59) run concurrently; calls that share a workspace are serialised -- host
 *     threads by a lock, their GPU work by an event the next forward waits for
 *     in-stream, whatever stream each caller passed.
 */
#ifndef DRAG_B200_H_
#define DRAG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRAG_ABI_VERSION 2   /* 2: drag_encoder_profile_end fills nine classes; drag_debug_mlp; attention trace entry points */

enum drag_status {
  DRAG_OK = 0,
  DRAG_ERR_INVALID = 1,     /* bad argument */
  DRAG_ERR_CUDA = 2,        /* CUDA runtime/driver error (message has details) */
  DRAG_ERR_UNSUPPORTED = 3, /* shape outside what the kernels implement */
  DRAG_ERR_NOMEM = 4,
  DRAG_ERR_DEVICE = 5       /* no sm_100 device / wrong architecture */
};

/* aidial_rag/retrievers/embeddings_metrics.py:7-11 (enum Metric) */
enum drag_metric {
  DRAG_METRIC_COSINE_SIM = 0,
  DRAG_METRIC_EUCLIDEAN_DIST = 1,
  DRAG_METRIC_SQEUCLIDEAN_DIST = 2,
  DRAG_METRIC_INNER_PRODUCT = 3
};

enum drag_dtype { DRAG_F32 = 0, DRAG_BF16 = 1 };

const char* drag_last_error(void);
int drag_abi_version(void);
/* Number of CUDA devices and compute capability (major*10+minor) of `device`. */
int drag_device_info(int device, int* n_devices, int* compute_capability, int* sm_count);

/* ------------------------------------------------------------------------- *
 *  Encoder: bge-small-en style BERT -> CLS -> L2 normalise                   *
 *  replaces HuggingFaceBgeEmbeddings/SentenceTransformer.encode as used by   *
 *  aidial_rag/embeddings/embeddings.py:52-66 (construction), :79-96 (calls). *
 * ------------------------------------------------------------------------- */

typedef struct drag_encoder drag_encoder;

typedef struct drag_bert_shape {
  int32_t vocab;      /* 30522 */
  int32_t hidden;     /* 384   (must be 384 in this build) */
  int32_t layers;     /* 12    */
  int32_t heads;      /* 12    (head_dim must be 32) */
  int32_t inter;      /* 1536  */
  int32_t max_pos;    /* 512   */
  int32_t type_vocab; /* 2     */
  float ln_eps;       /* 1e-12 */
} drag_bert_shape;

/*
 * h_tensors: n_tensors host fp32 arrays in HF BertModel state-dict order:
 *   word_embeddings, position_embeddings, token_type_embeddings, emb LN weight, emb LN bias,
 *   then per layer: q.w q.b k.w k.b v.w v.b attn_out.w attn_out.b inter.w inter.b out.w out.b
 *                   attn_LN.w attn_LN.b out_LN.w out_LN.b          (5 + 16*layers tensors)
 * Linear weights are [out, in] row-major exactly as stored by HF.  They are
 * converted to bf16 and uploaded to `device`; nothing is borrowed after return.
 * max_tokens bounds the packed token count of one forward call (workspace size).
 * Replaces bge_embedding_impl() (embeddings.py:52-66).
 */
int drag_encoder_create(const drag_bert_shape* shape, const float* const* h_tensors,
                        int n_tensors, int device, int64_t max_tokens, drag_encoder** out);
int drag_encoder_destroy(drag_encoder* enc);

/*
 * One forward over a packed (padding-free) batch resident on the device.
 *   d_ids        int32[total_tokens]  token ids, sequences back to back
 *   d_cu_seqlens int32[n_seq + 1]     prefix sums of sequence lengths (cu[0] = 0)
 *   h_cu_seqlens same values on the host (needed to plan the launch)
 *   d_out        f32[n_seq, hidden]   L2-normalised CLS embeddings
 * Every sequence length must be in [1, max_pos].  Asynchronous on `stream`.
 * Replaces SentenceTransformer.encode -> BertModel.forward -> Pooling(cls) ->
 * Normalize -> F.normalize (call sites embeddings.py:80-82, :94-96).
 */
int drag_encoder_forward(drag_encoder* enc, const int32_t* d_ids, const int32_t* d_cu_seqlens,
                         const int32_t* h_cu_seqlens, int n_seq, float* d_out, void* stream);

/*
 * Host-buffer form of the same call (what HuggingFaceBgeEmbeddings.embed_documents
 * hands back to embeddings.py:84-91): token ids in, float32 rows out, host memory on
 * both sides; copies, forward and synchronisation happen inside.
 */
int drag_encoder_embed_host(drag_encoder* enc, const int32_t* h_ids, const int32_t* h_cu_seqlens,
                            int n_seq, float* h_out);

/* Device-side debug taps used by the parity tests (layer = 0 -> embedding LayerNorm
 * output, l -> output of encoder layer l).  d_hidden: f32[total_tokens, hidden]. */
int drag_encoder_forward_debug(drag_encoder* enc, const int32_t* d_ids, const int32_t* d_cu_seqlens,
                               const int32_t* h_cu_seqlens, int n_seq, int stop_after_layer,
                               float* d_hidden, void* stream);

/*
 * Per-kernel-class device timing (CUDA events on the launch stream) for roofline reporting.
 * _begin arms up to max_launches event pairs; _end (after the caller synchronised the stream)
 * returns summed milliseconds and launch counts for the 9 classes
 *   {embed+LN, QKV GEMM, attention, out-proj GEMM+LN, FFN-up GEMM+GELU, FFN-down GEMM+LN, pool,
 *    CLS-only tail of the last layer, fused feed-forward kernel}.  Only forwards in the bulk workspace are recorded.
 */
/* Debug: phase timeline (clock64 stamps of CTA 0) of the tcgen05 attention kernel, see scripts/attn_trace.py. */
int drag_debug_attention_trace_words(void);
int drag_debug_set_attention_trace(int device, void* d_trace);
int drag_encoder_profile_begin(drag_encoder* enc, int max_launches);
int drag_encoder_profile_end(drag_encoder* enc, double* ms_by_class, int* launches_by_class);

/* ------------------------------------------------------------------------- *
 *  Index: row norms + exact top-k                                            *
 *  replaces ENUM_TO_METRIC[...] + np.argsort(kind="stable")[:limit] of       *
 *  aidial_rag/retrievers/embeddings_index.py:51-89 and                       *
 *  aidial_rag/retrievers/embeddings_metrics.py:14-50.                        *
 * ------------------------------------------------------------------------- */

/*
 * d_out[i] = float32 sum of squares of row i, evaluated in numpy's pairwise order
 * so that it equals `np.sum(docs**2, axis=1)` (embeddings_metrics.py:40) bit for bit.
 * Needed once per index for the (sq)euclidean metrics.
 */
int drag_row_sqnorm(const void* d_matrix, int dtype, int64_t n_rows, int dim,
                    float* d_out, void* stream);

/*
 * All distances of ONE query to every row: d_out[i] = ENUM_TO_METRIC[metric](query, docs)[i]
 * (embeddings_metrics.py:53-58), float64.  d_scratch16: 16 bytes of device scratch.
 */
int drag_distances(int device, const void* d_matrix, int dtype, int64_t n_rows, int dim,
                   const float* d_row_sqnorm, const double* d_query, int metric, double* d_out,
                   void* d_scratch16, void* stream);

/* Bytes of scratch drag_topk needs for (n_queries, k) on `device`. */
int drag_topk_workspace_bytes(int device, int n_queries, int k, size_t* bytes);

/*
 * Exact top-k of every query against one row-major matrix shard.
 *   d_matrix     [n_rows, dim] f32 or bf16, row-major, dense
 *   d_row_sqnorm f32[n_rows] from drag_row_sqnorm (may be NULL for cosine / inner product)
 *   d_queries    f64[n_queries, dim]  (the reference's query is float64,
 *                aidial_rag/retrievers/semantic_retriever.py:49,53)
 *   row_id_base  added to local row numbers (row-sharded index: shard offset)
 *   d_out_dist   f64[n_queries, k]  "smaller is better" distance, ascending
 *   d_out_row    i64[n_queries, k]  global row id; ties -> lowest row id
 *                (== earlier document, then lower row: embeddings_index.py:58,81)
 *   d_out_count  i32[n_queries]     min(k, n_rows) valid entries per query
 * Scores are accumulated in float64 like the reference's numpy path; NaN
 * distances (sqrt of a negative rounding residue, embeddings_metrics.py:50) sort
 * last, as numpy's argsort does.  Asynchronous on `stream`.
 */
int drag_topk(int device, const void* d_matrix, int dtype, int64_t n_rows, int dim,
              const float* d_row_sqnorm, const double* d_queries, int n_queries, int k,
              int metric, int64_t row_id_base, double* d_out_dist, int64_t* d_out_row,
              int32_t* d_out_count, void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * Merge per-shard results (after an all-gather) into the global top-k.
 *   d_in_dist/d_in_row/d_in_count : [n_shards, n_queries, k] / [n_shards, n_queries]
 * Same ordering rule as drag_topk.  Output arrays as in drag_topk.
 */
int drag_topk_merge(int device, const double* d_in_dist, const int64_t* d_in_row,
                    const int32_t* d_in_count, int n_shards, int n_queries, int k,
                    double* d_out_dist, int64_t* d_out_row, int32_t* d_out_count, void* stream);

/*
 * Map winning global rows to (document, chunk) pairs: doc = the i with
 * d_doc_offsets[i] <= row < d_doc_offsets[i+1] (empty documents are skipped, as in
 * embeddings_index.py:67-68), chunk = d_row_chunk_ids[row] (embeddings_index.py:60).
 */
int drag_rows_to_chunks(const int64_t* d_rows, int64_t n, const int64_t* d_doc_offsets,
                        int n_docs, const int64_t* d_row_chunk_ids, int64_t* d_out_doc,
                        int64_t* d_out_chunk, void* stream);


/* ------------------------------------------------------------------------- *
 *  Batched exact top-k (tensor cores + float64 re-rank)                      *
 *  same contract as drag_topk -- the reference's per-query find()            *
 *  (embeddings_index.py:62-89) for MANY queries sharing one pass over the    *
 *  matrix (langchain's retriever.batch / eval/eval_retriever.py:97).         *
 * ------------------------------------------------------------------------- */

/* bf16 (round-to-nearest-even) copy of a float32 matrix: the tensor-core scoring operand. */
int drag_rows_to_bf16(const float* d_in, int64_t n_elems, void* d_out, void* stream);

/*
 * d_stats4 <- {max |row|^2, min non-zero |row|^2, number of non-finite |row|^2, 0};
 * d_inv_norm (optional) <- 1 / max(|row|, 1e-8)   (cosine keys).  Input: drag_row_sqnorm output.
 */
int drag_row_norm_stats(const float* d_row_sqnorm, int64_t n_rows, float* d_inv_norm, float* d_stats4, void* stream);

int drag_topk_batch_workspace_bytes(int device, int n_queries, int k, int dim, size_t* bytes);

/*
 * Exact top-k of a batch of queries: candidates are generated with bf16 tcgen05 scores under a
 * certified error bound, survivors are re-scored in float64 exactly like drag_topk and sorted on
 * (distance, row id).  Results are identical to drag_topk's.
 *   d_matrix       the matrix the exact scores are taken from (f32 or bf16)
 *   d_shadow_bf16  bf16 copy of it (== d_matrix when dtype is bf16), 16-byte aligned
 *   d_row_inv_norm from drag_row_norm_stats (cosine only, else may be NULL)
 *   max_row_norm   sqrt(stats[0])
 *   d_out_status   i32[n_queries]: 0 = answered; 1 = the certificate did not cover this query
 *                  (candidate overflow, NaN distances) -> the caller re-runs it through drag_topk.
 * Requires dim % 64 == 0, dim <= 512, k <= 256, finite rows.  Asynchronous on `stream`.
 */
int drag_topk_batch(int device, const void* d_matrix, int dtype, const void* d_shadow_bf16, int64_t n_rows,
                    int dim, const float* d_row_sqnorm, const float* d_row_inv_norm, float max_row_norm,
                    const double* d_queries, int n_queries, int k, int metric, int64_t row_id_base,
                    double* d_out_dist, int64_t* d_out_row, int32_t* d_out_count, int32_t* d_out_status,
                    void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 *  Kernel-level entry points (parity tests only): run ONE fused kernel.      *
 * ------------------------------------------------------------------------- */

/*
 * OUT[M,N] = epilogue(A[M,K] . W[N,K]^T), bf16 in/out, fp32 accumulate (tcgen05), LayerNorm folded:
 * with (mu, rstd) of a row taken from d_in_stats ([M][3] float2 partial (sum, sum^2), width 1/inv_width)
 *   variant 0: out = rstd*(acc - mu*colc) + cold                       (N % 192 == 0)
 *   variant 1: out = gelu(rstd*(acc - mu*colc) + cold), stored as FP16  (N % 256 == 0)
 *   variant 2: out = acc + cold + (residual - mu)*rstd*gamma, and the row's per-128-column-tile
 *              (sum, sum^2) of `out` -> d_out_stats [M][3]             (N == 384)
 *   +10: the CTA-pair (cta_group::2) form (variant 12 leaves the third statistics slot untouched: 192-column tiles);
 *   +20: A and W hold fp16 instead of bf16 (the FFN-down GEMM reads the fp16 GELU output).
 * DRAG_GEMM_DBG=<mask> (probes only): 1 no output stores, 2 no epilogue, 4 stores land on one row block.
 */
int drag_debug_gemm(int device, int variant, const void* d_a, const void* d_w, const float* d_colc,
                    const float* d_cold, const float* d_gamma, const void* d_in_stats, const void* d_residual,
                    void* d_out, void* d_out_stats, int M, int N, int K, float inv_width, float ln_eps,
                    void* stream);

/*
 * The fused feed-forward block (csrc/drag_mlp.cuh), bge-small shape only (384 -> 1536 -> 384):
 *   out = gelu(LN_1(x) . W1^T + b_1) . W2^T + b_2 + LN_1(x)   (raw bf16 rows) and d_out_stats [M][3] partial (sum, sum^2)
 * d_x bf16 [M,384] raw rows with d_in_stats [M][3]; d_w1g bf16 [1536,384] = gamma_1 (.) W1 with the folded terms
 * d_up_c / d_up_d [1536]; d_w2 FP16 [384,1536]; d_down_cold = b_2 + beta_1, d_down_gamma = gamma_1 [384].
 */
int drag_debug_mlp(int device, const void* d_x, const void* d_in_stats, const void* d_w1g, const float* d_up_c,
                   const float* d_up_d, const void* d_w2, const float* d_down_cold, const float* d_down_gamma,
                   void* d_out, void* d_out_stats, int M, float ln_eps, void* stream);

/* ctx[T, heads*32] = softmax(Q K^T / sqrt(32)) V per packed sequence; d_qkv is bf16 [n_tokens, 3*heads*32].
 * variant 0 = mma.sync kernel (what the encoder uses for sequences of up to 256 tokens), 3 = tcgen05 kernel (longer
 * sequences; DRAG_ATTENTION=mma / tc3 force one for every length), 7 = 3 with the debug timeline. */
int drag_debug_attention(int device, int variant, const void* d_qkv, void* d_ctx, const int32_t* d_cu_seqlens,
                         int n_seq, int n_tokens, int max_len, int heads, void* stream);

/*
 * Approximate "bigger is better" keys of the batched path's score kernel for every (query, row):
 * inner product: s; (sq)euclidean: s - |d|^2/2 (d_colvec = row_sqnorm); cosine: s / |d| (d_colvec =
 * inverse norms), with s = bf16(q) . bf16(d) accumulated in fp32 by tcgen05.  n_rows <= 4096.
 * d_out_keys: f32[n_queries, n_rows].  Workspace: drag_topk_batch_workspace_bytes(.., k = 1, ..).
 */
int drag_debug_tc_keys(int device, const void* d_shadow_bf16, int64_t n_rows, int dim, const float* d_colvec,
                       int metric, const double* d_queries, int n_queries, float* d_out_keys,
                       void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 *  Host-side WordPiece: the ASCII fast path of the uncased     *
 *  BERT tokenizer the reference runs inside sentence-transformers            *
 *  (aidial_rag/embeddings/embeddings.py:57-65 -> BertTokenizerFast).         *
 * ------------------------------------------------------------------------- */
typedef struct drag_wordpiece drag_wordpiece;

/* vocab_path: BERT vocab.txt (id = line number; needs [UNK] [CLS] [SEP]). */
int drag_wordpiece_create(const char* vocab_path, int lowercase, drag_wordpiece** out);
int drag_wordpiece_destroy(drag_wordpiece* tk);

/*
 * Tokenise n_texts UTF-8 strings (text i = bytes[offsets[i] : offsets[i+1]]) on n_threads host threads
 * (<= 0: all cores): clean-up, lower-casing, whitespace / punctuation split, greedy WordPiece,
 * [CLS] ids[: max_len-2] [SEP].  out_ids: int32 [n_texts][max_len]; out_len[i] = number of ids of text i, or
 * -1 when the text holds a non-ASCII byte or a literal special token and must go through the reference
 * tokenizer (identical results by construction; the caller splices them in).  No GPU involved.
 */
int drag_wordpiece_encode(const drag_wordpiece* tk, const char* bytes, const int64_t* offsets, int n_texts,
                          int max_len, int n_threads, int32_t* out_ids, int32_t* out_len);

#ifdef __cplusplus
}
#endif
#endif /* DRAG_B200_H_ */
