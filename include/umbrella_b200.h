/*
 * umbrella_b200 -- C ABI of the B200-native speculative-decoding verify path.
 *
 * Drop-in boundary for the hot path of Infini-AI-Lab/UMbreLLa (reference @ ff1ac17).  Every entry
 * point takes plain device pointers, sizes and a cudaStream_t (as void*); no torch types, no
 * allocation, no host synchronisation, no default-stream use inside.  Return value: 0 = ok,
 * negative = argument / capability error (see UMB_ERR_*), positive = cudaError_t.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   (i)   awq_ext.gemm_forward_cuda / awq_ext.dequantize_weights_cuda, called from
 *         umbrella/quantization/awq_utils.py:70-77            -> umb_awq_repack, umb_awq_gemm_f16,
 *                                                                umb_awq_dequantize_f16
 *         torch F.linear for fp16 weights (lm_head llama.py:321, fp16 models llama.py:87-110)
 *                                                             -> umb_gemm_f16
 *   (ii)  flashinfer.single_prefill_with_kv_cache(custom_mask=...), called from
 *         umbrella/attn/cache.py:77-85; KV append cache.py:53-65 -> umb_tree_attn_f16, umb_rope_append_f16
 *   (iii) the accept block of verify() static_speculation_engine.py:312-349 /
 *         dynamic_speculation_engine.py:283-323 and KV_Cache.gather_kv_incremental
 *         cache.py:41-49,136-144                               -> umb_tree_accept, umb_kv_gather
 *   glue  flashinfer.rmsnorm (model_utils.py:62), apply_rotary_pos_emb (model_utils.py:24-52),
 *         silu*up / residual adds (llama.py:293-301), F.embedding (llama.py:312), argmax
 *         (static_speculation_engine.py:307)                   -> umb_rmsnorm_f16, umb_silu_mul_f16,
 *                                                                umb_add_f16, umb_embedding_f16, umb_argmax_f32
 */
#ifndef UMBRELLA_B200_H
#define UMBRELLA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* umb_stream_t; /* cudaStream_t */

#define UMB_OK 0
#define UMB_ERR_ARG (-1)         /* bad pointer / size / alignment */
#define UMB_ERR_UNSUPPORTED (-2) /* shape or option outside what the kernels implement */
#define UMB_ERR_WORKSPACE (-3)   /* workspace too small */
#define UMB_ERR_DRIVER (-4)      /* driver entry point (cuTensorMapEncodeTiled) unavailable */

const char* umb_version(void);
const char* umb_error_string(int code);
/* Number of kernel launches issued through this library since load (all entry points). */
unsigned long long umb_launch_count(void);
/* A replayed CUDA graph re-issues the launches recorded at capture: the host adds them here. */
void umb_launch_count_add(unsigned long long n);

/* ------------------------------------------------------------------------------------------------
 * (i) AWQ INT4 GEMM  --  y[M,N] = x[M,K] . W,  W[k, n] = (q[k,n] - z[k/g,n]) * s[k/g,n]
 *
 * Weights are repacked ONCE (at AwqLinear.init_parameters / copy time) from the AutoAWQ "GEMM"
 * checkpoint layout into a K-major tile stream laid out for TMA bulk copies:
 *   tile t (128 output features) x group g (128 input features) -> one 8512-byte record
 *   [2 k-blocks][2 chunks][128 rows][16 B nibbles] | [128] fp16 scales | [128] 4-bit zeros
 * Same 0.5 + 2/128 + 0.5/128 bytes per parameter as the checkpoint.  Several checkpoint matrices
 * that share K can be packed into one stream (q|k|v, gate|up) through the row mapping
 *   packed_row(n) = (n / row_block) * row_stride + row_offset + n % row_block .
 * ---------------------------------------------------------------------------------------------- */
size_t umb_awq_packed_bytes(int K, int N_total, int group);

/* qweight int32 [K, N/8], qzeros int32 [K/group, N/8], scales fp16 [K/group, N]  (device, AutoAWQ
 * nibble order [0,4,1,5,2,6,3,7]) -> rows of `packed` (device, umb_awq_packed_bytes(K, N_total)).
 * The caller zero-fills `packed` once before the first repack into it. */
int umb_awq_repack(const int32_t* qweight, const int32_t* qzeros, const void* scales, int K, int N, int group,
                   void* packed, int N_total, int row_block, int row_stride, int row_offset, umb_stream_t stream);

/* awq_ext.dequantize_weights_cuda drop-in: fp16 [K, N] = fp16((q - z) * s). */
int umb_awq_dequantize_f16(const int32_t* qweight, const int32_t* qzeros, const void* scales, int K, int N,
                           int group, void* out, umb_stream_t stream);

/* Inverse of the repack for tests: packed stream -> fp16 [N_total, K] (row = packed row). */
int umb_awq_unpack_f16(const void* packed, int K, int N_total, int group, void* out, umb_stream_t stream);

/* Epilogue applied by the CTA that completes an output tile. */
enum {
    UMB_EPI_STORE = 0,    /* y[m, n] = fp16(acc) (+ bias[n])                                   */
    UMB_EPI_RESIDUAL = 1, /* y[m, n] = fp16(acc) +_fp16 residual[m, n]   (llama.py:293,301)      */
    UMB_EPI_SILU_MUL = 2, /* packed rows interleave 64 gate | 64 up per tile;
                             y[m, 64 t + r] = fp16(silu(fp16(gate))) *_fp16 fp16(up)  (llama.py:296-299) */
    UMB_EPI_LOGITS = 3,   /* y is fp32 [M, N]: float(fp16(acc)); optional running arg-max keys   */
    UMB_EPI_ROPE_QKV = 4  /* fused q|k|v projection: fp16(acc) (+ bias), RoPE on the q and k heads (the three rounded
                             fp16 ops of apply_rotary_pos_emb, model_utils.py:17-52), q -> q_out [M, Hq, D], k / v
                             appended to the cache rows (update_kv_cache, cache.py:53-65); y is not written.
                             N_total == (Hq + 2 Hkv) * D, D in {64, 128}                                  */
};
/* Operands of UMB_EPI_ROPE_QKV (same meaning as umb_rope_append_f16's). */
typedef struct {
    void* q_out;
    void* k_cache;
    void* v_cache;
    const void* cos;
    const void* sin;
    const int64_t* position_ids;   /* n_dev == NULL: positions of the M rows; else the base array               */
    const int32_t* n_dev;          /* optional device int32 n: rows are tokens n + row_offset + m, appended at
                                      cache rows n + kv_offset + m (graph replay)                               */
    int row_offset;
    int kv_offset;                 /* n_dev == NULL: absolute append position                                   */
    int num_q_heads, num_kv_heads, head_dim;
    int64_t kv_tok_stride, kv_head_stride;
} umb_rope_epilogue_t;
typedef struct {
    int mode;
    const void* bias;            /* fp16 [N] or NULL                                              */
    const void* residual;        /* fp16 [M, ld_residual] (mode 1); may alias y                   */
    int64_t ld_residual;
    unsigned long long* argmax;  /* mode 3: [M] keys, (orderable(fp16 logit) << 32) | ~index; caller
                                    zero-fills; NULL = no arg-max                                  */
    const umb_rope_epilogue_t* rope;   /* mode 4 only (read at launch time)                          */
} umb_epilogue_t;

size_t umb_gemm_workspace_bytes(int M, int K, int N_total);

/* x fp16 [M, K] row-major (leading dimension ldx elements), y fp16 [M, N_out] (ldy elements),
 * 1 <= M <= 128 per call.  `workspace` must be zero-filled once after allocation; the kernel
 * leaves it clean.  N_total % 8 == 0, K % 128 == 0, group == 128. */
int umb_awq_gemm_f16(const void* x, int64_t ldx, const void* packed, void* y, int64_t ldy, int M, int K,
                     int N_total, int group, const umb_epilogue_t* epilogue, void* workspace, size_t workspace_bytes,
                     umb_stream_t stream);

/* fp16 weights W [N, K] row-major (torch Linear layout): y[M,N] = x . W^T.  Same contract. */
int umb_gemm_f16(const void* x, int64_t ldx, const void* w, int64_t ldw, void* y, int64_t ldy, int M, int K, int N,
                 const umb_epilogue_t* epilogue, void* workspace, size_t workspace_bytes, umb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (ii) tree-masked GQA attention over the KV cache
 * ---------------------------------------------------------------------------------------------- */
enum {
    UMB_MASK_TREE = 0,   /* row i sees [0, prefix_len) and prefix_len + j for every set bit j of
                            tree_bits[i, :] (u32 words, bit j%32 of word j/32), j < kv_len - prefix_len.
                            The rows may be a slice of the tree (one level of the draft's build_tree):
                            pass tree_bits + first_row * tree_words                               */
    UMB_MASK_DENSE = 1,  /* row i sees j iff dense_mask[i * mask_stride + j] != 0  (bool / u8)     */
    UMB_MASK_CAUSAL = 2  /* row i sees j <= prefix_len + i   (prefill / append)                    */
};
typedef struct {
    int mask_mode;
    int prefix_len;                /* committed prefix n (modes 0, 2)                              */
    const int32_t* dyn;            /* optional device int32* n (the committed prefix length, e.g. the
                                      state word umb_tree_accept maintains): prefix_len and kv_len
                                      are then OFFSETS added to n[0] at run time, so one captured
                                      CUDA graph serves a moving prefix                             */
    const uint32_t* tree_bits;     /* [qo, tree_words]                                             */
    int tree_words;
    const uint8_t* dense_mask;
    int64_t mask_stride;
    int kv_len_bound;              /* with dyn: upper bound of the run-time kv_len (grid sizing)   */
} umb_attn_mask_t;

size_t umb_tree_attn_workspace_bytes(int qo, int num_q_heads, int head_dim, int max_kv_len);

/* q fp16 [qo, Hq, D]; k/v: base of this layer's cache, token stride / head stride in elements
 * ([S,Hkv,D]: tok = Hkv*D, head = D;  [Hkv,S,D]: tok = D, head = S*D); out fp16 [qo, Hq, D].
 * kv_len = number of valid cache rows (after the append).  sm_scale <= 0 -> 1/sqrt(D).
 * `workspace` (umb_tree_attn_workspace_bytes) holds the split-KV partials; without it the KV range
 * is not split across CTAs. */
int umb_tree_attn_f16(const void* q, const void* k_cache, const void* v_cache, void* out, int qo, int kv_len,
                      int num_q_heads, int num_kv_heads, int head_dim, int64_t kv_tok_stride, int64_t kv_head_stride,
                      const umb_attn_mask_t* mask, float sm_scale, float soft_cap, void* workspace,
                      size_t workspace_bytes, umb_stream_t stream);

/* RoPE (HF half-split, fp16 arithmetic, fp16 tables [S, D]) on q and k, then append k (rotated) and
 * v at cache rows [kv_offset, kv_offset + qo).  qkv fp16 [qo, (Hq + 2 Hkv) * D] (q | k | v);
 * q_out fp16 [qo, Hq, D].  kv_offset_dev (optional device int32) overrides kv_offset. */
int umb_rope_append_f16(const void* qkv, int64_t ld_qkv, void* q_out, void* k_cache, void* v_cache, const void* cos,
                        const void* sin, const int64_t* position_ids, int qo, int num_q_heads, int num_kv_heads,
                        int head_dim, int64_t kv_tok_stride, int64_t kv_head_stride, int kv_offset,
                        const int32_t* kv_offset_dev, umb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (iii) tree accept + KV gather (integer, bit-exact with the reference)
 * ---------------------------------------------------------------------------------------------- */
/* In:  sampled int64[N] (target token after each node), tokens int64[S], position_ids int64[S],
 *      parents int32[N] (parents[0] = 0, parent < child), depth int32[N], num_nodes (or
 *      state[0] if state != NULL), capacity = elements of tokens / position_ids (no access at or past
 *      it: a prefix that leaves no room for the tree is clamped instead of corrupting memory; the
 *      engines refuse such a state on the host), eos int64[n_eos].
 * Out: tokens[n : n+len+1], position_ids updated exactly as the reference's verify() does;
 *      result int32[4 + N]: {accept_len, continue, new_num_nodes, bonus_lo32, path_abs[0..len)}
 *      (device or mapped-pinned memory); if state != NULL (device int32[4]): state[0] = new_num_nodes,
 *      state[1] = new_num_nodes + N, state[2] = the num_nodes this call started from (the gather
 *      offset: pass &state[2] as umb_kv_gather's offset_dev). */
int umb_tree_accept(const int64_t* sampled, int64_t* tokens, int64_t* position_ids, const int32_t* parents,
                    const int32_t* depth, int N, int num_nodes, int capacity, int32_t* state, const int64_t* eos,
                    int n_eos, int32_t* result, umb_stream_t stream);

/* cache[plane, offset + j, :] = cache[plane, path[j], :] for j < len, for every plane
 * (plane = layer for [L,S,Hkv,D]; layer*head for [L,Hkv,S,D]); K and V in one launch.
 * path/len are read from `result` as written by umb_tree_accept (device memory).  Rows
 * [offset + len, zero_upto) are zero-filled when zero_upto > 0 (the reference zero-fills the
 * whole tail, cache.py:46-47). */
int umb_kv_gather(void* k_cache, void* v_cache, int num_planes, int64_t plane_stride_bytes, int64_t row_stride_bytes,
                  int row_bytes, const int32_t* result, int offset, const int32_t* offset_dev, int zero_upto,
                  umb_stream_t stream);

/* One level of the dynamic (SpecExec-style) tree builder -- DynamicSpeculationEngine.build_tree,
 * reference umbrella/speculation/dynamic_speculation_engine.py:231-248 -- after the draft's top-`beams`
 * logits of the level's `rows` nodes (tree ids [level_first, level_first + rows)) have been selected:
 *   cand[r, b] = tree_score[level_first + r] + log(softmax(topk_logits[r, :])[b] + 1e-4)   (fp32)
 *   the `width` best candidates, best first, become nodes level_first + rows + j:
 *   tree_score[id], tokens[n + id] = topk_ids[flat], parents[id] = level_first + flat / beams,
 *   tree_bits[id, :] = tree_bits[parent, :] | bit id.
 * n = n_dev[0] if n_dev != NULL (device int32) else num_nodes.  rows <= 128, rows * beams <= 4096. */
int umb_dyn_tree_expand(const float* topk_logits, const int64_t* topk_ids, int rows, int beams, int width,
                        int level_first, int num_nodes, int capacity, const int32_t* n_dev, float* tree_score, int64_t* tokens,
                        int32_t* parents, uint32_t* tree_bits, int tree_words, umb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * glue
 * ---------------------------------------------------------------------------------------------- */
int umb_rmsnorm_f16(const void* x, const void* weight, void* y, int rows, int hidden, float eps, umb_stream_t stream);
int umb_silu_mul_f16(const void* gate, const void* up, void* y, int64_t n, umb_stream_t stream);
int umb_add_f16(const void* a, const void* b, void* y, int64_t n, umb_stream_t stream);
int umb_embedding_f16(const int64_t* ids, const void* table, void* y, int rows, int hidden, umb_stream_t stream);
/* first-occurrence arg-max over each row of fp32 logits [rows, vocab] -> int64[rows] */
int umb_argmax_f32(const float* logits, int64_t ld, int rows, int vocab, int64_t* out, umb_stream_t stream);
/* Fused target-side sampling of verify() over fp32 logits [rows, vocab] (one token per tree node):
 * repetition penalty -> top-k -> softmax(T) -> top-p -> draw, or the arg-max of the penalised logits.
 *   UMB_SAMPLE_STATIC   static_speculation_engine.py:305-310: flashinfer top_k_top_p_sampling_from_logits(
 *                       logits / T, uniform_samples[rounds, rows], top_k, top_p) -- top-k mask, softmax, then
 *                       top-p rejection rounds driven by the caller's FIXED uniforms (deterministic)
 *   UMB_SAMPLE_DYNAMIC  dynamic_speculation_engine.py:276-281: apply_topk, softmax(logits / T), top_p_renorm_prob,
 *                       then the inverse CDF (ascending token id) at uniforms[row] in place of torch.multinomial
 *   UMB_SAMPLE_ARGMAX   T < 0.05 with repetition_penalty > 1.01 (static...:298-307): first-occurrence arg-max
 * penalty_ids: the token ids seen so far (speculation_utils.py:340-345; shared by all rows), count =
 * penalty_count (+ penalty_count_dev[0] when that device pointer is given); ignored unless penalty > 1.01.
 * top_k <= 1024.  out int64[rows]. */
enum {
    UMB_SAMPLE_STATIC = 0,
    UMB_SAMPLE_DYNAMIC = 1,
    UMB_SAMPLE_ARGMAX = 2
};
int umb_sample_f32(const float* logits, int64_t ld, int rows, int vocab, int mode, float temperature, int top_k,
                   float top_p, const float* uniforms, int rounds, const int64_t* penalty_ids, int penalty_count,
                   const int32_t* penalty_count_dev, float penalty, int64_t* out, umb_stream_t stream);

/* Row-wise top-k of fp32 logits [rows, vocab], sorted by value descending (lower token id first among equal values):
 * out_vals fp32 [rows, k], out_ids int64 [rows, k].  Replaces ``logits.topk(k)`` on the draft side
 * (speculation_utils.py:52-61 sampling_argmax_gather, dynamic_speculation_engine.py:235).  k <= 64. */
int umb_topk_f32(const float* logits, int64_t ld, int rows, int vocab, int k, float* out_vals, int64_t* out_ids,
                 umb_stream_t stream);

/* decode the UMB_EPI_LOGITS arg-max keys into int64 token ids */
int umb_argmax_keys_to_ids(const unsigned long long* keys, int rows, int64_t* out, umb_stream_t stream);

/* Debug probes (tests only): one 128x16x64 UMMA with A from shared memory (mode 0) / tensor memory (mode 1); modes 2, 3:
 * b is fp16 [64 k, 128 n] (MN-major B operand), d fp32 [128, 128] -- the two readings of the operand descriptor. */
int umb_probe_umma(const void* a /*fp16 [128,64]*/, const void* b /*fp16 [16,64]*/, float* d /*[128,16]*/, int mode,
                   umb_stream_t stream);


/* Debug (tests / profiling only): with UMB_GEMM_TRACE=1 in the environment every GEMM launch records 16
 * %globaltimer stamps per CTA (entry, set-up done, weight copies issued, dependency resolved, first record
 * landed, first MMA, last MMA, first accumulator ready, epilogue done, exit, ...); this call synchronises the
 * device and copies the stamps of the most recent launch, [max_ctas][16], to host memory. */
int umb_debug_gemm_trace(unsigned long long* host_out, int max_ctas);

/* Cross-check of the packed AWQ stream on CUDA cores (tests only; not a product path). */
int umb_awq_gemm_f16_simt(const void* x, int64_t ldx, const void* packed, void* y, int64_t ldy, int M, int K, int N_total,
                          umb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * native forward driver: one call enqueues the whole Llama forward of LlamaAwq.inference /
 * Llama.inference (reference umbrella/models/llama.py:305-322, 117-134; layer_compute :262-303,
 * :75-114) on `stream` -- embedding, L x {RMSNorm, fused q|k|v GEMM, RoPE + KV append, tree
 * attention, o-proj GEMM (+residual), RMSNorm, fused gate|up GEMM (SiLU*up epilogue), down GEMM
 * (+residual)}, final RMSNorm, lm_head GEMM (fp32 logits, optional arg-max keys).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const void* input_norm;  /* fp16 [H]                                                             */
    const void* post_norm;   /* fp16 [H]                                                             */
    const void* wqkv;        /* awq: packed stream, rows q|k|v; fp16: [(Hq+2Hkv)*D, H] row-major      */
    const void* wo;          /* awq: packed stream [H rows, K = Hq*D]; fp16: [H, Hq*D]                */
    const void* wgu;         /* rows interleaved 64 gate | 64 up per 128-row tile, 2*I rows total     */
    const void* wdown;       /* [H rows, K = I]                                                       */
} umb_llama_layer_t;

typedef struct {
    int hidden_size, intermediate_size, num_layers, num_q_heads, num_kv_heads, head_dim, vocab;
    float rms_eps;
    int awq;                 /* 1 = layer weights are AWQ packed streams, 0 = fp16 matrices           */
    int max_length;          /* S: rows of the KV cache and of the cos/sin tables                     */
    const umb_llama_layer_t* layers; /* host array [num_layers]                                       */
    const void* embed;       /* fp16 [V, H]                                                           */
    const void* final_norm;  /* fp16 [H]                                                              */
    const void* lm_head;     /* fp16 [V, H] (never quantised, llama.py:321)                           */
    const void* cos;         /* fp16 [S, D]                                                           */
    const void* sin;         /* fp16 [S, D]                                                           */
    void* k_cache;           /* fp16 [L, S, Hkv, D]  (KV_Cache layout, cache.py:18-34)                */
    void* v_cache;
    /* activations for up to 128 rows: hidden [128,H], normed [128,H], qkv [128,(Hq+2Hkv)*D],
     * q_rot [128,Hq*D], attn_out [128,Hq*D], act [128,I]                                             */
    void* hidden;
    void* normed;
    void* qkv;
    void* q_rot;
    void* attn_out;
    void* act;
    void* gemm_ws;           /* zero-filled once by the caller; >= umb_llama_gemm_ws_bytes()          */
    size_t gemm_ws_bytes;
    void* attn_ws;
    size_t attn_ws_bytes;
} umb_llama_t;

size_t umb_llama_gemm_ws_bytes(int hidden, int intermediate, int num_q_heads, int num_kv_heads, int head_dim,
                               int vocab);

/* rows tokens appended at cache rows [kv_offset, kv_offset + rows) (KV_Cache.update_kv_cache appends,
 * cache.py:53-65).  rows > 128 are processed in chunks of 128 (chunked prefill); `mask` describes
 * the rows of the whole call (dense: row r of the call = mask row r; causal/tree: prefix_len is the
 * number of cache rows visible to every row of the call, normally kv_offset).
 * logits: fp32 [rows - logits_from_row, vocab] for call rows >= logits_from_row, or NULL.
 * argmax_keys: optional [rows - logits_from_row] u64 keys (zero-filled by the caller). */
int umb_llama_forward(const umb_llama_t* model, const int64_t* input_ids, const int64_t* position_ids, int rows,
                      int kv_offset, const umb_attn_mask_t* mask, float* logits, int logits_from_row,
                      unsigned long long* argmax_keys, umb_stream_t stream);

/* Layer-range slice of umb_llama_forward for paged (host-offloaded) weights -- the seam of the
 * reference's LlamaOffload.inference loop (umbrella/models/llama.py:196-219): the caller runs layers
 * [layer_begin, layer_end) of rows <= 128 (hidden state kept in model->hidden between calls), waiting on
 * its own weight-copy events in between.  do_embed != 0 starts from the embedding (first slice);
 * logits != NULL finishes with the final norm + lm_head (last slice; layer_begin == layer_end ==
 * num_layers runs the head alone). */
int umb_llama_forward_range(const umb_llama_t* model, const int64_t* input_ids, const int64_t* position_ids, int rows,
                            int kv_offset, const umb_attn_mask_t* mask, int layer_begin, int layer_end, int do_embed,
                            float* logits, int logits_from_row, unsigned long long* argmax_keys, umb_stream_t stream);

/* Graph-replay form of the tree forward: every position-dependent quantity is read on the device
 * from n_dev[0] (the committed prefix length n, maintained by umb_tree_accept through `state`):
 * the rows are tokens_base[n + row_offset + r] / position_base[n + row_offset + r], they are
 * appended at cache rows n + row_offset + r, and row r sees [0, n) plus the set bits of
 * tree_bits[r, :] (offset n).  row_offset = -1 starts one row before the tree (the last committed token): the
 * prefix is then n - 1 and the bits are relative to it.  kv_len_bound >= n + row_offset + rows sizes the attention grid.
 * rows <= 128.  One captured CUDA graph of this call is valid until n exceeds the bound. */
int umb_llama_forward_dyn(const umb_llama_t* model, const int64_t* tokens_base, const int64_t* position_base,
                          int rows, const int32_t* n_dev, int row_offset, int kv_len_bound,
                          const uint32_t* tree_bits, int tree_words, float* logits, unsigned long long* argmax_keys,
                          umb_stream_t stream);

/* tokens_base[n_dev[0] + offset + i] = src[i], i < count  (draft tokens of the next tree level) */
int umb_store_tokens(int64_t* tokens_base, const int32_t* n_dev, int offset, const int64_t* src, int count,
                     int capacity, umb_stream_t stream);
/* same with a gather: tokens_base[n + offset + i] = src[gather[i]]  (n = n_dev[0], or num_nodes when n_dev is NULL):
 * the growmap-ordered pick of the per-node top-k draft tokens (static_speculation_engine.py:115-123, 279-281) */
int umb_store_tokens_gather(int64_t* tokens_base, const int32_t* n_dev, int num_nodes, int offset, const int64_t* src,
                            const int64_t* gather, int count, int capacity, umb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UMBRELLA_B200_H */
