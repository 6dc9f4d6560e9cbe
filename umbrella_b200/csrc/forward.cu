// Native forward driver: enqueues the whole Llama verify forward from one C call (no Python between
// kernels).  Mirrors LlamaAwq.inference / layer_compute (reference umbrella/models/llama.py:262-322)
// and Llama.inference (llama.py:75-134) op for op; every fp16 rounding point of the reference is a
// kernel boundary or an explicit rounding inside a fused epilogue here.
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace umb {

static int linear(const umb_llama_t* m, const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows,
                  int K, int N, const umb_epilogue_t* epi, umb_stream_t st) {
    if (m->awq)
        return umb_awq_gemm_f16(x, ldx, w, y, ldy, rows, K, N, 128, epi, m->gemm_ws, m->gemm_ws_bytes, st);
    return umb_gemm_f16(x, ldx, w, K, y, ldy, rows, K, N, epi, m->gemm_ws, m->gemm_ws_bytes, st);
}

#define UMB_TRY(expr) do { int rc__ = (expr); if (rc__ != UMB_OK) return rc__; } while (0)

constexpr int kChunkRows = UMB_MAX_ROWS;

// UMB_FUSE_ROPE=1 runs the q|k|v projection with the fused RoPE + KV-append epilogue (UMB_EPI_ROPE_QKV) instead of
// the GEMM + rope_append kernel pair.  Off by default: bit-identical, one launch fewer, but measured slower
// (15.47 vs 14.91 ms per iteration) -- the rotation then sits on the serial tail of a 148-CTA persistent kernel,
// while the stand-alone kernel is a small parallel launch that programmatic dependent launch overlaps.
static bool fuse_rope() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("UMB_FUSE_ROPE");
        cached = (e && e[0] == '1') ? 1 : 0;
    }
    return cached != 0;
}

// n_dev == NULL: ids / pos point at the rows, kv_offset is the append position (host value).
// n_dev != NULL: ids / pos are base pointers, kv_offset is the row offset added to n_dev[0] on the device.
static int forward_chunk(const umb_llama_t* m, const int64_t* ids, const int64_t* pos, int rows, int kv_offset,
                         const umb_attn_mask_t* mask, float* logits, int logits_from, unsigned long long* keys,
                         umb_stream_t st, const int32_t* n_dev = nullptr, int layer_begin = 0, int layer_end = -1,
                         bool do_embed = true) {
    const int H = m->hidden_size, I = m->intermediate_size, Hq = m->num_q_heads, Hkv = m->num_kv_heads, D = m->head_dim;
    const int qkv_n = (Hq + 2 * Hkv) * D;
    const int64_t tok_stride = (int64_t)Hkv * D;
    const int64_t layer_elems = (int64_t)m->max_length * tok_stride;
    if (layer_end < 0) layer_end = m->num_layers;
    if (!do_embed) {
        // a later slice of a layer-by-layer (paged) forward: the hidden rows are already in m->hidden
    } else if (n_dev)
        UMB_TRY(embedding_dyn(ids, n_dev, kv_offset, m->embed, m->hidden, rows, H, (cudaStream_t)st));
    else
        UMB_TRY(umb_embedding_f16(ids, m->embed, m->hidden, rows, H, st));
    for (int l = layer_begin; l < layer_end; ++l) {
        const umb_llama_layer_t& L = m->layers[l];
        __half* kc = (__half*)m->k_cache + l * layer_elems;
        __half* vc = (__half*)m->v_cache + l * layer_elems;
        UMB_TRY(umb_rmsnorm_f16(m->hidden, L.input_norm, m->normed, rows, H, m->rms_eps, st));
        if (fuse_rope() && (D == 64 || D == 128)) {
            // q|k|v projection + RoPE + KV append in one kernel (UMB_EPI_ROPE_QKV)
            umb_rope_epilogue_t rp{};
            rp.q_out = m->q_rot;
            rp.k_cache = kc;
            rp.v_cache = vc;
            rp.cos = m->cos;
            rp.sin = m->sin;
            rp.position_ids = pos;
            rp.n_dev = n_dev;
            rp.row_offset = kv_offset;
            rp.kv_offset = kv_offset;
            rp.num_q_heads = Hq;
            rp.num_kv_heads = Hkv;
            rp.head_dim = D;
            rp.kv_tok_stride = tok_stride;
            rp.kv_head_stride = D;
            umb_epilogue_t qe{};
            qe.mode = UMB_EPI_ROPE_QKV;
            qe.rope = &rp;
            UMB_TRY(linear(m, m->normed, H, L.wqkv, m->qkv, qkv_n, rows, H, qkv_n, &qe, st));
        } else {
        UMB_TRY(linear(m, m->normed, H, L.wqkv, m->qkv, qkv_n, rows, H, qkv_n, nullptr, st));
        if (n_dev)
            UMB_TRY(rope_append_dyn(m->qkv, qkv_n, m->q_rot, kc, vc, m->cos, m->sin, pos, n_dev, kv_offset, rows, Hq, Hkv,
                                    D, tok_stride, D, (cudaStream_t)st));
        else
            UMB_TRY(umb_rope_append_f16(m->qkv, qkv_n, m->q_rot, kc, vc, m->cos, m->sin, pos, rows, Hq, Hkv, D,
                                        tok_stride, D, kv_offset, nullptr, st));
        }
        UMB_TRY(umb_tree_attn_f16(m->q_rot, kc, vc, m->attn_out, rows, kv_offset + rows, Hq, Hkv, D, tok_stride, D,
                                  mask, 0.f, 0.f, m->attn_ws, m->attn_ws_bytes, st));
        umb_epilogue_t res{};
        res.mode = UMB_EPI_RESIDUAL;
        res.residual = m->hidden;
        res.ld_residual = H;
        UMB_TRY(linear(m, m->attn_out, (int64_t)Hq * D, L.wo, m->hidden, H, rows, Hq * D, H, &res, st));
        UMB_TRY(umb_rmsnorm_f16(m->hidden, L.post_norm, m->normed, rows, H, m->rms_eps, st));
        umb_epilogue_t silu{};
        silu.mode = UMB_EPI_SILU_MUL;
        UMB_TRY(linear(m, m->normed, H, L.wgu, m->act, I, rows, H, 2 * I, &silu, st));
        UMB_TRY(linear(m, m->act, I, L.wdown, m->hidden, H, rows, I, H, &res, st));
    }
    if (logits || keys) {
        const int from = logits_from < 0 ? 0 : logits_from;
        if (from < rows) {
            const int n = rows - from;
            UMB_TRY(umb_rmsnorm_f16((const __half*)m->hidden + (int64_t)from * H, m->final_norm, m->normed, n, H,
                                    m->rms_eps, st));
            umb_epilogue_t lg{};
            lg.mode = UMB_EPI_LOGITS;
            lg.argmax = keys;
            // logits == NULL with keys only is not supported by the epilogue: the fp32 rows are always written
            UMB_TRY(umb_gemm_f16(m->normed, H, m->lm_head, H, logits, m->vocab, n, H, m->vocab, &lg, m->gemm_ws,
                                 m->gemm_ws_bytes, st));
        }
    }
    return UMB_OK;
}

}  // namespace umb

extern "C" size_t umb_llama_gemm_ws_bytes(int hidden, int intermediate, int num_q_heads, int num_kv_heads, int head_dim,
                                          int vocab) {
    size_t best = 0;
    const int ns[5] = {(num_q_heads + 2 * num_kv_heads) * head_dim, hidden, 2 * intermediate, hidden, vocab};
    for (int i = 0; i < 5; ++i) {
        const size_t b = umb_gemm_workspace_bytes(128, hidden, ns[i]);
        best = b > best ? b : best;
    }
    return best;
}

extern "C" int umb_llama_forward(const umb_llama_t* m, const int64_t* input_ids, const int64_t* position_ids, int rows,
                                 int kv_offset, const umb_attn_mask_t* mask, float* logits, int logits_from_row,
                                 unsigned long long* argmax_keys, umb_stream_t stream) {
    UMB_CHECK_ARG(m && input_ids && position_ids && mask && rows >= 1 && kv_offset >= 0);
    UMB_CHECK_ARG(m->layers && m->embed && m->final_norm && m->cos && m->sin && m->k_cache && m->v_cache);
    UMB_CHECK_ARG(m->hidden && m->normed && m->qkv && m->q_rot && m->attn_out && m->act && m->gemm_ws);
    UMB_CHECK_ARG(kv_offset + rows <= m->max_length);
    UMB_CHECK_ARG(!argmax_keys || logits);
    if (logits) UMB_CHECK_ARG(m->lm_head != nullptr);
    // rows > 128 (prefill, dynamic trees of 257..505 nodes): equal chunks of <= 128 rows, each a full pass
    // over the weights; later chunks see the KV rows the earlier ones appended.
    const int n_chunks = (rows + umb::kChunkRows - 1) / umb::kChunkRows;
    const int chunk = (rows + n_chunks - 1) / n_chunks;
    for (int r0 = 0; r0 < rows; r0 += chunk) {
        const int n = rows - r0 < chunk ? rows - r0 : chunk;
        umb_attn_mask_t mk = *mask;
        if (mk.mask_mode == UMB_MASK_DENSE) {
            mk.dense_mask = mask->dense_mask + (int64_t)r0 * mask->mask_stride;
        } else if (mk.mask_mode == UMB_MASK_CAUSAL) {
            mk.prefix_len = mask->prefix_len + r0;
        } else {
            mk.tree_bits = mask->tree_bits + (int64_t)r0 * mask->tree_words;   // rows r0.. of the ancestor bitmask
        }
        // logits rows of this chunk
        float* lg = nullptr;
        unsigned long long* keys = nullptr;
        int from = 0;
        if (logits && logits_from_row < r0 + n) {
            from = logits_from_row > r0 ? logits_from_row - r0 : 0;
            const int64_t out_row = (int64_t)(r0 + from) - logits_from_row;
            lg = logits + out_row * m->vocab;
            keys = argmax_keys ? argmax_keys + out_row : nullptr;
        }
        const int rc = umb::forward_chunk(m, input_ids + r0, position_ids + r0, n, kv_offset + r0, &mk, lg, from, keys,
                                          stream);
        if (rc != UMB_OK) return rc;
    }
    return UMB_OK;
}

extern "C" int umb_llama_forward_range(const umb_llama_t* m, const int64_t* input_ids, const int64_t* position_ids,
                                       int rows, int kv_offset, const umb_attn_mask_t* mask, int layer_begin,
                                       int layer_end, int do_embed, float* logits, int logits_from_row,
                                       unsigned long long* argmax_keys, umb_stream_t stream) {
    UMB_CHECK_ARG(m && input_ids && position_ids && mask && rows >= 1 && rows <= 128 && kv_offset >= 0);
    UMB_CHECK_ARG(m->layers && m->embed && m->final_norm && m->cos && m->sin && m->k_cache && m->v_cache);
    UMB_CHECK_ARG(m->hidden && m->normed && m->qkv && m->q_rot && m->attn_out && m->act && m->gemm_ws);
    UMB_CHECK_ARG(kv_offset + rows <= m->max_length);
    UMB_CHECK_ARG(layer_begin >= 0 && layer_begin <= layer_end && layer_end <= m->num_layers);
    UMB_CHECK_ARG(!argmax_keys || logits);
    if (logits) UMB_CHECK_ARG(m->lm_head != nullptr && logits_from_row >= 0 && logits_from_row < rows);
    return umb::forward_chunk(m, input_ids, position_ids, rows, kv_offset, mask, logits, logits_from_row, argmax_keys,
                              stream, nullptr, layer_begin, layer_end, do_embed != 0);
}

extern "C" int umb_llama_forward_dyn(const umb_llama_t* m, const int64_t* tokens_base, const int64_t* position_base,
                                     int rows, const int32_t* n_dev, int row_offset, int kv_len_bound,
                                     const uint32_t* tree_bits, int tree_words, float* logits,
                                     unsigned long long* argmax_keys, umb_stream_t stream) {
    UMB_CHECK_ARG(m && tokens_base && position_base && n_dev && tree_bits && rows >= 1 && rows <= 128);
    // row_offset = -1: the rows start at the last committed token (n - 1), which then counts as prefix n - 1 --
    // the draft's first step re-derives that token's K/V together with the root (lazily computed leaf rows)
    UMB_CHECK_ARG(row_offset >= -1 && tree_words >= 1 && kv_len_bound >= row_offset + rows && kv_len_bound <= m->max_length);
    UMB_CHECK_ARG(m->layers && m->embed && m->final_norm && m->cos && m->sin && m->k_cache && m->v_cache);
    UMB_CHECK_ARG(m->hidden && m->normed && m->qkv && m->q_rot && m->attn_out && m->act && m->gemm_ws);
    UMB_CHECK_ARG(!argmax_keys || logits);
    umb_attn_mask_t mk{};
    mk.mask_mode = UMB_MASK_TREE;
    mk.prefix_len = row_offset < 0 ? row_offset : 0;   // offsets relative to n_dev[0]
    mk.dyn = n_dev;
    mk.tree_bits = tree_bits;
    mk.tree_words = tree_words;
    mk.kv_len_bound = kv_len_bound;
    return umb::forward_chunk(m, tokens_base, position_base, rows, row_offset, &mk, logits, 0, argmax_keys, stream, n_dev);
}
