// (ii) tree-masked grouped-query attention over the KV cache.
//
// Replaces flashinfer.single_prefill_with_kv_cache(custom_mask=...) as called from
// KV_Cache.compute_attention (reference umbrella/attn/cache.py:67-87) and the torch path of
// StaticKV_Cache.compute_attention (cache.py:169-192).
//
// Shape of the problem at verify time: qo = tree size (13..505) query tokens, Hq/Hkv = 8 (or 4)
// query heads share each KV head, kv = prefix + tree rows.  Every KV row is therefore used by
// G * qo query rows: the kernel flattens (token, head-in-group) into one row index per KV head so
// that a 64-row CTA tile re-uses each K/V row 64 times from shared memory, streams K/V exactly once
// per (row tile, split) with 16-byte cp.async loads (coalesced, L1-bypassing), runs QK^T and PV on
// mma.sync m16n8k16 (fp16 in, fp32 accumulate) and keeps the softmax in fp32 with quad shuffles.
// The mask is never materialised for the tree case: rows see [0, prefix_len) unconditionally and
// prefix_len + j iff bit j of their ancestor bitmask is set.  Long contexts are split along KV
// (grid.z) and merged by a second kernel with a log-sum-exp reduction.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace umb {

struct AttnParams {
    const __half* q;
    const __half* k;
    const __half* v;
    __half* out;
    int qo, kv_len, Hq, Hkv, G;
    int64_t kv_tok_stride, kv_head_stride;
    int mask_mode, prefix_len;
    const int32_t* dyn;
    const uint32_t* tree_bits;
    int tree_words;
    const uint8_t* dense_mask;
    int64_t mask_stride;
    float scale_log2;   // sm_scale * log2(e)
    float soft_cap;     // 0 = off
    float sm_scale;
    int splits;
    float* part_o;      // [splits][Hkv * qo * G][D]
    float2* part_ml;    // [splits][Hkv * qo * G]  (running max in the log2 domain, running sum)
    int* counters;      // [Hkv * row_tiles] arrival counters of the split CTAs (zero between launches)
};

constexpr int kAttnThreads = 128;
constexpr int kBM = 64;   // query rows per CTA (16 per warp)
constexpr int kBN = 64;   // kv rows per tile

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// smem tile: kBN rows x D halves, 16-byte chunks XOR-swizzled by (row & 7)
template <int D>
__device__ __forceinline__ __half* tile_ptr(__half* base, int row, int chunk) {
    return base + row * D + ((chunk ^ (row & 7)) << 3);
}

template <int D>
__device__ __forceinline__ void load_tile(__half* smem, const __half* gbase, int64_t tok_stride, int row0, int kv_hi) {
    constexpr int kChunks = D / 8;
    for (int i = threadIdx.x; i < kBN * kChunks; i += kAttnThreads) {
        const int r = i / kChunks, c = i % kChunks;
        const int row = row0 + r;
        const bool ok = row < kv_hi;
        cp_async16(tile_ptr<D>(smem, r, c), gbase + (int64_t)(ok ? row : 0) * tok_stride + c * 8, ok);
    }
}

// 64-column visibility window of one query row for the kv tile starting at col0 (bit i = column col0 + i).
// Built once per (row, tile) -- the per-element work is then a constant-position bit test.
__device__ __forceinline__ unsigned long long row_window(const AttnParams& p, int tok, bool row_ok, int col0, int kv_hi,
                                                         int prefix_len) {
    if (!row_ok) return 0ull;
    const int n_valid = min(kBN, kv_hi - col0);                        // columns of this tile that exist
    const unsigned long long exist = n_valid >= 64 ? ~0ull : ((1ull << n_valid) - 1ull);
    unsigned long long allow;
    if (p.mask_mode == UMB_MASK_CAUSAL) {
        const int n = prefix_len + tok + 1 - col0;                     // columns j <= prefix_len + tok
        allow = n >= 64 ? ~0ull : (n <= 0 ? 0ull : ((1ull << n) - 1ull));
    } else if (p.mask_mode == UMB_MASK_TREE) {
        const int np = prefix_len - col0;                              // leading prefix columns: always visible
        allow = np >= 64 ? ~0ull : (np <= 0 ? 0ull : ((1ull << np) - 1ull));
        if (np < 64) {
            const uint32_t* bits = p.tree_bits + (int64_t)tok * p.tree_words;
            const int start = -np;                                     // tree bit index of column col0 (may be < 0)
#pragma unroll 1
            for (int w = 0; w < p.tree_words; ++w) {                   // place tree word w at columns [32w - start, +32)
                const int pos = 32 * w - start;
                if (pos <= -32 || pos >= 64) continue;
                const unsigned long long v = bits[w];
                allow |= pos >= 0 ? (v << pos) : (v >> (-pos));
            }
        }
    } else {
        allow = 0ull;
        const uint8_t* mrow = p.dense_mask + (int64_t)tok * p.mask_stride + col0;
#pragma unroll 1
        for (int i = 0; i < n_valid; ++i) allow |= (unsigned long long)(mrow[i] != 0) << i;
    }
    return allow & exist;
}

template <int D, bool kSoftCap>
__global__ void __launch_bounds__(kAttnThreads) tree_attn_kernel(const AttnParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __half* sK = reinterpret_cast<__half*>(smem_raw);                 // [2][kBN][D]
    __half* sV = sK + 2 * kBN * D;                                     // [2][kBN][D]
    pdl_launch_dependents();
    pdl_wait();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g4 = lane >> 2, t4 = lane & 3;
    const int hkv = blockIdx.y, split = blockIdx.z;
    const int rows_total = p.qo * p.G;
    const int prefix_len = p.dyn ? p.dyn[0] + p.prefix_len : p.prefix_len;
    const int kv_len = p.dyn ? p.dyn[0] + p.kv_len : p.kv_len;

    const int row_base = blockIdx.x * kBM + warp * 16;
    // the two rows this thread owns inside its warp's 16-row slab
    int row[2], tok[2];
    bool row_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        row[i] = row_base + g4 + 8 * i;
        row_ok[i] = row[i] < rows_total;
        tok[i] = row_ok[i] ? row[i] / p.G : 0;
    }

    // ---- Q fragments (A operand), straight from global ----
    uint32_t qf[D / 16][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int hq = hkv * p.G + (row_ok[i] ? row[i] % p.G : 0);
        const __half* qr = p.q + ((int64_t)tok[i] * p.Hq + hq) * D;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t lo = 0, hi = 0;
            if (row_ok[i]) {
                lo = *reinterpret_cast<const uint32_t*>(qr + ks * 16 + t4 * 2);
                hi = *reinterpret_cast<const uint32_t*>(qr + ks * 16 + 8 + t4 * 2);
            }
            qf[ks][i] = lo;
            qf[ks][2 + i] = hi;
        }
    }

    // ---- kv range of this CTA ----
    int kv_hi = kv_len;
    if (p.mask_mode == UMB_MASK_CAUSAL) {
        const int last_row = min(rows_total, (int)(blockIdx.x + 1) * kBM) - 1;
        kv_hi = min(kv_hi, prefix_len + last_row / p.G + 1);
    } else if (p.mask_mode == UMB_MASK_TREE) {
        kv_hi = min(kv_hi, prefix_len + p.tree_words * 32);   // columns the bitmask can describe
    }
    const int n_tiles = (kv_hi + kBN - 1) / kBN;
    const int tiles_per_split = (n_tiles + p.splits - 1) / p.splits;
    const int tile_lo = split * tiles_per_split;
    const int tile_hi = min(n_tiles, tile_lo + tiles_per_split);

    const __half* kg = p.k + (int64_t)hkv * p.kv_head_stride;
    const __half* vg = p.v + (int64_t)hkv * p.kv_head_stride;

    float o_acc[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    if (tile_lo < tile_hi) {
        load_tile<D>(sK, kg, p.kv_tok_stride, tile_lo * kBN, kv_hi);
        load_tile<D>(sV, vg, p.kv_tok_stride, tile_lo * kBN, kv_hi);
        cp_async_commit();
    }
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
        const int buf = (tile - tile_lo) & 1;
        if (tile + 1 < tile_hi) {
            load_tile<D>(sK + (buf ^ 1) * kBN * D, kg, p.kv_tok_stride, (tile + 1) * kBN, kv_hi);
            load_tile<D>(sV + (buf ^ 1) * kBN * D, vg, p.kv_tok_stride, (tile + 1) * kBN, kv_hi);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const __half* tK = sK + buf * kBN * D;
        const __half* tV = sV + buf * kBN * D;

        // ---- S = Q K^T  (16 x 64 per warp) ----
        float s[kBN / 8][4];
#pragma unroll
        for (int i = 0; i < kBN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
            for (int nb = 0; nb < kBN / 8; nb += 2) {
                uint32_t b[4];
                const int r = nb * 8 + (lane & 7) + ((lane >> 4) << 3);
                const int c = ks * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(b, tile_ptr<D>(const_cast<__half*>(tK), r, c));
                mma_16816(s[nb], qf[ks], b[0], b[1]);
                mma_16816(s[nb + 1], qf[ks], b[2], b[3]);
            }
        }

        // ---- soft-cap (Gemma only), mask ----
        const int col0 = tile * kBN;
        if constexpr (kSoftCap) {
#pragma unroll
            for (int nb = 0; nb < kBN / 8; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    s[nb][e] = p.soft_cap * tanhf(s[nb][e] * p.sm_scale / p.soft_cap) / p.sm_scale;   // back in raw units
        }
        const bool need_mask = (col0 + kBN > kv_hi) || (p.mask_mode == UMB_MASK_DENSE) || (col0 + kBN > prefix_len);
        if (need_mask) {
#pragma unroll
            for (int ri = 0; ri < 2; ++ri) {
                const unsigned long long win = row_window(p, tok[ri], row_ok[ri], col0, kv_hi, prefix_len) >> (t4 * 2);
                const uint32_t lo = (uint32_t)win, hi = (uint32_t)(win >> 32);
#pragma unroll
                for (int nb = 0; nb < kBN / 8; ++nb) {
                    const uint32_t wbits = nb < 4 ? lo : hi;
                    const int b = (nb & 3) * 8;
                    if (!((wbits >> b) & 1u)) s[nb][2 * ri] = -INFINITY;
                    if (!((wbits >> (b + 1)) & 1u)) s[nb][2 * ri + 1] = -INFINITY;
                }
            }
        }

        // ---- online softmax (rows g4 and g4 + 8), scores still unscaled: exp2((s - m) * scale_log2) ----
        float corr[2];
#pragma unroll
        for (int ri = 0; ri < 2; ++ri) {
            float mx = -INFINITY;
#pragma unroll
            for (int nb = 0; nb < kBN / 8; ++nb) mx = fmaxf(mx, fmaxf(s[nb][2 * ri], s[nb][2 * ri + 1]));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_new = fmaxf(m_run[ri], mx * p.scale_log2);       // running max lives in the log2 domain
            const float m_use = m_new == -INFINITY ? 0.f : m_new;
            corr[ri] = exp2f(m_run[ri] - m_use);                            // m_run = -inf -> 0
            float sum = 0.f;
#pragma unroll
            for (int nb = 0; nb < kBN / 8; ++nb) {
                const float p0 = exp2f(fmaf(s[nb][2 * ri], p.scale_log2, -m_use));
                const float p1 = exp2f(fmaf(s[nb][2 * ri + 1], p.scale_log2, -m_use));
                s[nb][2 * ri] = p0;
                s[nb][2 * ri + 1] = p1;
                sum += p0 + p1;
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            l_run[ri] = l_run[ri] * corr[ri] + sum;
            m_run[ri] = m_new;
        }
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
            o_acc[i][0] *= corr[0];
            o_acc[i][1] *= corr[0];
            o_acc[i][2] *= corr[1];
            o_acc[i][3] *= corr[1];
        }

        // ---- O += P V ----
#pragma unroll
        for (int ks = 0; ks < kBN / 16; ++ks) {
            uint32_t a[4];
            a[0] = pack_half2(s[2 * ks][0], s[2 * ks][1]);
            a[1] = pack_half2(s[2 * ks][2], s[2 * ks][3]);
            a[2] = pack_half2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
            a[3] = pack_half2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
            for (int db = 0; db < D / 8; db += 2) {
                uint32_t b[4];
                const int r = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int c = db + (lane >> 4);
                ldmatrix_x4_trans(b, tile_ptr<D>(const_cast<__half*>(tV), r, c));
                mma_16816(o_acc[db], a, b[0], b[1]);
                mma_16816(o_acc[db + 1], a, b[2], b[3]);
            }
        }
        __syncthreads();   // the buffer is re-filled two iterations later
    }

    // ---- write-out ----
    if (p.splits == 1) {
#pragma unroll
        for (int ri = 0; ri < 2; ++ri) {
            if (!row_ok[ri]) continue;
            const float inv = l_run[ri] > 0.f ? 1.f / l_run[ri] : 0.f;
            const int hq = hkv * p.G + row[ri] % p.G;
            __half* orow = p.out + ((int64_t)tok[ri] * p.Hq + hq) * D;
#pragma unroll
            for (int db = 0; db < D / 8; ++db)
                *reinterpret_cast<uint32_t*>(orow + db * 8 + t4 * 2) =
                    pack_half2(o_acc[db][2 * ri] * inv, o_acc[db][2 * ri + 1] * inv);
        }
    } else {
#pragma unroll
        for (int ri = 0; ri < 2; ++ri) {
            if (!row_ok[ri]) continue;
            const int64_t grow = (int64_t)split * p.Hkv * rows_total + (int64_t)hkv * rows_total + row[ri];
            float* orow = p.part_o + grow * D;
#pragma unroll
            for (int db = 0; db < D / 8; ++db)
                *reinterpret_cast<float2*>(orow + db * 8 + t4 * 2) = make_float2(o_acc[db][2 * ri], o_acc[db][2 * ri + 1]);
            if (t4 == 0) p.part_ml[grow] = make_float2(m_run[ri], l_run[ri]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Latency-oriented variant for small query counts (tree verify, draft levels, short prefill
// chunks): one CTA per (16-row slab, kv head[, kv split]); its four warps take kv tiles of 32 rows
// round-robin, each with a private double-buffered K/V ring in shared memory (no CTA-wide barrier
// in the loop), and merge their (m, l, o) partials through shared memory at the end.
// ------------------------------------------------------------------------------------------------
constexpr int kSBN = 32;
constexpr int kSlabTilesPerCta = 32;     // minimum kv tiles per slab CTA before the KV range is split across CTAs

template <int D>
__device__ __forceinline__ void load_tile_warp(__half* smem, const __half* gbase, int64_t tok_stride, int row0, int kv_hi,
                                               int lane) {
    constexpr int kChunks = D / 8;
#pragma unroll
    for (int i = lane; i < kSBN * kChunks; i += 32) {
        const int r = i / kChunks, c = i % kChunks;
        const int row = row0 + r;
        const bool ok = row < kv_hi;
        cp_async16(tile_ptr<D>(smem, r, c), gbase + (int64_t)(ok ? row : 0) * tok_stride + c * 8, ok);
    }
}

// 32-column visibility window (bit i = column col0 + i)
__device__ __forceinline__ uint32_t row_window32(const AttnParams& p, int tok, bool row_ok, int col0, int kv_hi,
                                                 int prefix_len) {
    if (!row_ok) return 0u;
    const int n_valid = min(kSBN, kv_hi - col0);
    const uint32_t exist = n_valid >= 32 ? ~0u : ((1u << n_valid) - 1u);
    uint32_t allow;
    if (p.mask_mode == UMB_MASK_CAUSAL) {
        const int n = prefix_len + tok + 1 - col0;
        allow = n >= 32 ? ~0u : (n <= 0 ? 0u : ((1u << n) - 1u));
    } else if (p.mask_mode == UMB_MASK_TREE) {
        const int np = prefix_len - col0;
        allow = np >= 32 ? ~0u : (np <= 0 ? 0u : ((1u << np) - 1u));
        if (np < 32) {
            const uint32_t* bits = p.tree_bits + (int64_t)tok * p.tree_words;
            const int start = -np;                                   // tree bit index of column col0
            const int w0 = start >> 5;                               // floor division (start may be negative)
            const int sh = start & 31;
            const uint32_t lo = (w0 >= 0 && w0 < p.tree_words) ? bits[w0] : 0u;
            const uint32_t hi = (w0 + 1 >= 0 && w0 + 1 < p.tree_words) ? bits[w0 + 1] : 0u;
            allow |= sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
        }
    } else {
        allow = 0u;
        const uint8_t* mrow = p.dense_mask + (int64_t)tok * p.mask_stride + col0;
#pragma unroll 1
        for (int i = 0; i < n_valid; ++i) allow |= (uint32_t)(mrow[i] != 0) << i;
    }
    return allow & exist;
}

// kWarps = 4: every warp owns a double-buffered ring.  kWarps = 8 (short contexts: 2-3 tiles per warp instead of 4-5, the
// K/V loads of the whole range in flight at once): double-buffered while the rings fit in shared memory (D = 64),
// single-buffered otherwise (D = 128: 8 x 16 KB).  The merge is done by the first four warps.
template <int D, int kWarps>
struct SlabCfg {
    static constexpr int kRing = (kWarps * 4 * kSBN * D * (int)sizeof(__half) <= 160 * 1024) ? 2 : 1;
    static constexpr size_t kSmem = (size_t)kWarps * kRing * 2 * kSBN * D * sizeof(__half);
};
template <int D, bool kSoftCap, int kWarps>
__global__ void __launch_bounds__(kWarps * 32) tree_attn_slab_kernel(const AttnParams p) {
    constexpr int kRing = SlabCfg<D, kWarps>::kRing;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g4 = lane >> 2, t4 = lane & 3;
    const int hkv = blockIdx.y, split = blockIdx.z;
    const int rows_total = p.qo * p.G;
    const int prefix_len = p.dyn ? p.dyn[0] + p.prefix_len : p.prefix_len;
    const int kv_len = p.dyn ? p.dyn[0] + p.kv_len : p.kv_len;
    // per-warp ring: [kRing stages][K | V][kSBN][D]
    __half* ring = reinterpret_cast<__half*>(smem_raw) + warp * (kRing * 2 * kSBN * D);

    const int row_base = blockIdx.x * 16;
    int row[2], tok[2];
    bool row_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        row[i] = row_base + g4 + 8 * i;
        row_ok[i] = row[i] < rows_total;
        tok[i] = row_ok[i] ? row[i] / p.G : 0;
    }
    uint32_t qf[D / 16][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int hq = hkv * p.G + (row_ok[i] ? row[i] % p.G : 0);
        const __half* qr = p.q + ((int64_t)tok[i] * p.Hq + hq) * D;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t lo = 0, hi = 0;
            if (row_ok[i]) {
                lo = *reinterpret_cast<const uint32_t*>(qr + ks * 16 + t4 * 2);
                hi = *reinterpret_cast<const uint32_t*>(qr + ks * 16 + 8 + t4 * 2);
            }
            qf[ks][i] = lo;
            qf[ks][2 + i] = hi;
        }
    }

    int kv_hi = kv_len;
    if (p.mask_mode == UMB_MASK_CAUSAL) {
        const int last_row = min(rows_total, row_base + 16) - 1;
        kv_hi = min(kv_hi, prefix_len + last_row / p.G + 1);
    } else if (p.mask_mode == UMB_MASK_TREE) {
        kv_hi = min(kv_hi, prefix_len + p.tree_words * 32);
    }
    const int n_tiles = (kv_hi + kSBN - 1) / kSBN;
    const int tiles_per_split = (n_tiles + p.splits - 1) / p.splits;
    const int tile_lo = split * tiles_per_split;
    const int tile_hi = min(n_tiles, tile_lo + tiles_per_split);
    const __half* kg = p.k + (int64_t)hkv * p.kv_head_stride;
    const __half* vg = p.v + (int64_t)hkv * p.kv_head_stride;

    float o_acc[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    int tile = tile_lo + warp;
    if (tile < tile_hi) {
        load_tile_warp<D>(ring, kg, p.kv_tok_stride, tile * kSBN, kv_hi, lane);
        load_tile_warp<D>(ring + kSBN * D, vg, p.kv_tok_stride, tile * kSBN, kv_hi, lane);
        cp_async_commit();
    }
    for (int it = 0; tile < tile_hi; tile += kWarps, ++it) {
        const int buf = kRing == 2 ? (it & 1) : 0;
        if (kRing == 2 && tile + kWarps < tile_hi) {
            __half* nxt = ring + (buf ^ 1) * (2 * kSBN * D);
            load_tile_warp<D>(nxt, kg, p.kv_tok_stride, (tile + kWarps) * kSBN, kv_hi, lane);
            load_tile_warp<D>(nxt + kSBN * D, vg, p.kv_tok_stride, (tile + kWarps) * kSBN, kv_hi, lane);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        __half* tK = ring + buf * (2 * kSBN * D);
        __half* tV = tK + kSBN * D;

        float s[kSBN / 8][4];
#pragma unroll
        for (int i = 0; i < kSBN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
            for (int nb = 0; nb < kSBN / 8; nb += 2) {
                uint32_t b[4];
                const int r = nb * 8 + (lane & 7) + ((lane >> 4) << 3);
                const int c = ks * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(b, tile_ptr<D>(tK, r, c));
                mma_16816(s[nb], qf[ks], b[0], b[1]);
                mma_16816(s[nb + 1], qf[ks], b[2], b[3]);
            }
        }
        const int col0 = tile * kSBN;
        if constexpr (kSoftCap) {
#pragma unroll
            for (int nb = 0; nb < kSBN / 8; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    s[nb][e] = p.soft_cap * tanhf(s[nb][e] * p.sm_scale / p.soft_cap) / p.sm_scale;
        }
        const bool need_mask = (col0 + kSBN > kv_hi) || (p.mask_mode == UMB_MASK_DENSE) || (col0 + kSBN > prefix_len);
        if (need_mask) {
#pragma unroll
            for (int ri = 0; ri < 2; ++ri) {
                const uint32_t win = row_window32(p, tok[ri], row_ok[ri], col0, kv_hi, prefix_len) >> (t4 * 2);
#pragma unroll
                for (int nb = 0; nb < kSBN / 8; ++nb) {
                    if (!((win >> (nb * 8)) & 1u)) s[nb][2 * ri] = -INFINITY;
                    if (!((win >> (nb * 8 + 1)) & 1u)) s[nb][2 * ri + 1] = -INFINITY;
                }
            }
        }
        float corr[2];
#pragma unroll
        for (int ri = 0; ri < 2; ++ri) {
            float mx = -INFINITY;
#pragma unroll
            for (int nb = 0; nb < kSBN / 8; ++nb) mx = fmaxf(mx, fmaxf(s[nb][2 * ri], s[nb][2 * ri + 1]));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_new = fmaxf(m_run[ri], mx * p.scale_log2);
            const float m_use = m_new == -INFINITY ? 0.f : m_new;
            corr[ri] = exp2f(m_run[ri] - m_use);
            float sum = 0.f;
#pragma unroll
            for (int nb = 0; nb < kSBN / 8; ++nb) {
                const float p0 = exp2f(fmaf(s[nb][2 * ri], p.scale_log2, -m_use));
                const float p1 = exp2f(fmaf(s[nb][2 * ri + 1], p.scale_log2, -m_use));
                s[nb][2 * ri] = p0;
                s[nb][2 * ri + 1] = p1;
                sum += p0 + p1;
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            l_run[ri] = l_run[ri] * corr[ri] + sum;
            m_run[ri] = m_new;
        }
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
            o_acc[i][0] *= corr[0];
            o_acc[i][1] *= corr[0];
            o_acc[i][2] *= corr[1];
            o_acc[i][3] *= corr[1];
        }
#pragma unroll
        for (int ks = 0; ks < kSBN / 16; ++ks) {
            uint32_t a[4];
            a[0] = pack_half2(s[2 * ks][0], s[2 * ks][1]);
            a[1] = pack_half2(s[2 * ks][2], s[2 * ks][3]);
            a[2] = pack_half2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
            a[3] = pack_half2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
            for (int db = 0; db < D / 8; db += 2) {
                uint32_t b[4];
                const int r = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int c = db + (lane >> 4);
                ldmatrix_x4_trans(b, tile_ptr<D>(tV, r, c));
                mma_16816(o_acc[db], a, b[0], b[1]);
                mma_16816(o_acc[db + 1], a, b[2], b[3]);
            }
        }
        __syncwarp();
        if (kRing == 1 && tile + kWarps < tile_hi) {       // single buffer: the next tile is requested once this one is consumed
            load_tile_warp<D>(ring, kg, p.kv_tok_stride, (tile + kWarps) * kSBN, kv_hi, lane);
            load_tile_warp<D>(ring + kSBN * D, vg, p.kv_tok_stride, (tile + kWarps) * kSBN, kv_hi, lane);
            cp_async_commit();
        }
    }

    // ---- merge the warps' partials through shared memory (first 128 threads) ----
    __syncthreads();                                   // every warp is done with its ring
    constexpr int kLd = D + 4;                         // padded fp32 row
    float* s_o = reinterpret_cast<float*>(smem_raw);   // [kWarps][16][kLd]
    float* s_m = s_o + kWarps * 16 * kLd;              // [kWarps][16]
    float* s_l = s_m + kWarps * 16;                    // [kWarps][16]
#pragma unroll
    for (int ri = 0; ri < 2; ++ri) {
        const int r = g4 + 8 * ri;
        float* orow = s_o + (warp * 16 + r) * kLd;
#pragma unroll
        for (int db = 0; db < D / 8; ++db)
            *reinterpret_cast<float2*>(orow + db * 8 + t4 * 2) = make_float2(o_acc[db][2 * ri], o_acc[db][2 * ri + 1]);
        if (t4 == 0) {
            s_m[warp * 16 + r] = m_run[ri];
            s_l[warp * 16 + r] = l_run[ri];
        }
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int r = threadIdx.x >> 3;                // 16 rows x 8 threads
        const int c0 = (threadIdx.x & 7) * (D / 8);    // D/8 consecutive columns per thread
        const int grow_in_head = row_base + r;
        if (grow_in_head < rows_total) {
            float m = -INFINITY;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) m = fmaxf(m, s_m[w * 16 + r]);
            const float m_use = m == -INFINITY ? 0.f : m;
            float wgt[kWarps], l = 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                wgt[w] = exp2f(s_m[w * 16 + r] - m_use);
                l += s_l[w * 16 + r] * wgt[w];
            }
            const int tk = grow_in_head / p.G, hq = hkv * p.G + grow_in_head % p.G;
            if (p.splits == 1) {
                const float inv = l > 0.f ? 1.f / l : 0.f;
                __half* orow = p.out + ((int64_t)tk * p.Hq + hq) * D + c0;
#pragma unroll
                for (int c = 0; c < D / 8; c += 2) {
                    float v0 = 0.f, v1 = 0.f;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) {
                        v0 += s_o[(w * 16 + r) * kLd + c0 + c] * wgt[w];
                        v1 += s_o[(w * 16 + r) * kLd + c0 + c + 1] * wgt[w];
                    }
                    *reinterpret_cast<uint32_t*>(orow + c) = pack_half2(v0 * inv, v1 * inv);
                }
            } else {
                const int64_t grow = (int64_t)split * p.Hkv * rows_total + (int64_t)hkv * rows_total + grow_in_head;
                float* orow = p.part_o + grow * D + c0;
#pragma unroll
                for (int c = 0; c < D / 8; ++c) {
                    float v = 0.f;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) v += s_o[(w * 16 + r) * kLd + c0 + c] * wgt[w];
                    orow[c] = v;
                }
                if ((threadIdx.x & 7) == 0) p.part_ml[grow] = make_float2(m, l);
            }
        }
    }
}

#ifndef UMB_ATTN_TC
#define UMB_ATTN_TC 1   // 0: build without the tensor-memory long-context kernel
#endif
#if UMB_ATTN_TC
// ------------------------------------------------------------------------------------------------
// Long-context variant on the 5th-generation tensor cores (head_dim 128, <= 128 flattened query rows: every tree of
// <= 16 nodes at G = 8): one CTA per (kv split, kv head) holds ALL query rows of the head group as one 128-row UMMA
// tile, so every K/V byte is read once per CTA and the two GEMMs run at tcgen05 rate instead of the legacy mma.sync
// rate that bounds the kernels above at long context (profiles/r2_hmma_rate_b200.txt: a quarter of it).
//   warps 4, 6  TMA producers (K tiles / V tiles): tiles of 64 kv rows ([2 atoms][64 rows][128 B], 128-byte swizzle), ONE
//               request per tile through 4-D tensor maps over the cache (d, kv row, atom, kv head), a 5-deep ring
//   warp 5      MMA issuer: S = Q K^T (A = Q and B = K from shared memory, both K-major, accumulator in tensor memory,
//               double-buffered), O += P V (A = P from TENSOR memory, B = the V tile as it lies in the cache, i.e.
//               MN-major: LBO = the 64-d atom stride, SBO = 1024 B between 8-row groups -- pinned by umb_probe_umma mode 2)
//   warps 0..3  one thread per query row: ONE tcgen05.ld of its 64-column S row, mask (skipped inside the committed
//               prefix), online softmax in fp32 (no shuffles: the row is the thread's; ex2.approx), P as fp16 pairs back to
//               tensor memory with tcgen05.st; the reference exponent is lazy (moves only when a tile exceeds it by 2^8),
//               so O in tensor memory is rescaled a handful of times per row instead of on every tile
// Tensor memory: S0 | S1 (64 columns each) | O (128, at column 256) | P0 | P1 (32 each, at 384).  Partials (m, l, O) per split go to the
// same workspace layout as the kernels above and are merged by tree_attn_combine_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kTcT = 64;                        // kv rows per tile
constexpr int kTcKv = 5;                        // K/V ring depth: 4 tiles (128 KB) in flight ahead of the one being consumed
constexpr int kTcChunks = kTcT / 32;            // 32-column chunks of a score row
constexpr int kTcThreads = 224;                 // 4 softmax warps, K producer, MMA issuer, V producer
constexpr int kTcTileBytes = kTcT * 256;        // one K (or V) tile: 2 atoms x kTcT rows x 128 B
constexpr int kTcSmem = 1024 + 32768 + kTcKv * 2 * kTcTileBytes + 256;

// 2^x on the special-function unit (ex2.approx.ftz: 2 ulp, the arguments are <= 0 here; -inf -> 0)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#ifdef UMB_ATTN_TC_DEBUG
// debugging build: every wait of block (0, 0) leaves its source line in a host-mapped progress buffer, one slot per warp
// (low bit 8 set = still waiting), so that a hang can be read from the host while the kernel is stuck
__device__ unsigned long long* g_tc_dbg = nullptr;
__device__ __forceinline__ void tc_mark(int slot, unsigned long long v) {
    if (g_tc_dbg && blockIdx.x == 0 && blockIdx.y == 0) {
        g_tc_dbg[slot] = v;
        __threadfence_system();
    }
}
__device__ __forceinline__ void tc_wait_dbg(int line, uint64_t* bar, uint32_t parity) {
#ifdef UMB_ATTN_TC_MARK_WAITS            // progress marks through PCIe: only when hunting a hang
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) tc_mark(w, ((unsigned long long)line << 32) | 0x100u | parity);
    mbar_wait(bar, parity);
    if ((threadIdx.x & 31) == 0) tc_mark(w, ((unsigned long long)line << 32) | parity);
#else
    (void)line;
    mbar_wait_spin(bar, parity);
#endif
}
#define TC_WAIT(line, bar, parity) tc_wait_dbg(line, bar, parity)
#define TC_MARK(slot, v) do { if ((threadIdx.x & 31) == 0) tc_mark(slot, v); } while (0)
// timeline of the first 12 tiles of block (0, 0): %globaltimer stamps kept in shared memory (8 per tile: softmax warp 0
// 0..4, MMA warp 5..7) and copied to the host-mapped buffer when the kernel ends -- stamping through PCIe costs
// microseconds per stamp and measures itself
__device__ __forceinline__ unsigned long long* tc_stamp_buf() {
    __shared__ unsigned long long s_stamp[96];
    return s_stamp;
}
#define TC_STAMP(tile, k) do { if ((threadIdx.x & 31) == 0 && (tile) < 12 && blockIdx.x == 0 && blockIdx.y == 0 && \
                                   ((threadIdx.x >> 5) == 0 || (threadIdx.x >> 5) == 5)) { \
        unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); tc_stamp_buf()[(tile) * 8 + (k)] = t__; } } while (0)
#else
#define TC_STAMP(tile, k) do { } while (0)
#define TC_WAIT(line, bar, parity) mbar_wait_spin(bar, parity)
#define TC_MARK(slot, v) do { } while (0)
#endif
__global__ void __launch_bounds__(kTcThreads, 1)
tree_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                    const AttnParams p) {
    constexpr int D = 128;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* s_q = smem;                                  // [2 atoms][128 rows][128 B]
    unsigned char* s_kv = smem + 32768;                         // [kTcKv stages][K | V][kTcTileBytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_kv + kTcKv * 2 * kTcTileBytes);
    uint64_t* k_full = bars;            // [kTcKv]
    uint64_t* v_full = bars + 8;        // [kTcKv]
    uint64_t* kv_empty = bars + 16;     // [kTcKv]  both tiles of the stage consumed (PV retired)
    uint64_t* s_full = bars + 24;       // [2]  S buffer written (QK retired)
    uint64_t* s_empty = bars + 26;      // [2]  S buffer read by the four softmax warps
    uint64_t* p_full = bars + 28;       // [2]  P buffer written, O rescaled
    uint64_t* o_ready = bars + 30;      //      PV of a tile retired
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 31);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, hkv = blockIdx.y;
    const int rows_total = p.qo * p.G;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
        for (int i = 0; i < kTcKv; ++i) {
            mbar_init(k_full + i, 1);
            mbar_init(v_full + i, 1);
            mbar_init(kv_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_full + i, 1);
            mbar_init(s_empty + i, 4);
            mbar_init(p_full + i, 4);
        }
        mbar_init(o_ready, 1);
        fence_barrier_init();
    }
    TC_MARK(8 + warp, 1);
    if (warp == 5) {
        tmem_alloc(s_tmem, 512);
        tmem_relinquish();
    }
    TC_MARK(8 + warp, 2);
    pdl_launch_dependents();
    pdl_wait();                                 // q and the appended K/V rows are the predecessors' output
    TC_MARK(8 + warp, 3);
    const int prefix_len = p.dyn ? p.dyn[0] + p.prefix_len : p.prefix_len;
    const int kv_len = p.dyn ? p.dyn[0] + p.kv_len : p.kv_len;
    int kv_hi = kv_len;
    if (p.mask_mode == UMB_MASK_CAUSAL)
        kv_hi = min(kv_hi, prefix_len + (rows_total - 1) / p.G + 1);
    else if (p.mask_mode == UMB_MASK_TREE)
        kv_hi = min(kv_hi, prefix_len + p.tree_words * 32);
    const int n_tiles = (kv_hi + kTcT - 1) / kTcT;
    const int tiles_per_split = (n_tiles + p.splits - 1) / p.splits;
    const int tile_lo = split * tiles_per_split;
    const int tile_hi = min(n_tiles, tile_lo + tiles_per_split);
    const int n_my = max(tile_hi - tile_lo, 0);

    // Q rows of this head group -> shared memory, K-major with the 128-byte swizzle (row r = token r / G, head r % G)
    if (warp < 4) {
        const int r = threadIdx.x;
        const bool ok = r < rows_total;
        const __half* qr = p.q + ((int64_t)(ok ? r / p.G : 0) * p.Hq + hkv * p.G + (ok ? r % p.G : 0)) * D;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (ok) v = *reinterpret_cast<const uint4*>(qr + c * 8);
            *reinterpret_cast<uint4*>(s_q + (c >> 3) * 16384 + (r >> 3) * 1024 + (r & 7) * 128 + (((c & 7) ^ (r & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    constexpr uint32_t kColS = 0, kColO = 256, kColP = 384;
    TC_MARK(8 + warp, 4 | ((unsigned long long)n_my << 8) | ((unsigned long long)tmem << 32));

    if (warp == 4 || warp == 6) {
        // ===================== TMA producers: warp 4 requests the K tiles, warp 6 the V tiles =====================
        // One request per 16 KB tile (4-D map: both 64-d atoms in one box) from its own thread: a thread gets a TMA request
        // out every ~230 ns whatever its size (profiles/r2_tma_rates_b200.txt); four requests per 64-row tile from one
        // thread WERE the tile period of the first form of this kernel (1.57 us, profiles/r2_attn_long_context.txt).
        const uint64_t pol = l2_policy_evict_first();
        const bool is_k = warp == 4;
        const CUtensorMap* tm = is_k ? &tmap_k : &tmap_v;
        uint64_t* full = is_k ? k_full : v_full;
        for (int i = 0; i < n_my; ++i) {
            const int st = i % kTcKv, use = i / kTcKv;
            if (use > 0) TC_WAIT(__LINE__, kv_empty + st, (use - 1) & 1);
            if (elect_one()) {
                unsigned char* dst = s_kv + st * (2 * kTcTileBytes) + (is_k ? 0 : kTcTileBytes);
                mbar_arrive_expect_tx(full + st, (uint32_t)kTcTileBytes);
                tma_tensor4d_g2s(dst, tm, 0, (tile_lo + i) * kTcT, 0, hkv, full + st, pol);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        constexpr uint32_t kIdescQK = umma_idesc_f16(128, kTcT);
        constexpr uint32_t kIdescPV = umma_idesc_f16(128, D) | (1u << 16);          // B (V) MN-major
        auto issue_qk = [&](int i) {
            const int st = i % kTcKv, use = i / kTcKv, sb = i & 1, su = i >> 1;
            TC_WAIT(__LINE__, k_full + st, use & 1);
            if (su > 0) TC_WAIT(__LINE__, s_empty + sb, (su - 1) & 1);                        // S buffer i & 1 has been read
            tc_fence_after();
            if (elect_one()) {
                const uint32_t kt = smem_u32(s_kv + st * (2 * kTcTileBytes));
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint64_t ad = umma_smem_desc_sw128(smem_u32(s_q) + (kk >> 2) * 16384 + (kk & 3) * 32);
                    const uint64_t bd = umma_smem_desc_sw128(kt + (kk >> 2) * (kTcT * 128) + (kk & 3) * 32);
                    umma_ss_f16(tmem + kColS + sb * kTcT, ad, bd, kIdescQK, kk ? 1u : 0u);
                }
                umma_commit(s_full + sb);
            }
            __syncwarp();
        };
        if (n_my > 0) issue_qk(0);
        for (int i = 0; i < n_my; ++i) {
            if (i + 1 < n_my) issue_qk(i + 1);
            const int st = i % kTcKv, use = i / kTcKv, sb = i & 1, su = i >> 1;
            TC_STAMP(i, 5);
            TC_WAIT(__LINE__, v_full + st, use & 1);
            TC_WAIT(__LINE__, p_full + sb, su & 1);
            tc_fence_after();
            TC_STAMP(i, 6);
            if (elect_one()) {
                const uint32_t vt = smem_u32(s_kv + st * (2 * kTcTileBytes) + kTcTileBytes);
#pragma unroll
                for (int kk = 0; kk < kTcT / 16; ++kk) {
                    uint64_t bd = 0;
                    bd |= (uint64_t)(((vt + kk * 2048) & 0x3FFFF) >> 4);              // 16 kv rows per k-step
                    bd |= (uint64_t)((kTcT * 128) >> 4) << 16;                        // LBO: next 64-d atom
                    bd |= (uint64_t)(1024 >> 4) << 32;                                // SBO: next group of 8 kv rows
                    bd |= (uint64_t)1 << 46;
                    bd |= (uint64_t)2 << 61;
                    umma_ts_f16(tmem + kColO, tmem + kColP + sb * (kTcT / 2) + kk * 8, bd, kIdescPV, (i > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(kv_empty + st);
                umma_commit(o_ready);
            }
            __syncwarp();
            TC_STAMP(i, 7);
        }
    } else {
        // ===================== softmax: thread = query row =====================
        const int r = threadIdx.x;
        const bool row_ok = r < rows_total;
        const int tok = row_ok ? r / p.G : 0;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        for (int i = 0; i < n_my; ++i) {
            const int st = i & 1, use = i >> 1;             // S / P buffer and its use count
            const int col0 = (tile_lo + i) * kTcT;
            // warp-uniform on purpose: the rows past qo * G of the last warp run the same unmasked path on their all-zero Q
            // (never stored) -- with `|| !row_ok` here that warp diverged on every tile, took twice as long as the other
            // three and set the pace of the whole pipeline (p_full needs all four warps)
            const bool need_mask = (col0 + kTcT > kv_hi) || (col0 + kTcT > prefix_len);
            TC_STAMP(i, 0);
            TC_WAIT(__LINE__, s_full + st, use & 1);
            tc_fence_after();
            TC_STAMP(i, 1);
            const uint32_t s_addr = tmem + lane_addr + kColS + st * kTcT;
            uint32_t win[kTcChunks];
#pragma unroll
            for (int c = 0; c < kTcChunks; ++c) win[c] = ~0u;
            if (need_mask) {
#pragma unroll
                for (int c = 0; c < kTcChunks; ++c) win[c] = row_window32(p, tok, row_ok, col0 + 32 * c, kv_hi, prefix_len);
            }
            // the whole score row (kTcT = 64 columns) in ONE tensor-memory load, kept in registers for both passes
            static_assert(kTcT == 64, "the softmax below holds one 64-column row per thread");
            uint32_t v[64];
            UMB_TMEM_LD_X64(s_addr, v);
            tc_wait_ld();
            // pass 1: row maximum.  Tiles inside the committed prefix need no mask: the common case at long context runs
            // without the per-element window tests.
            float mx = -INFINITY;
            if (need_mask) {
#pragma unroll
                for (int j = 0; j < 64; ++j)
                    if ((win[j >> 5] >> (j & 31)) & 1u) mx = fmaxf(mx, __uint_as_float(v[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 64; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
            // Lazy reference: the exponentials stay relative to the running reference m_run as long as the tile's maximum
            // exceeds it by less than 2^8 (P <= 256 in fp16, sums in fp32: exact up to rounding) -- O in tensor memory is
            // rescaled only when the reference really moves, not on every small increase of some row's maximum.
            const float m_tile = mx * p.scale_log2;
            const float m_new = (m_tile > m_run + 8.f || m_run == -INFINITY) ? fmaxf(m_run, m_tile) : m_run;
            const float m_use = m_new == -INFINITY ? 0.f : m_new;
            const float corr = m_new == m_run ? 1.f : exp2f(m_run - m_use);
            // pass 2: probabilities -> P (fp16 pairs, the A operand of the PV product)
            float l_tile = 0.f;
            const uint32_t p_addr = tmem + lane_addr + kColP + st * (kTcT / 2);
            uint32_t pk[32];
            if (need_mask) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float p0 = ((win[j >> 4] >> ((2 * j) & 31)) & 1u) ? ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -m_use)) : 0.f;
                    const float p1 = ((win[j >> 4] >> ((2 * j + 1) & 31)) & 1u) ? ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -m_use)) : 0.f;
                    l_tile += p0 + p1;
                    pk[j] = pack_half2(p0, p1);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -m_use));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -m_use));
                    l_tile += p0 + p1;
                    pk[j] = pack_half2(p0, p1);
                }
            }
            UMB_TMEM_ST_X32(p_addr, pk);
            TC_STAMP(i, 2);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + st);            // this warp's 32 rows of S are consumed
            // O so far belongs to the previous maximum: rescale it once the previous PV has retired
            if (i > 0) {
                TC_WAIT(__LINE__, o_ready, (i - 1) & 1);
                tc_fence_after();
                TC_STAMP(i, 3);
                if (__any_sync(0xffffffffu, corr != 1.f)) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32];
                        UMB_TMEM_LD_X32(tmem + lane_addr + kColO + c * 32, v);
                        tc_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * corr);
                        UMB_TMEM_ST_X32(tmem + lane_addr + kColO + c * 32, v);
                    }
                }
            }
            l_run = l_run * corr + l_tile;
            m_run = m_new;
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full + st);
            TC_STAMP(i, 4);
        }
        // ---- output: the normalised rows (one split) or the (m, l, O) partial of this split ----
        if (n_my > 0) {
            TC_WAIT(__LINE__, o_ready, (n_my - 1) & 1);
            tc_fence_after();
        }
        {
            // every lane executes the tensor-memory loads (.sync.aligned: a partially active warp is undefined behaviour --
            // the rows past rows_total of the last warp hung the first version here); only the stores are predicated
            const int hq = hkv * p.G + (row_ok ? r % p.G : 0);
            const float inv = (p.splits == 1) ? (l_run > 0.f ? 1.f / l_run : 0.f) : 1.f;
            __half* orow = p.out + ((int64_t)tok * p.Hq + hq) * D;
            const int64_t grow = (int64_t)split * p.Hkv * rows_total + (int64_t)hkv * rows_total + (row_ok ? r : 0);
            float* prow = p.splits == 1 ? nullptr : p.part_o + grow * D;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                if (n_my > 0) {
                    UMB_TMEM_LD_X32(tmem + lane_addr + kColO + c * 32, v);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                if (!row_ok) continue;
                if (p.splits == 1) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 o;
                        o.x = pack_half2(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
                        o.y = pack_half2(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
                        o.z = pack_half2(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
                        o.w = pack_half2(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + c * 32 + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(prow + c * 32 + j) =
                            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                }
            }
            if (row_ok && p.splits > 1) p.part_ml[grow] = make_float2(m_run, l_run);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
#ifdef UMB_ATTN_TC_DEBUG
    if (g_tc_dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 96) g_tc_dbg[64 + threadIdx.x] = tc_stamp_buf()[threadIdx.x];
#endif
}


#endif

// grid (Hkv * qo * G); block D threads.  Merges the per-split (m, l, o) triples.
template <int D>
__global__ void __launch_bounds__(D) tree_attn_combine_kernel(const AttnParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int rows_total = p.qo * p.G;
    const int grow = blockIdx.x;                 // hkv * rows_total + row
    const int hkv = grow / rows_total, row = grow % rows_total;
    const int64_t stride = (int64_t)p.Hkv * rows_total;
    float m = -INFINITY;
    for (int s = 0; s < p.splits; ++s) m = fmaxf(m, p.part_ml[s * stride + grow].x);
    const float m_use = m == -INFINITY ? 0.f : m;
    float l = 0.f, o = 0.f;
    for (int s = 0; s < p.splits; ++s) {
        const int64_t idx = s * stride + grow;
        const float2 ml = p.part_ml[idx];
        const float w = exp2f(ml.x - m_use);
        l += ml.y * w;
        o += p.part_o[idx * D + threadIdx.x] * w;
    }
    const int tok = row / p.G, hq = hkv * p.G + row % p.G;
    p.out[((int64_t)tok * p.Hq + hq) * D + threadIdx.x] = __float2half_rn(l > 0.f ? o / l : 0.f);
}

}  // namespace umb

using namespace umb;

#ifdef UMB_ATTN_TC_DEBUG
extern "C" int umb_debug_attn_tc_ptr(void* host_mapped) {
    unsigned long long* p = (unsigned long long*)host_mapped;
    return (int)cudaMemcpyToSymbol(umb::g_tc_dbg, &p, sizeof(p));
}
#endif

extern "C" size_t umb_tree_attn_workspace_bytes(int qo, int num_q_heads, int head_dim, int max_kv_len) {
    (void)max_kv_len;
    if (qo < 1 || num_q_heads < 1 || head_dim < 1) return 0;
    const size_t rows = (size_t)qo * num_q_heads;
    return 32 * rows * ((size_t)head_dim * sizeof(float) + sizeof(float2)) + 256;
}

#if UMB_ATTN_TC
// 4-D tensor map over a K (or V) cache plane with the caller's strides, 128-byte swizzle -- one box is one whole tile (both
// 64-d atoms) as tcgen05.mma reads it.
typedef CUresult (*PFN_encodeTiledAttn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_tmap_kv(CUtensorMap* map, const void* base, int D, int Hkv, int64_t rows, int64_t tok_stride,
                        int64_t head_stride) {
    static PFN_encodeTiledAttn enc = nullptr;
    if (!enc) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return UMB_ERR_DRIVER;
        enc = reinterpret_cast<PFN_encodeTiledAttn>(sym);
    }
    // (d within a 64-wide atom, kv row, atom, kv head): one box = [2 atoms][kTcT rows][128 B] in shared memory
    const cuuint64_t dims[4] = {64, (cuuint64_t)rows, (cuuint64_t)(D / 64), (cuuint64_t)Hkv};
    const cuuint64_t strides[3] = {(cuuint64_t)tok_stride * sizeof(__half), 64 * sizeof(__half),
                                   (cuuint64_t)head_stride * sizeof(__half)};
    const cuuint32_t box[4] = {64, (cuuint32_t)kTcT, (cuuint32_t)(D / 64), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? UMB_OK : arg_fail(__FILE__, __LINE__, "cuTensorMapEncodeTiled(kv)");
}

// Contexts of >= 2048 KV rows go to the tensor-memory kernel: head_dim 128, all query rows of a head group in one 128-row
// tile, implicit masks only.  Measured against the mma.sync kernels (profiles/r2_attn_long_context.txt): 21.9 vs 24.4 us at
// 2k, 24.8 vs 29.5 at 4k, 30.8 vs 33.6 at 8k, 61.2 vs 80.6 at 32k; slower below (20.1 vs 17.5 us at 1k).
// UMB_ATTN_TC=0 switches it off at run time, UMB_ATTN_TC_MIN_KV moves the threshold.
static int attn_tc_min_kv() {
    static int cached = -1;
    if (cached < 0) {
        const char* off = getenv("UMB_ATTN_TC");
        const char* e = getenv("UMB_ATTN_TC_MIN_KV");
        cached = (off && off[0] == '0') ? (1 << 30) : (e ? atoi(e) : 2048);
    }
    return cached;
}


#endif

template <int D, bool kSoftCap, int kWarps>
static int launch_slab(const AttnParams& p, dim3 grid, cudaStream_t st) {
    const size_t smem = SlabCfg<D, kWarps>::kSmem;
    // per-device attribute (a process may drive several GPUs): set on every launch, it is a cheap host call
    if (smem > 48 * 1024)
        UMB_CUDA(cudaFuncSetAttribute(tree_attn_slab_kernel<D, kSoftCap, kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UMB_LAUNCH_PDL((tree_attn_slab_kernel<D, kSoftCap, kWarps>), grid, dim3(kWarps * 32), smem, st, p);
    return UMB_OK;
}

template <int D, bool kSoftCap>
static int launch_attn(const AttnParams& p, bool slab, int slab_warps, dim3 grid, cudaStream_t st) {
    if (slab) {
        const int rc = slab_warps == 8 ? launch_slab<D, kSoftCap, 8>(p, grid, st) : launch_slab<D, kSoftCap, 4>(p, grid, st);
        if (rc != UMB_OK) return rc;
    } else {
        const size_t smem = (size_t)4 * kBN * D * sizeof(__half);
        if (smem > 48 * 1024)
            UMB_CUDA(cudaFuncSetAttribute(tree_attn_kernel<D, kSoftCap>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UMB_LAUNCH_PDL(tree_attn_kernel<D, kSoftCap>, grid, dim3(kAttnThreads), smem, st, p);
    }
    int launches = 1;
    if (p.splits > 1) {
        UMB_LAUNCH_PDL(tree_attn_combine_kernel<D>, dim3((unsigned)(p.Hkv * p.qo * p.G)), dim3(D), 0, st, p);
        launches = 2;
    }
    count_launches(launches);
    UMB_LAUNCH_CHECK();
    return UMB_OK;
}

extern "C" int umb_tree_attn_f16(const void* q, const void* k_cache, const void* v_cache, void* out, int qo, int kv_len,
                                 int num_q_heads, int num_kv_heads, int head_dim, int64_t kv_tok_stride,
                                 int64_t kv_head_stride, const umb_attn_mask_t* mask, float sm_scale, float soft_cap,
                                 void* workspace, size_t workspace_bytes, umb_stream_t stream) {
    UMB_CHECK_ARG(q && k_cache && v_cache && out && mask);
    UMB_CHECK_ARG(qo >= 1 && kv_len >= 1 && num_q_heads >= 1 && num_kv_heads >= 1);
    const int kv_bound = mask->dyn ? (mask->kv_len_bound > kv_len ? mask->kv_len_bound : kv_len) : kv_len;
    UMB_CHECK_ARG(num_q_heads % num_kv_heads == 0);
    UMB_CHECK_ARG((kv_tok_stride % 8) == 0 && (kv_head_stride % 8) == 0);
    UMB_CHECK_ARG((((uintptr_t)q | (uintptr_t)k_cache | (uintptr_t)v_cache | (uintptr_t)out) & 15) == 0);
    if (head_dim != 64 && head_dim != 128) return UMB_ERR_UNSUPPORTED;
    UMB_CHECK_ARG(mask->mask_mode >= 0 && mask->mask_mode <= 2);
    if (mask->mask_mode == UMB_MASK_TREE) UMB_CHECK_ARG(mask->tree_bits && mask->tree_words >= 1);
    if (mask->mask_mode == UMB_MASK_DENSE) UMB_CHECK_ARG(mask->dense_mask && mask->mask_stride >= kv_len && !mask->dyn);

    AttnParams p{};
    p.q = (const __half*)q;
    p.k = (const __half*)k_cache;
    p.v = (const __half*)v_cache;
    p.out = (__half*)out;
    p.qo = qo;
    p.kv_len = kv_len;
    p.Hq = num_q_heads;
    p.Hkv = num_kv_heads;
    p.G = num_q_heads / num_kv_heads;
    p.kv_tok_stride = kv_tok_stride;
    p.kv_head_stride = kv_head_stride;
    p.mask_mode = mask->mask_mode;
    p.prefix_len = mask->prefix_len;
    p.dyn = mask->dyn;
    p.tree_bits = mask->tree_bits;
    p.tree_words = mask->tree_words;
    p.dense_mask = mask->dense_mask;
    p.mask_stride = mask->mask_stride;
    p.sm_scale = sm_scale > 0.f ? sm_scale : 1.0f / sqrtf((float)head_dim);
    p.scale_log2 = p.sm_scale * 1.4426950408889634f;
    p.soft_cap = soft_cap;

#if UMB_ATTN_TC
    const int rows_total_tc = qo * p.G;
    if (head_dim == 128 && rows_total_tc <= 128 && soft_cap <= 0.f && mask->mask_mode != UMB_MASK_DENSE &&
        kv_bound >= attn_tc_min_kv() && (kv_tok_stride * sizeof(__half)) % 16 == 0 && (kv_head_stride * sizeof(__half)) % 16 == 0) {
        const int n_tiles = (kv_bound + kTcT - 1) / kTcT;
        int splits = sm_count() / num_kv_heads;
        splits = splits < 1 ? 1 : (splits > 32 ? 32 : splits);
        splits = splits > n_tiles ? n_tiles : splits;
        const size_t rows = (size_t)qo * num_q_heads;
        const size_t per_split = rows * ((size_t)head_dim * sizeof(float) + sizeof(float2));
        if (splits > 1) {
            const size_t avail = (workspace && workspace_bytes > 256) ? (workspace_bytes - 256) / per_split : 0;
            if ((size_t)splits > avail) splits = avail < 1 ? 1 : (int)avail;
        }
        p.splits = splits;
        if (splits > 1) {
            const uintptr_t base = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
            p.part_o = reinterpret_cast<float*>(base);
            p.part_ml = reinterpret_cast<float2*>(base + (size_t)splits * rows * head_dim * sizeof(float));
        }
        CUtensorMap tk, tv;
        int rc = make_tmap_kv(&tk, k_cache, head_dim, num_kv_heads, kv_bound, kv_tok_stride, kv_head_stride);
        if (rc == UMB_OK) rc = make_tmap_kv(&tv, v_cache, head_dim, num_kv_heads, kv_bound, kv_tok_stride, kv_head_stride);
        if (rc == UMB_OK) {
            cudaStream_t st_tc = (cudaStream_t)stream;
            UMB_CUDA(cudaFuncSetAttribute(tree_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
            UMB_LAUNCH_PDL(tree_attn_tc_kernel, dim3(splits, num_kv_heads), dim3(kTcThreads), (size_t)kTcSmem, st_tc, tk, tv, p);
            int launches = 1;
            if (splits > 1) {
                UMB_LAUNCH_PDL(tree_attn_combine_kernel<128>, dim3((unsigned)(p.Hkv * p.qo * p.G)), dim3(128), 0, st_tc, p);
                launches = 2;
            }
            count_launches(launches);
            UMB_LAUNCH_CHECK();
            return UMB_OK;
        }
        // the tensor maps could not be encoded (driver entry point missing, unusual strides): the mma.sync kernels below
        // take the call
    }

#endif
    // kernel + split selection.  slab kernel: K/V is re-read once per 16-row slab (L2), so it is used
    // while that re-read is cheap; long contexts with many slabs go to the 64-row tile kernel.
    const int rows_total = qo * p.G;
    const int slabs = (rows_total + 15) / 16;
    const bool slab = !(kv_bound > 8192 && slabs > 2);
    const int ctas_base = (slab ? slabs : (rows_total + kBM - 1) / kBM) * p.Hkv;
    const int kv_tiles = (kv_bound + (slab ? kSBN : kBN) - 1) / (slab ? kSBN : kBN);
    static int slab_tiles = 0;                                // UMB_ATTN_SLAB_TILES: tuning override (read once)
    if (slab_tiles == 0) {
        const char* e = getenv("UMB_ATTN_SLAB_TILES");
        slab_tiles = e ? atoi(e) : kSlabTilesPerCta;
        if (slab_tiles < 4) slab_tiles = kSlabTilesPerCta;
    }
    const int tiles_per_cta_min = slab ? slab_tiles : 4;      // slab CTAs run 4 tiles at a time (one per warp)
    int splits = (kv_tiles + tiles_per_cta_min - 1) / tiles_per_cta_min;
    const int by_ctas = (2 * sm_count()) / (ctas_base > 0 ? ctas_base : 1);
    splits = splits > by_ctas ? by_ctas : splits;
    splits = splits > 32 ? 32 : splits;
    splits = splits < 1 ? 1 : splits;
    const size_t rows = (size_t)qo * num_q_heads;
    const size_t per_split = rows * ((size_t)head_dim * sizeof(float) + sizeof(float2));
    if (splits > 1) {
        const size_t avail = (workspace && workspace_bytes > 256) ? (workspace_bytes - 256) / per_split : 0;
        if ((size_t)splits > avail) splits = avail < 1 ? 1 : (int)avail;
    }
    p.splits = splits;
    if (splits > 1) {
        const uintptr_t base = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
        p.part_o = reinterpret_cast<float*>(base);
        p.part_ml = reinterpret_cast<float2*>(base + (size_t)splits * rows * head_dim * sizeof(float));
    }
    const dim3 grid(slab ? slabs : (rows_total + kBM - 1) / kBM, p.Hkv, splits);
    // slab CTAs of 8 warps when each warp would otherwise walk more than two tiles (UMB_ATTN_WARPS=4 / 8: tuning override)
    static int warps_override = -1;
    if (warps_override < 0) {
        const char* e = getenv("UMB_ATTN_WARPS");
        warps_override = e ? atoi(e) : 0;
    }
    const int tiles_per_cta = (kv_tiles + splits - 1) / splits;
    int slab_warps = tiles_per_cta > 8 ? 8 : 4;
    if (warps_override == 4 || warps_override == 8) slab_warps = warps_override;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128)
        return soft_cap > 0.f ? launch_attn<128, true>(p, slab, slab_warps, grid, st)
                              : launch_attn<128, false>(p, slab, slab_warps, grid, st);
    return soft_cap > 0.f ? launch_attn<64, true>(p, slab, slab_warps, grid, st) : launch_attn<64, false>(p, slab, slab_warps, grid, st);
}
