// (i) AWQ INT4 weight-dequant fused into a tcgen05 tensor-core GEMM, plus its fp16-weight twin.
//
// Replaces awq_ext.gemm_forward_cuda / dequantize_weights_cuda as called from AwqLinear.apply
// (reference umbrella/quantization/awq_utils.py:63-86) and torch F.linear for fp16 weights
// (lm_head: umbrella/models/llama.py:321; fp16 models: llama.py:87-110).
//
// Orientation ("swap-AB"): the verify pass multiplies M <= 128 token rows by a huge weight matrix,
// so the WEIGHTS are the UMMA A operand (128 output features per tile = the 128 TMEM lanes) and the
// token rows are the UMMA B/N operand (N = M padded to 16/32/64/128):
//     D[128 features, MPAD tokens] (fp32, TMEM) += A[128, k] * B[MPAD, k]^T .
// Every weight byte is read from HBM exactly once per call.
//
// AWQ mainloop (per 128-feature x 128-k "unit" = one 8512-byte record of the repacked stream):
//   warp 0      TMA producer: cp.async.bulk of the packed record + cp.async.bulk.tensor (SW128)
//               of the x slab [MPAD, 128 k] into a kStages-deep shared-memory ring
//   warps 2..13 dequant, three groups of four warps taking units round-robin: thread = one feature
//               row x 128 k; one lop3 per nibble pair -> fp16 (1024+q | 1024+16q), an exact
//               hsub2 / hfma2 removes the bias and the zero point, one hmul2 by the group scale
//               == fp16((q - z) * s), the same rounding as the reference kernels; the fp16 row
//               goes straight to TENSOR MEMORY with tcgen05.st (no shared-memory round trip)
//   warp 1      one thread issues tcgen05.mma kind::f16 with A from TMEM, B from shared memory
//   warps 2..5  epilogue: tcgen05.ld the fp32 accumulator, fuse bias / residual / SiLU*up / fp32
//               logits, fp16 rounding placed where the reference has fp16 tensors
// Work is distributed stream-K: the (tile, k-group) space is cut into gridDim.x equal contiguous
// ranges (one persistent CTA per SM), partial tiles are reduced through a small fp32 workspace in a
// fixed order by the last-arriving CTA, so results are deterministic.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace umb {

constexpr int kTileN = 128;              // output features per tile (UMMA M)
constexpr int kGroup = 128;              // AWQ group size == k per AWQ unit
constexpr int kRecordBytes = 8512;       // 8192 nibble bytes + 256 scale bytes + 64 zero bytes (what a TMA copy moves)
#ifndef UMB_RECORD_STRIDE
#define UMB_RECORD_STRIDE 8512
#endif
constexpr int kRecordStride = UMB_RECORD_STRIDE;   // distance between records in the stream
constexpr int kRecScaleOff = 8192;
constexpr int kRecZeroOff = 8448;
#ifndef UMB_DQ_GROUPS
#define UMB_DQ_GROUPS 3
#endif
#ifndef UMB_DQ_GROUPS_16
#define UMB_DQ_GROUPS_16 UMB_DQ_GROUPS     // dequant groups of the <= 16-row instantiation (the verify pass of the static trees)
#endif
constexpr int kDqGroups = UMB_DQ_GROUPS;   // dequant warp groups (4 warps each, one per TMEM lane quarter)
constexpr int kGemmThreadsF16 = 96 + 128;           // producer + 2 MMA issuers + 4 epilogue warps (fp16 weights); AWQ: GemmCfg::kThreadsAwq
constexpr int kTmemCols = 512;
#ifndef UMB_WAIT_HINT
#define UMB_WAIT_HINT 1                    // 1: mbarrier waits carry a suspend-time hint (short in the pipeline, long in the epilogue)
#endif
#if UMB_WAIT_HINT
#define UMB_WAIT(bar, parity) mbar_wait_hint(bar, parity, 2000u)
#define UMB_WAIT_LONG(bar, parity) mbar_wait_hint(bar, parity, 100000u)
#else
#define UMB_WAIT(bar, parity) mbar_wait(bar, parity)
#define UMB_WAIT_LONG(bar, parity) mbar_wait(bar, parity)
#endif
#ifndef UMB_X_3D
#define UMB_X_3D 1                         // 1: one 3-D tensor copy per x slab (both 64-k atoms) instead of two 2-D copies (AWQ)
#endif
#ifndef UMB_W_PAIR
#define UMB_W_PAIR 0                       // 1: the weight producer requests two records per cp.async.bulk (AWQ)
#endif
#ifndef UMB_X_PRODUCER
#define UMB_X_PRODUCER 0                   // 1: the x slabs are requested by their own warp (AWQ), not by the weight producer
#endif
// Pipeline variants that were built, measured on B200 and rejected (sleeping mbarrier polls in the idle warps,
// software-pipelined dequant, x slabs ahead of the weight ring, early residual fetch / early partial-sum
// reduction, a fourth dequant group) are recorded in profiles/r1_gemm_experiments.txt; none beat this form,
// whose steady state is bound by the fp16x2 issue rate of the dequant warps (scripts/ubench/pipe_rates.cu).

// nibble position of the i-th of 8 consecutive values inside a 32-bit word: (0,4,1,5,2,6,3,7).
// Used by the AutoAWQ checkpoint layout (along N) and by our repacked stream (along K): nibbles p
// and p + 4 hold consecutive values, so (w >> 4p) & 0x000f000f yields an ordered pair.
__host__ __device__ __forceinline__ int nib_pos(int i) { return (i >> 1) + ((i & 1) << 2); }

// ------------------------------------------------------------------------------------------------
// repack / dequantize / unpack
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t packed_row(int n, int row_block, int row_stride, int row_offset) {
    return (int64_t)(n / row_block) * row_stride + row_offset + n % row_block;
}

// thread (kw, c): 8 k x 8 n block of the checkpoint -> 8 words (one per n) of 8 k-nibbles
__global__ void awq_repack_weights_kernel(const int32_t* __restrict__ qweight, int K, int N, unsigned char* packed,
                                          int row_block, int row_stride, int row_offset) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int kw = blockIdx.y;
    if (c >= N / 8) return;
    uint32_t in[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) in[i] = (uint32_t)qweight[(int64_t)(kw * 8 + i) * (N / 8) + c];
    const int k0 = kw * 8, g = k0 / kGroup, kk = k0 % kGroup;
    const int kb = kk >> 6, ch = (kk >> 5) & 1, w = (kk >> 3) & 3;
    const int G = K / kGroup;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t out = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) out |= ((in[i] >> (4 * nib_pos(j))) & 0xFu) << (4 * nib_pos(i));
        const int64_t pr = packed_row(c * 8 + j, row_block, row_stride, row_offset);
        const int64_t tile = pr / kTileN;
        const int r = (int)(pr % kTileN);
        unsigned char* rec = packed + (tile * G + g) * kRecordStride;
        *reinterpret_cast<uint32_t*>(rec + ((kb * 2 + ch) * kTileN + r) * 16 + w * 4) = out;
    }
}
// thread (g, c): scales and zeros of 8 consecutive output features of one group
__global__ void awq_repack_meta_kernel(const int32_t* __restrict__ qzeros, const __half* __restrict__ scales, int K,
                                       int N, unsigned char* packed, int row_block, int row_stride, int row_offset) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (c >= N / 8) return;
    const int G = K / kGroup;
    const uint32_t zin = (uint32_t)qzeros[(int64_t)g * (N / 8) + c];
    uint32_t zout = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) zout |= ((zin >> (4 * nib_pos(j))) & 0xFu) << (4 * j);
    const int64_t pr = packed_row(c * 8, row_block, row_stride, row_offset);   // 8-aligned run of rows
    const int64_t tile = pr / kTileN;
    const int r = (int)(pr % kTileN);
    unsigned char* rec = packed + (tile * G + g) * kRecordStride;
    *reinterpret_cast<uint32_t*>(rec + kRecZeroOff + (r >> 3) * 4) = zout;
    *reinterpret_cast<uint4*>(rec + kRecScaleOff + r * 2) =
        *reinterpret_cast<const uint4*>(scales + (int64_t)g * N + c * 8);
}

// awq_ext.dequantize_weights_cuda: out[k, n] = fp16((q - z) * s), checkpoint layout in, [K, N] out
__global__ void awq_dequantize_kernel(const int32_t* __restrict__ qweight, const int32_t* __restrict__ qzeros,
                                      const __half* __restrict__ scales, int K, int N, __half* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (c >= N / 8) return;
    const uint32_t q = (uint32_t)qweight[(int64_t)k * (N / 8) + c];
    const uint32_t z = (uint32_t)qzeros[(int64_t)(k / kGroup) * (N / 8) + c];
    union { uint4 u; __half h[8]; } s, o;
    s.u = *reinterpret_cast<const uint4*>(scales + (int64_t)(k / kGroup) * N + c * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = (int)((q >> (4 * nib_pos(j))) & 0xFu) - (int)((z >> (4 * nib_pos(j))) & 0xFu);
        o.h[j] = __hmul(__int2half_rn(d), s.h[j]);
    }
    *reinterpret_cast<uint4*>(out + (int64_t)k * N + c * 8) = o.u;
}

// packed stream -> fp16 [N_total, K]  (row = packed row); thread = (row, 8 k)
__global__ void awq_unpack_kernel(const unsigned char* __restrict__ packed, int K, int N_total,
                                  __half* __restrict__ out) {
    const int kw = blockIdx.x * blockDim.x + threadIdx.x;
    const int pr = blockIdx.y;
    if (kw >= K / 8) return;
    const int G = K / kGroup;
    const int k0 = kw * 8, g = k0 / kGroup, kk = k0 % kGroup;
    const int kb = kk >> 6, ch = (kk >> 5) & 1, w = (kk >> 3) & 3;
    const int tile = pr / kTileN, r = pr % kTileN;
    const unsigned char* rec = packed + ((int64_t)tile * G + g) * kRecordStride;
    const uint32_t word = *reinterpret_cast<const uint32_t*>(rec + ((kb * 2 + ch) * kTileN + r) * 16 + w * 4);
    const uint32_t zw = *reinterpret_cast<const uint32_t*>(rec + kRecZeroOff + (r >> 3) * 4);
    const int z = (int)((zw >> (4 * (r & 7))) & 0xFu);
    const __half s = *reinterpret_cast<const __half*>(rec + kRecScaleOff + r * 2);
    union { uint4 u; __half h[8]; } o;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.h[i] = __hmul(__int2half_rn((int)((word >> (4 * nib_pos(i))) & 0xFu) - z), s);
    *reinterpret_cast<uint4*>(out + (int64_t)pr * K + k0) = o.u;
}

// ------------------------------------------------------------------------------------------------
// stream-K schedule helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long sk_begin(long long c, long long U, long long P) { return c * U / P; }
__device__ __forceinline__ int sk_owner(long long u, long long U, long long P) {
    long long c = u * P / U;
    while (c + 1 < P && sk_begin(c + 1, U, P) <= u) ++c;
    while (c > 0 && sk_begin(c, U, P) > u) --c;
    return (int)c;
}

// One 32-bit word of the packed stream = 8 consecutive k of one feature row, nibble i at position
// (0,4,1,5,2,6,3,7)[i].  Nibbles p and p+4 are extracted together as an fp16 pair:
//   p = 0: (w      & 0x000f000f) | 0x6400 = 1024 +    q      p = 1: (w      & 0x00f000f0) | 0x6400 = 1024 + 16 q
//   p = 2: (w >> 8 & 0x000f000f) | 0x6400                     p = 3: (w >> 8 & 0x00f000f0) | 0x6400
// then   even: (1024 + q) - (1024 + z)        odd: (1024 + 16 q) * (1/16) - (64 + z)
// are both EXACT (q - z), and one multiply by the group scale rounds once: fp16((q - z) * s).
template <uint32_t kMask>
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "n"(kMask), "r"(c));
    return d;
}
__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }
__device__ __forceinline__ uint32_t as_u32(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ void dequant_word(uint32_t w, uint32_t magic, __half2 ce, __half2 neg_co, __half2 s2,
                                             uint32_t* out) {
    const __half2 sixteenth = as_h2(0x2C002C00u);                 // 1/16
    const uint32_t w8 = w >> 8;
    // (t - ce) and fma(t, 1/16, -co) are exact; the multiply by the scale is the single rounding
    out[0] = as_u32(__hmul2(__hsub2(as_h2(lop3_and_or<0x000f000fu>(w, magic)), ce), s2));
    out[1] = as_u32(__hmul2(__hfma2(as_h2(lop3_and_or<0x00f000f0u>(w, magic)), sixteenth, neg_co), s2));
    out[2] = as_u32(__hmul2(__hsub2(as_h2(lop3_and_or<0x000f000fu>(w8, magic)), ce), s2));
    out[3] = as_u32(__hmul2(__hfma2(as_h2(lop3_and_or<0x00f000f0u>(w8, magic)), sixteenth, neg_co), s2));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}

struct GemmParams {
    const unsigned char* packed;   // AWQ stream (kAwq) -- else unused
    int M, K, N_total;             // N_total = output features (rows of W)
    int units_per_tile;            // K / 128 (AWQ) or K / 64 (fp16)
    int tiles;
    void* y;
    int64_t ldy;
    int mode;
    const __half* bias;
    const __half* residual;
    int64_t ld_residual;
    int vec_ok;                    // y (and the residual) can be accessed with 16-byte vectors
    int rope_vec_ok;               // same for q_out / k_cache / v_cache / cos / sin of UMB_EPI_ROPE_QKV
    unsigned long long* argmax;
    int* counters;                 // [tiles]
    float* slots;                  // [tiles + grid][MPAD][128]
    umb_rope_epilogue_t rope;      // UMB_EPI_ROPE_QKV operands (by value)
    int stages;                    // depth of the shared-memory ring
    unsigned long long* trace;     // debug timeline [grid][16] (UMB_GEMM_TRACE=1), else NULL
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t) : : "memory");
    return t;
}
// The timeline stamps are compiled in only for the `trace` variant build (-DUMB_GEMM_TRACE_BUILD=1,
// scripts/gemm_trace.py): even a predicated-off stamp costs issue slots inside the dequant loop.
#ifdef UMB_GEMM_TRACE_BUILD
#define UMB_TRACE(slot) do { if (p.trace) p.trace[(size_t)blockIdx.x * 16 + (slot)] = gtimer(); } while (0)
#else
#define UMB_TRACE(slot) do { } while (0)
#endif

template <int MPAD>
struct GemmCfg {
    static constexpr int kDq = MPAD == 16 ? UMB_DQ_GROUPS_16 : kDqGroups;           // dequant groups per CTA (AWQ)
    static constexpr int kThreadsAwq = 96 + 128 * (kDq + 1) + (UMB_X_PRODUCER ? 32 : 0);
    static constexpr int kStagesAwq = MPAD == 16 ? 10 : (MPAD == 32 ? 8 : (MPAD == 64 ? 6 : 4));
    static constexpr int kStagesF16 = MPAD <= 32 ? 8 : (MPAD == 64 ? 6 : 4);
    static constexpr int kXBytesAwq = MPAD * 256;           // [2 atoms][MPAD][128 B]
    static constexpr int kXBytesF16 = MPAD * 128;           // [MPAD][128 B]
    static constexpr int kWBytesF16 = kTileN * 128;         // [128][128 B]
    static constexpr int kPackedStride = 8576;              // record padded to a 128-byte multiple
    static constexpr int kExchBytes = MPAD * 64 * 2 < 3 * 16 * 128 * 2 ? 3 * 16 * 128 * 2 : MPAD * 64 * 2;   // epilogue staging: output tile + cos / sin rows (RoPE)
    // Independent fp32 accumulators in TMEM (columns a * MPAD): consecutive k-steps rotate through them so
    // that back-to-back tcgen05.mma never wait on each other's accumulate (a dependent 128xNx16 MMA chain
    // runs at its ~150-cycle latency, not at its 8-cycle issue rate); the epilogue sums them.
    static constexpr int kNumAcc = MPAD <= 16 ? 8 : (MPAD == 32 ? 4 : 2);
    static constexpr int kTmemACol0 = kNumAcc * MPAD < 128 ? 128 : kNumAcc * MPAD;   // A operand stages start here
    static constexpr int kTmemAStages = (kTmemCols - kTmemACol0) / 64;             // 6 (4 at MPAD = 128)
    static constexpr int kBarBytes = 512;
    static constexpr int smem_bytes(bool awq, int stages) {
        return 1024 /*align slack*/ + kBarBytes + kExchBytes +
               stages * (awq ? (kXBytesAwq + kPackedStride) : (kXBytesF16 + kWBytesF16));
    }
};

__device__ __forceinline__ float silu_f(float f) { return f / (1.f + expf(-f)); }
__device__ __forceinline__ unsigned int orderable_f32_g(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Epilogue of 16 token columns [m0, m0 + 16) for feature lane `fl` (0..127) of tile `tile`.
template <int MPAD>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int tile, int fl, int m0, const float (&v)[16],
                                               __half* exch) {
    const int n = tile * kTileN + fl;
    if (p.mode == UMB_EPI_SILU_MUL) {
        // lanes 0..63 hold `gate`, lanes 64..127 `up` of output features tile*64 + (fl & 63).  Both halves of the
        // 16-row chunk are staged as fp16 (the rounding point of the reference's gate / up tensors), every thread
        // then finishes 8 of the 16 x 64 outputs in place, and the tile leaves with one 16-byte store per thread.
        __half* g_t = exch;                    // [16][64] gate, overwritten by the outputs
        __half* u_t = exch + 16 * 64;          // [16][64] up
        {
            __half* dst = (fl < 64 ? g_t : u_t) + (fl & 63);
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[j * 64] = __float2half_rn(v[j]);
        }
        named_bar_sync(2, 128);
        {
            const int col = fl & 63, r0 = fl >> 6;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + 2 * i;
                const __half g = g_t[r * 64 + col];
                const __half sl = __float2half_rn(silu_f(__half2float(g)));
                g_t[r * 64 + col] = __hmul(sl, u_t[r * 64 + col]);
            }
        }
        named_bar_sync(2, 128);
        {
            const int row = fl >> 3, seg = (fl & 7) * 8;
            const int no = tile * 64 + seg;
            if (row < p.M - m0 && no + 8 <= p.N_total / 2) {
                __half* yp = reinterpret_cast<__half*>(p.y) + (int64_t)(m0 + row) * p.ldy + no;
                if (p.vec_ok) {
                    *reinterpret_cast<uint4*>(yp) = *reinterpret_cast<const uint4*>(g_t + row * 64 + seg);
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) yp[c] = g_t[row * 64 + seg + c];
                }
            }
        }
        named_bar_sync(2, 128);
        return;
    }
    if (p.mode == UMB_EPI_ROPE_QKV) {
        // q|k|v projection with RoPE + KV append fused: the rotation partner of feature d is d +- D/2, i.e. lane
        // fl ^ (D/2) of the same tile (128 % D == 0); the pair is exchanged through shared memory as fp16 -- the
        // rounding point where the reference holds the fp16 q / k / v tensors.
        const umb_rope_epilogue_t& r = p.rope;
        const int D = r.head_dim, half_d = D >> 1;
        const bool n_ok = n < p.N_total;
        __half h[16];
        const bool has_bias = p.bias != nullptr && n_ok;
        const __half qb = has_bias ? p.bias[n] : __float2half_rn(0.f);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            h[j] = __float2half_rn(v[j]);
            if (has_bias) h[j] = __hadd(h[j], qb);
            exch[j * kTileN + fl] = h[j];
        }
        const int base = r.n_dev ? r.n_dev[0] : 0;
        const int mcols = p.M - m0;
        if (p.rope_vec_ok) {
            // Staged form: the cos / sin rows of the chunk's 16 positions are fetched with 16-byte loads by the whole
            // group (two dependent round trips in total instead of three per token through one register), the
            // rotated tile is finished in shared memory and leaves with two 16-byte stores per thread.
            __half* cos_t = exch + 16 * kTileN;          // [16][D]
            __half* sin_t = cos_t + 16 * kTileN;         // [16][D]
            const int segs = D >> 3;                     // 16-byte segments per table row
            const int64_t* pid = r.position_ids + (r.n_dev ? base + r.row_offset : 0) + m0;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int sidx = fl + it * 128;
                const int row = sidx / segs, seg = sidx - row * segs;
                if (sidx < 16 * segs && row < mcols) {
                    const int64_t off = pid[row] * D + seg * 8;
                    *reinterpret_cast<uint4*>(cos_t + row * D + seg * 8) =
                        *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(r.cos) + off);
                    *reinterpret_cast<uint4*>(sin_t + row * D + seg * 8) =
                        *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(r.sin) + off);
                }
            }
            named_bar_sync(2, 128);
            const int head = n / D, d = n - head * D;
            const bool rot = n_ok && head < r.num_q_heads + r.num_kv_heads;
            __half o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                o[j] = h[j];
                if (rot && j < mcols) {
                    const __half pr = exch[j * kTileN + (fl ^ half_d)];
                    // (x * cos) + (rotate_half(x) * sin): two products and one sum, each rounded (no fma)
                    o[j] = __hadd_rn(__hmul_rn(h[j], cos_t[j * D + d]), __hmul_rn(d < half_d ? __hneg(pr) : pr, sin_t[j * D + d]));
                }
            }
            named_bar_sync(2, 128);                  // every partner value has been read
#pragma unroll
            for (int j = 0; j < 16; ++j) exch[j * kTileN + fl] = o[j];
            named_bar_sync(2, 128);
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int sidx = fl + it * 128, row = sidx >> 4, c0 = (sidx & 15) * 8;
                const int n0 = tile * kTileN + c0;
                if (row < mcols && n0 + 8 <= p.N_total) {
                    const int hd = n0 / D, d0 = n0 - hd * D, m = m0 + row;
                    __half* dst;
                    if (hd < r.num_q_heads) {
                        dst = reinterpret_cast<__half*>(r.q_out) + ((int64_t)m * r.num_q_heads + hd) * D;
                    } else {
                        const bool is_k = hd < r.num_q_heads + r.num_kv_heads;
                        const int kvh = hd - r.num_q_heads - (is_k ? 0 : r.num_kv_heads);
                        dst = reinterpret_cast<__half*>(is_k ? r.k_cache : r.v_cache) +
                              (int64_t)(base + r.kv_offset + m) * r.kv_tok_stride + (int64_t)kvh * r.kv_head_stride;
                    }
                    *reinterpret_cast<uint4*>(dst + d0) = *reinterpret_cast<const uint4*>(exch + row * kTileN + c0);
                }
            }
            named_bar_sync(2, 128);
            return;
        }
        named_bar_sync(2, 128);
        if (n_ok) {
            const int head = n / D, d = n - head * D;
            const bool rot = head < r.num_q_heads + r.num_kv_heads;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (j >= mcols) continue;
                const int m = m0 + j;
                __half o = h[j];
                if (rot) {
                    const int64_t pos = r.position_ids[(r.n_dev ? base + r.row_offset : 0) + m];
                    const __half c = reinterpret_cast<const __half*>(r.cos)[pos * D + d];
                    const __half sn = reinterpret_cast<const __half*>(r.sin)[pos * D + d];
                    const __half pr = exch[j * kTileN + (fl ^ half_d)];
                    // (x * cos) + (rotate_half(x) * sin): two products and one sum, each rounded (no fma)
                    o = __hadd_rn(__hmul_rn(h[j], c), __hmul_rn(d < half_d ? __hneg(pr) : pr, sn));
                }
                __half* dst;
                if (head < r.num_q_heads) {
                    dst = reinterpret_cast<__half*>(r.q_out) + ((int64_t)m * r.num_q_heads + head) * D;
                } else {
                    const bool is_k = head < r.num_q_heads + r.num_kv_heads;
                    const int kvh = head - r.num_q_heads - (is_k ? 0 : r.num_kv_heads);
                    dst = reinterpret_cast<__half*>(is_k ? r.k_cache : r.v_cache) +
                          (int64_t)(base + r.kv_offset + m) * r.kv_tok_stride + (int64_t)kvh * r.kv_head_stride;
                }
                dst[d] = o;
            }
        }
        named_bar_sync(2, 128);
        return;
    }
    const bool n_ok = n < p.N_total;
    const int mcols = p.M - m0;                       // valid token columns of this chunk (may exceed 16)
    // fp16 outputs (plain store / residual add): the 16 x 128 tile goes through shared memory so that global memory
    // is touched with two independent 16-byte accesses per thread instead of one 2-byte access per token row.
    // Written as the plain per-row loop the residual loads compile into 13 serialised L2 round trips (4.2 us on the
    // tail of every o-proj / down-proj launch) and the stores into ~500 dependent instructions
    // (profiles/r1_gemm_tail_trace.txt).  y may alias the residual: all loads of the tile complete before any store.
    if ((p.mode == UMB_EPI_STORE || p.mode == UMB_EPI_RESIDUAL) && p.vec_ok) {
        const bool resid = p.mode == UMB_EPI_RESIDUAL;
        int row[2], col[2];
        bool ok[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = fl + i * 128;
            row[i] = idx >> 4;
            col[i] = (idx & 15) * 8;
            ok[i] = row[i] < mcols && tile * kTileN + col[i] + 8 <= p.N_total;
        }
        if (resid) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (ok[i])
                    *reinterpret_cast<uint4*>(exch + row[i] * kTileN + col[i]) = *reinterpret_cast<const uint4*>(
                        p.residual + (int64_t)(m0 + row[i]) * p.ld_residual + tile * kTileN + col[i]);
            named_bar_sync(2, 128);
        }
        const __half bb = (p.bias && n_ok) ? p.bias[n] : __float2half_rn(0.f);
        __half h[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            h[j] = __float2half_rn(v[j]);
            if (p.bias) h[j] = __hadd(h[j], bb);
            if (resid) h[j] = __hadd(exch[j * kTileN + fl], h[j]);
        }
        if (resid) named_bar_sync(2, 128);            // every thread has read its residual column
#pragma unroll
        for (int j = 0; j < 16; ++j) exch[j * kTileN + fl] = h[j];
        named_bar_sync(2, 128);
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (ok[i])
                *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.y) + (int64_t)(m0 + row[i]) * p.ldy + tile * kTileN + col[i]) =
                    *reinterpret_cast<const uint4*>(exch + row[i] * kTileN + col[i]);
        named_bar_sync(2, 128);                       // the staging tile is free for the next chunk
        return;
    }
    if (!n_ok && !(p.mode == UMB_EPI_LOGITS && p.argmax)) return;
    const __half b = (p.bias && n_ok) ? p.bias[n] : __float2half_rn(0.f);
    // residual rows first, all loads in flight together: y may alias residual, so a load placed after a
    // store of the previous row would be serialised behind it (one L2 round trip per token row)
    __half res[16];
    if (p.mode == UMB_EPI_RESIDUAL) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j < mcols) res[j] = p.residual[(int64_t)(m0 + j) * p.ld_residual + n];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j >= mcols) continue;
        const int m = m0 + j;
        __half h = __float2half_rn(v[j]);
        if (p.bias) h = __hadd(h, b);
        if (p.mode == UMB_EPI_RESIDUAL) {
            h = __hadd(res[j], h);
            reinterpret_cast<__half*>(p.y)[(int64_t)m * p.ldy + n] = h;
        } else if (p.mode == UMB_EPI_LOGITS) {
            const float f = __half2float(h);
            if (n_ok) reinterpret_cast<float*>(p.y)[(int64_t)m * p.ldy + n] = f;
            if (p.argmax) {
                // warp-level max first: one atomic per (warp, token)
                unsigned long long key =
                    n_ok ? (((unsigned long long)orderable_f32_g(f) << 32) | (unsigned int)(~n)) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other > key ? other : key;
                }
                if ((threadIdx.x & 31) == 0) atomicMax(p.argmax + m, key);
            }
        } else {
            reinterpret_cast<__half*>(p.y)[(int64_t)m * p.ldy + n] = h;
        }
    }
}

// v[j] = sum over the accumulators of column (m0 + j), in a fixed order (deterministic).  The
// accumulators are split between the two MMA warps (first / second half); `use_mask` bit w tells
// whether warp w issued anything into its half during this segment.
template <int kAcc, int MPAD>
__device__ __forceinline__ void load_acc_sum(uint32_t taddr, float (&v)[16], int use_mask) {
    uint32_t rr[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
#pragma unroll
    for (int a = 0; a < kAcc; ++a) {
        if (!((use_mask >> (a / (kAcc / 2))) & 1)) continue;
        UMB_TMEM_LD_X16(taddr + a * MPAD, rr);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(rr[j]);
    }
}

// Warp roles (AWQ: 19 warps, fp16: 7 warps)
//   warp 0        TMA producer.  The weight copies of the first kStages units are issued BEFORE
//                 griddepcontrol.wait (weights never depend on the stream predecessor), the x slabs after it
//   warps 1, 2    MMA issuers: CTA-local unit u is issued by warp 1 + (u & 1), each into its own half of
//                 the TMEM accumulators
//   warps 3..14   dequant groups (AWQ only): shared memory -> registers -> tensor memory; they start as
//                 soon as a weight record has landed, also before the predecessor kernel has finished
//   last 4 warps  epilogue only (AWQ: warps 15..18, fp16: warps 3..6): the partial-tile fix-up is a chain of
//                 global round trips and must not stall the dequant stream
template <int MPAD, bool kAwq>
__global__ void __launch_bounds__(kAwq ? GemmCfg<MPAD>::kThreadsAwq : kGemmThreadsF16, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
            const GemmParams p) {
    using Cfg = GemmCfg<MPAD>;
    const int kStages = p.stages;
    constexpr int kXBytes = kAwq ? Cfg::kXBytesAwq : Cfg::kXBytesF16;
    // paired weight requests: two records of the stream land back to back, so the shared-memory stride is the record's
    constexpr bool kPair = kAwq && UMB_W_PAIR && kRecordStride == kRecordBytes;
    constexpr int kWStride = kAwq ? (kPair ? kRecordBytes : Cfg::kPackedStride) : Cfg::kWBytesF16;
    constexpr uint32_t kWTxBytes = kAwq ? (uint32_t)kRecordBytes : (uint32_t)Cfg::kWBytesF16;
    constexpr uint32_t kIdesc = umma_idesc_f16(kTileN, MPAD);
    constexpr int kNumAcc = Cfg::kNumAcc;                       // >= 2, even
    constexpr int kAccUsed = kAwq ? kNumAcc : (kNumAcc < 8 ? kNumAcc : 8);
    constexpr int kAccHalf = kAccUsed / 2;                      // accumulators per MMA warp
    constexpr int kACol0 = Cfg::kTmemACol0;
    constexpr int kAStages = Cfg::kTmemAStages;
    // the MMA warps take alternate units, so a ring stage must always come back to the same warp (parity waits, see the
    // dequant warps' note): both rings need an even depth
    static_assert(!kAwq || (kAStages % 2 == 0 && Cfg::kStagesAwq % 2 == 0), "ring depths must be even");
    constexpr int kDq = Cfg::kDq;                               // dequant groups of this CTA

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* s_x = smem;                                   // kStages x kXBytes   (1024-aligned slabs)
    unsigned char* s_w = s_x + kStages * kXBytes;                // kStages x kWStride
    __half* s_exch = reinterpret_cast<__half*>(s_w + kStages * kWStride);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_exch) + Cfg::kExchBytes);
    uint64_t* full_w = bars;                  // [kStages]   weight record / tile landed
    uint64_t* full_x = bars + 16;             // [kStages]   x slab landed
    uint64_t* done = bars + 32;               // [kStages]   every MMA of the unit in this stage has retired
    uint64_t* a_full = bars + 48;             // [kAStages]  dequantised A operand is in tensor memory
    uint64_t* acc_full = bars + 56;
    uint64_t* acc_empty = bars + 57;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 58);
    int* s_flag = reinterpret_cast<int*>(bars + 59);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) UMB_TRACE(0);
    const long long U = (long long)p.tiles * p.units_per_tile, P = gridDim.x;
    const long long u_begin = sk_begin(blockIdx.x, U, P), u_end = sk_begin(blockIdx.x + 1, U, P);
    const int n_units = (int)(u_end - u_begin);
    const int in_tile0 = (int)u_begin % p.units_per_tile;        // position of the first unit inside its tile
    const int tile0 = (int)u_begin / p.units_per_tile;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        if (!kAwq || UMB_X_3D) tma_prefetch_desc(&tmap_w);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(full_w + i, 1);
            mbar_init(full_x + i, 1);
            mbar_init(done + i, 1);
        }
        for (int i = 0; i < kAStages; ++i) mbar_init(a_full + i, 4);
        mbar_init(acc_full, 2);
        mbar_init(acc_empty, 4);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(s_tmem, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();
    const uint32_t tmem_base = *s_tmem;
    if (threadIdx.x == 0) UMB_TRACE(1);

    if (warp == 0) {
        // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
        const uint64_t pol_w = l2_policy_evict_first(), pol_x = l2_policy_evict_last();
        const unsigned char* src = p.packed + (kAwq ? u_begin * kRecordStride : 0);
        const int n_pre = n_units < kStages ? n_units : kStages;
        // The weight copies of the whole ring are requested before griddepcontrol.wait (they never depend on the
        // stream predecessor), the x slabs after it.  (Requesting only a few records early, so that the first x
        // slab does not queue behind ~85 KB of weights, starts the first MMA 2 us sooner but measured no faster:
        // the dequant warps set the pace.)
        const int n_early = n_pre;
        int tile = tile0, g = in_tile0;                          // (tile, g) of the next weight copy
        auto issue_w = [&](int stage) {
            mbar_arrive_expect_tx(full_w + stage, kWTxBytes);
            if constexpr (kAwq)
                tma_bulk_g2s(s_w + stage * kWStride, src, kRecordBytes, full_w + stage, pol_w);
            else
                tma_tensor2d_g2s(s_w + stage * kWStride, &tmap_w, g * 64, tile * kTileN, full_w + stage, pol_w);
        };
        auto issue_x = [&](int stage, int gk) {
            if (kAwq && UMB_X_PRODUCER) return;                   // requested by the x-producer warp
            mbar_arrive_expect_tx(full_x + stage, (uint32_t)kXBytes);
            unsigned char* xs = s_x + stage * kXBytes;
            if constexpr (kAwq && UMB_X_3D) {
                tma_tensor3d_g2s(xs, &tmap_w, 0, 0, gk * 2, full_x + stage, pol_x);      // tmap_w = x as [64, rows, K / 64]
            } else if constexpr (kAwq) {
                tma_tensor2d_g2s(xs, &tmap_x, gk * kGroup, 0, full_x + stage, pol_x);
                tma_tensor2d_g2s(xs + MPAD * 128, &tmap_x, gk * kGroup + 64, 0, full_x + stage, pol_x);
            } else {
                tma_tensor2d_g2s(xs, &tmap_x, gk * 64, 0, full_x + stage, pol_x);
            }
        };
        auto advance_w = [&]() {
            src += kRecordStride;
            if (++g == p.units_per_tile) {
                g = 0;
                ++tile;
            }
        };
        if constexpr (kPair) {
            // One request per PAIR of units (17 KB): a thread gets a cp.async.bulk out every ~230 ns whatever its size
            // (scripts/ubench/tma_rates.cu), so single records cap an SM at 37 GB/s (5.4 TB/s for the chip).  Stages
            // 2j and 2j + 1 share the barrier full_w[2j]; kStages is even and the ring starts at stage 0, so pairs
            // never straddle the wrap.
            auto issue_pair = [&](int stage, int cnt) {
                mbar_arrive_expect_tx(full_w + stage, (uint32_t)(kRecordBytes * cnt));
                tma_bulk_g2s(s_w + stage * kWStride, src, (uint32_t)(kRecordBytes * cnt), full_w + stage, pol_w);
            };
            for (int i = 0; i < n_pre; i += 2) {                 // weights first: independent of the predecessor
                const int cnt = n_pre - i >= 2 ? 2 : 1;
                if (elect_one()) issue_pair(i, cnt);
                __syncwarp();
                src += (size_t)kRecordStride * cnt;
            }
            if (lane == 0) UMB_TRACE(2);
            pdl_wait();                                          // x is the predecessor's output
            if (lane == 0) UMB_TRACE(3);
            int gx = in_tile0;
            for (int i = 0; i < n_pre; ++i) {
                if (elect_one()) issue_x(i, gx);
                __syncwarp();
                if (++gx == p.units_per_tile) gx = 0;
            }
            int stage = 0, use = 1;
            for (int i = n_pre; i < n_units; i += 2) {
                const int cnt = n_units - i >= 2 ? 2 : 1;
                UMB_WAIT(done + stage, (use - 1) & 1);          // the units that last used these stages have retired
                if (cnt == 2) UMB_WAIT(done + stage + 1, (use - 1) & 1);
                if (elect_one()) {
                    issue_pair(stage, cnt);
                    issue_x(stage, gx);
                    if (cnt == 2) issue_x(stage + 1, gx + 1 == p.units_per_tile ? 0 : gx + 1);
                }
                __syncwarp();
                src += (size_t)kRecordStride * cnt;
                for (int c = 0; c < cnt; ++c)
                    if (++gx == p.units_per_tile) gx = 0;
                stage += 2;
                if (stage >= kStages) {
                    stage = 0;
                    ++use;
                }
            }
        } else {
        for (int i = 0; i < n_early; ++i) {                      // weights first: independent of the predecessor
            if (elect_one()) issue_w(i);
            __syncwarp();
            advance_w();
        }
        if (lane == 0) UMB_TRACE(2);
        pdl_wait();                                              // x is the predecessor's output
        if (lane == 0) UMB_TRACE(3);
        int gx = in_tile0;
        for (int i = 0; i < n_early; ++i) {
            if (elect_one()) issue_x(i, gx);
            __syncwarp();
            if (++gx == p.units_per_tile) gx = 0;
        }
        for (int i = n_early; i < n_pre; ++i) {                  // rest of the ring: stages still untouched
            if (elect_one()) {
                issue_w(i);
                issue_x(i, g);
            }
            __syncwarp();
            advance_w();
        }
        Ring r;
        r.phase = 1;                                             // every stage has been used once
        for (int i = n_pre; i < n_units; ++i) {
            UMB_WAIT(done + r.stage, r.phase ^ 1);              // the unit that last used this stage has retired
            if (elect_one()) {
                issue_w(r.stage);
                issue_x(r.stage, g);
            }
            __syncwarp();
            advance_w();
            r.advance(kStages);
        }
        }
        if (lane == 0) UMB_TRACE(10);
    } else if (kAwq && UMB_X_PRODUCER && warp == GemmCfg<MPAD>::kThreadsAwq / 32 - 1) {
        // ===================== x-slab producer (AWQ, UMB_X_PRODUCER): two tensor copies per unit =====================
        const uint64_t pol_x = l2_policy_evict_last();
        pdl_wait();                                              // x is the predecessor's output
        int gx = in_tile0, stage = 0, use = 0;
        for (int i = 0; i < n_units; ++i) {
            if (use > 0) UMB_WAIT(done + stage, (use - 1) & 1); // the unit that last used this stage has retired
            if (elect_one()) {
                mbar_arrive_expect_tx(full_x + stage, (uint32_t)kXBytes);
                unsigned char* xs = s_x + stage * kXBytes;
                if constexpr (UMB_X_3D) {
                    tma_tensor3d_g2s(xs, &tmap_w, 0, 0, gx * 2, full_x + stage, pol_x);
                } else {
                    tma_tensor2d_g2s(xs, &tmap_x, gx * kGroup, 0, full_x + stage, pol_x);
                    tma_tensor2d_g2s(xs + MPAD * 128, &tmap_x, gx * kGroup + 64, 0, full_x + stage, pol_x);
                }
            }
            __syncwarp();
            if (++gx == p.units_per_tile) gx = 0;
            if (++stage == kStages) {
                stage = 0;
                ++use;
            }
        }
    } else if (warp <= 2) {
        // ===================== MMA issuers: warp 1 takes even CTA-local units, warp 2 odd ones =====================
        const int par = warp - 1;
        const uint32_t d_base = tmem_base + par * kAccHalf * MPAD;   // this warp's accumulators
        constexpr uint64_t kDescHi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
        int st = par % kStages, ph = (par / kStages) & 1;            // smem ring position of unit `u`
        int ts = par % kAStages, pa = 0;                              // tensor-memory A ring position
        int u = par, seg = 0, seg_begin = 0, in_tile = in_tile0;
        while (seg_begin < n_units) {
            const int seg_end = seg_begin + min(n_units - seg_begin, p.units_per_tile - in_tile);
            UMB_WAIT_LONG(acc_empty, (seg & 1) ^ 1);
            tc_fence_after();
            bool first = true;
            for (; u < seg_end; u += 2) {
                const bool last = u + 2 >= seg_end;
                if constexpr (kAwq) {
                    UMB_WAIT(a_full + ts, pa);
                } else {
                    UMB_WAIT(full_w + st, ph);
                }
                UMB_WAIT(full_x + st, ph);
                tc_fence_after();
                if (u == 0 && lane == 0) UMB_TRACE(5);
                const uint32_t xs = smem_u32(s_x + st * kXBytes);
                const uint64_t bd0 = kDescHi | (uint64_t)((xs & 0x3FFFF) >> 4);
                if constexpr (kAwq) {
                    const uint32_t a_col = tmem_base + kACol0 + ts * 64;
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint64_t bd = bd0 + (uint64_t)(((kk >> 2) * (MPAD * 128) + (kk & 3) * 32) >> 4);
                            umma_ts_f16(d_base + (kk % kAccHalf) * MPAD, a_col + kk * 8, bd, kIdesc,
                                        (first && kk < kAccHalf) ? 0u : 1u);
                        }
                        umma_commit(done + st);
                        if (last) umma_commit(acc_full);
                    }
                } else {
                    const uint32_t ws = smem_u32(s_w + st * kWStride);
                    const uint64_t ad0 = kDescHi | (uint64_t)((ws & 0x3FFFF) >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_ss_f16(d_base + (kk % kAccHalf) * MPAD, ad0 + kk * 2, bd0 + kk * 2, kIdesc,
                                        (first && kk < kAccHalf) ? 0u : 1u);
                        umma_commit(done + st);
                        if (last) umma_commit(acc_full);
                    }
                }
                __syncwarp();
                first = false;
                st += 2;
                while (st >= kStages) {
                    st -= kStages;
                    ph ^= 1;
                }
                ts += 2;
                if (ts >= kAStages) {
                    ts -= kAStages;
                    pa ^= 1;
                }
            }
            if (first) {                     // this warp had no unit in the segment: plain arrival
                if (elect_one()) mbar_arrive(acc_full);
                __syncwarp();
            }
            seg_begin = seg_end;
            in_tile = 0;
            ++seg;
        }
        if (par == 0 && lane == 0) UMB_TRACE(6);
    } else {
        const int q = warp & 3;                      // TMEM lane quarter this warp may touch
        const int grp = (warp - 3) >> 2;             // AWQ: dequant group 0..kDqGroups-1, then the epilogue group
        const bool is_epi = kAwq ? (grp == kDq) : true;
        const int fl = q * 32 + lane;                // feature lane 0..127
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        if (!is_epi) {
            // ===================== dequant warps (AWQ): group g takes CTA-local units g, g + kDqGroups, ... =====================
            if constexpr (kAwq) {
                int dq_st = grp % kStages, dq_ph = 0, dq_ts = grp % kAStages;      // ring positions of unit `mine`
                int w_first = grp;                                                 // first unit of this group that re-uses a TMEM stage
                while (w_first < kAStages) w_first += kDq;
                int w_st = (w_first - kAStages) % kStages, w_ph = 0;               // ring position of unit `mine - kAStages`
                // the packed words / scale / zero of a unit are pulled into registers as soon as its record has landed
                uint4 w4[4];
                uint32_t sbits = 0, zw = 0;
                auto fetch = [&](int st_) {
                    const uint32_t rec = smem_u32(s_w + st_ * kWStride);
#pragma unroll
                    for (int c = 0; c < 4; ++c) w4[c] = lds128(rec + (c * kTileN + fl) * 16);    // 128 k of this feature row
                    sbits = lds16(rec + kRecScaleOff + fl * 2);
                    zw = lds32(rec + kRecZeroOff + (fl >> 3) * 4);
                };
                if (grp < n_units) {
                    UMB_WAIT(full_w + (kPair ? (dq_st & ~1) : dq_st), dq_ph);
                    if (grp == 0 && q == 0 && lane == 0) UMB_TRACE(4);
                    fetch(dq_st);
                }
                for (int mine = grp; mine < n_units; mine += kDq) {
                    const int ts = dq_ts;
                    dq_st += kDq;
                    if (dq_st >= kStages) {
                        dq_st -= kStages;
                        dq_ph ^= 1;
                    }
                    dq_ts += kDq;
                    if (dq_ts >= kAStages) dq_ts -= kAStages;
                    const uint32_t z = (zw >> (4 * (fl & 7))) & 0xFu;
                    const __half2 ce = as_h2((0x6400u | z) * 0x10001u);                 // (1024 + z) x2
                    const __half2 neg_co = as_h2((0xD400u | (z << 4)) * 0x10001u);      // -(64 + z) x2
                    const __half2 s2 = as_h2(sbits * 0x10001u);
                    const uint32_t magic = 0x64006400u;
                    const uint32_t taddr = tmem_base + lane_addr + kACol0 + ts * 64;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {                            // 4 chunks of 32 k (16 TMEM columns)
                        const uint32_t words[4] = {w4[c].x, w4[c].y, w4[c].z, w4[c].w};
                        uint32_t a[16];
#pragma unroll
                        for (int w = 0; w < 4; ++w) dequant_word(words[w], magic, ce, neg_co, s2, &a[w * 4]);
                        UMB_TMEM_ST_X16(taddr + c * 16, a);
                    }
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full + ts);
                    if (mine + kDq < n_units) {            // this group's next unit v: (dq_st, dq_ph) already point at it
                        // (1) The unit that last used v's tensor-memory stage (v - kAStages) must have retired before the
                        //     stage is overwritten.
                        // (2) The same wait makes the parity wait on full_w sound.  A parity wait lets a waiter through that
                        //     is TWO uses ahead of its barrier (the parity of "use j - 2 completed" equals that of "use j
                        //     completed"), and the previous occupant of v's shared-memory stage, unit v - kStages, belongs to
                        //     another group: with the 4-deep ring of the 128-column instantiation it is unit mine - 1, which
                        //     may still be in flight.  The group then read a stale record and signalled a_full early, the MMA
                        //     retired early, and the producer re-armed full_w while its previous copy was still open: the
                        //     "Warp Illegal Instruction" that killed one 128-row launch in ~150 (profiles/r2_m128_root_cause.txt).
                        //     done[v - kAStages] implies done[v - kStages]: both ring depths are even, so the two units were
                        //     issued and committed, in order, by the same MMA thread; a retired occupant has landed, so
                        //     full_w is at most one use behind.  (v < kAStages: first use of both stages, nothing to wait for.)
                        if (mine + kDq >= kAStages) {
                            UMB_WAIT(done + w_st, w_ph);
                            tc_fence_after();
                            w_st += kDq;
                            if (w_st >= kStages) {
                                w_st -= kStages;
                                w_ph ^= 1;
                            }
                        }
                        UMB_WAIT(full_w + (kPair ? (dq_st & ~1) : dq_st), dq_ph);
                        fetch(dq_st);
                    }
                }
            }
        } else {
            // ===================== epilogue warps =====================
            const int ew = kAwq ? warp - 3 - 4 * kDq : warp - 3;       // 0..3 inside the epilogue group
            pdl_wait();                                  // residual / workspace / y belong to the predecessors
            if (ew == 0 && lane == 0) UMB_TRACE(11);
            int seg = 0, seg_begin = 0, in_tile = in_tile0, tile = tile0;
            // acc[j] += the partial sums the other CTAs of `tile_` published for token columns [m0, m0 + 16), in
            // ascending CTA order (deterministic); the polled words are stored back as zeros.
            auto reduce_others = [&](float (&acc)[16], int m0, int tile_, int c_first, int n_other) {
                const int mcols = p.M - m0;
                uint32_t* sl0 = reinterpret_cast<uint32_t*>(p.slots) +
                                ((size_t)tile_ + c_first + 1) * (MPAD * kTileN) + m0 * kTileN + fl;
                if (mcols <= 4) {
                    // few token rows (draft levels): 8 slots in flight per round trip
                    for (int c0 = 0; c0 < n_other; c0 += 8) {
                        uint32_t t[8][4];
                        bool pending;
                        do {
                            pending = false;
#pragma unroll
                            for (int cc = 0; cc < 8; ++cc)
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const bool need = c0 + cc < n_other && j < mcols;
                                    t[cc][j] = need ? ld_relaxed_gpu_u32(sl0 + (size_t)(c0 + cc) * (MPAD * kTileN) + j * kTileN) : 0x80000000u;
                                }
#pragma unroll
                            for (int cc = 0; cc < 8; ++cc)
#pragma unroll
                                for (int j = 0; j < 4; ++j) pending |= t[cc][j] == 0u;
                        } while (pending);
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc[j] += __uint_as_float(t[cc][j]);
                                if (c0 + cc < n_other && j < mcols) sl0[(size_t)(c0 + cc) * (MPAD * kTileN) + j * kTileN] = 0u;
                            }
                    }
                } else if (mcols > 0) {
                    for (int c0 = 0; c0 < n_other; c0 += 2) {
                        uint32_t t[2][16];
                        bool pending;
                        do {
                            pending = false;
#pragma unroll
                            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const bool need = c0 + cc < n_other && j < mcols;
                                    t[cc][j] = need ? ld_relaxed_gpu_u32(sl0 + (size_t)(c0 + cc) * (MPAD * kTileN) + j * kTileN) : 0x80000000u;
                                }
#pragma unroll
                            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                                for (int j = 0; j < 16; ++j) pending |= t[cc][j] == 0u;
                        } while (pending);
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                acc[j] += __uint_as_float(t[cc][j]);
                                if (c0 + cc < n_other && j < mcols) sl0[(size_t)(c0 + cc) * (MPAD * kTileN) + j * kTileN] = 0u;
                            }
                    }
                }
            };
            while (seg_begin < n_units) {
                const int seg_end = seg_begin + min(n_units - seg_begin, p.units_per_tile - in_tile);
                const int seg_len = seg_end - seg_begin;
                const bool whole = seg_len == p.units_per_tile;
                // which MMA warps issued into their accumulators during this segment
                const int n_odd = (seg_end >> 1) - (seg_begin >> 1);               // odd CTA-local indices in [begin, end)
                const int use_mask = ((seg_len - n_odd) > 0 ? 1 : 0) | (n_odd > 0 ? 2 : 0);
                const long long tile_u0 = (long long)tile * p.units_per_tile;
                const int c_first = whole ? (int)blockIdx.x : sk_owner(tile_u0, U, P);
                const int c_last = whole ? (int)blockIdx.x : sk_owner(tile_u0 + p.units_per_tile - 1, U, P);
                const bool reducer = !whole && (int)blockIdx.x == c_first;
                UMB_WAIT_LONG(acc_full, seg & 1);
                tc_fence_after();
                if (ew == 0 && lane == 0) UMB_TRACE(seg == 0 ? 7 : 12);
                if (whole) {
#pragma unroll 1
                    for (int m0 = 0; m0 < MPAD; m0 += 16) {
                        float v[16];
                        load_acc_sum<kAccUsed, MPAD>(tmem_base + lane_addr + m0, v, use_mask);
                        if (m0 + 16 >= MPAD) {       // accumulators fully read: hand TMEM back to the MMA warps
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(acc_empty);
                        }
                        if (m0 < p.M || p.mode == UMB_EPI_SILU_MUL || p.mode == UMB_EPI_ROPE_QKV)
                            epilogue_chunk<MPAD>(p, tile, fl, m0, v, s_exch);
                    }
                } else {
                    // Partial tile.  The CTA that owns the HEAD of the tile (c_first; that part is the last
                    // segment of its range) is the reducer; every other part is the FIRST segment of its owner,
                    // i.e. published early in the kernel.  Publication needs no fence, counter or barrier: every
                    // 32-bit word of a slot is its own flag (0 = not yet written; a partial sum of +0.0f is stored
                    // as -0.0f, which adds identically), the reducer polls exactly the words it needs with relaxed
                    // gpu-scope loads, adds them (ascending CTA order -- deterministic) and stores the zeros back,
                    // which leaves the workspace clean for the next launch.
                    if (!reducer) {
                        uint32_t* slot = reinterpret_cast<uint32_t*>(p.slots) + ((size_t)tile + blockIdx.x) * (MPAD * kTileN);
#pragma unroll 1
                        for (int m0 = 0; m0 < MPAD; m0 += 16) {
                            float v[16];
                            load_acc_sum<kAccUsed, MPAD>(tmem_base + lane_addr + m0, v, use_mask);
                            if (m0 + 16 >= MPAD) {
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(acc_empty);
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (m0 + j < p.M) {
                                    const uint32_t u = __float_as_uint(v[j]);
                                    slot[(m0 + j) * kTileN + fl] = u ? u : 0x80000000u;
                                }
                        }
                    } else {
                        const int n_other = c_last - c_first;
#pragma unroll 1
                        for (int m0 = 0; m0 < MPAD; m0 += 16) {
                            float v[16];
                            load_acc_sum<kAccUsed, MPAD>(tmem_base + lane_addr + m0, v, use_mask);
                            if (m0 + 16 >= MPAD) {
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(acc_empty);
                            }
                            if (!(m0 < p.M || p.mode == UMB_EPI_SILU_MUL || p.mode == UMB_EPI_ROPE_QKV)) continue;
                            if (ew == 0 && lane == 0 && m0 == 0) UMB_TRACE(14);
                            reduce_others(v, m0, tile, c_first, n_other);
                            if (ew == 0 && lane == 0 && m0 == 0) UMB_TRACE(15);
                            epilogue_chunk<MPAD>(p, tile, fl, m0, v, s_exch);
                        }
                    }
                }
                seg_begin = seg_end;
                in_tile = 0;
                ++tile;
                ++seg;
                if (ew == 0 && lane == 0) UMB_TRACE(seg == 1 ? 8 : 13);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) UMB_TRACE(9);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// probe: one 128 x 16 x 64 UMMA, A from shared memory (mode 0) or tensor memory (mode 1)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
probe_umma_kernel(const __half* __restrict__ a, const __half* __restrict__ b, float* __restrict__ d, int mode) {
    __shared__ __align__(1024) unsigned char s_a[128 * 128];
    __shared__ __align__(1024) unsigned char s_b[16 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
    // manual SW128 K-major tiles: element (row, k) at (row/8)*1024 + (row%8)*128 + ((k/8) ^ (row%8))*16 + (k%8)*2
    for (int c = 0; c < 8; ++c) {
        const uint4 va = *reinterpret_cast<const uint4*>(a + r * 64 + c * 8);
        *reinterpret_cast<uint4*>(s_a + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = va;
        if (r < 16) {
            const uint4 vb = *reinterpret_cast<const uint4*>(b + r * 64 + c * 8);
            *reinterpret_cast<uint4*>(s_b + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = vb;
        }
    }
    fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&s_tmem, 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    if (mode == 1) {
        uint32_t regs[32];
        for (int c = 0; c < 32; ++c) regs[c] = *reinterpret_cast<const uint32_t*>(a + r * 64 + c * 2);
        UMB_TMEM_ST_X16(tmem + lane_addr + 32, (&regs[0]));
        UMB_TMEM_ST_X16(tmem + lane_addr + 48, (&regs[16]));
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(128, 16);
        for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bd = umma_smem_desc_sw128(smem_u32(s_b) + kk * 32);
            if (mode == 1) {
                umma_ts_f16(tmem, tmem + 32 + kk * 8, bd, idesc, kk ? 1u : 0u);
            } else {
                const uint64_t ad = umma_smem_desc_sw128(smem_u32(s_a) + kk * 32);
                umma_ss_f16(tmem, ad, bd, idesc, kk ? 1u : 0u);
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t rr[16];
    UMB_TMEM_LD_X16(tmem + lane_addr, rr);
    tc_wait_ld();
    for (int j = 0; j < 16; ++j) d[r * 16 + j] = __uint_as_float(rr[j]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 64);
    }
}

// probe: D[128, 128] = A[128, 64 k] * Bt[64 k, 128 n] with the B operand MN-major (n contiguous in memory, as the V rows
// of a KV cache are): Bt is staged as two 64-n atoms of [64 k rows][128 B], 128-byte swizzled -- the image a TMA box
// [64 n, 64 k rows] would leave -- and described with LBO = the atom stride along n, SBO = 1024 B between 8-row k groups
// (mode 2), or with the two swapped (mode 3: the reading that must NOT work).
__global__ void __launch_bounds__(128, 1)
probe_umma_mn_kernel(const __half* __restrict__ a, const __half* __restrict__ bt, float* __restrict__ d, int mode) {
    __shared__ __align__(1024) unsigned char s_a[128 * 128];
    __shared__ __align__(1024) unsigned char s_b[2 * 64 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, r = threadIdx.x;
    for (int c = 0; c < 8; ++c) {
        const uint4 va = *reinterpret_cast<const uint4*>(a + r * 64 + c * 8);
        *reinterpret_cast<uint4*>(s_a + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = va;
    }
    for (int i = r; i < 2 * 64 * 8; i += 128) {                 // (atom, k row, 16-byte chunk)
        const int atom = i / (64 * 8), row = (i / 8) % 64, c = i % 8;
        const uint4 vb = *reinterpret_cast<const uint4*>(bt + row * 128 + atom * 64 + c * 8);
        *reinterpret_cast<uint4*>(s_b + atom * 8192 + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4)) = vb;
    }
    fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&s_tmem, 128);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(128, 128) | (1u << 16);      // B MN-major
        const uint64_t lbo = mode == 2 ? 8192 : 1024, sbo = mode == 2 ? 1024 : 8192;
        for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = umma_smem_desc_sw128(smem_u32(s_a) + kk * 32);
            uint64_t bd = 0;
            bd |= (uint64_t)(((smem_u32(s_b) + kk * 2048) & 0x3FFFF) >> 4);    // 16 k rows = two 8-row groups further
            bd |= (uint64_t)(lbo >> 4) << 16;
            bd |= (uint64_t)(sbo >> 4) << 32;
            bd |= (uint64_t)1 << 46;
            bd |= (uint64_t)2 << 61;
            umma_ss_f16(tmem, ad, bd, idesc, kk ? 1u : 0u);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t rr[16];
        UMB_TMEM_LD_X16(tmem + lane_addr + c0, rr);
        tc_wait_ld();
        for (int j = 0; j < 16; ++j) d[r * 128 + c0 + j] = __uint_as_float(rr[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 128);
    }
}

// ------------------------------------------------------------------------------------------------
// SIMT cross-check of the packed stream (tests only): one CTA per tile, thread = feature row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
awq_gemm_simt_kernel(const unsigned char* __restrict__ packed, const __half* __restrict__ x, int64_t ldx,
                     __half* __restrict__ y, int64_t ldy, int M, int K, int N_total) {
    const int tile = blockIdx.x, r = threadIdx.x, G = K / kGroup;
    const int n = tile * kTileN + r;
    for (int m0 = 0; m0 < M; m0 += 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int g = 0; g < G; ++g) {
            const unsigned char* rec = packed + ((int64_t)tile * G + g) * kRecordStride;
            const uint32_t zw = *reinterpret_cast<const uint32_t*>(rec + kRecZeroOff + (r >> 3) * 4);
            const int z = (int)((zw >> (4 * (r & 7))) & 0xFu);
            const __half s = *reinterpret_cast<const __half*>(rec + kRecScaleOff + r * 2);
            for (int c = 0; c < 4; ++c) {
                const uint4 w4 = *reinterpret_cast<const uint4*>(rec + (c * kTileN + r) * 16);
                const uint32_t words[4] = {w4.x, w4.y, w4.z, w4.w};
                for (int w = 0; w < 4; ++w)
                    for (int i = 0; i < 8; ++i) {
                        const float wf = __half2float(__hmul(__int2half_rn((int)((words[w] >> (4 * nib_pos(i))) & 0xFu) - z), s));
                        const int k = g * kGroup + c * 32 + w * 8 + i;
                        for (int j = 0; j < 8; ++j)
                            if (m0 + j < M) acc[j] = fmaf(__half2float(x[(int64_t)(m0 + j) * ldx + k]), wf, acc[j]);
                    }
            }
        }
        if (n < N_total)
            for (int j = 0; j < 8; ++j)
                if (m0 + j < M) y[(int64_t)(m0 + j) * ldy + n] = __float2half_rn(acc[j]);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(sym);
    }
    return fn;
}
// fp16 row-major [rows, cols] (leading dimension ld elements) -> tiled map, box [64 cols, box_rows], SW128
static int make_tmap(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return UMB_ERR_DRIVER;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld * sizeof(__half)};
    const cuuint32_t box[2] = {64, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return arg_fail(__FILE__, __LINE__, "cuTensorMapEncodeTiled");
    return UMB_OK;
}

// fp16 row-major x [rows, cols] as (64 cols, rows, cols / 64 atoms): box [64, box_rows, 2] = the two 128-byte-swizzled
// atoms of one 128-k slab in ONE copy
static int make_tmap_x3d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return UMB_ERR_DRIVER;
    const cuuint64_t dims[3] = {64, rows, cols / 64};
    const cuuint64_t strides[2] = {ld * sizeof(__half), 64 * sizeof(__half)};
    const cuuint32_t box[3] = {64, box_rows, 2};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return arg_fail(__FILE__, __LINE__, "cuTensorMapEncodeTiled(3d)");
    return UMB_OK;
}

// debug timeline (UMB_GEMM_TRACE=1): one device buffer, 16 stamps per CTA of the LAST gemm launch
static unsigned long long* trace_buffer() {
    static unsigned long long* buf = nullptr;
    static int tried = 0;
    if (!tried) {
        tried = 1;
        const char* e = getenv("UMB_GEMM_TRACE");
        if (e && e[0] == '1' && cudaMalloc(&buf, 1024 * 16 * sizeof(unsigned long long)) != cudaSuccess) buf = nullptr;
        if (buf) cudaMemset(buf, 0, 1024 * 16 * sizeof(unsigned long long));
    }
    return buf;
}

// UMB_GEMM_TRACE_LAUNCH=k: stamp only the k-th launch (0-based) after the last umb_debug_gemm_trace read, so that a
// launch in the middle of a back-to-back chain can be looked at without its neighbours overwriting the stamps
static int g_trace_launches = 0;
static unsigned long long* trace_for_this_launch() {
    unsigned long long* buf = trace_buffer();
    if (!buf) return nullptr;
    static int want = -2;
    if (want == -2) {
        const char* e = getenv("UMB_GEMM_TRACE_LAUNCH");
        want = e ? atoi(e) : -1;
    }
    const int idx = g_trace_launches++;
    return (want < 0 || idx == want) ? buf : nullptr;
}

// Rows one launch takes (internal.h: UMB_MAX_ROWS; 64 splits 65..128 rows into two launches as round 1 did).
constexpr int kMaxRowsPerLaunch = UMB_MAX_ROWS;
static int mpad_of(int M) { return M <= 16 ? 16 : (M <= 32 ? 32 : (M <= 64 ? 64 : 128)); }

template <int MPAD, bool kAwq>
static int launch_gemm(const CUtensorMap& tx, const CUtensorMap& tw, GemmParams p, int grid, cudaStream_t st) {
    p.stages = kAwq ? GemmCfg<MPAD>::kStagesAwq : GemmCfg<MPAD>::kStagesF16;
    {
        static int env_stages = -1;
        if (env_stages < 0) {
            const char* e = getenv("UMB_GEMM_STAGES");
            env_stages = e ? atoi(e) : 0;
        }
        if (env_stages > 0 && env_stages <= 16 && (env_stages % 2) == 0 && GemmCfg<MPAD>::smem_bytes(kAwq, env_stages) <= 227 * 1024)
            p.stages = env_stages;                                   // even: see the static_assert on the ring depths
    }
    p.trace = trace_for_this_launch();
    const int smem = GemmCfg<MPAD>::smem_bytes(kAwq, p.stages);
    // per-device attribute (a process may drive several GPUs / threads): set on every launch, it is a cheap host call
    UMB_CUDA(cudaFuncSetAttribute(gemm_kernel<MPAD, kAwq>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    UMB_LAUNCH_PDL(gemm_kernel<MPAD, kAwq>, dim3(grid), dim3(kAwq ? GemmCfg<MPAD>::kThreadsAwq : kGemmThreadsF16), (size_t)smem, st, tx,
                   tw, p);
    count_launches(1);
    return UMB_OK;
}

static int gemm_common(bool awq, const void* x, int64_t ldx, const void* w, int64_t ldw, void* y, int64_t ldy, int M,
                       int K, int N_total, const umb_epilogue_t* epi, void* workspace, size_t workspace_bytes,
                       cudaStream_t st) {
    UMB_CHECK_ARG(x && w && y && M >= 1 && K >= 1 && N_total >= 1);
    UMB_CHECK_ARG((((uintptr_t)x | (uintptr_t)w) & 15) == 0 && (ldx % 8) == 0 && ldx >= K);
    if (M > 128) return UMB_ERR_UNSUPPORTED;
    if (M > kMaxRowsPerLaunch) {
        // only with -DUMB_MAX_ROWS=64: 65..128 rows run as two launches of <= 64 rows
        const int m0 = kMaxRowsPerLaunch;
        umb_epilogue_t e0{}, e1{};
        umb_rope_epilogue_t r1{};
        if (epi) {
            e0 = *epi;
            e1 = *epi;
            const size_t esz = (e1.mode == UMB_EPI_LOGITS) ? sizeof(float) : sizeof(__half);
            (void)esz;
            if (e1.residual) e1.residual = (const __half*)e1.residual + (int64_t)m0 * e1.ld_residual;
            if (e1.argmax) e1.argmax += m0;
            if (e1.mode == UMB_EPI_ROPE_QKV && e1.rope) {
                r1 = *e1.rope;
                r1.q_out = (__half*)r1.q_out + (int64_t)m0 * r1.num_q_heads * r1.head_dim;
                if (r1.n_dev) r1.row_offset += m0; else r1.position_ids += m0;
                r1.kv_offset += m0;
                e1.rope = &r1;
            }
        }
        const bool f32 = epi && epi->mode == UMB_EPI_LOGITS;
        int rc = gemm_common(awq, x, ldx, w, ldw, y, ldy, m0, K, N_total, epi ? &e0 : nullptr, workspace, workspace_bytes, st);
        if (rc != UMB_OK) return rc;
        void* y1 = f32 ? (void*)((float*)y + (int64_t)m0 * ldy) : (void*)((__half*)y + (int64_t)m0 * ldy);
        return gemm_common(awq, (const __half*)x + (int64_t)m0 * ldx, ldx, w, ldw, y1, ldy, M - m0, K, N_total,
                           epi ? &e1 : nullptr, workspace, workspace_bytes, st);
    }
    if (awq ? (K % kGroup) != 0 : (K % 64) != 0) return UMB_ERR_UNSUPPORTED;
    if (N_total % 8) return UMB_ERR_UNSUPPORTED;
    GemmParams p{};
    p.packed = awq ? (const unsigned char*)w : nullptr;
    p.M = M;
    p.K = K;
    p.N_total = N_total;
    p.units_per_tile = awq ? K / kGroup : K / 64;
    p.tiles = (N_total + kTileN - 1) / kTileN;
    p.y = y;
    p.ldy = ldy;
    p.mode = epi ? epi->mode : UMB_EPI_STORE;
    p.bias = epi ? (const __half*)epi->bias : nullptr;
    p.residual = epi ? (const __half*)epi->residual : nullptr;
    p.ld_residual = epi ? epi->ld_residual : 0;
    p.argmax = epi ? epi->argmax : nullptr;

    UMB_CHECK_ARG(p.mode >= UMB_EPI_STORE && p.mode <= UMB_EPI_ROPE_QKV);
    if (p.mode == UMB_EPI_ROPE_QKV) {
        UMB_CHECK_ARG(epi->rope != nullptr);
        p.rope = *epi->rope;
        const umb_rope_epilogue_t& r = p.rope;
        UMB_CHECK_ARG(r.q_out && r.k_cache && r.v_cache && r.cos && r.sin && r.position_ids);
        UMB_CHECK_ARG((r.head_dim == 64 || r.head_dim == 128) && r.num_q_heads >= 1 && r.num_kv_heads >= 1);
        UMB_CHECK_ARG(N_total == (r.num_q_heads + 2 * r.num_kv_heads) * r.head_dim);
        const uintptr_t ptrs = (uintptr_t)r.q_out | (uintptr_t)r.k_cache | (uintptr_t)r.v_cache | (uintptr_t)r.cos | (uintptr_t)r.sin;
        p.rope_vec_ok = ((ptrs & 15) == 0 && (r.kv_tok_stride % 8) == 0 && (r.kv_head_stride % 8) == 0) ? 1 : 0;
    }
    if (p.mode == UMB_EPI_RESIDUAL) UMB_CHECK_ARG(p.residual != nullptr);
    p.vec_ok = (((uintptr_t)y & 15) == 0 && (ldy % 8) == 0 && (N_total % 8) == 0 &&
                (p.mode != UMB_EPI_RESIDUAL || ((((uintptr_t)p.residual) & 15) == 0 && (p.ld_residual % 8) == 0))) ? 1 : 0;
    if (p.mode == UMB_EPI_SILU_MUL) UMB_CHECK_ARG((N_total % kTileN) == 0 && p.bias == nullptr);

    const int mpad = mpad_of(M);
    const long long U = (long long)p.tiles * p.units_per_tile;
    if (U >= (1ll << 31)) return UMB_ERR_UNSUPPORTED;
    int grid = sm_count();
    if ((long long)grid > U) grid = (int)U;
    static int aligned_split = -1;                   // UMB_GEMM_ALIGNED_SPLIT=0: plain stream-K for small problems too (tuning)
    if (aligned_split < 0) {
        const char* e = getenv("UMB_GEMM_ALIGNED_SPLIT");
        aligned_split = (e && e[0] == '0') ? 0 : 1;
    }
    if (aligned_split && U < 16ll * grid && p.tiles <= grid) {
        // small problem (draft-model linears): latency-bound, so cut every tile into the same power-of-two
        // number of k-slices (one segment per CTA, fewer partials per tile to reduce) instead of balancing units
        int S = 1;
        while (p.tiles * (S * 2) <= grid && p.units_per_tile % (S * 2) == 0) S *= 2;
        grid = p.tiles * S;
    }
    const size_t need = umb_gemm_workspace_bytes(M, K, N_total);
    if (!workspace || workspace_bytes < need) return UMB_ERR_WORKSPACE;
    UMB_CHECK_ARG(((uintptr_t)workspace & 15) == 0);
    p.counters = reinterpret_cast<int*>(workspace);
    p.slots = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + (((size_t)p.tiles * 4 + 255) & ~(size_t)255));

    CUtensorMap tx, tw;
    int rc = make_tmap(&tx, x, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, (uint32_t)mpad);
    if (rc != UMB_OK) return rc;
    if (!awq) {
        UMB_CHECK_ARG((ldw % 8) == 0 && ldw >= K);
        rc = make_tmap(&tw, w, (uint64_t)N_total, (uint64_t)K, (uint64_t)ldw, (uint32_t)kTileN);
        if (rc != UMB_OK) return rc;
    } else if (UMB_X_3D) {
        rc = make_tmap_x3d(&tw, x, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, (uint32_t)mpad);   // AWQ: tmap_w = x as slabs
        if (rc != UMB_OK) return rc;
    } else {
        tw = tx;
    }
    if (sync_debug())
        fprintf(stderr, "umbrella_b200: gemm awq=%d M=%d K=%d N=%d mode=%d grid=%d mpad=%d\n", (int)awq, M, K, N_total, p.mode,
                grid, mpad);
#define UMB_DISPATCH(MP)                                                       \
    return awq ? launch_gemm<MP, true>(tx, tw, p, grid, st) : launch_gemm<MP, false>(tx, tw, p, grid, st)
    switch (mpad) {
        case 16: UMB_DISPATCH(16);
        case 32: UMB_DISPATCH(32);
        case 64: UMB_DISPATCH(64);
        default:
#if UMB_MAX_ROWS > 64
            UMB_DISPATCH(128);
#else
            if (awq) return UMB_ERR_UNSUPPORTED;      // never reached: gemm_common splits more than 64 rows
            return launch_gemm<128, false>(tx, tw, p, grid, st);
#endif
    }
#undef UMB_DISPATCH
}

}  // namespace umb

using namespace umb;

extern "C" size_t umb_awq_packed_bytes(int K, int N_total, int group) {
    if (group != kGroup || K <= 0 || N_total <= 0 || (K % kGroup) != 0) return 0;
    return (size_t)((N_total + kTileN - 1) / kTileN) * (K / kGroup) * kRecordStride;
}

extern "C" size_t umb_gemm_workspace_bytes(int M, int K, int N_total) {
    (void)K;
    if (M < 1 || N_total < 1) return 0;
    const size_t tiles = (size_t)(N_total + kTileN - 1) / kTileN;
    const size_t mpad = (size_t)mpad_of(M > 128 ? 128 : M);
    const size_t slots = tiles + 1024;   // tiles + an upper bound on the grid size
    return ((tiles * 4 + 255) & ~(size_t)255) + slots * mpad * kTileN * sizeof(float);
}

extern "C" int umb_awq_repack(const int32_t* qweight, const int32_t* qzeros, const void* scales, int K, int N, int group,
                              void* packed, int N_total, int row_block, int row_stride, int row_offset,
                              umb_stream_t stream) {
    UMB_CHECK_ARG(qweight && qzeros && scales && packed);
    if (group != kGroup) return UMB_ERR_UNSUPPORTED;
    UMB_CHECK_ARG(K > 0 && N > 0 && (K % kGroup) == 0 && (N % 8) == 0);
    UMB_CHECK_ARG(row_block > 0 && (row_block % 8) == 0 && (row_offset % 8) == 0 && (row_stride % 8) == 0);
    UMB_CHECK_ARG((N % row_block) == 0 || N < row_block);
    const int64_t last = (int64_t)((N - 1) / row_block) * row_stride + row_offset + (N - 1) % row_block;
    UMB_CHECK_ARG(last < (int64_t)((N_total + kTileN - 1) / kTileN) * kTileN);
    cudaStream_t st = (cudaStream_t)stream;
    const int cols = N / 8;
    awq_repack_weights_kernel<<<dim3((cols + 127) / 128, K / 8), 128, 0, st>>>(qweight, K, N, (unsigned char*)packed,
                                                                                row_block, row_stride, row_offset);
    awq_repack_meta_kernel<<<dim3((cols + 127) / 128, K / kGroup), 128, 0, st>>>(
        qzeros, (const __half*)scales, K, N, (unsigned char*)packed, row_block, row_stride, row_offset);
    count_launches(2);
    UMB_LAUNCH_CHECK();
    return UMB_OK;
}

extern "C" int umb_awq_dequantize_f16(const int32_t* qweight, const int32_t* qzeros, const void* scales, int K, int N,
                                      int group, void* out, umb_stream_t stream) {
    UMB_CHECK_ARG(qweight && qzeros && scales && out);
    if (group != kGroup) return UMB_ERR_UNSUPPORTED;
    UMB_CHECK_ARG(K > 0 && N > 0 && (K % kGroup) == 0 && (N % 8) == 0);
    awq_dequantize_kernel<<<dim3((N / 8 + 127) / 128, K), 128, 0, (cudaStream_t)stream>>>(
        qweight, qzeros, (const __half*)scales, K, N, (__half*)out);
    count_launches(1);
    UMB_LAUNCH_CHECK();
    return UMB_OK;
}

extern "C" int umb_awq_unpack_f16(const void* packed, int K, int N_total, int group, void* out, umb_stream_t stream) {
    UMB_CHECK_ARG(packed && out);
    if (group != kGroup) return UMB_ERR_UNSUPPORTED;
    UMB_CHECK_ARG(K > 0 && N_total > 0 && (K % kGroup) == 0);
    awq_unpack_kernel<<<dim3((K / 8 + 127) / 128, N_total), 128, 0, (cudaStream_t)stream>>>(
        (const unsigned char*)packed, K, N_total, (__half*)out);
    count_launches(1);
    UMB_LAUNCH_CHECK();
    return UMB_OK;
}

extern "C" int umb_awq_gemm_f16(const void* x, int64_t ldx, const void* packed, void* y, int64_t ldy, int M, int K,
                                int N_total, int group, const umb_epilogue_t* epilogue, void* workspace,
                                size_t workspace_bytes, umb_stream_t stream) {
    if (group != kGroup) return UMB_ERR_UNSUPPORTED;
    return gemm_common(true, x, ldx, packed, 0, y, ldy, M, K, N_total, epilogue, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}

extern "C" int umb_gemm_f16(const void* x, int64_t ldx, const void* w, int64_t ldw, void* y, int64_t ldy, int M, int K,
                            int N, const umb_epilogue_t* epilogue, void* workspace, size_t workspace_bytes,
                            umb_stream_t stream) {
    return gemm_common(false, x, ldx, w, ldw, y, ldy, M, K, N, epilogue, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}

extern "C" int umb_debug_gemm_trace(unsigned long long* host_out, int max_ctas) {
    unsigned long long* buf = trace_buffer();
    if (!buf || !host_out || max_ctas < 1) return UMB_ERR_ARG;
    if (max_ctas > 1024) max_ctas = 1024;
    UMB_CUDA(cudaDeviceSynchronize());
    UMB_CUDA(cudaMemcpy(host_out, buf, (size_t)max_ctas * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    UMB_CUDA(cudaMemset(buf, 0, 1024 * 16 * sizeof(unsigned long long)));
    g_trace_launches = 0;
    return UMB_OK;
}

extern "C" int umb_probe_umma(const void* a, const void* b, float* d, int mode, umb_stream_t stream) {
    UMB_CHECK_ARG(a && b && d && mode >= 0 && mode <= 3);
    if (mode >= 2)      // B operand MN-major: b is fp16 [64 k, 128 n], d fp32 [128, 128]
        probe_umma_mn_kernel<<<1, 128, 0, (cudaStream_t)stream>>>((const __half*)a, (const __half*)b, d, mode);
    else
        probe_umma_kernel<<<1, 128, 0, (cudaStream_t)stream>>>((const __half*)a, (const __half*)b, d, mode);
    count_launches(1);
    UMB_LAUNCH_CHECK();
    return UMB_OK;
}

extern "C" int umb_awq_gemm_f16_simt(const void* x, int64_t ldx, const void* packed, void* y, int64_t ldy, int M, int K,
                                     int N_total, umb_stream_t stream) {
    UMB_CHECK_ARG(x && packed && y && M >= 1 && K > 0 && (K % kGroup) == 0 && N_total > 0);
    awq_gemm_simt_kernel<<<(N_total + kTileN - 1) / kTileN, 128, 0, (cudaStream_t)stream>>>(
        (const unsigned char*)packed, (const __half*)x, ldx, (__half*)y, ldy, M, K, N_total);
    count_launches(1);
    UMB_LAUNCH_CHECK();
    return UMB_OK;
}
