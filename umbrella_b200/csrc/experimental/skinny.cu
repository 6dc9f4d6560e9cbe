// EXPERIMENT RECORD -- NOT PART OF THE BUILD (umbrella_b200/build.py does not compile this file).
// Measured in round 2 and not adopted: shorter per launch under ncu than the tensor-memory kernel on the draft model's
// shapes, but the draft forward as the iteration graph replays it got slower (855 vs 802 us): this kernel stages weights in
// registers, fills the register file and cannot co-reside with its stream predecessor, while gemm.cu's kernel requests
// its whole weight share into shared memory before griddepcontrol.wait (profiles/r2_gemm_experiments.txt, section I).
// To rebuild the experiment: add "skinny.cu" to SOURCES, declare skinny_gemm_f16 in internal.h and call it from
// gemm_common for !awq && M <= 8 (git history of round 2 has the wiring, including the fused-norm forward driver).
//
// fp16-weight linear for very few token rows (M <= 8): the draft model's per-level forwards.
//
// Replaces torch F.linear on the draft side of build_tree (reference umbrella/models/llama.py:87-110 as driven by
// static_speculation_engine.py:257-281): 1-8 rows against 2048 x 2048 ... 8192 x 2048 matrices.  These launches are
// latency-bound, not bandwidth-bound -- the persistent tcgen05 kernel of gemm.cu spends 14-21 us on each (barrier and
// tensor-memory set-up, ring fill, stream-K fix-up) for 8-67 MB of weights (1.3-10 us at the HBM roofline).  This
// kernel has no set-up at all:
//   * a work item is 16 output features x all of K; the 8 warps of a CTA split K, every lane has its 16-byte weight
//     loads of a whole chunk in flight before it touches them (the first chunk even before griddepcontrol.wait: the
//     weights never depend on the stream predecessor);
//   * the products go through the warp-level tensor path: mma.sync.m16n8k16 with the WEIGHTS as the 16-row operand and
//     the (<= 8) token rows as the 8-column operand.  A lane's 16 bytes of a weight row are 8 consecutive k; the MMA's
//     k index is permuted accordingly (a lane's halves (0,1)(2,3) feed one MMA, (4,5)(6,7) the next), and the x operand
//     is read from shared memory with the same permutation, so no ldmatrix / transposition is needed;
//   * the per-warp partial sums (fp32) are added in warp order through shared memory (deterministic), and the epilogue
//     (store / residual add / SiLU(gate) * up, with the fp16 rounding points of gemm.cu's epilogue) runs in the MMA's
//     accumulator layout.
// Bound: HBM (2 * K * N bytes per launch); what it buys is the removal of the fixed microseconds per launch.
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace umb {

constexpr int kSkinnyWarps = 8;
constexpr int kSkinnyThreads = kSkinnyWarps * 32;
constexpr int kSkinnyChunk = 4;                 // 32-k steps per lane with their loads in flight together

struct SkinnyParams {
    const __half* x;
    int64_t ldx;
    const __half* w;
    int64_t ldw;
    __half* y;
    int64_t ldy;
    int M, K, N;                 // N = rows of W
    int mode;                    // UMB_EPI_STORE / RESIDUAL / SILU_MUL
    const __half* bias;
    const __half* residual;
    int64_t ld_residual;
    int items;                   // N / 16
    umb_rope_epilogue_t rope;    // UMB_EPI_ROPE_QKV operands (by value)
    const __half* norm_weight;   // != NULL: x is RMS-normalised on the way into shared memory (fused rmsnorm_reg_kernel)
    float norm_eps;
};

__device__ __forceinline__ void hmma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                           uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float silu_f32(float f) { return f / (1.f + expf(-f)); }

// W rows of item `it`: lane group g (0..7) owns rows r0 (MMA rows 0..7) and r1 (MMA rows 8..15)
__device__ __forceinline__ void skinny_rows(const SkinnyParams& p, int it, int g, int& r0, int& r1) {
    if (p.mode == UMB_EPI_SILU_MUL) {
        // gate | up interleaved 64 | 64 per 128-row tile (models/llama.py): item = 8 gate rows + their 8 up rows
        const int tile = it >> 3, j0 = (it & 7) * 8;
        r0 = tile * 128 + j0 + g;
        r1 = r0 + 64;
    } else if (p.mode == UMB_EPI_ROPE_QKV && it < (p.rope.num_q_heads + p.rope.num_kv_heads) * (p.rope.head_dim >> 4)) {
        // q / k heads: a lane owns feature d < D/2 of a head and its rotation partner d + D/2
        const int per_head = p.rope.head_dim >> 4;
        const int head = it / per_head, j0 = (it - head * per_head) * 8;
        r0 = head * p.rope.head_dim + j0 + g;
        r1 = r0 + (p.rope.head_dim >> 1);
    } else {
        r0 = it * 16 + g;
        r1 = r0 + 8;
    }
}

__global__ void __launch_bounds__(kSkinnyThreads, 4)
skinny_f16_kernel(const SkinnyParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int row_bytes = p.K * 2 + 64;                    // +64: the 8 token rows of a quarter-warp hit distinct banks
    unsigned char* s_x = smem_raw;                         // [M + 1][K + 32] fp16: the token rows and one zero row
    float4* s_red = reinterpret_cast<float4*>(smem_raw + (p.M + 1) * row_bytes);   // [warps][32] partial accumulators
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int k_per_warp = p.K / kSkinnyWarps;             // multiple of 32
    const int steps = k_per_warp >> 5;
    const int k_lane = warp * k_per_warp + t * 8;          // first k of this lane

    pdl_launch_dependents();

    int it = blockIdx.x;
    uint4 wa[kSkinnyChunk], wb[kSkinnyChunk];
    auto load_chunk = [&](int item, int s0) {
        int r0, r1;
        skinny_rows(p, item, g, r0, r1);
        const __half* pa = p.w + (int64_t)r0 * p.ldw + k_lane + s0 * 32;
        const __half* pb = p.w + (int64_t)r1 * p.ldw + k_lane + s0 * 32;
#pragma unroll
        for (int s = 0; s < kSkinnyChunk; ++s)
            if (s0 + s < steps) {
                wa[s] = ldg_nc_v4(pa + s * 32);
                wb[s] = ldg_nc_v4(pb + s * 32);
            }
    };
    if (it < p.items) load_chunk(it, 0);                   // weights first: they never depend on the predecessor

    pdl_wait();                                            // x (and the residual) are the predecessor's output
    {
        const int vecs_per_row = p.K >> 3;
        for (int c = threadIdx.x; c < vecs_per_row; c += kSkinnyThreads)      // the zero row (token columns >= M)
            *reinterpret_cast<uint4*>(s_x + p.M * row_bytes + c * 16) = make_uint4(0, 0, 0, 0);
        if (p.norm_weight == nullptr) {
            for (int v = threadIdx.x; v < p.M * vecs_per_row; v += kSkinnyThreads) {
                const int r = v / vecs_per_row, c = v - r * vecs_per_row;
                *reinterpret_cast<uint4*>(s_x + r * row_bytes + c * 16) =
                    *reinterpret_cast<const uint4*>(p.x + (int64_t)r * p.ldx + c * 8);
            }
        } else {
            // Fused RMSNorm (K == 2048: one 16-byte vector per thread and row).  Same arithmetic, same association order
            // as glue.cu's rmsnorm_reg_kernel<1> -- element by element inside a thread, xor-shuffle tree inside a warp,
            // the 8 warp sums in order -- so the normalised fp16 rows are bit-identical to the stand-alone kernel's.
            union H8 { uint4 u; __half h[8]; };
            float* s_part = reinterpret_cast<float*>(s_red);   // [8 rows][8 warps], free until the first reduction
            H8 wv;
            wv.u = reinterpret_cast<const uint4*>(p.norm_weight)[threadIdx.x];
#pragma unroll 4
            for (int r = 0; r < p.M; ++r) {
                H8 xv;
                xv.u = *reinterpret_cast<const uint4*>(p.x + (int64_t)r * p.ldx + threadIdx.x * 8);
                *reinterpret_cast<uint4*>(s_x + r * row_bytes + threadIdx.x * 16) = xv.u;
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float f = __half2float(xv.h[j]);
                    ss = fmaf(f, f, ss);
                }
                ss = warp_sum(ss);
                if (lane == 0) s_part[r * kSkinnyWarps + warp] = ss;
            }
            __syncthreads();
#pragma unroll 4
            for (int r = 0; r < p.M; ++r) {
                float tot = 0.f;
#pragma unroll
                for (int i = 0; i < kSkinnyWarps; ++i) tot += s_part[r * kSkinnyWarps + i];
                const float rs = rsqrtf(tot / (float)p.K + p.norm_eps);
                H8 xv, o;
                xv.u = *reinterpret_cast<const uint4*>(s_x + r * row_bytes + threadIdx.x * 16);   // this thread's own store
#pragma unroll
                for (int j = 0; j < 8; ++j) o.h[j] = __float2half_rn(__half2float(xv.h[j]) * rs * __half2float(wv.h[j]));
                *reinterpret_cast<uint4*>(s_x + r * row_bytes + threadIdx.x * 16) = o.u;
            }
        }
    }
    __syncthreads();

    const unsigned char* xrow = s_x + (g < p.M ? g : p.M) * row_bytes + k_lane * 2;      // token g (or zeros), this lane's k
    for (; it < p.items; it += gridDim.x) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s0 = 0; s0 < steps; s0 += kSkinnyChunk) {
            if (s0 > 0) load_chunk(it, s0);
#pragma unroll
            for (int s = 0; s < kSkinnyChunk; ++s)
                if (s0 + s < steps) {
                    const uint4 xv = *reinterpret_cast<const uint4*>(xrow + (s0 + s) * 64);
                    hmma_16816(c, wa[s].x, wb[s].x, wa[s].y, wb[s].y, xv.x, xv.y);
                    hmma_16816(c, wa[s].z, wb[s].z, wa[s].w, wb[s].w, xv.z, xv.w);
                }
        }
        const int next = it + gridDim.x;
        if (next < p.items) load_chunk(next, 0);           // the next item's first chunk flies during the reduction
        s_red[warp * 32 + lane] = make_float4(c[0], c[1], c[2], c[3]);
        __syncthreads();
        if (warp == 0) {
            float4 a = s_red[lane];
#pragma unroll
            for (int w = 1; w < kSkinnyWarps; ++w) {       // fixed order: deterministic
                const float4 b = s_red[w * 32 + lane];
                a.x += b.x;
                a.y += b.y;
                a.z += b.z;
                a.w += b.w;
            }
            // accumulator layout: (a.x, a.y) = MMA row g, token columns 2t, 2t+1; (a.z, a.w) = MMA row g + 8
            int r0, r1;
            skinny_rows(p, it, g, r0, r1);
            const float v0[2] = {a.x, a.y}, v1[2] = {a.z, a.w};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int m = 2 * t + j;
                if (m >= p.M) continue;
                if (p.mode == UMB_EPI_ROPE_QKV) {
                    // RoPE + KV append (glue.cu rope_append_kernel, same fp16 arithmetic: two products and one sum, each
                    // rounded): the projection is rounded to fp16 first, as the reference's q / k / v tensors are
                    const umb_rope_epilogue_t& r = p.rope;
                    const int D = r.head_dim, half_d = D >> 1;
                    const int base = r.n_dev ? r.n_dev[0] : 0;
                    const __half x0 = __float2half_rn(v0[j]), x1 = __float2half_rn(v1[j]);
                    const int head = r0 / D, d0 = r0 - head * D, d1 = r1 - head * D;
                    __half o0 = x0, o1 = x1;
                    __half* dst;
                    if (head < r.num_q_heads + r.num_kv_heads) {
                        const int64_t pos = r.position_ids[(r.n_dev ? base + r.row_offset : 0) + m];
                        const __half* ct = reinterpret_cast<const __half*>(r.cos) + pos * D;
                        const __half* sn = reinterpret_cast<const __half*>(r.sin) + pos * D;
                        o0 = __hadd_rn(__hmul_rn(x0, ct[d0]), __hmul_rn(__hneg(x1), sn[d0]));
                        o1 = __hadd_rn(__hmul_rn(x1, ct[d0 + half_d]), __hmul_rn(x0, sn[d0 + half_d]));
                        dst = head < r.num_q_heads
                                  ? reinterpret_cast<__half*>(r.q_out) + ((int64_t)m * r.num_q_heads + head) * D
                                  : reinterpret_cast<__half*>(r.k_cache) + (int64_t)(base + r.kv_offset + m) * r.kv_tok_stride +
                                        (int64_t)(head - r.num_q_heads) * r.kv_head_stride;
                    } else {
                        dst = reinterpret_cast<__half*>(r.v_cache) + (int64_t)(base + r.kv_offset + m) * r.kv_tok_stride +
                              (int64_t)(head - r.num_q_heads - r.num_kv_heads) * r.kv_head_stride;
                    }
                    dst[d0] = o0;
                    dst[d1] = o1;
                } else if (p.mode == UMB_EPI_SILU_MUL) {
                    // gate and up are rounded to fp16 (the reference holds them as fp16 tensors), SiLU in fp32 rounded
                    // once, the product rounded once -- gemm.cu's UMB_EPI_SILU_MUL epilogue
                    const __half gate = __float2half_rn(v0[j]), up = __float2half_rn(v1[j]);
                    const __half sl = __float2half_rn(silu_f32(__half2float(gate)));
                    p.y[(int64_t)m * p.ldy + it * 8 + g] = __hmul(sl, up);
                } else {
                    __half h0 = __float2half_rn(v0[j]), h1 = __float2half_rn(v1[j]);
                    if (p.bias) {
                        h0 = __hadd(h0, p.bias[r0]);
                        h1 = __hadd(h1, p.bias[r1]);
                    }
                    if (p.mode == UMB_EPI_RESIDUAL) {
                        h0 = __hadd(p.residual[(int64_t)m * p.ld_residual + r0], h0);
                        h1 = __hadd(p.residual[(int64_t)m * p.ld_residual + r1], h1);
                    }
                    p.y[(int64_t)m * p.ldy + r0] = h0;
                    p.y[(int64_t)m * p.ldy + r1] = h1;
                }
            }
        }
        __syncthreads();                                   // s_red is free again
    }
}

static bool skinny_enabled() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("UMB_SKINNY");
        cached = (e && e[0] == '0') ? 0 : 1;
    }
    return cached != 0;
}

// UMB_OK if the launch was taken, UMB_ERR_UNSUPPORTED if the shape / mode belongs to the tensor-memory kernel.
int skinny_gemm_f16(const void* x, int64_t ldx, const void* w, int64_t ldw, void* y, int64_t ldy, int M, int K, int N,
                    const umb_epilogue_t* epi, cudaStream_t st, const void* norm_weight, float norm_eps) {
    if (!skinny_enabled()) return UMB_ERR_UNSUPPORTED;
    if (norm_weight && (K != 8 * kSkinnyThreads || ((uintptr_t)norm_weight & 15) != 0)) return UMB_ERR_UNSUPPORTED;
    const int mode = epi ? epi->mode : UMB_EPI_STORE;
    if (M > 8 || (mode != UMB_EPI_STORE && mode != UMB_EPI_RESIDUAL && mode != UMB_EPI_SILU_MUL && mode != UMB_EPI_ROPE_QKV))
        return UMB_ERR_UNSUPPORTED;
    if (K % (32 * kSkinnyWarps) != 0 || K > 8192 || N % 16 != 0 || (mode == UMB_EPI_SILU_MUL && N % 128 != 0))
        return UMB_ERR_UNSUPPORTED;
    if ((((uintptr_t)x | (uintptr_t)w) & 15) != 0 || (ldx % 8) != 0 || (ldw % 8) != 0) return UMB_ERR_UNSUPPORTED;
    SkinnyParams p{};
    p.x = (const __half*)x;
    p.ldx = ldx;
    p.w = (const __half*)w;
    p.ldw = ldw;
    p.y = (__half*)y;
    p.ldy = ldy;
    p.M = M;
    p.K = K;
    p.N = N;
    p.mode = mode;
    p.bias = epi ? (const __half*)epi->bias : nullptr;
    p.residual = epi ? (const __half*)epi->residual : nullptr;
    p.ld_residual = epi ? epi->ld_residual : 0;
    p.items = N / 16;
    p.norm_weight = (const __half*)norm_weight;
    p.norm_eps = norm_eps;
    if (mode == UMB_EPI_SILU_MUL && p.bias) return UMB_ERR_UNSUPPORTED;
    if (mode == UMB_EPI_ROPE_QKV) {
        if (!epi->rope || p.bias) return UMB_ERR_UNSUPPORTED;
        p.rope = *epi->rope;
        const umb_rope_epilogue_t& r = p.rope;
        if (!(r.q_out && r.k_cache && r.v_cache && r.cos && r.sin && r.position_ids)) return UMB_ERR_ARG;
        if ((r.head_dim != 64 && r.head_dim != 128) || N != (r.num_q_heads + 2 * r.num_kv_heads) * r.head_dim)
            return UMB_ERR_UNSUPPORTED;
    }
    if (mode == UMB_EPI_RESIDUAL && !p.residual) return UMB_ERR_ARG;
    const int smem = (M + 1) * (K * 2 + 64) + kSkinnyWarps * 32 * (int)sizeof(float4);
    const int per_sm = smem <= 56 * 1024 ? 4 : (smem <= 112 * 1024 ? 2 : 1);
    int grid = sm_count() * per_sm;
    if (grid > p.items) grid = p.items;
    UMB_CUDA(cudaFuncSetAttribute(skinny_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    UMB_LAUNCH_PDL(skinny_f16_kernel, dim3(grid), dim3(kSkinnyThreads), (size_t)smem, st, p);
    count_launches(1);
    return UMB_OK;
}

}  // namespace umb
