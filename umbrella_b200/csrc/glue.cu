// Glue kernels of the verify forward: RMSNorm, RoPE + KV append, SiLU*up, residual add, embedding
// gather, arg-max.  All HBM/L2-bound elementwise or row-reduction work: 16-byte vector accesses,
// fp32 reductions with warp shuffles, fp16 rounding points placed where the reference has fp16
// tensors.
//
// Reference behaviour restated (paths relative to the reference tree):
//   layer_norm -> flashinfer.rmsnorm     umbrella/models/model_utils.py:54-64
//   apply_rotary_pos_emb / rotate_half   umbrella/models/model_utils.py:17-52
//   KV append                            umbrella/attn/cache.py:53-65
//   silu(gate) * up, residual adds       umbrella/models/llama.py:293-301
//   F.embedding                          umbrella/models/llama.py:312
//   argmax                               umbrella/speculation/static_speculation_engine.py:307
#include "common.cuh"
#include "internal.h"

namespace umb {

union H8 {
    uint4 u;
    __half2 h2[4];
    __half h[8];
};

// ------------------------------------------------------------------------------------------------
// RMSNorm: y = fp16(float(x) * rsqrt(mean(x^2) + eps) * float(w)); one CTA per row.
// ------------------------------------------------------------------------------------------------
constexpr int kNormThreads = 256;
__global__ void __launch_bounds__(kNormThreads)
rmsnorm_kernel(const __half* __restrict__ x, const __half* __restrict__ w, __half* __restrict__ y, int hidden,
               float eps) {
    __shared__ float s_part[kNormThreads / 32];
    pdl_launch_dependents();
    pdl_wait();
    const __half* xr = x + (int64_t)blockIdx.x * hidden;
    __half* yr = y + (int64_t)blockIdx.x * hidden;
    const int vecs = hidden >> 3;
    float ss = 0.f;
    for (int v = threadIdx.x; v < vecs; v += kNormThreads) {
        H8 a;
        a.u = reinterpret_cast<const uint4*>(xr)[v];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float f = __half2float(a.h[j]);
            ss = fmaf(f, f, ss);
        }
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < kNormThreads / 32; ++i) tot += s_part[i];
    const float r = rsqrtf(tot / (float)hidden + eps);
    for (int v = threadIdx.x; v < vecs; v += kNormThreads) {
        H8 a, b, o;
        a.u = reinterpret_cast<const uint4*>(xr)[v];
        b.u = reinterpret_cast<const uint4*>(w)[v];
#pragma unroll
        for (int j = 0; j < 8; ++j) o.h[j] = __float2half_rn(__half2float(a.h[j]) * r * __half2float(b.h[j]));
        reinterpret_cast<uint4*>(yr)[v] = o.u;
    }
}

// Register-resident form for hidden <= 8 * 256 * VPT: the row and the weight vector are fetched once, all loads in
// flight together (the weight loads even before griddepcontrol.wait -- they never depend on the predecessor), so
// the kernel is one global round trip + the block reduction instead of three dependent passes.  Same arithmetic.
template <int VPT>
__global__ void __launch_bounds__(kNormThreads)
rmsnorm_reg_kernel(const __half* __restrict__ x, const __half* __restrict__ w, __half* __restrict__ y, int hidden,
                   float eps) {
    __shared__ float s_part[kNormThreads / 32];
    pdl_launch_dependents();
    const int vecs = hidden >> 3;
    H8 a[VPT], b[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * kNormThreads;
        b[i].u = v < vecs ? reinterpret_cast<const uint4*>(w)[v] : make_uint4(0, 0, 0, 0);
    }
    pdl_wait();
    const __half* xr = x + (int64_t)blockIdx.x * hidden;
    __half* yr = y + (int64_t)blockIdx.x * hidden;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * kNormThreads;
        a[i].u = v < vecs ? reinterpret_cast<const uint4*>(xr)[v] : make_uint4(0, 0, 0, 0);
    }
    // same accumulation order as the streaming kernel: vector by vector, element by element
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float f = __half2float(a[i].h[j]);
            ss = fmaf(f, f, ss);
        }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < kNormThreads / 32; ++i) tot += s_part[i];
    const float r = rsqrtf(tot / (float)hidden + eps);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * kNormThreads;
        if (v < vecs) {
            H8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j) o.h[j] = __float2half_rn(__half2float(a[i].h[j]) * r * __half2float(b[i].h[j]));
            reinterpret_cast<uint4*>(yr)[v] = o.u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------------
__global__ void silu_mul_kernel(const __half* __restrict__ gate, const __half* __restrict__ up, __half* __restrict__ y,
                                int64_t nvec) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvec) return;
    H8 g, u, o;
    g.u = reinterpret_cast<const uint4*>(gate)[v];
    u.u = reinterpret_cast<const uint4*>(up)[v];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float f = __half2float(g.h[j]);
        const __half s = __float2half_rn(f / (1.f + expf(-f)));   // F.silu on an fp16 tensor: fp32 maths, one rounding
        o.h[j] = __hmul_rn(s, u.h[j]);                            // fp16 * fp16 tensor product
    }
    reinterpret_cast<uint4*>(y)[v] = o.u;
}

__global__ void add_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ y,
                           int64_t nvec) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvec) return;
    H8 p, q, o;
    p.u = reinterpret_cast<const uint4*>(a)[v];
    q.u = reinterpret_cast<const uint4*>(b)[v];
#pragma unroll
    for (int j = 0; j < 4; ++j) o.h2[j] = __hadd2(p.h2[j], q.h2[j]);
    reinterpret_cast<uint4*>(y)[v] = o.u;
}

// ids may live at a device-resident offset (CUDA-graph replay with a moving window)
__global__ void embedding_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ ids_offset_dev,
                                 int ids_offset_add, const __half* __restrict__ table, __half* __restrict__ y,
                                 int hidden) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t off = ids_offset_dev ? (int64_t)ids_offset_dev[0] + ids_offset_add : 0;
    const int64_t id = ids[off + blockIdx.x];
    const uint4* src = reinterpret_cast<const uint4*>(table + id * hidden);
    uint4* dst = reinterpret_cast<uint4*>(y + (int64_t)blockIdx.x * hidden);
    for (int v = threadIdx.x; v < (hidden >> 3); v += blockDim.x) dst[v] = src[v];
}

// ------------------------------------------------------------------------------------------------
// RoPE (HF half-split, fp16 arithmetic: two products and one sum, each rounded) + KV append.
// grid (qo, Hq + 2 Hkv); block D/2 threads: thread d owns elements d and d + D/2 of one head.
// ------------------------------------------------------------------------------------------------
__global__ void rope_append_kernel(const __half* __restrict__ qkv, int64_t ld_qkv, __half* __restrict__ q_out,
                                   __half* __restrict__ k_cache, __half* __restrict__ v_cache,
                                   const __half* __restrict__ cos_t, const __half* __restrict__ sin_t,
                                   const int64_t* __restrict__ position_ids, const int32_t* __restrict__ pos_offset_dev,
                                   int Hq, int Hkv, int D, int64_t kv_tok_stride, int64_t kv_head_stride,
                                   int kv_offset_host, const int32_t* __restrict__ kv_offset_dev) {
    // with the device words: kv row / position index = dev[0] + host offset
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x, h = blockIdx.y, d = threadIdx.x, half_d = D >> 1;
    const __half* src = qkv + (int64_t)t * ld_qkv + (int64_t)h * D;
    const int kv_offset = kv_offset_dev ? kv_offset_dev[0] + kv_offset_host : kv_offset_host;
    if (h >= Hq + Hkv) {  // V: plain copy
        __half* dst = v_cache + (int64_t)(kv_offset + t) * kv_tok_stride + (int64_t)(h - Hq - Hkv) * kv_head_stride;
        dst[d] = src[d];
        dst[d + half_d] = src[d + half_d];
        return;
    }
    const int64_t pos = position_ids[(pos_offset_dev ? pos_offset_dev[0] + kv_offset_host : 0) + t];
    const __half c0 = cos_t[pos * D + d], c1 = cos_t[pos * D + d + half_d];
    const __half s0 = sin_t[pos * D + d], s1 = sin_t[pos * D + d + half_d];
    const __half x0 = src[d], x1 = src[d + half_d];
    // (x * cos) + (rotate_half(x) * sin), rotate_half = cat(-x[D/2:], x[:D/2])
    // _rn forms: each product and the sum round separately (no contraction into an fma), as the
    // reference's three fp16 tensor ops do
    const __half o0 = __hadd_rn(__hmul_rn(x0, c0), __hmul_rn(__hneg(x1), s0));
    const __half o1 = __hadd_rn(__hmul_rn(x1, c1), __hmul_rn(x0, s1));
    __half* dst = h < Hq ? q_out + ((int64_t)t * Hq + h) * D
                         : k_cache + (int64_t)(kv_offset + t) * kv_tok_stride + (int64_t)(h - Hq) * kv_head_stride;
    dst[d] = o0;
    dst[d + half_d] = o1;
}

// ------------------------------------------------------------------------------------------------
// arg-max over fp32 logits: key = (orderable(value) << 32) | ~index, reduced with a 64-bit
// atomicMax -> largest value, smallest index among equals (first occurrence).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int orderable_f32(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
constexpr int kArgmaxThreads = 256;
__global__ void __launch_bounds__(kArgmaxThreads)
argmax_keys_kernel(const float* __restrict__ logits, int64_t ld, int vocab, unsigned long long* __restrict__ keys) {
    __shared__ unsigned long long s_best[kArgmaxThreads / 32];
    const float* row = logits + (int64_t)blockIdx.y * ld;
    const int chunk = (vocab + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * chunk, hi = min(vocab, lo + chunk);
    unsigned long long best = 0ull;
    for (int i = lo + threadIdx.x; i < hi; i += kArgmaxThreads) {
        const unsigned long long k = ((unsigned long long)orderable_f32(row[i]) << 32) | (unsigned int)(~i);
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < kArgmaxThreads / 32; ++i) best = s_best[i] > best ? s_best[i] : best;
        if (best) atomicMax(keys + blockIdx.y, best);
    }
}
// decode keys -> int64 ids and leave the key buffer zeroed for the next use
__global__ void argmax_decode_kernel(unsigned long long* keys, int rows, int64_t* out, int reset_keys) {
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const unsigned long long k = keys[r];
    if (reset_keys) keys[r] = 0ull;
    out[r] = (int64_t)(unsigned int)(~(unsigned int)(k & 0xffffffffull));
}

}  // namespace umb

using namespace umb;

extern "C" int umb_rmsnorm_f16(const void* x, const void* weight, void* y, int rows, int hidden, float eps,
                               umb_stream_t stream) {
    UMB_CHECK_ARG(x && weight && y && rows >= 1 && hidden >= 8 && (hidden % 8) == 0);
    UMB_CHECK_ARG((((uintptr_t)x | (uintptr_t)weight | (uintptr_t)y) & 15) == 0);
    const int vecs = hidden >> 3;
    cudaStream_t st = (cudaStream_t)stream;
    const __half* xh = (const __half*)x;
    const __half* wh = (const __half*)weight;
    __half* yh = (__half*)y;
    if (vecs <= kNormThreads) {
        UMB_LAUNCH_PDL(rmsnorm_reg_kernel<1>, dim3(rows), dim3(kNormThreads), 0, st, xh, wh, yh, hidden, eps);
    } else if (vecs <= 2 * kNormThreads) {
        UMB_LAUNCH_PDL(rmsnorm_reg_kernel<2>, dim3(rows), dim3(kNormThreads), 0, st, xh, wh, yh, hidden, eps);
    } else if (vecs <= 4 * kNormThreads) {
        UMB_LAUNCH_PDL(rmsnorm_reg_kernel<4>, dim3(rows), dim3(kNormThreads), 0, st, xh, wh, yh, hidden, eps);
    } else {
        UMB_LAUNCH_PDL(rmsnorm_kernel, dim3(rows), dim3(kNormThreads), 0, st, xh, wh, yh, hidden, eps);
    }
    count_launches(1);
    return UMB_OK;
}

extern "C" int umb_silu_mul_f16(const void* gate, const void* up, void* y, int64_t n, umb_stream_t stream) {
    UMB_CHECK_ARG(gate && up && y && n >= 8 && (n % 8) == 0);
    UMB_CHECK_ARG((((uintptr_t)gate | (uintptr_t)up | (uintptr_t)y) & 15) == 0);
    const int64_t nvec = n / 8;
    silu_mul_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)gate, (const __half*)up, (__half*)y, nvec);
    count_launches(1);
    UMB_LAUNCH_CHECK();
    return UMB_OK;
}

extern "C" int umb_add_f16(const void* a, const void* b, void* y, int64_t n, umb_stream_t stream) {
    UMB_CHECK_ARG(a && b && y && n >= 8 && (n % 8) == 0);
    UMB_CHECK_ARG((((uintptr_t)a | (uintptr_t)b | (uintptr_t)y) & 15) == 0);
    const int64_t nvec = n / 8;
    add_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)a, (const __half*)b,
                                                                                 (__half*)y, nvec);
    count_launches(1);
    UMB_LAUNCH_CHECK();
    return UMB_OK;
}

extern "C" int umb_embedding_f16(const int64_t* ids, const void* table, void* y, int rows, int hidden,
                                 umb_stream_t stream) {
    UMB_CHECK_ARG(ids && table && y && rows >= 1 && hidden >= 8 && (hidden % 8) == 0);
    UMB_CHECK_ARG((((uintptr_t)table | (uintptr_t)y) & 15) == 0);
    UMB_LAUNCH_PDL(embedding_kernel, dim3(rows), dim3(256), 0, (cudaStream_t)stream, ids, (const int32_t*)nullptr, 0,
                   (const __half*)table, (__half*)y, hidden);
    count_launches(1);
    return UMB_OK;
}

namespace umb {
int embedding_dyn(const int64_t* ids_base, const int32_t* n_dev, int offset, const void* table, void* y, int rows,
                  int hidden, cudaStream_t st) {
    UMB_LAUNCH_PDL(embedding_kernel, dim3(rows), dim3(256), 0, st, ids_base, n_dev, offset, (const __half*)table,
                   (__half*)y, hidden);
    count_launches(1);
    return UMB_OK;
}
// position_ids is the BASE pointer; rows are read at n_dev[0] + offset + t and appended at the same cache rows
int rope_append_dyn(const void* qkv, int64_t ld_qkv, void* q_out, void* k_cache, void* v_cache, const void* cos,
                    const void* sin, const int64_t* position_base, const int32_t* n_dev, int offset, int qo, int Hq,
                    int Hkv, int D, int64_t kv_tok_stride, int64_t kv_head_stride, cudaStream_t st) {
    UMB_LAUNCH_PDL(rope_append_kernel, dim3(qo, Hq + 2 * Hkv), dim3(D / 2), 0, st, (const __half*)qkv, ld_qkv,
                   (__half*)q_out, (__half*)k_cache, (__half*)v_cache, (const __half*)cos, (const __half*)sin,
                   position_base, n_dev, Hq, Hkv, D, kv_tok_stride, kv_head_stride, offset, n_dev);
    count_launches(1);
    return UMB_OK;
}
}  // namespace umb

extern "C" int umb_rope_append_f16(const void* qkv, int64_t ld_qkv, void* q_out, void* k_cache, void* v_cache,
                                   const void* cos, const void* sin, const int64_t* position_ids, int qo,
                                   int num_q_heads, int num_kv_heads, int head_dim, int64_t kv_tok_stride,
                                   int64_t kv_head_stride, int kv_offset, const int32_t* kv_offset_dev,
                                   umb_stream_t stream) {
    UMB_CHECK_ARG(qkv && q_out && k_cache && v_cache && cos && sin && position_ids);
    UMB_CHECK_ARG(qo >= 1 && num_q_heads >= 1 && num_kv_heads >= 1 && kv_offset >= 0);
    if (head_dim < 2 || head_dim > 512 || (head_dim & 1)) return UMB_ERR_UNSUPPORTED;
    UMB_LAUNCH_PDL(rope_append_kernel, dim3(qo, num_q_heads + 2 * num_kv_heads), dim3(head_dim / 2), 0,
                   (cudaStream_t)stream, (const __half*)qkv, ld_qkv, (__half*)q_out, (__half*)k_cache, (__half*)v_cache,
                   (const __half*)cos, (const __half*)sin, position_ids, (const int32_t*)nullptr, num_q_heads,
                   num_kv_heads, head_dim, kv_tok_stride, kv_head_stride, kv_offset, kv_offset_dev);
    count_launches(1);
    return UMB_OK;
}

extern "C" int umb_argmax_f32(const float* logits, int64_t ld, int rows, int vocab, int64_t* out, umb_stream_t stream) {
    UMB_CHECK_ARG(logits && out && rows >= 1 && vocab >= 1 && ld >= vocab);
    // out is int64[rows], the same size as the key array: keys are built and decoded in place
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(out);
    UMB_CUDA(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * rows, (cudaStream_t)stream));
    const int chunks = vocab >= 16384 ? 16 : 1;
    argmax_keys_kernel<<<dim3(chunks, rows), kArgmaxThreads, 0, (cudaStream_t)stream>>>(logits, ld, vocab, keys);
    argmax_decode_kernel<<<(rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(keys, rows, out, 0);
    count_launches(2);
    UMB_LAUNCH_CHECK();
    return UMB_OK;
}

extern "C" int umb_argmax_keys_to_ids(const unsigned long long* keys, int rows, int64_t* out, umb_stream_t stream) {
    UMB_CHECK_ARG(keys && out && rows >= 1);
    UMB_CHECK_ARG((const void*)keys != (const void*)out);
    UMB_LAUNCH_PDL(argmax_decode_kernel, dim3((rows + 127) / 128), dim3(128), 0, (cudaStream_t)stream,
                   const_cast<unsigned long long*>(keys), rows, out, 1);
    count_launches(1);
    return UMB_OK;
}
