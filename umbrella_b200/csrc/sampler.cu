// Fused target-side sampling of verify(): repetition penalty + top-k + softmax(T) + top-p + one sample per tree
// node, one CTA per node, straight from the fp32 logits the lm_head GEMM wrote.
//
// Replaces (reference paths):
//   apply_repetition_penalty            umbrella/speculation/speculation_utils.py:340-345
//   static engine, T >= 0.05            umbrella/speculation/static_speculation_engine.py:305-310
//       flashinfer.sampling.top_k_top_p_sampling_from_logits(logits / T, uniform_samples[3, N], topk, topp)
//       = top_k_mask_logits -> softmax -> top_p_sampling_from_probs (rejection rounds driven by the FIXED
//         uniform_samples the engine draws once, so the sample is a deterministic function of the logits)
//   dynamic engine, T >= 0.05           umbrella/speculation/dynamic_speculation_engine.py:276-281
//       apply_topk -> softmax(logits / T) -> flashinfer top_p_renorm_prob -> torch.multinomial
//       (multinomial's private Philox stream is replaced by one explicit uniform per node: inverse CDF in
//        ascending token id over the renormalised nucleus -- same distribution)
//   greedy with repetition penalty      static...:298-307, dynamic...:265-274  (arg-max of the penalised logits)
//
// The reference materialises [N, V] temporaries for the gather / scatter of the penalty, the top-k mask, the
// softmax, and (dynamic) the renormalised probabilities: ~10 launches and 6 x N x 513 KB of traffic.  Here the
// row is read 4-5 times from L2 by one CTA and nothing is written but the sampled id:
//   1. seen-token bitmap in shared memory (the penalty's gather / scatter is idempotent per token id)
//   2. exact k-th largest value by a 4-pass radix select over orderable keys (ties at the k-th value all stay,
//      as `logits < kth` / top_k_mask_logits keep them)
//   3. the survivors (top_k <= 1024 plus ties, <= 2048; index order) in shared memory; softmax, nucleus test and the draw run on them.
#include "common.cuh"
#include "internal.h"

namespace umb {

constexpr int kSampleThreads = 1024;
constexpr int kMaxTopK = 1024;
constexpr int kSurvivorCap = 2048;      // top_k + ties at the k-th value (fp16-valued logits tie often)

__device__ __forceinline__ uint32_t orderable_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct SampleParams {
    const float* logits;
    int64_t ld;
    int vocab;
    int mode;                 // UMB_SAMPLE_*
    float inv_t, top_p, penalty;
    int top_k;
    const float* uniforms;    // [rounds, rows]
    int rows, rounds;
    const int64_t* pen_ids;   // token ids seen so far (shared by all rows) or NULL
    int pen_count;            // host count, or the offset added to pen_count_dev[0]
    const int32_t* pen_count_dev;
    int64_t* out;
};

__global__ void __launch_bounds__(kSampleThreads, 1) sample_kernel(const SampleParams p) {
    extern __shared__ uint32_t s_seen[];                 // [ceil(vocab / 32)] when the penalty is on
    __shared__ int s_hist[256];
    __shared__ int s_idx[kSurvivorCap], s_tmp_idx[kSurvivorCap];
    __shared__ float s_val[kSurvivorCap], s_tmp_val[kSurvivorCap];
    __shared__ unsigned char s_keep[kSurvivorCap];
    __shared__ int s_count, s_digit, s_remaining;
    __shared__ unsigned long long s_best[kSampleThreads / 32];

    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* x_row = p.logits + (int64_t)row * p.ld;
    const bool pen = p.pen_ids != nullptr && p.penalty > 1.01f;
    if (pen) {
        const int words = (p.vocab + 31) >> 5;
        for (int i = tid; i < words; i += kSampleThreads) s_seen[i] = 0u;
        __syncthreads();
        const int n = p.pen_count_dev ? p.pen_count_dev[0] + p.pen_count : p.pen_count;
        for (int i = tid; i < n; i += kSampleThreads) {
            const long long t = p.pen_ids[i];
            if (t >= 0 && t < p.vocab) atomicOr(&s_seen[t >> 5], 1u << (t & 31));
        }
        __syncthreads();
    }
    const float inv_pen = 1.f / p.penalty;
    // the value the reference's top-k sees: penalised logit, divided by T first in the static engine
    auto value = [&](int j) -> float {
        float x = x_row[j];
        if (pen && ((s_seen[j >> 5] >> (j & 31)) & 1u)) x = x < 0.f ? x * p.penalty : x * inv_pen;
        return p.mode == UMB_SAMPLE_STATIC ? x * p.inv_t : x;
    };

    if (p.mode == UMB_SAMPLE_ARGMAX) {
        unsigned long long best = 0ull;
        for (int j = tid; j < p.vocab; j += kSampleThreads) {
            const unsigned long long key = ((unsigned long long)orderable_key(value(j)) << 32) | (unsigned int)(~j);
            best = key > best ? key : best;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if ((tid & 31) == 0) s_best[tid >> 5] = best;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kSampleThreads / 32; ++w) best = s_best[w] > best ? s_best[w] : best;
            p.out[row] = (long long)(~(unsigned int)(best & 0xffffffffull));
        }
        return;
    }

    // ---- exact k-th largest key: radix select, most significant digit first
    const int k = p.top_k < p.vocab ? p.top_k : p.vocab;
    uint32_t prefix = 0u, mask = 0u;
    if (tid == 0) s_remaining = k;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        for (int j = tid; j < p.vocab; j += kSampleThreads) {
            const uint32_t key = orderable_key(value(j));
            if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int cum = 0, d = 255, rem = s_remaining;
            for (; d > 0; --d) {
                if (cum + s_hist[d] >= rem) break;
                cum += s_hist[d];
            }
            s_digit = d;
            s_remaining = rem - cum;
        }
        __syncthreads();
        prefix |= (uint32_t)s_digit << shift;
        mask |= 255u << shift;
    }
    const uint32_t kth = prefix;

    // ---- survivors (key >= kth), then ordered by token id
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int j = tid; j < p.vocab; j += kSampleThreads) {
        const float v = value(j);
        if (orderable_key(v) >= kth) {
            const int at = atomicAdd(&s_count, 1);
            if (at < kSurvivorCap) {
                s_tmp_idx[at] = j;
                s_tmp_val[at] = p.mode == UMB_SAMPLE_STATIC ? v : v * p.inv_t;
            }
        }
    }
    __syncthreads();
    const int n = s_count < kSurvivorCap ? s_count : kSurvivorCap;   // > cap only for degenerate rows (mass ties)
    for (int t = tid; t < n; t += kSampleThreads) {
        const int mine = s_tmp_idx[t];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += s_tmp_idx[j] < mine;
        s_idx[rank] = mine;
        s_val[rank] = s_tmp_val[t];
    }
    __syncthreads();
    // ---- softmax over the survivors (everything else is exp(-inf) = 0 in the reference), fixed summation order
    __shared__ float s_max, s_sum;
    if (tid == 0) {
        float m = -INFINITY;
        for (int j = 0; j < n; ++j) m = fmaxf(m, s_val[j]);
        s_max = m;
    }
    __syncthreads();
    for (int t = tid; t < n; t += kSampleThreads) s_val[t] = expf(s_val[t] - s_max);
    __syncthreads();
    if (tid == 0) {
        float z = 0.f;
        for (int j = 0; j < n; ++j) z += s_val[j];
        s_sum = z;
    }
    __syncthreads();
    for (int t = tid; t < n; t += kSampleThreads) s_val[t] = s_val[t] / s_sum;
    __syncthreads();

    if (p.mode == UMB_SAMPLE_STATIC) {
        // top_p_sampling_from_probs: rejection rounds with the engine's fixed uniforms
        if (tid == 0) {
            float q = 1.f, pivot = 0.f;
            int sampled = n - 1;
            for (int r = 0; r < p.rounds; ++r) {
                const float u = p.uniforms[(int64_t)r * p.rows + row] * q;
                float run = 0.f;
                sampled = -1;
                int last_live = n - 1;
                for (int j = 0; j < n; ++j) {
                    if (s_val[j] > pivot) {
                        run += s_val[j];
                        last_live = j;
                        if (run > u) {
                            sampled = j;
                            break;
                        }
                    }
                }
                if (sampled < 0) sampled = last_live;
                pivot = fmaxf(pivot, s_val[sampled]);
                float gt = 0.f;
                for (int j = 0; j < n; ++j) gt += s_val[j] > pivot ? s_val[j] : 0.f;
                q = gt;
                if (q < p.top_p) break;
            }
            p.out[row] = s_idx[sampled];
        }
        return;
    }
    // ---- dynamic: top_p_renorm_prob's kept set {j : mass(probs > probs[j]) < top_p}, then the inverse CDF
    for (int t = tid; t < n; t += kSampleThreads) {
        const float mine = s_val[t];
        float gt = 0.f;
        for (int j = 0; j < n; ++j) gt += s_val[j] > mine ? s_val[j] : 0.f;
        s_keep[t] = gt < p.top_p;
    }
    __syncthreads();
    if (tid == 0) {
        float total = 0.f;
        int last = 0;
        for (int j = 0; j < n; ++j)
            if (s_keep[j]) {
                total += s_val[j];
                last = j;
            }
        const float u = p.uniforms[row] * total;
        float run = 0.f;
        int sampled = last;
        for (int j = 0; j < n; ++j)
            if (s_keep[j]) {
                run += s_val[j];
                if (run > u) {
                    sampled = j;
                    break;
                }
            }
        p.out[row] = s_idx[sampled];
    }
}


// ------------------------------------------------------------------------------------------------
// Row-wise top-k of fp32 logits, sorted (value descending, lower token id first among equals) -- replaces
// ``logits.topk(k)`` on the draft side: five mbtopk launches per call in torch, one here.
//   static engine   sampling_argmax_gather (speculation_utils.py:52-61): the next level's tokens are the per-node top
//                   ``max(branches)`` draft tokens gathered in growmap order  (static...:274-281)
//   dynamic engine  ``logits.topk(k=num_beams)`` feeding the beam scores (dynamic...:235)
// One CTA per row: exact k-th largest key by the sampler's radix select, survivors ranked in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kTopkMax = 64;
constexpr int kTopkCap = 512;      // top_k + ties at the k-th value

__global__ void __launch_bounds__(kSampleThreads, 1)
topk_kernel(const float* __restrict__ logits, int64_t ld, int vocab, int k, float* __restrict__ out_vals,
            int64_t* __restrict__ out_ids) {
    __shared__ int s_hist[256];
    __shared__ int s_idx[kTopkCap];
    __shared__ float s_val[kTopkCap];
    __shared__ int s_count, s_digit, s_remaining;
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* x_row = logits + (int64_t)row * ld;
    uint32_t prefix = 0u, mask = 0u;
    if (tid == 0) s_remaining = k;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        for (int j = tid; j < vocab; j += kSampleThreads) {
            const uint32_t key = orderable_key(x_row[j]);
            if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int cum = 0, d = 255, rem = s_remaining;
            for (; d > 0; --d) {
                if (cum + s_hist[d] >= rem) break;
                cum += s_hist[d];
            }
            s_digit = d;
            s_remaining = rem - cum;
        }
        __syncthreads();
        prefix |= (uint32_t)s_digit << shift;
        mask |= 255u << shift;
    }
    const uint32_t kth = prefix;
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int j = tid; j < vocab; j += kSampleThreads) {
        const float v = x_row[j];
        if (orderable_key(v) >= kth) {
            const int at = atomicAdd(&s_count, 1);
            if (at < kTopkCap) {
                s_idx[at] = j;
                s_val[at] = v;
            }
        }
    }
    __syncthreads();
    const int n = s_count < kTopkCap ? s_count : kTopkCap;
    for (int t = tid; t < n; t += kSampleThreads) {
        const float v = s_val[t];
        const int id = s_idx[t];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (s_val[j] > v) || (s_val[j] == v && s_idx[j] < id);
        if (rank < k) {
            out_vals[(int64_t)row * k + rank] = v;
            out_ids[(int64_t)row * k + rank] = id;
        }
    }
}

}  // namespace umb

extern "C" int umb_sample_f32(const float* logits, int64_t ld, int rows, int vocab, int mode, float temperature,
                              int top_k, float top_p, const float* uniforms, int rounds, const int64_t* penalty_ids,
                              int penalty_count, const int32_t* penalty_count_dev, float penalty, int64_t* out,
                              umb_stream_t stream) {
    UMB_CHECK_ARG(logits && out && rows >= 1 && vocab >= 1 && ld >= vocab);
    UMB_CHECK_ARG(mode == UMB_SAMPLE_STATIC || mode == UMB_SAMPLE_DYNAMIC || mode == UMB_SAMPLE_ARGMAX);
    UMB_CHECK_ARG(penalty_count >= 0 || penalty_count_dev);
    umb::SampleParams p{};
    p.logits = logits;
    p.ld = ld;
    p.vocab = vocab;
    p.mode = mode;
    p.penalty = penalty;
    p.pen_ids = penalty > 1.01f ? penalty_ids : nullptr;
    p.pen_count = penalty_count;
    p.pen_count_dev = penalty_count_dev;
    p.out = out;
    p.rows = rows;
    if (mode != UMB_SAMPLE_ARGMAX) {
        UMB_CHECK_ARG(uniforms && temperature > 0.f && top_k >= 1 && top_p > 0.f);
        UMB_CHECK_ARG(mode == UMB_SAMPLE_DYNAMIC || rounds >= 1);
        if (top_k > umb::kMaxTopK) return UMB_ERR_UNSUPPORTED;
        p.inv_t = 1.f / temperature;
        p.top_k = top_k;
        p.top_p = top_p;
        p.uniforms = uniforms;
        p.rounds = rounds;
    }
    const size_t smem = p.pen_ids ? (size_t)((vocab + 31) / 32) * 4 : 0;
    if (smem > 160 * 1024) return UMB_ERR_UNSUPPORTED;
    if (smem > 0)   // ~36 KB of static shared memory + the bitmap exceed the 48 KB default: opt in (per device, cheap)
        UMB_CUDA(cudaFuncSetAttribute(umb::sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    UMB_LAUNCH_PDL(umb::sample_kernel, dim3(rows), dim3(umb::kSampleThreads), smem, (cudaStream_t)stream, p);
    umb::count_launches(1);
    return UMB_OK;
}

extern "C" int umb_topk_f32(const float* logits, int64_t ld, int rows, int vocab, int k, float* out_vals, int64_t* out_ids,
                            umb_stream_t stream) {
    UMB_CHECK_ARG(logits && out_vals && out_ids && rows >= 1 && vocab >= 1 && ld >= vocab && k >= 1 && k <= vocab);
    if (k > umb::kTopkMax) return UMB_ERR_UNSUPPORTED;
    UMB_LAUNCH_PDL(umb::topk_kernel, dim3(rows), dim3(umb::kSampleThreads), 0, (cudaStream_t)stream, logits, ld, vocab, k,
                   out_vals, out_ids);
    umb::count_launches(1);
    return UMB_OK;
}
