// (iii) tree accept + KV gather -- integer kernels, bit-exact with the reference.
//
// Reference behaviour restated (paths relative to the reference tree):
//   accept block        umbrella/speculation/static_speculation_engine.py:312-327
//                       umbrella/speculation/dynamic_speculation_engine.py:283-302
//   EOS truncation      static_speculation_engine.py:329-334, speculation_utils.py:316-338
//   state advance       static_speculation_engine.py:340-349
//   KV gather           umbrella/attn/cache.py:41-49 ([L,S,Hkv,D]) and :136-144 ([L,Hkv,S,D])
#include "common.cuh"
#include "internal.h"

namespace umb {

constexpr int kAcceptThreads = 1024;
constexpr int kMaxTree = 1024;

// One CTA.  Thread i owns tree node i.
//   acc[i]   = (i == 0) | (sampled[parents[i]] == tokens[n + i])
//   valid[i] = acc[v] for every v on the path root..i  (== "row sum of accept*mask equals
//              node_in_path", the reference's test, because mask[i] is the ancestor-or-self set)
//   path     = ascending list of valid nodes (torch.nonzero order)
__global__ void __launch_bounds__(kAcceptThreads, 1)
tree_accept_kernel(const int64_t* __restrict__ sampled, int64_t* __restrict__ tokens,
                   int64_t* __restrict__ position_ids, const int32_t* __restrict__ parents,
                   const int32_t* __restrict__ depth, int N, int num_nodes_host, int capacity,
                   int32_t* __restrict__ state, const int64_t* __restrict__ eos, int n_eos,
                   int32_t* __restrict__ result) {
    __shared__ unsigned char s_acc[kMaxTree];
    __shared__ int s_path[kMaxTree];
    __shared__ long long s_pos[kMaxTree];
    __shared__ int s_warp_cnt[kAcceptThreads / 32];
    __shared__ int s_len, s_eos_pos;

    pdl_launch_dependents();
    pdl_wait();
    const int i = threadIdx.x;
    int n = state ? state[0] : num_nodes_host;
    // `capacity` = elements of tokens / position_ids: a prefix that leaves no room for the tree is a host-side
    // bookkeeping error (the engines refuse such a state); never index past the buffers because of it
    if (n + N > capacity) n = capacity - N > 0 ? capacity - N : 0;
    const int lane = i & 31, warp = i >> 5;

    long long spec = 0;
    int par = 0;
    if (i < N) {
        spec = tokens[n + i];
        par = parents[i];
        s_acc[i] = (i == 0) || (sampled[par] == spec);
    }
    if (i == 0) s_eos_pos = 0x7fffffff;
    __syncthreads();

    bool valid = false;
    if (i < N) {
        valid = s_acc[i];
        int v = par;
        // parents[child] < child, parents[0] == 0: the walk ends at the root
        while (valid && v != 0) {
            valid = s_acc[v];
            v = parents[v];
        }
    }
    // ordered compaction (ballot + per-warp prefix)
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) s_warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += s_warp_cnt[w];
    if (valid) s_path[base + __popc(ballot & ((1u << lane) - 1u))] = i;
    if (i == 0) {
        int tot = 0;
        for (int w = 0; w < kAcceptThreads / 32; ++w) tot += s_warp_cnt[w];
        s_len = tot;
    }
    __syncthreads();
    const int len = s_len;                       // >= 1 (the root is always valid)
    const long long bonus = sampled[s_path[len - 1]];

    // tokens[n + j] = spec[path[j]]; tokens[n + len] = bonus  (read all, then write)
    long long tok_j = 0;
    if (i < len) tok_j = tokens[n + s_path[i]];
    __syncthreads();
    if (i < len) tokens[n + i] = tok_j;
    if (i == len) tok_j = bonus;
    if (i == len && n + len < capacity) tokens[n + len] = bonus;
    // EOS scan over tokens[n : n + len + 1]
    if (i <= len) {
        bool hit = false;
        for (int e = 0; e < n_eos; ++e) hit |= (tok_j == eos[e]);
        if (hit) atomicMin(&s_eos_pos, i);
    }
    __syncthreads();
    int cont = 1, alen = len;
    if (s_eos_pos != 0x7fffffff) {
        cont = 0;
        alen = s_eos_pos;                        // accept_path[:eos_position]
    }
    const int new_n = n + alen;
    // position_ids[n : new_n] = position_ids[path_abs]; then the next tree's positions
    if (i < alen) s_pos[i] = position_ids[n + s_path[i]];
    __syncthreads();
    if (i < alen) position_ids[n + i] = s_pos[i];
    __syncthreads();
    if (i < N && new_n + i < capacity) position_ids[new_n + i] = (long long)new_n + depth[i];

    if (i == 0) {
        result[0] = alen;
        result[1] = cont;
        result[2] = new_n;
        result[3] = (int32_t)(bonus & 0xffffffffll);
        if (state) {
            state[0] = new_n;
            state[1] = new_n + N;
            state[2] = n;
        }
    }
    if (i < alen) result[4 + i] = n + s_path[i];
}

// grid (plane, 2): blockIdx.y = 0 -> K cache, 1 -> V cache.  Thread t owns 16-byte column t of a
// row.  Sources of a batch are loaded before its destinations are stored; across batches the
// ascending order is safe because dest(j) = offset + j <= path[j] < path[k] for k > j.
constexpr int kGatherBatch = 8;
__global__ void __launch_bounds__(256)
kv_gather_kernel(unsigned char* __restrict__ k_cache, unsigned char* __restrict__ v_cache, int64_t plane_stride,
                 int64_t row_stride, int row_bytes, const int32_t* __restrict__ result, int offset_host,
                 const int32_t* __restrict__ offset_dev, int zero_upto) {
    pdl_launch_dependents();
    pdl_wait();
    unsigned char* base = (blockIdx.y == 0 ? k_cache : v_cache) + (int64_t)blockIdx.x * plane_stride;
    const int len = result[0];
    const int offset = offset_dev ? offset_dev[0] : offset_host;
    const int32_t* path = result + 4;
    const int vecs = row_bytes >> 4;
    for (int c = threadIdx.x; c < vecs; c += blockDim.x) {
        for (int j0 = 0; j0 < len; j0 += kGatherBatch) {
            uint4 tmp[kGatherBatch];
#pragma unroll
            for (int b = 0; b < kGatherBatch; ++b)
                if (j0 + b < len)
                    tmp[b] = *reinterpret_cast<const uint4*>(base + (int64_t)path[j0 + b] * row_stride + (int64_t)c * 16);
#pragma unroll
            for (int b = 0; b < kGatherBatch; ++b)
                if (j0 + b < len)
                    *reinterpret_cast<uint4*>(base + (int64_t)(offset + j0 + b) * row_stride + (int64_t)c * 16) = tmp[b];
        }
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int r = offset + len; r < zero_upto; ++r)
            *reinterpret_cast<uint4*>(base + (int64_t)r * row_stride + (int64_t)c * 16) = z;
    }
}

}  // namespace umb

extern "C" int umb_tree_accept(const int64_t* sampled, int64_t* tokens, int64_t* position_ids, const int32_t* parents,
                               const int32_t* depth, int N, int num_nodes, int capacity, int32_t* state,
                               const int64_t* eos, int n_eos, int32_t* result, umb_stream_t stream) {
    UMB_CHECK_ARG(sampled && tokens && position_ids && parents && depth && result);
    UMB_CHECK_ARG(N >= 1 && num_nodes >= 0 && n_eos >= 0 && (n_eos == 0 || eos));
    UMB_CHECK_ARG(capacity >= N && (state || num_nodes + N <= capacity));
    if (N >= umb::kMaxTree) return UMB_ERR_UNSUPPORTED;
    UMB_LAUNCH_PDL(umb::tree_accept_kernel, dim3(1), dim3(umb::kAcceptThreads), 0, (cudaStream_t)stream, sampled, tokens,
                   position_ids, parents, depth, N, num_nodes, capacity, state, eos, n_eos, result);
    umb::count_launches(1);
    return UMB_OK;
}

extern "C" int umb_kv_gather(void* k_cache, void* v_cache, int num_planes, int64_t plane_stride_bytes,
                             int64_t row_stride_bytes, int row_bytes, const int32_t* result, int offset,
                             const int32_t* offset_dev, int zero_upto, umb_stream_t stream) {
    UMB_CHECK_ARG(k_cache && v_cache && result && num_planes >= 1);
    UMB_CHECK_ARG(row_bytes > 0 && (row_bytes % 16) == 0 && (row_stride_bytes % 16) == 0 && (plane_stride_bytes % 16) == 0);
    UMB_CHECK_ARG((((uintptr_t)k_cache | (uintptr_t)v_cache) & 15) == 0);
    const int threads = row_bytes / 16 >= 256 ? 256 : (row_bytes / 16 >= 128 ? 128 : 64);
    UMB_LAUNCH_PDL(umb::kv_gather_kernel, dim3(num_planes, 2), dim3(threads), 0, (cudaStream_t)stream,
                   (unsigned char*)k_cache, (unsigned char*)v_cache, plane_stride_bytes, row_stride_bytes, row_bytes,
                   result, offset, offset_dev, zero_upto);
    umb::count_launches(1);
    return UMB_OK;
}

namespace umb {
__global__ void store_tokens_kernel(int64_t* __restrict__ tokens, const int32_t* __restrict__ n_dev, int offset,
                                    const int64_t* __restrict__ src, int count, int capacity) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t at = (int64_t)n_dev[0] + offset + i;
    if (i < count && at < capacity) tokens[at] = src[i];
}
}  // namespace umb

namespace umb {
__global__ void store_tokens_gather_kernel(int64_t* __restrict__ tokens, const int32_t* __restrict__ n_dev, int num_nodes,
                                           int offset, const int64_t* __restrict__ src, const int64_t* __restrict__ gather,
                                           int count, int capacity) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t at = (int64_t)(n_dev ? n_dev[0] : num_nodes) + offset + i;
    if (i < count && at < capacity) tokens[at] = src[gather[i]];
}
}  // namespace umb

extern "C" int umb_store_tokens_gather(int64_t* tokens_base, const int32_t* n_dev, int num_nodes, int offset,
                                       const int64_t* src, const int64_t* gather, int count, int capacity,
                                       umb_stream_t stream) {
    UMB_CHECK_ARG(tokens_base && src && gather && count >= 1 && offset >= 0 && capacity >= 1 && num_nodes >= 0);
    UMB_LAUNCH_PDL(umb::store_tokens_gather_kernel, dim3((count + 127) / 128), dim3(128), 0, (cudaStream_t)stream,
                   tokens_base, n_dev, num_nodes, offset, src, gather, count, capacity);
    umb::count_launches(1);
    return UMB_OK;
}

extern "C" int umb_store_tokens(int64_t* tokens_base, const int32_t* n_dev, int offset, const int64_t* src, int count,
                                int capacity, umb_stream_t stream) {
    UMB_CHECK_ARG(tokens_base && n_dev && src && count >= 1 && offset >= 0 && capacity >= 1);
    UMB_LAUNCH_PDL(umb::store_tokens_kernel, dim3((count + 127) / 128), dim3(128), 0, (cudaStream_t)stream, tokens_base,
                   n_dev, offset, src, count, capacity);
    umb::count_launches(1);
    return UMB_OK;
}

// ------------------------------------------------------------------------------------------------
// Dynamic (SpecExec-style) tree growth: one level of DynamicSpeculationEngine.build_tree
// (reference umbrella/speculation/dynamic_speculation_engine.py:231-248) in one CTA.
//   cand[r, b]  = tree_score[level_first + r] + log(softmax(topk_logits[r, :])[b] + 1e-4)
//                 (softmax over the top-`beams` slice only, dynamic...:236-237)
//   the `width` best candidates, best first (torch.topk order; equal scores: lower flat index first)
//   become the next level: node id = level_first + rows + j
//   tree_score[id] = score, tokens[n + id] = topk_ids[flat], parents[id] = level_first + flat / beams,
//   tree_bits[id, :] = tree_bits[parent, :] | bit id   (the reference copies the parent's mask row and
//   sets the diagonal, dynamic...:247-248)
// ------------------------------------------------------------------------------------------------
namespace umb {
constexpr int kExpandThreads = 512;
constexpr int kMaxCand = 4096;

__global__ void __launch_bounds__(kExpandThreads, 1)
dyn_tree_expand_kernel(const float* __restrict__ topk_logits, const int64_t* __restrict__ topk_ids, int rows, int beams,
                       int width, int level_first, int num_nodes_host, int capacity,
                       const int32_t* __restrict__ n_dev, float* __restrict__ tree_score, int64_t* __restrict__ tokens, int32_t* __restrict__ parents,
                       uint32_t* __restrict__ tree_bits, int tree_words) {
    __shared__ float s_cand[kMaxCand];
    __shared__ float s_row_max[128], s_row_sum[128];
    pdl_launch_dependents();
    pdl_wait();
    const int n = n_dev ? n_dev[0] : num_nodes_host;
    const int total = rows * beams;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kExpandThreads / 32;
    // per-row max and sum(exp(x - max)) (torch softmax: fp32, max-subtracted)
    for (int r = warp; r < rows; r += nwarps) {
        float mx = -INFINITY;
        for (int b = lane; b < beams; b += 32) mx = fmaxf(mx, topk_logits[r * beams + b]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int b = lane; b < beams; b += 32) sum += expf(topk_logits[r * beams + b] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            s_row_max[r] = mx;
            s_row_sum[r] = sum;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < total; c += kExpandThreads) {
        const int r = c / beams;
        const float pr = expf(topk_logits[c] - s_row_max[r]) / s_row_sum[r];
        s_cand[c] = tree_score[level_first + r] + logf(pr + 1e-4f);
    }
    __syncthreads();
    // rank by counting: candidate c lands at output position #(better candidates)
    for (int c = threadIdx.x; c < total; c += kExpandThreads) {
        const float v = s_cand[c];
        int rank = 0;
        for (int j = 0; j < total; ++j) {
            const float o = s_cand[j];
            rank += (o > v) || (o == v && j < c);
        }
        if (rank < width) {
            const int id = level_first + rows + rank;
            const int par = level_first + c / beams;
            tree_score[id] = v;
            if (n + id < capacity) tokens[(int64_t)n + id] = topk_ids[c];
            parents[id] = par;
            for (int w = 0; w < tree_words; ++w) {
                uint32_t bits = tree_bits[(int64_t)par * tree_words + w];
                if (w == (id >> 5)) bits |= 1u << (id & 31);
                tree_bits[(int64_t)id * tree_words + w] = bits;
            }
        }
    }
}
}  // namespace umb

extern "C" int umb_dyn_tree_expand(const float* topk_logits, const int64_t* topk_ids, int rows, int beams, int width,
                                   int level_first, int num_nodes, int capacity, const int32_t* n_dev,
                                   float* tree_score, int64_t* tokens, int32_t* parents, uint32_t* tree_bits, int tree_words,
                                   umb_stream_t stream) {
    UMB_CHECK_ARG(topk_logits && topk_ids && tree_score && tokens && parents && tree_bits);
    UMB_CHECK_ARG(rows >= 1 && rows <= 128 && beams >= 1 && width >= 1 && level_first >= 0 && num_nodes >= 0);
    UMB_CHECK_ARG(capacity >= 1 && (n_dev || num_nodes + level_first + rows + width <= capacity));
    UMB_CHECK_ARG(width <= rows * beams && tree_words >= 1 && (level_first + rows + width) <= tree_words * 32);
    if (rows * beams > umb::kMaxCand) return UMB_ERR_UNSUPPORTED;
    UMB_LAUNCH_PDL(umb::dyn_tree_expand_kernel, dim3(1), dim3(umb::kExpandThreads), 0, (cudaStream_t)stream, topk_logits,
                   topk_ids, rows, beams, width, level_first, num_nodes, capacity, n_dev, tree_score, tokens, parents, tree_bits,
                   tree_words);
    umb::count_launches(1);
    return UMB_OK;
}
