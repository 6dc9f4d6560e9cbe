// Per-SM / whole-chip throughput of the copy engine feeding the AWQ GEMM's weight ring (B200): cp.async.bulk of one
// record per request into a shared-memory ring, nothing consuming but the mbarrier wait -- what the stream could
// deliver if dequant and MMA were free.  Variants: record size / stride (alignment of the global address), ring depth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../umbrella_b200/csrc -o tma_rates tma_rates.cu && ./tma_rates
#include <cstdio>
#include <cuda_runtime.h>

#include "common.cuh"
using namespace umb;

__global__ void __launch_bounds__(64, 1)
stream_kernel(const unsigned char* __restrict__ src, size_t bytes_per_cta, int rec_bytes, int rec_stride, int stages,
              int smem_stride, long long* cyc, int policy_kind, int split) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    unsigned char* ring = smem + 1024;
    const int n = (int)(bytes_per_cta / rec_stride);
    const unsigned char* base = src + (size_t)blockIdx.x * bytes_per_cta;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) mbar_init(full + i, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t pol = policy_kind ? l2_policy_evict_first() : l2_policy_evict_last();
        long long t0 = clock64();
        int ph = 0, st = 0;
        for (int i = 0; i < n + stages; ++i) {
            if (i >= stages) {
                mbar_wait(full + st, ph);          // the copy that used this stage has landed
            }
            if (i < n) {
                mbar_arrive_expect_tx(full + st, (uint32_t)rec_bytes);
                const int part = rec_bytes / split;
                for (int s = 0; s < split; ++s)
                    tma_bulk_g2s(ring + (size_t)st * smem_stride + s * part, base + (size_t)i * rec_stride + s * part, part,
                                 full + st, pol);
            }
            if (++st == stages) {
                st = 0;
                if (i >= stages) ph ^= 1;
            }
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) cyc[0] = t1 - t0;
    }
}

int main() {
    const size_t total = (size_t)1 << 30;
    unsigned char* buf;
    cudaMalloc(&buf, total + 65536);
    cudaMemset(buf, 1, total + 65536);
    long long* cyc;
    cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    struct V { int rec, stride, stages, split; size_t off; const char* name; };
    V vs[] = {
        {8512, 8512, 10, 1, 0, "8512 B records, stride 8512 (product: 64-B misaligned every other record)"},
        {8512, 8576, 10, 1, 0, "8512 B records, stride 8576 (128-B aligned)"},
        {8576, 8576, 10, 1, 0, "8576 B records, stride 8576"},
        {8192, 8192, 10, 1, 0, "8192 B records, stride 8192"},
        {8512, 8512, 10, 1, 16, "8512 B records, stride 8512, base + 16"},
        {8512, 8512, 20, 1, 0, "8512 B, 20 stages"},
        {8512, 8576, 20, 1, 0, "8512 B stride 8576, 20 stages"},
        {17024, 17024, 10, 1, 0, "17024 B records (2 per request)"},
        {17152, 17152, 10, 1, 0, "17152 B records aligned"},
        {8512, 8576, 10, 2, 0, "8512 B stride 8576, issued as 2 x 4256"},
        {8192, 8192, 10, 4, 0, "8192 B as 4 x 2048"},
        {4096, 4096, 20, 1, 0, "4096 B records, 20 stages"},
        {32768, 32768, 6, 1, 0, "32768 B records, 6 stages"},
    };
    for (int grid : {148, 1}) {
        printf("grid = %d CTAs\n", grid);
        for (auto& v : vs) {
            const size_t per_cta = ((total / 148) / v.stride) * v.stride;
            const int smem_stride = (v.rec + 127) / 128 * 128;
            const size_t smem = 2048 + (size_t)v.stages * smem_stride;
            for (int pol = 1; pol < 2; ++pol) {
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                stream_kernel<<<grid, 64, smem>>>(buf + v.off, per_cta, v.rec, v.stride, v.stages, smem_stride, cyc, pol, v.split);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                stream_kernel<<<grid, 64, smem>>>(buf + v.off, per_cta, v.rec, v.stride, v.stages, smem_stride, cyc, pol, v.split);
                cudaEventRecord(e1);
                cudaError_t err = cudaDeviceSynchronize();
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = (double)(per_cta / v.stride) * v.rec * grid;
                printf("  %-78s %8.1f GB/s  (%.1f GB/s per SM) %s\n", v.name, bytes / ms / 1e6, bytes / ms / 1e6 / grid,
                       err == cudaSuccess ? "" : cudaGetErrorString(err));
            }
        }
    }
    return 0;
}
