// How long does it take ONE CTA per SM to get `n` bulk requests issued and landed when (a) one thread issues them in a
// loop, (b) n lanes of one warp issue one request each in the same instruction?  (ring fill of every GEMM launch)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../umbrella_b200/csrc -o tma_issue tma_issue.cu && ./tma_issue
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace umb;

__global__ void __launch_bounds__(64, 1)
fill_kernel(const unsigned char* __restrict__ src, size_t bytes_per_cta, int rec_bytes, int n, int parallel, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    unsigned char* ring = smem + 1024;
    const int smem_stride = (rec_bytes + 127) / 128 * 128;
    const unsigned char* base = src + (size_t)blockIdx.x * bytes_per_cta;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < n; ++i) mbar_init(full + i, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint64_t pol = l2_policy_evict_first();
        __syncwarp();
        const long long t0 = clock64();
        if (parallel) {
            if (lane < n) {
                mbar_arrive_expect_tx(full + lane, (uint32_t)rec_bytes);
                tma_bulk_g2s(ring + (size_t)lane * smem_stride, base + (size_t)lane * rec_bytes, rec_bytes, full + lane, pol);
            }
        } else if (lane == 0) {
            for (int i = 0; i < n; ++i) {
                mbar_arrive_expect_tx(full + i, (uint32_t)rec_bytes);
                tma_bulk_g2s(ring + (size_t)i * smem_stride, base + (size_t)i * rec_bytes, rec_bytes, full + i, pol);
            }
        }
        __syncwarp();
        const long long t1 = clock64();
        mbar_wait(full + 0, 0);
        const long long t2 = clock64();
        for (int i = 1; i < n; ++i) mbar_wait(full + i, 0);
        const long long t3 = clock64();
        if (lane == 0 && blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
            out[2] = t3 - t0;
        }
    }
}

int main() {
    const size_t total = (size_t)1 << 30;
    unsigned char* buf;
    cudaMalloc(&buf, total);
    cudaMemset(buf, 1, total);
    long long* out;
    cudaMalloc(&out, 64);
    cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    long long h[3];
    for (int grid : {148, 1})
        for (int rec : {8512, 16384})
            for (int n : {4, 8, 10})
                for (int par = 0; par < 2; ++par) {
                    const size_t per_cta = total / 148 / 2;
                    for (int rep = 0; rep < 2; ++rep) {
                        fill_kernel<<<grid, 64, 2048 + (size_t)n * ((rec + 127) / 128 * 128)>>>(buf + (size_t)rep * (total / 2), per_cta, rec, n, par, out);
                        cudaDeviceSynchronize();
                    }
                    cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
                    printf("grid %3d  %5d B x %2d  %-8s: issued after %6.2f us, first landed %6.2f us, all landed %6.2f us\n", grid, rec, n,
                           par ? "parallel" : "serial", h[0] / 1965.0, h[1] / 1965.0, h[2] / 1965.0);
                }
    return 0;
}
