// Legacy warp-level tensor path on B200: issue rate of mma.sync.m16n8k16 (fp16 in, fp32 accumulate) per SM, alone and
// interleaved with the AWQ nibble -> fp16 conversion (1 shf + 4 lop3 + 4 hsub2/hfma2 per packed word), i.e. the inner
// loop of a register-level (Marlin-style) INT4 GEMV.  Answers: can 16-row mma.sync keep up with the HBM stream?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ void hmma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <uint32_t kMask>
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "n"(kMask), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t h2sub(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("sub.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t h2fma(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// mode 0: HMMA only (8 independent accumulators); mode 1: per packed word 9 conversion instructions + 2 HMMA
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(const uint32_t* __restrict__ in, float* out, int iters, long long* cyc) {
    float c[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
    uint32_t a[4] = {in[threadIdx.x], in[threadIdx.x + 1], in[threadIdx.x + 2], in[threadIdx.x + 3]};
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = in[threadIdx.x + 8 + j];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) {
                hmma(c[j], a, w[j], w[(j + 1) & 7]);
                hmma(c[j], a, w[(j + 2) & 7], w[(j + 3) & 7]);
            } else {
                const uint32_t x = w[j] + it;                    // a "fresh" packed word every iteration
                const uint32_t x8 = x >> 8;
                const uint32_t b0 = h2sub(lop3_and_or<0x000f000fu>(x, 0x64006400u), 0x64086408u);
                const uint32_t b1 = h2fma(lop3_and_or<0x00f000f0u>(x, 0x64006400u), 0x2C002C00u, 0xD480D480u);
                const uint32_t b2 = h2sub(lop3_and_or<0x000f000fu>(x8, 0x64006400u), 0x64086408u);
                const uint32_t b3 = h2fma(lop3_and_or<0x00f000f0u>(x8, 0x64006400u), 0x2C002C00u, 0xD480D480u);
                hmma(c[j], a, b0, b1);
                hmma(c[j], a, b2, b3);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    uint32_t* in;
    float* out;
    long long* cyc;
    cudaMalloc(&in, 8192);
    cudaMemset(in, 0x3c, 8192);
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(in, out, iters, cyc);
                else k<1><<<148, warps * 32>>>(in, out, iters, cyc);
                cudaDeviceSynchronize();
            }
            long long h;
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double hm = (double)iters * 16 * warps;        // HMMAs per SM
            printf("%s warps/SM %2d: %.2f cycles per HMMA per SM (%.0f dense fp16 MAC/clk/SM)%s\n",
                   mode == 0 ? "HMMA only          " : "convert(9) + 2 HMMA", warps, h / hm, 4096.0 * hm / h,
                   mode == 1 ? "  [per packed word of 8 weights: x2 cycles]" : "");
        }
    printf("budget at the HBM roofline (6.5 TB/s, 148 SMs, 1.965 GHz): 378 cycles per 128x128 unit per SM = 2.95 cycles per HMMA at 128 HMMA per unit\n");
    return 0;
}
