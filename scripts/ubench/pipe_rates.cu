// Issue-rate microbenchmark (B200): warp-instructions per cycle per SM sub-partition for the opcodes of the
// AWQ dequant inner loop.  One CTA per SM, `warps` warps, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void k(unsigned* out, long long* cyc, unsigned seed) {
    unsigned r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;
    unsigned a = seed | 0x3c003c00u, b = 0x3c013c01u;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (OP == 1) asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(a));
            if (OP == 2) asm volatile("sub.rn.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(a));
            if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&a), "f"(*(float*)&b));
            if (OP == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (OP == 6) asm volatile("shr.u32 %0, %0, 8;" : "+r"(r[i]));
            if (OP == 7) {   // the dequant mix for one packed word: 1 shf, 4 lop3, 2 sub, 2 fma, 4 mul
                unsigned w = r[i], w8, t0_, t1_, t2_, t3_;
                asm volatile("shr.u32 %0, %1, 8;" : "=r"(w8) : "r"(w));
                asm volatile("lop3.b32 %0, %1, 0x000f000f, %2, 0xEA;" : "=r"(t0_) : "r"(w), "r"(b));
                asm volatile("lop3.b32 %0, %1, 0x00f000f0, %2, 0xEA;" : "=r"(t1_) : "r"(w), "r"(b));
                asm volatile("lop3.b32 %0, %1, 0x000f000f, %2, 0xEA;" : "=r"(t2_) : "r"(w8), "r"(b));
                asm volatile("lop3.b32 %0, %1, 0x00f000f0, %2, 0xEA;" : "=r"(t3_) : "r"(w8), "r"(b));
                asm volatile("sub.rn.f16x2 %0, %0, %1;" : "+r"(t0_) : "r"(a));
                asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(t1_) : "r"(a), "r"(b));
                asm volatile("sub.rn.f16x2 %0, %0, %1;" : "+r"(t2_) : "r"(a));
                asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(t3_) : "r"(a), "r"(b));
                asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(t0_) : "r"(a));
                asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(t1_) : "r"(a));
                asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(t2_) : "r"(a));
                asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(t3_) : "r"(a));
                r[i] = t0_ ^ t1_ ^ t2_ ^ t3_;
            }
        }
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int warps, double per_iter) {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<148, warps * 32>>>(out, cyc, 1);
    k<OP><<<148, warps * 32>>>(out, cyc, 2);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    double inst_per_smsp = (double)ITERS * 8 * per_iter * warps / 4.0;
    printf("%-28s warps/SM %2d  %.3f warp-instr/cycle/SMSP  (%.2f cycles per instr)\n", name, warps, inst_per_smsp / c, c / inst_per_smsp);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int warps : {4, 8, 16}) {
        run<0>("HFMA2 (fma.f16x2, 3 reg)", warps, 1);
        run<1>("HMUL2", warps, 1);
        run<2>("HADD2 (sub.f16x2)", warps, 1);
        run<3>("LOP3", warps, 1);
        run<4>("FFMA", warps, 1);
        run<5>("IMAD", warps, 1);
        run<6>("SHF", warps, 1);
        run<7>("dequant mix (13 instr/word)", warps, 13);
    }
    return 0;
}
