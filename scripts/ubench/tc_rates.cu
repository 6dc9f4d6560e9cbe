// tcgen05 rate microbenchmark (B200): what the AWQ GEMM redesign needs to know.
//   (1) cycles per tcgen05.mma kind::f16 128 x N x 16 (A from tensor memory, B from shared memory) when the 8
//       MMAs of a unit rotate through `nacc` accumulators (dependent-chain latency vs issue rate), issued by one
//       or by two warps;
//   (2) tcgen05.ld / tcgen05.st throughput with 4 warps (32x32b.x16 / .x64).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../umbrella_b200/csrc -o tc_rates tc_rates.cu && ./tc_rates
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"

using namespace umb;

#define UNITS 512

// one CTA per SM; warps 0 (and 1 if two_warps) issue UNITS units of 8 MMAs each
template <int N>
__global__ void __launch_bounds__(128, 1) mma_chain(long long* cyc, int nacc, int two_warps, int commit_every) {
    __shared__ __align__(1024) unsigned char s_b[256 * 128];   // [N rows][64 halves] SW128 (contents irrelevant)
    __shared__ uint64_t bar[2];
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256 * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(s_b)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&s_tmem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // A operand region: columns 256.. (garbage is fine: timing only)
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
    const uint64_t bd0 = umma_smem_desc_sw128(smem_u32(s_b));
    long long t0 = 0, t1 = 0;
    if (warp < (two_warps ? 2 : 1)) {
        const uint32_t d_base = tmem + warp * (nacc * N);
        __syncwarp();
        t0 = clock64();
        uint32_t ph = 0;
        for (int u = 0; u < UNITS; ++u) {
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_ts_f16(d_base + (kk % nacc) * N, tmem + 256 + 64 * (u & 3) + kk * 8, bd0 + (kk & 3) * 2, idesc,
                                kk < nacc ? 0u : 1u);
                if (commit_every && (u % commit_every) == commit_every - 1) umma_commit(bar + warp);
            }
            __syncwarp();
            if (commit_every && (u % commit_every) == commit_every - 1) {
                mbar_wait(bar + warp, ph);
                ph ^= 1;
            }
        }
        if (!commit_every) {
            if (elect_one()) umma_commit(bar + warp);
            __syncwarp();
            mbar_wait(bar + warp, 0);
        }
        t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

#define TMEM_LD_X64(taddr, r)                                                                                     \
    asm volatile(                                                                                                  \
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "                                                                  \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31," \
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];" \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),     \
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),          \
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),         \
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),         \
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),         \
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),         \
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),         \
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])                      \
        : "r"(taddr))

// 4 (or nw) warps each issue ITER loads (mode 0: x16, 1: x64) or stores (mode 2: x16) of their lane quarter
__global__ void __launch_bounds__(512, 1) tmem_rw(long long* cyc, unsigned* sink, int mode, int iters) {
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        tmem_alloc(&s_tmem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = threadIdx.x + i;
    unsigned acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = (it * 64) & 255;
        if (mode == 0) {
            UMB_TMEM_LD_X16(tmem + col, r);
            tc_wait_ld();
            acc += r[0] + r[15];
        } else if (mode == 1) {
            TMEM_LD_X64(tmem + col, r);
            tc_wait_ld();
            acc += r[0] + r[63];
        } else if (mode == 3) {   // 4 x16 loads in flight before one wait
            UMB_TMEM_LD_X16(tmem + col, r);
            UMB_TMEM_LD_X16(tmem + col + 16, (r + 16));
            UMB_TMEM_LD_X16(tmem + col + 32, (r + 32));
            UMB_TMEM_LD_X16(tmem + col + 48, (r + 48));
            tc_wait_ld();
            acc += r[0] + r[63];
        } else {
            UMB_TMEM_ST_X16(tmem + col, r);
            UMB_TMEM_ST_X16(tmem + col + 16, (r + 16));
            UMB_TMEM_ST_X16(tmem + col + 32, (r + 32));
            UMB_TMEM_ST_X16(tmem + col + 48, (r + 48));
            tc_wait_st();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(s_tmem, 512);
    }
}


// (3) contention: 2 MMA warps issue units (A from TMEM columns 256.., B from smem, `nacc` accumulators each) while
//     `st_warps` other warps keep storing 32 KB "A operands" into tensor memory (tcgen05.st x16 x4 + wait::st per warp
//     iteration = 8 KB per warp) and / or `ld_warps` warps keep loading accumulators.
__global__ void __launch_bounds__(640, 1) mma_vs_st(long long* cyc, int nacc, int st_warps, int ld_warps, int units, int st_gap) {
    __shared__ __align__(1024) unsigned char s_b[64 * 128];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_b)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        fence_barrier_init();
        s_stop = 0;
    }
    if (warp == 0) {
        tmem_alloc(&s_tmem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    constexpr uint32_t idesc = umma_idesc_f16(128, 16);
    const uint64_t bd0 = umma_smem_desc_sw128(smem_u32(s_b));
    if (warp < 2) {
        const uint32_t d_base = tmem + warp * (nacc * 16);
        long long t0 = clock64();
        for (int u = 0; u < units; ++u) {
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_ts_f16(d_base + (kk % nacc) * 16, tmem + 128 + 64 * (u % 6) + kk * 8, bd0 + (kk & 3) * 2, idesc,
                                kk < nacc ? 0u : 1u);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(bar + warp);
        __syncwarp();
        mbar_wait(bar + warp, 0);
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
        if (lane == 0) atomicAdd((int*)&s_stop, 1);
    } else if (warp < 2 + st_warps) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128;
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = threadIdx.x + i;
        long long n = 0, t0 = clock64();
        while (s_stop < 2) {
            const uint32_t col = base + 64 * (uint32_t)(n % 6);
            UMB_TMEM_ST_X16(col, r);
            UMB_TMEM_ST_X16(col + 16, r);
            UMB_TMEM_ST_X16(col + 32, r);
            UMB_TMEM_ST_X16(col + 48, r);
            tc_wait_st();
            ++n;
            for (int g = 0; g < st_gap; ++g) __nanosleep(20);
        }
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0 && warp == 2) {
            cyc[2] = n;
            cyc[3] = t1 - t0;
        }
    } else if (warp < 2 + st_warps + ld_warps) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t r[16];
        unsigned acc = 0;
        while (s_stop < 2) {
            UMB_TMEM_LD_X16(base, r);
            tc_wait_ld();
            acc += r[0];
        }
        if (acc == 0x12345678u) cyc[7] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int main() {
    long long* cyc;
    unsigned* sink;
    cudaMalloc(&cyc, 64);
    cudaMalloc(&sink, 148 * 512 * 4);
    long long h[2];
    printf("tcgen05.mma kind::f16 128xNx16, A from TMEM: cycles per unit of 8 MMAs (one SM of 148 running)\n");
    for (int two = 0; two < 2; ++two)
        for (int ce = 0; ce < 2; ++ce)
            for (int nacc = 1; nacc <= 8; nacc *= 2) {
                cudaMemset(cyc, 0, 16);
                mma_chain<16><<<148, 128>>>(cyc, nacc, two, ce);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
                printf("  N=16 warps=%d commit/wait every unit=%d nacc=%d: %7.1f cyc/unit (per warp) %s\n", two + 1, ce, nacc,
                       (double)h[0] / UNITS, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    for (int nacc = 1; nacc <= 4; nacc *= 2) {
        cudaMemset(cyc, 0, 16);
        mma_chain<32><<<148, 128>>>(cyc, nacc, 1, 0);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        printf("  N=32 warps=2 nacc=%d: %7.1f cyc/unit (per warp)\n", nacc, (double)h[0] / UNITS);
    }
    for (int nacc = 1; nacc <= 2; nacc *= 2) {
        cudaMemset(cyc, 0, 16);
        mma_chain<64><<<148, 128>>>(cyc, nacc, 1, 0);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        printf("  N=64 warps=2 nacc=%d: %7.1f cyc/unit (per warp)\n", nacc, (double)h[0] / UNITS);
    }
    printf("tcgen05.ld / st, cycles per iteration (each warp moves its 32-lane quarter)\n");
    const char* names[4] = {"ld x16 (2 KB/warp)", "ld x64 (8 KB/warp)", "st 4 x x16 (8 KB/warp)", "ld 4 x x16 one wait (8 KB/warp)"};
    for (int mode = 0; mode < 4; ++mode)
        for (int threads = 128; threads <= 512; threads *= 2) {
            cudaMemset(cyc, 0, 16);
            tmem_rw<<<148, threads>>>(cyc, sink, mode, 2048);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
            printf("  %-32s warps=%2d: %7.1f cyc/iter %s\n", names[mode], threads / 32, (double)h[0] / 2048,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    printf("MMA (2 warps, N=16, 8 MMAs per unit) against concurrent tcgen05.st / tcgen05.ld warps\n");
    long long h4[8];
    for (int nacc : {4, 1})
        for (int stw : {0, 4, 8, 12, 16})
            for (int gap : {0, 4}) {
                if (stw == 0 && gap) continue;
                cudaMemset(cyc, 0, 64);
                mma_vs_st<<<148, 640>>>(cyc, nacc, stw, 0, 2048, gap);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h4, cyc, 64, cudaMemcpyDeviceToHost);
                printf("  nacc=%d st_warps=%2d gap=%d: %7.1f cyc/unit per MMA warp; one st warp: %lld x 8 KB in %lld cyc (%.1f cyc / 8 KB) %s\n",
                       nacc, stw, gap, (double)h4[0] / 2048, h4[2], h4[3], h4[2] ? (double)h4[3] / h4[2] : 0.0,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    for (int ldw : {4, 8}) {
        cudaMemset(cyc, 0, 64);
        mma_vs_st<<<148, 640>>>(cyc, 4, 12, ldw, 2048, 0);
        cudaDeviceSynchronize();
        cudaMemcpy(h4, cyc, 64, cudaMemcpyDeviceToHost);
        printf("  nacc=4 st_warps=12 ld_warps=%d: %7.1f cyc/unit per MMA warp\n", ldw, (double)h4[0] / 2048);
    }
    return 0;
}
