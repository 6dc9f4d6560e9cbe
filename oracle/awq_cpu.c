/* Oracle (TEST / MEASUREMENT INFRASTRUCTURE ONLY -- see oracle/__init__.py): the AWQ INT4 linear of the verify
 * forward as a plain C restatement for host cores -- the `cpu_baseline` / `--impl reference` leg of bench.py and a
 * third implementation the numpy oracle is cross-checked against (tests/test_oracle_numeric.py).
 *
 * Restates the arithmetic behind AwqLinear.apply (reference umbrella/quantization/awq_utils.py:63-86 ->
 * awq_ext.gemm_forward_cuda, AutoAWQ GEMM layout, oracle/awq.py for the format):
 *     W[k, 8c + j] = fp16((nib(qweight[k, c], p_j) - nib(qzeros[k / g, c], p_j)) * scales[k / g, 8c + j]),
 *     p_j = (0, 4, 1, 5, 2, 6, 3, 7)[j];        y = x @ W   (fp32 accumulation, caller rounds to fp16)
 * without materialising W: the weights stay INT4 in memory (37.7 GB for the 70B model; an fp32 copy would be 274 GB),
 * each thread owns a block of output columns and streams k, and the (q - z) * s product goes through a 16-entry
 * table per (group, column) that is rebuilt every 128 k.
 *
 * Build (oracle/Makefile):  gcc -O3 -march=x86-64-v3 -pthread -shared -fPIC awq_cpu.c -o _build/libawq_cpu.so
 */
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <stdlib.h>

#define MPAD 16 /* token rows handled per pass (the verify tree has 13); more rows -> several passes */

/* IEEE binary16 <-> binary32, round to nearest even (no F16C dependency: the table build is off the hot loop) */
static float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            int e = -1;
            do {
                ++e;
                man <<= 1;
            } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
static uint16_t float_to_half(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - exp);
        uint32_t half = man >> shift, rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) ++half;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)exp << 10) | (man >> 13), rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) ++half;
    return (uint16_t)(sign | half);
}

static const int kNib[8] = {0, 4, 1, 5, 2, 6, 3, 7};

/* One block of CB packed words (128 output columns = one 64-byte line of every qweight row): the INT4 stream is
 * read with unit stride and exactly once; x is fp32, transposed [K][MPAD] (one broadcast-FMA vector per weight). */
#define CB 16
typedef struct {
    const float* xt;
    const int32_t* qweight;
    const int32_t* qzeros;
    const uint16_t* scales;
    int mrows, K, N, group;
    float* y;
    int first, last, step; /* blocks first, first + step, ... < last */
} job_t;

static void run_blocks(const job_t* jb) {
    const int cols8 = jb->N / 8, K = jb->K, N = jb->N, group = jb->group;
    float(*acc)[MPAD] = malloc(sizeof(float) * CB * 8 * MPAD);
    float(*lut)[16] = malloc(sizeof(float) * CB * 8 * 16);
    for (int b = jb->first; b < jb->last; b += jb->step) {
        const int c0 = b * CB, nc = cols8 - c0 < CB ? cols8 - c0 : CB;
        memset(acc, 0, sizeof(float) * CB * 8 * MPAD);
        for (int k = 0; k < K; ++k) {
            if (k % group == 0) {
                for (int c = 0; c < nc; ++c) {
                    const uint32_t zw = (uint32_t)jb->qzeros[(size_t)(k / group) * cols8 + c0 + c];
                    for (int j = 0; j < 8; ++j) {
                        const int z = (int)((zw >> (4 * kNib[j])) & 0xf);
                        const float s = half_to_float(jb->scales[(size_t)(k / group) * N + (c0 + c) * 8 + j]);
                        for (int q = 0; q < 16; ++q) lut[c * 8 + j][q] = half_to_float(float_to_half((float)(q - z) * s));
                    }
                }
            }
            const uint32_t* wrow = (const uint32_t*)jb->qweight + (size_t)k * cols8 + c0;
            const float* xr = jb->xt + (size_t)k * MPAD;
            for (int c = 0; c < nc; ++c) {
                const uint32_t w = wrow[c];
                for (int j = 0; j < 8; ++j) {
                    const float wv = lut[c * 8 + j][(w >> (4 * kNib[j])) & 0xf];
                    float* a = acc[c * 8 + j];
                    for (int m = 0; m < MPAD; ++m) a[m] += wv * xr[m];
                }
            }
        }
        for (int c = 0; c < nc * 8; ++c)
            for (int m = 0; m < jb->mrows; ++m) jb->y[(size_t)m * N + c0 * 8 + c] = acc[c][m];
    }
    free(acc);
    free(lut);
}
static void* worker(void* arg) {
    run_blocks((const job_t*)arg);
    return NULL;
}

/* y[mrows, N] (fp32) = x @ dequant(qweight[K, N/8], qzeros[K/g, N/8], scales[K/g, N] fp16 bits), xt = x as fp32
 * [K][16] (rows >= mrows zero), mrows <= 16, on `threads` host threads (blocks dealt round-robin). */
int oracle_awq_gemm_xt(const float* xt, const int32_t* qweight, const int32_t* qzeros, const uint16_t* scales, int mrows,
                       int K, int N, int group, float* y, int threads) {
    if (mrows < 1 || mrows > MPAD || group < 1 || K % group || N % 8 || threads < 1) return -1;
    const int blocks = (N / 8 + CB - 1) / CB;
    if (threads > blocks) threads = blocks;
    if (threads > 256) threads = 256;
    job_t jobs[256];
    pthread_t tid[256];
    for (int t = 0; t < threads; ++t) {
        job_t j = {xt, qweight, qzeros, scales, mrows, K, N, group, y, t, blocks, threads};
        jobs[t] = j;
    }
    for (int t = 1; t < threads; ++t)
        if (pthread_create(&tid[t], NULL, worker, &jobs[t]) != 0) return -2;
    run_blocks(&jobs[0]);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
    return 0;
}

/* convenience for the tests: x as fp16 bits [M, K], any M (passes of 16 rows) */
int oracle_awq_gemm(const uint16_t* x, const int32_t* qweight, const int32_t* qzeros, const uint16_t* scales, int M, int K,
                    int N, int group, float* y, int threads) {
    if (M < 1) return -1;
    float* xt = malloc(sizeof(float) * (size_t)K * MPAD);
    int rc = 0;
    for (int m0 = 0; m0 < M && rc == 0; m0 += MPAD) {
        const int mrows = M - m0 < MPAD ? M - m0 : MPAD;
        for (int k = 0; k < K; ++k)
            for (int m = 0; m < MPAD; ++m) xt[(size_t)k * MPAD + m] = m < mrows ? half_to_float(x[(size_t)(m0 + m) * K + k]) : 0.f;
        rc = oracle_awq_gemm_xt(xt, qweight, qzeros, scales, mrows, K, N, group, y + (size_t)m0 * N, threads);
    }
    free(xt);
    return rc;
}
