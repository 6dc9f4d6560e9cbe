// Library bookkeeping: version, error strings, launch counter, device properties.
#include <atomic>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace umb {
static std::atomic<unsigned long long> g_launches{0};
void count_launches(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("UMB_PDL");
        cached = (e && e[0] == '0') ? 0 : 1;
    }
    return cached != 0;
}
bool sync_debug() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("UMB_SYNC_DEBUG");
        cached = (e && e[0] == '1') ? 1 : 0;
    }
    return cached != 0;
}
cudaError_t sync_debug_check(const char* name, dim3 grid, dim3 block, cudaStream_t st, cudaError_t launch_rc) {
    cudaError_t rc = launch_rc;
    if (rc == cudaSuccess) rc = cudaStreamSynchronize(st);
    if (rc != cudaSuccess)
        fprintf(stderr, "umbrella_b200: kernel %s grid (%u,%u,%u) block %u failed: %s\n", name, grid.x, grid.y, grid.z,
                block.x, cudaGetErrorString(rc));
    return rc;
}
int arg_fail(const char* file, int line, const char* what) {
    static int dbg = -1;
    if (dbg < 0) {
        const char* e = getenv("UMB_DEBUG");
        dbg = (e && e[0] == '1') ? 1 : 0;
    }
    if (dbg) fprintf(stderr, "umbrella_b200: argument check failed at %s:%d: %s\n", file, line, what);
    return UMB_ERR_ARG;
}
int sm_count() {
    // cached per device ordinal: one process may drive several GPUs (benign race: every thread computes the same value)
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}
}  // namespace umb

extern "C" const char* umb_version(void) { return "umbrella_b200 0.1 (sm_100a)"; }

extern "C" const char* umb_error_string(int code) {
    switch (code) {
        case UMB_OK: return "ok";
        case UMB_ERR_ARG: return "invalid argument (pointer, size or alignment)";
        case UMB_ERR_UNSUPPORTED: return "shape or option not supported by the sm_100a kernels";
        case UMB_ERR_WORKSPACE: return "workspace too small";
        case UMB_ERR_DRIVER: return "CUDA driver entry point unavailable (cuTensorMapEncodeTiled)";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

extern "C" void umb_launch_count_add(unsigned long long n) { umb::g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" unsigned long long umb_launch_count(void) { return umb::g_launches.load(std::memory_order_relaxed); }
