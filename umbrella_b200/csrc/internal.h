// Library-internal host helpers shared by the translation units of libumbrella_b200.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/umbrella_b200.h"

#include <utility>

// Token rows one GEMM launch takes (and the forward drivers' chunk size).  Round 1 held this at 64 because the
// 128-column AWQ instantiation died about once in 150-300 launches; the cause (a parity wait that could run two phases
// ahead of its barrier in the 4-deep ring) was found and fixed in round 2 -- profiles/r2_m128_root_cause.txt.
// -DUMB_MAX_ROWS=64 restores the split into two launches.
#ifndef UMB_MAX_ROWS
#define UMB_MAX_ROWS 128
#endif

namespace umb {
// Programmatic dependent launch (PDL): every kernel of the forward chain is launched with the
// programmatic-stream-serialization attribute, so its CTAs may start (barrier init, TMEM allocation,
// weight TMA + dequantisation -- nothing that depends on the predecessor) while the stream predecessor
// is still draining.  Each such kernel executes griddepcontrol.wait before touching global memory a
// predecessor wrote, and before completing (transitivity of the chain).  UMB_PDL=0 disables it.
bool pdl_enabled();
// UMB_SYNC_DEBUG=1: synchronise after every launch and name the kernel that faulted (debugging aid)
bool sync_debug();
cudaError_t sync_debug_check(const char* name, dim3 grid, dim3 block, cudaStream_t st, cudaError_t launch_rc);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(const char* name, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    const cudaError_t rc = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    return sync_debug() ? sync_debug_check(name, grid, block, st, rc) : rc;
}
#define UMB_LAUNCH_PDL(kernel, ...) do { cudaError_t e__ = umb::launch_pdl(#kernel, kernel, __VA_ARGS__); if (e__ != cudaSuccess) return (int)e__; } while (0)
// every host entry point reports the kernels it enqueued (umb_launch_count)
void count_launches(unsigned n);
// number of SMs of the current device (cached)
int sm_count();
// graph-replay variants (device-resident prefix length), used by the native forward driver
int embedding_dyn(const int64_t* ids_base, const int32_t* n_dev, int offset, const void* table, void* y, int rows,
                  int hidden, cudaStream_t st);
int rope_append_dyn(const void* qkv, int64_t ld_qkv, void* q_out, void* k_cache, void* v_cache, const void* cos,
                    const void* sin, const int64_t* position_base, const int32_t* n_dev, int offset, int qo, int Hq,
                    int Hkv, int D, int64_t kv_tok_stride, int64_t kv_head_stride, cudaStream_t st);
}  // namespace umb
