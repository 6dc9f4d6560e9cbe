// Library-internal host helpers shared by the translation units of libumbrella_b200.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/umbrella_b200.h"

namespace umb {
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
