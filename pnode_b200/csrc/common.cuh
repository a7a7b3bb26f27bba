// Shared helpers for the pnode_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pnode_b200.h"

namespace pnode {

void set_error(const char *fmt, ...);

#define PNODE_CUDA_OK(expr)                                                                                       \
    do {                                                                                                          \
        cudaError_t _e = (expr);                                                                                  \
        if (_e != cudaSuccess) {                                                                                  \
            pnode::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));               \
            return 1;                                                                                             \
        }                                                                                                         \
    } while (0)

#define PNODE_REQUIRE(cond, ...)                                                                                  \
    do {                                                                                                          \
        if (!(cond)) {                                                                                            \
            pnode::set_error(__VA_ARGS__);                                                                        \
            return 2;                                                                                             \
        }                                                                                                         \
    } while (0)

int sm_count();

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pnode
