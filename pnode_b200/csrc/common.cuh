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

// ---------------------------------------------------------------------------------------------------------------------
// In-kernel one-shot all-reduce over NVLink peer memory, fused into the tail of the adjoint sweeps (the compute step
// "reduce mu over this GPU's trajectories" is followed by the collective "sum mu over GPUs": one kernel does both).
// Every rank owns a symmetric buffer (torch symmetric memory, mapped into every peer):
//     double slots[2][world][PNODE_PEER_NP_MAX];   unsigned long long flags[2][world];
// The last block of rank r stores its partial mu into slot [set][r] of EVERY rank (plain st.global through the peer
// mapping), fences system-wide, raises flag [set][r] = epoch on every rank, waits until all of its own flags reached
// epoch, and sums its own slots in rank order (identical, bit-reproducible result on every rank).  set = epoch & 1
// double-buffers consecutive collectives; epochs only grow, so flags never need a reset.
struct PeerComm {
    const unsigned long long *peer_bufs;  // device array [world] of peer-mapped base addresses (NULL: single rank)
    int rank, world;
    unsigned long long epoch;
};

__device__ __forceinline__ double *peer_slot(unsigned long long base, int set, int world, int r) {
    return reinterpret_cast<double *>(base) + ((size_t)set * world + r) * PNODE_PEER_NP_MAX;
}
__device__ __forceinline__ unsigned long long *peer_flag(unsigned long long base, int set, int world, int r) {
    return reinterpret_cast<unsigned long long *>(reinterpret_cast<double *>(base) +
                                                  (size_t)2 * world * PNODE_PEER_NP_MAX) + set * world + r;
}

// Called by ALL threads of ONE block per rank.  `mine[p]` (shared or global, p < np) holds this rank's sums.
template <typename T>
__device__ void peer_allreduce_and_store(const double *mine, int np, const PeerComm &pc, T *out) {
    const int set = (int)(pc.epoch & 1ull);
    for (int r = 0; r < pc.world; ++r) {
        double *dst = peer_slot(pc.peer_bufs[r], set, pc.world, pc.rank);
        for (int p = threadIdx.x; p < np; p += blockDim.x) dst[p] = mine[p];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < pc.world) {
        unsigned long long *f = peer_flag(pc.peer_bufs[threadIdx.x], set, pc.world, pc.rank);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(pc.epoch) : "memory");
        // wait for rank threadIdx.x's contribution to land in MY buffer
        unsigned long long *w = peer_flag(pc.peer_bufs[pc.rank], set, pc.world, threadIdx.x);
        unsigned long long v = 0;
        for (long long spin = 0; spin < (1ll << 27); ++spin) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
            if (v >= pc.epoch) break;
            __nanosleep(64);
        }
    }
    __syncthreads();
    const unsigned long long my = pc.peer_bufs[pc.rank];
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < pc.world; ++r) s += __ldcg(peer_slot(my, set, pc.world, r) + p);
        out[p] = (T)s;
    }
}

}  // namespace pnode
