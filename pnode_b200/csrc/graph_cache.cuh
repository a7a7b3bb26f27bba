// Launch sequences replayed from CUDA graphs.
//
// One right-hand-side evaluation of a convolutional ODE block (the body of the reference's evalRHSFunction,
// pnode/petsc_adjoint.py:393-412) is 6 to 17 dependent kernels of 5-30 us, one vector-Jacobian product (RHSJacShell /
// RHSJacPShell.multTranspose, 52-82, 341-363) 20 to 42: at batch 256 the pass is bound by the launches, not by the kernels
// (DESIGN.md section 4b).  The entry points therefore record their launch sequence into a CUDA graph the second time they
// see the same arguments -- every pointer and scalar that reaches a kernel is part of the key -- and replay it afterwards:
// a training loop hands the same buffers over and over (PyTorch's caching allocator returns the same blocks when the
// allocation pattern repeats), so a handful of graphs serves the whole run.
//
//  * recording happens on a private stream (the caller's stream may be the legacy default stream, which cannot be captured);
//    the graph is launched on the caller's stream, so ordering against the caller's other work is unchanged;
//  * a call made while the caller's stream is itself being captured (the caller records its whole step) launches directly;
//  * first sighting of a key launches directly (warm-up: attribute raises, occupancy queries), second sighting records;
//  * bounded: at most MAX_GRAPHS graphs, oldest dropped; if recording keeps missing (arguments never repeat) the cache
//    switches itself off.
//
// OFF by default (PNODE_CONV_GRAPHS=1, `-pnode_conv_graphs 1` or pnode_graph_cache_enable(1) turn it on).  Measured on B200
// (tools/graph_cache_probe.py, one synchronised pass of BASELINE config 4): block 1 1.97 -> 1.89 ms, block 3 2.22 -> 2.12 ms
// once the caller's buffers repeat (third pass on), with one 4 ms recording pass before that -- the kernels inside a
// sequence already overlap through programmatic dependent launches, so what a per-call graph removes is host time the GPU
// was not waiting for.  A caller who wants the launch gaps gone records the whole step (1.63 / 1.83 ms, tests/test_gpu_graph.py).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <mutex>
#include <string>
#include <unordered_map>

namespace pnode {
namespace gcache {

constexpr size_t MAX_GRAPHS = 192;

struct Entry {
    cudaGraphExec_t exec = nullptr;
    unsigned long long stamp = 0;
    int seen = 0;
};

struct Cache {
    std::mutex mu;
    std::unordered_map<std::string, Entry> map;
    unsigned long long clock = 0;
    long long hits = 0, records = 0, direct = 0;
    cudaStream_t side = nullptr;
    int enabled = -1;
};

inline Cache &cache() {
    static Cache c;
    return c;
}

class Key {
   public:
    template <typename T>
    Key &add(const T &v) {
        s_.append(reinterpret_cast<const char *>(&v), sizeof(T));
        return *this;
    }
    const std::string &str() const { return s_; }

   private:
    std::string s_;
};

inline void drop_all(Cache &c) {
    for (auto &kv : c.map)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    c.map.clear();
}

inline void evict_old(Cache &c) {
    if (c.map.size() <= MAX_GRAPHS) return;
    const unsigned long long cut = c.clock - MAX_GRAPHS / 2;  // keep the younger half
    for (auto it = c.map.begin(); it != c.map.end();) {
        if (it->second.stamp < cut) {
            if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
            it = c.map.erase(it);
        } else {
            ++it;
        }
    }
}

// launch(stream) issues the whole sequence on `stream` and returns 0 or a pnode error code.
template <typename F>
int run(const Key &key, cudaStream_t st, F &&launch) {
    Cache &c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.enabled < 0) {
        const char *e = getenv("PNODE_CONV_GRAPHS");
        c.enabled = e ? (atoi(e) != 0) : 0;
    }
    cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
    if (!c.enabled || cudaStreamIsCapturing(st, &cst) != cudaSuccess || cst != cudaStreamCaptureStatusNone) {
        if (c.enabled) cudaGetLastError();
        ++c.direct;
        return launch(st);
    }
    Entry &e = c.map[key.str()];
    e.stamp = ++c.clock;
    if (e.exec != nullptr) {
        if (cudaGraphLaunch(e.exec, st) == cudaSuccess) {
            ++c.hits;
            return 0;
        }
        cudaGetLastError();
        cudaGraphExecDestroy(e.exec);
        e.exec = nullptr;
        ++c.direct;
        return launch(st);
    }
    if (++e.seen < 2) {
        evict_old(c);
        ++c.direct;
        return launch(st);
    }
    if (c.records >= 64 && c.hits < c.records) {  // the caller's buffers never repeat: recording only costs
        c.enabled = 0;
        drop_all(c);
        ++c.direct;
        return launch(st);
    }
    if (c.side == nullptr && cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        c.enabled = 0;
        return launch(st);
    }
    if (cudaStreamBeginCapture(c.side, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        c.enabled = 0;
        return launch(st);
    }
    const int rc = launch(c.side);
    cudaGraph_t g = nullptr;
    const cudaError_t ended = cudaStreamEndCapture(c.side, &g);
    cudaGraphExec_t ex = nullptr;
    if (rc != 0 || ended != cudaSuccess || g == nullptr || cudaGraphInstantiate(&ex, g, 0) != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        c.enabled = 0;  // something in the sequence cannot be recorded: stay on direct launches
        drop_all(c);
        return rc != 0 ? rc : launch(st);
    }
    cudaGraphDestroy(g);
    ++c.records;
    if (cudaGraphLaunch(ex, st) != cudaSuccess) {
        cudaGetLastError();
        cudaGraphExecDestroy(ex);
        c.enabled = 0;
        drop_all(c);
        return launch(st);
    }
    c.map[key.str()].exec = ex;
    return 0;
}

}  // namespace gcache
}  // namespace pnode
