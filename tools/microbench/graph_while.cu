// CUDA-graph WHILE node driven from the device (cudaGraphSetConditional called by the body kernel): does it work on this
// driver, and what does one loop iteration cost next to a plain stream launch?  Basis of the on-device time loop of
// csrc/cnf_rk.cu (pnode_cnf_rk_solve).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/gw tools/microbench/graph_while.cu && /tmp/gw
#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int *counter, cudaGraphConditionalHandle h, int limit) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int c = ++(*counter);
        cudaGraphSetConditional(h, c < limit ? 1u : 0u);
    }
}
int main() {
    int *d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h;
    cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node;
    cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
    printf("add node: %s\n", cudaGetErrorString(e));
    cudaGraph_t bodyg = p.conditional.phGraph_out[0];
    cudaStream_t s; cudaStreamCreate(&s);
    // capture the body into the conditional's graph
    e = cudaStreamBeginCaptureToGraph(s, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeGlobal);
    printf("begin capture: %s\n", cudaGetErrorString(e));
    body<<<4, 32, 0, s>>>(d, h, 7);
    e = cudaStreamEndCapture(s, nullptr);
    printf("end capture: %s\n", cudaGetErrorString(e));
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0);
    printf("instantiate: %s\n", cudaGetErrorString(e));
    cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
    int hc; cudaMemcpy(&hc, d, 4, cudaMemcpyDeviceToHost);
    printf("counter = %d (expect 7)\n", hc);
    cudaMemset(d, 0, 4);
    cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
    cudaMemcpy(&hc, d, 4, cudaMemcpyDeviceToHost);
    printf("counter = %d (expect 7) err=%s\n", hc, cudaGetErrorString(cudaGetLastError()));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemsetAsync(d, 0, 4, s);
        cudaEventRecord(a, s);
        for (int i = 0; i < 20; ++i) { cudaGraphLaunch(ex, s); cudaMemsetAsync(d, 0, 4, s); }
        cudaEventRecord(b, s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("graph WHILE: %.2f us per launch of 7 iterations = %.2f us per iteration\n", ms * 1e3 / 20, ms * 1e3 / 140);
        cudaEventRecord(a, s);
        for (int i = 0; i < 140; ++i) body<<<4, 32, 0, s>>>(d, h, 1 << 30);
        cudaEventRecord(b, s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("stream launches: %.2f us per kernel\n", ms * 1e3 / 140);
    }
    return 0;
}
