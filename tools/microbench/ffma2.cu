// Issue-slot microbenchmark: scalar FFMA against the packed FFMA2 (fma.rn.f32x2, sm_100a) -- alone and interleaved with
// shared-memory loads / MUFU, the mix of the fused CNF kernels (csrc/cnf_rk.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/ffma2 tools/microbench/ffma2.cu && /tmp/ffma2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
    float r;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
constexpr int CH = 8;
// MODE 0: FFMA x CH chains; 1: FFMA2 x CH chains; 2: FFMA + 1 LDS per 4; 3: FFMA2 + 1 LDS per 4; 4: FFMA + MUFU per 8;
// 5: FFMA2 + MUFU per 8
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float s) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = s + i;
    __syncthreads();
    float a[CH];
    u64 A[CH];
    for (int c = 0; c < CH; ++c) a[c] = s * c, A[c] = pk(s * c, s + c);
    float extra = s;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE & 1)
                A[c] = fma2(A[c], pk(s, s), A[(c + 1) % CH]);
            else
                a[c] = fma1(a[c], s, a[(c + 1) % CH]);
        }
        if (MODE == 2 || MODE == 3) {
            extra += sm[idx & 1023] + sm[(idx + 32) & 1023];
            idx += 64;
        }
        if (MODE == 4 || MODE == 5) {
            float e;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(extra));
            extra = e;
        }
    }
    float r = extra;
    for (int c = 0; c < CH; ++c) r += a[c] + (float)(A[c] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char *name, int fma_per_inst) {
    int iters = 20000, blocks = 148 * 8;
    float *out;
    cudaMalloc(&out, blocks * 256 * 4);
    k<MODE><<<blocks, 256>>>(out, 100, 1.0001f);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(out, iters, 1.0001f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double fmas = (double)blocks * 256 * iters * CH * fma_per_inst;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  (%.1f fma-instr/clk/SM at 1.9 GHz)\n", name, ms, 2 * fmas / ms / 1e9,
           (double)blocks * 256 / 32 * iters * CH / (ms * 1e-3 * 1.9e9) / 148);
    cudaFree(out);
}
int main() {
    run<0>("FFMA", 1);
    run<1>("FFMA2", 2);
    run<2>("FFMA + 2 LDS per 8", 1);
    run<3>("FFMA2 + 2 LDS per 8", 2);
    run<4>("FFMA + 1 MUFU per 8", 1);
    run<5>("FFMA2 + 1 MUFU per 8", 2);
    return 0;
}
