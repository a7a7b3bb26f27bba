// Packed pairs of fp32 values for sm_100a: one FFMA2 / FMUL2 / FADD2 (PTX fma/mul/add.rn.f32x2) does two IEEE fp32
// operations per lane in ONE issue slot.  The fp32 FMA rate per SM does not change (tools/microbench/ffma2.cu on a B200:
// FFMA 4.0 warp instructions/clk/SM, FFMA2 2.0 = the same 71.5 TFLOP/s) but the issue port is freed for the loads, MUFU
// and integer instructions around them -- the fused kernels of this library are issue-bound, not FMA-pipe-bound
// (profiles/r2_ncu_cnf_*).  A scalar operand is broadcast by the instruction itself (`FFMA2 R16, R8.F32, R20.F32x2, ...`),
// so "weight x (trajectory A, trajectory B)" costs no packing move.  Each half is rounded exactly like the scalar FFMA /
// FMUL / FADD, so a kernel that puts two trajectories in the two halves produces bit-identical per-trajectory results.
#pragma once
#include <cuda_runtime.h>

namespace pnode {

using ::fma;  // the packed overloads below join, not hide, the builtin scalar ones

struct F2 {
    unsigned long long v;
};

__device__ __forceinline__ F2 pk(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 splat(float x) { return pk(x, x); }
__device__ __forceinline__ void unpk(F2 a, float &x, float &y) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
}
__device__ __forceinline__ float lo(F2 a) { return __uint_as_float((unsigned int)a.v); }
__device__ __forceinline__ float hi(F2 a) { return __uint_as_float((unsigned int)(a.v >> 32)); }

__device__ __forceinline__ F2 fma(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ F2 fma(float a, F2 b, F2 c) { return fma(splat(a), b, c); }
__device__ __forceinline__ F2 fma(F2 a, float b, F2 c) { return fma(a, splat(b), c); }
__device__ __forceinline__ F2 fma(F2 a, float b, float c) { return fma(a, splat(b), splat(c)); }
__device__ __forceinline__ F2 fma(F2 a, F2 b, float c) { return fma(a, b, splat(c)); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) {
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 operator*(float a, F2 b) { return splat(a) * b; }
__device__ __forceinline__ F2 operator*(F2 a, float b) { return a * splat(b); }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) {
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 operator+(F2 a, float b) { return a + splat(b); }
__device__ __forceinline__ F2 operator-(F2 a) {
    F2 r;
    r.v = a.v ^ 0x8000000080000000ull;
    return r;
}
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return a + (-b); }
__device__ __forceinline__ F2 &operator+=(F2 &a, F2 b) {
    a = a + b;
    return a;
}

// Packing traits: how many trajectories one thread carries and the register type that carries them.
template <typename T>
struct Pack;
template <>
struct Pack<double> {
    typedef double V;
    static constexpr int W = 1;
    static __device__ __forceinline__ V make(double a, double) { return a; }
    static __device__ __forceinline__ V all(double a) { return a; }
    static __device__ __forceinline__ double get(V v, int) { return v; }
};
template <>
struct Pack<float> {
    typedef F2 V;
    static constexpr int W = 2;
    static __device__ __forceinline__ V make(float a, float b) { return pk(a, b); }
    static __device__ __forceinline__ V all(float a) { return splat(a); }
    static __device__ __forceinline__ float get(V v, int w) { return w == 0 ? lo(v) : hi(v); }
};

}  // namespace pnode
