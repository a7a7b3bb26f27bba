// Fused sweeps for tiny-state MLP right-hand sides  f(t,y) = W2 tanh(W1 phi(y) + b1) + b2  (spiral model,
// /root/reference/examples-pnode/ode_demo_petsc.py:207-230) under fixed-step explicit Runge-Kutta.
//
// What it replaces in the reference: the whole TSSolve / TSAdjointSolve host loops of PETSc plus the Python callbacks
// evalRHSFunction (pnode/petsc_adjoint.py:393-412), RHSJacShell.multTranspose (52-82) and RHSJacPShell.multTranspose
// (341-363), i.e. >= 10 interpreter calls and >= 6 kernel launches per stage for a 2-float state.  Here ONE launch
// does every step and stage of the forward sweep and ONE launch does the whole discrete adjoint.
//
// B200 mapping
//  * forward: TPT trajectories per thread, weights (252 scalars, packed per hidden unit) broadcast from shared memory,
//    the s stage slopes in registers, stage values Y_i streamed to HBM as [step][stage][dim][traj] (coalesced; 8 scalars
//    per trajectory-step for RK4) -- the "stage checkpoints in HBM" that replace -ts_trajectory_type memory.
//  * adjoint: lambda recurrence and state VJP with TPT trajectories per thread; the parameter gradient (a
//    [batch x 5] x [batch x H] outer-product sum) is reduced without atomics or shuffles: each warp parks tanh(z_j) and s_j
//    of its 32*TPT trajectories in a warp-private shared-memory tile, then re-reads the tile TRANSPOSED (lane = hidden
//    unit j, 16-byte loads along the trajectory axis) so that every lane accumulates "its" five parameter gradients in
//    registers.  Only __syncwarp() separates the phases, warps never wait for each other, the per-lane accumulators
//    (double) live across all stages, steps and tiles of the persistent kernel.  Per-block partials are combined in a
//    fixed order by the last block (bit-reproducible mu).
//  * both kernels are bound by instruction issue of tanh (fp64: FP64 pipe; fp32: MUFU/issue), not by HBM (arithmetic
//    intensity ~80 flop/B): the work went into shrinking tanh -- fp64: 1024-entry 2^(j/1024) table in
//    SMEM + degree-3 polynomial + cubic reciprocal refinement = 11 FP64 ops (see PNODE_F64_TANH_V below); fp32: one
//    MUFU.EX2 per unit and ONE shared MUFU.RCP per four units (product trick).  Grids are persistent: SMs x resident CTAs.
#include <stdlib.h>

#include "common.cuh"

namespace pnode {

// ---------------------------------------------------------------------------------------------------------------------
// tanh

// Two evaluations of the fp64 tanh (PNODE_F64_TANH_V).  ncu on the fp64 sweeps: the warp schedulers are busy 98.5 % of the
// cycles in the forward sweep (an FP64 instruction holds the dispatch port for two cycles: issue-active 64.7 % + FP64 33.9 %)
// and 81 % in the adjoint -- the sweeps are bound by the NUMBER of instructions, two slots per FP64 instruction, one per
// integer / load instruction, and a tanh is four fifths of them.
//   0  relative accuracy everywhere: em1 / (em1 + 2), 256-entry table, degree-4 polynomial, two-step argument reduction:
//      15 FP64 + 9 integer instructions + 1 table load + 1 MUFU = 41 slots
//   1  absolute accuracy (6e-16, all the state and the gradients need: tanh enters through W2 a and 1 - a^2):
//      1 - 2 / (e^{2x} + 1), 1024-entry table (8 KB), degree-3 polynomial whose r^2 coefficient absorbs most of the r^4 term,
//      one-step argument reduction (the product kf * c is exact inside the FMA; c is ln2/2048 correctly rounded), clamp as
//      one FMNMX + one LOP3 on the high word: 11 FP64 + 7 integer + 1 load + 1 MUFU = 31 slots
#ifndef PNODE_F64_TANH_V
#define PNODE_F64_TANH_V 1
#endif
#if PNODE_F64_TANH_V == 0
constexpr int EXP_TAB = 256;  // 2^(j/256), j = 0..255  (2 KB of shared memory per CTA)
constexpr int EXP_TAB_LOG2 = 8;
#else
constexpr int EXP_TAB = 1024;  // 2^(j/1024)  (8 KB of shared memory per CTA)
constexpr int EXP_TAB_LOG2 = 10;
#endif

// |x| clamped to 24 on the high word (tanh(24) rounds to 1; keeps k = rint(x 512/ln2) small; NaN also saturates -- the
// state that produced it stays NaN through the AXPYs, so divergence is still visible to the caller)
__device__ __forceinline__ double clamp24(double x) {
#if PNODE_F64_TANH_V == 0
    const int hx = __double2hiint(x);
    const int ha = min(hx & 0x7fffffff, 0x40380000);
    return __hiloint2double(ha | (hx & 0x80000000), __double2loint(x));
#else
    // the magnitude of a double's high word orders like a float with the same bits: 0x40380000 (24.0) reads 2.875f; one FMNMX
    // with |.| on its source (high words that read as float NaN -- |x| >= 2^1017, inf, NaN -- come out as 2.875f), one LOP3
    const int hx = __double2hiint(x);
    const float ha = fminf(fabsf(__int_as_float(hx)), 2.875f);
    return __hiloint2double(__float_as_int(ha) | (hx & 0x80000000), __double2loint(x));
#endif
}

#if PNODE_F64_TANH_V == 0
// fp64: tanh(x) = em1 / (em1 + 2), em1 = e^{2x} - 1 without cancellation.
//   2x = k ln2/256 + r, |r| <= ln2/512;  e^{2x} = 2^(k>>8) * T[k&255] * (1 + P(r)),  P(r) = e^r - 1 (degree 4: truncation
//   r^5/120 < 4e-17 absolute, < 3e-14 relative to tanh near 0)
//   em1 = (s - 1) + s P with s = T 2^(k>>8): exact for k == 0, so small |x| keeps its relative accuracy.
__device__ __forceinline__ double tanh_acc(double xin, const double *__restrict__ tab) {
    const double MAGIC = 6755399441055744.0;          // 1.5 * 2^52
    const double x = clamp24(xin);
    double kf = fma(x, 738.6598609351493, MAGIC);     // 512 / ln2
    const int k = __double2loint(kf);
    kf -= MAGIC;
    double rh = fma(kf, -0.001353803086658445, x);    // ln2/512 hi (0x1.62e42fee00000p-10: 21 trailing zero bits)
    rh = fma(kf, -3.7269822837316166e-13, rh);        // ln2/512 lo ;  r = 2 rh
    // P(r), r = 2 rh:  r + r^2/2 + r^3/6 + r^4/24 = rh (2 + rh (2 + rh (4/3 + rh 2/3)))
    double p = fma(rh, 0.6666666666666666, 1.3333333333333333);
    p = fma(p, rh, 2.0);
    p = fma(p, rh, 2.0);
    p *= rh;
    const double T = tab[k & (EXP_TAB - 1)];
    const double s = __hiloint2double(__double2hiint(T) + ((k >> EXP_TAB_LOG2) << 20), __double2loint(T));
    const double em1 = fma(s, p, s - 1.0);
    const double d = em1 + 2.0;
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
    const double e0 = fma(-d, r0, 1.0);
    const double rc = fma(r0, fma(e0, e0, e0), r0);   // cubic refinement: error e0^3
    return em1 * rc;
}
#else
// fp64: tanh(x) = 1 - 2 / (e^{2x} + 1).
//   2x = k ln2/1024 + r, |r| <= ln2/2048;  e^{2x} = 2^(k>>10) * T[k&1023] * (1 + P(r));  with rh = r/2:
//   e^r = 1 + rh (2 + rh (C2 + rh 4/3)),  C2 = 2 + 4 delta, delta = (sqrt2 - 1) h^2 / 12 (h = ln2/2048): the r^2 coefficient takes
//   the r^4/24 term of e^r with it, leaving 0.0071 h^4 = 1e-16;  e^{2x} + 1 = fma(s, e^r, 1).
constexpr double TANH_C = 2954.639443740597;         // 2048 / ln2
constexpr double TANH_NEG_STEP = -0.0003384507717577858;  // -ln2 / 2048, correctly rounded
constexpr double TANH_C2 = 2.000000015815906;
// scaled table value 2^(k/1024): the exponent part of k is added into the high word ((k >> 10) << 20 = (8k - 8(k & 1023)) << 7)
__device__ __forceinline__ double exp_tab(const double *__restrict__ tab, int k) {
    const int a8 = k << 3, i8 = a8 & ((EXP_TAB - 1) << 3);
    const double T = *reinterpret_cast<const double *>(reinterpret_cast<const char *>(tab) + i8);
    return __hiloint2double(__double2hiint(T) + ((a8 - i8) << (20 - EXP_TAB_LOG2 - 3)), __double2loint(T));
}
__device__ __forceinline__ double tanh_acc(double xin, const double *__restrict__ tab) {
    const double MAGIC = 6755399441055744.0;          // 1.5 * 2^52
    const double x = clamp24(xin);
    double kf = fma(x, TANH_C, MAGIC);
    const int k = __double2loint(kf);
    kf -= MAGIC;
    const double rh = fma(kf, TANH_NEG_STEP, x);
    double p = fma(rh, 1.3333333333333333, TANH_C2);
    p = fma(p, rh, 2.0);
    p = fma(p, rh, 1.0);                               // e^r
    const double s = exp_tab(tab, k);
    const double d = fma(s, p, 1.0);
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
    const double e0 = fma(-d, r0, 1.0);
    const double rc = fma(r0, fma(e0, e0, e0), r0);   // cubic refinement: error e0^3
    return fma(-2.0, rc, 1.0);
}
#endif

__device__ __forceinline__ float ex2_approx(float a) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
    return e;
}
__device__ __forceinline__ float rcp_approx(float d) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
}

// fp32: tanh(z) = 1 - 2/(1 + 2^(2 log2e z)); the argument is clamped to +-30 (tanh == +-1 in fp32 beyond) so that products
// of up to four denominators stay finite.
__device__ __forceinline__ float tanh_den(float z) {
    float a = fminf(fmaxf(z * 2.8853900817779268f, -30.0f), 30.0f);
    return 1.0f + ex2_approx(a);
}
__device__ __forceinline__ float tanh_acc(float z, const float *) { return fmaf(-2.0f, rcp_approx(tanh_den(z)), 1.0f); }

// G tanh's at once.  fp32: one reciprocal for the whole group; fp64: independent.
template <int G>
__device__ __forceinline__ void tanh_group(const float (&z)[G], float (&a)[G], const float *) {
    float d[G];
#pragma unroll
    for (int g = 0; g < G; ++g) d[g] = tanh_den(z[g]);
    if (G == 4) {
        const float p01 = d[0] * d[1], p23 = d[2] * d[3];
        const float r = rcp_approx(p01 * p23);
        const float r01 = r * p23, r23 = r * p01;
        a[0] = fmaf(-2.0f, r01 * d[1], 1.0f);
        a[1] = fmaf(-2.0f, r01 * d[0], 1.0f);
        a[2 % G] = fmaf(-2.0f, r23 * d[3 % G], 1.0f);
        a[3 % G] = fmaf(-2.0f, r23 * d[2 % G], 1.0f);
    } else if (G == 2) {
        const float r = rcp_approx(d[0] * d[1 % G]);
        a[0] = fmaf(-2.0f, r * d[1 % G], 1.0f);
        a[1 % G] = fmaf(-2.0f, r * d[0], 1.0f);
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) a[g] = fmaf(-2.0f, rcp_approx(d[g]), 1.0f);
    }
}
// fp64 group: the same arithmetic as tanh_acc(double), written step-by-step ACROSS the group so that the G dependent
// chains are interleaved in program order (ptxas keeps them interleaved; per-element inlining left Horner chains
// back-to-back and the warps stalled on the FP64 pipe latency -- profiles/r1 "stall_wait").
template <int G>
__device__ __forceinline__ void tanh_group(const double (&z)[G], double (&a)[G], const double *__restrict__ tab) {
    const double MAGIC = 6755399441055744.0;
#if PNODE_F64_TANH_V == 0
    double x[G], kf[G], rh[G], p[G], s[G], em1[G], d[G], r0[G], e0[G];
    int k[G];
#pragma unroll
    for (int g = 0; g < G; ++g) x[g] = clamp24(z[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) kf[g] = fma(x[g], 738.6598609351493, MAGIC);
#pragma unroll
    for (int g = 0; g < G; ++g) {
        k[g] = __double2loint(kf[g]);
        kf[g] -= MAGIC;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const double T = tab[k[g] & (EXP_TAB - 1)];
        s[g] = __hiloint2double(__double2hiint(T) + ((k[g] >> EXP_TAB_LOG2) << 20), __double2loint(T));
    }
#pragma unroll
    for (int g = 0; g < G; ++g) rh[g] = fma(kf[g], -0.001353803086658445, x[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) rh[g] = fma(kf[g], -3.7269822837316166e-13, rh[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) p[g] = fma(rh[g], 0.6666666666666666, 1.3333333333333333);
#pragma unroll
    for (int g = 0; g < G; ++g) p[g] = fma(p[g], rh[g], 2.0);
#pragma unroll
    for (int g = 0; g < G; ++g) p[g] = fma(p[g], rh[g], 2.0);
#pragma unroll
    for (int g = 0; g < G; ++g) p[g] *= rh[g];
#pragma unroll
    for (int g = 0; g < G; ++g) em1[g] = fma(s[g], p[g], s[g] - 1.0);
#pragma unroll
    for (int g = 0; g < G; ++g) d[g] = em1[g] + 2.0;
#pragma unroll
    for (int g = 0; g < G; ++g) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0[g]) : "d"(d[g]));
#pragma unroll
    for (int g = 0; g < G; ++g) e0[g] = fma(-d[g], r0[g], 1.0);
#pragma unroll
    for (int g = 0; g < G; ++g) r0[g] = fma(r0[g], fma(e0[g], e0[g], e0[g]), r0[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) a[g] = em1[g] * r0[g];
#else
    double x[G], kf[G], rh[G], p[G], s[G], d[G], r0[G], e0[G];
    int k[G];
#pragma unroll
    for (int g = 0; g < G; ++g) x[g] = clamp24(z[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) kf[g] = fma(x[g], TANH_C, MAGIC);
#pragma unroll
    for (int g = 0; g < G; ++g) {
        k[g] = __double2loint(kf[g]);
        kf[g] -= MAGIC;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) s[g] = exp_tab(tab, k[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) rh[g] = fma(kf[g], TANH_NEG_STEP, x[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) p[g] = fma(rh[g], 1.3333333333333333, TANH_C2);
#pragma unroll
    for (int g = 0; g < G; ++g) p[g] = fma(p[g], rh[g], 2.0);
#pragma unroll
    for (int g = 0; g < G; ++g) p[g] = fma(p[g], rh[g], 1.0);
#pragma unroll
    for (int g = 0; g < G; ++g) d[g] = fma(s[g], p[g], 1.0);
#pragma unroll
    for (int g = 0; g < G; ++g) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0[g]) : "d"(d[g]));
#pragma unroll
    for (int g = 0; g < G; ++g) e0[g] = fma(-d[g], r0[g], 1.0);
#pragma unroll
    for (int g = 0; g < G; ++g) r0[g] = fma(r0[g], fma(e0[g], e0[g], e0[g]), r0[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) a[g] = fma(-2.0, r0[g], 1.0);
#endif
}

// ---------------------------------------------------------------------------------------------------------------------

template <typename T>
struct MlpPtrs {
    const T *w1, *b1, *w2, *b2;
    int slot;  // __constant__ weight slot of this launch
};

// per-hidden-unit weights packed for vector LDS: w1[j][0..D), b1[j], w2[0..D)[j]
template <typename T, int D>
struct alignas(2 * sizeof(T)) Unit {
    T w1[D];
    T b1;
    T w2[D];
    T pad[(2 * D + 1) % 2];
};

// tuning knobs (overridable with -D for tools/tune_spiral.py)
#ifndef PNODE_F32_TPT
#define PNODE_F32_TPT 2
#endif
#ifndef PNODE_F32_GROUP
#define PNODE_F32_GROUP 4
#endif
#ifndef PNODE_F32_ADJ_GROUP
#define PNODE_F32_ADJ_GROUP PNODE_F32_GROUP
#endif
#ifndef PNODE_F64_ADJ_GROUP
#define PNODE_F64_ADJ_GROUP PNODE_F64_GROUP
#endif
#ifndef PNODE_F32_CHUNKS
#define PNODE_F32_CHUNKS 2
#endif
#ifndef PNODE_F32_FWD_CTAS
#define PNODE_F32_FWD_CTAS 8
#endif
#ifndef PNODE_F32_ADJ_CTAS
#define PNODE_F32_ADJ_CTAS 3
#endif
#ifndef PNODE_F64_TPT
#define PNODE_F64_TPT 1
#endif
#ifndef PNODE_F64_GROUP
#define PNODE_F64_GROUP 5
#endif
#ifndef PNODE_F64_CHUNKS
#define PNODE_F64_CHUNKS 2
#endif
#ifndef PNODE_F64_FWD_CTAS
#define PNODE_F64_FWD_CTAS 5
#endif
#ifndef PNODE_F64_ADJ_CTAS
#define PNODE_F64_ADJ_CTAS 3
#endif
#ifndef PNODE_ADJ_ROLL
#define PNODE_ADJ_ROLL 1
#endif

template <typename T>
struct Cfg;
template <>
struct Cfg<float> {
    static constexpr int TPT = PNODE_F32_TPT;            // trajectories per thread
    static constexpr int GROUP = PNODE_F32_GROUP;        // hidden units sharing one reciprocal
    static constexpr int ADJ_GROUP = PNODE_F32_ADJ_GROUP;
    static constexpr int ADJ_CHUNKS = PNODE_F32_CHUNKS;  // hidden units in chunks of <= 32 (lane = unit in phase 2)
    static constexpr int FWD_MIN_CTAS = PNODE_F32_FWD_CTAS;
    static constexpr int ADJ_MIN_CTAS = PNODE_F32_ADJ_CTAS;
};
template <>
struct Cfg<double> {
    static constexpr int TPT = PNODE_F64_TPT;
    static constexpr int GROUP = PNODE_F64_GROUP;
    static constexpr int ADJ_GROUP = PNODE_F64_ADJ_GROUP;
    static constexpr int ADJ_CHUNKS = PNODE_F64_CHUNKS;
    static constexpr int FWD_MIN_CTAS = PNODE_F64_FWD_CTAS;
    static constexpr int ADJ_MIN_CTAS = PNODE_F64_ADJ_CTAS;
};

// EXPERIMENT (off by default, -DPNODE_CONST_WEIGHTS=1): weights in __constant__ memory, fed to FFMA/DFMA through the
// uniform datapath (LDCU -> UR operands) instead of shared-memory broadcasts.  Measured on B200 at 2^20 trajectories: fp32
// forward 1.07 -> 1.03 ms, adjoint 2.60 -> 2.75 ms; fp64 forward 2.99 -> 11.4 ms (252 doubles + kernel parameters overflow
// the per-SM immediate-constant cache and every hidden unit misses), adjoint 5.62 -> 5.79 ms.  Shared memory stays.
// Layout per slot = the flat parameter order W1[H][D] | b1[H] | W2[D][H] | b2[D]; slots are used round-robin so that sweeps
// enqueued on different streams would not overwrite each other's weights.
#ifndef PNODE_CONST_WEIGHTS
#define PNODE_CONST_WEIGHTS 0
#endif
constexpr int W_SLOTS = 4;
#if PNODE_CONST_WEIGHTS
constexpr int W_MAXP = 2 * 50 * 2 + 50 + 2;
__constant__ float cWf[W_SLOTS][W_MAXP];
__constant__ double cWd[W_SLOTS][W_MAXP];
template <typename T>
__device__ __forceinline__ T cwget(int slot, int idx);
template <>
__device__ __forceinline__ float cwget<float>(int slot, int idx) {
    return cWf[slot][idx];
}
template <>
__device__ __forceinline__ double cwget<double>(int slot, int idx) {
    return cWd[slot][idx];
}
#endif

template <typename T, int D, int H>
__device__ __forceinline__ Unit<T, D> get_unit(const Unit<T, D> *__restrict__ sW, int slot, int j) {
#if PNODE_CONST_WEIGHTS
    Unit<T, D> u;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        u.w1[d] = cwget<T>(slot, j * D + d);
        u.w2[d] = cwget<T>(slot, H * D + H + d * H + j);
    }
    u.b1 = cwget<T>(slot, H * D + j);
    return u;
#else
    return sW[j];
#endif
}
template <typename T, int D, int H>
__device__ __forceinline__ T get_b2(const T *__restrict__ sB2, int slot, int d) {
#if PNODE_CONST_WEIGHTS
    return cwget<T>(slot, H * D + H + D * H + d);
#else
    return sB2[d];
#endif
}

template <typename T, int D, int H>
__device__ __forceinline__ void load_weights(Unit<T, D> *sW, T *sB2, T *sTab, const MlpPtrs<T> &w) {
#if !PNODE_CONST_WEIGHTS
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        Unit<T, D> u;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            u.w1[d] = w.w1[j * D + d];
            u.w2[d] = w.w2[d * H + j];
        }
        u.b1 = w.b1[j];
        sW[j] = u;
    }
    if (threadIdx.x < D) sB2[threadIdx.x] = w.b2[threadIdx.x];
#endif
    if (sizeof(T) == 8)
        for (int j = threadIdx.x; j < EXP_TAB; j += blockDim.x) sTab[j] = (T)exp2((double)j / EXP_TAB);
}

template <typename T, int D, int PHI>
__device__ __forceinline__ void apply_phi(const T (&y)[D], T (&x)[D]) {
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = (PHI == 1) ? y[d] * y[d] * y[d] : y[d];
}

// out[q] = f(y[q]) for the thread's TPT trajectories; hidden units in groups of G (weights loaded once per group).
template <typename T, int D, int H, int PHI, int TPT, int G>
__device__ __forceinline__ void mlp_eval_units(const Unit<T, D> *__restrict__ sW, const T *__restrict__ sTab, int slot,
                                               int j0, const T (&x)[TPT][D], T (&out)[TPT][D]) {
    Unit<T, D> u[G];
#pragma unroll
    for (int g = 0; g < G; ++g) u[g] = get_unit<T, D, H>(sW, slot, j0 + g);
#pragma unroll
    for (int q = 0; q < TPT; ++q) {
        T z[G], a[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            z[g] = u[g].b1;
#pragma unroll
            for (int d = 0; d < D; ++d) z[g] = fma(u[g].w1[d], x[q][d], z[g]);
        }
        tanh_group<G>(z, a, sTab);
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int d = 0; d < D; ++d) out[q][d] = fma(u[g].w2[d], a[g], out[q][d]);
    }
}

template <typename T, int D, int H, int PHI, int TPT>
__device__ __forceinline__ void mlp_eval(const Unit<T, D> *__restrict__ sW, const T *__restrict__ sB2,
                                         const T *__restrict__ sTab, int slot, const T (&y)[TPT][D],
                                         T (&out)[TPT][D]) {
    constexpr int G = Cfg<T>::GROUP;
    T x[TPT][D];
#pragma unroll
    for (int q = 0; q < TPT; ++q) {
        apply_phi<T, D, PHI>(y[q], x[q]);
#pragma unroll
        for (int d = 0; d < D; ++d) out[q][d] = get_b2<T, D, H>(sB2, slot, d);
    }
    constexpr int NG = H / G;
#pragma unroll 1
    for (int gi = 0; gi < NG; ++gi) mlp_eval_units<T, D, H, PHI, TPT, G>(sW, sTab, slot, gi * G, x, out);
    constexpr int R = H - NG * G;
    if (R >= 2) mlp_eval_units<T, D, H, PHI, TPT, (R >= 2 ? 2 : 1)>(sW, sTab, slot, NG * G, x, out);
    if (R == 3) mlp_eval_units<T, D, H, PHI, TPT, 1>(sW, sTab, slot, NG * G + 2, x, out);
    if (R == 1) mlp_eval_units<T, D, H, PHI, TPT, 1>(sW, sTab, slot, NG * G, x, out);
    static_assert(R < 4 || G > 4, "remainder handling assumes GROUP <= 4 or H % GROUP == 0");
    static_assert(G <= 4 || H % G == 0, "GROUP > 4 needs H % GROUP == 0");
}

// ---------------------------------------------------------------------------------------------------------------------
// forward sweep

constexpr int FWD_THREADS = 128;

template <typename T, int D, int H, int S, int PHI>
__global__ void __launch_bounds__(FWD_THREADS, (D == 2 && H == 50) ? Cfg<T>::FWD_MIN_CTAS : 2)
mlp_rk_fwd_kernel(const MlpPtrs<T> w, const pnode_rk_tableau tab, const T *__restrict__ u0, const int64_t ntraj,
                  const pnode_step *__restrict__ sched, const int nsteps, T *__restrict__ sol, T *__restrict__ ckpt,
                  const int solution_only) {
    constexpr int TPT = Cfg<T>::TPT;
    __shared__ Unit<T, D> sW[H];
    __shared__ T sB2[D];
    __shared__ T sTab[EXP_TAB];
    load_weights<T, D, H>(sW, sB2, sTab, w);
    __syncthreads();

    const int64_t tile = (int64_t)FWD_THREADS * TPT;
    const int64_t ntiles = (ntraj + tile - 1) / tile;
    for (int64_t tidx = blockIdx.x; tidx < ntiles; tidx += gridDim.x) {
        int64_t traj[TPT];
        bool valid[TPT];
        T y[TPT][D];
#pragma unroll
        for (int q = 0; q < TPT; ++q) {
            traj[q] = tidx * tile + q * FWD_THREADS + threadIdx.x;
            valid[q] = traj[q] < ntraj;
#pragma unroll
            for (int d = 0; d < D; ++d) y[q][d] = valid[q] ? u0[traj[q] * D + d] : T(0);
        }
        T K[S][TPT][D];
        for (int n = 0; n < nsteps; ++n) {
            const double h = sched[n].h;
            const int out_slot = sched[n].out_slot;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                T Y[TPT][D];
#pragma unroll
                for (int q = 0; q < TPT; ++q)
#pragma unroll
                    for (int d = 0; d < D; ++d) Y[q][d] = y[q][d];
#pragma unroll
                for (int j = 0; j < i; ++j) {
                    const T ha = (T)(h * tab.a[i][j]);
#pragma unroll
                    for (int q = 0; q < TPT; ++q)
#pragma unroll
                        for (int d = 0; d < D; ++d) Y[q][d] = fma(ha, K[j][q][d], Y[q][d]);
                }
                // -ts_trajectory_solution_only: keep u_n alone ([step][dim][traj]); the adjoint sweep recomputes the stages
                if (ckpt != nullptr && (!solution_only || i == 0)) {
                    const int64_t row = solution_only ? (int64_t)n : (int64_t)n * S + i;
#pragma unroll
                    for (int q = 0; q < TPT; ++q)
                        if (valid[q]) {
#pragma unroll
                            for (int d = 0; d < D; ++d) ckpt[(row * D + d) * ntraj + traj[q]] = Y[q][d];
                        }
                }
                if (i == 0 && tab.fsal && n > 0) {
#pragma unroll
                    for (int q = 0; q < TPT; ++q)
#pragma unroll
                        for (int d = 0; d < D; ++d) K[0][q][d] = K[S - 1][q][d];
                } else {
                    mlp_eval<T, D, H, PHI, TPT>(sW, sB2, sTab, w.slot, Y, K[i]);
                }
            }
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const T hb = (T)(h * tab.b[j]);
#pragma unroll
                for (int q = 0; q < TPT; ++q)
#pragma unroll
                    for (int d = 0; d < D; ++d) y[q][d] = fma(hb, K[j][q][d], y[q][d]);
            }
            if (out_slot >= 0) {
#pragma unroll
                for (int q = 0; q < TPT; ++q)
                    if (valid[q]) {
#pragma unroll
                        for (int d = 0; d < D; ++d) sol[((int64_t)out_slot * ntraj + traj[q]) * D + d] = y[q][d];
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// adjoint sweep

constexpr int ADJ_WARPS = 4;
constexpr int ADJ_THREADS = ADJ_WARPS * 32;
constexpr int ADJ_MAX_BLOCKS = 148 * 8;

template <typename T, int D, int H>
struct AdjShape {
    static constexpr int TPT = Cfg<T>::TPT;
    // hidden units in chunks of <= 32 (lane = unit in phase 2): the tuned count for H = 50, one chunk per 25 units beyond
    static constexpr int NCHUNK = (H <= 50) ? Cfg<T>::ADJ_CHUNKS : (H + 24) / 25;
    static constexpr int G = Cfg<T>::ADJ_GROUP;
    // hidden units per chunk: <= 32 (lane = unit in phase 2), a multiple of the tanh group width
    static constexpr int JH = ((H + NCHUNK - 1) / NCHUNK + G - 1) / G * G;
    static constexpr int NK = 32 * TPT;                       // trajectories per warp tile
    static constexpr int VEC = 16 / sizeof(T);                // scalars per 16-byte shared load
    static constexpr int PITCH = NK + VEC;                    // rows 16B-aligned; (PITCH/VEC) odd => conflict-free
    static constexpr int NP = 2 * H * D + H + D;
    static_assert(JH <= 32, "chunk too wide");
    static_assert(((PITCH / VEC) & 1) == 1, "pitch must be an odd number of 16-byte words");
};

struct AdjWork {
    unsigned int ticket;
    unsigned int pad[15];
    double partial[1];  // [blocks][NP]
};

// PNODE_ADJ_RECOMPUTE_S 1: phase 2 forms s_j = (W2[:,j] . v)(1 - a_j^2) again from the parked a_j, the lane's own W2 column and
// the broadcast v instead of reading a parked copy -- the same three instructions as phase 1, so the same bits.  The fp64
// adjoint sweep issues 0.91 shared-memory wavefronts per cycle and is bound there: the parked copy costs 2.1 (store) + 2.6
// (transposed re-read) of its 25.2 wavefronts per hidden unit and warp, the recomputation four FP64 instructions per
// (unit, trajectory) in lanes that were waiting for the loads.
// fp64 only: measured at 2^20 trajectories 5.55 -> 5.30 ms; the fp32 sweep (MUFU / issue bound) 2.61 -> 2.66 ms, so it keeps
// the parked copy.
#ifndef PNODE_ADJ_RECOMPUTE_S
#define PNODE_ADJ_RECOMPUTE_S 1
#endif
template <typename T>
struct AdjRecompute {
    static constexpr bool value = PNODE_ADJ_RECOMPUTE_S && sizeof(T) == 8;
};

// With s recomputed the tile has room for the tanh values of BOTH 25-unit chunks of a stage (the footprint the parked copy of
// s had): one phase 2 per stage in which a lane serves its unit of every chunk, so the v / x broadcasts -- 8 bytes per
// shared-memory wavefront whatever the load width, 5.1 wavefronts per hidden unit and warp -- are read once per stage.
template <typename T, int D, int H>
struct AdjMerge {
    static constexpr bool value = AdjRecompute<T>::value && AdjShape<T, D, H>::NCHUNK == 2 && AdjShape<T, D, H>::TPT == 1;
    static constexpr int ROWS = (value ? AdjShape<T, D, H>::NCHUNK : 1) * AdjShape<T, D, H>::JH;
};

template <typename T, int D, int H>
struct alignas(16) WarpTile {
    T A[AdjMerge<T, D, H>::ROWS * AdjShape<T, D, H>::PITCH];  // tanh(z_j)           per (unit, trajectory)
    // s_j = g_j(1-a_j^2) per (unit, trajectory); one 16-byte placeholder when phase 2 recomputes it
    T Sg[AdjRecompute<T>::value ? AdjShape<T, D, H>::VEC : AdjShape<T, D, H>::JH * AdjShape<T, D, H>::PITCH];
    T V[D][AdjShape<T, D, H>::NK];                           // stage cotangent v, per trajectory
    T X[D][AdjShape<T, D, H>::NK];                           // phi(Y_i)
};

template <typename T>
struct VecT;
template <>
struct VecT<float> {
    typedef float4 type;
};
template <>
struct VecT<double> {
    typedef double2 type;
};
template <typename T>
__device__ __forceinline__ void lds16(const T *p, T (&r)[16 / sizeof(T)]) {
    typename VecT<T>::type v = *reinterpret_cast<const typename VecT<T>::type *>(p);
    memcpy(r, &v, 16);
}

template <typename T, int NP>
__device__ void adj_finish(double *blk, int nwarps, AdjWork *__restrict__ work, T *__restrict__ mu_out, const PeerComm &pc,
                           bool *is_last);

// SO: the forward sweep kept u_n per step only (-ts_trajectory_solution_only 1, [step][dim][traj]); the stage values of a
// step are recomputed from u_n with the forward sweep's own arithmetic before its adjoint stages run.
template <typename T, int D, int H, int S, int PHI, bool SO>
__global__ void __launch_bounds__(ADJ_THREADS, (D == 2 && H == 50) ? Cfg<T>::ADJ_MIN_CTAS : 2)
mlp_rk_adj_kernel(const MlpPtrs<T> w, const pnode_rk_tableau tab, const int64_t ntraj,
                  const pnode_step *__restrict__ sched, const int nsteps, const int last_slot,
                  const T *__restrict__ gout, const T *__restrict__ ckpt, T *__restrict__ lambda_out,
                  T *__restrict__ mu_out, AdjWork *__restrict__ work, const PeerComm pc) {
    typedef AdjShape<T, D, H> Sh;
    constexpr int TPT = Sh::TPT, NCHUNK = Sh::NCHUNK, JH = Sh::JH, PITCH = Sh::PITCH, NP = Sh::NP, NK = Sh::NK;
    constexpr int VEC = Sh::VEC, G = Cfg<T>::ADJ_GROUP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Unit<T, D> *sW = reinterpret_cast<Unit<T, D> *>(smem_raw);
    T *sB2 = reinterpret_cast<T *>(sW + H);
    T *sTab = sB2 + D;
    WarpTile<T, D, H> *tiles = reinterpret_cast<WarpTile<T, D, H> *>(
        smem_raw + ((sizeof(Unit<T, D>) * H + sizeof(T) * (D + EXP_TAB) + 15) / 16) * 16);
    __shared__ bool is_last;

    load_weights<T, D, H>(sW, sB2, sTab, w);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpTile<T, D, H> &tile = tiles[warp];

    // per-lane parameter-gradient accumulators (lane = hidden unit within chunk), double regardless of T
    double accW1[NCHUNK][D], accB1[NCHUNK], accW2[NCHUNK][D], accB2[D];
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        accB1[c] = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) accW1[c][d] = accW2[c][d] = 0.0;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) accB2[d] = 0.0;

    const int64_t tile_traj = (int64_t)ADJ_THREADS * TPT;
    const int64_t ntiles = (ntraj + tile_traj - 1) / tile_traj;
    for (int64_t tidx = blockIdx.x; tidx < ntiles; tidx += gridDim.x) {
        // trajectory q of this thread sits at tile column k = q*32 + lane of the warp's tile
        int64_t traj[TPT];
        bool valid[TPT];
        T lam[TPT][D];
#pragma unroll
        for (int q = 0; q < TPT; ++q) {
            traj[q] = tidx * tile_traj + (int64_t)warp * NK + q * 32 + lane;
            valid[q] = traj[q] < ntraj;
#pragma unroll
            for (int d = 0; d < D; ++d)
                lam[q][d] = valid[q] ? gout[((int64_t)last_slot * ntraj + traj[q]) * D + d] : T(0);
        }

        // stage values are fetched one stage ahead of their use (hides the HBM latency of the checkpoint stream)
        const int s_top = (tab.fsal ? S - 2 : S - 1);
        T Ynext[TPT][D];
        auto fetch_Y = [&](int nn, int ii) {
            const int64_t row = SO ? (int64_t)nn : (int64_t)nn * S + ii;
#pragma unroll
            for (int q = 0; q < TPT; ++q)
#pragma unroll
                for (int d = 0; d < D; ++d)
                    Ynext[q][d] = (valid[q] && nn >= 0) ? ckpt[(row * D + d) * ntraj + traj[q]] : T(0);
        };
        fetch_Y(nsteps - 1, s_top);

        for (int n = nsteps - 1; n >= 0; --n) {
            const double h = sched[n].h;
            const int in_slot = sched[n].in_slot;
            T ls[S][TPT][D];
            T Yst[SO ? S : 1][TPT][D];
            if (SO) {
                T yn[TPT][D], K[S][TPT][D];
#pragma unroll
                for (int q = 0; q < TPT; ++q)
#pragma unroll
                    for (int d = 0; d < D; ++d) yn[q][d] = Ynext[q][d];
                fetch_Y(n - 1, 0);  // u_{n-1} is on its way while this step's stages are recomputed
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    T Yi[TPT][D];
#pragma unroll
                    for (int q = 0; q < TPT; ++q)
#pragma unroll
                        for (int d = 0; d < D; ++d) Yi[q][d] = yn[q][d];
#pragma unroll
                    for (int j = 0; j < i; ++j) {
                        const T ha = (T)(h * tab.a[i][j]);
#pragma unroll
                        for (int q = 0; q < TPT; ++q)
#pragma unroll
                            for (int d = 0; d < D; ++d) Yi[q][d] = fma(ha, K[j][q][d], Yi[q][d]);
                    }
#pragma unroll
                    for (int q = 0; q < TPT; ++q)
#pragma unroll
                        for (int d = 0; d < D; ++d) Yst[SO ? i : 0][q][d] = Yi[q][d];
                    if (i < S - 1) mlp_eval<T, D, H, PHI, TPT>(sW, sB2, sTab, w.slot, Yi, K[i]);
                }
            }
            // the stage loop is deliberately NOT unrolled (ls becomes a small local array, touched a few times per stage):
            // the loop body stays resident in the instruction cache; the chunk loop IS unrolled so that the
            // parameter-gradient accumulators stay in registers
#if PNODE_ADJ_ROLL
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int i = S - 1; i >= 0; --i) {
                if (tab.fsal && i == S - 1) {
#pragma unroll
                    for (int q = 0; q < TPT; ++q)
#pragma unroll
                        for (int d = 0; d < D; ++d) ls[i][q][d] = T(0);
                    continue;
                }
                // cotangent of the stage slope, pre-multiplied by the step coefficient: v = c * w   
                T v[TPT][D];
                const double bi = tab.b[i];
                const bool has_b = bi != 0.0;
                const T cstep = (T)(has_b ? h * bi : h);
#pragma unroll
                for (int q = 0; q < TPT; ++q) {
#pragma unroll
                    for (int d = 0; d < D; ++d) v[q][d] = has_b ? lam[q][d] : T(0);
                }
                for (int j = i + 1; j < S; ++j) {
                    const T r = (T)(has_b ? tab.a[j][i] / bi : tab.a[j][i]);
#pragma unroll
                    for (int q = 0; q < TPT; ++q)
#pragma unroll
                        for (int d = 0; d < D; ++d) v[q][d] = fma(r, ls[j][q][d], v[q][d]);
                }
                T Y[TPT][D], x[TPT][D], dx[TPT][D];
#pragma unroll
                for (int q = 0; q < TPT; ++q) {
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        Y[q][d] = SO ? Yst[SO ? i : 0][q][d] : Ynext[q][d];
                        v[q][d] = valid[q] ? v[q][d] * cstep : T(0);
                        dx[q][d] = T(0);
                    }
                    apply_phi<T, D, PHI>(Y[q], x[q]);
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        tile.V[d][q * 32 + lane] = v[q][d];
                        tile.X[d][q * 32 + lane] = x[q][d];
                        accB2[d] += (double)v[q][d];
                    }
                }
                if (!SO) {
                    if (i > 0) fetch_Y(n, i - 1); else fetch_Y(n - 1, s_top);
                }
                if constexpr (AdjMerge<T, D, H>::value) {
                    // ---- phase 1 (lane = trajectory): VJP through every hidden unit of the stage, tanh parked per unit ----
#pragma unroll
                    for (int c = 0; c < NCHUNK; ++c) {
                        const int j0 = c * JH;
                        const int jn = (H - j0 < JH) ? (H - j0) : JH;
#pragma unroll 1
                        for (int jb = 0; jb < jn; jb += G) {
                            Unit<T, D> u[G];
#pragma unroll
                            for (int g = 0; g < G; ++g)
                                u[g] = get_unit<T, D, H>(sW, w.slot, j0 + ((jb + g < jn) ? jb + g : jn - 1));
                            T z[G], a[G];
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                z[g] = u[g].b1;
#pragma unroll
                                for (int d = 0; d < D; ++d) z[g] = fma(u[g].w1[d], x[0][d], z[g]);
                            }
                            tanh_group<G>(z, a, sTab);
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                if (jb + g < jn) {
                                    T gg = T(0);
#pragma unroll
                                    for (int d = 0; d < D; ++d) gg = fma(u[g].w2[d], v[0][d], gg);
                                    const T s = gg * fma(-a[g], a[g], T(1));
#pragma unroll
                                    for (int d = 0; d < D; ++d) dx[0][d] = fma(s, u[g].w1[d], dx[0][d]);
                                    tile.A[(j0 + jb + g) * PITCH + lane] = a[g];
                                }
                            }
                        }
                    }
                    __syncwarp();
                    // ---- phase 2 (lane = one hidden unit of every chunk): outer products over the warp's trajectories -----
                    if (lane < JH) {
                        T pW2[NCHUNK][D], pW1[NCHUNK][D], pB1[NCHUNK];
                        Unit<T, D> own[NCHUNK];
                        bool act[NCHUNK];
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c) {
                            const int jn = (H - c * JH < JH) ? (H - c * JH) : JH;
                            act[c] = lane < jn;
                            own[c] = get_unit<T, D, H>(sW, w.slot, c * JH + (act[c] ? lane : 0));
                            pB1[c] = T(0);
#pragma unroll
                            for (int d = 0; d < D; ++d) pW2[c][d] = pW1[c][d] = T(0);
                        }
#pragma unroll 2
                        for (int k = 0; k < NK; k += VEC) {
                            T vv[D][VEC], xx[D][VEC];
#pragma unroll
                            for (int d = 0; d < D; ++d) {
                                lds16(&tile.V[d][k], vv[d]);
                                lds16(&tile.X[d][k], xx[d]);
                            }
#pragma unroll
                            for (int c = 0; c < NCHUNK; ++c) {
                                if (act[c]) {
                                    T a[VEC];
                                    lds16(&tile.A[(c * JH + lane) * PITCH + k], a);
#pragma unroll
                                    for (int e = 0; e < VEC; ++e) {  // phase 1's three instructions, operand for operand
                                        T gg = T(0);
#pragma unroll
                                        for (int d = 0; d < D; ++d) gg = fma(own[c].w2[d], vv[d][e], gg);
                                        const T s = gg * fma(-a[e], a[e], T(1));
                                        pB1[c] += s;
#pragma unroll
                                        for (int d = 0; d < D; ++d) {
                                            pW2[c][d] = fma(vv[d][e], a[e], pW2[c][d]);
                                            pW1[c][d] = fma(s, xx[d][e], pW1[c][d]);
                                        }
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c) {
                            if (act[c]) {
                                accB1[c] += (double)pB1[c];
#pragma unroll
                                for (int d = 0; d < D; ++d) {
                                    accW2[c][d] += (double)pW2[c][d];
                                    accW1[c][d] += (double)pW1[c][d];
                                }
                            }
                        }
                    }
                    __syncwarp();
                } else {
#pragma unroll
                for (int c = 0; c < NCHUNK; ++c) {
                    const int j0 = c * JH;
                    const int jn = (H - j0 < JH) ? (H - j0) : JH;
                    // ---- phase 1 (lane = trajectory): VJP through every hidden unit of the chunk ---------------------
#pragma unroll 1
                    for (int jb = 0; jb < jn; jb += G) {
                        Unit<T, D> u[G];
#pragma unroll
                        for (int g = 0; g < G; ++g)
                            u[g] = get_unit<T, D, H>(sW, w.slot, j0 + ((jb + g < jn) ? jb + g : jn - 1));
#pragma unroll
                        for (int q = 0; q < TPT; ++q) {
                            T z[G], a[G];
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                z[g] = u[g].b1;
#pragma unroll
                                for (int d = 0; d < D; ++d) z[g] = fma(u[g].w1[d], x[q][d], z[g]);
                            }
                            tanh_group<G>(z, a, sTab);
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                if (jb + g < jn) {
                                    T gg = T(0);
#pragma unroll
                                    for (int d = 0; d < D; ++d) gg = fma(u[g].w2[d], v[q][d], gg);
                                    const T s = gg * fma(-a[g], a[g], T(1));
#pragma unroll
                                    for (int d = 0; d < D; ++d) dx[q][d] = fma(s, u[g].w1[d], dx[q][d]);
                                    tile.A[(jb + g) * PITCH + q * 32 + lane] = a[g];
                                    if (!AdjRecompute<T>::value) tile.Sg[(jb + g) * PITCH + q * 32 + lane] = s;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    // ---- phase 2 (lane = hidden unit): outer products summed over the warp's trajectories -------------
                    if (lane < jn) {
                        T pW2[D], pW1[D], pB1 = T(0);
#pragma unroll
                        for (int d = 0; d < D; ++d) pW2[d] = pW1[d] = T(0);
                        const Unit<T, D> own = get_unit<T, D, H>(sW, w.slot, j0 + lane);  // used when s is recomputed
#pragma unroll 4
                        for (int k = 0; k < NK; k += VEC) {
                            T a[VEC], s[VEC], vv[D][VEC], xx[D][VEC];
                            lds16(&tile.A[lane * PITCH + k], a);
                            if (!AdjRecompute<T>::value) lds16(&tile.Sg[lane * PITCH + k], s);
#pragma unroll
                            for (int d = 0; d < D; ++d) {
                                lds16(&tile.V[d][k], vv[d]);
                                lds16(&tile.X[d][k], xx[d]);
                            }
                            if (AdjRecompute<T>::value) {
#pragma unroll
                                for (int e = 0; e < VEC; ++e) {  // phase 1's three instructions, operand for operand
                                    T gg = T(0);
#pragma unroll
                                    for (int d = 0; d < D; ++d) gg = fma(own.w2[d], vv[d][e], gg);
                                    s[e] = gg * fma(-a[e], a[e], T(1));
                                }
                            }
#pragma unroll
                            for (int e = 0; e < VEC; ++e) {
                                pB1 += s[e];
#pragma unroll
                                for (int d = 0; d < D; ++d) {
                                    pW2[d] = fma(vv[d][e], a[e], pW2[d]);
                                    pW1[d] = fma(s[e], xx[d][e], pW1[d]);
                                }
                            }
                        }
                        accB1[c] += (double)pB1;
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            accW2[c][d] += (double)pW2[d];
                            accW1[c][d] += (double)pW1[d];
                        }
                    }
                    __syncwarp();
                }
                }
#pragma unroll
                for (int q = 0; q < TPT; ++q)
#pragma unroll
                    for (int d = 0; d < D; ++d)
                        ls[i][q][d] = (PHI == 1) ? dx[q][d] * (T(3) * Y[q][d] * Y[q][d]) : dx[q][d];
            }
#pragma unroll
            for (int q = 0; q < TPT; ++q) {
#pragma unroll
                for (int i = 0; i < S; ++i)
#pragma unroll
                    for (int d = 0; d < D; ++d) lam[q][d] += ls[i][q][d];
                if (in_slot >= 0 && valid[q]) {
#pragma unroll
                    for (int d = 0; d < D; ++d) lam[q][d] += gout[((int64_t)in_slot * ntraj + traj[q]) * D + d];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < TPT; ++q)
            if (valid[q]) {
#pragma unroll
                for (int d = 0; d < D; ++d) lambda_out[traj[q] * D + d] = lam[q][d];
            }
    }

    // ---- block-level combine of the per-lane accumulators (fixed order), then grid-level by the last block ----------
    __syncthreads();
    double *blk = reinterpret_cast<double *>(tiles);  // reuse tile storage: [ADJ_WARPS][NP]
    static_assert(sizeof(WarpTile<T, D, H>) * ADJ_WARPS >= sizeof(double) * ADJ_WARPS * NP, "tile storage too small");
#pragma unroll
    for (int d = 0; d < D; ++d) accB2[d] = warp_sum(accB2[d]);
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int j = c * JH + lane;
        if (lane < JH && j < H) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                blk[warp * NP + j * D + d] = accW1[c][d];
                blk[warp * NP + H * D + H + d * H + j] = accW2[c][d];
            }
            blk[warp * NP + H * D + j] = accB1[c];
        }
    }
    if (lane < D) blk[warp * NP + 2 * H * D + H + lane] = accB2[lane];
    adj_finish<T, NP>(blk, ADJ_WARPS, work, mu_out, pc, &is_last);
}

// ---------------------------------------------------------------------------------------------------------------------
// small batches (the reference's own spiral run is batch 20)
//
// With one trajectory per thread a batch of 20 is ONE warp that issues the 50 hidden units of every stage evaluation one
// after the other: 70 us forward + 148 us adjoint of pure FP64-issue latency on one scheduler, the rest of the GPU idle
// (profiles: tools/prof_pass.py --config 1).  Below SMALL_MAX_TRAJ trajectories a slot (one trajectory) is spread over
// SMALL_LPT = 10 adjacent lanes, five hidden units each (one tanh group); the per-unit sums -- the slope, the state
// cotangent -- are added up in lane order by every lane of the slot (bit-identical copies, so the ten lanes carry the same
// state), lane 0 of the slot stores.  Three slots per warp, twelve per CTA: batch 20 becomes two CTAs of fully used warps.
constexpr int SMALL_LPT = 10;
constexpr int SMALL_SPW = 32 / SMALL_LPT;             // slots per warp
constexpr int SMALL_WARPS = 4;
constexpr int SMALL_THREADS = SMALL_WARPS * 32;
constexpr int SMALL_SLOTS = SMALL_WARPS * SMALL_SPW;  // slots per CTA
constexpr int64_t SMALL_MAX_TRAJ = 4096;

template <typename T>
__device__ __forceinline__ T slot_sum(T v, int base) {  // sum over the SMALL_LPT lanes base .. base + 9, fixed order
    T t = __shfl_sync(0xffffffffu, v, base);
#pragma unroll
    for (int k = 1; k < SMALL_LPT; ++k) t += __shfl_sync(0xffffffffu, v, base + k);
    return t;
}

template <typename T, int D, int H, int S, int PHI>
__global__ void __launch_bounds__(SMALL_THREADS)
mlp_rk_fwd_small_kernel(const MlpPtrs<T> w, const pnode_rk_tableau tab, const T *__restrict__ u0, const int64_t ntraj,
                        const pnode_step *__restrict__ sched, const int nsteps, T *__restrict__ sol, T *__restrict__ ckpt) {
    static_assert(H % SMALL_LPT == 0, "hidden units must split evenly over the lanes of a slot");
    constexpr int UPL = H / SMALL_LPT;
    __shared__ Unit<T, D> sW[H];
    __shared__ T sB2[D];
    __shared__ T sTab[EXP_TAB];
    load_weights<T, D, H>(sW, sB2, sTab, w);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sl = lane / SMALL_LPT, sub = lane % SMALL_LPT, base = sl * SMALL_LPT;
    const int64_t nt = (ntraj + SMALL_SLOTS - 1) / SMALL_SLOTS;
    for (int64_t tidx = blockIdx.x; tidx < nt; tidx += gridDim.x) {
        const int64_t traj = tidx * SMALL_SLOTS + warp * SMALL_SPW + sl;
        const bool valid = sl < SMALL_SPW && traj < ntraj;
        const bool writer = valid && sub == 0;
        T y[1][D], K[S][1][D];
#pragma unroll
        for (int d = 0; d < D; ++d) y[0][d] = valid ? u0[traj * D + d] : T(0);
        for (int n = 0; n < nsteps; ++n) {
            const double h = sched[n].h;
            const int out_slot = sched[n].out_slot;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                T Y[1][D];
#pragma unroll
                for (int d = 0; d < D; ++d) Y[0][d] = y[0][d];
#pragma unroll
                for (int j = 0; j < i; ++j) {
                    const T ha = (T)(h * tab.a[i][j]);
#pragma unroll
                    for (int d = 0; d < D; ++d) Y[0][d] = fma(ha, K[j][0][d], Y[0][d]);
                }
                if (ckpt != nullptr && writer) {
#pragma unroll
                    for (int d = 0; d < D; ++d) ckpt[(((int64_t)n * S + i) * D + d) * ntraj + traj] = Y[0][d];
                }
                if (i == 0 && tab.fsal && n > 0) {
#pragma unroll
                    for (int d = 0; d < D; ++d) K[0][0][d] = K[S - 1][0][d];
                } else {
                    T x[1][D], part[1][D];
                    apply_phi<T, D, PHI>(Y[0], x[0]);
#pragma unroll
                    for (int d = 0; d < D; ++d) part[0][d] = sub == 0 ? get_b2<T, D, H>(sB2, w.slot, d) : T(0);
                    mlp_eval_units<T, D, H, PHI, 1, UPL>(sW, sTab, w.slot, sub * UPL, x, part);
#pragma unroll
                    for (int d = 0; d < D; ++d) K[i][0][d] = slot_sum(part[0][d], base);
                }
            }
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const T hb = (T)(h * tab.b[j]);
#pragma unroll
                for (int d = 0; d < D; ++d) y[0][d] = fma(hb, K[j][0][d], y[0][d]);
            }
            if (out_slot >= 0 && writer) {
#pragma unroll
                for (int d = 0; d < D; ++d) sol[((int64_t)out_slot * ntraj + traj) * D + d] = y[0][d];
            }
        }
    }
}

// tile of one warp of the small-batch adjoint: all H units x (slots of the warp, padded to a 16-byte vector)
template <typename T, int D, int H>
struct alignas(16) SmallTile {
    static constexpr int VEC = 16 / sizeof(T);
    static constexpr int NKP = (SMALL_SPW + VEC - 1) / VEC * VEC;
    T A[H][NKP], Sg[H][NKP], V[D][NKP], X[D][NKP];
};

template <typename T, int D, int H, int S, int PHI>
__global__ void __launch_bounds__(SMALL_THREADS)
mlp_rk_adj_small_kernel(const MlpPtrs<T> w, const pnode_rk_tableau tab, const int64_t ntraj,
                        const pnode_step *__restrict__ sched, const int nsteps, const int last_slot,
                        const T *__restrict__ gout, const T *__restrict__ ckpt, T *__restrict__ lambda_out,
                        T *__restrict__ mu_out, AdjWork *__restrict__ work, const PeerComm pc) {
    typedef SmallTile<T, D, H> Tile;
    constexpr int UPL = H / SMALL_LPT, NKP = Tile::NKP, VEC = Tile::VEC, NP = 2 * H * D + H + D;
    constexpr int NU = (H + 31) / 32;  // hidden units per lane in phase 2 (lane, lane + 32)
    __shared__ Unit<T, D> sW[H];
    __shared__ T sB2[D];
    __shared__ T sTab[EXP_TAB];
    __shared__ Tile tiles[SMALL_WARPS];
    __shared__ double blk[SMALL_WARPS * NP];
    __shared__ bool is_last;
    __shared__ T sR[S][S];  // a_ji / b_i (a_ji where b_i = 0): the stage recurrences' coefficients, once per launch
    load_weights<T, D, H>(sW, sB2, sTab, w);
    if (threadIdx.x < S * S) {
        const int j = threadIdx.x / S, i = threadIdx.x % S;
        const double bi = tab.b[i];
        sR[j][i] = (T)(bi != 0.0 ? tab.a[j][i] / bi : tab.a[j][i]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Tile &tile = tiles[warp];
    for (int i = lane; i < (int)(sizeof(Tile) / sizeof(T)); i += 32) reinterpret_cast<T *>(&tile)[i] = T(0);  // pad columns
    __syncthreads();
    const int sl = lane / SMALL_LPT, sub = lane % SMALL_LPT, base = sl * SMALL_LPT;
    const int col = sl < SMALL_SPW ? sl : 0;

    double accW1[NU][D], accW2[NU][D], accB1[NU], accB2[D];
#pragma unroll
    for (int c = 0; c < NU; ++c) {
        accB1[c] = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) accW1[c][d] = accW2[c][d] = 0.0;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) accB2[d] = 0.0;

    const int64_t nt = (ntraj + SMALL_SLOTS - 1) / SMALL_SLOTS;
    for (int64_t tidx = blockIdx.x; tidx < nt; tidx += gridDim.x) {
        const int64_t traj = tidx * SMALL_SLOTS + warp * SMALL_SPW + sl;
        const bool valid = sl < SMALL_SPW && traj < ntraj;
        const bool writer = valid && sub == 0;
        T lam[D];
#pragma unroll
        for (int d = 0; d < D; ++d) lam[d] = valid ? gout[((int64_t)last_slot * ntraj + traj) * D + d] : T(0);
        // stage values are fetched one stage ahead of their use
        const int s_top = (tab.fsal ? S - 2 : S - 1);
        T Ynext[D];
        auto fetch_Y = [&](int nn, int ii) {
#pragma unroll
            for (int d = 0; d < D; ++d)
                Ynext[d] = (valid && nn >= 0) ? ckpt[(((int64_t)nn * S + ii) * D + d) * ntraj + traj] : T(0);
        };
        fetch_Y(nsteps - 1, s_top);
        for (int n = nsteps - 1; n >= 0; --n) {
            const double h = sched[n].h;
            const int in_slot = sched[n].in_slot;
            T ls[S][D];
#pragma unroll
            for (int i = S - 1; i >= 0; --i) {
                if (tab.fsal && i == S - 1) {
#pragma unroll
                    for (int d = 0; d < D; ++d) ls[i][d] = T(0);
                    continue;
                }
                T v[D];
                const double bi = tab.b[i];
                const bool has_b = bi != 0.0;
                const T cstep = (T)(has_b ? h * bi : h);
#pragma unroll
                for (int d = 0; d < D; ++d) v[d] = has_b ? lam[d] : T(0);
#pragma unroll
                for (int j = i + 1; j < S; ++j) {
                    const T r = sR[j][i];
#pragma unroll
                    for (int d = 0; d < D; ++d) v[d] = fma(r, ls[j][d], v[d]);
                }
                T Y[D], x[D], dx[D];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    Y[d] = Ynext[d];
                    v[d] = valid ? v[d] * cstep : T(0);
                    dx[d] = T(0);
                }
                if (i > 0) fetch_Y(n, i - 1); else fetch_Y(n - 1, s_top);
                apply_phi<T, D, PHI>(Y, x);
                if (sub == 0 && sl < SMALL_SPW) {
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        tile.V[d][col] = v[d];
                        tile.X[d][col] = x[d];
                        accB2[d] += (double)v[d];
                    }
                }
                // ---- phase 1 (lane = one tenth of a trajectory's hidden units) ------------------------------------------
                {
                    Unit<T, D> u[UPL];
                    T z[UPL], a[UPL];
#pragma unroll
                    for (int g = 0; g < UPL; ++g) {
                        u[g] = get_unit<T, D, H>(sW, w.slot, sub * UPL + g);
                        z[g] = u[g].b1;
#pragma unroll
                        for (int d = 0; d < D; ++d) z[g] = fma(u[g].w1[d], x[d], z[g]);
                    }
                    tanh_group<UPL>(z, a, sTab);
#pragma unroll
                    for (int g = 0; g < UPL; ++g) {
                        T gg = T(0);
#pragma unroll
                        for (int d = 0; d < D; ++d) gg = fma(u[g].w2[d], v[d], gg);
                        const T sg = gg * fma(-a[g], a[g], T(1));
#pragma unroll
                        for (int d = 0; d < D; ++d) dx[d] = fma(sg, u[g].w1[d], dx[d]);
                        if (sl < SMALL_SPW) {
                            tile.A[sub * UPL + g][col] = a[g];
                            tile.Sg[sub * UPL + g][col] = sg;
                        }
                    }
                }
#pragma unroll
                for (int d = 0; d < D; ++d) dx[d] = slot_sum(dx[d], base);
                __syncwarp();
                // ---- phase 2 (lane = hidden unit): outer products summed over the warp's trajectories -------------------
#pragma unroll
                for (int c = 0; c < NU; ++c) {
                    const int j = c * 32 + lane;
                    if (j < H) {
                        T pW2[D], pW1[D], pB1 = T(0);
#pragma unroll
                        for (int d = 0; d < D; ++d) pW2[d] = pW1[d] = T(0);
#pragma unroll
                        for (int k = 0; k < NKP; k += VEC) {
                            T a[VEC], s[VEC], vv[D][VEC], xx[D][VEC];
                            lds16(&tile.A[j][k], a);
                            lds16(&tile.Sg[j][k], s);
#pragma unroll
                            for (int d = 0; d < D; ++d) {
                                lds16(&tile.V[d][k], vv[d]);
                                lds16(&tile.X[d][k], xx[d]);
                            }
#pragma unroll
                            for (int e = 0; e < VEC; ++e) {
                                pB1 += s[e];
#pragma unroll
                                for (int d = 0; d < D; ++d) {
                                    pW2[d] = fma(vv[d][e], a[e], pW2[d]);
                                    pW1[d] = fma(s[e], xx[d][e], pW1[d]);
                                }
                            }
                        }
                        accB1[c] += (double)pB1;
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            accW2[c][d] += (double)pW2[d];
                            accW1[c][d] += (double)pW1[d];
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int d = 0; d < D; ++d) ls[i][d] = (PHI == 1) ? dx[d] * (T(3) * Y[d] * Y[d]) : dx[d];
            }
#pragma unroll
            for (int i = 0; i < S; ++i)
#pragma unroll
                for (int d = 0; d < D; ++d) lam[d] += ls[i][d];
            if (in_slot >= 0 && valid) {
#pragma unroll
                for (int d = 0; d < D; ++d) lam[d] += gout[((int64_t)in_slot * ntraj + traj) * D + d];
            }
        }
        if (writer) {
#pragma unroll
            for (int d = 0; d < D; ++d) lambda_out[traj * D + d] = lam[d];
        }
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < D; ++d) accB2[d] = warp_sum(accB2[d]);
#pragma unroll
    for (int c = 0; c < NU; ++c) {
        const int j = c * 32 + lane;
        if (j < H) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                blk[warp * NP + j * D + d] = accW1[c][d];
                blk[warp * NP + H * D + H + d * H + j] = accW2[c][d];
            }
            blk[warp * NP + H * D + j] = accB1[c];
        }
    }
    if (lane < D) blk[warp * NP + 2 * H * D + H + lane] = accB2[lane];
    adj_finish<T, NP>(blk, SMALL_WARPS, work, mu_out, pc, &is_last);
}

// block partial (fixed order over the warps) -> grid partials -> the last block sums them in block order (and, in sharded
// runs, all-reduces the result over the peers' inboxes) -- shared by both adjoint kernels
template <typename T, int NP>
__device__ void adj_finish(double *blk, int nwarps, AdjWork *__restrict__ work, T *__restrict__ mu_out, const PeerComm &pc,
                           bool *is_last) {
    __syncthreads();
    for (int p = threadIdx.x; p < NP; p += blockDim.x) {
        double s = 0.0;
        for (int wi = 0; wi < nwarps; ++wi) s += blk[wi * NP + p];
        work->partial[(int64_t)blockIdx.x * NP + p] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(&work->ticket, 1u);
        *is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (*is_last) {
        __threadfence();
        const bool dp = pc.peer_bufs != nullptr && pc.world > 1;
        for (int p = threadIdx.x; p < NP; p += blockDim.x) {
            double s = 0.0;
            for (int b = 0; b < (int)gridDim.x; ++b) s += ((volatile double *)work->partial)[(int64_t)b * NP + p];
            if (dp)
                blk[p] = s;  // this GPU's mu; the cross-GPU sum follows in the same kernel
            else
                mu_out[p] = (T)s;
        }
        if (threadIdx.x == 0) work->ticket = 0u;
        if (dp) {
            __syncthreads();
            peer_allreduce_and_store<T>(blk, NP, pc, mu_out);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host dispatch

template <typename T, int D, int H>
static size_t adj_smem_bytes() {
    return ((sizeof(Unit<T, D>) * H + sizeof(T) * (D + EXP_TAB) + 15) / 16) * 16 + sizeof(WarpTile<T, D, H>) * ADJ_WARPS;
}

static int g_slot_counter = 0;

// the small-batch kernels (ten lanes per trajectory) exist for the spiral shape only; other shapes run the large-batch
// kernels at every batch size
template <int D, int H>
struct HasSmall {
    static constexpr bool value = (D == 2 && H == 50);
};

static bool small_batch_enabled() {
    static const int on = [] {
        const char *e = getenv("PNODE_MLP_SMALL");
        return e ? atoi(e) : 1;
    }();
    return on != 0;
}

template <typename T>
static int upload_weights(const pnode_mlp_desc *m, int *slot_out, cudaStream_t st) {
    const int slot = (g_slot_counter++) % W_SLOTS;
    *slot_out = slot;
#if PNODE_CONST_WEIGHTS
    const int D = m->dim, H = m->hidden;
    const size_t base = (size_t)slot * W_MAXP * sizeof(T);
    const void *sym = sizeof(T) == 4 ? (const void *)cWf : (const void *)cWd;
    const void *src[4] = {m->d_w1, m->d_b1, m->d_w2, m->d_b2};
    const size_t cnt[4] = {(size_t)H * D, (size_t)H, (size_t)D * H, (size_t)D};
    size_t off = 0;
    for (int i = 0; i < 4; ++i) {
        PNODE_CUDA_OK(cudaMemcpyToSymbolAsync(sym, src[i], cnt[i] * sizeof(T), base + off * sizeof(T),
                                              cudaMemcpyDeviceToDevice, st));
        off += cnt[i];
    }
#endif
    return 0;
}

template <typename T, int D, int H, int S, int PHI>
static int launch_fwd(const pnode_mlp_desc *m, const pnode_rk_tableau *tab, const void *d_u0, int64_t ntraj,
                      const pnode_step *d_sched, int nsteps, void *d_sol, void *d_ckpt, int so, cudaStream_t st) {
    int slot = 0;
    if (int rc = upload_weights<T>(m, &slot, st)) return rc;
    MlpPtrs<T> w{static_cast<const T *>(m->d_w1), static_cast<const T *>(m->d_b1), static_cast<const T *>(m->d_w2),
                 static_cast<const T *>(m->d_b2), slot};
    if constexpr (HasSmall<D, H>::value) {
        if (ntraj <= SMALL_MAX_TRAJ && small_batch_enabled() && !so) {  // bounded storage is for large batches
            const int64_t want = (ntraj + SMALL_SLOTS - 1) / SMALL_SLOTS;
            const int grid = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
            mlp_rk_fwd_small_kernel<T, D, H, S, PHI><<<grid, SMALL_THREADS, 0, st>>>(
                w, *tab, static_cast<const T *>(d_u0), ntraj, d_sched, nsteps, static_cast<T *>(d_sol),
                static_cast<T *>(d_ckpt));
            PNODE_CUDA_OK(cudaGetLastError());
            return 0;
        }
    }
    auto kern = mlp_rk_fwd_kernel<T, D, H, S, PHI>;
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        PNODE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, FWD_THREADS, 0));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    const int64_t tile = (int64_t)FWD_THREADS * Cfg<T>::TPT;
    int64_t want = (ntraj + tile - 1) / tile;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, FWD_THREADS, 0, st>>>(w, *tab, static_cast<const T *>(d_u0), ntraj, d_sched, nsteps,
                                       static_cast<T *>(d_sol), static_cast<T *>(d_ckpt), so);
    PNODE_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T, int D, int H, int S, int PHI, bool SO = false>
static int launch_adj(const pnode_mlp_desc *m, const pnode_rk_tableau *tab, int64_t ntraj, const pnode_step *d_sched,
                      int nsteps, int last_slot, const void *d_gout, const void *d_ckpt, void *d_lambda, void *d_mu,
                      void *d_work, const PeerComm &pc, cudaStream_t st) {
    int slot = 0;
    if (int rc = upload_weights<T>(m, &slot, st)) return rc;
    MlpPtrs<T> w{static_cast<const T *>(m->d_w1), static_cast<const T *>(m->d_b1), static_cast<const T *>(m->d_w2),
                 static_cast<const T *>(m->d_b2), slot};
    if constexpr (HasSmall<D, H>::value) {
        if (ntraj <= SMALL_MAX_TRAJ && small_batch_enabled() && !SO) {
            const int64_t want = (ntraj + SMALL_SLOTS - 1) / SMALL_SLOTS;
            int64_t cap = (int64_t)sm_count() * 4;
            if (cap > ADJ_MAX_BLOCKS) cap = ADJ_MAX_BLOCKS;
            const int grid = (int)(want < cap ? want : cap);
            mlp_rk_adj_small_kernel<T, D, H, S, PHI><<<grid, SMALL_THREADS, 0, st>>>(
                w, *tab, ntraj, d_sched, nsteps, last_slot, static_cast<const T *>(d_gout),
                static_cast<const T *>(d_ckpt), static_cast<T *>(d_lambda), static_cast<T *>(d_mu),
                static_cast<AdjWork *>(d_work), pc);
            PNODE_CUDA_OK(cudaGetLastError());
            return 0;
        }
    }
    auto kern = mlp_rk_adj_kernel<T, D, H, S, PHI, SO>;
    const size_t smem = adj_smem_bytes<T, D, H>();
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        PNODE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PNODE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, ADJ_THREADS, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    const int64_t tile = (int64_t)ADJ_THREADS * Cfg<T>::TPT;
    int64_t want = (ntraj + tile - 1) / tile;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (cap > ADJ_MAX_BLOCKS) cap = ADJ_MAX_BLOCKS;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, ADJ_THREADS, smem, st>>>(w, *tab, ntraj, d_sched, nsteps, last_slot, static_cast<const T *>(d_gout),
                                          static_cast<const T *>(d_ckpt), static_cast<T *>(d_lambda),
                                          static_cast<T *>(d_mu), static_cast<AdjWork *>(d_work), pc);
    PNODE_CUDA_OK(cudaGetLastError());
    return 0;
}

// The compiled instantiations.  (dim, hidden) = (2, 50) is the spiral model; the other shapes widen the recogniser to
// 1- to 4-dimensional states and up to 100 hidden units (the host zero-pads narrower layers to the next compiled width --
// a padded unit contributes fma(0, tanh(0), acc) = acc, so results are unchanged bit for bit).  Every shape is compiled for
// every explicit tableau PETSc's TSRK offers through pnode's method= table (1fe, 2a/2b, 3, 3bs/4, 5dp) and phi in
// {identity, cube}.  The file is compiled once per PART, in parallel (pnode_b200/build.py): part 0 holds the C entry points.
#define PNODE_FOR_STAGES(X) X(1) X(2) X(3) X(4) X(7)
#ifndef PNODE_MLP_PART
#define PNODE_MLP_PART 0
#endif
#if PNODE_MLP_PART == 0
#define PNODE_FOR_SHAPES(X) X(2, 50)
#define PNODE_PART_NAME(f) f##_part0
#elif PNODE_MLP_PART == 1
#define PNODE_FOR_SHAPES(X) X(2, 100)
#define PNODE_PART_NAME(f) f##_part1
#elif PNODE_MLP_PART == 2
#define PNODE_FOR_SHAPES(X) X(3, 50)
#define PNODE_PART_NAME(f) f##_part2
#elif PNODE_MLP_PART == 3
#define PNODE_FOR_SHAPES(X) X(4, 50) X(1, 50)
#define PNODE_PART_NAME(f) f##_part3
#endif
constexpr int MLP_PARTS = 4;
constexpr int NO_KERNEL_HERE = -1000;  // this part holds no kernel for the shape: the caller tries the next part

static bool shape_ok(int dim, int hidden, int phi, int stages) {
    bool s_ok = stages == 1 || stages == 2 || stages == 3 || stages == 4 || stages == 7;
    bool dh_ok = (hidden == 50 && dim >= 1 && dim <= 4) || (hidden == 100 && dim == 2);
    return dh_ok && (phi == 0 || phi == 1) && s_ok;
}

struct FwdArgs {
    const pnode_mlp_desc *m;
    const pnode_rk_tableau *tab;
    const void *d_u0;
    int64_t ntraj;
    const pnode_step *d_sched;
    int nsteps;
    void *d_sol, *d_ckpt;
    int so;
    cudaStream_t st;
};
struct AdjArgs {
    const pnode_mlp_desc *m;
    const pnode_rk_tableau *tab;
    int64_t ntraj;
    const pnode_step *d_sched;
    int nsteps, last_slot;
    const void *d_gout, *d_ckpt;
    void *d_lambda, *d_mu, *d_work;
    PeerComm pc;
    bool so;
    cudaStream_t st;
};

template <typename T, int D, int H>
static int dispatch_fwd(const FwdArgs &a) {
#define X(SS)                                                                                                      \
    if (a.tab->s == SS) {                                                                                          \
        if (a.m->phi == 1)                                                                                         \
            return launch_fwd<T, D, H, SS, 1>(a.m, a.tab, a.d_u0, a.ntraj, a.d_sched, a.nsteps, a.d_sol, a.d_ckpt, \
                                              a.so, a.st);                                                         \
        return launch_fwd<T, D, H, SS, 0>(a.m, a.tab, a.d_u0, a.ntraj, a.d_sched, a.nsteps, a.d_sol, a.d_ckpt,     \
                                          a.so, a.st);                                                             \
    }
    PNODE_FOR_STAGES(X)
#undef X
    PNODE_REQUIRE(false, "pnode_mlp_rk_forward: no kernel for %d stages", a.tab->s);
}

template <typename T, int D, int H, bool SO>
static int dispatch_adj(const AdjArgs &a) {
#define X(SS)                                                                                                      \
    if (a.tab->s == SS) {                                                                                          \
        if (a.m->phi == 1)                                                                                         \
            return launch_adj<T, D, H, SS, 1, SO>(a.m, a.tab, a.ntraj, a.d_sched, a.nsteps, a.last_slot, a.d_gout, \
                                                  a.d_ckpt, a.d_lambda, a.d_mu, a.d_work, a.pc, a.st);             \
        return launch_adj<T, D, H, SS, 0, SO>(a.m, a.tab, a.ntraj, a.d_sched, a.nsteps, a.last_slot, a.d_gout,     \
                                              a.d_ckpt, a.d_lambda, a.d_mu, a.d_work, a.pc, a.st);                 \
    }
    PNODE_FOR_STAGES(X)
#undef X
    PNODE_REQUIRE(false, "pnode_mlp_rk_adjoint: no kernel for %d stages", a.tab->s);
}

int PNODE_PART_NAME(mlp_rk_fwd)(const FwdArgs &a) {
#define X(DD, HH)                                                                 \
    if (a.m->dim == DD && a.m->hidden == HH) {                                    \
        if (a.m->dtype == PNODE_F32) return dispatch_fwd<float, DD, HH>(a);       \
        if (a.m->dtype == PNODE_F64) return dispatch_fwd<double, DD, HH>(a);      \
    }
    PNODE_FOR_SHAPES(X)
#undef X
    return NO_KERNEL_HERE;
}

int PNODE_PART_NAME(mlp_rk_adj)(const AdjArgs &a) {
#define X(DD, HH)                                                                                                   \
    if (a.m->dim == DD && a.m->hidden == HH) {                                                                      \
        if (a.m->dtype == PNODE_F32) return a.so ? dispatch_adj<float, DD, HH, true>(a) : dispatch_adj<float, DD, HH, false>(a);   \
        if (a.m->dtype == PNODE_F64) return a.so ? dispatch_adj<double, DD, HH, true>(a) : dispatch_adj<double, DD, HH, false>(a); \
    }
    PNODE_FOR_SHAPES(X)
#undef X
    return NO_KERNEL_HERE;
}

#if PNODE_MLP_PART == 0
int mlp_rk_fwd_part1(const FwdArgs &a);
int mlp_rk_fwd_part2(const FwdArgs &a);
int mlp_rk_fwd_part3(const FwdArgs &a);
int mlp_rk_adj_part1(const AdjArgs &a);
int mlp_rk_adj_part2(const AdjArgs &a);
int mlp_rk_adj_part3(const AdjArgs &a);

static int mlp_rk_fwd_any_part(const FwdArgs &a) {
    int (*const parts[MLP_PARTS])(const FwdArgs &) = {mlp_rk_fwd_part0, mlp_rk_fwd_part1, mlp_rk_fwd_part2, mlp_rk_fwd_part3};
    for (int p = 0; p < MLP_PARTS; ++p) {
        const int rc = parts[p](a);
        if (rc != NO_KERNEL_HERE) return rc;
    }
    PNODE_REQUIRE(false, "pnode_mlp_rk_forward: unsupported dtype %d", a.m->dtype);
}
static int mlp_rk_adj_any_part(const AdjArgs &a) {
    int (*const parts[MLP_PARTS])(const AdjArgs &) = {mlp_rk_adj_part0, mlp_rk_adj_part1, mlp_rk_adj_part2, mlp_rk_adj_part3};
    for (int p = 0; p < MLP_PARTS; ++p) {
        const int rc = parts[p](a);
        if (rc != NO_KERNEL_HERE) return rc;
    }
    PNODE_REQUIRE(false, "pnode_mlp_rk_adjoint: unsupported dtype %d", a.m->dtype);
}

template <typename T>
__global__ void tanh_probe_kernel(const T *in, T *out, int64_t n) {
    __shared__ T sTab[EXP_TAB];
    if (sizeof(T) == 8)
        for (int j = threadIdx.x; j < EXP_TAB; j += blockDim.x) sTab[j] = (T)exp2((double)j / EXP_TAB);
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = tanh_acc(in[i], sTab);
}

#endif  // PNODE_MLP_PART == 0

}  // namespace pnode

#if PNODE_MLP_PART == 0
using namespace pnode;

extern "C" {

int pnode_mlp_rk_supported(int dim, int hidden, int phi, int dtype, int stages) {
    return (dtype == PNODE_F32 || dtype == PNODE_F64) && shape_ok(dim, hidden, phi, stages) ? 1 : 0;
}

static int mlp_rk_forward_any(const pnode_mlp_desc *mlp, const pnode_rk_tableau *tab, const void *d_u0, int64_t ntraj,
                              const pnode_step *d_sched, int nsteps, void *d_sol, void *d_ckpt, int so, void *stream) {
    PNODE_REQUIRE(mlp && tab && d_sched, "pnode_mlp_rk_forward: null argument");
    PNODE_REQUIRE(shape_ok(mlp->dim, mlp->hidden, mlp->phi, tab->s),
                  "pnode_mlp_rk_forward: unsupported shape dim=%d hidden=%d phi=%d stages=%d", mlp->dim, mlp->hidden,
                  mlp->phi, tab->s);
    if (ntraj == 0 || nsteps == 0) return 0;
    PNODE_REQUIRE(mlp->dtype == PNODE_F32 || mlp->dtype == PNODE_F64, "pnode_mlp_rk_forward: unsupported dtype %d",
                  mlp->dtype);
    return mlp_rk_fwd_any_part(FwdArgs{mlp, tab, d_u0, ntraj, d_sched, nsteps, d_sol, d_ckpt, so,
                                       static_cast<cudaStream_t>(stream)});
}

int pnode_mlp_rk_forward(const pnode_mlp_desc *mlp, const pnode_rk_tableau *tab, const void *d_u0, int64_t ntraj,
                         const pnode_step *d_sched, int nsteps, void *d_sol, void *d_ckpt, void *stream) {
    return mlp_rk_forward_any(mlp, tab, d_u0, ntraj, d_sched, nsteps, d_sol, d_ckpt, 0, stream);
}

int pnode_mlp_rk_forward_so(const pnode_mlp_desc *mlp, const pnode_rk_tableau *tab, const void *d_u0, int64_t ntraj,
                            const pnode_step *d_sched, int nsteps, void *d_sol, void *d_usteps, void *stream) {
    PNODE_REQUIRE(d_usteps != nullptr || nsteps == 0, "pnode_mlp_rk_forward_so: solution checkpoints missing");
    return mlp_rk_forward_any(mlp, tab, d_u0, ntraj, d_sched, nsteps, d_sol, d_usteps, 1, stream);
}

int64_t pnode_mlp_rk_adjoint_work_bytes(const pnode_mlp_desc *mlp) {
    int64_t np = 2 * (int64_t)mlp->hidden * mlp->dim + mlp->hidden + mlp->dim;
    return 64 + (int64_t)ADJ_MAX_BLOCKS * np * (int64_t)sizeof(double);
}

static int mlp_rk_adjoint_any(const pnode_mlp_desc *mlp, const pnode_rk_tableau *tab, int64_t ntraj,
                              const pnode_step *d_sched, int nsteps, int last_slot, const void *d_gout, const void *d_ckpt,
                              void *d_lambda, void *d_mu, void *d_work, const uint64_t *d_peer_bufs, int rank, int world,
                              uint64_t epoch, bool so, void *stream) {
    PNODE_REQUIRE(mlp && tab && d_sched && d_work, "pnode_mlp_rk_adjoint: null argument");
    PNODE_REQUIRE(shape_ok(mlp->dim, mlp->hidden, mlp->phi, tab->s),
                  "pnode_mlp_rk_adjoint: unsupported shape dim=%d hidden=%d phi=%d stages=%d", mlp->dim, mlp->hidden,
                  mlp->phi, tab->s);
    PNODE_REQUIRE(d_ckpt != nullptr || nsteps == 0, "pnode_mlp_rk_adjoint: stage checkpoints missing");
    PNODE_REQUIRE(world <= 1 || d_peer_bufs == nullptr || (epoch >= 1 && rank >= 0 && rank < world && world <= 64),
                  "pnode_mlp_rk_adjoint_dp: bad rank/world/epoch");
    PNODE_REQUIRE(mlp->dtype == PNODE_F32 || mlp->dtype == PNODE_F64, "pnode_mlp_rk_adjoint: unsupported dtype %d",
                  mlp->dtype);
    PeerComm pc{reinterpret_cast<const unsigned long long *>(d_peer_bufs), rank, world, (unsigned long long)epoch};
    return mlp_rk_adj_any_part(AdjArgs{mlp, tab, ntraj, d_sched, nsteps, last_slot, d_gout, d_ckpt, d_lambda, d_mu, d_work,
                                       pc, so, static_cast<cudaStream_t>(stream)});
}

int pnode_mlp_rk_adjoint_dp(const pnode_mlp_desc *mlp, const pnode_rk_tableau *tab, int64_t ntraj,
                            const pnode_step *d_sched, int nsteps, int last_slot, const void *d_gout, const void *d_ckpt,
                            void *d_lambda, void *d_mu, void *d_work, const uint64_t *d_peer_bufs, int rank, int world,
                            uint64_t epoch, void *stream) {
    return mlp_rk_adjoint_any(mlp, tab, ntraj, d_sched, nsteps, last_slot, d_gout, d_ckpt, d_lambda, d_mu, d_work,
                              d_peer_bufs, rank, world, epoch, false, stream);
}

int pnode_mlp_rk_adjoint_so(const pnode_mlp_desc *mlp, const pnode_rk_tableau *tab, int64_t ntraj,
                            const pnode_step *d_sched, int nsteps, int last_slot, const void *d_gout,
                            const void *d_usteps, void *d_lambda, void *d_mu, void *d_work, const uint64_t *d_peer_bufs,
                            int rank, int world, uint64_t epoch, void *stream) {
    return mlp_rk_adjoint_any(mlp, tab, ntraj, d_sched, nsteps, last_slot, d_gout, d_usteps, d_lambda, d_mu, d_work,
                              d_peer_bufs, rank, world, epoch, true, stream);
}

int pnode_mlp_rk_adjoint(const pnode_mlp_desc *mlp, const pnode_rk_tableau *tab, int64_t ntraj,
                         const pnode_step *d_sched, int nsteps, int last_slot, const void *d_gout, const void *d_ckpt,
                         void *d_lambda, void *d_mu, void *d_work, void *stream) {
    return pnode_mlp_rk_adjoint_dp(mlp, tab, ntraj, d_sched, nsteps, last_slot, d_gout, d_ckpt, d_lambda, d_mu, d_work,
                                   nullptr, 0, 1, 0, stream);
}

int64_t pnode_peer_buffer_bytes(int world) {
    return (int64_t)2 * world * PNODE_PEER_NP_MAX * (int64_t)sizeof(double) + (int64_t)2 * world * 8;
}

int pnode_tanh_probe(const void *d_in, void *d_out, int64_t n, int dtype, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == PNODE_F32)
        tanh_probe_kernel<float><<<148, 256, 0, st>>>(static_cast<const float *>(d_in), static_cast<float *>(d_out), n);
    else
        tanh_probe_kernel<double><<<148, 256, 0, st>>>(static_cast<const double *>(d_in), static_cast<double *>(d_out),
                                                       n);
    PNODE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
#endif  // PNODE_MLP_PART == 0
