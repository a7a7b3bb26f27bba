// Convolutional ODE-block right-hand side and its vector-Jacobian products as hand-written sm_100a kernels
// (BASELINE config 4: the SqueezeNext block of /root/reference/examples-pnode/models/sqnxt_PETSc.py:70-121 -- a chain of
// relu(bn_k(conv_k(.))) with 1x1, (1,3) and (3,1) stride-1 "same" convolutions and nn.BatchNorm2d in TRAIN mode).
//
// What it replaces: the arithmetic of the reference's evalRHSFunction (pnode/petsc_adjoint.py:393-412: one forward of the
// module) and RHSJacShell.multTranspose (52-82: one forward re-evaluation + one autograd backward) for this module family.
//
// Design (the first two blocks are HBM-bound, AI 6-12 flop/B; channel counts are 8..32, far too small
// for MMA tiles, so these are CUDA-core FMA kernels whose job is to touch every activation once):
//   * every layer is ONE kernel: it reads the raw output z_{k-1} of the previous convolution, applies that layer's
//     BatchNorm scale/shift + ReLU in registers on load, convolves, adds the bias, writes z_k, and accumulates the
//     per-channel sum / sum of squares of z_k for the NEXT BatchNorm in its epilogue.  The last CTA to finish (ticket
//     counter) combines the per-CTA partial sums in a fixed order (deterministic), writes mean / inv-std / scale / shift and
//     updates running_mean / running_var / num_batches_tracked like nn.BatchNorm2d.  relu(bn(z)) is never materialised
//     between layers: per layer the HBM traffic is (C_in + C_out) scalars per pixel, the lower bound of section 8d.
//   * thread = 4 consecutive pixels of one image row (one 16-byte access per channel) x RC output channels; weights of the
//     CTA's channel tile sit in shared memory and are read as broadcast LDS.128; horizontal taps come from the neighbouring
//     lanes by warp shuffle, vertical taps from the rows above / below (L1 hits).
//   * the backward of a layer is two kernels over the same inputs (g_k = dL/dy_k, z_k, z_{k-1}): a data-gradient kernel --
//     the same convolution kernel with the BatchNorm+ReLU backward  dz = c0 [y>0] g + c1 (z - mean) + c2  applied on load,
//     flipped taps / transposed weights, and the NEXT layer's backward reductions (sum [y>0] g, sum [y>0] g xhat) in its
//     epilogue -- and a weight-gradient kernel (pixels are the reduction dimension: per-CTA register tiles of
//     16 x 32 (c_out x c_in*tap) accumulated over a grid-stride loop of 128-pixel chunks staged in shared memory,
//     per-CTA partials combined in a fixed order by one final kernel that can add  coef * gradient  straight into mu).
#include <stdlib.h>

#include "common.cuh"

namespace pnode {

constexpr int CB_PGX = 64;    // pixel groups (4 pixels each) per CTA tile = blockDim.x
constexpr int CB_MAXCG = 4;   // channel groups per CTA = blockDim.y (<= 256 threads)
enum { COEF_SCALE = 0, COEF_SHIFT, COEF_MEAN, COEF_INVSTD, COEF_C0, COEF_C1, COEF_C2, COEF_DGAMMA, COEF_DBETA, COEF_N };
enum { SRC_RAW = 0, SRC_ACT = 1, SRC_DZ = 2 };
enum { EPI_NONE = 0, EPI_FWD = 1, EPI_DGRAD = 2 };
#define FULL 0xffffffffu

template <typename T>
struct V4 {
    T v[4];
};

__device__ __forceinline__ V4<float> ld4(const float *p) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    V4<float> r;
    r.v[0] = t.x, r.v[1] = t.y, r.v[2] = t.z, r.v[3] = t.w;
    return r;
}
__device__ __forceinline__ V4<double> ld4(const double *p) {
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    V4<double> r;
    r.v[0] = a.x, r.v[1] = a.y, r.v[2] = b.x, r.v[3] = b.y;
    return r;
}
__device__ __forceinline__ void st4(float *p, const V4<float> &r) {
    *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
__device__ __forceinline__ void st4(double *p, const V4<double> &r) {
    *reinterpret_cast<double2 *>(p) = make_double2(r.v[0], r.v[1]);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(r.v[2], r.v[3]);
}
template <typename T>
__device__ __forceinline__ V4<T> zero4() {
    V4<T> r;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = T(0);
    return r;
}

// ---- what a kernel reads: the raw tensor, relu(bn(z)) formed on load, or the BatchNorm+ReLU backward formed on load -------
// Split into coefs(channel) / fetch(offset) / finish(coefs, data) so that a kernel can issue the loads of several channels
// back to back (memory-level parallelism) before it starts consuming them.
template <typename T, int SRC>
struct Source;

template <typename T>
struct Source<T, SRC_RAW> {
    struct Coef {};
    struct Data {
        V4<T> a;
    };
    const T *p;
    __device__ __forceinline__ Source(const T *a, const T *, const T *, int) : p(a) {}
    __device__ __forceinline__ Coef coefs(int) const { return Coef(); }
    __device__ __forceinline__ Data fetch(int64_t off) const {
        Data d;
        d.a = ld4(p + off);
        return d;
    }
    __device__ __forceinline__ V4<T> finish(const Coef &, const Data &d) const { return d.a; }
};

template <typename T>
struct Source<T, SRC_ACT> {
    struct Coef {
        T sc, sh;
    };
    struct Data {
        V4<T> a;
    };
    const T *p, *coef;
    int cs;
    __device__ __forceinline__ Source(const T *a, const T *, const T *c, int s) : p(a), coef(c), cs(s) {}
    __device__ __forceinline__ Coef coefs(int ch) const {
        Coef c;
        c.sc = __ldg(coef + COEF_SCALE * cs + ch);
        c.sh = __ldg(coef + COEF_SHIFT * cs + ch);
        return c;
    }
    __device__ __forceinline__ Data fetch(int64_t off) const {
        Data d;
        d.a = ld4(p + off);
        return d;
    }
    __device__ __forceinline__ V4<T> finish(const Coef &c, const Data &d) const {
        V4<T> v;
#pragma unroll
        for (int e = 0; e < 4; ++e) v.v[e] = fmax(fma(c.sc, d.a.v[e], c.sh), T(0));
        return v;
    }
};

template <typename T>
struct Source<T, SRC_DZ> {
    struct Coef {
        T sc, sh, mean, c0, c1, c2;
    };
    struct Data {
        V4<T> g, z;
    };
    const T *g, *z, *coef;
    int cs;
    __device__ __forceinline__ Source(const T *a, const T *b, const T *c, int s) : g(a), z(b), coef(c), cs(s) {}
    __device__ __forceinline__ Coef coefs(int ch) const {
        Coef c;
        c.sc = __ldg(coef + COEF_SCALE * cs + ch);
        c.sh = __ldg(coef + COEF_SHIFT * cs + ch);
        c.mean = __ldg(coef + COEF_MEAN * cs + ch);
        c.c0 = __ldg(coef + COEF_C0 * cs + ch);
        c.c1 = __ldg(coef + COEF_C1 * cs + ch);
        c.c2 = __ldg(coef + COEF_C2 * cs + ch);
        return c;
    }
    __device__ __forceinline__ Data fetch(int64_t off) const {
        Data d;
        d.g = ld4(g + off);
        d.z = ld4(z + off);
        return d;
    }
    __device__ __forceinline__ V4<T> finish(const Coef &c, const Data &d) const {
        V4<T> r;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const T base = fma(c.c1, d.z.v[e] - c.mean, c.c2);
            r.v[e] = fma(c.sc, d.z.v[e], c.sh) > T(0) ? fma(c.c0, d.g.v[e], base) : base;
        }
        return r;
    }
};

// KIND 0: 1x1.  KIND 1: (1,3) pad (0,1): offsets -1, 0, +1 along W.  KIND 2: (3,1) pad (1,0): offsets -W, 0, +W.
// fetch_offsets issues the loads (KIND 2: the rows above / below, clamped at the image border so that every load is
// unconditional); finish_offsets turns them into o[i] = the 4 source values at offset (i - 1) of the thread's 4 pixels, zero
// outside the image.  finish_offsets must be called by all 32 lanes of a warp (KIND 1 shuffles); lanes hold consecutive pixel
// groups and a warp covers whole image rows.  Lanes past the end of the tensor read pixel group 0 and are never stored.
template <typename T, int KIND, typename S>
__device__ __forceinline__ void fetch_offsets(const S &src, int64_t off, int h, int H, int W,
                                              typename S::Data (&d)[KIND == 2 ? 3 : 1]) {
    if constexpr (KIND == 2) {
        d[0] = src.fetch(off - (h > 0 ? W : 0));
        d[1] = src.fetch(off);
        d[2] = src.fetch(off + (h < H - 1 ? W : 0));
    } else {
        d[0] = src.fetch(off);
    }
}

template <typename T, int KIND, typename S>
__device__ __forceinline__ void finish_offsets(const S &src, const typename S::Coef &cf,
                                               const typename S::Data (&d)[KIND == 2 ? 3 : 1], int h, int w0, int H, int W,
                                               V4<T> (&o)[KIND == 0 ? 1 : 3]) {
    if constexpr (KIND == 0) {
        o[0] = src.finish(cf, d[0]);
    } else if constexpr (KIND == 1) {
        const V4<T> c = src.finish(cf, d[0]);
        T l = __shfl_up_sync(FULL, c.v[3], 1), r = __shfl_down_sync(FULL, c.v[0], 1);
        if (w0 == 0) l = T(0);
        if (w0 + 4 == W) r = T(0);
        o[0].v[0] = l, o[0].v[1] = c.v[0], o[0].v[2] = c.v[1], o[0].v[3] = c.v[2];
        o[1] = c;
        o[2].v[0] = c.v[1], o[2].v[1] = c.v[2], o[2].v[2] = c.v[3], o[2].v[3] = r;
    } else {
        o[0] = h > 0 ? src.finish(cf, d[0]) : zero4<T>();
        o[1] = src.finish(cf, d[1]);
        o[2] = h < H - 1 ? src.finish(cf, d[2]) : zero4<T>();
    }
}

// channels whose loads are in flight together, per thread (registers: UN * NF * (1 or 2) * 4 scalars)
template <int KIND, int SRC>
struct Unroll {
    static constexpr int N = KIND == 2 ? (SRC == SRC_DZ ? 2 : 4) : 4;
};

struct PixelCoord {
    int n, rem, h, w0;
    bool active;
};
__device__ __forceinline__ PixelCoord pixel_coord(int64_t pg, int64_t npg, int HW, int W) {
    PixelCoord c;
    c.active = pg < npg;
    const int64_t q = c.active ? pg * 4 : 0;
    c.n = (int)(q / HW);
    c.rem = (int)(q - (int64_t)c.n * HW);
    c.h = c.rem / W;
    c.w0 = c.rem - c.h * W;
    return c;
}

template <typename T>
struct ConvArgs {
    const T *in, *in2;      // forward: z_{k-1} (or x), unused;  data gradient: g_k, z_k
    const T *coef_in;       // per-channel coefficients of the BatchNorm on the INPUT side (forward: layer k-1; dgrad: layer k)
    const T *w, *bias;      // conv weight [Cout][Cin][taps]; bias [Cout] (forward only)
    T *out;                 // forward: z_k;  data gradient: g_{k-1} (or the VJP w.r.t. the block input)
    const T *zprev;         // EPI_DGRAD: z_{k-1} at the output pixels
    T *coef_out;            // coefficients finalised by the last CTA (forward: layer k; dgrad: layer k-1)
    double *partial;        // [gridDim.x][CB][2]
    unsigned *counter;
    const T *gamma, *beta;  // forward finalise
    T *rmean, *rvar;
    long long *nbt;
    double eps, momentum;
    int CA, CB;             // reduction channels, output channels
    int H, W, HW, cs;
    int64_t npg;            // pixel groups = N*H*W/4
    int tiles;
    double M;               // N*H*W
};

// 2*RC per-thread values summed over the 32 lanes with a transposing butterfly (2*RC + log-ish shuffles instead of 5 per
// value): on return v[0] of lane l holds the warp total of value index idx(l) = bits 4..(5-log2(NV)) of l, mirrored.
template <typename T, int NV>
__device__ __forceinline__ T warp_multi_sum(T (&v)[NV], int lane, int &idx) {
    idx = 0;
    int bit = 16;
#pragma unroll
    for (int n = NV / 2; n >= 1; n >>= 1, bit >>= 1) {
        const bool hi = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const T send = hi ? v[i] : v[i + n];
            const T keep = hi ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, bit);
        }
        idx = idx * 2 + (hi ? 1 : 0);
    }
    T t = v[0];
    for (; bit >= 1; bit >>= 1) t += __shfl_xor_sync(FULL, t, bit);
    return t;
}

// Per-CTA partial sums -> global; the last CTA combines them (fixed order) and finalises the per-channel coefficients.
template <typename T, int RC, int EPI>
__device__ __forceinline__ void stats_epilogue(const ConvArgs<T> &a, T (&s)[RC], T (&q)[RC], int cb0) {
    __shared__ double red[2][CB_MAXCG][2 * RC];
    __shared__ double tot[2 * 512];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, wx = threadIdx.x >> 5;
    const int tid = threadIdx.y * CB_PGX + threadIdx.x, nthr = CB_PGX * blockDim.y;
    T v[2 * RC];
#pragma unroll
    for (int r = 0; r < RC; ++r) v[r] = s[r], v[RC + r] = q[r];
    int idx;
    const T t = warp_multi_sum<T, 2 * RC>(v, lane, idx);
    // lanes sharing idx hold the same total; one of them publishes it
    constexpr int REP = 32 / (2 * RC);
    if ((lane & (REP - 1)) == 0) red[wx][threadIdx.y][idx] = (double)t;
    __syncthreads();
    if (threadIdx.x < 2 * RC) {
        const int which = threadIdx.x / RC, r = threadIdx.x % RC;
        const int ch = cb0 + threadIdx.y * RC + r;
        a.partial[((int64_t)blockIdx.x * a.CB + ch) * 2 + which] = red[0][threadIdx.y][threadIdx.x] + red[1][threadIdx.y][threadIdx.x];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned ticket = atomicAdd(a.counter, 1u);
        is_last = ticket == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int warp = tid >> 5, nwarp = nthr >> 5;
    for (int base = 0; base < a.CB; base += 512) {  // tot[] holds up to 512 channels at a time
        const int nch = min(512, a.CB - base);
        for (int pair = warp; pair < 2 * nch; pair += nwarp) {
            const int ch = base + (pair >> 1), which = pair & 1;
            double acc = 0.0;
            for (int b = lane; b < (int)gridDim.x; b += 32) acc += __ldcg(a.partial + ((int64_t)b * a.CB + ch) * 2 + which);
            acc = warp_sum(acc);
            if (lane == 0) tot[pair] = acc;
        }
        __syncthreads();
        for (int c = tid; c < nch; c += nthr) {
            const int ch = base + c;
            const double S = tot[2 * c], Q = tot[2 * c + 1];
            T *co = a.coef_out;
            if (EPI == EPI_FWD) {
                const double mean = S / a.M;
                const double var = fmax(Q / a.M - mean * mean, 0.0);
                const double invstd = 1.0 / sqrt(var + a.eps);
                const double scale = (double)a.gamma[ch] * invstd;
                co[COEF_SCALE * a.cs + ch] = (T)scale;
                co[COEF_SHIFT * a.cs + ch] = (T)((double)a.beta[ch] - mean * scale);
                co[COEF_MEAN * a.cs + ch] = (T)mean;
                co[COEF_INVSTD * a.cs + ch] = (T)invstd;
                if (a.rmean != nullptr) {
                    const double unbiased = a.M > 1.0 ? var * a.M / (a.M - 1.0) : var;
                    a.rmean[ch] = (T)((1.0 - a.momentum) * (double)a.rmean[ch] + a.momentum * mean);
                    a.rvar[ch] = (T)((1.0 - a.momentum) * (double)a.rvar[ch] + a.momentum * unbiased);
                }
            } else {
                // S = sum [y>0] g = dbeta, Q = sum [y>0] g xhat = dgamma;  dz = c0 [y>0] g + c1 (z - mean) + c2
                const double c0 = (double)co[COEF_SCALE * a.cs + ch], invstd = (double)co[COEF_INVSTD * a.cs + ch];
                co[COEF_C0 * a.cs + ch] = (T)c0;
                co[COEF_C1 * a.cs + ch] = (T)(-c0 * invstd * Q / a.M);
                co[COEF_C2 * a.cs + ch] = (T)(-c0 * S / a.M);
                co[COEF_DGAMMA * a.cs + ch] = (T)Q;
                co[COEF_DBETA * a.cs + ch] = (T)S;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (EPI == EPI_FWD && a.nbt != nullptr) *a.nbt += 1;
        *a.counter = 0u;
    }
}

// sum / sum-of-squares (forward) or the two BatchNorm-backward sums (dgrad) of one thread's RC x 4 outputs
template <typename T, int RC, int EPI>
__device__ __forceinline__ void stats_accumulate(const ConvArgs<T> &a, const T (&acc)[RC][4], int64_t off_out, int ch0,
                                                 T (&s)[RC], T (&q)[RC]) {
    if constexpr (EPI == EPI_FWD) {
#pragma unroll
        for (int r = 0; r < RC; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[r] += acc[r][j];
                q[r] = fma(acc[r][j], acc[r][j], q[r]);
            }
    } else {
        const T *co = a.coef_out;
        V4<T> z[RC];
        T sc[RC], sh[RC], mean[RC], invstd[RC];
#pragma unroll
        for (int r = 0; r < RC; ++r) {
            z[r] = ld4(a.zprev + off_out + (int64_t)r * a.HW);
            sc[r] = __ldg(co + COEF_SCALE * a.cs + ch0 + r), sh[r] = __ldg(co + COEF_SHIFT * a.cs + ch0 + r);
            mean[r] = __ldg(co + COEF_MEAN * a.cs + ch0 + r), invstd[r] = __ldg(co + COEF_INVSTD * a.cs + ch0 + r);
        }
#pragma unroll
        for (int r = 0; r < RC; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const T gm = fma(sc[r], z[r].v[j], sh[r]) > T(0) ? acc[r][j] : T(0);
                s[r] += gm;
                q[r] = fma(gm, (z[r].v[j] - mean[r]) * invstd[r], q[r]);
            }
    }
}

template <typename T>
__device__ __forceinline__ V4<T> lds4(const T *p);
template <>
__device__ __forceinline__ V4<float> lds4<float>(const float *p) {
    return ld4(p);
}
template <>
__device__ __forceinline__ V4<double> lds4<double>(const double *p) {
    return ld4(p);
}

// out[b][pixel] = bias[b] + sum_{a, offset} Wsel[a][offset][b] * src(a, pixel + offset)
template <typename T, int KIND, int SRC, int EPI, int RC>
__global__ void __launch_bounds__(CB_PGX *CB_MAXCG) conv_kernel(const ConvArgs<T> a) {
    constexpr int TAPS = KIND == 0 ? 1 : 3;
    constexpr bool FWD = SRC != SRC_DZ;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *ws = reinterpret_cast<T *>(smem_raw);  // [CA * TAPS][CT]
    const int CT = blockDim.y * RC;
    const int cb0 = blockIdx.y * CT;
    const int tid = threadIdx.y * CB_PGX + threadIdx.x, nthr = CB_PGX * blockDim.y;
    for (int i = tid; i < a.CA * TAPS * CT; i += nthr) {
        const int bl = i % CT, ao = i / CT, o = ao % TAPS, ach = ao / TAPS, b = cb0 + bl;
        // forward: W[co = b][ci = ach][tap = o];  data gradient: W[co = ach][ci = b][tap = TAPS-1-o]
        ws[i] = FWD ? a.w[((int64_t)b * a.CA + ach) * TAPS + o] : a.w[((int64_t)ach * a.CB + b) * TAPS + (TAPS - 1 - o)];
    }
    __syncthreads();
    Source<T, SRC> src(a.in, a.in2, a.coef_in, a.cs);
    const int ch0 = cb0 + threadIdx.y * RC;
    T bias[RC], s[RC], q[RC];
#pragma unroll
    for (int r = 0; r < RC; ++r) {
        bias[r] = (FWD && a.bias != nullptr) ? a.bias[ch0 + r] : T(0);
        s[r] = q[r] = T(0);
    }
    for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
        const PixelCoord pc = pixel_coord((int64_t)tile * CB_PGX + threadIdx.x, a.npg, a.HW, a.W);
        const int64_t off_in = (int64_t)pc.n * a.CA * a.HW + pc.rem;
        T acc[RC][4];
#pragma unroll
        for (int r = 0; r < RC; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][j] = bias[r];
        constexpr int UN = Unroll<KIND, SRC>::N, NF = KIND == 2 ? 3 : 1;
        for (int cb = 0; cb < a.CA; cb += UN) {  // host guarantees CA % 4 == 0
            typename Source<T, SRC>::Coef cf[UN];
            typename Source<T, SRC>::Data dd[UN][NF];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                cf[u] = src.coefs(cb + u);
                fetch_offsets<T, KIND>(src, off_in + (int64_t)(cb + u) * a.HW, pc.h, a.H, a.W, dd[u]);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                V4<T> o[TAPS];
                finish_offsets<T, KIND>(src, cf[u], dd[u], pc.h, pc.w0, a.H, a.W, o);
#pragma unroll
                for (int oi = 0; oi < TAPS; ++oi) {
                    const T *wr = ws + ((cb + u) * TAPS + oi) * CT + threadIdx.y * RC;
#pragma unroll
                    for (int r4 = 0; r4 < RC; r4 += 4) {
                        const V4<T> wv = lds4<T>(wr + r4);
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[r4 + r][j] = fma(wv.v[r], o[oi].v[j], acc[r4 + r][j]);
                    }
                }
            }
        }
        if (pc.active) {
            const int64_t off_out = ((int64_t)pc.n * a.CB + ch0) * a.HW + pc.rem;
#pragma unroll
            for (int r = 0; r < RC; ++r) {
                V4<T> v;
#pragma unroll
                for (int j = 0; j < 4; ++j) v.v[j] = acc[r][j];
                st4(a.out + off_out + (int64_t)r * a.HW, v);
            }
            if (EPI != EPI_NONE) stats_accumulate<T, RC, EPI>(a, acc, off_out, ch0, s, q);
        }
    }
    if (EPI != EPI_NONE) stats_epilogue<T, RC, EPI>(a, s, q, cb0);
}

// BatchNorm-backward sums of the TOP layer (g = the cotangent handed to the VJP, z = z_L): same epilogue, no convolution.
template <typename T, int RC>
__global__ void __launch_bounds__(CB_PGX *CB_MAXCG) top_stats_kernel(const ConvArgs<T> a) {
    const int CT = blockDim.y * RC;
    const int cb0 = blockIdx.y * CT, ch0 = cb0 + threadIdx.y * RC;
    T s[RC], q[RC];
#pragma unroll
    for (int r = 0; r < RC; ++r) s[r] = q[r] = T(0);
    for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
        const PixelCoord pc = pixel_coord((int64_t)tile * CB_PGX + threadIdx.x, a.npg, a.HW, a.W);
        if (!pc.active) continue;
        const int64_t off_out = ((int64_t)pc.n * a.CB + ch0) * a.HW + pc.rem;
        T acc[RC][4];
        V4<T> g[RC];
#pragma unroll
        for (int r = 0; r < RC; ++r) g[r] = ld4(a.in + off_out + (int64_t)r * a.HW);
#pragma unroll
        for (int r = 0; r < RC; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][j] = g[r].v[j];
        stats_accumulate<T, RC, EPI_DGRAD>(a, acc, off_out, ch0, s, q);
    }
    stats_epilogue<T, RC, EPI_DGRAD>(a, s, q, cb0);
}

// out = base_coef * base + coef * relu(scale_c * z + shift_c)      (base may be NULL: plain activation of the last layer;
// with base it is the RK stage combination Y_{i+1} = u + h a_{i+1,i} k_i fused into the block's output pass)
template <typename T>
__global__ void __launch_bounds__(256) act_out_kernel(const T *__restrict__ z, const T *__restrict__ coef, int cs, int C,
                                                      int HW, int64_t nvec, T *__restrict__ out, const T *__restrict__ base,
                                                      T base_coef, T kcoef, T *__restrict__ kout) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const int64_t e0 = i * 4;
        const int c = (int)((e0 / HW) % C);
        const T sc = __ldg(coef + COEF_SCALE * cs + c), sh = __ldg(coef + COEF_SHIFT * cs + c);
        V4<T> v = ld4(z + e0);
#pragma unroll
        for (int e = 0; e < 4; ++e) v.v[e] = fmax(fma(sc, v.v[e], sh), T(0));
        if (kout != nullptr) st4(kout + e0, v);
        if (base != nullptr) {
            const V4<T> b = ld4(base + e0);
#pragma unroll
            for (int e = 0; e < 4; ++e) v.v[e] = fma(kcoef, v.v[e], base_coef * b.v[e]);
        }
        if (out != nullptr) st4(out + e0, v);
    }
}

// ---- weight gradient ------------------------------------------------------------------------------------------------------
constexpr int WG_CO = 16, WG_CIT = 32, WG_PX = 128, WG_LD = 132, WG_THREADS = 256, WG_TILE = WG_CO * WG_CIT + WG_CO;

template <typename T>
struct WgradArgs {
    const T *g, *z, *coef_k;   // dz_k formed on load from g_k, z_k and layer k's coefficients
    const T *yin, *coef_in;    // y_{k-1}: raw block input (k = 1) or relu(bn(z_{k-1}))
    T *partial;                // [gridDim.x][gridDim.y][WG_TILE]
    int Cin, Cout, H, W, HW, cs, tiles_cit;
    int64_t npg;
    int nchunks;
};

template <typename T, int KIND, int YSRC>
__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_kernel(const WgradArgs<T> a) {
    constexpr int TAPS = KIND == 0 ? 1 : 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_dz = reinterpret_cast<T *>(smem_raw);  // [WG_CO][WG_LD]
    T *s_y = s_dz + WG_CO * WG_LD;              // [WG_CIT][WG_LD]
    const int tile = blockIdx.y, tco = tile / a.tiles_cit, tcit = tile % a.tiles_cit;
    const int co0 = tco * WG_CO, cit0 = tcit * WG_CIT;
    const int ci_lo = cit0 / TAPS, ci_hi = min(a.Cin, (cit0 + WG_CIT - 1) / TAPS + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cgp = lane & 3, cg8 = lane >> 2;
    for (int i = threadIdx.x; i < (WG_CO + WG_CIT) * WG_LD; i += WG_THREADS) s_dz[i] = T(0);
    __syncthreads();
    Source<T, SRC_DZ> dsrc(a.g, a.z, a.coef_k, a.cs);
    Source<T, YSRC> ysrc(a.yin, nullptr, a.coef_in, a.cs);
    T acc[4][4], accb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        accb[i] = T(0);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    }
    for (int chunk = blockIdx.x; chunk < a.nchunks; chunk += gridDim.x) {
        const PixelCoord pc = pixel_coord((int64_t)chunk * 32 + lane, a.npg, a.HW, a.W);
        // issue every load of this warp's rows (2 dz rows, up to NY source channels) before consuming any of them
        constexpr int NY = TAPS == 1 ? WG_CIT / 8 : 2, NF = KIND == 2 ? 3 : 1;
        typename Source<T, SRC_DZ>::Coef cdz[2];
        typename Source<T, SRC_DZ>::Data ddz[2];
        typename Source<T, YSRC>::Coef cy[NY];
        typename Source<T, YSRC>::Data dy[NY][NF];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int co = min(co0 + warp + 8 * i, a.Cout - 1);
            cdz[i] = dsrc.coefs(co);
            ddz[i] = dsrc.fetch(((int64_t)pc.n * a.Cout + co) * a.HW + pc.rem);
        }
#pragma unroll
        for (int u = 0; u < NY; ++u) {
            const int ci = min(ci_lo + warp + 8 * u, ci_hi - 1);
            cy[u] = ysrc.coefs(ci);
            fetch_offsets<T, KIND>(ysrc, ((int64_t)pc.n * a.Cin + ci) * a.HW + pc.rem, pc.h, a.H, a.W, dy[u]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = warp + 8 * i;
            if (co0 + r < a.Cout) st4(s_dz + r * WG_LD + lane * 4, pc.active ? dsrc.finish(cdz[i], ddz[i]) : zero4<T>());
        }
#pragma unroll
        for (int u = 0; u < NY; ++u) {
            const int ci = ci_lo + warp + 8 * u;  // warp-uniform
            V4<T> o[TAPS];
            finish_offsets<T, KIND>(ysrc, cy[u], dy[u], pc.h, pc.w0, a.H, a.W, o);
            if (ci < ci_hi) {
#pragma unroll
                for (int oi = 0; oi < TAPS; ++oi) {
                    const int r = ci * TAPS + oi - cit0;
                    if (r >= 0 && r < WG_CIT) st4(s_y + r * WG_LD + lane * 4, pc.active ? o[oi] : zero4<T>());
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int pp = 0; pp < WG_PX / 8; pp += 4) {
            const int p = warp * (WG_PX / 8) + pp;
            V4<T> av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = lds4<T>(s_dz + (cgp + 4 * i) * WG_LD + p);
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = lds4<T>(s_y + (cg8 + 8 * j) * WG_LD + p);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int e = 0; e < 4; ++e) accb[i] += av[i].v[e];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[i][j] = fma(av[i].v[e], bv[j].v[e], acc[i][j]);
            }
        }
        __syncthreads();
    }
    // combine the 8 warps (each saw different pixels) in a fixed order, one partial tile per CTA
    T *red = s_dz;  // 8 * WG_TILE scalars <= (WG_CO + WG_CIT) * WG_LD
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) red[warp * WG_TILE + (cgp + 4 * i) * WG_CIT + cg8 + 8 * j] = acc[i][j];
        if (cg8 == 0) red[warp * WG_TILE + WG_CO * WG_CIT + cgp + 4 * i] = accb[i];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < WG_TILE; idx += WG_THREADS) {
        T t = T(0);
#pragma unroll
        for (int w = 0; w < WG_THREADS / 32; ++w) t += red[w * WG_TILE + idx];
        a.partial[((int64_t)blockIdx.x * gridDim.y + tile) * WG_TILE + idx] = t;
    }
}

template <typename T>
struct GradLayer {
    const T *partial, *coef;
    int PX, ntiles, tiles_cit, Cin, Cout, taps;
    int64_t base;  // offset of this layer's first parameter (conv.weight, conv.bias, bn.weight, bn.bias follow each other)
};
template <typename T>
struct GradArgs {
    GradLayer<T> L[PNODE_CONV_MAX_LAYERS];
    int nl, cs, accumulate;
    int64_t np;
    T *out;
    double coef;
};

// one group of G lanes per parameter: the lanes stride over the per-CTA partial tiles, then combine in a fixed order
template <typename T, int G>
__global__ void __launch_bounds__(256) grads_finalize_kernel(const GradArgs<T> a) {
    const int sub = threadIdx.x % G;
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x / G);
    const int64_t rounds = (a.np + stride - 1) / stride;
    for (int64_t rd = 0; rd < rounds; ++rd) {
        const int64_t i = rd * stride + (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
        const bool live = i < a.np;
        double val = 0.0;
        if (live) {
            int k = 0;
            while (k + 1 < a.nl && i >= a.L[k + 1].base) ++k;
            const GradLayer<T> &l = a.L[k];
            int64_t loc = i - l.base;
            const int64_t nw = (int64_t)l.Cout * l.Cin * l.taps;
            if (loc < nw + l.Cout) {
                int tile, idx;
                if (loc < nw) {
                    const int co = (int)(loc / (l.Cin * l.taps)), cit = (int)(loc % (l.Cin * l.taps));
                    tile = (co / WG_CO) * l.tiles_cit + cit / WG_CIT;
                    idx = (co % WG_CO) * WG_CIT + cit % WG_CIT;
                } else {
                    const int co = (int)(loc - nw);
                    tile = (co / WG_CO) * l.tiles_cit;
                    idx = WG_CO * WG_CIT + co % WG_CO;
                }
                for (int b = sub; b < l.PX; b += G) val += (double)__ldcg(l.partial + ((int64_t)b * l.ntiles + tile) * WG_TILE + idx);
            } else if (sub == 0) {
                loc -= nw + l.Cout;
                val = loc < l.Cout ? (double)l.coef[COEF_DGAMMA * a.cs + loc] : (double)l.coef[COEF_DBETA * a.cs + (loc - l.Cout)];
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) val += __shfl_xor_sync(FULL, val, o);
        if (live && sub == 0) a.out[i] = a.accumulate ? (T)((double)a.out[i] + a.coef * val) : (T)val;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
struct CbPlan {
    int L, Cmax, cs;
    int64_t M, npg;
    size_t esz;
    size_t off_z[PNODE_CONV_MAX_LAYERS], off_g[2], off_coef[PNODE_CONV_MAX_LAYERS], off_stat, off_wpart[PNODE_CONV_MAX_LAYERS];
    size_t off_counter, total;
    int kind[PNODE_CONV_MAX_LAYERS], taps[PNODE_CONV_MAX_LAYERS];
    int wg_px[PNODE_CONV_MAX_LAYERS], wg_tiles_cit[PNODE_CONV_MAX_LAYERS], wg_ntiles[PNODE_CONV_MAX_LAYERS];
    int64_t pbase[PNODE_CONV_MAX_LAYERS + 1];
};

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct ConvGrid {
    int rc, cg, gy, gx;
    size_t smem;
};

// thread = 4 pixels x rc output channels; rc = 4 when 8 channels per thread would leave the SMs short of resident warps
static ConvGrid conv_grid(int CA, int CB, int taps, int64_t npg, size_t esz) {
    ConvGrid g;
    const int sm = sm_count();
    g.rc = (CB % 8 == 0 && (int64_t)(CB / 8) * npg >= (int64_t)sm * 1024) ? 8 : 4;
    static const int force_rc = getenv("PNODE_CONV_RC") ? atoi(getenv("PNODE_CONV_RC")) : 0;  // tuning knob
    if (force_rc == 4 || (force_rc == 8 && CB % 8 == 0)) g.rc = force_rc;
    int groups = CB / g.rc;
    g.cg = groups < CB_MAXCG ? groups : CB_MAXCG;
    while (groups % g.cg != 0) --g.cg;
    g.gy = groups / g.cg;
    const int tiles = (int)((npg + CB_PGX - 1) / CB_PGX);
    static const int occ = getenv("PNODE_CONV_OCC") ? atoi(getenv("PNODE_CONV_OCC")) : 1024;  // tuning knob
    int cap = sm * (occ / (CB_PGX * g.cg)) / g.gy;
    if (cap < 1) cap = 1;
    g.gx = tiles < cap ? tiles : cap;
    g.smem = (size_t)CA * taps * g.cg * g.rc * esz;
    return g;
}

static int cb_plan(const pnode_convblock_desc *d, CbPlan &p) {
    PNODE_REQUIRE(d != nullptr, "convblock: null descriptor");
    PNODE_REQUIRE(d->nlayers >= 1 && d->nlayers <= PNODE_CONV_MAX_LAYERS, "convblock: 1..%d layers (got %d)",
                  PNODE_CONV_MAX_LAYERS, d->nlayers);
    PNODE_REQUIRE(d->dtype == PNODE_F32 || d->dtype == PNODE_F64, "convblock: unsupported dtype %d", d->dtype);
    PNODE_REQUIRE(d->N >= 1 && d->H >= 1 && d->W >= 4 && d->W % 4 == 0 && (d->W & (d->W - 1)) == 0 && d->W <= 128,
                  "convblock: W must be a power of two in 4..128 (got N=%d H=%d W=%d)", d->N, d->H, d->W);
    p.L = d->nlayers;
    p.esz = d->dtype == PNODE_F32 ? 4 : 8;
    p.M = (int64_t)d->N * d->H * d->W;
    p.npg = p.M / 4;
    p.Cmax = 0;
    p.pbase[0] = 0;
    for (int k = 0; k < p.L; ++k) {
        const pnode_conv_layer &l = d->layer[k];
        PNODE_REQUIRE(l.cin >= 4 && l.cout >= 4 && l.cin % 4 == 0 && l.cout % 4 == 0,
                      "convblock: channel counts must be multiples of 4 (layer %d: %d -> %d)", k, l.cin, l.cout);
        PNODE_REQUIRE(k == 0 || l.cin == d->layer[k - 1].cout, "convblock: layer %d input channels do not chain", k);
        if (l.kh == 1 && l.kw == 1 && l.ph == 0 && l.pw == 0) p.kind[k] = 0;
        else if (l.kh == 1 && l.kw == 3 && l.ph == 0 && l.pw == 1) p.kind[k] = 1;
        else if (l.kh == 3 && l.kw == 1 && l.ph == 1 && l.pw == 0) p.kind[k] = 2;
        else PNODE_REQUIRE(false, "convblock: layer %d kernel (%d,%d) pad (%d,%d) is not 1x1 / (1,3)p(0,1) / (3,1)p(1,0)", k,
                           l.kh, l.kw, l.ph, l.pw);
        PNODE_REQUIRE(l.d_weight && l.d_bias && l.d_gamma && l.d_beta, "convblock: layer %d has a null parameter", k);
        p.taps[k] = p.kind[k] == 0 ? 1 : 3;
        if (l.cin > p.Cmax) p.Cmax = l.cin;
        if (l.cout > p.Cmax) p.Cmax = l.cout;
        p.pbase[k + 1] = p.pbase[k] + (int64_t)l.cout * l.cin * p.taps[k] + 3 * (int64_t)l.cout;
    }
    p.cs = p.Cmax;
    size_t off = 0;
    for (int k = 0; k < p.L; ++k) {
        p.off_z[k] = off;
        off = align_up(off + (size_t)p.M * d->layer[k].cout * p.esz);
    }
    for (int i = 0; i < 2; ++i) {
        p.off_g[i] = off;
        off = align_up(off + (size_t)p.M * p.Cmax * p.esz);
    }
    for (int k = 0; k < p.L; ++k) {
        p.off_coef[k] = off;
        off = align_up(off + (size_t)COEF_N * p.cs * p.esz);
    }
    size_t stat = 0;
    const int sm = sm_count();
    for (int k = 0; k < p.L; ++k) {
        const pnode_conv_layer &l = d->layer[k];
        const ConvGrid f = conv_grid(l.cin, l.cout, p.taps[k], p.npg, p.esz);
        const ConvGrid b = conv_grid(l.cout, l.cin, p.taps[k], p.npg, p.esz);
        const ConvGrid t = conv_grid(0, l.cout, 1, p.npg, p.esz);
        size_t need = (size_t)f.gx * l.cout;
        if ((size_t)b.gx * l.cin > need) need = (size_t)b.gx * l.cin;
        if ((size_t)t.gx * l.cout > need) need = (size_t)t.gx * l.cout;
        if (need > stat) stat = need;
        PNODE_REQUIRE(f.smem <= 200 * 1024 && b.smem <= 200 * 1024, "convblock: layer %d weights do not fit in shared memory", k);
        p.wg_tiles_cit[k] = (l.cin * p.taps[k] + WG_CIT - 1) / WG_CIT;
        p.wg_ntiles[k] = ((l.cout + WG_CO - 1) / WG_CO) * p.wg_tiles_cit[k];
        const int nchunks = (int)((p.npg + 31) / 32);
        int px = sm * 4 / p.wg_ntiles[k];
        if (px < 1) px = 1;
        if (px > nchunks) px = nchunks;
        p.wg_px[k] = px;
    }
    p.off_stat = off;
    off = align_up(off + stat * 2 * sizeof(double));
    for (int k = 0; k < p.L; ++k) {
        p.off_wpart[k] = off;
        off = align_up(off + (size_t)p.wg_px[k] * p.wg_ntiles[k] * WG_TILE * p.esz);
    }
    p.off_counter = off;
    off = align_up(off + 256);
    p.total = off;
    return 0;
}

template <typename T, int KIND, int SRC, int EPI>
static int launch_conv_rc(const ConvArgs<T> &a, const ConvGrid &g, cudaStream_t st) {
    dim3 grid(g.gx, g.gy), block(CB_PGX, g.cg);
    if (g.rc == 8) {
        if (g.smem > 32 * 1024)
            PNODE_CUDA_OK(cudaFuncSetAttribute(conv_kernel<T, KIND, SRC, EPI, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        conv_kernel<T, KIND, SRC, EPI, 8><<<grid, block, g.smem, st>>>(a);
    } else {
        if (g.smem > 32 * 1024)
            PNODE_CUDA_OK(cudaFuncSetAttribute(conv_kernel<T, KIND, SRC, EPI, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        conv_kernel<T, KIND, SRC, EPI, 4><<<grid, block, g.smem, st>>>(a);
    }
    return 0;
}

template <typename T, int SRC, int EPI>
static int launch_conv(int kind, const ConvArgs<T> &a, const ConvGrid &g, cudaStream_t st) {
    if (kind == 0) return launch_conv_rc<T, 0, SRC, EPI>(a, g, st);
    if (kind == 1) return launch_conv_rc<T, 1, SRC, EPI>(a, g, st);
    return launch_conv_rc<T, 2, SRC, EPI>(a, g, st);
}

template <typename T, int YSRC>
static int launch_wgrad(int kind, const WgradArgs<T> &a, int px, int ntiles, cudaStream_t st) {
    const size_t smem = (size_t)(WG_CO + WG_CIT) * WG_LD * sizeof(T);
    dim3 grid(px, ntiles);
#define PNODE_WG(K)                                                                                                        \
    do {                                                                                                                   \
        if (smem > 48 * 1024)                                                                                              \
            PNODE_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<T, K, YSRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)smem));                                                               \
        conv_wgrad_kernel<T, K, YSRC><<<grid, WG_THREADS, smem, st>>>(a);                                                  \
    } while (0)
    if (kind == 0) PNODE_WG(0);
    else if (kind == 1) PNODE_WG(1);
    else PNODE_WG(2);
#undef PNODE_WG
    return 0;
}

template <typename T>
struct Bufs {
    unsigned char *w;
    const CbPlan &p;
    Bufs(void *work, const CbPlan &pl) : w(static_cast<unsigned char *>(work)), p(pl) {}
    T *z(int k) const { return reinterpret_cast<T *>(w + p.off_z[k]); }
    T *g(int i) const { return reinterpret_cast<T *>(w + p.off_g[i]); }
    T *coef(int k) const { return reinterpret_cast<T *>(w + p.off_coef[k]); }
    double *stat() const { return reinterpret_cast<double *>(w + p.off_stat); }
    T *wpart(int k) const { return reinterpret_cast<T *>(w + p.off_wpart[k]); }
    unsigned *counter() const { return reinterpret_cast<unsigned *>(w + p.off_counter); }
};

template <typename T>
static void fill_common(ConvArgs<T> &a, const pnode_convblock_desc *d, const CbPlan &p, const Bufs<T> &b) {
    a.H = d->H, a.W = d->W, a.HW = d->H * d->W, a.cs = p.cs;
    a.npg = p.npg;
    a.tiles = (int)((p.npg + CB_PGX - 1) / CB_PGX);
    a.M = (double)p.M;
    a.partial = b.stat();
    a.counter = b.counter();
}

// z_1 .. z_L of the chain (and the BatchNorm coefficients / running statistics of every layer)
template <typename T>
static int forward_chain(const pnode_convblock_desc *d, const CbPlan &p, const Bufs<T> &b, const T *x, cudaStream_t st) {
    for (int k = 0; k < p.L; ++k) {
        const pnode_conv_layer &l = d->layer[k];
        ConvArgs<T> a = {};
        fill_common(a, d, p, b);
        a.in = k == 0 ? x : b.z(k - 1);
        a.coef_in = k == 0 ? nullptr : b.coef(k - 1);
        a.w = static_cast<const T *>(l.d_weight), a.bias = static_cast<const T *>(l.d_bias);
        a.out = b.z(k);
        a.coef_out = b.coef(k);
        a.gamma = static_cast<const T *>(l.d_gamma), a.beta = static_cast<const T *>(l.d_beta);
        a.rmean = static_cast<T *>(l.d_running_mean), a.rvar = static_cast<T *>(l.d_running_var);
        a.nbt = static_cast<long long *>(l.d_num_batches_tracked);
        a.eps = l.eps, a.momentum = l.momentum;
        a.CA = l.cin, a.CB = l.cout;
        const ConvGrid g = conv_grid(l.cin, l.cout, p.taps[k], p.npg, p.esz);
        int rc = k == 0 ? launch_conv<T, SRC_RAW, EPI_FWD>(p.kind[k], a, g, st) : launch_conv<T, SRC_ACT, EPI_FWD>(p.kind[k], a, g, st);
        if (rc) return rc;
    }
    PNODE_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
static int act_out(const pnode_convblock_desc *d, const CbPlan &p, const Bufs<T> &b, T *out, const T *base, double base_coef,
                   double kcoef, T *kout, cudaStream_t st) {
    const int C = d->layer[p.L - 1].cout;
    const int64_t nvec = p.M * C / 4;
    int64_t blocks = (nvec + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    act_out_kernel<T><<<(int)blocks, 256, 0, st>>>(b.z(p.L - 1), b.coef(p.L - 1), p.cs, C, d->H * d->W, nvec, out, base,
                                                    (T)base_coef, (T)kcoef, kout);
    PNODE_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
static int vjp_chain(const pnode_convblock_desc *d, const CbPlan &p, const Bufs<T> &b, const T *x, const T *w, T *vu, T *gout,
                     double coef, int accumulate, cudaStream_t st) {
    const int L = p.L;
    {  // BatchNorm-backward sums of the top layer
        const pnode_conv_layer &l = d->layer[L - 1];
        ConvArgs<T> a = {};
        fill_common(a, d, p, b);
        a.in = w, a.zprev = b.z(L - 1), a.coef_out = b.coef(L - 1);
        a.CA = 0, a.CB = l.cout;
        const ConvGrid g = conv_grid(0, l.cout, 1, p.npg, p.esz);
        dim3 grid(g.gx, g.gy), block(CB_PGX, g.cg);
        if (g.rc == 8) top_stats_kernel<T, 8><<<grid, block, 0, st>>>(a);
        else top_stats_kernel<T, 4><<<grid, block, 0, st>>>(a);
    }
    const T *gk = w;
    for (int k = L - 1; k >= 0; --k) {
        const pnode_conv_layer &l = d->layer[k];
        if (gout != nullptr) {
            WgradArgs<T> wa = {};
            wa.g = gk, wa.z = b.z(k), wa.coef_k = b.coef(k);
            wa.yin = k == 0 ? x : b.z(k - 1), wa.coef_in = k == 0 ? nullptr : b.coef(k - 1);
            wa.partial = b.wpart(k);
            wa.Cin = l.cin, wa.Cout = l.cout, wa.H = d->H, wa.W = d->W, wa.HW = d->H * d->W, wa.cs = p.cs;
            wa.tiles_cit = p.wg_tiles_cit[k];
            wa.npg = p.npg;
            wa.nchunks = (int)((p.npg + 31) / 32);
            int rc = k == 0 ? launch_wgrad<T, SRC_RAW>(p.kind[k], wa, p.wg_px[k], p.wg_ntiles[k], st)
                            : launch_wgrad<T, SRC_ACT>(p.kind[k], wa, p.wg_px[k], p.wg_ntiles[k], st);
            if (rc) return rc;
        }
        if (k == 0 && vu == nullptr) break;
        ConvArgs<T> a = {};
        fill_common(a, d, p, b);
        a.in = gk, a.in2 = b.z(k), a.coef_in = b.coef(k);
        a.w = static_cast<const T *>(l.d_weight);
        a.CA = l.cout, a.CB = l.cin;
        const ConvGrid g = conv_grid(l.cout, l.cin, p.taps[k], p.npg, p.esz);
        int rc;
        if (k == 0) {
            a.out = vu;
            rc = launch_conv<T, SRC_DZ, EPI_NONE>(p.kind[k], a, g, st);
        } else {
            a.out = b.g(k & 1);
            a.zprev = b.z(k - 1), a.coef_out = b.coef(k - 1);
            rc = launch_conv<T, SRC_DZ, EPI_DGRAD>(p.kind[k], a, g, st);
            gk = a.out;
        }
        if (rc) return rc;
    }
    if (gout != nullptr) {
        GradArgs<T> ga = {};
        ga.nl = L, ga.cs = p.cs, ga.accumulate = accumulate, ga.np = p.pbase[L], ga.out = gout, ga.coef = coef;
        for (int k = 0; k < L; ++k) {
            ga.L[k].partial = b.wpart(k), ga.L[k].coef = b.coef(k);
            ga.L[k].PX = p.wg_px[k], ga.L[k].ntiles = p.wg_ntiles[k], ga.L[k].tiles_cit = p.wg_tiles_cit[k];
            ga.L[k].Cin = d->layer[k].cin, ga.L[k].Cout = d->layer[k].cout, ga.L[k].taps = p.taps[k];
            ga.L[k].base = p.pbase[k];
        }
        int pxmax = 1;
        for (int k = 0; k < L; ++k) pxmax = p.wg_px[k] > pxmax ? p.wg_px[k] : pxmax;
        const int G = pxmax > 8 ? 32 : 4;
        int64_t blocks = (ga.np * G + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        if (G == 32) grads_finalize_kernel<T, 32><<<(int)blocks, 256, 0, st>>>(ga);
        else grads_finalize_kernel<T, 4><<<(int)blocks, 256, 0, st>>>(ga);
    }
    PNODE_CUDA_OK(cudaGetLastError());
    return 0;
}

static bool cb_aligned(const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace pnode

using namespace pnode;

extern "C" {

int64_t pnode_convblock_work_bytes(const pnode_convblock_desc *desc) {
    CbPlan p;
    if (cb_plan(desc, p) != 0) return -1;
    return (int64_t)p.total;
}

int64_t pnode_convblock_param_count(const pnode_convblock_desc *desc) {
    CbPlan p;
    if (cb_plan(desc, p) != 0) return -1;
    return p.pbase[p.L];
}

int pnode_convblock_forward(const pnode_convblock_desc *desc, const void *d_x, void *d_out, const void *d_base,
                            double base_coef, double k_coef, void *d_k, void *d_work, void *stream) {
    CbPlan p;
    int rc = cb_plan(desc, p);
    if (rc) return rc;
    PNODE_REQUIRE(d_x && d_work && (d_out || d_k), "pnode_convblock_forward: null argument");
    PNODE_REQUIRE(cb_aligned(d_x) && cb_aligned(d_out) && cb_aligned(d_base) && cb_aligned(d_k) && cb_aligned(d_work),
                  "pnode_convblock_forward: tensors must be 16-byte aligned");
    PNODE_REQUIRE(desc->layer[p.L - 1].cout == desc->layer[0].cin, "pnode_convblock_forward: an ODE right-hand side maps C -> C");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (desc->dtype == PNODE_F32) {
        Bufs<float> b(d_work, p);
        rc = forward_chain<float>(desc, p, b, static_cast<const float *>(d_x), st);
        if (rc) return rc;
        return act_out<float>(desc, p, b, static_cast<float *>(d_out), static_cast<const float *>(d_base), base_coef, k_coef,
                              static_cast<float *>(d_k), st);
    }
    Bufs<double> b(d_work, p);
    rc = forward_chain<double>(desc, p, b, static_cast<const double *>(d_x), st);
    if (rc) return rc;
    return act_out<double>(desc, p, b, static_cast<double *>(d_out), static_cast<const double *>(d_base), base_coef, k_coef,
                           static_cast<double *>(d_k), st);
}

int pnode_convblock_vjp(const pnode_convblock_desc *desc, const void *d_x, const void *d_w, void *d_vu, void *d_grads,
                        double coef, int accumulate, void *d_work, void *stream) {
    CbPlan p;
    int rc = cb_plan(desc, p);
    if (rc) return rc;
    PNODE_REQUIRE(d_x && d_w && d_work && (d_vu || d_grads), "pnode_convblock_vjp: null argument");
    PNODE_REQUIRE(cb_aligned(d_x) && cb_aligned(d_w) && cb_aligned(d_vu) && cb_aligned(d_work),
                  "pnode_convblock_vjp: tensors must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (desc->dtype == PNODE_F32) {
        Bufs<float> b(d_work, p);
        rc = forward_chain<float>(desc, p, b, static_cast<const float *>(d_x), st);
        if (rc) return rc;
        return vjp_chain<float>(desc, p, b, static_cast<const float *>(d_x), static_cast<const float *>(d_w),
                                static_cast<float *>(d_vu), static_cast<float *>(d_grads), coef, accumulate, st);
    }
    Bufs<double> b(d_work, p);
    rc = forward_chain<double>(desc, p, b, static_cast<const double *>(d_x), st);
    if (rc) return rc;
    return vjp_chain<double>(desc, p, b, static_cast<const double *>(d_x), static_cast<const double *>(d_w),
                             static_cast<double *>(d_vu), static_cast<double *>(d_grads), coef, accumulate, st);
}

}  // extern "C"
