// Train-mode BatchNorm2d + ReLU, forward and backward, for the SqueezeNext ODE block of BASELINE config 4
// (/root/reference/examples-pnode/models/sqnxt_PETSc.py:70-121: five times  relu(bn(conv(x)))  on NCHW tensors).
//
// Why: on the reference's path (and on this repo's generic path) the block's RHS evaluation and its VJP spend 69 % of their
// GPU time in cuDNN's bn_fw_tr_1C11 / bn_bw_1C11 kernels, which launch ONE CTA PER CHANNEL -- 16 to 32 CTAs on a 148-SM
// B200, ~135 us per call for a 17-34 MB tensor that HBM3e streams in ~5 us (profiles/r1_launches_cfg4.csv).  These are
// HBM-bound streaming reductions: the kernels below split every channel over many CTAs (grid = C x chunks ~ 8 CTAs/SM),
// read 16 bytes per thread with fully coalesced HW-contiguous rows, reduce in double with a fixed two-level order
// (per-CTA partials, then every consumer CTA re-sums the partials of its channel in the same order: deterministic, no atomics),
// and fuse ReLU (forward) / the ReLU mask (backward) into the same passes.
//   forward : stats pass (read x) + apply pass (read x, write y; also saves mean / inv-std and updates the running statistics
//             exactly like nn.BatchNorm2d: momentum, unbiased running variance)
//   backward: reduce pass (read dy, x, y) + apply pass (read dy, x, y, write dx; writes dgamma, dbeta)
#include "common.cuh"

namespace pnode {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_CHUNKS = 256;

template <typename T>
struct BnVec;
template <>
struct BnVec<float> {
    typedef float4 type;
    static constexpr int N = 4;
};
template <>
struct BnVec<double> {
    typedef double2 type;
    static constexpr int N = 2;
};

__device__ __forceinline__ void block_sum2(double &a, double &b, double *sh) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        sh[2 * w] = a;
        sh[2 * w + 1] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sb = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            sa += sh[2 * i];
            sb += sh[2 * i + 1];
        }
        sh[0] = sa;
        sh[1] = sb;
    }
    __syncthreads();
    a = sh[0];
    b = sh[1];
    __syncthreads();
}

// images [n0, n1) of channel c, HW contiguous scalars each: loop helper
template <typename T, typename F>
__device__ __forceinline__ void for_each_vec(const int n0, const int n1, const int C, const int c, const int HW, F body) {
    constexpr int V = BnVec<T>::N;
    const int hwv = HW / V;  // host guarantees HW % V == 0 and 16-byte aligned bases
    for (int n = n0; n < n1; ++n) {
        const int64_t base = ((int64_t)n * C + c) * HW;
        for (int i = threadIdx.x; i < hwv; i += blockDim.x) body(base + (int64_t)i * V);
    }
}

// partial[(c * nchunk + chunk) * 2 + {0,1}] = (sum x, sum x^2) over the chunk's images
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const T *__restrict__ x, int N, int C, int HW, int nchunk,
                                                              double *__restrict__ partial) {
    constexpr int V = BnVec<T>::N;
    typedef typename BnVec<T>::type VT;
    __shared__ double sh[2 * BN_THREADS / 32];
    const int c = blockIdx.x, chunk = blockIdx.y;
    const int per = (N + nchunk - 1) / nchunk;
    const int n0 = chunk * per, n1 = min(N, n0 + per);
    double s = 0.0, q = 0.0;
    for_each_vec<T>(n0, n1, C, c, HW, [&](int64_t off) {
        VT v = *reinterpret_cast<const VT *>(x + off);
        T e[V];
        memcpy(e, &v, sizeof(VT));
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const double d = (double)e[k];
            s += d;
            q = fma(d, d, q);
        }
    });
    block_sum2(s, q, sh);
    if (threadIdx.x == 0) {
        partial[((int64_t)c * nchunk + chunk) * 2] = s;
        partial[((int64_t)c * nchunk + chunk) * 2 + 1] = q;
    }
}

// y = relu(gamma (x - mean) * invstd + beta); chunk 0 of every channel also records mean / invstd and updates running stats
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_relu_apply_kernel(const T *__restrict__ x, T *__restrict__ y, const T *__restrict__ gamma, const T *__restrict__ beta,
                     const double *__restrict__ partial, int N, int C, int HW, int nchunk, double eps, double momentum,
                     T *__restrict__ running_mean, T *__restrict__ running_var, T *__restrict__ save_mean,
                     T *__restrict__ save_invstd) {
    constexpr int V = BnVec<T>::N;
    typedef typename BnVec<T>::type VT;
    const int c = blockIdx.x, chunk = blockIdx.y;
    double s = 0.0, q = 0.0;
    for (int k = 0; k < nchunk; ++k) {  // same order in every CTA of the channel
        s += partial[((int64_t)c * nchunk + k) * 2];
        q += partial[((int64_t)c * nchunk + k) * 2 + 1];
    }
    const double M = (double)N * HW;
    const double mean = s / M;
    const double var = fmax(q / M - mean * mean, 0.0);
    const double invstd = rsqrt(var + eps);
    if (chunk == 0 && threadIdx.x == 0) {
        save_mean[c] = (T)mean;
        save_invstd[c] = (T)invstd;
        if (running_mean != nullptr) {
            const double unbiased = M > 1.0 ? var * M / (M - 1.0) : var;
            running_mean[c] = (T)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
            running_var[c] = (T)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
        }
    }
    const T a = (T)((double)gamma[c] * invstd);
    const T b = (T)((double)beta[c] - (double)gamma[c] * invstd * mean);
    const int per = (N + nchunk - 1) / nchunk;
    const int n0 = chunk * per, n1 = min(N, n0 + per);
    for_each_vec<T>(n0, n1, C, c, HW, [&](int64_t off) {
        VT v = *reinterpret_cast<const VT *>(x + off);
        T e[V];
        memcpy(e, &v, sizeof(VT));
#pragma unroll
        for (int k = 0; k < V; ++k) e[k] = fmax(fma(a, e[k], b), T(0));
        memcpy(&v, e, sizeof(VT));
        *reinterpret_cast<VT *>(y + off) = v;
    });
}

// partial[(c*nchunk+chunk)*2 + {0,1}] = (sum dz, sum dz * xhat), dz = dy * [y > 0], xhat = (x - mean) invstd
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_relu_bwd_reduce_kernel(const T *__restrict__ dy, const T *__restrict__ x, const T *__restrict__ y,
                          const T *__restrict__ save_mean, const T *__restrict__ save_invstd, int N, int C, int HW,
                          int nchunk, double *__restrict__ partial) {
    constexpr int V = BnVec<T>::N;
    typedef typename BnVec<T>::type VT;
    __shared__ double sh[2 * BN_THREADS / 32];
    const int c = blockIdx.x, chunk = blockIdx.y;
    const int per = (N + nchunk - 1) / nchunk;
    const int n0 = chunk * per, n1 = min(N, n0 + per);
    const T mean = save_mean[c], invstd = save_invstd[c];
    double s = 0.0, q = 0.0;
    for_each_vec<T>(n0, n1, C, c, HW, [&](int64_t off) {
        VT vd = *reinterpret_cast<const VT *>(dy + off), vx = *reinterpret_cast<const VT *>(x + off),
           vy = *reinterpret_cast<const VT *>(y + off);
        T d[V], xx[V], yy[V];
        memcpy(d, &vd, sizeof(VT));
        memcpy(xx, &vx, sizeof(VT));
        memcpy(yy, &vy, sizeof(VT));
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const double dz = yy[k] > T(0) ? (double)d[k] : 0.0;
            s += dz;
            q = fma(dz, (double)((xx[k] - mean) * invstd), q);
        }
    });
    block_sum2(s, q, sh);
    if (threadIdx.x == 0) {
        partial[((int64_t)c * nchunk + chunk) * 2] = s;
        partial[((int64_t)c * nchunk + chunk) * 2 + 1] = q;
    }
}

// dx = gamma invstd (dz - sum(dz)/M - xhat sum(dz xhat)/M);  dgamma = sum(dz xhat), dbeta = sum(dz)
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_relu_bwd_apply_kernel(const T *__restrict__ dy, const T *__restrict__ x, const T *__restrict__ y,
                         const T *__restrict__ gamma, const T *__restrict__ save_mean,
                         const T *__restrict__ save_invstd, const double *__restrict__ partial, int N, int C, int HW,
                         int nchunk, T *__restrict__ dx, T *__restrict__ dgamma, T *__restrict__ dbeta) {
    constexpr int V = BnVec<T>::N;
    typedef typename BnVec<T>::type VT;
    const int c = blockIdx.x, chunk = blockIdx.y;
    double s = 0.0, q = 0.0;
    for (int k = 0; k < nchunk; ++k) {
        s += partial[((int64_t)c * nchunk + k) * 2];
        q += partial[((int64_t)c * nchunk + k) * 2 + 1];
    }
    if (chunk == 0 && threadIdx.x == 0) {
        dbeta[c] = (T)s;
        dgamma[c] = (T)q;
    }
    const double M = (double)N * HW;
    const T mean = save_mean[c], invstd = save_invstd[c];
    const T g = (T)((double)gamma[c] * (double)invstd);
    const T ms = (T)(s / M), mq = (T)(q / M);
    const int per = (N + nchunk - 1) / nchunk;
    const int n0 = chunk * per, n1 = min(N, n0 + per);
    for_each_vec<T>(n0, n1, C, c, HW, [&](int64_t off) {
        VT vd = *reinterpret_cast<const VT *>(dy + off), vx = *reinterpret_cast<const VT *>(x + off),
           vy = *reinterpret_cast<const VT *>(y + off);
        T d[V], xx[V], yy[V];
        memcpy(d, &vd, sizeof(VT));
        memcpy(xx, &vx, sizeof(VT));
        memcpy(yy, &vy, sizeof(VT));
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const T dz = yy[k] > T(0) ? d[k] : T(0);
            const T xh = (xx[k] - mean) * invstd;
            d[k] = g * (dz - ms - xh * mq);
        }
        memcpy(&vd, d, sizeof(VT));
        *reinterpret_cast<VT *>(dx + off) = vd;
    });
}

static int bn_chunks(int N, int C) {
    int want = (sm_count() * 8 + C - 1) / C;
    if (want > N) want = N;
    if (want > BN_MAX_CHUNKS) want = BN_MAX_CHUNKS;
    if (want < 1) want = 1;
    return want;
}

static bool bn_layout_ok(int HW, int dtype, const void *a, const void *b, const void *c, const void *d) {
    const int V = dtype == PNODE_F32 ? 4 : 2;
    auto al = [](const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    return HW % V == 0 && al(a) && al(b) && al(c) && al(d);
}

}  // namespace pnode

using namespace pnode;

extern "C" {

int64_t pnode_bn_work_bytes(int channels) { return (int64_t)channels * BN_MAX_CHUNKS * 2 * (int64_t)sizeof(double); }

int pnode_bn_relu_forward(const void *d_x, void *d_y, const void *d_gamma, const void *d_beta, void *d_running_mean,
                          void *d_running_var, void *d_save_mean, void *d_save_invstd, int N, int C, int HW, double eps,
                          double momentum, void *d_work, int dtype, void *stream) {
    PNODE_REQUIRE(d_x && d_y && d_gamma && d_beta && d_save_mean && d_save_invstd && d_work, "pnode_bn_relu_forward: null argument");
    PNODE_REQUIRE(bn_layout_ok(HW, dtype, d_x, d_y, nullptr, nullptr),
                  "pnode_bn_relu_forward: H*W must be a multiple of 16 bytes and tensors 16-byte aligned (HW=%d)", HW);
    if (N == 0 || C == 0 || HW == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nchunk = bn_chunks(N, C);
    dim3 grid(C, nchunk);
    double *partial = static_cast<double *>(d_work);
    if (dtype == PNODE_F32) {
        bn_stats_kernel<float><<<grid, BN_THREADS, 0, st>>>(static_cast<const float *>(d_x), N, C, HW, nchunk, partial);
        bn_relu_apply_kernel<float><<<grid, BN_THREADS, 0, st>>>(
            static_cast<const float *>(d_x), static_cast<float *>(d_y), static_cast<const float *>(d_gamma),
            static_cast<const float *>(d_beta), partial, N, C, HW, nchunk, eps, momentum,
            static_cast<float *>(d_running_mean), static_cast<float *>(d_running_var), static_cast<float *>(d_save_mean),
            static_cast<float *>(d_save_invstd));
    } else if (dtype == PNODE_F64) {
        bn_stats_kernel<double><<<grid, BN_THREADS, 0, st>>>(static_cast<const double *>(d_x), N, C, HW, nchunk, partial);
        bn_relu_apply_kernel<double><<<grid, BN_THREADS, 0, st>>>(
            static_cast<const double *>(d_x), static_cast<double *>(d_y), static_cast<const double *>(d_gamma),
            static_cast<const double *>(d_beta), partial, N, C, HW, nchunk, eps, momentum,
            static_cast<double *>(d_running_mean), static_cast<double *>(d_running_var),
            static_cast<double *>(d_save_mean), static_cast<double *>(d_save_invstd));
    } else {
        PNODE_REQUIRE(false, "pnode_bn_relu_forward: unsupported dtype %d", dtype);
    }
    PNODE_CUDA_OK(cudaGetLastError());
    return 0;
}

int pnode_bn_relu_backward(const void *d_dy, const void *d_x, const void *d_y, const void *d_gamma,
                           const void *d_save_mean, const void *d_save_invstd, void *d_dx, void *d_dgamma,
                           void *d_dbeta, int N, int C, int HW, void *d_work, int dtype, void *stream) {
    PNODE_REQUIRE(d_dy && d_x && d_y && d_gamma && d_save_mean && d_save_invstd && d_dx && d_dgamma && d_dbeta && d_work,
                  "pnode_bn_relu_backward: null argument");
    PNODE_REQUIRE(bn_layout_ok(HW, dtype, d_dy, d_x, d_y, d_dx),
                  "pnode_bn_relu_backward: H*W must be a multiple of 16 bytes and tensors 16-byte aligned (HW=%d)", HW);
    if (N == 0 || C == 0 || HW == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nchunk = bn_chunks(N, C);
    dim3 grid(C, nchunk);
    double *partial = static_cast<double *>(d_work);
    if (dtype == PNODE_F32) {
        typedef float T;
        bn_relu_bwd_reduce_kernel<T><<<grid, BN_THREADS, 0, st>>>(
            static_cast<const T *>(d_dy), static_cast<const T *>(d_x), static_cast<const T *>(d_y),
            static_cast<const T *>(d_save_mean), static_cast<const T *>(d_save_invstd), N, C, HW, nchunk, partial);
        bn_relu_bwd_apply_kernel<T><<<grid, BN_THREADS, 0, st>>>(
            static_cast<const T *>(d_dy), static_cast<const T *>(d_x), static_cast<const T *>(d_y),
            static_cast<const T *>(d_gamma), static_cast<const T *>(d_save_mean), static_cast<const T *>(d_save_invstd),
            partial, N, C, HW, nchunk, static_cast<T *>(d_dx), static_cast<T *>(d_dgamma), static_cast<T *>(d_dbeta));
    } else if (dtype == PNODE_F64) {
        typedef double T;
        bn_relu_bwd_reduce_kernel<T><<<grid, BN_THREADS, 0, st>>>(
            static_cast<const T *>(d_dy), static_cast<const T *>(d_x), static_cast<const T *>(d_y),
            static_cast<const T *>(d_save_mean), static_cast<const T *>(d_save_invstd), N, C, HW, nchunk, partial);
        bn_relu_bwd_apply_kernel<T><<<grid, BN_THREADS, 0, st>>>(
            static_cast<const T *>(d_dy), static_cast<const T *>(d_x), static_cast<const T *>(d_y),
            static_cast<const T *>(d_gamma), static_cast<const T *>(d_save_mean), static_cast<const T *>(d_save_invstd),
            partial, N, C, HW, nchunk, static_cast<T *>(d_dx), static_cast<T *>(d_dgamma), static_cast<T *>(d_dbeta));
    } else {
        PNODE_REQUIRE(false, "pnode_bn_relu_backward: unsupported dtype %d", dtype);
    }
    PNODE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
