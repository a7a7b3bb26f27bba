// Fused sweeps for tiny-state MLP right-hand sides  f(t,y) = W2 tanh(W1 phi(y) + b1) + b2  (spiral model,
// /root/reference/examples-pnode/ode_demo_petsc.py:207-230) under fixed-step explicit Runge-Kutta.
//
// What it replaces in the reference: the whole TSSolve / TSAdjointSolve host loops of PETSc plus the Python callbacks
// evalRHSFunction (pnode/petsc_adjoint.py:393-412), RHSJacShell.multTranspose (52-82) and RHSJacPShell.multTranspose
// (341-363), i.e. >= 10 interpreter calls and >= 6 kernel launches per stage for a 2-float state.  Here ONE launch
// does every step and stage of the forward sweep and ONE launch does the whole discrete adjoint.
//
// B200 mapping
//  * forward: one trajectory per thread, weights (252 scalars) broadcast from shared memory, the s stage slopes in
//    registers, stage values Y_i streamed to HBM as [step][stage][dim][traj] (coalesced, 8 scalars per
//    trajectory-step for RK4) -- the "stage checkpoints in HBM" that replace -ts_trajectory_type memory.
//  * adjoint: one trajectory per thread for the lambda recurrence and the VJP w.r.t. the state; the parameter gradient
//    (a [batch x 5] x [batch x H] outer-product sum) is reduced without atomics or shuffles: each warp parks tanh(z_j)
//    and s_j of its 32 trajectories in a warp-private shared-memory tile, then re-reads the tile TRANSPOSED (lane = hidden
//    unit j) so that every lane accumulates "its" five parameter gradients over the 32 trajectories in registers.  Only
//    __syncwarp() separates the two phases, warps never wait for each other, and the per-lane accumulators live across
//    all stages, steps and tiles of the persistent kernel.  Per-block partials are combined in a fixed order by the last
//    block (bit-reproducible mu).
//  * the kernels are bound by the FP64 / FP32+MUFU instruction issue rate of tanh (arithmetic intensity ~80 flop/B,
//    SURVEY.md section 8d), not by HBM: grids are persistent, sized SMs x resident CTAs.
#include "common.cuh"

namespace pnode {

// ---------------------------------------------------------------------------------------------------------------------
// tanh, accurate to a few ulp in fp64 and ~1.5e-7 absolute in fp32, branch-free

__device__ __forceinline__ float tanh_acc(float x) {
    // 1 - 2/(1 + e^{2x}):  one MUFU.EX2 + one MUFU.RCP
    float e, r;
    float a = x * 2.8853900817779268f;  // 2*log2(e)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
    float d = 1.0f + e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ double tanh_acc(double x) {
    // tanh|x| = em1 / (em1 + 2),  em1 = e^{2|x|} - 1 evaluated without cancellation
    double ax = fmin(fabs(x), 20.0);  // tanh(20) rounds to 1
    double y = ax + ax;
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: round-to-nearest-integer trick
    double nf = fma(y, 1.4426950408889634, MAGIC);
    int n = __double2loint(nf);
    nf -= MAGIC;
    double r = fma(nf, -6.93147180369123816490e-01, y);
    r = fma(nf, -1.90821492927058770002e-10, r);
    // q(r) = (e^r - 1)/r, |r| <= ln2/2, Taylor through r^12 (truncation 4e-18)
    double q = 1.6059043836821613e-10;   // 1/13!
    q = fma(q, r, 2.08767569878681e-09);   // 1/12!
    q = fma(q, r, 2.505210838544172e-08);  // 1/11!
    q = fma(q, r, 2.755731922398589e-07);  // 1/10!
    q = fma(q, r, 2.7557319223985893e-06); // 1/9!
    q = fma(q, r, 2.48015873015873e-05);   // 1/8!
    q = fma(q, r, 1.984126984126984e-04);  // 1/7!
    q = fma(q, r, 1.388888888888889e-03);  // 1/6!
    q = fma(q, r, 8.333333333333333e-03);  // 1/5!
    q = fma(q, r, 4.1666666666666664e-02); // 1/4!
    q = fma(q, r, 1.6666666666666666e-01); // 1/3!
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    double rq = r * q;  // e^r - 1
    double p = 1.0 + rq;
    double e = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));  // p * 2^n, 0 <= n <= 58
    double em1 = (n == 0) ? rq : e - 1.0;
    double d = em1 + 2.0;
    double rc;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(d));
    rc = fma(fma(-d, rc, 1.0), rc, rc);
    rc = fma(fma(-d, rc, 1.0), rc, rc);
    double t = em1 * rc;
    t = fma(fma(-d, t, em1), rc, t);
    return copysign(t, x);
}

// ---------------------------------------------------------------------------------------------------------------------

template <typename T>
struct MlpPtrs {
    const T *w1, *b1, *w2, *b2;
};

// per-hidden-unit weights packed for vector LDS: w1[j][0..D), b1[j], w2[0..D)[j]
template <typename T, int D>
struct alignas(2 * sizeof(T)) Unit {
    T w1[D];
    T b1;
    T w2[D];
    T pad[(2 * D + 1) % 2];
};

template <typename T, int D, int H>
__device__ __forceinline__ void load_weights(Unit<T, D> *sW, T *sB2, const MlpPtrs<T> &w) {
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
}

template <typename T, int D, int PHI>
__device__ __forceinline__ void apply_phi(const T (&y)[D], T (&x)[D]) {
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = (PHI == 1) ? y[d] * y[d] * y[d] : y[d];
}

template <typename T, int D, int H, int PHI>
__device__ __forceinline__ void mlp_eval(const Unit<T, D> *__restrict__ sW, const T *__restrict__ sB2,
                                         const T (&y)[D], T (&out)[D]) {
    T x[D];
    apply_phi<T, D, PHI>(y, x);
#pragma unroll
    for (int d = 0; d < D; ++d) out[d] = sB2[d];
#pragma unroll 5
    for (int j = 0; j < H; ++j) {
        const Unit<T, D> u = sW[j];
        T z = u.b1;
#pragma unroll
        for (int d = 0; d < D; ++d) z = fma(u.w1[d], x[d], z);
        const T a = tanh_acc(z);
#pragma unroll
        for (int d = 0; d < D; ++d) out[d] = fma(u.w2[d], a, out[d]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward sweep

constexpr int FWD_THREADS = 128;

template <typename T, int D, int H, int S, int PHI>
__global__ void __launch_bounds__(FWD_THREADS)
mlp_rk_fwd_kernel(const MlpPtrs<T> w, const pnode_rk_tableau tab, const T *__restrict__ u0, const int64_t ntraj,
                  const pnode_step *__restrict__ sched, const int nsteps, T *__restrict__ sol, T *__restrict__ ckpt) {
    __shared__ Unit<T, D> sW[H];
    __shared__ T sB2[D];
    load_weights<T, D, H>(sW, sB2, w);
    __syncthreads();

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t traj = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; traj < ntraj; traj += stride) {
        T y[D];
#pragma unroll
        for (int d = 0; d < D; ++d) y[d] = u0[traj * D + d];
        T K[S][D];
        for (int n = 0; n < nsteps; ++n) {
            const double h = sched[n].h;
            const int out_slot = sched[n].out_slot;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                T Y[D];
#pragma unroll
                for (int d = 0; d < D; ++d) Y[d] = y[d];
#pragma unroll
                for (int j = 0; j < i; ++j) {
                    const T ha = (T)(h * tab.a[i][j]);
#pragma unroll
                    for (int d = 0; d < D; ++d) Y[d] = fma(ha, K[j][d], Y[d]);
                }
                if (ckpt != nullptr) {
#pragma unroll
                    for (int d = 0; d < D; ++d) ckpt[(((int64_t)n * S + i) * D + d) * ntraj + traj] = Y[d];
                }
                if (i == 0 && tab.fsal && n > 0) {
#pragma unroll
                    for (int d = 0; d < D; ++d) K[0][d] = K[S - 1][d];
                } else {
                    mlp_eval<T, D, H, PHI>(sW, sB2, Y, K[i]);
                }
            }
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const T hb = (T)(h * tab.b[j]);
#pragma unroll
                for (int d = 0; d < D; ++d) y[d] = fma(hb, K[j][d], y[d]);
            }
            if (out_slot >= 0) {
#pragma unroll
                for (int d = 0; d < D; ++d) sol[((int64_t)out_slot * ntraj + traj) * D + d] = y[d];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// adjoint sweep

constexpr int ADJ_WARPS = 4;
constexpr int ADJ_THREADS = ADJ_WARPS * 32;
constexpr int ADJ_MAX_BLOCKS = 148 * 8;

template <int D, int H>
struct AdjShape {
    static constexpr int NCHUNK = (H + 31) / 32;
    static constexpr int JH = (H + NCHUNK - 1) / NCHUNK;  // hidden units per chunk (<= 32)
    static constexpr int PITCH = 33;                      // odd pitch: transposed re-read is bank-conflict free
    static constexpr int NP = 2 * H * D + H + D;
};

struct AdjWork {
    unsigned int ticket;
    unsigned int pad[15];
    double partial[1];  // [blocks][NP]
};

template <typename T, int D, int H>
struct WarpTile {
    T A[AdjShape<D, H>::JH * AdjShape<D, H>::PITCH];  // tanh(z_j) per (unit, trajectory)
    T Sg[AdjShape<D, H>::JH * AdjShape<D, H>::PITCH];  // s_j = g_j (1 - a_j^2)
    alignas(16) T VX[32 * 2 * D];                      // per trajectory: v[0..D), x[0..D)
};

template <typename T, int D, int H, int S, int PHI>
__global__ void __launch_bounds__(ADJ_THREADS)
mlp_rk_adj_kernel(const MlpPtrs<T> w, const pnode_rk_tableau tab, const int64_t ntraj,
                  const pnode_step *__restrict__ sched, const int nsteps, const int last_slot,
                  const T *__restrict__ gout, const T *__restrict__ ckpt, T *__restrict__ lambda_out,
                  T *__restrict__ mu_out, AdjWork *__restrict__ work) {
    typedef AdjShape<D, H> Sh;
    constexpr int NCHUNK = Sh::NCHUNK, JH = Sh::JH, PITCH = Sh::PITCH, NP = Sh::NP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Unit<T, D> *sW = reinterpret_cast<Unit<T, D> *>(smem_raw);
    T *sB2 = reinterpret_cast<T *>(sW + H);
    WarpTile<T, D, H> *tiles = reinterpret_cast<WarpTile<T, D, H> *>(
        smem_raw + ((sizeof(Unit<T, D>) * H + sizeof(T) * D + 15) / 16) * 16);
    __shared__ bool is_last;

    load_weights<T, D, H>(sW, sB2, w);
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

    const int64_t ntiles = (ntraj + ADJ_THREADS - 1) / ADJ_THREADS;
    for (int64_t tidx = blockIdx.x; tidx < ntiles; tidx += gridDim.x) {
        const int64_t traj = tidx * ADJ_THREADS + threadIdx.x;
        const bool valid = traj < ntraj;
        T lam[D];
#pragma unroll
        for (int d = 0; d < D; ++d) lam[d] = valid ? gout[((int64_t)last_slot * ntraj + traj) * D + d] : T(0);

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
                // cotangent of the stage slope, pre-multiplied by the step coefficient: v = c * w
                T v[D];
                const double bi = tab.b[i];
                if (bi != 0.0) {
#pragma unroll
                    for (int d = 0; d < D; ++d) v[d] = lam[d];
#pragma unroll
                    for (int j = i + 1; j < S; ++j) {
                        const T r = (T)(tab.a[j][i] / bi);
#pragma unroll
                        for (int d = 0; d < D; ++d) v[d] = fma(r, ls[j][d], v[d]);
                    }
                    const T c = (T)(h * bi);
#pragma unroll
                    for (int d = 0; d < D; ++d) v[d] *= c;
                } else {
#pragma unroll
                    for (int d = 0; d < D; ++d) v[d] = T(0);
#pragma unroll
                    for (int j = i + 1; j < S; ++j) {
                        const T r = (T)tab.a[j][i];
#pragma unroll
                        for (int d = 0; d < D; ++d) v[d] = fma(r, ls[j][d], v[d]);
                    }
                    const T c = (T)h;
#pragma unroll
                    for (int d = 0; d < D; ++d) v[d] *= c;
                }
                T Y[D], x[D];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    Y[d] = valid ? ckpt[(((int64_t)n * S + i) * D + d) * ntraj + traj] : T(0);
                    if (!valid) v[d] = T(0);
                }
                apply_phi<T, D, PHI>(Y, x);
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    tile.VX[lane * 2 * D + d] = v[d];
                    tile.VX[lane * 2 * D + D + d] = x[d];
                    accB2[d] += (double)v[d];
                }
                T dx[D];
#pragma unroll
                for (int d = 0; d < D; ++d) dx[d] = T(0);
#pragma unroll
                for (int c = 0; c < NCHUNK; ++c) {
                    const int j0 = c * JH;
                    const int jn = (H - j0 < JH) ? (H - j0) : JH;
                    // phase 1 (lane = trajectory): VJP through every hidden unit of the chunk
#pragma unroll 5
                    for (int jj = 0; jj < jn; ++jj) {
                        const Unit<T, D> u = sW[j0 + jj];
                        T z = u.b1;
#pragma unroll
                        for (int d = 0; d < D; ++d) z = fma(u.w1[d], x[d], z);
                        const T a = tanh_acc(z);
                        T g = T(0);
#pragma unroll
                        for (int d = 0; d < D; ++d) g = fma(u.w2[d], v[d], g);
                        const T s = g * fma(-a, a, T(1));
#pragma unroll
                        for (int d = 0; d < D; ++d) dx[d] = fma(s, u.w1[d], dx[d]);
                        tile.A[jj * PITCH + lane] = a;
                        tile.Sg[jj * PITCH + lane] = s;
                    }
                    __syncwarp();
                    // phase 2 (lane = hidden unit): reduce the outer products over the warp's 32 trajectories
                    if (lane < jn) {
                        T pW2[D], pW1[D], pB1 = T(0);
#pragma unroll
                        for (int d = 0; d < D; ++d) pW2[d] = pW1[d] = T(0);
#pragma unroll 8
                        for (int k = 0; k < 32; ++k) {
                            const T a = tile.A[lane * PITCH + k];
                            const T s = tile.Sg[lane * PITCH + k];
                            pB1 += s;
#pragma unroll
                            for (int d = 0; d < D; ++d) {
                                pW2[d] = fma(tile.VX[k * 2 * D + d], a, pW2[d]);
                                pW1[d] = fma(s, tile.VX[k * 2 * D + D + d], pW1[d]);
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
#pragma unroll
                for (int d = 0; d < D; ++d)
                    ls[i][d] = (PHI == 1) ? dx[d] * (T(3) * Y[d] * Y[d]) : dx[d];
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
        if (valid) {
#pragma unroll
            for (int d = 0; d < D; ++d) lambda_out[traj * D + d] = lam[d];
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
    __syncthreads();
    for (int p = threadIdx.x; p < NP; p += blockDim.x) {
        double s = 0.0;
#pragma unroll
        for (int wi = 0; wi < ADJ_WARPS; ++wi) s += blk[wi * NP + p];
        work->partial[(int64_t)blockIdx.x * NP + p] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(&work->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int p = threadIdx.x; p < NP; p += blockDim.x) {
            double s = 0.0;
            for (int b = 0; b < (int)gridDim.x; ++b) s += ((volatile double *)work->partial)[(int64_t)b * NP + p];
            mu_out[p] = (T)s;
        }
        if (threadIdx.x == 0) work->ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host dispatch

template <typename T, int D, int H>
static size_t adj_smem_bytes() {
    return ((sizeof(Unit<T, D>) * H + sizeof(T) * D + 15) / 16) * 16 + sizeof(WarpTile<T, D, H>) * ADJ_WARPS;
}

template <typename T, int D, int H, int S, int PHI>
static int launch_fwd(const pnode_mlp_desc *m, const pnode_rk_tableau *tab, const void *d_u0, int64_t ntraj,
                      const pnode_step *d_sched, int nsteps, void *d_sol, void *d_ckpt, cudaStream_t st) {
    MlpPtrs<T> w{static_cast<const T *>(m->d_w1), static_cast<const T *>(m->d_b1), static_cast<const T *>(m->d_w2),
                 static_cast<const T *>(m->d_b2)};
    auto kern = mlp_rk_fwd_kernel<T, D, H, S, PHI>;
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        PNODE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, FWD_THREADS, 0));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    int64_t want = (ntraj + FWD_THREADS - 1) / FWD_THREADS;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, FWD_THREADS, 0, st>>>(w, *tab, static_cast<const T *>(d_u0), ntraj, d_sched, nsteps,
                                       static_cast<T *>(d_sol), static_cast<T *>(d_ckpt));
    PNODE_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T, int D, int H, int S, int PHI>
static int launch_adj(const pnode_mlp_desc *m, const pnode_rk_tableau *tab, int64_t ntraj, const pnode_step *d_sched,
                      int nsteps, int last_slot, const void *d_gout, const void *d_ckpt, void *d_lambda, void *d_mu,
                      void *d_work, cudaStream_t st) {
    MlpPtrs<T> w{static_cast<const T *>(m->d_w1), static_cast<const T *>(m->d_b1), static_cast<const T *>(m->d_w2),
                 static_cast<const T *>(m->d_b2)};
    auto kern = mlp_rk_adj_kernel<T, D, H, S, PHI>;
    const size_t smem = adj_smem_bytes<T, D, H>();
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        PNODE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PNODE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, ADJ_THREADS, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    int64_t want = (ntraj + ADJ_THREADS - 1) / ADJ_THREADS;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (cap > ADJ_MAX_BLOCKS) cap = ADJ_MAX_BLOCKS;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, ADJ_THREADS, smem, st>>>(w, *tab, ntraj, d_sched, nsteps, last_slot, static_cast<const T *>(d_gout),
                                          static_cast<const T *>(d_ckpt), static_cast<T *>(d_lambda),
                                          static_cast<T *>(d_mu), static_cast<AdjWork *>(d_work));
    PNODE_CUDA_OK(cudaGetLastError());
    return 0;
}

// the compiled instantiations: (dim, hidden) = (2, 50) -- the spiral model -- for every explicit tableau PETSc's
// TSRK offers through pnode's method= table (1fe, 2a/2b, 3, 3bs/4, 5dp) and phi in {identity, cube}
#define PNODE_FOR_STAGES(X) X(1) X(2) X(3) X(4) X(7)

static bool shape_ok(int dim, int hidden, int phi, int stages) {
    bool s_ok = stages == 1 || stages == 2 || stages == 3 || stages == 4 || stages == 7;
    return dim == 2 && hidden == 50 && (phi == 0 || phi == 1) && s_ok;
}

template <typename T>
static int dispatch_fwd(const pnode_mlp_desc *m, const pnode_rk_tableau *tab, const void *d_u0, int64_t ntraj,
                        const pnode_step *d_sched, int nsteps, void *d_sol, void *d_ckpt, cudaStream_t st) {
#define X(SS)                                                                                                     \
    if (tab->s == SS) {                                                                                           \
        if (m->phi == 1)                                                                                          \
            return launch_fwd<T, 2, 50, SS, 1>(m, tab, d_u0, ntraj, d_sched, nsteps, d_sol, d_ckpt, st);          \
        return launch_fwd<T, 2, 50, SS, 0>(m, tab, d_u0, ntraj, d_sched, nsteps, d_sol, d_ckpt, st);              \
    }
    PNODE_FOR_STAGES(X)
#undef X
    PNODE_REQUIRE(false, "pnode_mlp_rk_forward: no kernel for %d stages", tab->s);
}

template <typename T>
static int dispatch_adj(const pnode_mlp_desc *m, const pnode_rk_tableau *tab, int64_t ntraj, const pnode_step *d_sched,
                        int nsteps, int last_slot, const void *d_gout, const void *d_ckpt, void *d_lambda, void *d_mu,
                        void *d_work, cudaStream_t st) {
#define X(SS)                                                                                                     \
    if (tab->s == SS) {                                                                                           \
        if (m->phi == 1)                                                                                          \
            return launch_adj<T, 2, 50, SS, 1>(m, tab, ntraj, d_sched, nsteps, last_slot, d_gout, d_ckpt,         \
                                               d_lambda, d_mu, d_work, st);                                       \
        return launch_adj<T, 2, 50, SS, 0>(m, tab, ntraj, d_sched, nsteps, last_slot, d_gout, d_ckpt, d_lambda,   \
                                           d_mu, d_work, st);                                                     \
    }
    PNODE_FOR_STAGES(X)
#undef X
    PNODE_REQUIRE(false, "pnode_mlp_rk_adjoint: no kernel for %d stages", tab->s);
}

template <typename T>
__global__ void tanh_probe_kernel(const T *in, T *out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = tanh_acc(in[i]);
}

}  // namespace pnode

using namespace pnode;

extern "C" {

int pnode_mlp_rk_supported(int dim, int hidden, int phi, int dtype, int stages) {
    return (dtype == PNODE_F32 || dtype == PNODE_F64) && shape_ok(dim, hidden, phi, stages) ? 1 : 0;
}

int pnode_mlp_rk_forward(const pnode_mlp_desc *mlp, const pnode_rk_tableau *tab, const void *d_u0, int64_t ntraj,
                         const pnode_step *d_sched, int nsteps, void *d_sol, void *d_ckpt, void *stream) {
    PNODE_REQUIRE(mlp && tab && d_sched, "pnode_mlp_rk_forward: null argument");
    PNODE_REQUIRE(shape_ok(mlp->dim, mlp->hidden, mlp->phi, tab->s),
                  "pnode_mlp_rk_forward: unsupported shape dim=%d hidden=%d phi=%d stages=%d", mlp->dim, mlp->hidden,
                  mlp->phi, tab->s);
    if (ntraj == 0 || nsteps == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mlp->dtype == PNODE_F32) return dispatch_fwd<float>(mlp, tab, d_u0, ntraj, d_sched, nsteps, d_sol, d_ckpt, st);
    if (mlp->dtype == PNODE_F64)
        return dispatch_fwd<double>(mlp, tab, d_u0, ntraj, d_sched, nsteps, d_sol, d_ckpt, st);
    PNODE_REQUIRE(false, "pnode_mlp_rk_forward: unsupported dtype %d", mlp->dtype);
}

int64_t pnode_mlp_rk_adjoint_work_bytes(const pnode_mlp_desc *mlp) {
    int64_t np = 2 * (int64_t)mlp->hidden * mlp->dim + mlp->hidden + mlp->dim;
    return 64 + (int64_t)ADJ_MAX_BLOCKS * np * (int64_t)sizeof(double);
}

int pnode_mlp_rk_adjoint(const pnode_mlp_desc *mlp, const pnode_rk_tableau *tab, int64_t ntraj,
                         const pnode_step *d_sched, int nsteps, int last_slot, const void *d_gout, const void *d_ckpt,
                         void *d_lambda, void *d_mu, void *d_work, void *stream) {
    PNODE_REQUIRE(mlp && tab && d_sched && d_work, "pnode_mlp_rk_adjoint: null argument");
    PNODE_REQUIRE(shape_ok(mlp->dim, mlp->hidden, mlp->phi, tab->s),
                  "pnode_mlp_rk_adjoint: unsupported shape dim=%d hidden=%d phi=%d stages=%d", mlp->dim, mlp->hidden,
                  mlp->phi, tab->s);
    PNODE_REQUIRE(d_ckpt != nullptr || nsteps == 0, "pnode_mlp_rk_adjoint: stage checkpoints missing");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mlp->dtype == PNODE_F32)
        return dispatch_adj<float>(mlp, tab, ntraj, d_sched, nsteps, last_slot, d_gout, d_ckpt, d_lambda, d_mu, d_work,
                                   st);
    if (mlp->dtype == PNODE_F64)
        return dispatch_adj<double>(mlp, tab, ntraj, d_sched, nsteps, last_slot, d_gout, d_ckpt, d_lambda, d_mu,
                                    d_work, st);
    PNODE_REQUIRE(false, "pnode_mlp_rk_adjoint: unsupported dtype %d", mlp->dtype);
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
