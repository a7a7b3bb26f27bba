// Host-side interface of csrc/umma_gemm.cu for the other translation units (sliced operands + tensor-core products).
#pragma once
#include "common.cuh"
#include "pdl.cuh"

namespace pnode {
namespace umma {

constexpr int KIND_I8 = PNODE_SLICED_I8, KIND_TF32 = PNODE_SLICED_TF32, KIND_I8X = PNODE_SLICED_I8X;

inline int slices_of(int kind) { return kind == KIND_I8 ? PNODE_I8_SLICES : (kind == KIND_I8X ? PNODE_I8X_SLICES : 2); }
inline long long pitch_bytes(int kind, int k) {  // bytes of one operand row: k elements rounded up to 128 bytes
    const long long b = (long long)k * (kind == KIND_TF32 ? 4 : 1);
    return (b + 127) / 128 * 128;
}
inline long long sliced_bytes(int kind, int rows, int k) { return (long long)slices_of(kind) * rows * pitch_bytes(kind, k); }

// Epilogue of one product.  Zero-initialise, then set what is needed.
struct Epilogue {
    void *C;
    long long ldc;
    const int *ea, *eb;   // row exponents of A / of B (int8 kinds only)
    double alpha;
    const void *bias;     // [N] added before alpha (NULL: none)
    const void *mask;     // [M][ldmask]: result kept where mask_a[n] * mask + mask_b[n] > 0 (NULL arrays: 1 / 0), else 0
    long long ldmask;
    int relu, accumulate;
    const void *mask_a, *mask_b;
    // column statistics of the stored tile, per CTA row block: stats[blockIdx.y][n][2] = {sum x, sum x * q},
    // stat_mode 1: q = x (sum of squares);  2: q = q_a[n] * mask + q_b[n];  0: none
    double *stats;
    const void *q_a, *q_b;
    int stat_mode;
    // split-K: kb_per_split > 0 cuts the reduction into grid.z slices of that many K blocks (64 bytes of int8 / 32 floats);
    // slice z writes C + z * split_stride (elements)
    int kb_per_split;
    long long split_stride;
};
int gemm_ex(int kind, const void *a, const void *b, int M, int N, int K, const Epilogue &ep, cudaStream_t stream);
int split_k_blocks(int kind, int K, int splits);  // K blocks per slice for (about) `splits` slices

// C[m][n] (+)= mask(relu(alpha * (sum_k A[m][k] B[n][k] + bias[n])))
int gemm(int kind, const void *a, const int *ea, const void *b, const int *eb, int M, int N, int K, void *c, long long ldc,
         double alpha, const void *bias, int relu, const void *mask, long long ldmask, int accumulate, cudaStream_t stream);
int slice_rows(int kind, const void *x, long long ldx, int rows, int k, void *out, int *exps, cudaStream_t stream);
int slice_cols(int kind, const void *x, long long ldx, int rows, int cols, void *out, int *exps, void *colsum, double coef,
               cudaStream_t stream);

// Both slicings of one source in one pass: rows (operand = x) and columns (operand = x^T).  colpart: scratch of
// ceil(rows / 32) * cols doubles, needed when colsum != NULL (colsum += coef * column sums, fixed summation order).
int slice_both(int kind, const void *x, long long ldx, int rows, int cols, void *out_r, int *exp_r, void *out_c, int *exp_c,
               void *colsum, double coef, double *colpart, cudaStream_t stream);

using pdl::launch_pdl;
using pdl::pdl_launch_dependents;
using pdl::pdl_wait;

}  // namespace umma
}  // namespace pnode
