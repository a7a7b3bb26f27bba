// Host-side interface of csrc/umma_gemm.cu for the other translation units (sliced operands + tensor-core products).
#pragma once
#include "common.cuh"

namespace pnode {
namespace umma {

constexpr int KIND_I8 = PNODE_SLICED_I8, KIND_TF32 = PNODE_SLICED_TF32, KIND_I8X = PNODE_SLICED_I8X;

inline int slices_of(int kind) { return kind == KIND_I8 ? PNODE_I8_SLICES : (kind == KIND_I8X ? PNODE_I8X_SLICES : 2); }
inline long long pitch_bytes(int kind, int k) {  // bytes of one operand row: k elements rounded up to 128 bytes
    const long long b = (long long)k * (kind == KIND_TF32 ? 4 : 1);
    return (b + 127) / 128 * 128;
}
inline long long sliced_bytes(int kind, int rows, int k) { return (long long)slices_of(kind) * rows * pitch_bytes(kind, k); }

// C[m][n] (+)= mask(relu(alpha * (sum_k A[m][k] B[n][k] + bias[n])))
int gemm(int kind, const void *a, const int *ea, const void *b, const int *eb, int M, int N, int K, void *c, long long ldc,
         double alpha, const void *bias, int relu, const void *mask, long long ldmask, int accumulate, cudaStream_t stream);
int slice_rows(int kind, const void *x, long long ldx, int rows, int k, void *out, int *exps, cudaStream_t stream);
int slice_cols(int kind, const void *x, long long ldx, int rows, int cols, void *out, int *exps, void *colsum, double coef,
               cudaStream_t stream);

}  // namespace umma
}  // namespace pnode
