// Declarations for the layer-wise training backward (gemm.cu, backward.cu).
#pragma once
#include "common.cuh"

namespace cneus {

enum { GEMM_NT = 0, GEMM_NN = 1, GEMM_TN = 2 };

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  const float* bias;   // [N] added to every row (nullable)
  const float* mask;   // C = (mask > 0) ? C : 0 (nullable), row-major with ldmask
  int M, N;
  int64_t K;
  int lda, ldb, ldc, ldmask;
  float alpha;
  int accumulate;      // C += result (ignored for split-K partials)
  int relu;
  int64_t split_stride;
};

int launch_gemm(int mode, const GemmArgs& g, cudaStream_t st);
int launch_gemm_tn_splitk(GemmArgs g, float* partial, int splits, cudaStream_t st);

}  // namespace cneus
