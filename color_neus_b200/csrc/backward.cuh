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
// C (+)= sum_z partial[z] ([splits][M][N], fixed order), honouring g.accumulate / g.ldc
int splitk_reduce(const float* partial, int splits, const GemmArgs& g, cudaStream_t st);

// tensor-core versions (gemm_tc.cu): fp16 hi/lo 3-pass tcgen05 with per-block power-of-two scaling
size_t tc_gemm_ws_floats();
size_t tc_gemm_tn_partial_floats();
bool tc_gemm_supported(int mode, const GemmArgs& g);      // NT / NN
bool tc_gemm_tn_supported(const GemmArgs& g);
int launch_gemm_tc(int mode, const GemmArgs& g, float* ws, cudaStream_t st);
int launch_gemm_tn_tc(const GemmArgs& g, float* ws, float* partial, cudaStream_t st);
int tensor_amax(const float* a, int64_t lda, int64_t rows, int ncols, float* out, cudaStream_t st);  // max |a| -> out[0]
// narrow products on CUDA cores (memory-bound): see gemm_tc.cu
int launch_small_tn(const float* A, int64_t lda, int M, const float* B, int64_t ldb, int J, int64_t K, float* C, int64_t cs_m, int64_t cs_j,
                    int accumulate, float* partial, cudaStream_t st);
int launch_small_k_nn(const GemmArgs& g, cudaStream_t st);  // NN with K <= 4, no bias, alpha = 1
int launch_small_nt(const float* A, int64_t lda, int64_t M, int K, const float* B, int64_t bs_j, int64_t bs_k, int J, const float* bias,
                    float* C, int64_t ldc, int accumulate, cudaStream_t st);

}  // namespace cneus
