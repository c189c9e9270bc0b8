// Generic fp32 SGEMM on CUDA cores for the layer-wise training backward (backward.cu): row-major operands with
// arbitrary leading dimensions, three operand modes, fused bias / ReLU / mask epilogues, split-K for the weight
// gradients (reduction over all points of the batch).
//
//   NT: C[m,n] = sum_k A[m,k] * B[n,k]      (y = x W^T, the nn.Linear forward)
//   NN: C[m,n] = sum_k A[m,k] * B[k,n]      (dL/dx = dL/dy W)
//   TN: C[m,n] = sum_k A[k,m] * B[k,n]      (dL/dW = dL/dy^T x, k runs over points; split over blockIdx.z)
#include "backward.cuh"

namespace cneus {

constexpr int GBM = 128, GBN = 128, GBK = 8, GTHREADS = 256;

template <int MODE>
__global__ void __launch_bounds__(GTHREADS) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][GBK][GBM + 4];
  __shared__ __align__(16) float Bs[2][GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  // split-K range of this z-slice
  const int64_t kper = (g.K + gridDim.z - 1) / gridDim.z;
  const int64_t kbeg = (int64_t)blockIdx.z * kper;
  const int64_t kend = (kbeg + kper < g.K) ? kbeg + kper : g.K;
  const int tm = (tid & 15) * 8, tn = (tid >> 4) * 8;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // each thread stages 4 elements of A and 4 of B per k-tile
  float ra[4], rb[4];
  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * GTHREADS;  // 0..1023
      int kk, mm;
      if (MODE == GEMM_TN) { kk = e / GBM; mm = e % GBM; }        // A[k][m]: m contiguous
      else { mm = e / GBK; kk = e % GBK; }                          // A[m][k]: k contiguous
      const int64_t k = k0 + kk;
      const int m = m0 + mm;
      float v = 0.f;
      if (k < kend && m < g.M) v = (MODE == GEMM_TN) ? g.A[k * g.lda + m] : g.A[(int64_t)m * g.lda + k];
      ra[i] = v;
      int kb, nn;
      if (MODE == GEMM_NT) { nn = e / GBK; kb = e % GBK; }         // B[n][k]
      else { kb = e / GBN; nn = e % GBN; }                          // B[k][n]
      const int64_t k2 = k0 + kb;
      const int n = n0 + nn;
      float w = 0.f;
      if (k2 < kend && n < g.N) w = (MODE == GEMM_NT) ? g.B[(int64_t)n * g.ldb + k2] : g.B[k2 * g.ldb + n];
      rb[i] = w;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * GTHREADS;
      int kk, mm;
      if (MODE == GEMM_TN) { kk = e / GBM; mm = e % GBM; } else { mm = e / GBK; kk = e % GBK; }
      As[buf][kk][mm] = ra[i];
      int kb, nn;
      if (MODE == GEMM_NT) { nn = e / GBK; kb = e % GBK; } else { kb = e / GBN; nn = e % GBN; }
      Bs[buf][kb][nn] = rb[i];
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += GBK) {
    const bool more = k0 + GBK < kend;
    if (more) load_tiles(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tm]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][tm + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tn]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tn + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) store_tiles(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // epilogue
  float* C = g.C + (gridDim.z > 1 ? (int64_t)blockIdx.z * g.split_stride : 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + tm + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tn + j;
      if (n >= g.N) continue;
      float v = acc[i][j] * g.alpha;
      if (g.bias) v += g.bias[n];
      const int64_t ci = (int64_t)m * g.ldc + n;
      if (g.accumulate && gridDim.z == 1) v += C[ci];
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.mask) v = (g.mask[(int64_t)m * g.ldmask + n] > 0.f) ? v : 0.f;
      C[ci] = v;
    }
  }
}

// C[m,n] (+)= sum_z partial[z][m,n]
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t stride, int M, int N, int ldc,
                                     float* __restrict__ C, int accumulate) {
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), n = (int)(i % N);
    double s = 0.0;
    for (int z = 0; z < splits; ++z) s += (double)part[(int64_t)z * stride + (int64_t)m * N + n];
    const int64_t ci = (int64_t)m * ldc + n;
    C[ci] = (float)(accumulate ? (double)C[ci] + s : s);
  }
}

int launch_gemm(int mode, const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return CNEUS_OK;
  dim3 grid((g.N + GBN - 1) / GBN, (g.M + GBM - 1) / GBM, 1);
  if (mode == GEMM_NT) sgemm_kernel<GEMM_NT><<<grid, GTHREADS, 0, st>>>(g);
  else if (mode == GEMM_NN) sgemm_kernel<GEMM_NN><<<grid, GTHREADS, 0, st>>>(g);
  else sgemm_kernel<GEMM_TN><<<grid, GTHREADS, 0, st>>>(g);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return CNEUS_OK;
}

int splitk_reduce(const float* partial, int splits, const GemmArgs& g, cudaStream_t st) {
  const int64_t total = (int64_t)g.M * g.N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  splitk_reduce_kernel<<<blocks, 256, 0, st>>>(partial, splits, (int64_t)g.M * g.N, g.M, g.N, g.ldc, g.C, g.accumulate);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return CNEUS_OK;
}

// Weight gradient: C[M,N] (+)= A[K,M]^T B[K,N] with K = number of points; split over `splits` z-slices whose partial
// sums go to `partial` ([splits][M][N]) and are reduced in fixed order (deterministic).
int launch_gemm_tn_splitk(GemmArgs g, float* partial, int splits, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return CNEUS_OK;
  if (splits < 1) splits = 1;
  GemmArgs p = g;
  p.C = partial; p.ldc = g.N; p.split_stride = (int64_t)g.M * g.N; p.accumulate = 0;
  dim3 grid((g.N + GBN - 1) / GBN, (g.M + GBM - 1) / GBM, splits);
  if (splits == 1) p.split_stride = 0;
  sgemm_kernel<GEMM_TN><<<grid, GTHREADS, 0, st>>>(p);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return splitk_reduce(partial, splits, g, st);
}

}  // namespace cneus
