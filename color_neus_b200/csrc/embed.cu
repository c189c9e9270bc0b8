// Stand-alone positional encoding: get_embedder / Embedder.embed
// (lib/models/tools/PositionEncoding.py:45-94) -- out[p] = [x | sin(2^0 x) | cos(2^0 x) | ... | sin(2^(L-1) x) | cos(2^(L-1) x)],
// every block d wide.  Inside the renderer the encoding never exists in memory (it is computed in the point-shading kernels'
// prologue); this entry point serves user code that calls the embedder directly.  HBM-bound: 4 d bytes in, 4 d (1 + 2 L) out.
#include "common.cuh"

namespace cneus {

__global__ void embed_kernel(const float* __restrict__ x, int64_t P, int d, int L, float* __restrict__ out) {
  const int od = d * (1 + 2 * L);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P * od; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / od;
    const int q = (int)(i % od);
    const int blk = q / d, c = q % d;
    const float v = x[p * d + c];
    float r = v;
    if (blk > 0) {
      const float f = (float)(1 << ((blk - 1) >> 1));   // log-sampled bands 2^0 .. 2^(L-1) (exact in fp32)
      r = ((blk - 1) & 1) ? cosf(v * f) : sinf(v * f);
    }
    out[i] = r;
  }
}

}  // namespace cneus

extern "C" int cneus_embed(const float* x, int64_t P, int32_t input_dims, int32_t multires, float* out, void* stream) {
  CNEUS_NVTX_RANGE();
  using namespace cneus;
  if (P <= 0) return CNEUS_OK;
  if (!x || !out || input_dims <= 0 || multires < 0 || multires > 24) { set_error("embed: bad argument"); return CNEUS_EINVAL; }
  const int64_t total = P * input_dims * (1 + 2 * multires);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  embed_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, P, input_dims, multires, out);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return CNEUS_OK;
}
