// NeuS_Trainer.compute_loss (lib/models/NeuS_Trainer.py:129-171) and the seeds of its backward
// in two launches: launch 1 = fp64 partial sums (rgb term, BCE term, masked sum of delta_relight) per block; launch 2 =
// every block re-adds the partials in fixed order (deterministic, no atomics), block 0 writes the five loss terms, all
// blocks write d loss / d {color_fine, weight_sum, delta_relight}.  d loss / d gradient_error = lambda_eikonal is a constant.
//   rgb      = mean((c - gt)^2)  or  mean|c - gt|                       (torch.nn.MSELoss / L1Loss, :71-74)
//   mask     = mean(-(m log p + (1 - m) log(1 - p))), p = clip(weight_sum, 1e-3, 1 - 1e-3)   (:141-143)
//   relight  = (mean(delta_relight * mask[:,None,None]))^2             (:146-155; mask only when INCLUDE_MASK)
//   loss     = l_fine rgb + l_eik gradient_error + l_mask mask + l_relight relight
#include "common.cuh"

namespace cneus {

constexpr int LOSS_BLOCKS = 148;
constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < LOSS_THREADS / 32; ++i) t += sm[i];
  return t;
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_partial_kernel(const float* __restrict__ color, const float* __restrict__ gt,
                                                                      const float* __restrict__ wsum, const float* __restrict__ mask,
                                                                      const float* __restrict__ delta, int64_t B, int32_t S,
                                                                      int32_t l1, int32_t use_bce, int32_t mask_relight,
                                                                      double* __restrict__ partial) {
  __shared__ double sm[LOSS_THREADS / 32];
  const int64_t tid = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x, nth = (int64_t)gridDim.x * LOSS_THREADS;
  double s_rgb = 0.0, s_bce = 0.0, s_rel = 0.0;
  for (int64_t i = tid; i < B * 3; i += nth) {
    const float d = color[i] - gt[i];
    s_rgb += l1 ? (double)fabsf(d) : (double)(d * d);
  }
  if (use_bce) {
    for (int64_t i = tid; i < B; i += nth) {
      const float p = fminf(fmaxf(wsum[i], 1e-3f), 1.0f - 1e-3f), m = mask[i];
      // F.binary_cross_entropy clamps the logs at -100; unreachable after the clip, kept for fidelity
      s_bce -= (double)(m * fmaxf(logf(p), -100.0f) + (1.0f - m) * fmaxf(log1pf(-p), -100.0f));
    }
  }
  if (delta) {
    const int64_t per_ray = (int64_t)S * 3, n = B * per_ray;
    for (int64_t i = tid; i < n; i += nth) {
      const float m = mask_relight ? mask[i / per_ray] : 1.0f;
      s_rel += (double)(delta[i] * m);
    }
  }
  s_rgb = block_sum(s_rgb, sm);
  s_bce = block_sum(s_bce, sm);
  s_rel = block_sum(s_rel, sm);
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 3 + 0] = s_rgb;
    partial[blockIdx.x * 3 + 1] = s_bce;
    partial[blockIdx.x * 3 + 2] = s_rel;
  }
}

struct LossLambdas {
  float fine, eikonal, mask, relight;
};

__global__ void __launch_bounds__(LOSS_THREADS) loss_finish_kernel(const float* __restrict__ color, const float* __restrict__ gt,
                                                                     const float* __restrict__ wsum, const float* __restrict__ mask,
                                                                     const float* __restrict__ delta, const float* __restrict__ eik,
                                                                     int64_t B, int32_t S, int32_t l1, int32_t use_bce,
                                                                     int32_t mask_relight, const __grid_constant__ LossLambdas lam,
                                                                     const double* __restrict__ partial, int32_t n_partial,
                                                                     float* __restrict__ terms, float* __restrict__ g_color,
                                                                     float* __restrict__ g_wsum, float* __restrict__ g_delta) {
  double s_rgb = 0.0, s_bce = 0.0, s_rel = 0.0;
  for (int i = 0; i < n_partial; ++i) {
    s_rgb += partial[i * 3 + 0];
    s_bce += partial[i * 3 + 1];
    s_rel += partial[i * 3 + 2];
  }
  const float rgb = (float)(s_rgb / (double)(B * 3));
  const float bce = use_bce ? (float)(s_bce / (double)B) : 0.0f;
  const double n_rel = (double)B * S * 3;
  const float rel_mean = delta ? (float)(s_rel / n_rel) : 0.0f;
  const float rel = rel_mean * rel_mean;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const float e = eik[0];
    float loss = lam.fine * rgb + lam.eikonal * e;   // NeuS_Trainer.py:133-139, same order of accumulation
    if (use_bce) loss += lam.mask * bce;
    if (delta) loss += lam.relight * rel;
    terms[0] = loss; terms[1] = rgb; terms[2] = e; terms[3] = bce; terms[4] = rel;
  }
  const int64_t tid = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x, nth = (int64_t)gridDim.x * LOSS_THREADS;
  const float c_rgb = lam.fine / (float)(B * 3);
  for (int64_t i = tid; i < B * 3; i += nth) {
    const float d = color[i] - gt[i];
    g_color[i] = l1 ? c_rgb * (float)((d > 0.0f) - (d < 0.0f)) : 2.0f * c_rgb * d;
  }
  if (g_wsum) {
    const float c_m = lam.mask / (float)B;
    for (int64_t i = tid; i < B; i += nth) {
      float g = 0.0f;
      if (use_bce) {
        const float w = wsum[i];
        if (w >= 1e-3f && w <= 1.0f - 1e-3f)   // clamp's backward passes the gradient on the closed interval
          g = c_m * (w - mask[i]) / fmaxf((1.0f - w) * w, 1e-12f);
      }
      g_wsum[i] = g;
    }
  }
  if (delta && g_delta) {
    const float c_r = lam.relight * 2.0f * rel_mean / (float)n_rel;
    const int64_t per_ray = (int64_t)S * 3, n = B * per_ray;
    for (int64_t i = tid; i < n; i += nth) g_delta[i] = mask_relight ? c_r * mask[i / per_ray] : c_r;
  }
}

}  // namespace cneus

extern "C" size_t cneus_loss_workspace_bytes(void) { return (size_t)cneus::LOSS_BLOCKS * 3 * sizeof(double) + 256; }

extern "C" int cneus_neus_loss(const float* color_fine, const float* rgb_gt, const float* weight_sum, const float* mask,
                               const float* gradient_error, const float* delta_relight, int64_t B, int32_t S, int32_t rgb_l1,
                               float lambda_fine, float lambda_eikonal, float lambda_mask, float lambda_relight,
                               int32_t mask_relight, float* terms, float* g_color_fine, float* g_weight_sum,
                               float* g_delta_relight, void* ws, size_t ws_bytes, void* stream) {
  CNEUS_NVTX_RANGE();
  using namespace cneus;
  if (!color_fine || !rgb_gt || !gradient_error || !terms || !g_color_fine || !ws || B <= 0) {
    set_error("neus_loss: bad argument");
    return CNEUS_EINVAL;
  }
  const int use_bce = lambda_mask != 0.0f;
  if ((use_bce || (delta_relight && mask_relight)) && !mask) { set_error("neus_loss: mask required"); return CNEUS_EINVAL; }
  if (use_bce && (!weight_sum || !g_weight_sum)) { set_error("neus_loss: weight_sum / g_weight_sum required"); return CNEUS_EINVAL; }
  if (delta_relight && (S <= 0 || !g_delta_relight)) { set_error("neus_loss: delta_relight needs S and g_delta_relight"); return CNEUS_EINVAL; }
  if (ws_bytes < cneus_loss_workspace_bytes()) { set_error("neus_loss: workspace too small"); return CNEUS_ENOSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ws) + 7) & ~(uintptr_t)7);
  const int64_t work = delta_relight ? B * S * 3 : B * 3;
  int blocks = (int)((work + LOSS_THREADS * 4 - 1) / (LOSS_THREADS * 4));
  if (blocks > LOSS_BLOCKS) blocks = LOSS_BLOCKS;
  if (blocks < 1) blocks = 1;
  const float* dl = lambda_relight != 0.0f ? delta_relight : nullptr;
  loss_partial_kernel<<<blocks, LOSS_THREADS, 0, st>>>(color_fine, rgb_gt, weight_sum, mask, dl, B, S, rgb_l1, use_bce, mask_relight,
                                                        partial);
  LossLambdas lam{lambda_fine, lambda_eikonal, lambda_mask, lambda_relight};
  loss_finish_kernel<<<blocks, LOSS_THREADS, 0, st>>>(color_fine, rgb_gt, weight_sum, mask, dl, gradient_error, B, S, rgb_l1, use_bce,
                                                       mask_relight, lam, partial, blocks, terms, g_color_fine, g_weight_sum,
                                                       g_delta_relight);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch(2);
  return CNEUS_OK;
}
