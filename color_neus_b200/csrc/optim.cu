// per-parameter gradient-norm clipping + Adam as two multi-tensor launches.
// The reference clips every parameter tensor on its own (clip_grad_norm_(p, max_norm, 2) in a Python loop,
// lib/utils/net_utils.py:174-184: ~4 small launches per tensor, 53 tensors) and then runs torch.optim.Adam
// (betas 0.9 / 0.99, eps 1e-8, net_utils.py:88).  Here: launch 1 = sums of squares of every gradient (fp64 partials,
// fixed-order reduction -> deterministic); launch 2 = clip coefficient min(1, max_norm / (norm + 1e-6)), optional
// write-back of the clipped gradient (clip_grad_norm_ scales .grad in place), Adam moment and parameter update with
// torch's operation order (lerp, mul + addcmul, sqrt / bias_correction2_sqrt + eps, addcdiv).
#include "common.cuh"

namespace cneus {

constexpr int OPT_MAX_TENSORS = 64;   // per launch (descriptor table travels as a kernel parameter)
constexpr int OPT_BPT = 8;            // blocks per tensor

struct OptTable {
  CneusAdamTensor t[OPT_MAX_TENSORS];
  int32_t n;
};

__global__ void opt_sumsq_kernel(const __grid_constant__ OptTable tab, double* __restrict__ partial) {
  __shared__ double sm[8];
  const CneusAdamTensor& T = tab.t[blockIdx.y];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T.n; i += (int64_t)gridDim.x * blockDim.x) {
    const double g = (double)T.grad[i];
    s += g * g;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
    partial[blockIdx.y * OPT_BPT + blockIdx.x] = t;
  }
}

struct AdamScalars {
  float max_norm;             // <= 0: no clipping
  float beta1, beta2, eps;
  float one_minus_beta1, one_minus_beta2;
  float neg_step_size;        // -(lr / (1 - beta1^step))
  float bias_correction2_sqrt;
  float weight_decay;
  int32_t write_grad;
};

__global__ void opt_clip_adam_kernel(const __grid_constant__ OptTable tab, const double* __restrict__ partial,
                                     const __grid_constant__ AdamScalars a, float* __restrict__ norms_out) {
  const CneusAdamTensor& T = tab.t[blockIdx.y];
  double ss = 0.0;
  for (int i = 0; i < OPT_BPT; ++i) ss += partial[blockIdx.y * OPT_BPT + i];
  const float norm = (float)sqrt(ss);
  float coef = 1.0f;
  if (a.max_norm > 0.0f) coef = fminf(a.max_norm / (norm + 1e-6f), 1.0f);
  if (norms_out && blockIdx.x == 0 && threadIdx.x == 0) norms_out[blockIdx.y] = norm;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T.n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = T.grad[i] * coef;
    if (a.write_grad && coef != 1.0f) T.grad[i] = g;
    float p = T.param[i];
    if (a.weight_decay != 0.0f) g = fmaf(p, a.weight_decay, g);
    float m = T.exp_avg[i], v = T.exp_avg_sq[i];
    m = fmaf(a.one_minus_beta1, g - m, m);                 // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(a.one_minus_beta2 * g, g, v * a.beta2);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float denom = sqrtf(v) / a.bias_correction2_sqrt + a.eps;
    p = fmaf(a.neg_step_size, m / denom, p);               // param.addcdiv_(exp_avg, denom, value = -step_size)
    T.exp_avg[i] = m; T.exp_avg_sq[i] = v; T.param[i] = p;
  }
}

}  // namespace cneus

extern "C" size_t cneus_clip_adam_workspace_bytes(int32_t n_tensors) {
  return (size_t)(n_tensors > 0 ? n_tensors : 0) * cneus::OPT_BPT * sizeof(double) + 256;
}

extern "C" int cneus_clip_adam_step(const CneusAdamTensor* tensors, int32_t n_tensors, float max_norm, float lr, float beta1, float beta2,
                                    float eps, float weight_decay, int64_t step, int32_t write_clipped_grad, float* norms_out, void* ws,
                                    size_t ws_bytes, void* stream) {
  CNEUS_NVTX_RANGE();
  using namespace cneus;
  if (n_tensors <= 0) return CNEUS_OK;
  if (!tensors || !ws || step < 1) { set_error("clip_adam_step: bad argument"); return CNEUS_EINVAL; }
  if (ws_bytes < cneus_clip_adam_workspace_bytes(n_tensors)) { set_error("clip_adam_step: workspace too small"); return CNEUS_ENOSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ws) + 7) & ~(uintptr_t)7);
  AdamScalars a;
  a.max_norm = max_norm; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.one_minus_beta1 = 1.0f - beta1; a.one_minus_beta2 = 1.0f - beta2;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  a.neg_step_size = (float)(-((double)lr / bc1));
  a.bias_correction2_sqrt = (float)sqrt(bc2);
  a.weight_decay = weight_decay; a.write_grad = write_clipped_grad;
  for (int32_t base = 0; base < n_tensors; base += OPT_MAX_TENSORS) {
    OptTable tab;
    tab.n = (n_tensors - base < OPT_MAX_TENSORS) ? n_tensors - base : OPT_MAX_TENSORS;
    for (int i = 0; i < tab.n; ++i) {
      tab.t[i] = tensors[base + i];
      if (!tab.t[i].param || !tab.t[i].grad || !tab.t[i].exp_avg || !tab.t[i].exp_avg_sq || tab.t[i].n < 0) {
        set_error("clip_adam_step: tensor %d has a null pointer", base + i);
        return CNEUS_EINVAL;
      }
    }
    opt_sumsq_kernel<<<dim3(OPT_BPT, tab.n), 256, 0, st>>>(tab, partial + (size_t)base * OPT_BPT);
    opt_clip_adam_kernel<<<dim3(OPT_BPT, tab.n), 256, 0, st>>>(tab, partial + (size_t)base * OPT_BPT, a,
                                                                 norms_out ? norms_out + base : nullptr);
    CNEUS_CUDA_CHECK(cudaGetLastError());
    count_launch(2);
  }
  return CNEUS_OK;
}
