// NeuS alpha-from-SDF + compositing + Eikonal partial sums: one warp per ray.
// Restates NeuS.render_core (NeuS.py:233-279), Color_NeuS.render_core (Color_NeuS.py:66-123) and the derived
// outputs of NeuS.forward (NeuS.py:382-399).
#include "common.cuh"

namespace cneus {

constexpr int CW = 8;        // warps per CTA
constexpr int CMAXS = 512;

__global__ void __launch_bounds__(CW * 32) composite_kernel(const float* __restrict__ variance,
                                                            const float* __restrict__ ro, const float* __restrict__ rd,
                                                            const float* __restrict__ z, int64_t B, int S,
                                                            float cos_anneal, const __grid_constant__ CneusRenderOut o,
                                                            float* __restrict__ partials) {
  __shared__ float sm[CW][2][CMAXS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* al_s = sm[warp][0];
  float* w_s = sm[warp][1];
  // SingleVarianceNetwork.forward (fields.py:284-286) + clip (NeuS.py:233)
  const float inv_s = fminf(fmaxf(expf(__fmul_rn(variance[0], 10.0f)), 1e-6f), 1e6f);
  for (int64_t r = (int64_t)blockIdx.x * CW + warp; r < B; r += (int64_t)gridDim.x * CW) {
    const float ox = ro[r * 3], oy = ro[r * 3 + 1], oz = ro[r * 3 + 2];
    const float dx = rd[r * 3], dy = rd[r * 3 + 1], dz = rd[r * 3 + 2];
    const int64_t base = r * S;
    double eik_num = 0.0, eik_den = 0.0;
    __syncwarp();
    for (int i = lane; i < S; i += 32) {
      const float mid = o.mid_z[base + i], dist = o.dists[base + i], sdf = o.sdf[base + i];
      const float gx = o.gradients[(base + i) * 3], gy = o.gradients[(base + i) * 3 + 1], gz = o.gradients[(base + i) * 3 + 2];
      const float px = ray_point(ox, dx, mid), py = ray_point(oy, dy, mid), pz = ray_point(oz, dz, mid);
      const float pn = norm3(px, py, pz);
      const float tc = __fadd_rn(__fadd_rn(__fmul_rn(dx, gx), __fmul_rn(dy, gy)), __fmul_rn(dz, gz));
      const float ic = -__fadd_rn(__fmul_rn(fmaxf(__fadd_rn(__fmul_rn(-tc, 0.5f), 0.5f), 0.0f), __fsub_rn(1.0f, cos_anneal)),
                                  __fmul_rn(fmaxf(-tc, 0.0f), cos_anneal));
      const float h = __fmul_rn(__fmul_rn(ic, dist), 0.5f);
      const float pc = sigmoidf_(__fmul_rn(__fsub_rn(sdf, h), inv_s));
      const float nc = sigmoidf_(__fmul_rn(__fadd_rn(sdf, h), inv_s));
      float al = __fdiv_rn(__fadd_rn(__fsub_rn(pc, nc), 1e-5f), __fadd_rn(pc, 1e-5f));
      al = fminf(fmaxf(al, 0.0f), 1.0f);
      al_s[i] = al;
      if (o.cdf) o.cdf[base + i] = pc;
      if (o.alpha) o.alpha[base + i] = al;
      if (o.inside_sphere) o.inside_sphere[base + i] = pn < 1.0f ? 1.0f : 0.0f;
      const float relax = pn < 1.2f ? 1.0f : 0.0f;
      const float gn = __fsub_rn(norm3(gx, gy, gz), 1.0f);
      eik_num += (double)__fmul_rn(relax, __fmul_rn(gn, gn));
      eik_den += (double)relax;
    }
    __syncwarp();
    {  // exclusive cumprod of (1 - alpha + 1e-7) (NeuS.py:269-270): warp-level shuffle scan per 32-sample chunk with a carried
       // prefix, in double like torch's CPU accumulator
      double carry = 1.0;
      for (int i0 = 0; i0 < S; i0 += 32) {
        const int i = i0 + lane;
        const float al = i < S ? al_s[i] : 0.0f;
        const double f = i < S ? (double)__fadd_rn(__fsub_rn(1.0f, al), 1e-7f) : 1.0;
        const double incl = warp_scan_mul(f, lane) * carry;
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = carry;
        if (i < S) w_s[i] = __fmul_rn(al, (float)excl);
        carry = __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncwarp();
    double cr = 0, cg = 0, cb = 0, gr = 0, gg = 0, gb = 0, ws = 0, dep = 0;
    float wmax = 0.0f;
    for (int i = lane; i < S; i += 32) {
      const float w = w_s[i];
      if (o.weights) o.weights[base + i] = w;
      const float* c = o.sampled_color + (base + i) * 3;
      cr += (double)__fmul_rn(c[0], w); cg += (double)__fmul_rn(c[1], w); cb += (double)__fmul_rn(c[2], w);
      if (o.global_color) {
        const float* g = o.global_sampled + (base + i) * 3;
        gr += (double)__fmul_rn(g[0], w); gg += (double)__fmul_rn(g[1], w); gb += (double)__fmul_rn(g[2], w);
      }
      ws += (double)w;
      dep += (double)__fmul_rn(w, z[base + i]);
      wmax = fmaxf(wmax, w);
    }
    for (int off = 16; off > 0; off >>= 1) {
      cr += __shfl_xor_sync(0xffffffffu, cr, off); cg += __shfl_xor_sync(0xffffffffu, cg, off); cb += __shfl_xor_sync(0xffffffffu, cb, off);
      gr += __shfl_xor_sync(0xffffffffu, gr, off); gg += __shfl_xor_sync(0xffffffffu, gg, off); gb += __shfl_xor_sync(0xffffffffu, gb, off);
      ws += __shfl_xor_sync(0xffffffffu, ws, off); dep += __shfl_xor_sync(0xffffffffu, dep, off);
      wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
      eik_num += __shfl_xor_sync(0xffffffffu, eik_num, off); eik_den += __shfl_xor_sync(0xffffffffu, eik_den, off);
    }
    if (lane == 0) {
      if (o.color_fine) { o.color_fine[r * 3] = (float)cr; o.color_fine[r * 3 + 1] = (float)cg; o.color_fine[r * 3 + 2] = (float)cb; }
      if (o.global_color) { o.global_color[r * 3] = (float)gr; o.global_color[r * 3 + 1] = (float)gg; o.global_color[r * 3 + 2] = (float)gb; }
      if (o.weight_sum) o.weight_sum[r] = (float)ws;
      if (o.weight_max) o.weight_max[r] = wmax;
      if (o.depth) o.depth[r] = (float)dep;
      partials[r * 2] = (float)eik_num;
      partials[r * 2 + 1] = (float)eik_den;
    }
  }
}

// gradient_error = sum(relax * (|n|-1)^2) / (sum(relax) + 1e-5) over the whole batch (NeuS.py:275-277); fixed order.
__global__ void eikonal_reduce_kernel(const float* __restrict__ variance, const float* __restrict__ partials, int64_t B,
                                      float* __restrict__ scalars) {
  __shared__ double sn[256], sd[256];
  double n = 0.0, d = 0.0;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) { n += (double)partials[i * 2]; d += (double)partials[i * 2 + 1]; }
  sn[threadIdx.x] = n; sd[threadIdx.x] = d;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { sn[threadIdx.x] += sn[threadIdx.x + s]; sd[threadIdx.x] += sd[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float num = (float)sn[0], den = (float)sd[0];
    scalars[0] = __fdiv_rn(num, __fadd_rn(den, 1e-5f));
    scalars[1] = num;
    scalars[2] = den;
    const float inv_s = fminf(fmaxf(expf(__fmul_rn(variance[0], 10.0f)), 1e-6f), 1e6f);
    scalars[3] = __fdiv_rn(1.0f, inv_s);
  }
}

int launch_composite(const float* variance, const float* ro, const float* rd, const float* z, int64_t B, int S,
                     float cos_anneal, const CneusRenderOut& o, float* partials, cudaStream_t st) {
  if (B <= 0) return CNEUS_OK;
  if (S > CMAXS) { set_error("composite: S=%d exceeds %d", S, CMAXS); return CNEUS_EUNSUPPORTED; }
  if (!o.gradients || !o.sdf || !o.sampled_color || !o.mid_z || !o.dists || !o.scalars || (o.global_color && !o.global_sampled)) {
    set_error("composite: a required CneusRenderOut pointer is null");
    return CNEUS_EINVAL;
  }
  int64_t g = (B + CW - 1) / CW;
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap <= 0) cap = 148 * 8;
  composite_kernel<<<(int)(g > cap ? cap : g), CW * 32, 0, st>>>(variance, ro, rd, z, B, S, cos_anneal, o, partials);
  count_launch();
  CNEUS_CUDA_CHECK(cudaGetLastError());
  eikonal_reduce_kernel<<<1, 256, 0, st>>>(variance, partials, B, o.scalars);
  count_launch();
  CNEUS_CUDA_CHECK(cudaGetLastError());
  return CNEUS_OK;
}

}  // namespace cneus
