// Per-ray hierarchical sampling kernels: one warp owns one ray; all per-ray state lives in shared memory.
// Restates NeuS.forward's coarse depths (NeuS.py:311-326), NeuS.up_sample (NeuS.py:136-181),
// sample_pdf(det=True) (ray_utils.py:123-154) and the sort/gather of NeuS.cat_z_vals (NeuS.py:183-197).
#include "common.cuh"

namespace cneus {

constexpr int WARPS_PER_CTA = 8;
constexpr int MAXS = 512;  // max samples per ray handled by the per-ray kernels

// z[r][k] = near + (far - near) * lin[k] + (t_rand[r] - 0.5) * 2 / n_s      (NeuS.py:311-326)
__global__ void coarse_z_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                const float* __restrict__ t_rand, const float* __restrict__ lin, int64_t B, int n_s,
                                float* __restrict__ z) {
  int64_t total = B * n_s;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / n_s;
    int k = (int)(i % n_s);
    float nr = near[r], fr = far[r];
    float v = __fadd_rn(nr, __fmul_rn(__fsub_rn(fr, nr), lin[k]));
    if (t_rand != nullptr) {
      float t = __fsub_rn(t_rand[r], 0.5f);
      v = __fadd_rn(v, __fdiv_rn(__fmul_rn(t, 2.0f), (float)n_s));
    }
    z[i] = v;
  }
}

// One warp per ray.  The scans (exclusive cumprod of the transmittance, cumsum of the pdf) are warp-level shuffle scans over
// 32-sample chunks with a carried prefix, in double like torch's CPU cumprod / cumsum (acc_type<float> = double, stored back
// as float).
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) up_sample_kernel(const float* __restrict__ ro,
                                                                       const float* __restrict__ rd,
                                                                       const float* __restrict__ z_g,
                                                                       const float* __restrict__ sdf_g, int64_t B, int n,
                                                                       int m, float inv_s, const float* __restrict__ u_g,
                                                                       float* __restrict__ new_z) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* zs = sm + (size_t)warp * (4 * MAXS);
  float* ss = zs + MAXS;    // sdf, later weights
  float* cs = ss + MAXS;    // cos, later cdf
  float* as = cs + MAXS;    // inside flag (per sample), later alpha
  for (int64_t r = (int64_t)blockIdx.x * WARPS_PER_CTA + warp; r < B; r += (int64_t)gridDim.x * WARPS_PER_CTA) {
    const float ox = ro[r * 3], oy = ro[r * 3 + 1], oz = ro[r * 3 + 2];
    const float dx = rd[r * 3], dy = rd[r * 3 + 1], dz = rd[r * 3 + 2];
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      float zv = z_g[r * n + i];
      zs[i] = zv;
      ss[i] = sdf_g[r * n + i];
      float rad = norm3(ray_point(ox, dx, zv), ray_point(oy, dy, zv), ray_point(oz, dz, zv));
      as[i] = rad < 1.0f ? 1.0f : 0.0f;
    }
    __syncwarp();
    for (int i = lane; i < n - 1; i += 32)
      cs[i] = __fdiv_rn(__fsub_rn(ss[i + 1], ss[i]), __fadd_rn(__fsub_rn(zs[i + 1], zs[i]), 1e-5f));
    __syncwarp();
    float alpha_reg[MAXS / 32];
#pragma unroll
    for (int j = 0; j < MAXS / 32; ++j) {
      int i = lane + 32 * j;
      float al = 0.0f;
      if (i < n - 1) {
        float prev = i > 0 ? cs[i - 1] : 0.0f;
        float c = fminf(prev, cs[i]);
        c = fminf(fmaxf(c, -1e3f), 0.0f);
        float inside = (as[i] > 0.5f || as[i + 1] > 0.5f) ? 1.0f : 0.0f;
        c = __fmul_rn(c, inside);
        float dist = __fsub_rn(zs[i + 1], zs[i]);
        float mid = __fmul_rn(__fadd_rn(ss[i], ss[i + 1]), 0.5f);
        float h = __fmul_rn(__fmul_rn(c, dist), 0.5f);
        float pc = sigmoidf_(__fmul_rn(__fsub_rn(mid, h), inv_s));
        float nc = sigmoidf_(__fmul_rn(__fadd_rn(mid, h), inv_s));
        al = __fdiv_rn(__fadd_rn(__fsub_rn(pc, nc), 1e-5f), __fadd_rn(pc, 1e-5f));
      }
      alpha_reg[j] = al;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < MAXS / 32; ++j) {
      int i = lane + 32 * j;
      if (i < n - 1) as[i] = alpha_reg[j];
    }
    __syncwarp();
    // weights = alpha * exclusive_cumprod(1 - alpha + 1e-7), then w += 1e-5 (sample_pdf) and its sum
    {
      double carry = 1.0;   // product of all factors before this chunk
      for (int i0 = 0; i0 < n - 1; i0 += 32) {
        const int i = i0 + lane;
        const float al = i < n - 1 ? as[i] : 0.0f;
        const double f = i < n - 1 ? (double)__fadd_rn(__fsub_rn(1.0f, al), 1e-7f) : 1.0;
        const double incl = warp_scan_mul(f, lane) * carry;          // prod_{j <= i}
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);           // prod_{j < i}
        if (lane == 0) excl = carry;
        if (i < n - 1) ss[i] = __fadd_rn(__fmul_rn(al, (float)excl), 1e-5f);
        carry = __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncwarp();
    // torch.sum(weights, -1): pairwise-ish float reduction on CPU; a double sum rounded once is within its error
    double part = 0.0;
    for (int i = lane; i < n - 1; i += 32) part += (double)ss[i];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const float wsum = (float)part;
    for (int i = lane; i < n - 1; i += 32) ss[i] = __fdiv_rn(ss[i], wsum);  // pdf
    __syncwarp();
    {  // cdf = [0, cumsum(pdf)]  (n entries)
      double carry = 0.0;
      if (lane == 0) cs[0] = 0.0f;
      for (int i0 = 0; i0 < n - 1; i0 += 32) {
        const int i = i0 + lane;
        const double incl = warp_scan_add(i < n - 1 ? (double)ss[i] : 0.0, lane) + carry;
        if (i < n - 1) cs[i + 1] = (float)incl;
        carry = __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncwarp();
    for (int j = lane; j < m; j += 32) {
      const float u = u_g[j];
      // searchsorted(cdf, u, right=True): first index with cdf[idx] > u
      int lo = 0, hi = n;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cs[mid] > u) hi = mid; else lo = mid + 1;
      }
      int below = lo - 1 < 0 ? 0 : lo - 1;
      int above = lo > n - 1 ? n - 1 : lo;
      float c0 = cs[below], c1 = cs[above], b0 = zs[below], b1 = zs[above];
      float denom = __fsub_rn(c1, c0);
      if (denom < 1e-5f) denom = 1.0f;
      float t = __fdiv_rn(__fsub_rn(u, c0), denom);
      new_z[r * m + j] = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
    }
  }
}

// Merge two ascending lists per ray (old depths first on ties, like a stable sort of cat([z, new_z])) and carry the
// SDF values along.  One warp per ray; positions by binary search.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) merge_kernel(const float* __restrict__ z,
                                                                   const float* __restrict__ nz,
                                                                   const float* __restrict__ sdf,
                                                                   const float* __restrict__ nsdf, int64_t B, int n,
                                                                   int m, float* __restrict__ z_out,
                                                                   float* __restrict__ sdf_out) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* zs = sm + (size_t)warp * (2 * MAXS);
  float* ns = zs + MAXS;
  for (int64_t r = (int64_t)blockIdx.x * WARPS_PER_CTA + warp; r < B; r += (int64_t)gridDim.x * WARPS_PER_CTA) {
    __syncwarp();
    for (int i = lane; i < n; i += 32) zs[i] = z[r * n + i];
    for (int j = lane; j < m; j += 32) ns[j] = nz[r * m + j];
    __syncwarp();
    const int64_t ob = r * (n + m);
    for (int i = lane; i < n; i += 32) {  // # new strictly below z[i]
      float v = zs[i];
      int lo = 0, hi = m;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (ns[mid] < v) lo = mid + 1; else hi = mid; }
      z_out[ob + i + lo] = v;
      if (sdf_out != nullptr) sdf_out[ob + i + lo] = sdf[r * n + i];
    }
    for (int j = lane; j < m; j += 32) {  // # old <= new_z[j]
      float v = ns[j];
      int lo = 0, hi = n;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (zs[mid] <= v) lo = mid + 1; else hi = mid; }
      z_out[ob + j + lo] = v;
      if (sdf_out != nullptr) sdf_out[ob + j + lo] = nsdf[r * m + j];
    }
  }
}

// dists / mid-points of the sections (NeuS.py:215-218)
__global__ void sections_kernel(const float* __restrict__ z, int64_t B, int S, float sample_dist, float* __restrict__ mid,
                                float* __restrict__ dists) {
  int64_t total = B * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(i % S);
    float zi = z[i];
    float dd = (k + 1 < S) ? __fsub_rn(z[i + 1], zi) : sample_dist;
    dists[i] = dd;
    mid[i] = __fadd_rn(zi, __fmul_rn(dd, 0.5f));
  }
}

static int grid_1d(int64_t total, int block) {
  int64_t g = (total + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 16;
  if (cap <= 0) cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int launch_coarse_z(const float* near, const float* far, const float* t_rand, const float* lin, int64_t B, int n_s,
                    float* z, cudaStream_t st) {
  if (B <= 0) return CNEUS_OK;
  coarse_z_kernel<<<grid_1d(B * n_s, 256), 256, 0, st>>>(near, far, t_rand, lin, B, n_s, z);
  count_launch();
  CNEUS_CUDA_CHECK(cudaGetLastError());
  return CNEUS_OK;
}

int launch_up_sample(const float* ro, const float* rd, const float* z, const float* sdf, int64_t B, int n, int m,
                     float inv_s, const float* u, float* new_z, cudaStream_t st) {
  if (B <= 0) return CNEUS_OK;
  if (n < 2 || n > MAXS || m < 1 || m > MAXS) { set_error("up_sample: n=%d m=%d out of range (max %d)", n, m, MAXS); return CNEUS_EUNSUPPORTED; }
  static bool attr[CNEUS_MAX_DEVICES] = {false};
  const size_t smem = (size_t)WARPS_PER_CTA * 4 * MAXS * sizeof(float);
  if (first_use_on_device(attr)) CNEUS_CUDA_CHECK(cudaFuncSetAttribute(up_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  up_sample_kernel<<<grid_1d(B, WARPS_PER_CTA), WARPS_PER_CTA * 32, smem, st>>>(ro, rd, z, sdf, B, n, m, inv_s, u, new_z);
  count_launch();
  CNEUS_CUDA_CHECK(cudaGetLastError());
  return CNEUS_OK;
}

int launch_merge(const float* z, const float* new_z, const float* sdf, const float* new_sdf, int64_t B, int n, int m,
                 float* z_out, float* sdf_out, cudaStream_t st) {
  if (B <= 0) return CNEUS_OK;
  if (n > MAXS || m > MAXS) { set_error("merge: n=%d m=%d out of range", n, m); return CNEUS_EUNSUPPORTED; }
  const size_t smem = (size_t)WARPS_PER_CTA * 2 * MAXS * sizeof(float);
  merge_kernel<<<grid_1d(B, WARPS_PER_CTA), WARPS_PER_CTA * 32, smem, st>>>(z, new_z, sdf, new_sdf, B, n, m, z_out, sdf_out);
  count_launch();
  CNEUS_CUDA_CHECK(cudaGetLastError());
  return CNEUS_OK;
}

int launch_sections(const float* z, int64_t B, int S, float sample_dist, float* mid, float* dists, cudaStream_t st) {
  if (B <= 0) return CNEUS_OK;
  sections_kernel<<<grid_1d(B * S, 256), 256, 0, st>>>(z, B, S, sample_dist, mid, dists);
  count_launch();
  CNEUS_CUDA_CHECK(cudaGetLastError());
  return CNEUS_OK;
}

}  // namespace cneus
