// Internal definitions shared by the libcneus.so translation units (not part of the C ABI).
#pragma once
#include <nvtx3/nvToolsExt.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "cneus.h"

namespace cneus {

constexpr int TM = 64;        // points per SIMT tile
constexpr int NT = 256;       // threads per CTA of the point-shading kernel
constexpr int KC = 16;        // K rows per streamed weight chunk
constexpr int MAXH = 256;     // widest hidden layer supported
constexpr int SMALLK = 48;    // rows of the "small input" segment (pts | PE(dir) | normal)

// A GEMM-shaped layer in the packed buffer.  Forward operand Wt is [K0+K1][Np] (row k = input feature,
// contiguous over outputs); the input features come from two shared-memory segments: K0 rows of the small
// segment then K1 rows of the main activation buffer.  Backward operand Wb (SDF layers only) is the plain
// row-major [Nb][Kb] matrix (row = output feature) used for the input-gradient chain.
struct PLayer {
  int64_t wt_off;    // float offset of Wt, -1 if absent
  int64_t wb_off;    // float offset of Wb, -1 if absent
  int64_t bias_off;  // float offset of bias [Np]
  int32_t K0, K1;    // padded row counts of the two input segments (multiples of KC; K0 may be 0)
  int32_t k0v, k1v;  // valid rows in each segment
  int32_t N, Np;     // valid / padded (multiple of 64) outputs
  int32_t Nb, Kb;    // backward operand: rows (multiple of 16), cols (multiple of 64)
};

// A narrow layer (N <= 4) evaluated as dot products; W is row-major [N][K0+K1] with zero padding.
struct RowLayer {
  int64_t w_off;
  int64_t bias_off;
  int32_t K0, K1, N, pad_;
};

struct NetPack {
  CneusNetDesc d;
  PLayer sdf[CNEUS_MAX_SDF_LIN];      // [0 .. n_lin-2] hidden layers; [n_lin-1] = feature rows (1..d_out-1) of the last layer
  RowLayer sdf_row;                   // row 0 of the last layer (the sdf column)
  PLayer color[CNEUS_MAX_COLOR_LIN];  // [0 .. n_lin-2]
  RowLayer color_row;                 // last colour layer (d_out = 3)
  PLayer rl_in;
  PLayer rl[CNEUS_MAX_RELIGHT_LIN];   // [0 .. n_layers-2]
  RowLayer rl_row;                    // last relight layer
  int32_t pe_dim;                     // 3*(1+2*multires) of the SDF input
  int32_t color_k0v;                  // valid rows of the colour small segment
  int32_t relight_k0v;
  int64_t total_floats;
  // tensor-core weight images (byte offsets from the start of the packed buffer; see mlp_tc.cu)
  int32_t tc_eligible;
  int64_t tc_sdf_fwd[CNEUS_MAX_SDF_LIN];  // [0 .. n_lin-2] hidden layers, [n_lin-1] feature block of the last layer
  int64_t tc_sdf_bwd[CNEUS_MAX_SDF_LIN];  // transposed images of the hidden layers (gradient chain)
  int64_t tc_color[CNEUS_MAX_COLOR_LIN];
  int64_t tc_rl_in;
  int64_t tc_rl[CNEUS_MAX_RELIGHT_LIN];
};

int build_netpack(const CneusNetDesc* d, NetPack* np);  // host; returns CNEUS_* status

void set_error(const char* fmt, ...);

#define CNEUS_CUDA_CHECK(expr)                                                        \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      cneus::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return CNEUS_ECUDA;                                                             \
    }                                                                                 \
  } while (0)

__host__ __device__ inline int pad_to(int x, int m) { return (x + m - 1) / m * m; }

// ---- device math shared by every kernel (fp32, IEEE rounding, no fast-math) ---------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// nn.Softplus(beta=100, threshold=20)  (fields.py:77)
__device__ __forceinline__ float softplus100(float a) {
  float z = 100.0f * a;
  return z > 20.0f ? a : log1pf(expf(z)) / 100.0f;
}
__device__ __forceinline__ float softplus100_grad(float a) {
  float z = 100.0f * a;
  return z > 20.0f ? 1.0f : sigmoidf_(z);
}
// o + d * t with torch's two roundings (no FMA contraction) -- NeuS.py:220
// Warp-level inclusive scans of one double per lane (Kogge-Stone over shuffles).  The scans of the per-ray compositing and of
// the inverse-CDF sampler keep torch's CPU accumulator type (acc_type<float> = double): the association order differs from
// torch's sequential loop, which moves the double result by ~1e-16 relative -- invisible after the rounding to float.
__device__ __forceinline__ double warp_scan_mul(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ double warp_scan_add(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__device__ __forceinline__ float ray_point(float o, float d, float t) { return __fadd_rn(o, __fmul_rn(d, t)); }
__device__ __forceinline__ float norm3(float x, float y, float z) {
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// Optional per-point dumps of the SDF forward pass and of the reverse chain, written by the tensor-core kernel for the
// training backward (backward.cu), all fp32 row-major with 256 floats per point (gx0: ld_gx0):
//   in[l] (l >= 1)  input of SDF layer l: softplus(a_{l-1}), / sqrt(2) with the encoding appended when l is the skip layer
//   d[l]            softplus'(a_l)
//   gh[l], ga[l]    d sdf / d h_{l+1} (nullable: not needed by the fused backward) and gh[l] (.) softplus'(a_l),
//                   l < n_hidden - 1 (the last pair is a constant row)
//   gx0             d sdf / d (encoded input), [P, ld_gx0]
struct TrainDump {
  float* in[CNEUS_MAX_SDF_LIN];
  float* d[CNEUS_MAX_SDF_LIN];
  float* gh[CNEUS_MAX_SDF_LIN];
  float* ga[CNEUS_MAX_SDF_LIN];
  float* gx0;
  int32_t ld_gx0;
  int32_t on;
};

// ---- kernels' host launchers (defined in the .cu files) -------------------------------------------------------
struct ShadeArgs {
  // point source: 0 explicit pts[P,3]; 1 rays: p -> (r = p / n_per_ray), x = o[r] + d[r] * t[p]; 2 grid axes
  int32_t src_mode;
  int32_t n_per_ray;
  const float* pts;
  const float* rays_o;
  const float* rays_d;
  const float* t;
  const float* gx;
  const float* gy;
  const float* gz;
  int32_t res;
  int32_t viewdir_mode;  // 0: in_viewdirs if given else ray direction; 1: minus the SDF gradient (extract_color)
  int64_t lin_begin;
  int64_t P;
  // optional explicit inputs (stand-alone sub-module calls)
  const float* in_normals;
  const float* in_viewdirs;
  const float* in_feats;
  const float* in_rgb;
  // stages
  int32_t run_sdf;  // 0 none, 1 sdf column only, 2 sdf + feature
  int32_t run_grad;
  int32_t run_color;
  int32_t run_relight;
  // outputs (nullable)
  float* out_sdf;    // [P]
  float* out_full;   // [P, d_out]
  float* out_grad;   // [P,3]
  float* out_color;  // [P,3]  colour-network output
  float* out_relit;  // [P,3]
  float* out_drgb;   // [P,3]
  float out_sdf_sign;
  float* dscratch;  // grid * (n_hidden * MAXH * TM) floats
  TrainDump dump;   // tensor-core path only
  // Tangent pass of the SDF double backward (training, tensor-core path, dump.on required): t_{l+1} = softplus'(a_l) (.)
  // (W_l t_l), seeded by tan_t0 [P, pe_dim]; reads dump.d[l] (softplus') and dump.gh[l] (here: the adjoint AFTER the multiplication by softplus', ga_l); writes dump.in[l + 1] = t_{l+1}
  // ([P, 256], the skip layer's input with t_0 / sqrt(2) appended) and dump.ga[l] = softplus''(a_l) (.) (W_l t_l) (.) gh_l.
  // tan_amax: device float, max |tan_t0| (the A operand is scaled by a power of two taken from it).
  int32_t run_tangent;
  const float* tan_t0;
  const float* tan_amax;
};
int launch_shade(const NetPack& np, const float* packed, const ShadeArgs& a, int grid, cudaStream_t st);
size_t shade_scratch_floats_per_cta(const NetPack& np);
int shade_grid_for(int64_t P);

int launch_coarse_z(const float* near, const float* far, const float* t_rand, const float* lin, int64_t B, int n_s,
                    float* z, cudaStream_t st);
int launch_up_sample(const float* ro, const float* rd, const float* z, const float* sdf, int64_t B, int n, int m,
                     float inv_s, const float* u, float* new_z, cudaStream_t st);
int launch_merge(const float* z, const float* new_z, const float* sdf, const float* new_sdf, int64_t B, int n, int m,
                 float* z_out, float* sdf_out, cudaStream_t st);
int launch_sections(const float* z, int64_t B, int S, float sample_dist, float* mid, float* dists, cudaStream_t st);
int launch_composite(const float* variance, const float* ro, const float* rd, const float* z, int64_t B, int S,
                     float cos_anneal, const CneusRenderOut& o, float* partials, cudaStream_t st);

// NVTX range around a C-ABI entry point (tracing): shows up as a named host-side range in nsys / ncu
// timelines; a few ns when no tool is attached (header-only NVTX3, no library to link).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define CNEUS_NVTX_RANGE() ::cneus::NvtxRange cneus_nvtx_range_(__func__)

int sm_count();  // of the current device (cached per device)
void count_launch(int n = 1);
// Per-device one-time setup: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the device that is current when
// it is called, so a process that renders on several GPUs has to opt in on each of them.  `flags` is one array per call
// site; true the first time the site sees the current device.
constexpr int CNEUS_MAX_DEVICES = 64;
inline bool first_use_on_device(bool (&flags)[CNEUS_MAX_DEVICES]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= CNEUS_MAX_DEVICES) return true;  // unknown: just set it again
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

// tensor-core path (mlp_tc.cu)
constexpr int64_t TC_STAGE_BYTES_HOST = 32768;  // one (K-block, N-half) stage image: hi + lo slab
bool tc_supports(const NetPack& np, const ShadeArgs& a);
int launch_shade_tc(const NetPack& np, const float* packed, const ShadeArgs& a, float* gxscratch, cudaStream_t st);
int pack_tc_weights(const NetPack& np, const CneusParams* P, const int64_t* scale_off, float* packed, cudaStream_t st);
size_t tc_scratch_floats_per_cta(const NetPack& np);  // softplus' slots + encoding-adjoint scratch
extern int g_force_simt;

}  // namespace cneus
