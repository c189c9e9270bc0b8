// On-device ray generation for selected pixels (SURVEY.md section 8f "next" #1): the reference builds ALL N*H*W rays of
// every camera each training step and then gathers n_rays of them (lib/models/tools/ray_utils.py:16-87); this kernel
// generates only the selected rays, in the caller's index order (flat index = (cam * H + y) * W + x, the reference's
// reshape(-1, 3) order), with the origin/radius normalisation of NeuS_Trainer.render (NeuS_Trainer.py:121-122) and
// near_far_from_sphere (ray_utils.py:7-13) fused in.  Same fp32 operation order as the torch expressions
// (IEEE mul/add/div/sqrt, no FMA contraction).
#include "common.cuh"

namespace cneus {

struct RayGenArgs {
  const float* c2w;      // [n_cam, 4, 4] row-major (rows 0..2 used)
  const float* focal;    // [2]
  const int64_t* index;  // [n] flat pixel indices, or nullptr: index = first + t
  const float* origin;   // [3] nullable
  const float* radius;   // [1] nullable
  const float* image;    // [n_cam*H*W, 3] nullable: rgb gather
  const float* mask;     // [n_cam*H*W] nullable: mask gather
  float* rays_o; float* rays_d; float* near; float* far; float* rgb; float* mask_out;
  int64_t n, first;
  int32_t H, W, n_cam, normalize, opengl;
};

__global__ void gen_rays_kernel(const __grid_constant__ RayGenArgs a) {
  const int64_t hw = (int64_t)a.H * a.W;
  const float fx = a.focal[0], fy = a.focal[1];
  const float sy = a.opengl ? -1.0f : 1.0f, sz = a.opengl ? -1.0f : 1.0f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = a.index ? a.index[t] : a.first + t;
    const int64_t cam = idx / hw, pix = idx - cam * hw;
    const int y = (int)(pix / a.W), x = (int)(pix - (int64_t)y * a.W);
    // dirs = [(i - W/2) / fx, ys (j - H/2) / fy, zs]   (ray_utils.py:47, :111)
    float dx = __fdiv_rn(__fsub_rn((float)x, __fmul_rn((float)a.W, 0.5f)), fx);
    float dy = __fdiv_rn(__fmul_rn(sy, __fsub_rn((float)y, __fmul_rn((float)a.H, 0.5f))), fy);
    float dz = sz;
    if (a.normalize) {
      const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      dx = __fdiv_rn(dx, nrm); dy = __fdiv_rn(dy, nrm); dz = __fdiv_rn(dz, nrm);
    }
    const float* R = a.c2w + cam * 16;
    float d[3], o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      d[c] = __fadd_rn(__fadd_rn(__fmul_rn(dx, R[c * 4]), __fmul_rn(dy, R[c * 4 + 1])), __fmul_rn(dz, R[c * 4 + 2]));
      o[c] = R[c * 4 + 3];
      if (a.origin) o[c] = __fsub_rn(o[c], a.origin[c]);
      if (a.radius) o[c] = __fdiv_rn(o[c], a.radius[0]);
      a.rays_o[t * 3 + c] = o[c];
      a.rays_d[t * 3 + c] = d[c];
    }
    if (a.near) {
      const float aa = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
      const float bb = __fmul_rn(2.0f, __fadd_rn(__fadd_rn(__fmul_rn(o[0], d[0]), __fmul_rn(o[1], d[1])), __fmul_rn(o[2], d[2])));
      const float mid = __fdiv_rn(__fmul_rn(0.5f, -bb), aa);
      a.near[t] = __fsub_rn(mid, 1.0f);
      a.far[t] = __fadd_rn(mid, 1.0f);
    }
    if (a.rgb && a.image) {
#pragma unroll
      for (int c = 0; c < 3; ++c) a.rgb[t * 3 + c] = a.image[idx * 3 + c];
    }
    if (a.mask_out && a.mask) a.mask_out[t] = a.mask[idx];
  }
}

// the training images stay resident on the device as the uint8 they are stored as; the pixels of the
// selected rays are converted on the fly with the float pipeline of the reference datasets' get_image (lib/datasets/dtu.py:
// 98-113, bmvs.py:99-114, omniobject3d.py:98-111), bit for bit:
//   x = float(u8) / 255                       tvF.to_tensor
//   x = (x - 0.5) / std                       tvF.normalize(image, [0.5]*3, [std]*3)
//   x = x * 0.5 + 0.5
//   m = float(mask_u8) / 255                  tvF.to_tensor(mask)
//   x = x * m                                 (dtu / bmvs only: premultiply)
struct PixGatherArgs {
  const uint8_t* images;   // [n_img, H, W, 3] RGB
  const uint8_t* masks;    // [n_img, H, W] nullable
  const int64_t* cam_map;  // [n_cam] batch position -> dataset image (nullable: identity)
  const int64_t* index;    // [n] flat index into the BATCH: (cam * H + y) * W + x
  float* rgb;              // [n, 3]
  float* mask_out;         // [n] nullable
  int64_t n, hw;
  float std;
  int32_t premultiply;
};

__global__ void gather_pixels_u8_kernel(const __grid_constant__ PixGatherArgs a) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = a.index[t];
    const int64_t cam = idx / a.hw, pix = idx - cam * a.hw;
    const int64_t src = (a.cam_map ? a.cam_map[cam] : cam) * a.hw + pix;
    float m = 1.0f;
    if (a.masks) {
      m = __fdiv_rn((float)a.masks[src], 255.0f);
      if (a.mask_out) a.mask_out[t] = m;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = __fdiv_rn((float)a.images[src * 3 + c], 255.0f);
      x = __fdiv_rn(__fsub_rn(x, 0.5f), a.std);
      x = __fadd_rn(__fmul_rn(x, 0.5f), 0.5f);
      if (a.premultiply && a.masks) x = __fmul_rn(x, m);
      a.rgb[t * 3 + c] = x;
    }
  }
}

}  // namespace cneus

extern "C" int cneus_gen_rays(const float* c2w, int32_t n_cam, const float* focal, int32_t H, int32_t W, const int64_t* index,
                              int64_t first, int64_t n, int32_t normalize, int32_t opengl, const float* origin, const float* radius,
                              const float* image, const float* mask, float* rays_o, float* rays_d, float* near, float* far,
                              float* rgb, float* mask_out, void* stream) {
  CNEUS_NVTX_RANGE();
  using namespace cneus;
  if (!c2w || !focal || !rays_o || !rays_d || n_cam <= 0 || H <= 0 || W <= 0) { set_error("gen_rays: bad argument"); return CNEUS_EINVAL; }
  if ((near == nullptr) != (far == nullptr)) { set_error("gen_rays: near and far go together"); return CNEUS_EINVAL; }
  if (n <= 0) return CNEUS_OK;
  RayGenArgs a;
  a.c2w = c2w; a.focal = focal; a.index = index; a.origin = origin; a.radius = radius; a.image = image; a.mask = mask;
  a.rays_o = rays_o; a.rays_d = rays_d; a.near = near; a.far = far; a.rgb = rgb; a.mask_out = mask_out;
  a.n = n; a.first = first; a.H = H; a.W = W; a.n_cam = n_cam; a.normalize = normalize; a.opengl = opengl;
  const int64_t blocks = (n + 255) / 256;
  gen_rays_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(a);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return CNEUS_OK;
}

extern "C" int cneus_gather_pixels_u8(const uint8_t* images, const uint8_t* masks, const int64_t* cam_map, const int64_t* index,
                                      int64_t n, int32_t H, int32_t W, float std, int32_t premultiply_mask, float* rgb,
                                      float* mask_out, void* stream) {
  CNEUS_NVTX_RANGE();
  using namespace cneus;
  if (!images || !index || !rgb || H <= 0 || W <= 0 || !(std > 0.0f)) { set_error("gather_pixels_u8: bad argument"); return CNEUS_EINVAL; }
  if (mask_out && !masks) { set_error("gather_pixels_u8: mask_out without masks"); return CNEUS_EINVAL; }
  if (n <= 0) return CNEUS_OK;
  PixGatherArgs a;
  a.images = images; a.masks = masks; a.cam_map = cam_map; a.index = index; a.rgb = rgb; a.mask_out = mask_out;
  a.n = n; a.hw = (int64_t)H * W; a.std = std; a.premultiply = premultiply_mask;
  const int64_t blocks = (n + 255) / 256;
  gather_pixels_u8_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(a);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return CNEUS_OK;
}
