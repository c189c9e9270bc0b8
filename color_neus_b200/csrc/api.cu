// C-ABI entry points (include/cneus.h): argument checking, workspace carving and kernel sequencing.
// No allocation, no host synchronisation: everything is enqueued on the caller's stream.
#include <string.h>

#include "common.cuh"

using namespace cneus;

namespace {

struct Workspace {
  float* dscratch;   // per-CTA softplus' storage of the point-shading kernel
  float* ray;        // per-ray scratch
  size_t ray_floats;
};

size_t dscratch_floats(const NetPack& np) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  size_t per = shade_scratch_floats_per_cta(np);
  if (np.tc_eligible && tc_scratch_floats_per_cta(np) > per) per = tc_scratch_floats_per_cta(np);
  return (size_t)sms * per;
}

int carve(const NetPack& np, void* ws, size_t ws_bytes, size_t ray_floats_needed, Workspace* w) {
  size_t need = (dscratch_floats(np) + ray_floats_needed) * sizeof(float);
  if (ws == nullptr || ws_bytes < need) {
    set_error("workspace too small: need %zu bytes, have %zu", need, ws_bytes);
    return CNEUS_ENOSPACE;
  }
  w->dscratch = (float*)ws;
  w->ray = w->dscratch + dscratch_floats(np);
  w->ray_floats = ray_floats_needed;
  return CNEUS_OK;
}

ShadeArgs blank_args() {
  ShadeArgs a;
  memset(&a, 0, sizeof(a));
  a.out_sdf_sign = 1.0f;
  return a;
}

#define CHECK_RC(expr) do { int rc__ = (expr); if (rc__ != CNEUS_OK) return rc__; } while (0)
#define REQUIRE(cond, msg) do { if (!(cond)) { set_error(msg); return CNEUS_EINVAL; } } while (0)

}  // namespace

extern "C" size_t cneus_workspace_bytes(const CneusNetDesc* desc, int64_t n_rays, int32_t n_total_samples, int64_t n_points) {
  NetPack np;
  if (desc == nullptr || build_netpack(desc, &np) != CNEUS_OK) return 0;
  (void)n_points;
  size_t ray = (size_t)(n_rays > 0 ? n_rays : 0) * (size_t)(6 * (n_total_samples > 0 ? n_total_samples : 0) + 2) + 256;
  return (dscratch_floats(np) + ray) * sizeof(float);
}

extern "C" int cneus_sdf_forward(const CneusNetDesc* desc, const void* packed, const float* pts, int64_t P, float* out,
                                 int32_t out_cols, void* ws, size_t ws_bytes, void* stream) {
  CNEUS_NVTX_RANGE();
  if (P <= 0) return CNEUS_OK;  // empty input: nothing to do (empty torch tensors carry null pointers)
  REQUIRE(desc && packed && pts && out, "sdf_forward: null argument");
  NetPack np;
  CHECK_RC(build_netpack(desc, &np));
  REQUIRE(out_cols == 1 || out_cols == desc->sdf_d_out, "sdf_forward: out_cols must be 1 or sdf_d_out");
  Workspace w;
  CHECK_RC(carve(np, ws, ws_bytes, 0, &w));
  ShadeArgs a = blank_args();
  a.src_mode = 0; a.pts = pts; a.P = P; a.dscratch = w.dscratch;
  if (out_cols == 1) { a.run_sdf = 1; a.out_sdf = out; } else { a.run_sdf = 2; a.out_full = out; }
  return launch_shade(np, (const float*)packed, a, shade_grid_for(P), (cudaStream_t)stream);
}

extern "C" int cneus_sdf_gradient(const CneusNetDesc* desc, const void* packed, const float* pts, int64_t P, float* grad,
                                  void* ws, size_t ws_bytes, void* stream) {
  CNEUS_NVTX_RANGE();
  if (P <= 0) return CNEUS_OK;  // empty input: nothing to do (empty torch tensors carry null pointers)
  REQUIRE(desc && packed && pts && grad, "sdf_gradient: null argument");
  NetPack np;
  CHECK_RC(build_netpack(desc, &np));
  Workspace w;
  CHECK_RC(carve(np, ws, ws_bytes, 0, &w));
  ShadeArgs a = blank_args();
  a.src_mode = 0; a.pts = pts; a.P = P; a.run_sdf = 1; a.run_grad = 1; a.out_grad = grad; a.dscratch = w.dscratch;
  return launch_shade(np, (const float*)packed, a, shade_grid_for(P), (cudaStream_t)stream);
}

extern "C" int cneus_color_forward(const CneusNetDesc* desc, const void* packed, const float* pts, const float* normals,
                                   const float* view_dirs, const float* feats, int64_t P, float* rgb, void* ws,
                                   size_t ws_bytes, void* stream) {
  CNEUS_NVTX_RANGE();
  if (P <= 0) return CNEUS_OK;  // empty input: nothing to do (empty torch tensors carry null pointers)
  REQUIRE(desc && packed && pts && feats && rgb, "color_forward: null argument");
  REQUIRE(desc->color_n_lin > 0, "color_forward: descriptor has no colour network");
  REQUIRE(normals || desc->color_mode == CNEUS_COLOR_NO_NORMAL, "color_forward: normals required for this mode");
  REQUIRE(view_dirs || desc->color_mode == CNEUS_COLOR_NO_VIEW_DIR, "color_forward: view_dirs required for this mode");
  NetPack np;
  CHECK_RC(build_netpack(desc, &np));
  (void)ws; (void)ws_bytes;
  ShadeArgs a = blank_args();
  a.src_mode = 0; a.pts = pts; a.P = P; a.in_normals = normals; a.in_viewdirs = view_dirs; a.in_feats = feats;
  a.run_color = 1; a.out_color = rgb;
  return launch_shade(np, (const float*)packed, a, shade_grid_for(P), (cudaStream_t)stream);
}

extern "C" int cneus_relight_forward(const CneusNetDesc* desc, const void* packed, const float* rgb, const float* pts,
                                     const float* dirs, const float* grads, int64_t P, float* rgb_out, float* drgb_out,
                                     void* ws, size_t ws_bytes, void* stream) {
  CNEUS_NVTX_RANGE();
  if (P <= 0) return CNEUS_OK;  // empty input: nothing to do (empty torch tensors carry null pointers)
  REQUIRE(desc && packed && rgb && pts && dirs && rgb_out && drgb_out, "relight_forward: null argument");
  REQUIRE(desc->has_relight, "relight_forward: descriptor has no relight network");
  REQUIRE(grads || !desc->relight_include_grad, "relight_forward: gradients required (INCLUDE_GRAD)");
  NetPack np;
  CHECK_RC(build_netpack(desc, &np));
  (void)ws; (void)ws_bytes;
  ShadeArgs a = blank_args();
  a.src_mode = 0; a.pts = pts; a.P = P; a.in_normals = grads; a.in_viewdirs = dirs; a.in_rgb = rgb;
  a.run_relight = 1; a.out_relit = rgb_out; a.out_drgb = drgb_out;
  return launch_shade(np, (const float*)packed, a, shade_grid_for(P), (cudaStream_t)stream);
}

extern "C" int cneus_up_sample(const float* rays_o, const float* rays_d, const float* z, const float* sdf, int64_t B,
                               int32_t n, int32_t m, float inv_s, const float* u, float* new_z, void* stream) {
  CNEUS_NVTX_RANGE();
  if (B <= 0) return CNEUS_OK;  // empty input: nothing to do (empty torch tensors carry null pointers)
  REQUIRE(rays_o && rays_d && z && sdf && u && new_z, "up_sample: null argument");
  return launch_up_sample(rays_o, rays_d, z, sdf, B, n, m, inv_s, u, new_z, (cudaStream_t)stream);
}

static int sdf_on_rays(const NetPack& np, const float* packed, const float* ro, const float* rd, const float* t,
                       int64_t B, int n, float* sdf_out, float* dscratch, cudaStream_t st) {
  ShadeArgs a = blank_args();
  a.dscratch = dscratch;
  a.src_mode = 1; a.rays_o = ro; a.rays_d = rd; a.t = t; a.n_per_ray = n; a.P = B * n;
  a.run_sdf = 1; a.out_sdf = sdf_out;
  return launch_shade(np, packed, a, shade_grid_for(a.P), st);
}

extern "C" int cneus_cat_z_vals(const CneusNetDesc* desc, const void* packed, const float* rays_o, const float* rays_d,
                                const float* z, const float* new_z, const float* sdf, int64_t B, int32_t n, int32_t m,
                                int32_t last, float* z_out, float* sdf_out, void* ws, size_t ws_bytes, void* stream) {
  CNEUS_NVTX_RANGE();
  if (B <= 0) return CNEUS_OK;  // empty input: nothing to do (empty torch tensors carry null pointers)
  REQUIRE(desc && packed && rays_o && rays_d && z && new_z && z_out, "cat_z_vals: null argument");
  REQUIRE(last || (sdf && sdf_out), "cat_z_vals: sdf / sdf_out required unless last");
  NetPack np;
  CHECK_RC(build_netpack(desc, &np));
  cudaStream_t st = (cudaStream_t)stream;
  if (last) return launch_merge(z, new_z, nullptr, nullptr, B, n, m, z_out, nullptr, st);
  Workspace w;
  CHECK_RC(carve(np, ws, ws_bytes, (size_t)B * m, &w));
  CHECK_RC(sdf_on_rays(np, (const float*)packed, rays_o, rays_d, new_z, B, m, w.ray, w.dscratch, st));
  return launch_merge(z, new_z, sdf, w.ray, B, n, m, z_out, sdf_out, st);
}

extern "C" int cneus_sample_z(const CneusNetDesc* desc, const void* packed, const float* rays_o, const float* rays_d,
                              const float* near, const float* far, const float* t_rand, const float* lin, const float* u,
                              int64_t B, int32_t n_samples, int32_t n_importance, int32_t up_steps, float* z_out, void* ws,
                              size_t ws_bytes, void* stream) {
  CNEUS_NVTX_RANGE();
  if (B <= 0) return CNEUS_OK;  // empty input: nothing to do (empty torch tensors carry null pointers)
  REQUIRE(desc && packed && rays_o && rays_d && near && far && lin && z_out, "sample_z: null argument");
  REQUIRE(n_samples >= 2, "sample_z: n_samples must be >= 2");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_importance <= 0) return launch_coarse_z(near, far, t_rand, lin, B, n_samples, z_out, st);
  REQUIRE(u != nullptr, "sample_z: u (inverse-CDF abscissae) required");
  REQUIRE(up_steps >= 1 && n_importance % up_steps == 0, "sample_z: n_importance must be divisible by up_steps");
  NetPack np;
  CHECK_RC(build_netpack(desc, &np));
  const int S = n_samples + n_importance, m = n_importance / up_steps;
  Workspace w;
  CHECK_RC(carve(np, ws, ws_bytes, (size_t)B * 6 * S, &w));
  float* zA = w.ray;
  float* zB = zA + (size_t)B * S;
  float* sA = zB + (size_t)B * S;
  float* sB = sA + (size_t)B * S;
  float* nz = sB + (size_t)B * S;
  float* nsdf = nz + (size_t)B * S;
  const float* pk = (const float*)packed;
  CHECK_RC(launch_coarse_z(near, far, t_rand, lin, B, n_samples, zA, st));
  CHECK_RC(sdf_on_rays(np, pk, rays_o, rays_d, zA, B, n_samples, sA, w.dscratch, st));
  int n = n_samples;
  for (int i = 0; i < up_steps; ++i) {  // NeuS.py:347-355: inv_s = 64 * 2**i
    const bool last = (i + 1 == up_steps);
    CHECK_RC(launch_up_sample(rays_o, rays_d, zA, sA, B, n, m, 64.0f * (float)(1 << i), u, nz, st));
    if (!last) {
      CHECK_RC(sdf_on_rays(np, pk, rays_o, rays_d, nz, B, m, nsdf, w.dscratch, st));
      CHECK_RC(launch_merge(zA, nz, sA, nsdf, B, n, m, zB, sB, st));
      float* t = zA; zA = zB; zB = t;
      t = sA; sA = sB; sB = t;
    } else {
      CHECK_RC(launch_merge(zA, nz, nullptr, nullptr, B, n, m, z_out, nullptr, st));
    }
    n += m;
  }
  return CNEUS_OK;
}

extern "C" int cneus_render_core(const CneusNetDesc* desc, const void* packed, const float* variance, const float* rays_o,
                                 const float* rays_d, const float* z, int64_t B, int32_t S, float sample_dist,
                                 float cos_anneal_ratio, const CneusRenderOut* out, void* ws, size_t ws_bytes,
                                 void* stream) {
  CNEUS_NVTX_RANGE();
  REQUIRE(desc && packed && variance && rays_o && rays_d && z && out, "render_core: null argument");
  REQUIRE(out->gradients && out->sdf && out->sampled_color && out->mid_z && out->dists && out->scalars,
          "render_core: required output pointer is null");
  REQUIRE(desc->color_n_lin > 0, "render_core: descriptor has no colour network");
  REQUIRE(!desc->has_relight || out->global_sampled, "render_core: global_sampled required with a relight network");
  NetPack np;
  CHECK_RC(build_netpack(desc, &np));
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w;
  CHECK_RC(carve(np, ws, ws_bytes, (size_t)B * 2 + 64, &w));
  CHECK_RC(launch_sections(z, B, S, sample_dist, out->mid_z, out->dists, st));
  ShadeArgs a = blank_args();
  a.src_mode = 1; a.rays_o = rays_o; a.rays_d = rays_d; a.t = out->mid_z; a.n_per_ray = S; a.P = B * (int64_t)S;
  a.run_sdf = 2; a.run_grad = 1; a.run_color = 1; a.run_relight = desc->has_relight ? 1 : 0;
  a.out_sdf = out->sdf; a.out_grad = out->gradients; a.dscratch = w.dscratch;
  if (desc->has_relight) { a.out_color = out->global_sampled; a.out_relit = out->sampled_color; a.out_drgb = out->delta_relight; }
  else { a.out_color = out->sampled_color; }
  CHECK_RC(launch_shade(np, (const float*)packed, a, shade_grid_for(a.P), st));
  CneusRenderOut o = *out;
  if (!desc->has_relight) o.global_color = nullptr;
  return launch_composite(variance, rays_o, rays_d, z, B, S, cos_anneal_ratio, o, w.ray, st);
}

extern "C" int cneus_sdf_grid(const CneusNetDesc* desc, const void* packed, const float* xs, const float* ys,
                              const float* zs, int32_t res, int64_t lin_begin, int64_t lin_end, float* u, void* ws,
                              size_t ws_bytes, void* stream) {
  CNEUS_NVTX_RANGE();
  REQUIRE(desc && packed && xs && ys && zs && u, "sdf_grid: null argument");
  REQUIRE(res >= 1 && lin_begin >= 0 && lin_end >= lin_begin && lin_end <= (int64_t)res * res * res, "sdf_grid: bad range");
  NetPack np;
  CHECK_RC(build_netpack(desc, &np));
  Workspace w;
  CHECK_RC(carve(np, ws, ws_bytes, 0, &w));
  ShadeArgs a = blank_args();
  a.dscratch = w.dscratch;
  a.src_mode = 2; a.gx = xs; a.gy = ys; a.gz = zs; a.res = res; a.lin_begin = lin_begin; a.P = lin_end - lin_begin;
  a.run_sdf = 1; a.out_sdf = u; a.out_sdf_sign = -1.0f;
  return launch_shade(np, (const float*)packed, a, shade_grid_for(a.P), (cudaStream_t)stream);
}

extern "C" int cneus_vertex_color(const CneusNetDesc* desc, const void* packed, const float* vertices, int64_t V,
                                  float* rgb, void* ws, size_t ws_bytes, void* stream) {
  CNEUS_NVTX_RANGE();
  if (V <= 0) return CNEUS_OK;  // empty input: nothing to do (empty torch tensors carry null pointers)
  REQUIRE(desc && packed && vertices && rgb, "vertex_color: null argument");
  REQUIRE(desc->color_n_lin > 0, "vertex_color: descriptor has no colour network");
  NetPack np;
  CHECK_RC(build_netpack(desc, &np));
  Workspace w;
  CHECK_RC(carve(np, ws, ws_bytes, 0, &w));
  ShadeArgs a = blank_args();
  a.src_mode = 0; a.pts = vertices; a.P = V; a.run_sdf = 2; a.run_grad = 1; a.run_color = 1; a.viewdir_mode = 1;
  a.out_color = rgb; a.dscratch = w.dscratch;
  return launch_shade(np, (const float*)packed, a, shade_grid_for(V), (cudaStream_t)stream);
}

namespace cneus { void profile_enable(int on); int profile_read(int kind, double* total_ms, int64_t* launches); }
extern "C" void cneus_profile_enable(int on) { cneus::profile_enable(on); }
extern "C" int cneus_profile_read(int kind, double* total_ms, int64_t* launches) { return cneus::profile_read(kind, total_ms, launches); }

extern "C" void cneus_force_simt(int on) { cneus::g_force_simt = on ? 1 : 0; }
