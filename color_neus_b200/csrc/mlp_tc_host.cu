// Tensor-core path, host side: weight images (pack kernel), step program, launch.
#include <cuda_fp16.h>
#include <string.h>

#include "mlp_tc.cuh"

namespace cneus {

// ---------------------------------------------------------------------------------------------------------
// weight images: [kb][nh][hi slab | lo slab], slab element (r, kk) at SWIZZLE_128B position
// ---------------------------------------------------------------------------------------------------------
struct TcPackJob {
  const float* v;        // source weight_v / weight  [src_out][src_in]
  int64_t scale_off;     // float offset of the per-row weight-norm scale in the packed buffer
  int64_t dst_off;       // byte offset of the first stage image
  int32_t src_in;
  int32_t transposed;    // 0: image row = output n, image k = input col ; 1: image row = input col, image k = output n
  int32_t row_start;     // forward: first source row ; transposed: unused
  int32_t n_valid;       // valid image rows (outputs for forward, inputs for transposed)
  int32_t n_halves, n_kb;
  int32_t kstart[5];     // forward: first source column of the K-block ; transposed: first source row (output)
  int32_t kvalid[5];
};
constexpr int MAX_TC_JOBS = 28;
struct TcPackJobs { TcPackJob j[MAX_TC_JOBS]; int32_t n; };

__global__ void pack_tc_kernel(const __grid_constant__ TcPackJobs jobs, float* packed) {
  const TcPackJob& J = jobs.j[blockIdx.y];
  uint8_t* base = reinterpret_cast<uint8_t*>(packed) + J.dst_off;
  // image = [kb][sh] stages of 32 KB: hi slab then lo slab, each [256 rows][32 halfs] in the SWIZZLE_64B layout
  const int64_t total = (int64_t)J.n_kb * 2 * 256 * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i & 31);
    const int r = (int)((i >> 5) & 255);
    const int stage = (int)(i >> 13);  // kb * 2 + sh
    const int kb = stage >> 1, sh = stage & 1;
    const int kin = sh * 32 + kk;      // position inside the 64-wide K-block
    float val = 0.0f;
    if (r < J.n_valid && kin < J.kvalid[kb]) {
      int srow, scol;
      if (!J.transposed) { srow = J.row_start + r; scol = J.kstart[kb] + kin; }
      else { srow = J.kstart[kb] + kin; scol = r; }
      val = J.v[(int64_t)srow * J.src_in + scol] * packed[J.scale_off + srow] * W_SCALE;
    }
    const __half h = __float2half_rn(val);
    const __half l = __float2half_rn(val - __half2float(h));
    const int chunk = (kk >> 3) ^ ((r >> 1) & 3);
    const size_t off = (size_t)stage * STAGE_BYTES + (size_t)(r >> 3) * 512 + (size_t)(r & 7) * 64 + (size_t)chunk * 16 + (size_t)(kk & 7) * 2;
    *reinterpret_cast<__half*>(base + off) = h;
    *reinterpret_cast<__half*>(base + off + SLAB_BYTES) = l;
  }
}

}  // namespace cneus

// ---------------------------------------------------------------------------------------------------------
// host side: weight-image jobs, step program, launch
// ---------------------------------------------------------------------------------------------------------
namespace cneus {

int g_force_simt = 0;
int g_tc_prof_on = 0;

static int ceil16(int x) { return (x + 15) / 16; }

int pack_tc_weights(const NetPack& np, const CneusParams* P, const int64_t* scale_off, float* packed, cudaStream_t st) {
  if (!np.tc_eligible) return CNEUS_OK;
  const CneusNetDesc& d = np.d;
  const int nl = d.sdf_n_lin, cn = d.color_n_lin;
  TcPackJobs* jobs = new TcPackJobs();
  memset(jobs, 0, sizeof(*jobs));
  auto add = [&](const CneusLinear& S, int64_t soff, int64_t dst, bool transposed, int row_start, int n_valid, int n_halves,
                 int n_kb, const int* kstart, const int* kvalid) {
    TcPackJob& j = jobs->j[jobs->n++];
    j.v = S.weight_v; j.scale_off = soff; j.dst_off = dst; j.src_in = S.in; j.transposed = transposed ? 1 : 0;
    j.row_start = row_start; j.n_valid = n_valid; j.n_halves = n_halves; j.n_kb = n_kb;
    for (int i = 0; i < n_kb; ++i) { j.kstart[i] = kstart[i]; j.kvalid[i] = kvalid[i]; }
  };
  int li = 0;
  for (int l = 0; l < nl; ++l, ++li) {
    const CneusLinear& S = P->sdf[l];
    const int in = S.in;  // 39 or 256
    int ks[5], kv[5];
    const int kbs = (l == 0) ? 1 : 4;
    for (int i = 0; i < kbs; ++i) { ks[i] = 64 * i; kv[i] = (in - 64 * i) < 64 ? (in - 64 * i) : 64; }
    if (l < nl - 1) {
      add(S, scale_off[li], np.tc_sdf_fwd[l], false, 0, S.out, 2, kbs, ks, kv);
      // gradient chain: image rows = inputs, K' = outputs
      int bs[5], bv[5];
      for (int i = 0; i < 4; ++i) { bs[i] = 64 * i; int rem = S.out - 64 * i; bv[i] = rem < 0 ? 0 : (rem < 64 ? rem : 64); }
      add(S, scale_off[li], np.tc_sdf_bwd[l], true, 0, in, (l == 0) ? 1 : 2, 4, bs, bv);
    } else {
      add(S, scale_off[li], np.tc_sdf_fwd[l], false, 1, S.out - 1, 2, kbs, ks, kv);  // feature rows 1..256
    }
  }
  for (int l = 0; l < cn; ++l, ++li) {
    if (l >= cn - 1) continue;
    const CneusLinear& S = P->color[l];
    int ks[5], kv[5];
    if (l == 0) {
      for (int i = 0; i < 4; ++i) { ks[i] = np.color_k0v + 64 * i; kv[i] = 64; }
      ks[4] = 0; kv[4] = np.color_k0v;
      add(S, scale_off[li], np.tc_color[l], false, 0, S.out, 2, 5, ks, kv);
    } else {
      for (int i = 0; i < 4; ++i) { ks[i] = 64 * i; kv[i] = 64; }
      add(S, scale_off[li], np.tc_color[l], false, 0, S.out, 2, 4, ks, kv);
    }
  }
  if (d.has_relight) {
    {
      int ks[1] = {0}, kv[1] = {np.relight_k0v};
      add(P->relight_in, scale_off[li], np.tc_rl_in, false, 0, P->relight_in.out, 2, 1, ks, kv);
      ++li;
    }
    for (int i = 0; i < d.relight_n_layers; ++i, ++li) {
      if (i >= d.relight_n_layers - 1) continue;
      const CneusLinear& S = P->relight_mlp[i];
      int ks[5], kv[5];
      if (i == d.relight_y_in_layer - 1) {
        for (int q = 0; q < 4; ++q) { ks[q] = 3 + 64 * q; kv[q] = 64; }
        ks[4] = 0; kv[4] = 3;
        add(S, scale_off[li], np.tc_rl[i], false, 0, S.out, 2, 5, ks, kv);
      } else {
        for (int q = 0; q < 4; ++q) { ks[q] = 64 * q; kv[q] = 64; }
        add(S, scale_off[li], np.tc_rl[i], false, 0, S.out, 2, 4, ks, kv);
      }
    }
  }
  pack_tc_kernel<<<dim3(160, jobs->n), 256, 0, st>>>(*jobs, packed);
  delete jobs;
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return CNEUS_OK;
}

bool tc_supports(const NetPack& np, const ShadeArgs& a) {
  if (!np.tc_eligible || g_force_simt) return false;
  if (a.run_tangent) return a.dump.on && a.tan_t0 && a.tan_amax;
  if (a.in_normals || a.in_viewdirs || a.in_feats || a.in_rgb) return false;  // stand-alone sub-module calls
  if (!a.run_sdf) return false;
  if (a.run_color && (!a.run_grad || a.run_sdf != 2)) return false;  // one code path: colour always follows the gradient
  if (a.run_relight && !a.run_color) return false;
  if (a.out_full && a.run_sdf != 2) return false;
  return true;
}

size_t tc_scratch_floats_per_cta(const NetPack& np) { return (size_t)np.d.sdf_n_lin * 256 * TCM + (size_t)TC_GXS_ROWS * TCM; }

static void build_program(const NetPack& np, const ShadeArgs& a, TcProgram* pg) {
  memset(pg, 0, sizeof(*pg));
  const CneusNetDesc& d = np.d;
  const int nl = d.sdf_n_lin, nh = nl - 1;
  pg->n_hidden = nh; pg->multires = d.sdf_multires; pg->pe_dim = np.pe_dim; pg->sdf_scale = d.sdf_scale;
  pg->seed_row_off = (int32_t)np.sdf_row.w_off;
  pg->feat_bias_off = (int32_t)np.sdf[nl - 1].bias_off;
  pg->feat_inv_scale = 1.0f / W_SCALE;
  pg->color_mode = d.color_mode; pg->color_multires_view = d.color_multires_view; pg->color_squeeze = d.color_squeeze_out;
  pg->relight_multires_view = d.relight_multires_view; pg->relight_include_grad = d.relight_include_grad;
  pg->relight_inv_sigmoid = d.relight_inv_sigmoid; pg->has_skip = d.sdf_skip >= 0 ? 1 : 0;
  // few-input blocks as fp32 rank updates in the epilogue (frees the small-input slab for a third weight stage)
  const bool rank_path = (!a.run_color || d.color_mode == CNEUS_COLOR_NO_VIEW_DIR) &&
                         (!a.run_relight || d.relight_y_in_layer - 1 == d.relight_n_layers - 2);
  pg->n_stages = rank_path ? 3 : 2;
  int n = 0;
  auto base = [&](int64_t w_off, int n_kb, int n_halves) -> TcStep& {
    TcStep& S = pg->s[n++];
    S.w_off = w_off; S.bias_off = -1; S.row_off = -1; S.row_bias_off = -1; S.row_n = 0; S.n_valid = 256;
    S.n_kb = (int8_t)n_kb; S.n_halves = (int8_t)n_halves; S.acc = 0; S.epi = EPI_HIDDEN; S.act = TACT_RELU;
    S.prep_next = PREP_NONE; S.post = POST_NONE; S.flags = 0; S.d_layer = -1; S.n_small = 0; S.small_off = 0;
    for (int i = 0; i < 5; ++i) { S.slab[i] = (int8_t)i; S.ksteps[i] = 4; }
    S.inv_scale = 1.0f / W_SCALE; S.out_scale = 1.0f;
    return S;
  };
  if (a.run_tangent) {
    // tangent pass: the forward weights again, no bias, "activation" = multiplication by the stored softplus'
    for (int l = 0; l < nh; ++l) {
      TcStep& S = base(np.tc_sdf_fwd[l], l == 0 ? 1 : 4, 2);
      S.epi = EPI_TAN; S.d_layer = (int8_t)l; S.n_valid = (int16_t)np.sdf[l].N;
      if (l == 0) S.ksteps[0] = (int8_t)ceil16(np.pe_dim);
      if (l + 1 == d.sdf_skip) { S.flags |= TF_FEEDS_SKIP; S.out_scale = 0.70710678118654752440f; }
    }
    pg->n_steps = n;
    pg->tangent = 1;
    pg->prof = g_tc_prof_on;
    return;
  }
  // ---- SDF forward
  for (int l = 0; l < nh; ++l) {
    TcStep& S = base(np.tc_sdf_fwd[l], l == 0 ? 1 : 4, 2);
    S.bias_off = (int32_t)np.sdf[l].bias_off; S.act = TACT_SOFTPLUS; S.n_valid = (int16_t)np.sdf[l].N;
    if (l == 0) S.ksteps[0] = (int8_t)ceil16(np.pe_dim);
    if (l + 1 == d.sdf_skip) { S.flags |= TF_FEEDS_SKIP; S.out_scale = 0.70710678118654752440f; }
    if (a.run_grad) S.d_layer = (int8_t)l;
    if (l == nh - 1) {
      S.row_off = (int32_t)np.sdf_row.w_off; S.row_bias_off = (int32_t)np.sdf_row.bias_off; S.row_n = 1; S.post = POST_SDF;
      if (a.run_grad && a.run_sdf != 2) S.prep_next = PREP_SEED;
    }
  }
  if (a.run_sdf == 2) {
    TcStep& S = base(np.tc_sdf_fwd[nl - 1], 4, 2);
    S.epi = EPI_PARK;
    if (a.run_grad) S.prep_next = PREP_SEED;
  }
  if (a.run_grad) {
    for (int l = nh - 1; l >= 0; --l) {
      // K' = outputs of layer l (ga rows), N' = its inputs
      const int outs = np.sdf[l].N;
      TcStep& S = base(np.tc_sdf_bwd[l], 4, l == 0 ? 1 : 2);
      for (int i = 0; i < 4; ++i) { int rem = outs - 64 * i; S.ksteps[i] = (int8_t)(rem <= 0 ? 0 : ceil16(rem < 64 ? rem : 64)); }
      S.inv_scale = 1.0f / (W_SCALE * BWD_ASCALE);
      if (l > 0) {
        S.epi = EPI_BWD; S.d_layer = (int8_t)(l - 1); S.n_valid = (int16_t)np.sdf[l - 1].N; S.out_scale = BWD_ASCALE;
        if (l == d.sdf_skip) S.flags |= TF_SKIP_BWD;
      } else {
        S.epi = EPI_BWD_LAST;
        if (a.run_color) S.prep_next = PREP_COLOR_IN;
      }
    }
  }
  if (a.run_color) {
    const int cn = d.color_n_lin;
    for (int l = 0; l < cn - 1; ++l) {
      TcStep& S = base(np.tc_color[l], (l == 0 && !rank_path) ? 5 : 4, 2);
      S.bias_off = (int32_t)np.color[l].bias_off;
      if (l == 0 && !rank_path) { S.slab[4] = SMALL_SLAB; S.ksteps[4] = (int8_t)ceil16(np.color_k0v); }
      if (l == 0 && rank_path) { S.n_small = 6; S.small_off = (int32_t)np.color[0].wt_off; }
      if (l == cn - 2) {
        S.row_off = (int32_t)np.color_row.w_off; S.row_bias_off = (int32_t)np.color_row.bias_off; S.row_n = 3; S.post = POST_CG;
        if (a.run_relight) S.prep_next = PREP_RELIGHT_IN;
      }
    }
  }
  if (a.run_relight) {
    const int rn = d.relight_n_layers, y = d.relight_y_in_layer;
    {
      TcStep& S = base(np.tc_rl_in, 1, 2);
      S.bias_off = (int32_t)np.rl_in.bias_off; S.slab[0] = rank_path ? 0 : SMALL_SLAB; S.ksteps[0] = (int8_t)ceil16(np.relight_k0v);
      if (y - 1 == 0 && !rank_path) S.prep_next = PREP_CG;
    }
    for (int i = 0; i < rn - 1; ++i) {
      const bool yin = (i == y - 1);
      TcStep& S = base(np.tc_rl[i], (yin && !rank_path) ? 5 : 4, 2);
      S.bias_off = (int32_t)np.rl[i].bias_off;
      if (yin && !rank_path) { S.slab[4] = SMALL_SLAB; S.ksteps[4] = 1; }
      if (yin && rank_path) { S.n_small = 3; S.small_off = (int32_t)np.rl[i].wt_off; }
      if (i + 1 == y - 1 && !rank_path) S.prep_next = PREP_CG;
      if (i == rn - 2) { S.row_off = (int32_t)np.rl_row.w_off; S.row_bias_off = (int32_t)np.rl_row.bias_off; S.row_n = 3; S.post = POST_DRGB; }
    }
  }
  pg->n_steps = n;
  pg->prof = g_tc_prof_on;
}



int launch_shade_tc(const NetPack& np, const float* packed, const ShadeArgs& a, float* gxscratch, cudaStream_t st) {
  static bool attr_set[CNEUS_MAX_DEVICES] = {false};
  if (first_use_on_device(attr_set)) {
    CNEUS_CUDA_CHECK(cudaFuncSetAttribute(shade_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    CNEUS_CUDA_CHECK(cudaFuncSetAttribute(shade_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
  }
  if (a.P <= 0) return CNEUS_OK;
  if (a.dscratch == nullptr || gxscratch == nullptr) { set_error("tensor-core shading needs the per-CTA scratch (workspace)"); return CNEUS_EINVAL; }
  TcProgram pg;
  build_program(np, a, &pg);
  if (pg.n_steps > MAX_TC_STEPS) { set_error("tensor-core program too long"); return CNEUS_EUNSUPPORTED; }
  int64_t tiles = (a.P + TCM - 1) / TCM;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  if (kPair) {
    // CTA pairs: clusters of two CTAs (the two SMs of a TPC), each pair walks pairs of tiles
    const int64_t tile_pairs = (tiles + 1) / 2;
    const int grid = 2 * (int)(tile_pairs < sms / 2 ? tile_pairs : sms / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_KERNEL_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    static bool cluster_checked[CNEUS_MAX_DEVICES] = {false};
    if (first_use_on_device(cluster_checked)) {  // fail loudly where a pair cannot be co-scheduled (no silent fallback)
      int n_clusters = 0;
      CNEUS_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&n_clusters, shade_tc_kernel<false>, &cfg));
      if (n_clusters < 1) { set_error("this device cannot co-schedule a cluster of two CTAs with 227 KB of shared memory each"); return CNEUS_EUNSUPPORTED; }
    }
    if (a.dump.on) CNEUS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, shade_tc_kernel<true>, pg, packed, a, gxscratch));
    else CNEUS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, shade_tc_kernel<false>, pg, packed, a, gxscratch));
  } else {
    const int grid = (int)(tiles < sms ? tiles : sms);
    if (a.dump.on) shade_tc_kernel<true><<<grid, TC_KERNEL_THREADS, TC_SMEM_BYTES, st>>>(pg, packed, a, gxscratch);
    else shade_tc_kernel<false><<<grid, TC_KERNEL_THREADS, TC_SMEM_BYTES, st>>>(pg, packed, a, gxscratch);
  }
  CNEUS_CUDA_CHECK(cudaGetLastError());
  return CNEUS_OK;
}

}  // namespace cneus

extern "C" void cneus_tc_prof_enable(int on) { cneus::g_tc_prof_on = on; }  // 1 + epilogue warp to trace; >= 100: weight-wait distribution
