// Weight packing: effective weight-norm matrices (fields.py:72-73, 152-153) re-laid-out for the kernels.
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace cneus {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[CNEUS_MAX_DEVICES] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  const bool slot = dev >= 0 && dev < CNEUS_MAX_DEVICES;
  if (slot && cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  if (slot) cached[dev] = n;
  return n;
}

static int view_dim(int multires) { return multires > 0 ? 3 * (1 + 2 * multires) : 3; }

int build_netpack(const CneusNetDesc* d, NetPack* np) {
  memset(np, 0, sizeof(*np));
  np->d = *d;
  const int H = d->sdf_d_hidden, nl = d->sdf_n_lin;
  if (nl < 2 || nl > CNEUS_MAX_SDF_LIN) { set_error("sdf_n_lin %d out of range", nl); return CNEUS_EINVAL; }
  if (H < 64 || H > MAXH || H % 64) { set_error("sdf_d_hidden %d must be a multiple of 64 in [64,256]", H); return CNEUS_EUNSUPPORTED; }
  if (d->sdf_d_out < 2 || d->sdf_d_out - 1 > MAXH || (d->sdf_d_out - 1) % 16) {
    set_error("sdf_d_out %d unsupported (feature width must be a multiple of 16, <= 256)", d->sdf_d_out);
    return CNEUS_EUNSUPPORTED;
  }
  const int pe = view_dim(d->sdf_multires);
  if (pe > SMALLK) { set_error("sdf_multires %d too large", d->sdf_multires); return CNEUS_EUNSUPPORTED; }
  if (d->sdf_skip >= 0 && (d->sdf_skip < 1 || d->sdf_skip > nl - 1 || H - pe <= 0)) {
    set_error("sdf_skip %d unsupported", d->sdf_skip);
    return CNEUS_EUNSUPPORTED;
  }
  np->pe_dim = pe;
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off += (n + 63) / 64 * 64; return o; };

  for (int l = 0; l < nl; ++l) {
    PLayer& L = np->sdf[l];
    const int in = (l == 0) ? pe : H;
    int out = (l == nl - 1) ? d->sdf_d_out - 1 : ((l + 1 == d->sdf_skip) ? H - pe : H);
    L.K0 = 0; L.k0v = 0;
    L.k1v = in; L.K1 = pad_to(in, KC);
    L.N = out; L.Np = pad_to(out, 64);
    L.wt_off = take((int64_t)(L.K0 + L.K1) * L.Np);
    L.bias_off = take(L.Np);
    if (l < nl - 1) {
      L.Nb = pad_to(out, KC); L.Kb = pad_to(in, 64);
      L.wb_off = take((int64_t)L.Nb * L.Kb);
    } else {
      L.wb_off = -1;
    }
  }
  np->sdf_row.K0 = 0; np->sdf_row.K1 = pad_to(H, KC); np->sdf_row.N = 1;
  np->sdf_row.w_off = take(np->sdf_row.K1); np->sdf_row.bias_off = take(1);

  // ---- colour network
  const int cn = d->color_n_lin, Hc = d->color_d_hidden;
  if (cn > 0) {
    if (cn < 2 || cn > CNEUS_MAX_COLOR_LIN || Hc < 64 || Hc > MAXH || Hc % 64) { set_error("colour net shape unsupported"); return CNEUS_EUNSUPPORTED; }
    if (d->color_d_feature != d->sdf_d_out - 1) { set_error("color_d_feature must equal sdf_d_out-1"); return CNEUS_EINVAL; }
    const int vd = view_dim(d->color_multires_view);
    int k0v = 3;
    if (d->color_mode == CNEUS_COLOR_IDR) k0v += vd + 3;
    else if (d->color_mode == CNEUS_COLOR_NO_VIEW_DIR) k0v += 3;
    else if (d->color_mode == CNEUS_COLOR_NO_NORMAL) k0v += vd;
    else { set_error("no such mode: %d", d->color_mode); return CNEUS_EINVAL; }
    if (k0v > SMALLK) { set_error("colour input too wide"); return CNEUS_EUNSUPPORTED; }
    np->color_k0v = k0v;
    for (int l = 0; l < cn - 1; ++l) {
      PLayer& L = np->color[l];
      if (l == 0) { L.k0v = k0v; L.K0 = pad_to(k0v, KC); L.k1v = d->color_d_feature; L.K1 = pad_to(L.k1v, KC); }
      else { L.k0v = 0; L.K0 = 0; L.k1v = Hc; L.K1 = Hc; }
      L.N = Hc; L.Np = Hc; L.wb_off = -1;
      L.wt_off = take((int64_t)(L.K0 + L.K1) * L.Np);
      L.bias_off = take(L.Np);
    }
    np->color_row.K0 = 0; np->color_row.K1 = Hc; np->color_row.N = 3;
    np->color_row.w_off = take(3 * Hc); np->color_row.bias_off = take(3);
  }
  // ---- relight network
  if (d->has_relight) {
    const int n = d->relight_n_layers, y = d->relight_y_in_layer, Hr = d->relight_d_hidden;
    if (n < 1 || n > CNEUS_MAX_RELIGHT_LIN || Hr < 64 || Hr > MAXH || Hr % 64) { set_error("relight net shape unsupported"); return CNEUS_EUNSUPPORTED; }
    int k0v = 3 + view_dim(d->relight_multires_view) + (d->relight_include_grad ? 3 : 0);
    if (k0v > SMALLK) { set_error("relight input too wide"); return CNEUS_EUNSUPPORTED; }
    np->relight_k0v = k0v;
    PLayer& I = np->rl_in;
    I.k0v = k0v; I.K0 = pad_to(k0v, KC); I.k1v = 0; I.K1 = 0; I.N = Hr; I.Np = Hr; I.wb_off = -1;
    I.wt_off = take((int64_t)I.K0 * I.Np); I.bias_off = take(I.Np);
    for (int i = 0; i < n - 1; ++i) {
      PLayer& L = np->rl[i];
      if (i == y - 1) { L.k0v = 3; L.K0 = KC; } else { L.k0v = 0; L.K0 = 0; }
      L.k1v = Hr; L.K1 = Hr; L.N = Hr; L.Np = Hr; L.wb_off = -1;
      L.wt_off = take((int64_t)(L.K0 + L.K1) * L.Np); L.bias_off = take(L.Np);
    }
    RowLayer& R = np->rl_row;
    R.K0 = (y == n) ? KC : 0; R.K1 = Hr; R.N = 3;
    R.w_off = take(3 * (R.K0 + R.K1)); R.bias_off = take(3);
  }
  // ---- tensor-core weight images (mlp_tc.cu): 256-wide stacks only
  np->tc_eligible = (H == 256 && d->sdf_d_out == 257 && nl - 1 <= 8 && pe <= 64 &&
                     (d->sdf_skip < 0 || (d->sdf_skip >= 1 && d->sdf_skip <= nl - 2)) &&
                     (cn == 0 || (Hc == 256 && d->color_d_feature == 256)) &&
                     (!d->has_relight || (d->relight_d_hidden == 256 && d->relight_y_in_layer != d->relight_n_layers &&
                                          d->relight_n_layers >= 2)))
                        ? 1 : 0;
  if (np->tc_eligible) {
    off = (off + 255) / 256 * 256;  // 1024-byte alignment of every stage image
    const int64_t stage_floats = TC_STAGE_BYTES_HOST / 4;
    auto take_stages = [&](int n_stages) { int64_t o = off * 4; off += (int64_t)n_stages * stage_floats; return o; };
    for (int l = 0; l < nl; ++l) {
      const int kbs = (l == 0) ? 1 : 4;
      np->tc_sdf_fwd[l] = take_stages(kbs * 2);
      if (l < nl - 1) np->tc_sdf_bwd[l] = take_stages(8);  // K' = outputs (<=256): 4 K-blocks x 2 half-block stages
    }
    for (int l = 0; l < cn - 1; ++l) np->tc_color[l] = take_stages((l == 0 ? 5 : 4) * 2);
    if (d->has_relight) {
      np->tc_rl_in = take_stages(2);
      for (int i = 0; i < d->relight_n_layers - 1; ++i)
        np->tc_rl[i] = take_stages((i == d->relight_y_in_layer - 1 ? 5 : 4) * 2);
    }
  }
  np->total_floats = off;
  return CNEUS_OK;
}

// -------------------------------------------------------------------------------------------------------------
struct PackJob {
  const float* v;
  const float* g;
  const float* bias;
  int64_t dst_off;
  int64_t scale_off;
  int32_t src_out, src_in;
  int32_t mode;  // 0 Wt[k][n], 1 Wb[n][k], 2 bias[n], 3 rows [n][K0+K1]
  int32_t row_start, n_valid, Np;
  int32_t k0v, K0, k1v, K1;
  int32_t Nb, Kb;
};
constexpr int MAX_JOBS = 40;
struct PackJobs { PackJob j[MAX_JOBS]; int32_t n; };
struct ScaleJob { const float* v; const float* g; int64_t scale_off; int32_t out, in; };
struct ScaleJobs { ScaleJob j[32]; int32_t n; };

// one warp per source row: s = g / ||v_row||  (1 when the layer is not weight-normed)
__global__ void wn_scale_kernel(ScaleJobs jobs, float* packed) {
  const ScaleJob& J = jobs.j[blockIdx.y];
  int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x & 31;
  if (row >= J.out) return;
  float s = 1.0f;
  if (J.g != nullptr) {
    double acc = 0.0;
    const float* r = J.v + (int64_t)row * J.in;
    for (int k = lane; k < J.in; k += 32) { double x = r[k]; acc += x * x; }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    s = J.g[row] / (float)sqrt(acc);
  }
  if (lane == 0) packed[J.scale_off + row] = s;
}

__global__ void pack_kernel(PackJobs jobs, float* packed) {
  const PackJob& J = jobs.j[blockIdx.y];
  int64_t total;
  if (J.mode == 0) total = (int64_t)(J.K0 + J.K1) * J.Np;
  else if (J.mode == 1) total = (int64_t)J.Nb * J.Kb;
  else if (J.mode == 2) total = J.Np;
  else total = (int64_t)J.n_valid * (J.K0 + J.K1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int n, k;  // destination output index / destination input row
    if (J.mode == 0) { k = (int)(i / J.Np); n = (int)(i % J.Np); }
    else if (J.mode == 1) { n = (int)(i / J.Kb); k = (int)(i % J.Kb); }
    else if (J.mode == 2) { n = (int)i; k = 0; }
    else { n = (int)(i / (J.K0 + J.K1)); k = (int)(i % (J.K0 + J.K1)); }
    float val = 0.0f;
    if (n < J.n_valid) {
      int srow = J.row_start + n;
      if (J.mode == 2) {
        val = J.bias[srow];
      } else {
        int scol = -1;
        if (J.mode == 1) { if (k < J.src_in) scol = k; }
        else if (k < J.K0) { if (k < J.k0v) scol = k; }
        else { if (k - J.K0 < J.k1v) scol = J.k0v + (k - J.K0); }
        if (scol >= 0) val = J.v[(int64_t)srow * J.src_in + scol] * packed[J.scale_off + srow];
      }
    }
    packed[J.dst_off + i] = val;
  }
}

static int64_t g_launches = 0;
void count_launch(int n) { g_launches += n; }
int64_t launches_total() { return g_launches; }

static bool check_linear(const CneusLinear& L, int out, int in, const char* what, int idx) {
  if (L.weight_v == nullptr || L.bias == nullptr || L.out != out || L.in != in) {
    set_error("%s[%d]: expected a [%d,%d] linear, got [%d,%d]%s", what, idx, out, in, L.out, L.in,
              (L.weight_v == nullptr || L.bias == nullptr) ? " (null pointer)" : "");
    return false;
  }
  return true;
}

}  // namespace cneus

using namespace cneus;

extern "C" int cneus_abi_version(void) { return CNEUS_ABI_VERSION; }
extern "C" const char* cneus_last_error(void) { return g_err; }
extern "C" int cneus_device_sm_count(void) { return sm_count(); }
namespace cneus { int64_t launches_total(); }
extern "C" int64_t cneus_launch_count(void) { return cneus::launches_total(); }

static const int64_t kScaleFloats = 8192;

extern "C" size_t cneus_packed_bytes(const CneusNetDesc* desc) {
  NetPack np;
  if (desc == nullptr || build_netpack(desc, &np) != CNEUS_OK) return 0;
  return (size_t)(np.total_floats + kScaleFloats) * sizeof(float);
}

extern "C" int cneus_pack_weights(const CneusNetDesc* desc, const CneusParams* P, void* packed_dev, size_t packed_bytes,
                                  void* stream) {
  CNEUS_NVTX_RANGE();
  if (!desc || !P || !packed_dev) { set_error("null argument"); return CNEUS_EINVAL; }
  NetPack np;
  int rc = build_netpack(desc, &np);
  if (rc != CNEUS_OK) return rc;
  if (packed_bytes < (size_t)(np.total_floats + kScaleFloats) * sizeof(float)) { set_error("packed buffer too small"); return CNEUS_ENOSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  float* packed = (float*)packed_dev;
  const CneusNetDesc& d = *desc;
  const int nl = d.sdf_n_lin, H = d.sdf_d_hidden, pe = np.pe_dim;

  // ---- validate shapes against the topology
  for (int l = 0; l < nl; ++l) {
    int in = l == 0 ? pe : H;
    int out = (l == nl - 1) ? d.sdf_d_out : ((l + 1 == d.sdf_skip) ? H - pe : H);
    if (!check_linear(P->sdf[l], out, in, "sdf", l)) return CNEUS_EINVAL;
  }
  if (d.color_n_lin > 0) {
    for (int l = 0; l < d.color_n_lin; ++l) {
      int in = l == 0 ? np.color_k0v + d.color_d_feature : d.color_d_hidden;
      int out = l == d.color_n_lin - 1 ? 3 : d.color_d_hidden;
      if (!check_linear(P->color[l], out, in, "color", l)) return CNEUS_EINVAL;
    }
  }
  if (d.has_relight) {
    if (!check_linear(P->relight_in, d.relight_d_hidden, np.relight_k0v, "relight_in", 0)) return CNEUS_EINVAL;
    for (int i = 0; i < d.relight_n_layers; ++i) {
      int in = d.relight_d_hidden + (i == d.relight_y_in_layer - 1 ? 3 : 0);
      int out = i == d.relight_n_layers - 1 ? 3 : d.relight_d_hidden;
      if (!check_linear(P->relight_mlp[i], out, in, "relight_mlp", i)) return CNEUS_EINVAL;
    }
  }

  // ---- scales
  std::vector<const CneusLinear*> lins;
  for (int l = 0; l < nl; ++l) lins.push_back(&P->sdf[l]);
  for (int l = 0; l < d.color_n_lin; ++l) lins.push_back(&P->color[l]);
  if (d.has_relight) {
    lins.push_back(&P->relight_in);
    for (int i = 0; i < d.relight_n_layers; ++i) lins.push_back(&P->relight_mlp[i]);
  }
  std::vector<int64_t> scale_off(lins.size());
  int64_t so = np.total_floats;
  ScaleJobs sj;
  memset(&sj, 0, sizeof(sj));
  int max_out = 0;
  for (size_t i = 0; i < lins.size(); ++i) {
    scale_off[i] = so;
    so += lins[i]->out;
    if (so > np.total_floats + kScaleFloats || i >= 32) { set_error("too many rows for the scale scratch"); return CNEUS_EUNSUPPORTED; }
    sj.j[i] = ScaleJob{lins[i]->weight_v, lins[i]->weight_g, scale_off[i], lins[i]->out, lins[i]->in};
    if (lins[i]->out > max_out) max_out = lins[i]->out;
  }
  sj.n = (int)lins.size();
  wn_scale_kernel<<<dim3((max_out + 7) / 8, sj.n), 256, 0, st>>>(sj, packed);
  count_launch();
  CNEUS_CUDA_CHECK(cudaGetLastError());

  // ---- pack jobs
  std::vector<PackJob> jobs;
  auto add_player = [&](const PLayer& L, const CneusLinear& S, int64_t soff, int row_start) {
    PackJob j;
    memset(&j, 0, sizeof(j));
    j.v = S.weight_v; j.g = S.weight_g; j.bias = S.bias; j.scale_off = soff; j.src_out = S.out; j.src_in = S.in;
    j.row_start = row_start; j.n_valid = L.N; j.Np = L.Np; j.k0v = L.k0v; j.K0 = L.K0; j.k1v = L.k1v; j.K1 = L.K1;
    j.Nb = L.Nb; j.Kb = L.Kb;
    j.mode = 0; j.dst_off = L.wt_off; jobs.push_back(j);
    j.mode = 2; j.dst_off = L.bias_off; jobs.push_back(j);
    if (L.wb_off >= 0) { j.mode = 1; j.dst_off = L.wb_off; jobs.push_back(j); }
  };
  auto add_row = [&](const RowLayer& R, const CneusLinear& S, int64_t soff, int row_start, int k0v) {
    PackJob j;
    memset(&j, 0, sizeof(j));
    j.v = S.weight_v; j.g = S.weight_g; j.bias = S.bias; j.scale_off = soff; j.src_out = S.out; j.src_in = S.in;
    j.row_start = row_start; j.n_valid = R.N; j.Np = R.N; j.k0v = k0v; j.K0 = R.K0; j.k1v = S.in - k0v; j.K1 = R.K1;
    j.mode = 3; j.dst_off = R.w_off; jobs.push_back(j);
    j.mode = 2; j.dst_off = R.bias_off; jobs.push_back(j);
  };
  size_t li = 0;
  for (int l = 0; l < nl; ++l, ++li) {
    if (l < nl - 1) add_player(np.sdf[l], P->sdf[l], scale_off[li], 0);
    else { add_player(np.sdf[l], P->sdf[l], scale_off[li], 1); add_row(np.sdf_row, P->sdf[l], scale_off[li], 0, 0); }
  }
  for (int l = 0; l < d.color_n_lin; ++l, ++li) {
    if (l < d.color_n_lin - 1) add_player(np.color[l], P->color[l], scale_off[li], 0);
    else add_row(np.color_row, P->color[l], scale_off[li], 0, 0);
  }
  if (d.has_relight) {
    add_player(np.rl_in, P->relight_in, scale_off[li], 0); ++li;
    for (int i = 0; i < d.relight_n_layers; ++i, ++li) {
      if (i < d.relight_n_layers - 1) add_player(np.rl[i], P->relight_mlp[i], scale_off[li], 0);
      else add_row(np.rl_row, P->relight_mlp[i], scale_off[li], 0, (d.relight_y_in_layer == d.relight_n_layers) ? 3 : 0);
    }
  }
  { int rc_tc = pack_tc_weights(np, P, scale_off.data(), packed, st); if (rc_tc != CNEUS_OK) return rc_tc; }
  for (size_t b = 0; b < jobs.size(); b += MAX_JOBS) {
    PackJobs pj;
    memset(&pj, 0, sizeof(pj));
    pj.n = (int)((jobs.size() - b) < (size_t)MAX_JOBS ? (jobs.size() - b) : MAX_JOBS);
    for (int i = 0; i < pj.n; ++i) pj.j[i] = jobs[b + i];
    pack_kernel<<<dim3(80, pj.n), 256, 0, st>>>(pj, packed);
    count_launch();
    CNEUS_CUDA_CHECK(cudaGetLastError());
  }
  return CNEUS_OK;
}
