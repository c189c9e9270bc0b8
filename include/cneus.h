/*
 * cneus.h -- C ABI of libcneus.so: the B200-native (sm_100a) implementation of the Color-NeuS
 * volume-rendering hot path.  Each entry point replaces one reference interface (cited as
 * file:line relative to the reference tree).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every pointer marked "dev" is a device pointer to contiguous row-major fp32 unless said otherwise;
 *   - the library never allocates device memory: the caller passes the packed-weight buffer and a
 *     workspace (sizes from cneus_packed_bytes / cneus_workspace_bytes);
 *   - all work is stream-ordered on `stream` (a cudaStream_t passed as void*); no host sync inside;
 *   - return value 0 = ok, negative = error (text via cneus_last_error(), thread-local); never throws.
 */
#ifndef CNEUS_H_
#define CNEUS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNEUS_ABI_VERSION 1
#define CNEUS_MAX_SDF_LIN 12
#define CNEUS_MAX_COLOR_LIN 8
#define CNEUS_MAX_RELIGHT_LIN 8

enum { CNEUS_OK = 0, CNEUS_EINVAL = -1, CNEUS_ECUDA = -2, CNEUS_ENOSPACE = -3, CNEUS_EUNSUPPORTED = -4 };

enum { CNEUS_COLOR_IDR = 0, CNEUS_COLOR_NO_VIEW_DIR = 1, CNEUS_COLOR_NO_NORMAL = 2 };

/* Network topology: the cfg keys read by SDFNetwork.__init__ (lib/models/renderers/fields.py:19-29),
 * RenderingNetwork.__init__ (fields.py:126-134) and RelightNetwork.__init__ (fields.py:296-303). */
typedef struct CneusNetDesc {
  int32_t sdf_n_lin;             /* number of linear layers = N_LAYERS + 1 (9) */
  int32_t sdf_d_hidden;          /* D_HIDDEN (<= 256, multiple of 64) */
  int32_t sdf_d_out;             /* D_OUT (257): column 0 = sdf, rest = feature */
  int32_t sdf_multires;          /* MULTIRES (6; 0 = raw xyz) */
  int32_t sdf_skip;              /* linear index whose input is cat([h, pe])/sqrt(2); -1 if none */
  float sdf_scale;               /* SCALE */
  int32_t color_mode;            /* CNEUS_COLOR_* */
  int32_t color_n_lin;           /* N_LAYERS + 1 (5) */
  int32_t color_d_hidden;
  int32_t color_d_feature;       /* must equal sdf_d_out - 1 */
  int32_t color_multires_view;   /* 0 or L */
  int32_t color_squeeze_out;     /* final sigmoid */
  int32_t has_relight;
  int32_t relight_n_layers;      /* entries of rl_mlp (4) */
  int32_t relight_y_in_layer;    /* Y_IN_LAYER (3) */
  int32_t relight_d_hidden;
  int32_t relight_multires_view;
  int32_t relight_include_grad;
  int32_t relight_inv_sigmoid;
  int32_t reserved[5];
} CneusNetDesc;

/* One nn.Linear as PyTorch stores it.  weight_g == NULL means a plain (not weight-normed) layer and
 * weight_v is the weight itself; otherwise W[i,:] = g[i] * v[i,:] / ||v[i,:]|| (legacy weight_norm, dim=0). */
typedef struct CneusLinear {
  const float* weight_g; /* dev [out] or NULL */
  const float* weight_v; /* dev [out, in] */
  const float* bias;     /* dev [out] */
  int32_t out, in;
} CneusLinear;

typedef struct CneusParams {
  CneusLinear sdf[CNEUS_MAX_SDF_LIN];           /* sdf_network.lin{l} */
  CneusLinear color[CNEUS_MAX_COLOR_LIN];       /* color_network.lin{l} */
  CneusLinear relight_in;                       /* relight_network.in_layer */
  CneusLinear relight_mlp[CNEUS_MAX_RELIGHT_LIN]; /* relight_network.rl_mlp.{i} */
} CneusParams;

/* Per-call outputs of render_core / forward.  Any pointer may be NULL (= not wanted) except where noted. */
typedef struct CneusRenderOut {
  float* color_fine;     /* dev [B,3]   sum_k w_k c_k                        (NeuS.py:273, Color_NeuS.py:113) */
  float* global_color;   /* dev [B,3]   sum_k w_k cg_k  (Color_NeuS only)    (Color_NeuS.py:116) */
  float* weight_sum;     /* dev [B]                                           (NeuS.py:384) */
  float* weight_max;     /* dev [B]                                           (NeuS.py:393) */
  float* depth;          /* dev [B]     sum_k w_k z_k  (section starts)       (NeuS.py:398) */
  float* weights;        /* dev [B,S]                                         (NeuS.py:269) */
  float* cdf;            /* dev [B,S]   prev_cdf                              (NeuS.py:288) */
  float* inside_sphere;  /* dev [B,S]   0/1                                   (NeuS.py:261) */
  float* gradients;      /* dev [B,S,3] d sdf / d x   (REQUIRED)              (NeuS.py:231) */
  float* delta_relight;  /* dev [B,S,3] (Color_NeuS only)                     (Color_NeuS.py:58) */
  float* sdf;            /* dev [B,S]   (REQUIRED)                            (NeuS.py:228) */
  float* sampled_color;  /* dev [B,S,3] colour that is composited (REQUIRED) */
  float* global_sampled; /* dev [B,S,3] colour-network output before relight (REQUIRED for Color_NeuS) */
  float* alpha;          /* dev [B,S] */
  float* mid_z;          /* dev [B,S]   (REQUIRED) section mid-points         (NeuS.py:218) */
  float* dists;          /* dev [B,S]   (REQUIRED) section lengths            (NeuS.py:216-217) */
  float* scalars;        /* dev [4]     {gradient_error, eikonal numerator, eikonal denominator, 1/inv_s} (REQUIRED) */
} CneusRenderOut;

int cneus_abi_version(void);
const char* cneus_last_error(void);
/* Number of SMs of the current device (grid sizing is a multiple of it); <0 on error. */
int cneus_device_sm_count(void);

/* Packed weights: effective weight-norm matrices in the kernels' layouts.  Re-pack whenever parameters change. */
size_t cneus_packed_bytes(const CneusNetDesc* desc);
int cneus_pack_weights(const CneusNetDesc* desc, const CneusParams* params, void* packed_dev, size_t packed_bytes,
                       void* stream);

/* Workspace (bytes) large enough for any entry point below on B rays x S total samples / P points. */
size_t cneus_workspace_bytes(const CneusNetDesc* desc, int64_t n_rays, int32_t n_total_samples, int64_t n_points);

/* SDFNetwork.forward / .sdf (fields.py:81-100): out_cols = 1 (sdf only) or sdf_d_out. */
int cneus_sdf_forward(const CneusNetDesc* desc, const void* packed, const float* pts, int64_t P, float* out,
                      int32_t out_cols, void* ws, size_t ws_bytes, void* stream);
/* SDFNetwork.gradient (fields.py:105-115): grad dev [P,3]. */
int cneus_sdf_gradient(const CneusNetDesc* desc, const void* packed, const float* pts, int64_t P, float* grad,
                       void* ws, size_t ws_bytes, void* stream);
/* RenderingNetwork.forward (fields.py:161-188): points, normals, view_dirs [P,3], feature_vectors [P,d_feature]. */
int cneus_color_forward(const CneusNetDesc* desc, const void* packed, const float* pts, const float* normals,
                        const float* view_dirs, const float* feats, int64_t P, float* rgb, void* ws, size_t ws_bytes,
                        void* stream);
/* RelightNetwork.forward (fields.py:361-368): returns relit rgb and drgb, both [P,3]. */
int cneus_relight_forward(const CneusNetDesc* desc, const void* packed, const float* rgb, const float* pts,
                          const float* dirs, const float* grads, int64_t P, float* rgb_out, float* drgb_out, void* ws,
                          size_t ws_bytes, void* stream);

/* get_embedder(multires, input_dims) / Embedder.embed (lib/models/tools/PositionEncoding.py:45-94), stand-alone:
 * x dev [P, input_dims] -> out dev [P, input_dims * (1 + 2 * multires)] = [x | sin(2^k x) | cos(2^k x)]_{k < multires}. */
int cneus_embed(const float* x, int64_t P, int32_t input_dims, int32_t multires, float* out, void* stream);

/* NeuS.up_sample (NeuS.py:136-181) + sample_pdf(det=True) (ray_utils.py:123-154).
 * z, sdf dev [B,n]; u dev [m] = linspace(.5/m, 1-.5/m, m); new_z dev [B,m]. */
int cneus_up_sample(const float* rays_o, const float* rays_d, const float* z, const float* sdf, int64_t B, int32_t n,
                    int32_t m, float inv_s, const float* u, float* new_z, void* stream);
/* NeuS.cat_z_vals (NeuS.py:183-197): merge + (unless last) SDF of the new points gathered into sorted order. */
int cneus_cat_z_vals(const CneusNetDesc* desc, const void* packed, const float* rays_o, const float* rays_d,
                     const float* z, const float* new_z, const float* sdf, int64_t B, int32_t n, int32_t m,
                     int32_t last, float* z_out, float* sdf_out, void* ws, size_t ws_bytes, void* stream);
/* The no_grad sampling block of NeuS.forward (NeuS.py:311-357).  lin dev [n_samples] = linspace(0,1,n_samples);
 * t_rand dev [B] raw U[0,1) draws or NULL (perturb == 0); u dev [n_importance/up_steps]; z_out dev [B, n_samples+n_importance]. */
int cneus_sample_z(const CneusNetDesc* desc, const void* packed, const float* rays_o, const float* rays_d,
                   const float* near, const float* far, const float* t_rand, const float* lin, const float* u,
                   int64_t B, int32_t n_samples, int32_t n_importance, int32_t up_steps, float* z_out, void* ws,
                   size_t ws_bytes, void* stream);

/* NeuS.render_core (NeuS.py:199-292) / Color_NeuS.render_core (Color_NeuS.py:24-138), background off.
 * z dev [B,S]; variance dev [1] (deviation_network.variance); sample_dist = 2/N_SAMPLES. */
int cneus_render_core(const CneusNetDesc* desc, const void* packed, const float* variance, const float* rays_o,
                      const float* rays_d, const float* z, int64_t B, int32_t S, float sample_dist,
                      float cos_anneal_ratio, const CneusRenderOut* out, void* ws, size_t ws_bytes, void* stream);

/* extract_fields (NeuS.py:14-28) with query -sdf (NeuS.py:416): u[lin] for lin in [lin_begin, lin_end),
 * lin = (ix*res + iy)*res + iz; xs/ys/zs dev [res] are the caller's linspace axes. */
int cneus_sdf_grid(const CneusNetDesc* desc, const void* packed, const float* xs, const float* ys, const float* zs,
                   int32_t res, int64_t lin_begin, int64_t lin_end, float* u, void* ws, size_t ws_bytes, void* stream);
/* extract_color (NeuS.py:44-64): global colour color_network(p, n, -n, feat) per vertex; rgb dev [V,3]. */
int cneus_vertex_color(const CneusNetDesc* desc, const void* packed, const float* vertices, int64_t V, float* rgb,
                       void* ws, size_t ws_bytes, void* stream);

/* ---- training backward: what loss.backward() computes in the reference through
 * render_core (autograd double-backward through SDFNetwork.gradient, fields.py:105-115; train.py:70) ------------- */
typedef struct CneusBackwardIn {
  /* saved from the forward call (same pointers as CneusRenderOut / the inputs of cneus_render_core) */
  const float* rays_o; const float* rays_d; const float* z; const float* mid_z; const float* dists;
  const float* sdf; const float* gradients; const float* sampled_color; const float* global_sampled;
  const float* alpha; const float* weights; const float* variance; const float* eikonal_den; /* scalars[2] */
  /* upstream gradients of the returned dict; NULL = zero */
  const float* g_color_fine;     /* [B,3] */
  const float* g_global_color;   /* [B,3] */
  const float* g_weight_sum;     /* [B] */
  const float* g_weight_max;     /* [B] */
  const float* g_depth;          /* [B] */
  const float* g_weights;        /* [B,S] */
  const float* g_cdf;            /* [B,S] */
  const float* g_gradients;      /* [B,S,3] */
  const float* g_delta_relight;  /* [B,S,3] */
  const float* g_gradient_error; /* [1] */
  const float* g_s_val_sum;      /* [1] sum over rays of the gradient of s_val */
} CneusBackwardIn;

typedef struct CneusLinearGrad { float* weight; /* dev [out,in], gradient of the EFFECTIVE weight */ float* bias; } CneusLinearGrad;
typedef struct CneusParamGrads {
  CneusLinearGrad sdf[CNEUS_MAX_SDF_LIN];
  CneusLinearGrad color[CNEUS_MAX_COLOR_LIN];
  CneusLinearGrad relight_in;
  CneusLinearGrad relight_mlp[CNEUS_MAX_RELIGHT_LIN];
  float* variance; /* dev [1] */
} CneusParamGrads;

size_t cneus_backward_workspace_bytes(const CneusNetDesc* desc, int64_t n_rays, int32_t n_total_samples);
/* `eff` carries the EFFECTIVE weights (weight_g = NULL, weight_v = W [out,in] row-major, bias).  Every gradient buffer
 * must be zero-initialised by the caller: contributions are accumulated.  d_rays_o / d_rays_d ([B,3], zero-initialised)
 * may be NULL.  z_vals are treated as constants (they are built under no_grad when N_IMPORTANCE > 0, NeuS.py:343-355). */
int cneus_render_backward(const CneusNetDesc* desc, const CneusParams* eff, const CneusBackwardIn* in, int64_t B, int32_t S,
                          float cos_anneal_ratio, const CneusParamGrads* grads, float* d_rays_o, float* d_rays_d, void* ws,
                          size_t ws_bytes, void* stream);

/* Rays per pass of cneus_render_backward (default 0 = one pass).  A memory knob: the workspace is sized for one pass
 * (query cneus_backward_workspace_bytes after setting it); results are identical up to the order of the fp32 accumulation
 * of the parameter gradients across passes; more passes are slower (per-launch fixed costs). */
void cneus_backward_chunk_rays(int rays);
/* Which parts of cneus_render_backward run as launches of the fused tensor-core kernel where the topology allows: bit 0 =
 * the recompute of the SDF forward pass and of the reverse chain (training dumps), bit 1 = the tangent pass of the double
 * backward (needs bit 0; measured no faster than the layer-wise tangent pass, so off by default).  Default 1; 0 = everything
 * layer by layer (GEMMs + element-wise kernels). */
void cneus_backward_fused_recompute(int mode);

/* ---- measurement hooks (bench.py): CUDA-event timing of the point-shading kernel on its own stream ----------
 * kind 0 = SDF-only launches (sampling), 1 = full launches (render_core / vertex colour).  When enabled, every
 * launch is bracketed by cudaEventRecord on the launch stream; cneus_profile_read synchronises those events
 * and returns (and clears) the accumulated device time and launch count. */
void cneus_profile_enable(int on);
int cneus_profile_read(int kind, double* total_ms, int64_t* launches);
/* Validation switch: 1 = evaluate every layer with the fp32 CUDA-core kernel instead of the tcgen05 kernel
 * (same entry points, same outputs; used by the tests to A/B the split-precision tensor-core path). */
void cneus_force_simt(int on);
/* Caller-side ray generation for selected pixels: replaces get_rays_multicam's "build all N*H*W
 * rays, then gather" (lib/models/tools/ray_utils.py:16-87) and get_rays_at (:90-119).  index[n] holds flat pixel indices
 * (cam * H + y) * W + x in the reference's reshape(-1, 3) order (the caller draws them with the reference's CPU RNG
 * sequence); index == NULL generates the n consecutive indices starting at `first` (full images, y*W+x order).
 * c2w [n_cam,4,4] row-major, focal [2] (device).  Optional (NULL to skip): origin [3] / radius [1] normalisation of the
 * origins (NeuS_Trainer.py:121-122), near / far (ray_utils.py:7-13), rgb / mask gathered from image [n_cam*H*W,3] /
 * mask [n_cam*H*W].  All pointers device, fp32 (index int64). */
int cneus_gen_rays(const float* c2w, int32_t n_cam, const float* focal, int32_t H, int32_t W, const int64_t* index,
                   int64_t first, int64_t n, int32_t normalize, int32_t opengl, const float* origin, const float* radius,
                   const float* image, const float* mask, float* rays_o, float* rays_d, float* near, float* far,
                   float* rgb, float* mask_out, void* stream);

/* pixels of the selected rays from uint8 images / masks resident on the device, converted with
 * the float pipeline of the reference datasets' get_image (lib/datasets/dtu.py:98-113: to_tensor, normalize(0.5, std),
 * * 0.5 + 0.5, premultiplied by mask / 255 for dtu / bmvs), bit for bit.  images dev u8 [n_img,H,W,3] RGB; masks dev u8
 * [n_img,H,W] or NULL; cam_map dev int64 [n_cam] (batch position -> dataset image, NULL = identity); index dev int64 [n]
 * flat indices (cam*H + y)*W + x into the batch (what cneus_gen_rays takes); rgb dev [n,3]; mask_out dev [n] or NULL. */
int cneus_gather_pixels_u8(const uint8_t* images, const uint8_t* masks, const int64_t* cam_map, const int64_t* index, int64_t n,
                           int32_t H, int32_t W, float std, int32_t premultiply_mask, float* rgb, float* mask_out, void* stream);

/* per-parameter gradient-norm clipping (net_utils.py:174-184: clip_grad_norm_(p, max_norm, 2) for
 * every parameter tensor) + torch.optim.Adam's update (net_utils.py:88) for all tensors in two launches.  `tensors` is a
 * HOST array of device-pointer descriptors; max_norm <= 0 disables clipping; `step` is the 1-based Adam step count;
 * write_clipped_grad != 0 also scales .grad in place like clip_grad_norm_ does; norms_out (device, [n_tensors], nullable)
 * receives the un-clipped L2 norms. */
typedef struct CneusAdamTensor {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
} CneusAdamTensor;
size_t cneus_clip_adam_workspace_bytes(int32_t n_tensors);
int cneus_clip_adam_step(const CneusAdamTensor* tensors, int32_t n_tensors, float max_norm, float lr, float beta1, float beta2,
                         float eps, float weight_decay, int64_t step, int32_t write_clipped_grad, float* norms_out, void* ws,
                         size_t ws_bytes, void* stream);

/* NeuS_Trainer.compute_loss (lib/models/NeuS_Trainer.py:129-171) and the seeds of its backward in
 * two launches.  terms[5] (device) = {loss, rgb_fine_loss, eikonal_loss, mask_loss, relight_loss}; g_* (device, same shapes
 * as the inputs) receive d loss / d input; d loss / d gradient_error is the constant lambda_eikonal.  mask (dev [B]) is
 * required when lambda_mask != 0 (BCE on clip(weight_sum, 1e-3, 1-1e-3), :141-143) or mask_relight != 0 (INCLUDE_MASK,
 * :147-151); delta_relight (dev [B,S,3]) may be NULL (plain NeuS); rgb_l1 selects L1Loss instead of MSELoss (:71-74). */
size_t cneus_loss_workspace_bytes(void);
int cneus_neus_loss(const float* color_fine, const float* rgb_gt, const float* weight_sum, const float* mask,
                    const float* gradient_error, const float* delta_relight, int64_t B, int32_t S, int32_t rgb_l1,
                    float lambda_fine, float lambda_eikonal, float lambda_mask, float lambda_relight, int32_t mask_relight,
                    float* terms, float* g_color_fine, float* g_weight_sum, float* g_delta_relight, void* ws, size_t ws_bytes,
                    void* stream);

/* marching cubes on the device over the fp32 grid u[nx][ny][nz] (x-major, as cneus_sdf_grid /
 * extract_fields NeuS.py:14-28 produce it); replaces mcubes.marching_cubes(u, threshold) (NeuS.py:35).  Two phases because
 * the mesh size is data dependent: cneus_mc_count classifies the grid into the workspace and writes {n_vertices,
 * n_triangles} to counts (dev int64[2]); the caller reads them, allocates, and cneus_mc_emit (same u / iso / workspace)
 * writes vertices (dev float64 [V,3], grid-index coordinates like PyMCubes) and triangles (dev int32 [F,3]).  Conventions
 * (corner inside <=> u < iso, vertex / triangle order, orientation towards u < iso) in csrc/marching_cubes.cu.
 * cneus_mc_tables copies the derived case table to HOST arrays n_tri[256], tri[256*16] (-1 padded); it needs no GPU. */
size_t cneus_mc_workspace_bytes(int32_t nx, int32_t ny, int32_t nz);
int cneus_mc_count(const float* u, int32_t nx, int32_t ny, int32_t nz, double iso, void* ws, size_t ws_bytes, int64_t* counts,
                   void* stream);
int cneus_mc_emit(const float* u, int32_t nx, int32_t ny, int32_t nz, double iso, void* ws, size_t ws_bytes, int64_t n_vertices,
                  int64_t n_triangles, double* vertices, int32_t* triangles, void* stream);
int cneus_mc_tables(uint8_t* n_tri, int8_t* tri);

/* Validation entry for the GEMMs of the training backward (color_neus_b200/csrc/gemm.cu, gemm_tc.cu), fp32 row-major device
 * operands: mode 0 (NT) C[M,N] = A[M,K] B[N,K]^T, 1 (NN) C = A[M,K] B[K,N], 2 (TN) C[M,N] = A[K,M]^T B[K,N]; optional bias[N],
 * ReLU, mask ((mask > 0) ? C : 0), accumulate.  use_tc = 1 dispatches like cneus_render_backward does (tcgen05 split-precision
 * kernels where the shape allows), 0 forces the fp32 CUDA-core SGEMM. */
size_t cneus_gemm_test_workspace_bytes(void);
int cneus_gemm_test(int mode, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                    int64_t ldc, const float* bias, int relu, const float* mask, int64_t ldmask, int accumulate, int use_tc,
                    void* ws, size_t ws_bytes, void* stream);
/* Role-level cycle counters of CTA 0 of the tensor-core kernel (development aid): out32 = {MMA wait-A, MMA wait-weights,
 * MMA total, steps, epilogue wait-accumulator, epilogue total, producer wait-slot, producer total, MMA wait for A slab 0..3,
 * wait-accumulator / total of epilogue warp 12, 2 spare, then (builds with -DCNEUS_TC_EPI_PROF only) the epilogue
 * timeline of thread 0 (16 slots, see EpiProf in mlp_tc_kernel.cu)}. */
void cneus_tc_prof_enable(int on);
int cneus_tc_prof_read(unsigned long long* out32, int reset);
/* profiling builds (-DCNEUS_TC_EPI_PROF): epilogue cycles / step counts by step type, out16[2 t], out16[2 t + 1] */
int cneus_tc_prof_read_types(unsigned long long* out16, int reset);
/* Number of kernels this library has launched since load (all kinds). */
int64_t cneus_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CNEUS_H_ */
