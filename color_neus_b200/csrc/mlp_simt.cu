// Fused per-point network evaluation (fp32, CUDA cores): positional encoding -> SDF MLP -> analytic input
// gradient (reverse chain) -> colour MLP -> relight MLP for a tile of TM points per CTA, all activations in
// shared memory, weights streamed from the (L2-resident) packed buffer in KC-row chunks with cp.async.
//
// Restates: SDFNetwork.forward/gradient (fields.py:81-115), Embedder.embed (PositionEncoding.py:51-76),
// RenderingNetwork.forward (fields.py:161-188), RelightNetwork.relight (fields.py:332-359),
// inverse_sigmoid (lib/utils/transform.py:304-320).
#include "common.cuh"

namespace cneus {

// ---------------------------------------------------------------------------------------------------------
// shared-memory carve-up (floats)
// ---------------------------------------------------------------------------------------------------------
constexpr int BUF_FLOATS = MAXH * TM;        // one activation buffer: [256 features][64 points]
constexpr int WBUF_FLOATS = 2 * KC * MAXH;   // double-buffered weight chunk
constexpr int SMALL_FLOATS = SMALLK * TM;
constexpr int SM_BUFA = 0;
constexpr int SM_BUFB = SM_BUFA + BUF_FLOATS;
constexpr int SM_WBUF = SM_BUFB + BUF_FLOATS;
constexpr int SM_SMALL = SM_WBUF + WBUF_FLOATS;  // small input segment of colour / relight
constexpr int SM_X0 = SM_SMALL + SMALL_FLOATS;   // positional encoding of the scaled point (lin0 input, skip input)
constexpr int SM_GX0 = SM_X0 + SMALL_FLOATS;     // adjoint of the positional encoding
constexpr int SM_PTS = SM_GX0 + SMALL_FLOATS;    // [3][TM]
constexpr int SM_DIR = SM_PTS + 3 * TM;
constexpr int SM_NRM = SM_DIR + 3 * TM;
constexpr int SM_CG = SM_NRM + 3 * TM;           // colour-network output
constexpr int SM_DRGB = SM_CG + 4 * TM;
constexpr int SM_SDF = SM_DRGB + 4 * TM;
constexpr int SM_TOTAL = SM_SDF + TM;
constexpr size_t SHADE_SMEM_BYTES = (size_t)SM_TOTAL * sizeof(float);

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// acc[i][j] = sum_k X[k][m0+i] * Wt[k][n0+j];  X rows come from seg0 (K0 rows) then seg1 (K1 rows).
// Wt is dense [K0+K1][Np] in global memory, so chunk c is the contiguous block of KC*Np floats at c*KC*Np.
__device__ __forceinline__ void gemm_tile(float (&acc)[8][8], const float* __restrict__ Wt, int Np, const float* seg0,
                                          int K0, const float* seg1, int K1, float* wbuf, int n0, int m0) {
  const int tid = threadIdx.x;
  const int nchunks = (K0 + K1) / KC;
  const int chunk_f4 = KC * Np / 4;
  const bool active = n0 < Np;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  for (int i = tid; i < chunk_f4; i += NT) cp_async16(wbuf + 4 * i, Wt + 4 * i);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      float* dst = wbuf + ((c + 1) & 1) * (KC * MAXH);
      const float* src = Wt + (size_t)(c + 1) * KC * Np;
      for (int i = tid; i < chunk_f4; i += NT) cp_async16(dst + 4 * i, src + 4 * i);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int krow = c * KC;
    const float* X = (krow < K0) ? (seg0 + krow * TM) : (seg1 + (krow - K0) * TM);
    const float* W = wbuf + (c & 1) * (KC * MAXH);
    if (active) {
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        float4 a0 = *reinterpret_cast<const float4*>(X + kk * TM + m0);
        float4 a1 = *reinterpret_cast<const float4*>(X + kk * TM + m0 + 4);
        float4 b0 = *reinterpret_cast<const float4*>(W + kk * Np + n0);
        float4 b1 = *reinterpret_cast<const float4*>(W + kk * Np + n0 + 4);
        float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
}

enum { ACT_NONE = 0, ACT_SOFTPLUS = 1, ACT_RELU = 2 };

// Y[row_off + n][m] = act(acc + bias[n]) * out_scale for n < Nvalid; optionally D[n][m] = softplus'(a) to global.
__device__ __forceinline__ void epilogue_store(const float (&acc)[8][8], const float* __restrict__ bias, int n0, int m0,
                                               int Nvalid, int act, float out_scale, float* Y, int row_off,
                                               float* __restrict__ dsave) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = n0 + j;
    if (n >= Nvalid) continue;
    const float b = bias[n];
    float v[8], dv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = acc[i][j] + b;
      if (act == ACT_SOFTPLUS) {
        v[i] = softplus100(a) * out_scale;
        dv[i] = softplus100_grad(a);
      } else if (act == ACT_RELU) {
        v[i] = fmaxf(a, 0.0f) * out_scale;
      } else {
        v[i] = a * out_scale;
      }
    }
    float* y = Y + (row_off + n) * TM + m0;
    *reinterpret_cast<float4*>(y) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(y + 4) = make_float4(v[4], v[5], v[6], v[7]);
    if (act == ACT_SOFTPLUS && dsave != nullptr) {
      float* dg = dsave + n * TM + m0;
      *reinterpret_cast<float4*>(dg) = make_float4(dv[0], dv[1], dv[2], dv[3]);
      *reinterpret_cast<float4*>(dg + 4) = make_float4(dv[4], dv[5], dv[6], dv[7]);
    }
  }
}

// out[j][m] = bias[j] + sum_k X[k][m] * W[j][k]   (N <= 4 narrow layers)
__device__ __forceinline__ void rows_eval(const RowLayer& R, const float* __restrict__ packed, const float* seg0,
                                          const float* seg1, float* out /*[4][TM]*/) {
  const int j = threadIdx.x / TM, m = threadIdx.x % TM;
  if (j < R.N) {
    const float* w = packed + R.w_off + (size_t)j * (R.K0 + R.K1);
    float acc = 0.0f;
    for (int k = 0; k < R.K0; ++k) acc = fmaf(seg0[k * TM + m], __ldg(w + k), acc);
    for (int k = 0; k < R.K1; ++k) acc = fmaf(seg1[k * TM + m], __ldg(w + R.K0 + k), acc);
    out[j * TM + m] = acc + __ldg(packed + R.bias_off + j);
  }
}

// rows [row0, row0 + 3*(1+2L)) of dst <- [x | sin(2^k x) | cos(2^k x)]_k  (PositionEncoding.py:51-76), x = src[3][TM]*scale
__device__ __forceinline__ void fill_pe(float* dst, int row0, const float* src, int L, float scale) {
  const int per = 1 + 2 * L;
  for (int idx = threadIdx.x; idx < 3 * per * TM; idx += NT) {
    int m = idx % TM, r = idx / TM;  // r in [0, 3*per)
    int blk = r / 3, dim = r % 3;
    float x = src[dim * TM + m] * scale;
    float v;
    if (blk == 0) v = x;
    else {
      int k = (blk - 1) >> 1;
      float xf = x * (float)(1 << k);
      v = ((blk - 1) & 1) ? cosf(xf) : sinf(xf);
    }
    dst[(row0 + r) * TM + m] = v;
  }
}
__device__ __forceinline__ void copy_rows3(float* dst, int row0, const float* src, float sign) {
  for (int idx = threadIdx.x; idx < 3 * TM; idx += NT) dst[row0 * TM + idx] = sign * src[idx];
}
__device__ __forceinline__ void zero_rows(float* dst, int row_begin, int row_end) {
  for (int idx = threadIdx.x + row_begin * TM; idx < row_end * TM; idx += NT) dst[idx] = 0.0f;
}
// global [P,3] (explicit array) -> smem [3][TM]
__device__ __forceinline__ void load_p3(float* dst, const float* __restrict__ src, int64_t p0, int cnt) {
  for (int idx = threadIdx.x; idx < 3 * TM; idx += NT) {
    int m = idx / 3, c = idx % 3;
    dst[c * TM + m] = (m < cnt) ? src[(p0 + m) * 3 + c] : 0.0f;
  }
}
__device__ __forceinline__ void store_p3(float* __restrict__ dst, const float* src, int64_t p0, int cnt, float sign = 1.0f) {
  for (int idx = threadIdx.x; idx < 3 * cnt; idx += NT) {
    int m = idx / 3, c = idx % 3;
    dst[p0 * 3 + idx] = sign * src[c * TM + m];
  }
}

__global__ void __launch_bounds__(NT, 1) shade_kernel(const __grid_constant__ NetPack np, const float* __restrict__ packed,
                                                      const __grid_constant__ ShadeArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem + SM_BUFA;
  float* bufB = smem + SM_BUFB;
  float* wbuf = smem + SM_WBUF;
  float* small = smem + SM_SMALL;
  float* x0 = smem + SM_X0;
  float* gx0 = smem + SM_GX0;
  float* ptsS = smem + SM_PTS;
  float* dirS = smem + SM_DIR;
  float* nrmS = smem + SM_NRM;
  float* cgS = smem + SM_CG;
  float* drgbS = smem + SM_DRGB;
  float* sdfS = smem + SM_SDF;

  const CneusNetDesc& d = np.d;
  const int tid = threadIdx.x;
  const int n0 = (tid >> 3) * 8, m0 = (tid & 7) * 8;
  const int nl = d.sdf_n_lin;
  const int n_hidden = nl - 1;
  const float inv_sqrt2 = 0.70710678118654752440f;
  float* dscr = a.dscratch ? a.dscratch + (size_t)blockIdx.x * n_hidden * MAXH * TM : nullptr;
  const int64_t n_tiles = (a.P + TM - 1) / TM;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p0 = tile * TM;
    const int cnt = (int)((a.P - p0) < TM ? (a.P - p0) : TM);
    __syncthreads();
    // ------------------------------------------------------------------ points and directions
    if (a.src_mode == 0) {
      load_p3(ptsS, a.pts, p0, cnt);
    } else if (a.src_mode == 1) {
      for (int idx = tid; idx < 3 * TM; idx += NT) {
        int m = idx % TM, c = idx / TM;
        float v = 0.0f, dv = 0.0f;
        if (m < cnt) {
          int64_t p = p0 + m, r = p / a.n_per_ray;
          dv = a.rays_d[r * 3 + c];
          v = ray_point(a.rays_o[r * 3 + c], dv, a.t[p]);
        }
        ptsS[c * TM + m] = v;
        dirS[c * TM + m] = dv;
      }
    } else {
      for (int idx = tid; idx < 3 * TM; idx += NT) {
        int m = idx % TM, c = idx / TM;
        float v = 0.0f;
        if (m < cnt) {
          int64_t lin = a.lin_begin + p0 + m;
          int iz = (int)(lin % a.res), iy = (int)((lin / a.res) % a.res), ix = (int)(lin / ((int64_t)a.res * a.res));
          v = c == 0 ? a.gx[ix] : (c == 1 ? a.gy[iy] : a.gz[iz]);
        }
        ptsS[c * TM + m] = v;
      }
    }
    if (a.in_viewdirs != nullptr) load_p3(dirS, a.in_viewdirs, p0, cnt);
    else if (a.src_mode != 1) zero_rows(dirS, 0, 3);
    if (a.in_normals != nullptr) load_p3(nrmS, a.in_normals, p0, cnt);
    __syncthreads();

    float acc[8][8];
    float* featbuf = nullptr;  // where the feature vector [d_feature][TM] lives after the SDF stage

    // ------------------------------------------------------------------ SDF forward (fields.py:81-97)
    if (a.run_sdf) {
      if (d.sdf_multires > 0) fill_pe(x0, 0, ptsS, d.sdf_multires, d.sdf_scale);
      else for (int idx = tid; idx < 3 * TM; idx += NT) x0[idx] = ptsS[idx] * d.sdf_scale;
      zero_rows(x0, np.pe_dim, pad_to(np.pe_dim, KC));
      __syncthreads();
      float* cur = x0;     // input of layer l
      float* nxt = bufA;
      for (int l = 0; l < n_hidden; ++l) {
        const PLayer& L = np.sdf[l];
        gemm_tile(acc, packed + L.wt_off, L.Np, nullptr, 0, cur, L.K1, wbuf, n0, m0);
        const bool feeds_skip = (l + 1 == d.sdf_skip);
        epilogue_store(acc, packed + L.bias_off, n0, m0, L.N, ACT_SOFTPLUS, feeds_skip ? inv_sqrt2 : 1.0f, nxt, 0,
                       (a.run_grad && dscr) ? dscr + (size_t)l * MAXH * TM : nullptr);
        if (feeds_skip) {  // x = cat([x, inputs]) / sqrt(2)  (fields.py:90-91)
          for (int idx = tid; idx < np.pe_dim * TM; idx += NT) nxt[L.N * TM + idx] = x0[idx] * inv_sqrt2;
        }
        __syncthreads();
        cur = nxt;
        nxt = (cur == bufA) ? bufB : bufA;
      }
      // last layer: column 0 (sdf) as a dot product, columns 1.. (feature) as a GEMM
      rows_eval(np.sdf_row, packed, nullptr, cur, sdfS);
      __syncthreads();
      if (tid < TM) sdfS[tid] = sdfS[tid] / d.sdf_scale;
      if (a.run_sdf == 2) {
        const PLayer& L = np.sdf[nl - 1];
        gemm_tile(acc, packed + L.wt_off, L.Np, nullptr, 0, cur, L.K1, wbuf, n0, m0);
        epilogue_store(acc, packed + L.bias_off, n0, m0, L.N, ACT_NONE, 1.0f, nxt, 0, nullptr);
        featbuf = nxt;
      }
      __syncthreads();
      if (a.out_sdf != nullptr && tid < cnt) a.out_sdf[p0 + tid] = a.out_sdf_sign * sdfS[tid];
      if (a.out_full != nullptr) {
        const int dout = d.sdf_d_out;
        for (int idx = tid; idx < cnt * dout; idx += NT) {
          int m = idx / dout, c = idx % dout;
          a.out_full[p0 * dout + idx] = (c == 0) ? sdfS[m] : featbuf[(c - 1) * TM + m];
        }
      }

      // ---------------------------------------------------------------- d sdf / d x by the reverse chain
      if (a.run_grad) {
        // The chain runs in place in `cur` (the last hidden activation, dead by now); the other buffer may hold
        // the feature vector for the colour stage.  ga[n][m] = adjoint of the pre-activation n of hidden layer l.
        float* ga = cur;
        zero_rows(gx0, 0, SMALLK);
        __syncthreads();
        {  // seed: d sdf / d (input of the last linear) = W_last[0,:] / scale
          const float* w = packed + np.sdf_row.w_off;
          const float* D = dscr + (size_t)(n_hidden - 1) * MAXH * TM;
          const PLayer& Lh = np.sdf[n_hidden - 1];
          const bool last_is_skip = (d.sdf_skip == nl - 1);
          const float sc = last_is_skip ? inv_sqrt2 : 1.0f;
          const int rows = last_is_skip ? Lh.N + np.pe_dim : Lh.N;
          const int rows_all = rows > Lh.Nb ? rows : Lh.Nb;
          for (int idx = tid; idx < rows_all * TM; idx += NT) {
            const int n = idx / TM;
            const float g = (n < rows) ? (__ldg(w + n) / d.sdf_scale) * sc : 0.0f;
            if (n < Lh.N) ga[idx] = g * D[idx];
            else {
              if (n < Lh.Nb) ga[idx] = 0.0f;
              if (last_is_skip && n < rows) gx0[idx - Lh.N * TM] = g;
            }
          }
        }
        __syncthreads();
        for (int l = n_hidden - 1; l >= 0; --l) {
          const PLayer& L = np.sdf[l];
          // g_in[k][m] = sum_n Wb[n][k] ga[n][m]   (rows n = outputs of layer l, cols k = its inputs).
          // gemm_tile ends with __syncthreads(), so every read of `ga` is complete before it is overwritten.
          gemm_tile(acc, packed + L.wb_off, L.Kb, nullptr, 0, ga, L.Nb, wbuf, n0, m0);
          if (l > 0) {
            const PLayer& Lp = np.sdf[l - 1];
            const float* D = dscr + (size_t)(l - 1) * MAXH * TM;
            const bool is_skip = (l == d.sdf_skip);
            const float sc = is_skip ? inv_sqrt2 : 1.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = n0 + j;
              if (k >= L.Kb) continue;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int m = m0 + i;
                const float g = acc[i][j] * sc;
                if (k < Lp.N) ga[k * TM + m] = g * D[k * TM + m];
                else {
                  if (k < Lp.Nb) ga[k * TM + m] = 0.0f;
                  if (is_skip && k < Lp.N + np.pe_dim) gx0[(k - Lp.N) * TM + m] = g;
                }
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = n0 + j;
              if (k < np.pe_dim) {
#pragma unroll
                for (int i = 0; i < 8; ++i) gx0[k * TM + m0 + i] += acc[i][j];
              }
            }
          }
          __syncthreads();
        }
        // chain rule through the encoding and the input scaling
        if (tid < 3 * TM) {
          const int m = tid % TM, dim = tid / TM;
          const float xs = ptsS[dim * TM + m] * d.sdf_scale;
          float g = gx0[dim * TM + m];
          for (int k = 0; k < d.sdf_multires; ++k) {
            const float f = (float)(1 << k);
            const float gs = gx0[(3 + 6 * k + dim) * TM + m], gc = gx0[(6 + 6 * k + dim) * TM + m];
            g += f * (cosf(xs * f) * gs - sinf(xs * f) * gc);
          }
          nrmS[dim * TM + m] = g * d.sdf_scale;
        }
        __syncthreads();
        if (a.out_grad != nullptr) store_p3(a.out_grad, nrmS, p0, cnt);
      }
    }

    // ------------------------------------------------------------------ colour network (fields.py:161-188)
    if (a.run_color) {
      if (featbuf == nullptr) {  // stand-alone call: features come from the caller
        featbuf = bufA;
        const int F = d.color_d_feature;
        for (int idx = tid; idx < TM * F; idx += NT) {
          int m = idx / F, c = idx % F;
          featbuf[c * TM + m] = (m < cnt) ? a.in_feats[(p0 + m) * F + c] : 0.0f;
        }
      }
      // small segment: [pts | PE(view) | normal] according to the mode
      copy_rows3(small, 0, ptsS, 1.0f);
      int row = 3;
      const float vsign = (a.viewdir_mode == 1) ? -1.0f : 1.0f;
      const float* vsrc = (a.viewdir_mode == 1) ? nrmS : dirS;
      if (d.color_mode != CNEUS_COLOR_NO_VIEW_DIR) {
        if (d.color_multires_view > 0) {
          if (vsign < 0.0f) {  // PE of -n: negate into drgbS scratch first
            __syncthreads();
            for (int idx = tid; idx < 3 * TM; idx += NT) drgbS[idx] = -vsrc[idx];
            __syncthreads();
            fill_pe(small, row, drgbS, d.color_multires_view, 1.0f);
          } else {
            fill_pe(small, row, vsrc, d.color_multires_view, 1.0f);
          }
          row += 3 * (1 + 2 * d.color_multires_view);
        } else {
          copy_rows3(small, row, vsrc, vsign);
          row += 3;
        }
      }
      if (d.color_mode != CNEUS_COLOR_NO_NORMAL) { copy_rows3(small, row, nrmS, 1.0f); row += 3; }
      zero_rows(small, row, np.color[0].K0);
      __syncthreads();
      float* cur = featbuf;
      float* nxt = (featbuf == bufA) ? bufB : bufA;
      for (int l = 0; l < d.color_n_lin - 1; ++l) {
        const PLayer& L = np.color[l];
        gemm_tile(acc, packed + L.wt_off, L.Np, small, L.K0, cur, L.K1, wbuf, n0, m0);
        epilogue_store(acc, packed + L.bias_off, n0, m0, L.N, ACT_RELU, 1.0f, nxt, 0, nullptr);
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
      }
      rows_eval(np.color_row, packed, nullptr, cur, cgS);
      __syncthreads();
      if (d.color_squeeze_out && tid < 3 * TM) cgS[tid] = sigmoidf_(cgS[tid]);
      __syncthreads();
      if (a.out_color != nullptr) store_p3(a.out_color, cgS, p0, cnt);
    } else if (a.run_relight && a.in_rgb != nullptr) {
      load_p3(cgS, a.in_rgb, p0, cnt);
      __syncthreads();
    }

    // ------------------------------------------------------------------ relight network (fields.py:332-359)
    if (a.run_relight) {
      copy_rows3(small, 0, ptsS, 1.0f);
      int row = 3;
      if (d.relight_multires_view > 0) { fill_pe(small, row, dirS, d.relight_multires_view, 1.0f); row += 3 * (1 + 2 * d.relight_multires_view); }
      else { copy_rows3(small, row, dirS, 1.0f); row += 3; }
      if (d.relight_include_grad) { copy_rows3(small, row, nrmS, 1.0f); row += 3; }
      zero_rows(small, row, np.rl_in.K0);
      __syncthreads();
      float* cur = bufA;
      float* nxt = bufB;
      {
        const PLayer& L = np.rl_in;
        gemm_tile(acc, packed + L.wt_off, L.Np, small, L.K0, nullptr, 0, wbuf, n0, m0);
        epilogue_store(acc, packed + L.bias_off, n0, m0, L.N, ACT_RELU, 1.0f, cur, 0, nullptr);
      }
      __syncthreads();
      // from here on the small segment carries the colour that is re-injected at Y_IN_LAYER
      copy_rows3(small, 0, cgS, 1.0f);
      zero_rows(small, 3, KC);
      __syncthreads();
      for (int i = 0; i < d.relight_n_layers - 1; ++i) {
        const PLayer& L = np.rl[i];
        gemm_tile(acc, packed + L.wt_off, L.Np, small, L.K0, cur, L.K1, wbuf, n0, m0);
        epilogue_store(acc, packed + L.bias_off, n0, m0, L.N, ACT_RELU, 1.0f, nxt, 0, nullptr);
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
      }
      rows_eval(np.rl_row, packed, small, cur, drgbS);
      __syncthreads();
      if (a.out_drgb != nullptr) store_p3(a.out_drgb, drgbS, p0, cnt);
      if (a.out_relit != nullptr) {
        for (int idx = tid; idx < 3 * cnt; idx += NT) {
          int m = idx / 3, c = idx % 3;
          float rgb = cgS[c * TM + m], dr = drgbS[c * TM + m], v;
          if (d.relight_inv_sigmoid) {  // sigmoid(inverse_sigmoid(rgb) + drgb), eps 1e-5 (transform.py:317-320)
            float x = fminf(fmaxf(rgb, 0.0f), 1.0f);
            float x1 = fmaxf(x, 1e-5f), x2 = fmaxf(1.0f - x, 1e-5f);
            v = sigmoidf_(logf(x1 / x2) + dr);
          } else {
            v = fminf(fmaxf(rgb + sigmoidf_(dr) - 0.5f, 0.0f), 1.0f);
          }
          a.out_relit[p0 * 3 + idx] = v;
        }
      }
    }
  }
}

size_t shade_scratch_floats_per_cta(const NetPack& np) { return (size_t)(np.d.sdf_n_lin - 1) * MAXH * TM; }

int shade_grid_for(int64_t P) {
  int64_t tiles = (P + TM - 1) / TM;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  return (int)(tiles < sms ? (tiles > 0 ? tiles : 1) : sms);
}

// ---- CUDA-event profiling of the shading launches (bench.py's roofline leg) -------------------------------
namespace {
constexpr int PROF_RING = 2048;
struct ProfKind { cudaEvent_t beg[PROF_RING], end[PROF_RING]; int n = 0; bool made = false; double carry_ms = 0; int64_t carry_n = 0; };
ProfKind g_prof[2];
bool g_prof_on = false;
void prof_drain(ProfKind& k) {
  for (int i = 0; i < k.n; ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(k.end[i]) == cudaSuccess && cudaEventElapsedTime(&ms, k.beg[i], k.end[i]) == cudaSuccess) { k.carry_ms += ms; k.carry_n += 1; }
  }
  k.n = 0;
}
}  // namespace

void profile_enable(int on) { g_prof_on = on != 0; }
int profile_read(int kind, double* total_ms, int64_t* launches) {
  if (kind < 0 || kind > 1) return CNEUS_EINVAL;
  ProfKind& k = g_prof[kind];
  prof_drain(k);
  if (total_ms) *total_ms = k.carry_ms;
  if (launches) *launches = k.carry_n;
  k.carry_ms = 0; k.carry_n = 0;
  return CNEUS_OK;
}

int launch_shade(const NetPack& np, const float* packed, const ShadeArgs& a, int grid, cudaStream_t st) {
  static bool attr_set[CNEUS_MAX_DEVICES] = {false};
  if (first_use_on_device(attr_set)) {
    CNEUS_CUDA_CHECK(cudaFuncSetAttribute(shade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SHADE_SMEM_BYTES));
  }
  if (a.P <= 0) return CNEUS_OK;
  if (a.run_grad && a.dscratch == nullptr) { set_error("gradient stage needs the activation-derivative scratch"); return CNEUS_EINVAL; }
  ProfKind* pk = nullptr;
  if (g_prof_on) {
    pk = &g_prof[(a.run_sdf == 2 || a.run_color || a.run_relight) ? 1 : 0];
    if (!pk->made) { for (int i = 0; i < PROF_RING; ++i) { cudaEventCreate(&pk->beg[i]); cudaEventCreate(&pk->end[i]); } pk->made = true; }
    if (pk->n == PROF_RING) prof_drain(*pk);
    cudaEventRecord(pk->beg[pk->n], st);
  }
  int rc = CNEUS_OK;
  if (tc_supports(np, a)) {
    // tensor-core path: 128-point tiles; the encoding-adjoint scratch follows the softplus' slots of all CTAs
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    float* gx = a.dscratch ? a.dscratch + (size_t)sms * np.d.sdf_n_lin * 256 * 128 : nullptr;
    rc = launch_shade_tc(np, packed, a, gx, st);
  } else {
    shade_kernel<<<grid, NT, SHADE_SMEM_BYTES, st>>>(np, packed, a);
  }
  if (pk) { cudaEventRecord(pk->end[pk->n], st); pk->n++; }
  if (rc != CNEUS_OK) return rc;
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return CNEUS_OK;
}

}  // namespace cneus
