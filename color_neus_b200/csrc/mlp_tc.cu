// Tensor-core point-shading kernel for sm_100a: the same network evaluation as mlp_simt.cu (PE -> SDF MLP ->
// reverse-chain gradient -> colour MLP -> relight MLP) with every 256-wide layer on tcgen05.mma.
//
//   * one CTA per SM, tile = 128 points = the 128 TMEM lanes; thread r of the 4 epilogue warps owns point r;
//   * fp32 fidelity on fp16 tensor cores: every operand is split x = hi + lo (two fp16 planes) and each K-step
//     issues three MMAs (hi*hi + lo*hi + hi*lo) into the fp32 TMEM accumulator (measured rel. error ~1e-6,
//     tools/tc_probe.cu); weights are pre-scaled by 2^6 so their lo plane stays in the fp16 normal range;
//   * A operand (activations) lives in shared memory as K-major SWIZZLE_128B slabs written by the epilogue
//     threads; B operand (weights) is streamed from the L2-resident packed buffer by cp.async.bulk (TMA engine)
//     through a 2-stage mbarrier ring, already in its shared-memory image (pack_tc_kernel);
//   * warp roles: 0-3 epilogue (TMEM -> registers -> bias/activation -> fp16 split -> A slabs), 4 bulk-copy
//     producer, 5 single-thread MMA issuer.  The 257-wide last SDF layer is split: the feature block is an MMA
//     parked in the second TMEM accumulator, the sdf column / the 3-wide colour and relight outputs are fp32 dot
//     products folded into the preceding epilogue.
#include <cuda_fp16.h>
#include <string.h>

#include "common.cuh"

namespace cneus {

constexpr int TCM = 128;
constexpr int TC_THREADS = 192;
constexpr int SLAB_BYTES = 16384;          // [128 rows][64 halfs], SWIZZLE_128B
constexpr int A_SLABS = 5;                 // 4 main K-blocks + 1 small-input block
constexpr int TC_STAGES = 2;
constexpr int STAGE_BYTES = 2 * SLAB_BYTES;  // hi slab + lo slab of one (K-block, N-half)
constexpr float W_SCALE = 64.0f;
constexpr float BWD_ASCALE = 256.0f;       // scale of the A operand in the gradient chain
constexpr int MAX_TC_STEPS = 28;
constexpr int SMALL_SLAB = 4;

enum { EPI_HIDDEN = 0, EPI_PARK = 1, EPI_BWD = 2, EPI_BWD_LAST = 3 };
enum { TACT_SOFTPLUS = 1, TACT_RELU = 2 };
enum { PREP_NONE = 0, PREP_PE = 1, PREP_SEED = 2, PREP_COLOR_IN = 3, PREP_RELIGHT_IN = 4, PREP_CG = 5 };
enum { POST_NONE = 0, POST_SDF = 1, POST_CG = 2, POST_DRGB = 3 };
enum { TF_FEEDS_SKIP = 1, TF_SKIP_BWD = 2 };

struct TcStep {
  int64_t w_off;          // byte offset of the stage images [kb][nh][hi|lo] in the packed buffer
  int32_t bias_off;       // float offset of the fp32 bias (-1: none)
  int32_t row_off;        // float offset of a narrow layer [row_n][256] folded into this epilogue (-1: none)
  int32_t row_bias_off;
  int16_t row_n;
  int16_t n_valid;        // valid output columns
  int8_t n_kb, n_halves, acc, epi;
  int8_t act, prep_next, post, flags;
  int8_t d_layer;         // softplus' slot saved (forward) or loaded (gradient chain); -1 none
  int8_t slab[5];
  int8_t ksteps[5];
  int8_t pad_;
  float inv_scale;        // 1 / (weight scale * A-operand scale)
  float out_scale;        // factor applied to what is written to the next A operand
};

struct TcProgram {
  TcStep s[MAX_TC_STEPS];
  int32_t n_steps;
  int32_t n_hidden;           // hidden SDF layers = softplus' slots
  int32_t multires, pe_dim;
  float sdf_scale;
  int32_t seed_row_off;       // sdf row weights (gradient seed)
  int32_t feat_bias_off;
  float feat_inv_scale;
  int32_t color_mode, color_multires_view, color_squeeze;
  int32_t relight_multires_view, relight_include_grad, relight_inv_sigmoid;
  int32_t has_skip;
};

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  uint32_t spins = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();  // watchdog: a protocol bug must abort, never hang the GPU
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&f)[32]) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
}

// main accumulator + correction accumulator (256 columns further)
__device__ __forceinline__ void tmem_ld32_sum(uint32_t taddr, float (&f)[32]) {
  float c[32];
  tmem_ld32(taddr, f);
  tmem_ld32(taddr + 256u, c);
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] += c[i];
}

// ---------------------------------------------------------------------------------------------------------
// A-operand writers (row = the calling thread's point)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t a_chunk_offset(int slab, int row, int chunk) {
  return (uint32_t)slab * SLAB_BYTES + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u +
         (uint32_t)((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ uint32_t pack_h2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
// 8 consecutive K values -> one 16-byte chunk in each plane
__device__ __forceinline__ void write_a8(uint8_t* a_hi, uint8_t* a_lo, int slab, int row, int chunk, const float (&x)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    hi[i] = pack_h2(h);
    lo[i] = pack_h2(l);
  }
  const uint32_t off = a_chunk_offset(slab, row, chunk);
  *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
__device__ __forceinline__ void zero_a_row(uint8_t* a_hi, uint8_t* a_lo, int slab, int row) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t off = a_chunk_offset(slab, row, c);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(0, 0, 0, 0);
  }
}
// single element (small-input staging; a few dozen per tile)
__device__ __forceinline__ void put_a(uint8_t* a_hi, uint8_t* a_lo, int slab, int row, int k, float x) {
  const uint32_t off = a_chunk_offset(slab, row, k >> 3) + (uint32_t)(k & 7) * 2u;
  __half h = __float2half_rn(x);
  __half l = __float2half_rn(x - __half2float(h));
  *reinterpret_cast<__half*>(a_hi + off) = h;
  *reinterpret_cast<__half*>(a_lo + off) = l;
}
// [x | sin(2^k x) | cos(2^k x)]_k of a 3-vector, element q of 3*(1+2L)
__device__ __forceinline__ float pe_elem(const float (&x)[3], int q) {
  const int blk = q / 3, dim = q - 3 * blk;
  const float v = x[dim];
  if (blk == 0) return v;
  const int k = (blk - 1) >> 1;
  const float xf = v * (float)(1 << k);
  return ((blk - 1) & 1) ? cosf(xf) : sinf(xf);
}

__global__ void __launch_bounds__(TC_THREADS, 1) shade_tc_kernel(const __grid_constant__ TcProgram prog,
                                                                 const float* __restrict__ packed,
                                                                 const __grid_constant__ ShadeArgs a,
                                                                 float* __restrict__ gxscratch) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + A_SLABS * SLAB_BYTES;
  uint8_t* wring = smem + 2 * A_SLABS * SLAB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wring + TC_STAGES * STAGE_BYTES);
  uint64_t* bar_full = bars;                 // [TC_STAGES]
  uint64_t* bar_empty = bars + TC_STAGES;    // [TC_STAGES]
  uint64_t* bar_acc = bars + 2 * TC_STAGES;  // accumulator complete (MMA -> epilogue)
  uint64_t* bar_a = bars + 2 * TC_STAGES + 1;  // A operand ready (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_a, TCM);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t n_tiles = (a.P + TCM - 1) / TCM;
  const uint8_t* packed_b = reinterpret_cast<const uint8_t*>(packed);

  if (warp == 4) {
    // ================================================================ weight producer (bulk async copies)
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int s = 0; s < prog.n_steps; ++s) {
          const TcStep& S = prog.s[s];
          const int nst = S.n_kb * S.n_halves;
          const uint8_t* src = packed_b + S.w_off;
          for (int q = 0; q < nst; ++q, ++it) {
            const int stg = it % TC_STAGES;
            mbar_wait(&bar_empty[stg], ((it / TC_STAGES) & 1) ^ 1);
            mbar_expect_tx(&bar_full[stg], STAGE_BYTES);
            bulk_g2s(wring + stg * STAGE_BYTES, src + (size_t)q * STAGE_BYTES, STAGE_BYTES, &bar_full[stg]);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ================================================================ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TCM >> 4) << 24);  // f16 x f16 -> f32, N=128
      uint32_t it = 0, step_count = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int s = 0; s < prog.n_steps; ++s, ++step_count) {
          const TcStep& S = prog.s[s];
          mbar_wait(bar_a, step_count & 1);
          tc_fence_after();
          // Two fp32 accumulators: the hi*hi products go to columns [0,256), the 2^-11-sized correction products to
          // [256,512).  The tensor core truncates on every accumulate, so keeping the small terms away from the large
          // running sum cuts the systematic truncation bias 3x (16 instead of 48 accumulate steps on the big sum).
          const uint32_t d_base = tmem;
          for (int kb = 0; kb < S.n_kb; ++kb) {
            const uint32_t ah = smem_u32(a_hi + S.slab[kb] * SLAB_BYTES), al = smem_u32(a_lo + S.slab[kb] * SLAB_BYTES);
            for (int nh = 0; nh < S.n_halves; ++nh, ++it) {
              const int stg = it % TC_STAGES;
              mbar_wait(&bar_full[stg], (it / TC_STAGES) & 1);
              tc_fence_after();
              const uint32_t bh = smem_u32(wring + stg * STAGE_BYTES), bl = bh + SLAB_BYTES;
              const uint32_t d = d_base + (uint32_t)nh * 128u;
              for (int k = 0; k < S.ksteps[kb]; ++k) {
                const uint32_t koff = (uint32_t)k * 32u;
                const uint64_t dAh = make_desc_sw128(ah + koff), dAl = make_desc_sw128(al + koff);
                const uint64_t dBh = make_desc_sw128(bh + koff), dBl = make_desc_sw128(bl + koff);
                mma_f16(d, dAh, dBh, idesc, (kb | k) ? 1u : 0u);
                mma_f16(d + 256u, dAl, dBh, idesc, (kb | k) ? 1u : 0u);
                mma_f16(d + 256u, dAh, dBl, idesc, 1u);
              }
              mma_commit(&bar_empty[stg]);  // frees the ring slot when these MMAs retire
            }
          }
          mma_commit(bar_acc);
        }
      }
    }
  } else {
    // ================================================================ epilogue: thread = point
    const int row = threadIdx.x;  // 0..127 == TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    float* dscr = a.dscratch ? a.dscratch + (size_t)blockIdx.x * (prog.n_hidden + 1) * 256 * TCM : nullptr;  // +1: feature slot
    float* gxs = gxscratch + (size_t)blockIdx.x * 64 * TCM;
    const float inv_sqrt2 = 0.70710678118654752440f;
    uint32_t acc_count = 0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t p = tile * TCM + row;
      const bool valid = p < a.P;
      float pt[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f}, nrm[3] = {0.f, 0.f, 0.f}, cg[3] = {0.f, 0.f, 0.f};
      float sdfv = 0.f;
      if (valid) {
        if (a.src_mode == 0) {
          pt[0] = a.pts[p * 3]; pt[1] = a.pts[p * 3 + 1]; pt[2] = a.pts[p * 3 + 2];
        } else if (a.src_mode == 1) {
          const int64_t r = p / a.n_per_ray;
          const float t = a.t[p];
#pragma unroll
          for (int c = 0; c < 3; ++c) { dir[c] = a.rays_d[r * 3 + c]; pt[c] = ray_point(a.rays_o[r * 3 + c], dir[c], t); }
        } else {
          const int64_t lin = a.lin_begin + p;
          const int iz = (int)(lin % a.res), iy = (int)((lin / a.res) % a.res), ix = (int)(lin / ((int64_t)a.res * a.res));
          pt[0] = a.gx[ix]; pt[1] = a.gy[iy]; pt[2] = a.gz[iz];
        }
      }
      float xs[3] = {pt[0] * prog.sdf_scale, pt[1] * prog.sdf_scale, pt[2] * prog.sdf_scale};
      // ---- A operand of the first layer: positional encoding of the scaled point (PositionEncoding.py:51-76)
      zero_a_row(a_hi, a_lo, 0, row);
      for (int q = 0; q < prog.pe_dim; ++q) put_a(a_hi, a_lo, 0, row, q, prog.multires > 0 ? pe_elem(xs, q) : xs[q]);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_a);

      for (int s = 0; s < prog.n_steps; ++s, ++acc_count) {
        const TcStep& S = prog.s[s];
        mbar_wait(bar_acc, acc_count & 1);
        tc_fence_after();
        const uint32_t t_acc = t_lane;
        float dot[3] = {0.f, 0.f, 0.f};

        if (S.epi == EPI_HIDDEN) {
          const float* bias = packed + S.bias_off;
          float* dsave = (S.d_layer >= 0 && dscr) ? dscr + (size_t)S.d_layer * 256 * TCM : nullptr;
          const int nchunks = S.n_halves * 4;
          for (int c = 0; c < nchunks; ++c) {
            float v[32];
            tmem_ld32_sum(t_acc + c * 32, v);
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float o[8];
              const int nb = c * 32 + g8 * 8;
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + nb));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + nb + 4));
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int n = nb + j;
                float h = 0.f;
                if (n < S.n_valid) {
                  const float pre = fmaf(v[g8 * 8 + j], S.inv_scale, bb[j]);
                  if (S.act == TACT_SOFTPLUS) {
                    const float zz = 100.0f * pre;
                    float dd = 1.0f;
                    h = pre;
                    if (zz <= 20.0f) {
                      const float e = __expf(zz);
                      h = __logf(1.0f + e) * 0.01f;
                      dd = __fdividef(e, 1.0f + e);
                    }
                    if (dsave) dsave[n * TCM + row] = dd;
                  } else {
                    h = fmaxf(pre, 0.0f);
                  }
                  if (S.row_off >= 0) {
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj)
                      if (jj < S.row_n) dot[jj] = fmaf(h, __ldg(packed + S.row_off + jj * 256 + n), dot[jj]);
                  }
                  h *= S.out_scale;
                } else if ((S.flags & TF_FEEDS_SKIP) && n < S.n_valid + prog.pe_dim) {
                  // x = cat([x, inputs]) / sqrt(2)  (fields.py:90-91)
                  h = (prog.multires > 0 ? pe_elem(xs, n - S.n_valid) : xs[n - S.n_valid]) * inv_sqrt2;
                }
                o[j] = h;
              }
              write_a8(a_hi, a_lo, nb >> 6, row, (nb & 63) >> 3, o);
            }
          }
        } else if (S.epi == EPI_BWD) {
          // g_in = ga * W ; next adjoint = g_in (.) softplus'(a_{l-1})   (reverse chain of fields.py:105-115)
          const float* D = dscr + (size_t)S.d_layer * 256 * TCM;
          const float sc = (S.flags & TF_SKIP_BWD) ? S.inv_scale * inv_sqrt2 : S.inv_scale;
          for (int c = 0; c < 8; ++c) {
            float v[32];
            tmem_ld32_sum(t_acc + c * 32, v);
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float o[8];
              const int nb = c * 32 + g8 * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int k = nb + j;
                const float g = v[g8 * 8 + j] * sc;
                float w = 0.f;
                if (k < S.n_valid) w = g * D[k * TCM + row] * S.out_scale;
                else if ((S.flags & TF_SKIP_BWD) && k < S.n_valid + prog.pe_dim) gxs[(k - S.n_valid) * TCM + row] = g;
                o[j] = w;
              }
              write_a8(a_hi, a_lo, nb >> 6, row, (nb & 63) >> 3, o);
            }
          }
        } else if (S.epi == EPI_BWD_LAST) {
          // adjoint of the encoding -> d sdf / d x
          float g0[32], g1[32];
          tmem_ld32_sum(t_acc, g0);
          tmem_ld32_sum(t_acc + 32, g1);
          float gq[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int q = 0; q < 64; ++q) {
            if (q < prog.pe_dim) {
              float g = (q < 32 ? g0[q & 31] : g1[q & 31]) * S.inv_scale;
              if (prog.has_skip) g += gxs[q * TCM + row];
              const int blk = q / 3, dim = q - 3 * blk;
              float coef = 1.0f;
              if (blk > 0) {
                const int kf = (blk - 1) >> 1;
                const float f = (float)(1 << kf);
                const float xf = xs[dim] * f;
                coef = ((blk - 1) & 1) ? -f * sinf(xf) : f * cosf(xf);
              }
              gq[dim] = fmaf(coef, g, gq[dim]);
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) nrm[c] = gq[c] * prog.sdf_scale;
          if (valid && a.out_grad) { a.out_grad[p * 3] = nrm[0]; a.out_grad[p * 3 + 1] = nrm[1]; a.out_grad[p * 3 + 2] = nrm[2]; }
        } else {  // EPI_PARK: feature block of the last SDF layer -> fp32 scratch slot (read back by the colour stage)
          float* fslot = dscr ? dscr + (size_t)prog.n_hidden * 256 * TCM : nullptr;
          for (int c = 0; c < 8; ++c) {
            float v[32];
            tmem_ld32_sum(t_acc + c * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float f = fmaf(v[j], prog.feat_inv_scale, __ldg(packed + prog.feat_bias_off + c * 32 + j));
              if (fslot) fslot[(c * 32 + j) * TCM + row] = f;
              if (valid && a.out_full) a.out_full[p * 257 + 1 + c * 32 + j] = f;
            }
          }
          if (valid && a.out_full) a.out_full[p * 257] = sdfv;
        }

        // ---------------------------------------------------------------- narrow layers folded into this epilogue
        if (S.post == POST_SDF) {
          sdfv = (dot[0] + __ldg(packed + S.row_bias_off)) / prog.sdf_scale;
          if (valid && a.out_sdf) a.out_sdf[p] = a.out_sdf_sign * sdfv;
        } else if (S.post == POST_CG) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            cg[c] = dot[c] + __ldg(packed + S.row_bias_off + c);
            if (prog.color_squeeze) cg[c] = sigmoidf_(cg[c]);
          }
          if (valid && a.out_color) { a.out_color[p * 3] = cg[0]; a.out_color[p * 3 + 1] = cg[1]; a.out_color[p * 3 + 2] = cg[2]; }
        } else if (S.post == POST_DRGB) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float dr = dot[c] + __ldg(packed + S.row_bias_off + c);
            float rel;
            if (prog.relight_inv_sigmoid) {  // sigmoid(inverse_sigmoid(rgb) + drgb), eps 1e-5 (transform.py:317-320)
              const float x = fminf(fmaxf(cg[c], 0.0f), 1.0f);
              rel = sigmoidf_(logf(fmaxf(x, 1e-5f) / fmaxf(1.0f - x, 1e-5f)) + dr);
            } else {
              rel = fminf(fmaxf(cg[c] + sigmoidf_(dr) - 0.5f, 0.0f), 1.0f);
            }
            if (valid && a.out_drgb) a.out_drgb[p * 3 + c] = dr;
            if (valid && a.out_relit) a.out_relit[p * 3 + c] = rel;
          }
        }

        // ---------------------------------------------------------------- stage the A operand of the next step
        if (S.prep_next == PREP_SEED) {
          // d sdf / d a_last = W_last[0,:] / scale (.) softplus'(a_last)
          const float* D = dscr + (size_t)(prog.n_hidden - 1) * 256 * TCM;
          for (int nb = 0; nb < 256; nb += 8) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = (__ldg(packed + prog.seed_row_off + nb + j) / prog.sdf_scale) * D[(nb + j) * TCM + row] * BWD_ASCALE;
            write_a8(a_hi, a_lo, nb >> 6, row, (nb & 63) >> 3, o);
          }
        } else if (S.prep_next == PREP_COLOR_IN) {
          // colour input = [small: pts | PE(view) | normal] + [feature vector from the parked accumulator]
          const float* fslot = dscr + (size_t)prog.n_hidden * 256 * TCM;
          for (int nb = 0; nb < 256; nb += 8) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fslot[(nb + j) * TCM + row];
            write_a8(a_hi, a_lo, nb >> 6, row, (nb & 63) >> 3, o);
          }
          zero_a_row(a_hi, a_lo, SMALL_SLAB, row);
          int k = 0;
          for (int c = 0; c < 3; ++c) put_a(a_hi, a_lo, SMALL_SLAB, row, k++, pt[c]);
          if (prog.color_mode != CNEUS_COLOR_NO_VIEW_DIR) {
            float vd[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) vd[c] = (a.viewdir_mode == 1) ? -nrm[c] : dir[c];
            const int nv = prog.color_multires_view > 0 ? 3 * (1 + 2 * prog.color_multires_view) : 3;
            for (int q = 0; q < nv; ++q) put_a(a_hi, a_lo, SMALL_SLAB, row, k++, prog.color_multires_view > 0 ? pe_elem(vd, q) : vd[q]);
          }
          if (prog.color_mode != CNEUS_COLOR_NO_NORMAL)
            for (int c = 0; c < 3; ++c) put_a(a_hi, a_lo, SMALL_SLAB, row, k++, nrm[c]);
        } else if (S.prep_next == PREP_RELIGHT_IN) {
          zero_a_row(a_hi, a_lo, SMALL_SLAB, row);
          int k = 0;
          for (int c = 0; c < 3; ++c) put_a(a_hi, a_lo, SMALL_SLAB, row, k++, pt[c]);
          const int nv = prog.relight_multires_view > 0 ? 3 * (1 + 2 * prog.relight_multires_view) : 3;
          for (int q = 0; q < nv; ++q) put_a(a_hi, a_lo, SMALL_SLAB, row, k++, prog.relight_multires_view > 0 ? pe_elem(dir, q) : dir[q]);
          if (prog.relight_include_grad)
            for (int c = 0; c < 3; ++c) put_a(a_hi, a_lo, SMALL_SLAB, row, k++, nrm[c]);
        } else if (S.prep_next == PREP_CG) {
          zero_a_row(a_hi, a_lo, SMALL_SLAB, row);
          for (int c = 0; c < 3; ++c) put_a(a_hi, a_lo, SMALL_SLAB, row, c, cg[c]);
        }

        if (s + 1 < prog.n_steps) {
          fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
          tc_fence_before();
          mbar_arrive(bar_a);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

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
  const int64_t total = (int64_t)J.n_kb * J.n_halves * 128 * 64;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i & 63);
    const int r = (int)((i >> 6) & 127);
    const int stage = (int)(i >> 13);  // kb * n_halves + nh
    const int kb = stage / J.n_halves, nh = stage - kb * J.n_halves;
    const int nidx = nh * 128 + r;
    float val = 0.0f;
    if (nidx < J.n_valid && kk < J.kvalid[kb]) {
      int srow, scol;
      if (!J.transposed) { srow = J.row_start + nidx; scol = J.kstart[kb] + kk; }
      else { srow = J.kstart[kb] + kk; scol = nidx; }
      val = J.v[(int64_t)srow * J.src_in + scol] * packed[J.scale_off + srow] * W_SCALE;
    }
    const __half h = __float2half_rn(val);
    const __half l = __float2half_rn(val - __half2float(h));
    const int chunk = (kk >> 3) ^ (r & 7);
    const size_t off = (size_t)stage * STAGE_BYTES + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)chunk * 16 + (size_t)(kk & 7) * 2;
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
  if (a.in_normals || a.in_viewdirs || a.in_feats || a.in_rgb) return false;  // stand-alone sub-module calls
  if (!a.run_sdf) return false;
  if (a.run_color && (!a.run_grad || a.run_sdf != 2)) return false;  // one code path: colour always follows the gradient
  if (a.run_relight && !a.run_color) return false;
  if (a.out_full && a.run_sdf != 2) return false;
  return true;
}

size_t tc_scratch_floats_per_cta(const NetPack& np) { return (size_t)np.d.sdf_n_lin * 256 * TCM + 64 * TCM; }

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
  int n = 0;
  auto base = [&](int64_t w_off, int n_kb, int n_halves) -> TcStep& {
    TcStep& S = pg->s[n++];
    S.w_off = w_off; S.bias_off = -1; S.row_off = -1; S.row_bias_off = -1; S.row_n = 0; S.n_valid = 256;
    S.n_kb = (int8_t)n_kb; S.n_halves = (int8_t)n_halves; S.acc = 0; S.epi = EPI_HIDDEN; S.act = TACT_RELU;
    S.prep_next = PREP_NONE; S.post = POST_NONE; S.flags = 0; S.d_layer = -1;
    for (int i = 0; i < 5; ++i) { S.slab[i] = (int8_t)i; S.ksteps[i] = 4; }
    S.inv_scale = 1.0f / W_SCALE; S.out_scale = 1.0f;
    return S;
  };
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
      TcStep& S = base(np.tc_color[l], l == 0 ? 5 : 4, 2);
      S.bias_off = (int32_t)np.color[l].bias_off;
      if (l == 0) { S.slab[4] = SMALL_SLAB; S.ksteps[4] = (int8_t)ceil16(np.color_k0v); }
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
      S.bias_off = (int32_t)np.rl_in.bias_off; S.slab[0] = SMALL_SLAB; S.ksteps[0] = (int8_t)ceil16(np.relight_k0v);
      if (y - 1 == 0) S.prep_next = PREP_CG;
    }
    for (int i = 0; i < rn - 1; ++i) {
      const bool yin = (i == y - 1);
      TcStep& S = base(np.tc_rl[i], yin ? 5 : 4, 2);
      S.bias_off = (int32_t)np.rl[i].bias_off;
      if (yin) { S.slab[4] = SMALL_SLAB; S.ksteps[4] = 1; }
      if (i + 1 == y - 1) S.prep_next = PREP_CG;
      if (i == rn - 2) { S.row_off = (int32_t)np.rl_row.w_off; S.row_bias_off = (int32_t)np.rl_row.bias_off; S.row_n = 3; S.post = POST_DRGB; }
    }
  }
  pg->n_steps = n;
}

constexpr size_t TC_SMEM_BYTES = 2 * A_SLABS * SLAB_BYTES + TC_STAGES * STAGE_BYTES + 128 + 1024;

int launch_shade_tc(const NetPack& np, const float* packed, const ShadeArgs& a, float* gxscratch, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CNEUS_CUDA_CHECK(cudaFuncSetAttribute(shade_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    attr_set = true;
  }
  if (a.P <= 0) return CNEUS_OK;
  if (a.run_grad && a.dscratch == nullptr) { set_error("gradient stage needs the activation-derivative scratch"); return CNEUS_EINVAL; }
  TcProgram pg;
  build_program(np, a, &pg);
  if (pg.n_steps > MAX_TC_STEPS) { set_error("tensor-core program too long"); return CNEUS_EUNSUPPORTED; }
  int64_t tiles = (a.P + TCM - 1) / TCM;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int grid = (int)(tiles < sms ? tiles : sms);
  shade_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(pg, packed, a, gxscratch);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  return CNEUS_OK;
}

}  // namespace cneus
