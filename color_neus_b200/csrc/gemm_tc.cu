// Tensor-core GEMMs of the training backward (backward.cu) on tcgen05: fp32 operands in global memory, split on the
// fly into fp16 hi/lo planes (x = hi + lo), three MMAs per k-step (hi*hi + lo*hi + hi*lo) into one fp32 TMEM accumulator.
//
//   NT: C[m,n] = sum_k A[m,k] * B[n,k]      (y = x W^T)            M = points, N <= 256, K <= 320
//   NN: C[m,n] = sum_k A[m,k] * B[k,n]      (dL/dx = dL/dy W)      M = points, N <= 256, K <= 320
//   TN: C[m,n] = sum_k A[k,m] * B[k,n]      (dL/dW = dL/dy^T x)    K = points, M <= 256, N <= 256 (split over points)
//
// fp16 has a narrow exponent range and the adjoints of the backward pass are tiny, so every block of rows that is
// converted together (a 128-row tile of A; a CTA's range of points for TN) is scaled by a power of two taken from the
// block's largest magnitude (range_amax_kernel) and the factor is undone in the epilogue.  Weights (the B operand of
// NT / NN) are converted once per call into the same pre-swizzled stage images the forward kernel streams
// (pack_b_kernel, scale 2^6) and fetched with cp.async.bulk.
//
// NT / NN kernel: persistent CTAs, tile = 128 rows; 16 worker warps convert the A tile slab by slab (64 k) straight
// from global memory into SWIZZLE_128B shared-memory slabs and, while the MMAs of that tile run, write the previous
// tile's accumulators (TMEM double buffer: 2 x 256 columns) to global memory with the fused epilogue
// (scale, bias, ReLU, mask, accumulate); warp 16 streams the weight stages, warp 17 issues the MMAs.
// These GEMMs are HBM-bound (128 KB in + 128 KB out per tile against 6.1 k cycles of MMA work).
#include "backward.cuh"
#include "tc_ptx.cuh"

namespace cneus {

namespace {

constexpr int GT_M = 128;
constexpr int GT_SLAB_BYTES = 16384;           // [128 rows][64 halfs], SWIZZLE_128B
constexpr int GT_SLABS = 5;                    // K <= 320
constexpr int GT_STAGE_BYTES = 32768;          // weight stage: hi + lo slab of [256 rows][32 halfs], SWIZZLE_64B
constexpr int GT_STAGES = 2;
constexpr int GT_WORKERS = 16;
constexpr int GT_THREADS = GT_WORKERS * 32 + 64;
constexpr float GT_WSCALE = 64.0f;
constexpr size_t GT_SMEM = 2 * GT_SLABS * GT_SLAB_BYTES + GT_STAGES * GT_STAGE_BYTES + 256 + 1024;

// power-of-two scale that maps a block's largest magnitude into [2^12, 2^13)
__device__ __forceinline__ float block_scale(float amax) {
  if (!(amax > 0.0f) || !isfinite(amax)) return 1.0f;
  int e = ilogbf(amax);
  e = 12 - e;
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  return ldexpf(1.0f, e);
}

// ---------------------------------------------------------------------------------------------------------
// largest |a[r, c]| of every block of `rows_per_block` consecutive rows (c < ncols)
// ---------------------------------------------------------------------------------------------------------
// grid = (ranges, sub-blocks per range); out must be zeroed; non-negative floats order like their bit patterns
__global__ void range_amax_kernel(const float* __restrict__ a, int64_t lda, int64_t rows, int ncols, int64_t rows_per_block, int vec4,
                                  float* __restrict__ out) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < rows) ? r0 + rows_per_block : rows;
  const int lane = threadIdx.x & 31;
  const int wid = (int)(blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5)), nw = (int)(gridDim.y * (blockDim.x >> 5));
  float m = 0.0f;
  for (int64_t r = r0 + wid; r < r1; r += nw) {
    const float* row = a + r * lda;
    if (vec4) {
      for (int c = lane * 4; c < ncols; c += 128) {
        const float4 t = *reinterpret_cast<const float4*>(row + c);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(t.x), fabsf(t.y))), fmaxf(fabsf(t.z), fabsf(t.w)));
      }
    } else {
      for (int c = lane; c < ncols; c += 32) m = fmaxf(m, fabsf(row[c]));
    }
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0 && m > 0.0f) atomicMax(reinterpret_cast<unsigned int*>(out) + blockIdx.x, __float_as_uint(m));
}
int launch_range_amax(const float* a, int64_t lda, int64_t rows, int ncols, int64_t rows_per_block, float* out, cudaStream_t st) {
  const int64_t ranges = (rows + rows_per_block - 1) / rows_per_block;
  CNEUS_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)ranges * sizeof(float), st));
  int sub = (int)((148 * 8 + ranges - 1) / ranges);
  const int64_t max_sub = (rows_per_block + 7) / 8;
  if (sub > max_sub) sub = (int)max_sub;
  if (sub < 1) sub = 1;
  const int vec4 = (lda % 4 == 0) && (ncols % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0);
  range_amax_kernel<<<dim3((unsigned)ranges, (unsigned)sub), 256, 0, st>>>(a, lda, rows, ncols, rows_per_block, vec4, out);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  return CNEUS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// weight operand -> stage images [kb][sh][hi | lo], element (n, k): NT B[n][k], NN B[k][n]; rows >= N and k >= K are 0
// ---------------------------------------------------------------------------------------------------------
__global__ void pack_b_kernel(const float* __restrict__ B, int ldb, int N, int K, int transposed, int n_kb, uint8_t* __restrict__ img) {
  const int64_t total = (int64_t)n_kb * 2 * 256 * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i & 31);
    const int r = (int)((i >> 5) & 255);
    const int stage = (int)(i >> 13);  // kb * 2 + sh
    const int k = (stage >> 1) * 64 + (stage & 1) * 32 + kk;
    float val = 0.0f;
    if (r < N && k < K) val = (transposed ? B[(int64_t)k * ldb + r] : B[(int64_t)r * ldb + k]) * GT_WSCALE;
    const __half h = __float2half_rn(val);
    const __half l = __float2half_rn(val - __half2float(h));
    const int chunk = (kk >> 3) ^ ((r >> 1) & 3);
    const size_t off = (size_t)stage * GT_STAGE_BYTES + (size_t)(r >> 3) * 512 + (size_t)(r & 7) * 64 + (size_t)chunk * 16 + (size_t)(kk & 7) * 2;
    *reinterpret_cast<__half*>(img + off) = h;
    *reinterpret_cast<__half*>(img + off + GT_SLAB_BYTES) = l;
  }
}

struct TcGemmArgs {
  const float* A;
  float* C;
  const float* bias;
  const float* mask;
  const float* amax;     // [tiles]
  const uint8_t* bimg;   // stage images
  int64_t M;
  int N, K;
  int64_t lda, ldc, ldmask;
  float alpha;
  int accumulate, relu;
  int n_kb;              // 64-wide K blocks
  int a_vec4;            // rows of A are 16-byte aligned: 128-bit loads
  int c_vec4;            // rows of C (and mask) are 16-byte aligned
};

// 4 consecutive k of one row -> 8 bytes in each plane of a K-major SWIZZLE_128B slab
__device__ __forceinline__ void store_a4(uint8_t* a_hi, uint8_t* a_lo, int slab, int row, int kq, float x0, float x1, float x2, float x3) {
  const __half2 h0 = __floats2half2_rn(x0, x1), h1 = __floats2half2_rn(x2, x3);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(x0 - f0.x, x1 - f0.y), l1 = __floats2half2_rn(x2 - f1.x, x3 - f1.y);
  const uint32_t off = (uint32_t)slab * GT_SLAB_BYTES + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u +
                       (uint32_t)(((kq >> 3) ^ (row & 7)) << 4) + (uint32_t)(kq & 7) * 2u;
  *reinterpret_cast<uint2*>(a_hi + off) = make_uint2(pack_h2(h0), pack_h2(h1));
  *reinterpret_cast<uint2*>(a_lo + off) = make_uint2(pack_h2(l0), pack_h2(l1));
}

__global__ void __launch_bounds__(GT_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + GT_SLABS * GT_SLAB_BYTES;
  uint8_t* wring = smem + 2 * GT_SLABS * GT_SLAB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wring + GT_STAGES * GT_STAGE_BYTES);
  uint64_t* a_full = bars;          // [5] slab converted (workers -> MMA)
  uint64_t* a_empty = bars + 5;     // [5] slab consumed (MMA -> workers)
  uint64_t* b_full = bars + 10;     // [2]
  uint64_t* b_empty = bars + 12;    // [2]
  uint64_t* acc_full = bars + 14;   // [2] accumulators complete (MMA -> workers)
  uint64_t* acc_empty = bars + 16;  // [2] accumulators drained (workers -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 5; ++i) { mbar_init(&a_full[i], GT_WORKERS); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], GT_WORKERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GT_WORKERS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t n_tiles = (g.M + GT_M - 1) / GT_M;
  const int n_kb = g.n_kb;
  auto ksteps_of = [&](int kb) -> int { const int rem = g.K - 64 * kb; return rem >= 64 ? 4 : (rem + 15) / 16; };

  if (warp == GT_WORKERS) {
    // ================================================================ weight stages (bulk async copies)
    if (lane == 0) {
      uint32_t stg = 0, ph = 1;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < n_kb; ++kb) {
          const int nks = ksteps_of(kb);
          for (int sh = 0; sh < 2; ++sh) {
            if (nks <= 2 * sh) continue;
            mbar_wait(&b_empty[stg], ph);
            mbar_expect_tx(&b_full[stg], GT_STAGE_BYTES);
            const uint8_t* img = g.bimg + (size_t)(kb * 2 + sh) * GT_STAGE_BYTES;
            bulk_g2s(wring + stg * GT_STAGE_BYTES, img, GT_SLAB_BYTES, &b_full[stg]);
            bulk_g2s(wring + stg * GT_STAGE_BYTES + GT_SLAB_BYTES, img + GT_SLAB_BYTES, GT_SLAB_BYTES, &b_full[stg]);
            if (++stg == GT_STAGES) { stg = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == GT_WORKERS + 1) {
    // ================================================================ MMA issuer (warp-uniform, one elected lane)
    const uint32_t n_mma = g.N <= 128 ? 128u : 256u;
    const uint32_t idesc = (1u << 4) | ((n_mma >> 3) << 17) | ((uint32_t)(GT_M >> 4) << 24);
    constexpr uint32_t HI_SW128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t HI_SW64 = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);
    const uint32_t a_hi_lo32 = ((smem_u32(a_hi) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t a_lo_lo32 = ((smem_u32(a_lo) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t b_lo32 = ((smem_u32(wring) >> 4) & 0x3FFFu) | 0x10000u;
    auto desc = [](uint32_t hi, uint32_t lo) -> uint64_t { return ((uint64_t)hi << 32) | lo; };
    const bool leader = elect_one();
    uint32_t stg = 0, ph = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem + buf * 256u;
      uint32_t accum = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&a_full[kb], it & 1u);
        tc_fence_after();
        const uint32_t a_off = (uint32_t)kb * (GT_SLAB_BYTES >> 4);
        const int nks = ksteps_of(kb);
        for (int sh = 0; sh < 2; ++sh) {
          const int nk = nks - 2 * sh;
          if (nk <= 0) continue;
          mbar_wait(&b_full[stg], ph);
          tc_fence_after();
          const uint32_t bh = b_lo32 + stg * (GT_STAGE_BYTES >> 4), bl = bh + (GT_SLAB_BYTES >> 4);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (k < nk) {
                const uint32_t ka = a_off + (uint32_t)(sh * 2 + k) * 2u;
                const uint64_t dAh = desc(HI_SW128, a_hi_lo32 + ka), dAl = desc(HI_SW128, a_lo_lo32 + ka);
                const uint64_t dBh = desc(HI_SW64, bh + (uint32_t)k * 2u), dBl = desc(HI_SW64, bl + (uint32_t)k * 2u);
                mma_f16(d_tmem, dAh, dBh, idesc, accum);
                mma_f16(d_tmem, dAl, dBh, idesc, 1u);
                mma_f16(d_tmem, dAh, dBl, idesc, 1u);
                accum = 1u;
              }
            }
            mma_commit(&b_empty[stg]);
          }
          __syncwarp();
          if (++stg == GT_STAGES) { stg = 0; ph ^= 1u; }
        }
        if (leader) mma_commit(&a_empty[kb]);  // the slab may be refilled with the next tile
        __syncwarp();
      }
      if (leader) mma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ================================================================ workers: convert the A tile, drain the previous one
    const int q = warp & 3, cgp = warp >> 2;   // TMEM lane quarter / 64-column group of the epilogue
    const int erow = q * 32 + lane;
    auto epilogue = [&](int64_t tile, uint32_t it, float tile_sc) {
      const uint32_t buf = it & 1u;
      mbar_wait(&acc_full[buf], (it >> 1) & 1u);
      tc_fence_after();
      const float osc = g.alpha / (GT_WSCALE * tile_sc);
      const int64_t gm = tile * GT_M + erow;
      const uint32_t taddr = tmem + buf * 256u + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int n0 = cgp * 64 + c * 16;
        if (n0 >= g.N) break;
        uint32_t raw[16];
        tmem_ld16_nowait(taddr + n0, raw);
        tmem_ld_wait();
        if (gm < g.M) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]) * osc;
          if (g.bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (n0 + i < g.N) v[i] += __ldg(g.bias + n0 + i);
          }
          // same order as the SGEMM epilogue (gemm.cu): accumulate, ReLU, mask
          float* crow = g.C + gm * g.ldc + n0;
          const float* mrow = g.mask ? g.mask + gm * g.ldmask + n0 : nullptr;
          const bool vec = g.c_vec4 && n0 + 16 <= g.N;
          if (g.accumulate) {
            if (vec) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 c4 = *reinterpret_cast<const float4*>(crow + 4 * i);
                v[4 * i] += c4.x; v[4 * i + 1] += c4.y; v[4 * i + 2] += c4.z; v[4 * i + 3] += c4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (n0 + i < g.N) v[i] += crow[i];
            }
          }
          if (g.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
          }
          if (mrow) {
            if (vec) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 m4 = *reinterpret_cast<const float4*>(mrow + 4 * i);
                v[4 * i] = m4.x > 0.f ? v[4 * i] : 0.f; v[4 * i + 1] = m4.y > 0.f ? v[4 * i + 1] : 0.f;
                v[4 * i + 2] = m4.z > 0.f ? v[4 * i + 2] : 0.f; v[4 * i + 3] = m4.w > 0.f ? v[4 * i + 3] : 0.f;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (n0 + i < g.N) v[i] = mrow[i] > 0.f ? v[i] : 0.f;
            }
          }
          if (vec) {
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(crow + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (n0 + i < g.N) crow[i] = v[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    };

    uint32_t it = 0;
    int64_t prev_tile = -1;
    float prev_sc = 1.0f;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const float sc = block_scale(__ldg(g.amax + tile));
      // slabs in pairs: the loads of both (8 x 128 bits per thread, 64 KB per CTA) are in flight before the first use,
      // which is enough to keep this SM's share of HBM busy; rows warp*8 .. warp*8+7 of a slab, two rows per pass
      // (one per half-warp), 4 consecutive k per lane
      for (int kb0 = 0; kb0 < n_kb; kb0 += 2) {
        float x[2][4][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int p2 = 0; p2 < 4; ++p2) {
            const int row = warp * 8 + p2 * 2 + (lane >> 4);
            const int k = (kb0 + j) * 64 + (lane & 15) * 4;
            const int64_t gm = tile * GT_M + row;
            x[j][p2][0] = 0.f; x[j][p2][1] = 0.f; x[j][p2][2] = 0.f; x[j][p2][3] = 0.f;
            if (gm < g.M && kb0 + j < n_kb) {
              const float* src = g.A + gm * g.lda + k;
              if (g.a_vec4 && k + 3 < g.K) {
                const float4 t = *reinterpret_cast<const float4*>(src);
                x[j][p2][0] = t.x; x[j][p2][1] = t.y; x[j][p2][2] = t.z; x[j][p2][3] = t.w;
              } else {
                if (k < g.K) x[j][p2][0] = src[0];
                if (k + 1 < g.K) x[j][p2][1] = src[1];
                if (k + 2 < g.K) x[j][p2][2] = src[2];
                if (k + 3 < g.K) x[j][p2][3] = src[3];
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int kb = kb0 + j;
          if (kb < n_kb) {
            mbar_wait(&a_empty[kb], (it & 1u) ^ 1u);
#pragma unroll
            for (int p2 = 0; p2 < 4; ++p2) {
              const int row = warp * 8 + p2 * 2 + (lane >> 4);
              store_a4(a_hi, a_lo, kb, row, (lane & 15) * 4, x[j][p2][0] * sc, x[j][p2][1] * sc, x[j][p2][2] * sc, x[j][p2][3] * sc);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[kb]);
          }
        }
      }
      if (prev_tile >= 0) epilogue(prev_tile, it - 1, prev_sc);
      prev_tile = tile;
      prev_sc = sc;
    }
    if (prev_tile >= 0) epilogue(prev_tile, it - 1, prev_sc);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == GT_WORKERS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}


// ---------------------------------------------------------------------------------------------------------
// TN: C[m, n] (partial over a range of points) = sum_p A[p, m] * B[p, n].  Both operands have the reduction index as
// their slow dimension, so the workers transpose 8x8 blocks on the way into the K-major slabs with stmatrix.trans
// (one 16-byte row = 8 consecutive points of one column; the swizzled row addresses are conflict-free).
// grid = (m tiles of 128, splits); each CTA writes its [M, N] partial, reduced in fixed order afterwards.
// ---------------------------------------------------------------------------------------------------------
struct TcTnArgs {
  const float* A;        // [K, M] (lda)
  const float* B;        // [K, N] (ldb)
  float* partial;        // [splits][M][N]
  const float* amax_a;   // [splits]
  const float* amax_b;   // [splits]
  int64_t K, per;        // points, points per split (multiple of 64)
  int M, N;
  int64_t lda, ldb;
  int a_vec2, b_vec2;
};
constexpr int TN_STAGE_BYTES = 2 * GT_SLAB_BYTES + 2 * 2 * GT_SLAB_BYTES;  // A hi|lo (128 rows) + B hi|lo (256 rows)
constexpr int TN_STAGES = 2;
constexpr size_t TN_SMEM = (size_t)TN_STAGES * TN_STAGE_BYTES + 256 + 1024;

__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
               : "memory");
}

// one unit: 8 points (k0 .. k0+7 of the stage) x 32 columns (c0 .. c0+31) of a [points, cols] fp32 matrix -> 4 transposed
// 8x8 blocks in the hi and lo plane of a K-major SWIZZLE_128B slab whose rows are the columns.  Loads and stores are
// separate so that the loads of all units of a stage are in flight together.
__device__ __forceinline__ void tn_unit_load(const float* __restrict__ src, int64_t ld, int64_t p_base, int64_t p_end, int ncols, int vec2,
                                             int k0, int c0, int lane, float (&x)[8]) {
  const int64_t p = p_base + k0 + (lane >> 2);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + 8 * j + 2 * (lane & 3);
    x[2 * j] = 0.f; x[2 * j + 1] = 0.f;
    if (p < p_end) {
      const float* q = src + p * ld + c;
      if (vec2 && c + 1 < ncols) {
        const float2 t = *reinterpret_cast<const float2*>(q);
        x[2 * j] = t.x; x[2 * j + 1] = t.y;
      } else {
        if (c < ncols) x[2 * j] = q[0];
        if (c + 1 < ncols) x[2 * j + 1] = q[1];
      }
    }
  }
}
__device__ __forceinline__ void tn_unit_store(const float (&x)[8], int k0, int c0, float sc, uint32_t hi_base, uint32_t lo_base, int lane) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float x0 = x[2 * j] * sc, x1 = x[2 * j + 1] * sc;
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    hi[j] = pack_h2(h);
    lo[j] = pack_h2(__floats2half2_rn(x0 - hf.x, x1 - hf.y));
  }
  // lane l supplies the address of stored row (l & 7) of block (l >> 3): slab row = column index, chunk = k0 / 8
  const int r = c0 + 8 * (lane >> 3) + (lane & 7);
  const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)(((k0 >> 3) ^ (r & 7)) << 4);
  stmatrix_x4_trans(hi_base + off, hi[0], hi[1], hi[2], hi[3]);
  stmatrix_x4_trans(lo_base + off, lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __launch_bounds__(GT_THREADS, 1) tc_gemm_tn_kernel(const __grid_constant__ TcTnArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TN_STAGES * TN_STAGE_BYTES);
  uint64_t* full = bars;       // [2] stage converted (workers -> MMA)
  uint64_t* empty = bars + 2;  // [2] stage consumed (MMA -> workers)
  uint64_t* acc_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], GT_WORKERS); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GT_WORKERS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int mt = blockIdx.x, z = blockIdx.y;
  const int64_t p0 = (int64_t)z * g.per;
  const int64_t p1 = (p0 + g.per < g.K) ? p0 + g.per : g.K;
  const int n_blk = (int)((p1 - p0 + 63) / 64);
  const int n_bunits = (g.N + 31) / 32 * 8;  // units of the B part of a stage

  if (warp == GT_WORKERS + 1) {
    const uint32_t n_mma = g.N <= 128 ? 128u : 256u;
    const uint32_t idesc = (1u << 4) | ((n_mma >> 3) << 17) | ((uint32_t)(GT_M >> 4) << 24);
    constexpr uint32_t HI_SW128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    auto desc = [](uint32_t hi, uint32_t lo) -> uint64_t { return ((uint64_t)hi << 32) | lo; };
    const bool leader = elect_one();
    uint32_t accum = 0;
    for (int b = 0; b < n_blk; ++b) {
      const uint32_t stg = b & 1u;
      mbar_wait(&full[stg], (b >> 1) & 1u);
      tc_fence_after();
      const uint32_t base = ((smem_u32(smem + stg * TN_STAGE_BYTES) >> 4) & 0x3FFFu) | 0x10000u;
      const uint32_t ah = base, al = base + (GT_SLAB_BYTES >> 4), bh = base + 2 * (GT_SLAB_BYTES >> 4), bl = bh + 2 * (GT_SLAB_BYTES >> 4);
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t ko = (uint32_t)k * 2u;
          mma_f16(tmem, desc(HI_SW128, ah + ko), desc(HI_SW128, bh + ko), idesc, accum);
          mma_f16(tmem, desc(HI_SW128, al + ko), desc(HI_SW128, bh + ko), idesc, 1u);
          mma_f16(tmem, desc(HI_SW128, ah + ko), desc(HI_SW128, bl + ko), idesc, 1u);
          accum = 1u;
        }
        mma_commit(&empty[stg]);
      }
      __syncwarp();
    }
    if (leader) mma_commit(acc_full);
    __syncwarp();
  } else if (warp < GT_WORKERS) {
    const float sa = block_scale(__ldg(g.amax_a + z)), sb = block_scale(__ldg(g.amax_b + z));
    for (int b = 0; b < n_blk; ++b) {
      const uint32_t stg = b & 1u;
      uint8_t* st = smem + stg * TN_STAGE_BYTES;
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + GT_SLAB_BYTES, b_hi = a_hi + 2 * GT_SLAB_BYTES, b_lo = b_hi + 2 * GT_SLAB_BYTES;
      const int64_t pb = p0 + (int64_t)b * 64;
      // units 0..31: the 128 columns [mt*128, mt*128+128) of A (this m tile); 32..: the columns of B.  Up to 6 units per
      // warp; all their loads are issued before the stage is claimed
      float x[6][8];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int u = warp + GT_WORKERS * i;
        const bool is_a = u < 32;
        const int v = is_a ? u : u - 32;
        if (u < 32 + n_bunits)
          tn_unit_load(is_a ? g.A + (int64_t)mt * GT_M : g.B, is_a ? g.lda : g.ldb, pb, p1, is_a ? g.M - mt * GT_M : g.N,
                       is_a ? g.a_vec2 : g.b_vec2, (v & 7) * 8, (v >> 3) * 32, lane, x[i]);
      }
      mbar_wait(&empty[stg], ((b >> 1) & 1u) ^ 1u);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int u = warp + GT_WORKERS * i;
        const bool is_a = u < 32;
        const int v = is_a ? u : u - 32;
        if (u < 32 + n_bunits)
          tn_unit_store(x[i], (v & 7) * 8, (v >> 3) * 32, is_a ? sa : sb, is_a ? a_hi : b_hi, is_a ? a_lo : b_lo, lane);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stg]);
    }
    // epilogue: this CTA's partial of rows [mt*128, mt*128+128)
    const int q = warp & 3, cgp = warp >> 2;
    const int m = mt * GT_M + q * 32 + lane;
    float* prow = g.partial + ((int64_t)z * g.M + m) * g.N;
    if (n_blk > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    const float osc = 1.0f / (sa * sb);
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int n0 = cgp * 64 + c * 16;
      if (n0 >= g.N) break;
      uint32_t raw[16];
      if (n_blk > 0) {
        tmem_ld16_nowait(taddr + n0, raw);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) raw[i] = 0u;
      }
      if (m < g.M) {
#pragma unroll
        for (int i = 0; i < 16; ++i) if (n0 + i < g.N) prow[n0 + i] = __uint_as_float(raw[i]) * osc;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == GT_WORKERS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---------------------------------------------------------------------------------------------------------
// narrow products (one operand has at most 8 columns): memory-bound CUDA-core kernels
// ---------------------------------------------------------------------------------------------------------
// partial[z][m][j] = sum_{p in chunk z} A[p, m] * B[p, j]      (M arbitrary, J <= 8)
__global__ void small_tn_kernel(const float* __restrict__ A, int64_t lda, int M, const float* __restrict__ B, int64_t ldb, int J,
                                int64_t K, float* __restrict__ partial) {
  const int64_t per = (K + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = (int64_t)blockIdx.x * per, p1 = (p0 + per < K) ? p0 + per : K;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 8
    for (int64_t p = p0; p < p1; ++p) {
      const float a = A[p * lda + m];
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < J) acc[j] = fmaf(a, __ldg(B + p * ldb + j), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) if (j < J) partial[((int64_t)blockIdx.x * M + m) * J + j] = acc[j];
  }
}
// C[m * cs_m + j * cs_j] (+)= sum_z partial[z][m][j]
__global__ void small_tn_reduce_kernel(const float* __restrict__ partial, int chunks, int M, int J, float* __restrict__ C, int64_t cs_m,
                                       int64_t cs_j, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * J) return;
  double s = 0.0;
  for (int zc = 0; zc < chunks; ++zc) s += (double)partial[(int64_t)zc * M * J + i];
  const int m = i / J, j = i % J;
  float* c = C + m * cs_m + j * cs_j;
  *c = (float)(accumulate ? (double)*c + s : s);
}
// C[p, j] (+)= sum_k A[p, k] * B[j * bs_j + k * bs_k] (+ bias[j])      (J <= 4; one warp per row)
__global__ void small_nt_kernel(const float* __restrict__ A, int64_t lda, int64_t M, int K, const float* __restrict__ B, int64_t bs_j,
                                int64_t bs_k, int J, const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < M; p += nwarps) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = lane; k < K; k += 32) {
      const float a = A[p * lda + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) if (j < J) acc[j] = fmaf(a, __ldg(B + j * bs_j + k * bs_k), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    if (lane < J) {
      float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
      if (bias) v += bias[lane];
      float* c = C + p * ldc + lane;
      *c = accumulate ? *c + v : v;
    }
  }
}

}  // namespace

// largest |a[r, c]| of the whole [rows, ncols] tensor -> out[0] (device)
int tensor_amax(const float* a, int64_t lda, int64_t rows, int ncols, float* out, cudaStream_t st) {
  return launch_range_amax(a, lda, rows, ncols, rows, out, st);
}

// workspace layout of the tensor-core GEMMs (floats): [amax: 4096][weight stage images: 5 * 2 * 32 KB]
size_t tc_gemm_ws_floats() { return 4096 + (size_t)GT_SLABS * 2 * GT_STAGE_BYTES / sizeof(float) + 64; }

bool tc_gemm_supported(int mode, const GemmArgs& g) {
  if (mode == GEMM_TN) return false;
  if (g.N > 256 || g.N < 16 || g.K > 64 * GT_SLABS || g.K < 16 || g.M < GT_M) return false;
  if ((g.M + GT_M - 1) / GT_M > 4096) return false;
  return true;
}

int launch_gemm_tc(int mode, const GemmArgs& g, float* ws, cudaStream_t st) {
  static bool attr_set[CNEUS_MAX_DEVICES] = {false};
  if (first_use_on_device(attr_set)) {
    CNEUS_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
  }
  float* amax = ws;
  uint8_t* bimg = reinterpret_cast<uint8_t*>(ws + 4096);
  const int64_t tiles = (g.M + GT_M - 1) / GT_M;
  const int n_kb = (int)((g.K + 63) / 64);
  { int rc = launch_range_amax(g.A, g.lda, g.M, (int)g.K, GT_M, amax, st); if (rc != CNEUS_OK) return rc; }
  pack_b_kernel<<<80, 256, 0, st>>>(g.B, g.ldb, g.N, (int)g.K, mode == GEMM_NN ? 1 : 0, n_kb, bimg);
  TcGemmArgs a;
  a.A = g.A; a.C = g.C; a.bias = g.bias; a.mask = g.mask; a.amax = amax; a.bimg = bimg;
  a.M = g.M; a.N = g.N; a.K = (int)g.K; a.lda = g.lda; a.ldc = g.ldc; a.ldmask = g.ldmask;
  a.alpha = g.alpha; a.accumulate = g.accumulate; a.relu = g.relu; a.n_kb = n_kb;
  a.a_vec4 = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  a.c_vec4 = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
             (!g.mask || ((g.ldmask % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.mask) & 15) == 0)));
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  tc_gemm_kernel<<<(unsigned)(tiles < sms ? tiles : sms), GT_THREADS, GT_SMEM, st>>>(a);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch(3);
  return CNEUS_OK;
}


// ---- TN (weight gradients): C[M, N] (+)= A[K, M]^T B[K, N], 16 <= M, N <= 256; `partial` holds [splits][M][N]
constexpr int TN_SPLITS = 74;
size_t tc_gemm_tn_partial_floats() { return (size_t)TN_SPLITS * 256 * 256; }
bool tc_gemm_tn_supported(const GemmArgs& g) { return g.M >= 16 && g.M <= 256 && g.N >= 16 && g.N <= 256 && g.K >= 64; }

int launch_gemm_tn_tc(const GemmArgs& g, float* ws, float* partial, cudaStream_t st) {
  static bool attr_set[CNEUS_MAX_DEVICES] = {false};
  if (first_use_on_device(attr_set)) {
    CNEUS_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TN_SMEM));
  }
  const int64_t blocks64 = (g.K + 63) / 64;
  const int64_t per = (blocks64 + TN_SPLITS - 1) / TN_SPLITS * 64;
  const int splits = (int)((g.K + per - 1) / per);
  float* amax_a = ws;
  float* amax_b = ws + 2048;
  { int rc = launch_range_amax(g.A, g.lda, g.K, g.M, per, amax_a, st); if (rc != CNEUS_OK) return rc; }
  { int rc = launch_range_amax(g.B, g.ldb, g.K, g.N, per, amax_b, st); if (rc != CNEUS_OK) return rc; }
  TcTnArgs a;
  a.A = g.A; a.B = g.B; a.partial = partial; a.amax_a = amax_a; a.amax_b = amax_b; a.K = g.K; a.per = per; a.M = g.M; a.N = g.N;
  a.lda = g.lda; a.ldb = g.ldb;
  a.a_vec2 = (g.lda % 2 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 7) == 0);
  a.b_vec2 = (g.ldb % 2 == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 7) == 0);
  tc_gemm_tn_kernel<<<dim3((g.M + GT_M - 1) / GT_M, splits), GT_THREADS, TN_SMEM, st>>>(a);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch(3);
  return splitk_reduce(partial, splits, g, st);
}

// ---- narrow products
int launch_small_tn(const float* A, int64_t lda, int M, const float* B, int64_t ldb, int J, int64_t K, float* C, int64_t cs_m, int64_t cs_j,
                    int accumulate, float* partial, cudaStream_t st) {
  if (J > 8) { set_error("small_tn: J > 8"); return CNEUS_EINVAL; }
  const int chunks = 296;
  small_tn_kernel<<<chunks, 256, 0, st>>>(A, lda, M, B, ldb, J, K, partial);
  small_tn_reduce_kernel<<<(M * J + 127) / 128, 128, 0, st>>>(partial, chunks, M, J, C, cs_m, cs_j, accumulate);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch(2);
  return CNEUS_OK;
}
// C[p, n] = sum_{k < K} A[p, k] * B[k * ldb + n], K <= 4 (rank-K update of a [M, N] block; same epilogue order as the GEMMs)
__global__ void small_k_nn_kernel(const float* __restrict__ A, int64_t lda, int64_t M, int K, const float* __restrict__ B, int64_t ldb,
                                  int N, float* __restrict__ C, int64_t ldc, const float* __restrict__ mask, int64_t ldmask,
                                  int accumulate, int relu) {
  const int64_t total = M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / N;
    const int n = (int)(i % N);
    float v = 0.f;
    for (int k = 0; k < K; ++k) v = fmaf(A[p * lda + k], __ldg(B + k * ldb + n), v);
    if (accumulate) v += C[p * ldc + n];
    if (relu) v = fmaxf(v, 0.f);
    if (mask) v = mask[p * ldmask + n] > 0.f ? v : 0.f;
    C[p * ldc + n] = v;
  }
}
int launch_small_k_nn(const GemmArgs& g, cudaStream_t st) {
  small_k_nn_kernel<<<148 * 16, 256, 0, st>>>(g.A, g.lda, g.M, (int)g.K, g.B, g.ldb, g.N, g.C, g.ldc, g.mask, g.ldmask, g.accumulate, g.relu);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return CNEUS_OK;
}
int launch_small_nt(const float* A, int64_t lda, int64_t M, int K, const float* B, int64_t bs_j, int64_t bs_k, int J, const float* bias,
                    float* C, int64_t ldc, int accumulate, cudaStream_t st) {
  if (J > 4) { set_error("small_nt: J > 4"); return CNEUS_EINVAL; }
  small_nt_kernel<<<148 * 8, 256, 0, st>>>(A, lda, M, K, B, bs_j, bs_k, J, bias, C, ldc, accumulate);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return CNEUS_OK;
}

}  // namespace cneus
