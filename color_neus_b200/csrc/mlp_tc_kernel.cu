// Tensor-core point-shading kernel for sm_100a: the same network evaluation as mlp_simt.cu (PE -> SDF MLP ->
// reverse-chain gradient -> colour MLP -> relight MLP) with every 256-wide layer on tcgen05.mma.
//
//   * one CTA per SM, tile = 128 points = the 128 TMEM lanes; the CTAs run as PAIRS (clusters of two on the SMs of a TPC):
//     every MMA is a cta_group::2 instruction with M = 256 over both tiles, issued by the leader, and each CTA stages only
//     its half of the rows of every weight slab (half the L2 -> shared-memory weight bytes and two thirds of the
//     tensor core's shared-memory operand reads per CTA).  The peer's epilogue warps announce their A slabs on the
//     leader's barriers (mapa + remote arrive), the peer's MMA warp relays the arrival of the peer's weight halves,
//     tcgen05.commit multicasts "slot free" / "accumulators ready" to both CTAs (-DCNEUS_TC_SINGLE: the one-CTA variant);
//   * fp32 fidelity on fp16 tensor cores: every operand is split x = hi + lo (two fp16 planes) and every algorithmic MAC
//     is three MMAs: lo*hi + hi*lo (corrections) and hi*hi (main).  The tensor core truncates on every accumulate, so
//     2^-11 sized terms must not be added to a big running sum: a layer issues ALL its correction products first (two
//     per k-step, small among themselves) and the main products last, into ONE 256-column fp32 TMEM accumulator.  Same
//     accuracy as round 1's separate correction accumulator (measured on B200, profiles/r2b_parity_errors*.json), but a
//     layer needs 256 instead of 512 TMEM columns, so consecutive layers alternate between the two halves of TMEM and
//     the epilogue reads its accumulator in place, section by section, while the next layer's MMAs fill the other
//     half: no drain into registers, no register rotation, no correction add.  Price: the hi plane of the weights is
//     streamed twice (1.5x the L2 -> shared-memory traffic).  Weights are pre-scaled by 2^6 so their lo plane stays in
//     the fp16 normal range;
//   * A operand (activations) lives in shared memory as K-major SWIZZLE_128B slabs written by the epilogue threads;
//     B operand (weights) is streamed from the L2-resident packed buffer by cp.async.bulk (TMA engine) through an
//     mbarrier ring (pairs: three to five 16 KB slots = this CTA's half of a stage; one CTA: two or three 32 KB stages),
//     already in their shared-memory image (pack_tc_kernel; [256 outputs]
//     [32 inputs] SWIZZLE_64B slabs): correction pass = hi + lo plane of a 32-wide half K-block per stage, main pass = the hi
//     planes of both halves of a K-block per stage; one N=256 MMA per product keeps operand reads at 96 B/clk so the
//     concurrent bulk-copy writes fit under the 128 B/clk shared-memory bandwidth;
//   * warp roles: 0-15 epilogue (warp w: TMEM lanes 32*(w%4).., columns [16 (w/4), +16) of every 64-column slab), 16
//     bulk-copy producer, 17 MMA issuer (one elected lane).  The 257-wide last SDF layer is split: the feature block is an
//     MMA whose result is kept in an fp32 scratch slot, the sdf column and the 3-wide colour / relight outputs are fp32 dot
//     products folded into the preceding epilogue (partial sums of a row's four threads exchanged through the consumed TMEM
//     accumulator).
#include <cuda_fp16.h>

#include "mlp_tc.cuh"
#include "tc_ptx.cuh"

namespace cneus {

// cycle counters of CTA 0 (cneus_tc_prof_read): [0] MMA thread waiting for the A operand, [1] waiting for weights,
// [2] MMA thread total, [3] steps, [4] epilogue thread 0 waiting for accumulators, [5] epilogue total,
// [6] producer waiting for a free ring slot, [7] producer total
__device__ unsigned long long g_tc_prof_type[16];  // profiling builds: see the end of the epilogue's step loop
__device__ unsigned long long g_tc_prof[32];  // [8..11] MMA thread waiting for slab barrier 0..3; [12..13] wait_acc / total of warp 12

// fine-grained epilogue timeline of thread 0 of CTA 0 (profiling builds only: -DCNEUS_TC_EPI_PROF), slots [16..23]:
// sections 0+1, (unused: was the drain), announce 0, section 2, announces 1+2, section 3, skip feed, PARK body, BWD_LAST body, announce 3,
// narrow-layer exchange + post, announce all (late steps), -, seed / colour-in / relight-in / cg staging
struct EpiProf {
#ifdef CNEUS_TC_EPI_PROF
  bool on;
  long long t;
  __device__ __forceinline__ void start() { if (on) t = clock64(); }
  __device__ __forceinline__ void mark(int i) {
    if (on) { const long long n = clock64(); g_tc_prof[16 + i] += (unsigned long long)(n - t); t = n; }
  }
#else
  __device__ __forceinline__ void start() {}
  __device__ __forceinline__ void mark(int) {}
#endif
};

// The role-level cycle counters (cneus_tc_prof_read) exist in profiling builds only (tools/build_prof.sh): as a run-time flag
// their bookkeeping was spilled to local memory and re-read right after the accumulator wait -- on the critical path of
// every step (2.4 % of the epilogue's samples in profiles/r2j).
#ifdef CNEUS_TC_EPI_PROF
constexpr bool kTcProf = true;
#else
constexpr bool kTcProf = false;
#endif

// kPair (mlp_tc.cuh; the product build, -DCNEUS_TC_SINGLE builds the one-CTA variant for A/B runs): the kernel is launched as
// clusters of two CTAs (the two SMs of a TPC) and every MMA is a cta_group::2 instruction over the tiles of both CTAs
// (M = 256); each CTA stages only its half of the rows of every weight slab.

constexpr float SP_K1 = 144.26950408889634f;    // 100 * log2(e)
constexpr float SP_K2 = 0.0069314718055994531f;  // ln(2) / 100

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory"); }

__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&f)[16]) {
  uint32_t m[16];
  tmem_ld16_nowait(taddr, m);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(m[i]);
}

// ---------------------------------------------------------------------------------------------------------
// A-operand writers (row = the calling thread's point)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t a_chunk_offset(int slab, int row, int chunk) {
  return (uint32_t)slab * SLAB_BYTES + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u +
         (uint32_t)((chunk ^ (row & 7)) << 4);
}
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
// Elements [q_lo, q_lo + 16) of the encoding [v | sin(2^k v) | cos(2^k v)]_{k < L} (3 (1 + 2L) values,
// PositionEncoding.py:51-76) of a 3-vector, written to out[q - q_lo] (q_lo may be negative).  One sincosf serves the sine
// and the cosine of a (frequency, dim) pair when both fall into the range; the range tests are warp-uniform.
__device__ __forceinline__ void pe_range16(const float (&v)[3], int L, int q_lo, float* out) {
#pragma unroll
  for (int d = 0; d < 3; ++d)
    if (d >= q_lo && d < q_lo + 16) out[d - q_lo] = v[d];
#pragma unroll 1
  for (int k = 0; k < L; ++k) {
    const float f = (float)(1 << k);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int qs = 3 + 6 * k + d, qc = qs + 3;
      const bool ws = qs >= q_lo && qs < q_lo + 16, wc = qc >= q_lo && qc < q_lo + 16;
      if (ws || wc) {
        float sn, cs;
        sincosf(v[d] * f, &sn, &cs);
        if (ws) out[qs - q_lo] = sn;
        if (wc) out[qc - q_lo] = cs;
      }
    }
  }
}

// softplus' in [0,1] is kept as 16-bit fixed point, two consecutive features per 32-bit word, layout [feature/2][point]:
// halves the scratch traffic so the slots of all CTAs (148 x 8 x 64 KB) stay L2-resident; abs. error <= 7.6e-6.
__device__ __forceinline__ uint32_t d_pack(float a, float b) {
  // round(x * 65535) sits in the low mantissa bits of x * 65535 + 2^23 (no conversion instruction)
  return __byte_perm(__float_as_uint(fmaf(a, 65535.0f, 8388608.0f)), __float_as_uint(fmaf(b, 65535.0f, 8388608.0f)), 0x5410);
}
__device__ __forceinline__ void d_unpack(uint32_t w, float& a, float& b) {
  a = (float)(w & 0xFFFFu) * (1.0f / 65535.0f);
  b = (float)(w >> 16) * (1.0f / 65535.0f);
}

// per-tile state of one point (held in registers by both threads that serve the row)
struct RowState {
  float pt[3], dir[3], xs[3], nrm[3], cg[3];
  float sdf;
};

enum { SMALL_PE = 0, SMALL_COLOR = 1, SMALL_RELIGHT = 2, SMALL_CG = 3 };

// elements [16 cq, 16 cq + 16) of the "small" input vectors staged in front of a layer (zero padded)
__device__ __forceinline__ void small_block16(const TcProgram& prog, const RowState& st, int kind, int viewdir_mode, int cq, float* out) {
  const int k0 = 16 * cq;
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = 0.f;
  if (kind == SMALL_PE) {
    pe_range16(st.xs, prog.multires, k0, out);  // multires == 0: the identity part only
    return;
  }
  if (kind == SMALL_CG) {
    if (cq == 0) { out[0] = st.cg[0]; out[1] = st.cg[1]; out[2] = st.cg[2]; }
    return;
  }
  // [pts | PE(view dir) | normal] with mode-dependent members (fields.py:167-172, :341-350)
  if (cq == 0) { out[0] = st.pt[0]; out[1] = st.pt[1]; out[2] = st.pt[2]; }
  int k = 3;
  const bool has_view = (kind == SMALL_RELIGHT) || prog.color_mode != CNEUS_COLOR_NO_VIEW_DIR;
  const int L = (kind == SMALL_RELIGHT) ? prog.relight_multires_view : prog.color_multires_view;
  if (has_view) {
    const bool neg_n = (kind == SMALL_COLOR && viewdir_mode == 1);  // extract_color: view dir = -normal (NeuS.py:60)
    float v[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) v[d] = neg_n ? -st.nrm[d] : st.dir[d];
    pe_range16(v, L, k0 - k, out);
    k += 3 * (1 + 2 * L);
  }
  const bool has_n = (kind == SMALL_RELIGHT) ? (prog.relight_include_grad != 0) : (prog.color_mode != CNEUS_COLOR_NO_NORMAL);
  if (has_n) {
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (k + d >= k0 && k + d < k0 + 16) out[k + d - k0] = st.nrm[d];
  }
}
// this thread's quarter (16 K values) of a 64-wide small-input slab
__device__ __forceinline__ void stage_small(uint8_t* a_hi, uint8_t* a_lo, int slab, int row, int cq, const TcProgram& prog,
                                            const RowState& st, int kind, int viewdir_mode) {
  float v[16];  // local array (dynamic indexing in the encoders; once or twice per tile)
  small_block16(prog, st, kind, viewdir_mode, cq, v);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = v[c * 8 + j];
    write_a8(a_hi, a_lo, slab, row, cq * 2 + c, o);
  }
}

// ---------------------------------------------------------------------------------------------------------
// hot epilogue loops.  A thread serves one point (row = TMEM lane) and, of every 64-column slab, the 16 columns
// [16 g, 16 g + 16) (g = column group of its warp), so the K-blocks of the next layer's A operand complete one after
// the other and the MMA issuer can start on slab 0 while the later slabs are still being computed.  Every section is
// read from this step's TMEM accumulator when its turn comes (the next step's MMAs fill the other half of TMEM) and its
// slab is announced (bar_slab) as soon as the thread's part of it is stored.
// ---------------------------------------------------------------------------------------------------------
// announce that this warp's part of a slab (and everything before it) is in place
__device__ __forceinline__ void slab_fence(int lane) {
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();    // this thread's TMEM accesses ordered before whatever the barrier releases
  __syncwarp();
}
// pair build: the slab barriers that count are the leader's (its MMA issuer waits for the warps of both CTAs)
__device__ __forceinline__ void slab_arrive(uint64_t* bar) {
  if constexpr (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
  else mbar_arrive(bar);
}
__device__ __forceinline__ void slab_ready(uint64_t* bar, int lane) {
  slab_fence(lane);
  if (lane == 0) slab_arrive(bar);
}
__device__ __forceinline__ void slabs_ready_all(uint64_t* bar_slab, int lane) {
  slab_fence(lane);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) slab_arrive(&bar_slab[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Lean hot loops (round 2).  The round-1 loops spent ~28 issue slots per element (SASS): a third of them were register
// moves (rotation of accumulators drained into registers through the rolled section loop, lane indices of per-element
// shuffles that fetched the bias) and predicate tests.  Now (together with the single accumulator read in place, see the
// header):
//   * the layer's bias lives in 1 KB of shared memory (staged by the epilogue threads at the start of the step) and is
//     read with broadcast LDS.128, four columns per instruction, instead of one SHFL + MOV per element;
//   * A-operand stores are st.shared.v4 to per-thread precomputed addresses advanced by one add per section;
//   * the section loop stays ROLLED (one 16-column body per variant, ~5 KB, fits the instruction caches): unrolling the
//     four sections was tried while the accumulators were still drained (it removed their rotation) but tripled the hot
//     code to 300 KB, and the measured instruction-fetch stalls (28 % of the samples, profiles/r2c_*) ate the whole gain;
//   * ReLU is folded into the fp32 -> fp16x2 conversions (cvt.rz.relu for hi, cvt.rn.relu for lo);
//   * masking (skip layer: 217 of 256 columns valid) is a template parameter, not per-element tests.
// ---------------------------------------------------------------------------------------------------------
// Shared-space byte addresses of the thread's row in slab 0 of the hi / lo plane, for the two 16-byte chunks (8 K values
// each) the thread owns in every 64-column slab; slab s is s * SLAB_BYTES further (one add per section).
struct ARow { uint32_t hi0, hi1, lo0, lo1; };
__device__ __forceinline__ ARow a_row_addrs(const uint8_t* a_hi, const uint8_t* a_lo, int row, int cq) {
  const uint32_t base = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
  const uint32_t c0 = (uint32_t)(((2 * cq) ^ (row & 7)) << 4), c1 = (uint32_t)(((2 * cq + 1) ^ (row & 7)) << 4);
  ARow r;
  r.hi0 = smem_u32(a_hi) + base + c0; r.hi1 = smem_u32(a_hi) + base + c1;
  r.lo0 = smem_u32(a_lo) + base + c0; r.lo1 = smem_u32(a_lo) + base + c1;
  return r;
}
// (constant offsets added by the callers end up in the instructions' immediate fields once the loops are unrolled)
__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t (&v)[4]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {  // volatile: the buffer is rewritten every step
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// 8 consecutive K values -> fp16 hi / lo planes (one 16-byte chunk each).  RELU: max(x, 0) is folded into the conversions:
// hi = rz(x) clamped at 0 (truncation keeps lo = x - hi >= 0 for x >= 0, so the clamp of lo only acts on x < 0).
template <bool RELU>
__device__ __forceinline__ void split_store8(const float (&x)[8], uint32_t a_hi, uint32_t a_lo) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (RELU) asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[i]));
    const float r0 = x[2 * i] - hf.x, r1 = x[2 * i + 1] - hf.y;
    if (RELU) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo[i]) : "f"(r1), "f"(r0));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo[i]) : "f"(r1), "f"(r0));
  }
  sts128(a_hi, hi);
  sts128(a_lo, lo);
}

// Rarely needed per-column constants (narrow-layer weight rows, rank-update rows: 5 of the 25 steps of a render tile): every
// lane holds those of two of the warp's 64 columns (lane l: columns 64 (l / 8) + 16 g + 2 (l % 8) + {0, 1}), read with a
// shuffle whose lane index is an immediate in the unrolled code.
template <int NROW, int NSMALL>
struct StepConsts {
  float2 row[NROW > 0 ? NROW : 1];
  float2 small[NSMALL > 0 ? NSMALL : 1];
};
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float col_const(const float2& c, int sec, int i) {  // constant of column i (0..15) of section sec
  return __shfl_sync(0xffffffffu, (i & 1) ? c.y : c.x, sec * 8 + (i >> 1));
}
__device__ __forceinline__ void st8(float* dst, const float (&x)[8]) {
  *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
}

// 16 columns [64 sec + 16 cq, +16) of a hidden layer: activation(acc * inv_scale + bias) -> fp16 hi/lo A operand; optional
// softplus' store; optional fp32 dot products with up to NROW narrow-layer weight rows.  Everything that depends on the
// section arrives pre-offset: bias_a (shared address of the 16 biases), ar (A-operand rows of this slab), dmp_h / dmp_d (training dumps, DUMP instantiations only;
// nullable: this point's rows of the next layer's input and of softplus' at column 64 sec + 16 cq).
template <int ACT, bool SAVE_D, int NROW, int NSMALL, bool MASKED, bool DUMP>
__device__ __forceinline__ void hidden_sec(const TcStep& S, const float (&v)[16], const StepConsts<NROW, NSMALL>& K, int sec, int n0, uint32_t bias_a,
                                           uint32_t cst_a,
                                           const ARow& ar, uint32_t (&dpk)[8], float (&dot)[3], const float (&sv)[6], float* dmp_h,
                                           float* dmp_d) {
  // softplus(beta=100) in base 2: t = 100*log2(e)*a ; sp = log2(1 + 2^t) * ln2/100 ; linear above the threshold
  // (softplus(x) >= x, and with t clamped at 20*log2(e) the formula stays below x beyond it, so h = max(sp, a)).
  // (SP_K1 = 100 log2(e) is folded into the staged bias and into the accumulator scale of softplus steps: the epilogue
  //  forms t = SP_K1 * a directly, one multiply per element less; h = max(log2(1 + 2^t), t) * SP_K2, SP_K1 * SP_K2 = 1)
  constexpr float TMAX = 28.853900817779268f;  // 20 * log2(e)
  constexpr bool RELU_IN_CVT = (ACT == TACT_RELU) && NROW == 0 && !DUMP && !MASKED;
  static_assert(!(MASKED && NROW > 0), "narrow rows are folded into unmasked steps only");
  const float inv = (ACT == TACT_SOFTPLUS) ? S.inv_scale * SP_K1 : S.inv_scale;
  const float osc = S.out_scale;
  const int n_valid = S.n_valid;
#pragma unroll
  for (int g8 = 0; g8 < 2; ++g8) {
    const float4 b0 = lds128(bias_a + g8 * 32), b1 = lds128(bias_a + g8 * 32 + 16);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
    float dv[8];
    float pre8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pre8[j] = fmaf(v[g8 * 8 + j], inv, bb[j]);
    if (NSMALL > 0) {  // few-input block ([pts | normal] or the re-injected colour) as an fp32 rank-NSMALL update
#pragma unroll
      for (int q = 0; q < NSMALL; ++q) {
        if constexpr (kSmemConsts) {  // rows [NROW, NROW + NSMALL) of the staged constants, these 8 columns
          const float4 c0 = lds128(cst_a + (NROW + q) * 1024 + g8 * 32), c1 = lds128(cst_a + (NROW + q) * 1024 + g8 * 32 + 16);
          const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) pre8[j] = fmaf(sv[q], cc[j], pre8[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) pre8[j] = fmaf(sv[q], col_const(K.small[q], sec, g8 * 8 + j), pre8[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = g8 * 8 + j;
      const float pre = pre8[j];
      float h;
      if (ACT == TACT_SOFTPLUS) {  // pre = t = 100 log2(e) * a here
        const float e = ex2_ftz(fminf(pre, TMAX));
        const float ope = 1.0f + e;
        h = fmaxf(lg2_ftz(ope), pre) * SP_K2;
        if (SAVE_D) dv[j] = e * rcp_ftz(ope);  // sigmoid(100 a); -> 1 - 2e-9 in the linear region
      } else {
        h = RELU_IN_CVT ? pre : fmaxf(pre, 0.0f);
      }
      if (MASKED) h = (n0 + i < n_valid) ? h : 0.0f;
      if (NROW > 0 && !kSmemConsts) {
#pragma unroll
        for (int jj = 0; jj < NROW; ++jj) dot[jj] = fmaf(h, col_const(K.row[jj], sec, i), dot[jj]);
      }
      o[j] = MASKED ? h * osc : h;
    }
    if (NROW > 0 && kSmemConsts) {  // (o == h here: MASKED steps have no narrow rows)
#pragma unroll
      for (int jj = 0; jj < NROW; ++jj) {
        const float4 c0 = lds128(cst_a + jj * 1024 + g8 * 32), c1 = lds128(cst_a + jj * 1024 + g8 * 32 + 16);
        const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) dot[jj] = fmaf(o[j], cc[j], dot[jj]);
      }
    }
    if (SAVE_D) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dpk[g8 * 4 + j] = d_pack(dv[2 * j], dv[2 * j + 1]);  // stored by the caller, after the slab fence
      if (DUMP && dmp_d) st8(dmp_d + g8 * 8, dv);
    }
    if (DUMP && dmp_h) st8(dmp_h + g8 * 8, o);
    split_store8<RELU_IN_CVT>(o, g8 ? ar.hi1 : ar.hi0, g8 ? ar.lo1 : ar.lo0);
  }
}

// One step of a hidden layer: one rolled loop over the four 64-column sections, each read from this step's TMEM accumulator
// when its turn comes (the next layer's MMAs write the other half of TMEM) and announced as soon as its slab of the next
// A operand is in place, so the next layer's K-blocks run under the remaining sections.
template <int ACT, bool SAVE_D, int NROW, int NSMALL, bool MASKED, bool DUMP = false>
__device__ __forceinline__ void epi_hidden(const TcStep& S, const float* __restrict__ packed, uint32_t t_acc, int row, int g, ARow ar,
                                           uint32_t bias_a, uint32_t cst_a, uint32_t* dsave, float (&dot)[3], const float (&sv)[6], bool early,
                                           uint64_t* bar_slab, int lane, EpiProf& ep, float* dmp_h, float* dmp_d) {
  StepConsts<NROW, NSMALL> K;
  if constexpr (!kSmemConsts) {
    const int col = 64 * (lane >> 3) + 16 * g + 2 * (lane & 7);
#pragma unroll
    for (int jj = 0; jj < NROW; ++jj) K.row[jj] = ldg2(packed + S.row_off + jj * 256 + col);
#pragma unroll
    for (int q = 0; q < NSMALL; ++q) K.small[q] = ldg2(packed + S.small_off + q * 256 + col);
  }
  uint32_t* dsave_t = SAVE_D ? dsave + g * 8 * TCM + row : nullptr;
  if (DUMP) { if (dmp_h) dmp_h += 16 * g; if (dmp_d) dmp_d += 16 * g; }
  float w[16];
#pragma unroll 1
  for (int sec = 0; sec < 4; ++sec) {
    uint32_t dpk[8];
    tmem_ld16(t_acc + sec * 64 + g * 16, w);
    hidden_sec<ACT, SAVE_D, NROW, NSMALL, MASKED, DUMP>(S, w, K, sec, sec * 64 + g * 16, bias_a, cst_a, ar, dpk, dot, sv, dmp_h, dmp_d);
    ep.mark(sec < 2 ? 0 : (sec == 2 ? 3 : 5));
    bias_a += 256u;
    cst_a += 256u;
    ar.hi0 += SLAB_BYTES; ar.hi1 += SLAB_BYTES; ar.lo0 += SLAB_BYTES; ar.lo1 += SLAB_BYTES;
    if (DUMP) { if (dmp_h) dmp_h += 64; if (dmp_d) dmp_d += 64; }
    if (early && sec < 3) slab_ready(&bar_slab[sec], lane);
    ep.mark(sec == 0 ? 2 : 4);
    if (SAVE_D) {  // softplus' words of this section (slot + (32 sec + 8 g) TCM + row): global stores issued AFTER the slab's
                   // fence, which would otherwise wait for them to be performed
#pragma unroll
      for (int i = 0; i < 8; ++i) dsave_t[i * TCM] = dpk[i];
      dsave_t += 32 * TCM;
    }
  }
}

// gradient chain: next adjoint = (acc * scale) (.) softplus'(a_{l-1}); encoding part of a skip layer -> scratch.
// dw: the 8 softplus' words (16 values, 16-bit fixed point) of this section; MASKED: the skip layer's steps (n_valid < 256);
// n0 = 64 sec + 16 cq; ar, dmp_gh / dmp_ga pre-offset to this section.
template <bool MASKED, bool DUMP>
__device__ __forceinline__ void bwd_sec(const TcStep& S, const TcProgram& prog, const float (&v)[16], const uint32_t (&dw)[8], int n0, int row,
                                        const ARow& ar, float* gxs, float sc, float sco, bool skip, float* dmp_gh, float* dmp_ga) {
  const int n_valid = S.n_valid;
#pragma unroll
  for (int g8 = 0; g8 < 2; ++g8) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // 16-bit fixed point -> float without a conversion instruction: 0x4B000000 | q is the float 2^23 + q
      const uint32_t w = dw[g8 * 4 + (j >> 1)];
      const float q = __uint_as_float(__byte_perm(w, 0x4B000000u, (j & 1) ? 0x7632 : 0x7610)) - 8388608.0f;
      const float t = (v[g8 * 8 + j] * sco) * q;
      if (MASKED) {
        const int k = n0 + g8 * 8 + j;
        o[j] = (k < n_valid) ? t : 0.0f;
        if (skip && k >= n_valid && k < n_valid + prog.pe_dim) gxs[(k - n_valid) * TCM + row] = v[g8 * 8 + j] * sc;
      } else {
        o[j] = t;
      }
    }
    if (DUMP && dmp_gh) {  // training dumps: the adjoint before / after the multiplication by softplus' (un-scaled)
      float gh[8], ga[8];
      const float un = 1.0f / S.out_scale;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        gh[j] = (!MASKED || n0 + g8 * 8 + j < n_valid) ? v[g8 * 8 + j] * sc : 0.0f;
        ga[j] = o[j] * un;
      }
      if (dmp_gh != dmp_ga) st8(dmp_gh + g8 * 8, gh);   // optional
      st8(dmp_ga + g8 * 8, ga);
    }
    split_store8<false>(o, g8 ? ar.hi1 : ar.hi0, g8 ? ar.lo1 : ar.lo0);
  }
}

// `cur` / `nxt` hold the softplus' words of sections 0 and 1 (loaded by the caller before it waited for the accumulators); the
// words of section s + 2 are requested before section s is computed (the scratch is evicted to DRAM between its write and
// this read: one section of lead time did not cover that latency, profiles/r2j).
template <bool MASKED, bool DUMP>
__device__ __forceinline__ void epi_bwd(const TcStep& S, const TcProgram& prog, uint32_t t_acc, int row, int g, ARow ar,
                                        const uint32_t* D, float* gxs, bool early, uint64_t* bar_slab, int lane, EpiProf& ep,
                                        uint32_t (&cur)[8], uint32_t (&nxt)[8], float* dmp_gh, float* dmp_ga) {
  const bool skip = MASKED && (S.flags & TF_SKIP_BWD) != 0;
  const float sc = (S.flags & TF_SKIP_BWD) ? S.inv_scale * 0.70710678118654752440f : S.inv_scale;
  const float sco = sc * S.out_scale * (1.0f / 65535.0f);
  const uint32_t* D_t = D + g * 8 * TCM + row;
  if (DUMP) { if (dmp_gh) dmp_gh += 16 * g; if (dmp_ga) dmp_ga += 16 * g; }
  float w[16];
#pragma unroll 1
  for (int sec = 0; sec < 4; ++sec) {
    uint32_t nxt2[8];
    D_t += 32 * TCM;
    if (sec < 2) {  // softplus' words of section sec + 2: in flight during this section and the next
#pragma unroll
      for (int i = 0; i < 8; ++i) nxt2[i] = D_t[(32 + i) * TCM];
    }
    tmem_ld16(t_acc + sec * 64 + g * 16, w);
    bwd_sec<MASKED, DUMP>(S, prog, w, cur, sec * 64 + g * 16, row, ar, gxs, sc, sco, skip, dmp_gh, dmp_ga);
    ep.mark(sec < 2 ? 0 : (sec == 2 ? 3 : 5));
    ar.hi0 += SLAB_BYTES; ar.hi1 += SLAB_BYTES; ar.lo0 += SLAB_BYTES; ar.lo1 += SLAB_BYTES;
    if (DUMP) { if (dmp_gh) dmp_gh += 64; if (dmp_ga) dmp_ga += 64; }
    if (early && sec < 3) slab_ready(&bar_slab[sec], lane);
    ep.mark(sec == 0 ? 2 : 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) { cur[i] = nxt[i]; nxt[i] = nxt2[i]; }
  }
}

// Tangent pass (training instantiation only): u = W_l t_l (accumulators hold u * ts * W_SCALE, ts = the launch's power-of-two
// scale of the tangents), e = softplus''(a_l) (.) u (.) gh_l -> El row, t_{l+1} = softplus'(a_l) (.) u (* 1/sqrt(2) into the
// skip layer) -> next A operand (scaled by ts) and T row (un-scaled).  softplus' and gh come from the recompute launch's dumps.
__device__ __forceinline__ void epi_tan(const TcStep& S, uint32_t t_acc, int row, int g, uint8_t* a_hi, uint8_t* a_lo,
                                        const float* __restrict__ drow, const float* __restrict__ ghrow, float* erow, float* trow,
                                        float inv_ts, bool early, uint64_t* bar_slab, int lane) {
  const float sc = S.inv_scale, osc = S.out_scale;
  const int n_valid = S.n_valid;
  float w[16];
#pragma unroll 1
  for (int sec = 0; sec < 4; ++sec) {
    const int n0 = sec * 64 + g * 16;
    tmem_ld16(t_acc + n0, w);
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
      const int nb = n0 + g8 * 8;
      float dv[8], gv[8];
      if (drow) {
        const float4 d0 = *reinterpret_cast<const float4*>(drow + nb), d1 = *reinterpret_cast<const float4*>(drow + nb + 4);
        const float4 g0 = *reinterpret_cast<const float4*>(ghrow + nb), g1 = *reinterpret_cast<const float4*>(ghrow + nb + 4);
        dv[0] = d0.x; dv[1] = d0.y; dv[2] = d0.z; dv[3] = d0.w; dv[4] = d1.x; dv[5] = d1.y; dv[6] = d1.z; dv[7] = d1.w;
        gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w; gv[4] = g1.x; gv[5] = g1.y; gv[6] = g1.z; gv[7] = g1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { dv[j] = 0.f; gv[j] = 0.f; }
      }
      float o[8], ev[8], tv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = nb + j < n_valid;
        const float us = w[g8 * 8 + j] * sc;          // u * ts
        const float u = us * inv_ts;
        const float d = ok ? dv[j] : 0.0f;
        ev[j] = ok ? 100.0f * (1.0f - d) * u * gv[j] : 0.0f;   // gv = ga = gh d: softplus'' u gh = 100 (1 - d) u ga
        o[j] = us * d * osc;
        tv[j] = o[j] * inv_ts;
      }
      if (erow) { st8(erow + nb, ev); st8(trow + nb, tv); }
      write_a8(a_hi, a_lo, nb >> 6, row, (nb & 63) >> 3, o);
    }
    if (early && sec < 3) slab_ready(&bar_slab[sec], lane);
  }
}

// DUMP: the training instantiation (TrainDump stores compiled in); the inference instantiation carries none of it
template <bool DUMP>
__global__ void __launch_bounds__(TC_KERNEL_THREADS, 1) shade_tc_kernel(const __grid_constant__ TcProgram prog,
                                                                        const float* __restrict__ packed,
                                                                        const __grid_constant__ ShadeArgs a,
                                                                        float* __restrict__ gxscratch) {
  // 1 KB-aligned by declaration (the swizzled slabs need it): the base is then a link-time constant, where rounding it up at run
  // time was re-materialised all over the epilogue (4 % of the kernel's instructions, profiles/r2o); checked once below
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + A_SLABS * SLAB_BYTES;
  uint8_t* wring = smem + 2 * A_SLABS * SLAB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wring + TC_STAGES * STAGE_BYTES);
  // ring slot s: hi slab / lo slab addresses (slot 2 = the small-input slab of the two A planes)
  auto ring_hi = [&](int s_) -> uint8_t* { return s_ < 2 ? wring + s_ * STAGE_BYTES : a_hi + SMALL_SLAB * SLAB_BYTES; };
  auto ring_lo = [&](int s_) -> uint8_t* { return s_ < 2 ? wring + s_ * STAGE_BYTES + SLAB_BYTES : a_lo + SMALL_SLAB * SLAB_BYTES; };
  // pair build: this CTA stages half of every slab, so a slot is 16 KB ([first half slab | second half slab]): four in the
  // ring, and the small-input slabs of the two A planes are slots 4 and 5 when the program does not need them
  auto pslot = [&](int s_) -> uint8_t* {
    return s_ < TC_PAIR_RING_SLOTS ? wring + s_ * SLAB_BYTES
                                   : (s_ == TC_PAIR_RING_SLOTS ? a_hi + SMALL_SLAB * SLAB_BYTES : a_lo + SMALL_SLAB * SLAB_BYTES);
  };
  const int n_stages = kPair ? TC_PAIR_RING_SLOTS + (prog.n_stages == 3 ? 2 : 0) : prog.n_stages;
  // kSmemConsts: the step's narrow-layer / rank-update rows ([row_n + n_small][256] fp32), in the ring's fourth 16 KB
  float* cst_s = reinterpret_cast<float*>(wring + 3 * SLAB_BYTES);
  uint64_t* bar_full = bars;        // [6]
  uint64_t* bar_empty = bars + 6;   // [6]
  uint64_t* bar_acc = bars + 12;    // accumulators complete (MMA -> epilogue)
  uint64_t* bar_slab = bars + 13;   // [4] A-operand slab ready (epilogue warps -> MMA)
  uint64_t* bar_pfull = bars + 17;  // [6] pair build, leader: the peer's half of a weight stage has landed (relayed by the peer)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [256]: the current step's bias (epilogue)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;  // pair build: 0 = leader (issues the MMAs), 1 = peer
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); mbar_init(&bar_pfull[i], 1); }
    mbar_init(bar_acc, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_slab[i], kPair ? 2 * TC_EPI_WARPS : TC_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS) {
    if constexpr (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t n_tiles = (a.P + TCM - 1) / TCM;
  // tile iteration: CTA b takes tiles b, b + grid, ...; pair build: pair q = b / 2 takes the tile pairs q, q + grid / 2, ... and
  // the CTA of rank r the r-th tile of each (both CTAs walk the same number of steps; a tile beyond the end has no valid row)
  const int64_t it_first = kPair ? (blockIdx.x >> 1) : blockIdx.x;
  const int64_t it_stride = kPair ? (gridDim.x >> 1) : gridDim.x;
  const int64_t it_end = kPair ? (n_tiles + 1) / 2 : n_tiles;
  const uint8_t* packed_b = reinterpret_cast<const uint8_t*>(packed);
  // register hand-over between the warpgroups (the kernel launches with 96 per thread: five warps per sub-partition);
  // each role's branch starts with its setmaxnreg so that the allocator sees the budget of that branch

  if (warp >= TC_EPI_WARPS) {
  // one setmaxnreg for the whole warpgroup (.aligned: every warp of the group executes this very instruction)
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_OTHER));
  if (warp == TC_EPI_WARPS) {
    // ================================================================ weight producer (bulk async copies)
    if (lane == 0) {
      uint32_t stg = 0, ph = 1;  // ring slot and the parity its "empty" barrier is waited with (fresh barrier: passes)
      const bool prof = kTcProf && prog.prof && blockIdx.x == 0;
      long long t_wait = 0, t_begin = clock64();
      for (int64_t it = it_first; it < it_end; it += it_stride) {
        for (int s = 0; s < prog.n_steps; ++s) {
          const TcStep& S = prog.s[s];
          const uint8_t* src = packed_b + S.w_off;
          // pair build: this CTA's rows of every [N outputs][32 inputs] slab = N / 2 rows of 64 B from row (N / 2) rank on
          const uint32_t roff = kPair ? crank * (S.n_halves == 2 ? 8192u : 4096u) : 0u;
          // pass 0 (correction products): hi + lo plane of every 32-wide half-block, one stage each
          for (int kb = 0; kb < S.n_kb; ++kb) {
            for (int sh = 0; sh < 2; ++sh) {  // two half-block stages per K-block; empty ones are skipped
              if (S.ksteps[kb] <= 2 * sh) continue;
              const long long t0 = prof ? clock64() : 0;
              mbar_wait(&bar_empty[stg], ph);
              if (prof) t_wait += clock64() - t0;
              const uint8_t* img = src + (size_t)(kb * 2 + sh) * STAGE_BYTES;
              if constexpr (kPair) {
                mbar_expect_tx(&bar_full[stg], SLAB_BYTES);
                bulk_g2s(pslot(stg), img + roff, SLAB_BYTES / 2, &bar_full[stg]);
                bulk_g2s(pslot(stg) + SLAB_BYTES / 2, img + SLAB_BYTES + roff, SLAB_BYTES / 2, &bar_full[stg]);
              } else {
                mbar_expect_tx(&bar_full[stg], STAGE_BYTES);
                bulk_g2s(ring_hi(stg), img, SLAB_BYTES, &bar_full[stg]);
                bulk_g2s(ring_lo(stg), img + SLAB_BYTES, SLAB_BYTES, &bar_full[stg]);
              }
              if (++stg == (uint32_t)n_stages) { stg = 0; ph ^= 1u; }
            }
          }
          // pass 1 (main products): the hi planes again, both half-blocks of a K-block in one stage
          for (int kb = 0; kb < S.n_kb; ++kb) {
            const int nks = S.ksteps[kb];
            if (nks <= 0) continue;
            const long long t0 = prof ? clock64() : 0;
            mbar_wait(&bar_empty[stg], ph);
            if (prof) t_wait += clock64() - t0;
            const uint8_t* img = src + (size_t)(kb * 2) * STAGE_BYTES;
            if constexpr (kPair) {
              mbar_expect_tx(&bar_full[stg], nks > 2 ? SLAB_BYTES : SLAB_BYTES / 2);
              bulk_g2s(pslot(stg), img + roff, SLAB_BYTES / 2, &bar_full[stg]);
              if (nks > 2) bulk_g2s(pslot(stg) + SLAB_BYTES / 2, img + STAGE_BYTES + roff, SLAB_BYTES / 2, &bar_full[stg]);
            } else {
              mbar_expect_tx(&bar_full[stg], nks > 2 ? 2 * SLAB_BYTES : SLAB_BYTES);
              bulk_g2s(ring_hi(stg), img, SLAB_BYTES, &bar_full[stg]);
              if (nks > 2) bulk_g2s(ring_lo(stg), img + STAGE_BYTES, SLAB_BYTES, &bar_full[stg]);
            }
            if (++stg == (uint32_t)n_stages) { stg = 0; ph ^= 1u; }
          }
        }
      }
      if (prof) { g_tc_prof[6] += (unsigned long long)t_wait; g_tc_prof[7] += (unsigned long long)(clock64() - t_begin); }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ================================================================ MMA issuer: the whole warp walks the program
    // (warp-uniform control flow), one elected lane issues.  The loop is kept lean -- this warp shares its scheduler
    // with four busy epilogue warps, so every instruction here delays the tensor core: descriptors are a precomputed
    // low word plus an offset, ring slot / phase are counters (no division).
    // f16 x f16 -> f32, M=128 (pair build: M=256 over the two CTAs), N=256 (N=128 for layers whose image has <= 128 valid rows)
    constexpr uint32_t MMA_M = kPair ? 2 * TCM : TCM;
    const uint32_t idesc256 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(MMA_M >> 4) << 24);
    const uint32_t idesc128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(MMA_M >> 4) << 24);
    constexpr uint32_t HI_SW128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // descriptor bits [32,64)
    constexpr uint32_t HI_SW64 = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);
    const uint32_t a_hi_lo32 = ((smem_u32(a_hi) >> 4) & 0x3FFFu) | 0x10000u;  // descriptor bits [0,32) of slab 0, k-step 0
    const uint32_t a_lo_lo32 = ((smem_u32(a_lo) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t bh0 = ((smem_u32(ring_hi(0)) >> 4) & 0x3FFFu) | 0x10000u, bl0 = ((smem_u32(ring_lo(0)) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t bh1 = ((smem_u32(ring_hi(1)) >> 4) & 0x3FFFu) | 0x10000u, bl1 = ((smem_u32(ring_lo(1)) >> 4) & 0x3FFFu) | 0x10000u;
    const uint32_t bh2 = ((smem_u32(ring_hi(2)) >> 4) & 0x3FFFu) | 0x10000u, bl2 = ((smem_u32(ring_lo(2)) >> 4) & 0x3FFFu) | 0x10000u;
    // descriptor low word of the first / second 8 KB of weight slot `g` (pair build: six 16 KB slots, see pslot)
    auto slot_b0 = [&](uint32_t g) -> uint32_t {
      if constexpr (kPair) return g < (uint32_t)TC_PAIR_RING_SLOTS ? bh0 + g * (uint32_t)(SLAB_BYTES >> 4) : (g == (uint32_t)TC_PAIR_RING_SLOTS ? bh2 : bl2);
      else return g == 0 ? bh0 : (g == 1 ? bh1 : bh2);
    };
    auto slot_b1 = [&](uint32_t g) -> uint32_t {
      if constexpr (kPair) return slot_b0(g) + (uint32_t)(SLAB_BYTES >> 5);
      else return g == 0 ? bl0 : (g == 1 ? bl1 : bl2);
    };
    auto desc = [](uint32_t hi, uint32_t lo) -> uint64_t { return ((uint64_t)hi << 32) | lo; };
    auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
      if constexpr (kPair) mma_f16_pair(d, da, db, id, acc); else mma_f16(d, da, db, id, acc);
    };
    auto commit = [&](uint64_t* bar) { if constexpr (kPair) mma_commit_pair(bar); else mma_commit(bar); };
    const bool leader = elect_one();
    uint32_t stg = 0, ph = 0, step_par = 0, step_count = 0;
    if (kPair && crank != 0) {
      // ---- peer CTA of a pair: no MMAs to issue.  This warp relays "my half of weight stage g has landed" to the leader's
      // issuer: it walks the stages in the producer's order, waits for its own full barrier and arrives on the leader's
      // bar_pfull[g] (the commit's multicast then frees the slot in both CTAs).
      const uint32_t pfull0 = mapa_u32(smem_u32(bar_pfull), 0);
      for (int64_t it = it_first; it < it_end; it += it_stride) {
        for (int s = 0; s < prog.n_steps; ++s) {
          const TcStep& S = prog.s[s];
          int n_st = 0;  // stages of this step: two (or one) per K-block in pass 0, one per K-block in pass 1
          for (int kb = 0; kb < S.n_kb; ++kb) n_st += S.ksteps[kb] <= 0 ? 0 : (S.ksteps[kb] > 2 ? 3 : 2);
          for (int i = 0; i < n_st; ++i) {
            mbar_wait(&bar_full[stg], ph);
            if (leader) mbar_arrive_cluster(pfull0 + stg * 8u);
            __syncwarp();
            if (++stg == (uint32_t)n_stages) { stg = 0; ph ^= 1u; }
          }
        }
      }
    } else {
    const bool prof = kTcProf && prog.prof && blockIdx.x == 0 && lane == 0;
    long long t_wa = 0, t_wf = 0, t_begin = clock64();
    long long t_ws[4] = {0, 0, 0, 0};
    auto wait_slab = [&](int sb, uint32_t par) { mbar_wait(&bar_slab[sb], par); };
    // Weight stages are normally resident long before their turn: every stage's full barrier is probed one stage ahead
    // (`w_ready`, the probe's latency overlaps the issue of the current stage's MMAs), the blocking wait is the fallback.
    bool w_ready = false;
    auto wait_full = [&](uint32_t g, uint32_t par) {
      if (!w_ready) {
        mbar_wait(&bar_full[g], par);
        if constexpr (kPair) mbar_wait(&bar_pfull[g], par);
      }
    };
    auto probe_next = [&](uint32_t g, uint32_t par) {  // the stage after (g, par)
      uint32_t ng = g + 1, np_ = par;
      if (ng == (uint32_t)n_stages) { ng = 0; np_ ^= 1u; }
      w_ready = mbar_test(&bar_full[ng], np_);
      if constexpr (kPair) w_ready = mbar_test(&bar_pfull[ng], np_) && w_ready;
    };
    for (int64_t it = it_first; it < it_end; it += it_stride) {
      for (int s = 0; s < prog.n_steps; ++s, ++step_count, step_par ^= 1u) {
        const TcStep& S = prog.s[s];
        long long t0;
        const uint32_t idesc = S.n_halves == 2 ? idesc256 : idesc128;
        uint32_t accum = 0;
        uint32_t waited = 0;
        const uint32_t acc_t = tmem + ((step_count & 1u) << 8);  // consecutive steps alternate between the two halves of TMEM
        // ---- pass 0: the correction products of the whole layer (lo * hi + hi * lo), K-block by K-block as the epilogue
        // of the previous step announces the slabs of the A operand
        for (int kb = 0; kb < S.n_kb; ++kb) {
          t0 = prof ? clock64() : 0;
          const int sb = S.slab[kb] < 4 ? S.slab[kb] : 3;  // the small-input slab is staged last, together with slab 3
          wait_slab(sb, step_par);
          waited |= 1u << sb;
          if (prof) { const long long dt = clock64() - t0; t_wa += dt; t_ws[sb] += dt; }
          tc_fence_after();
          const uint32_t a_off = (uint32_t)S.slab[kb] * (SLAB_BYTES >> 4);
          const int nks = S.ksteps[kb];
          for (int sh = 0; sh < 2; ++sh) {
            const int nk = nks - 2 * sh;  // k-steps in this half-block stage
            if (nk <= 0) continue;
            t0 = prof ? clock64() : 0;
            wait_full(stg, ph);
            if (prof) t_wf += clock64() - t0;
            tc_fence_after();
            probe_next(stg, ph);
            const uint32_t bh = slot_b0(stg), bl = slot_b1(stg);
            if (leader) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                if (k < nk) {
                  const uint32_t ka = a_off + (uint32_t)(sh * 2 + k) * 2u;  // 32 bytes per k-step, in 16-byte units
                  const uint64_t dAh = desc(HI_SW128, a_hi_lo32 + ka), dAl = desc(HI_SW128, a_lo_lo32 + ka);
                  const uint64_t dBh = desc(HI_SW64, bh + (uint32_t)k * 2u), dBl = desc(HI_SW64, bl + (uint32_t)k * 2u);
                  mma(acc_t, dAl, dBh, idesc, accum);  // lo * hi
                  mma(acc_t, dAh, dBl, idesc, 1u);     // hi * lo
                  accum = 1u;
                }
              }
              commit(&bar_empty[stg]);  // frees the ring slot when these MMAs retire
            }
            __syncwarp();
            if (++stg == (uint32_t)n_stages) { stg = 0; ph ^= 1u; }
          }
        }
        // ---- pass 1: the main products (hi * hi) on top; one stage = the hi planes of both half-blocks of a K-block
        for (int kb = 0; kb < S.n_kb; ++kb) {
          const int nks = S.ksteps[kb];
          if (nks <= 0) continue;
          const uint32_t a_off = (uint32_t)S.slab[kb] * (SLAB_BYTES >> 4);
          t0 = prof ? clock64() : 0;
          wait_full(stg, ph);
          if (prof) t_wf += clock64() - t0;
          tc_fence_after();
          probe_next(stg, ph);
          const uint32_t b0 = slot_b0(stg), b1 = slot_b1(stg);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < nks) {
                const uint64_t dAh = desc(HI_SW128, a_hi_lo32 + a_off + (uint32_t)k * 2u);
                const uint64_t dBh = desc(HI_SW64, (k < 2 ? b0 : b1) + (uint32_t)(k & 1) * 2u);
                mma(acc_t, dAh, dBh, idesc, 1u);
              }
            }
            commit(&bar_empty[stg]);
          }
          __syncwarp();
          if (++stg == (uint32_t)n_stages) { stg = 0; ph ^= 1u; }
        }
        // every slab barrier completes exactly one phase per step; wait for the ones no K-block of this step used as well
        // (already complete or about to be: the epilogue announces them together with the used ones), so that no barrier can
        // run a phase ahead of its consumer (compute-sanitizer synccheck: "missing wait")
#pragma unroll
        for (int sb = 0; sb < 4; ++sb)
          if (!((waited >> sb) & 1u)) wait_slab(sb, step_par);
        if (leader) commit(bar_acc);
        __syncwarp();
      }
    }
    if (prof) {
      g_tc_prof[0] += (unsigned long long)t_wa; g_tc_prof[1] += (unsigned long long)t_wf;
      g_tc_prof[2] += (unsigned long long)(clock64() - t_begin); g_tc_prof[3] += step_count;
      for (int i = 0; i < 4; ++i) g_tc_prof[8 + i] += (unsigned long long)t_ws[i];
    }
    }  // leader / single-CTA issuer
  }  // warps 18, 19 only complete the warpgroup
  } else {
    // ================================================================ epilogue: 4 threads per point (column quarters)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_EPI));
    const int cq = warp >> 2;  // column group: columns [16 cq, 16 cq + 16) of every 64-column slab
    const int row = (warp & 3) * 32 + lane;  // == TMEM lane
    const uint32_t t_acc0 = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const ARow ar = a_row_addrs(a_hi, a_lo, row, cq);
    const uint32_t bias_a = smem_u32(bias_s) + (uint32_t)cq * 64u;  // bias of this thread's 16 columns of section 0
    const uint32_t cst_a = smem_u32(cst_s) + (uint32_t)cq * 64u;    // staged constants, row 0, same columns
    float* dscr = a.dscratch ? a.dscratch + (size_t)blockIdx.x * (prog.n_hidden + 1) * 256 * TCM : nullptr;  // +1: feature slot
    float* gxs = gxscratch + (size_t)blockIdx.x * TC_GXS_ROWS * TCM;
    float* pes = gxs + TC_GXS_PE * TCM;  // [pe_dim][TCM]: the tile's encoding, computed once (first layer's input) and re-read
                                         // by the skip layer and by the encoding's adjoint
    uint32_t acc_count = 0;
    const bool prof = kTcProf && prog.prof && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 12 * 32);
    long long t_wacc = 0;
    const long long t_begin = clock64();
    EpiProf ep;
#ifdef CNEUS_TC_EPI_PROF
    ep.on = prog.prof && prog.prof < 100 && blockIdx.x == 0 && threadIdx.x == (prog.prof - 1) * 32;   // cneus_tc_prof_enable(1 + warp)
    ep.t = 0;
#endif

    float tan_ts = 1.0f, tan_inv_ts = 1.0f;  // tangent pass: power-of-two scale of the A operand (largest seed magnitude -> 2^6)
    if (DUMP && prog.tangent) {
      const float amax = *a.tan_amax;
      if (amax > 0.0f && isfinite(amax)) {
        int e = 6 - ilogbf(amax);
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
        tan_ts = ldexpf(1.0f, e);
        tan_inv_ts = ldexpf(1.0f, -e);
      }
    }
    for (int64_t it = it_first; it < it_end; it += it_stride) {
      const int64_t tile = kPair ? 2 * it + (int64_t)crank : it;
      const int64_t p = tile * TCM + row;
      const bool valid = p < a.P;
      const bool writer = valid && cq == 0;
      RowState st;
#pragma unroll
      for (int c = 0; c < 3; ++c) { st.pt[c] = 0.f; st.dir[c] = 0.f; st.nrm[c] = 0.f; st.cg[c] = 0.f; }
      st.sdf = 0.f;
      if (valid) {
        if (a.src_mode == 0) {
          st.pt[0] = a.pts[p * 3]; st.pt[1] = a.pts[p * 3 + 1]; st.pt[2] = a.pts[p * 3 + 2];
        } else if (a.src_mode == 1) {
          const int64_t r = p / a.n_per_ray;
          const float t = a.t[p];
#pragma unroll
          for (int c = 0; c < 3; ++c) { st.dir[c] = a.rays_d[r * 3 + c]; st.pt[c] = ray_point(a.rays_o[r * 3 + c], st.dir[c], t); }
        } else {
          const int64_t lin = a.lin_begin + p;
          const int iz = (int)(lin % a.res), iy = (int)((lin / a.res) % a.res), ix = (int)(lin / ((int64_t)a.res * a.res));
          st.pt[0] = a.gx[ix]; st.pt[1] = a.gy[iy]; st.pt[2] = a.gz[iz];
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) st.xs[c] = st.pt[c] * prog.sdf_scale;
      // ---- A operand of the first layer: positional encoding of the scaled point (PositionEncoding.py:51-76); each of
      // the row's four threads computes a quarter of it (one sincosf per (frequency, dim)) and publishes it for the others
      {
        float v[16];
        if (DUMP && prog.tangent) {  // tangent pass: the seed t_0 takes the encoding's place (also in the skip feed below)
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (valid && 16 * cq + j < prog.pe_dim) ? a.tan_t0[p * prog.pe_dim + 16 * cq + j] : 0.0f;
        } else {
          small_block16(prog, st, SMALL_PE, 0, cq, v);
        }
        epi_bar_sync();  // the previous tile's readers are done
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (16 * cq + j < prog.pe_dim) pes[(16 * cq + j) * TCM + row] = v[j];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (DUMP && prog.tangent) ? v[c * 8 + j] * tan_ts : v[c * 8 + j];
          write_a8(a_hi, a_lo, 0, row, cq * 2 + c, o);
        }
        __threadfence_block();
        epi_bar_sync();
      }
      slabs_ready_all(bar_slab, lane);
      ep.start();

      for (int s = 0; s < prog.n_steps; ++s, ++acc_count) {
        const TcStep& S = prog.s[s];
        // section-0 constants (biases, or softplus' words of the gradient chain) are requested before the wait
        float bias_v = 0.f;
        uint32_t pre8[8], pre8b[8];
        const bool stage_bias = S.epi == EPI_HIDDEN || S.epi == EPI_PARK;
        if (stage_bias) {
          // The step's bias -> shared memory (single 1 KB buffer), staged HERE, under the tail of this step's MMAs (the warp has
          // nothing else to do until the accumulators are complete): the first named barrier orders every warp's last read of
          // the previous bias before the overwrite, the second publishes the new values to the 16 warps.
          const int boff = S.epi == EPI_PARK ? prog.feat_bias_off : S.bias_off;
          if (threadIdx.x < 256 && boff >= 0) bias_v = __ldg(packed + boff + threadIdx.x);
          if (S.epi == EPI_HIDDEN && S.act == TACT_SOFTPLUS) bias_v *= SP_K1;  // softplus steps work on t = 100 log2(e) * a
          // ... and, same protocol, the narrow-layer rows and the rank-update rows this step folds in (0 to 6 rows of 256)
          float cst_v[3] = {0.f, 0.f, 0.f};
          const int n_cst = (kSmemConsts && S.epi == EPI_HIDDEN)
                                ? (((S.row_off >= 0 && !(S.act == TACT_RELU && S.n_small == 6)) ? (int)S.row_n : 0) + (int)S.n_small) * 256 : 0;
          if (kSmemConsts && n_cst > 0) {
            const int n_row = (S.row_off >= 0 && !(S.act == TACT_RELU && S.n_small == 6)) ? (int)S.row_n * 256 : 0;  // as dispatched below
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const int e = (int)threadIdx.x + i * TC_EPI_THREADS;
              if (e < n_cst) cst_v[i] = __ldg(packed + (e < n_row ? S.row_off + e : S.small_off + (e - n_row)));
            }
          }
          ep.mark(12);
          epi_bar_sync();
          ep.mark(1);   // waiting for the slowest epilogue warp
          if (threadIdx.x < 256) asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(bias_s) + threadIdx.x * 4u), "f"(bias_v) : "memory");
          if (kSmemConsts && n_cst > 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const int e = (int)threadIdx.x + i * TC_EPI_THREADS;
              if (e < n_cst) cst_s[e] = cst_v[i];
            }
          }
          epi_bar_sync();
          ep.mark(12);
        } else if (S.epi == EPI_BWD) {
          const uint32_t* D0 = reinterpret_cast<const uint32_t*>(dscr + (size_t)S.d_layer * 256 * TCM);
#pragma unroll
          for (int i = 0; i < 8; ++i) { pre8[i] = D0[(cq * 8 + i) * TCM + row]; pre8b[i] = D0[(32 + cq * 8 + i) * TCM + row]; }
        }
        const long long t0 = prof ? clock64() : 0;
        mbar_wait(bar_acc, acc_count & 1);
        if (prof) t_wacc += clock64() - t0;
        const uint32_t t_acc = t_acc0 + ((acc_count & 1u) << 8);  // this step's half of TMEM (the MMA issuer alternates the same way)
        tc_fence_after();
        ep.start();
#ifdef CNEUS_TC_EPI_PROF
        const long long t_step0 = ep.on ? clock64() : 0;
#endif
        float dot[3] = {0.f, 0.f, 0.f};
        float sv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        // training dumps of this step (rows of this thread's point)
        float* dmp0 = nullptr;
        float* dmp1 = nullptr;
        if (DUMP && valid && S.d_layer >= 0) {
          if (S.epi == EPI_HIDDEN && S.act == TACT_SOFTPLUS) {
            dmp0 = a.dump.in[S.d_layer + 1] + p * 256;
            dmp1 = a.dump.d[S.d_layer] + p * 256;
          } else if (S.epi == EPI_BWD) {
            dmp1 = a.dump.ga[S.d_layer] + p * 256;
            dmp0 = a.dump.gh[S.d_layer] ? a.dump.gh[S.d_layer] + p * 256 : dmp1;   // gh is optional (bwd16 skips it when equal)
          }
        }
        // slabs are announced as they complete unless something is staged into the A operand after the main loop
        const bool early = (s + 1 < prog.n_steps) && S.prep_next == PREP_NONE && (S.epi == EPI_HIDDEN || S.epi == EPI_BWD || S.epi == EPI_TAN);

        if (S.epi == EPI_HIDDEN) {
          uint32_t* dsave = (S.d_layer >= 0 && dscr) ? reinterpret_cast<uint32_t*>(dscr + (size_t)S.d_layer * 256 * TCM) : nullptr;
          if (S.act == TACT_SOFTPLUS) {
            if (S.row_off >= 0) {
              if (DUMP && dsave) epi_hidden<TACT_SOFTPLUS, true, 1, 0, false, true>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, dsave, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
              else if (dsave) epi_hidden<TACT_SOFTPLUS, true, 1, 0, false>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, dsave, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
              else epi_hidden<TACT_SOFTPLUS, false, 1, 0, false>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, nullptr, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
            } else if (S.n_valid < 256 || S.out_scale != 1.0f) {
              if (DUMP && dsave) epi_hidden<TACT_SOFTPLUS, true, 0, 0, true, true>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, dsave, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
              else if (dsave) epi_hidden<TACT_SOFTPLUS, true, 0, 0, true>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, dsave, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
              else epi_hidden<TACT_SOFTPLUS, false, 0, 0, true>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, nullptr, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
            } else {
              if (DUMP && dsave) epi_hidden<TACT_SOFTPLUS, true, 0, 0, false, true>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, dsave, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
              else if (dsave) epi_hidden<TACT_SOFTPLUS, true, 0, 0, false>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, dsave, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
              else epi_hidden<TACT_SOFTPLUS, false, 0, 0, false>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, nullptr, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
            }
            if (S.flags & TF_FEEDS_SKIP) {
              // x = cat([x, inputs]) / sqrt(2): encoding columns behind the n_valid outputs (fields.py:90-91)
              // (this thread owns columns [64 sl + 16 cq, +16) of slab sl: encoding elements q = column - n_valid)
              for (int sl = S.n_valid >> 6; sl < 4; ++sl) {
                const int q_lo = sl * 64 + 16 * cq - S.n_valid;
                if (q_lo + 16 <= 0 || q_lo >= prog.pe_dim) continue;
                float pe[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) pe[j] = (q_lo + j >= 0 && q_lo + j < prog.pe_dim) ? pes[(q_lo + j) * TCM + row] : 0.f;
#pragma unroll 1
                for (int j = 0; j < 16; ++j) {
                  const int q = q_lo + j;
                  if (q < 0 || q >= prog.pe_dim) continue;
                  const int n = S.n_valid + q;
                  const float x = pe[j] * 0.70710678118654752440f;
                  const __half h = __float2half_rn(x);
                  const uint32_t off = a_chunk_offset(n >> 6, row, (n & 63) >> 3) + (uint32_t)(n & 7) * 2u;
                  *reinterpret_cast<__half*>(a_hi + off) = h;
                  *reinterpret_cast<__half*>(a_lo + off) = __float2half_rn(x - __half2float(h));
                  if (DUMP && dmp0) dmp0[n] = x;
                }
              }
            }
          } else {
            if (S.n_small == 6) {
              sv[0] = st.pt[0]; sv[1] = st.pt[1]; sv[2] = st.pt[2]; sv[3] = st.nrm[0]; sv[4] = st.nrm[1]; sv[5] = st.nrm[2];
              epi_hidden<TACT_RELU, false, 0, 6, false>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, nullptr, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
            } else if (S.n_small == 3) {
              sv[0] = st.cg[0]; sv[1] = st.cg[1]; sv[2] = st.cg[2];
              epi_hidden<TACT_RELU, false, 3, 3, false>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, nullptr, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
            } else if (S.row_off >= 0) {
              epi_hidden<TACT_RELU, false, 3, 0, false>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, nullptr, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
            } else {
              epi_hidden<TACT_RELU, false, 0, 0, false>(S, packed, t_acc, row, cq, ar, bias_a, cst_a, nullptr, dot, sv, early, bar_slab, lane, ep, dmp0, dmp1);
            }
          }
        } else if (DUMP && S.epi == EPI_TAN) {
          const int64_t ro_ = p * 256;
          epi_tan(S, t_acc, row, cq, a_hi, a_lo, valid ? a.dump.d[S.d_layer] + ro_ : nullptr, valid ? a.dump.gh[S.d_layer] + ro_ : nullptr,
                  valid ? a.dump.ga[S.d_layer] + ro_ : nullptr, valid ? a.dump.in[S.d_layer + 1] + ro_ : nullptr, tan_inv_ts, early,
                  bar_slab, lane);
          if (S.flags & TF_FEEDS_SKIP) {  // t_0 / sqrt(2) behind the n_valid outputs (same ownership as the forward skip feed)
            for (int sl = S.n_valid >> 6; sl < 4; ++sl) {
              const int q_lo = sl * 64 + 16 * cq - S.n_valid;
              if (q_lo + 16 <= 0 || q_lo >= prog.pe_dim) continue;
#pragma unroll 1
              for (int j = 0; j < 16; ++j) {
                const int q = q_lo + j;
                if (q < 0 || q >= prog.pe_dim) continue;
                const int n = S.n_valid + q;
                const float t = pes[q * TCM + row] * 0.70710678118654752440f;
                const float x = t * tan_ts;
                const __half h = __float2half_rn(x);
                const uint32_t off = a_chunk_offset(n >> 6, row, (n & 63) >> 3) + (uint32_t)(n & 7) * 2u;
                *reinterpret_cast<__half*>(a_hi + off) = h;
                *reinterpret_cast<__half*>(a_lo + off) = __float2half_rn(x - __half2float(h));
                if (valid) a.dump.in[S.d_layer + 1][ro_ + n] = t;
              }
            }
          }
        } else if (S.epi == EPI_BWD) {
          const uint32_t* Dl = reinterpret_cast<const uint32_t*>(dscr + (size_t)S.d_layer * 256 * TCM);
          const bool masked = S.n_valid < 256 || (S.flags & TF_SKIP_BWD) != 0;   // the skip layer's steps only
          if (DUMP) {
            if (masked) epi_bwd<true, true>(S, prog, t_acc, row, cq, ar, Dl, gxs, early, bar_slab, lane, ep, pre8, pre8b, dmp0, dmp1);
            else epi_bwd<false, true>(S, prog, t_acc, row, cq, ar, Dl, gxs, early, bar_slab, lane, ep, pre8, pre8b, dmp0, dmp1);
          } else {
            if (masked) epi_bwd<true, false>(S, prog, t_acc, row, cq, ar, Dl, gxs, early, bar_slab, lane, ep, pre8, pre8b, dmp0, dmp1);
            else epi_bwd<false, false>(S, prog, t_acc, row, cq, ar, Dl, gxs, early, bar_slab, lane, ep, pre8, pre8b, dmp0, dmp1);
          }
        } else if (S.epi == EPI_BWD_LAST) {
          // adjoint of the encoding -> d sdf / d x (all four threads of the row compute it): one sincosf per
          // (frequency, dim) serves the sin and the cos column (PositionEncoding.py:51-76: [x | sin f x | cos f x]_f)
          // Streaming form with static column indices (no local array: at ~226 KB of shared memory there is next to no L1, so
          // local memory is an L2 round trip): column q of the encoding adjoint is x_d (q < 3), sin(2^k x_d) (q = 3 + 6k + d) or
          // cos(2^k x_d) (q = 6 + 6k + d); its contribution to d sdf / d x_d needs the partner value from the tile's cached
          // encoding.  All loads are independent (issued up front by the unrolled code).
          float gq[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (cc * 16 < prog.pe_dim) {   // warp-uniform
              // every load of the chunk up front, unconditionally (scratch rows [0, 64) exist; the partner index is clamped), so
              // that the L2 latency is paid once per chunk instead of once per column
              float g0[16], sk[16], pr[16];
              tmem_ld16(t_acc + cc * 16, g0);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int q = cc * 16 + i;   // compile-time
                const int k = q >= 3 ? (q - 3) / 6 : 0, r = q >= 3 ? (q - 3) % 6 : 0, d = r % 3;
                const int partner = q < 3 ? 0 : (r < 3 ? 6 + 6 * k + d : 3 + 6 * k + d);
                sk[i] = prog.has_skip ? gxs[q * TCM + row] : 0.0f;
                pr[i] = pes[(partner < 64 ? partner : 63) * TCM + row];
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int q = cc * 16 + i;
                // scratch rows beyond the encoding are never written: select, do not multiply (they may hold anything)
                const float gv = (q < prog.pe_dim) ? fmaf(g0[i], S.inv_scale, sk[i]) : 0.0f;
                if (DUMP && writer && q < prog.pe_dim) a.dump.gx0[p * a.dump.ld_gx0 + q] = gv;
                if (q < 3) {
                  gq[q] += gv;
                } else {
                  const int k = (q - 3) / 6, r = (q - 3) % 6, d = r % 3;
                  const float f = (r < 3) ? (float)(1 << k) : -(float)(1 << k);
                  // d/dx sin(f x) = f cos(f x) ; d/dx cos(f x) = -f sin(f x); columns beyond the encoding contribute nothing
                  gq[d] = fmaf(f * pr[i], gv, gq[d]);
                }
              }
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) st.nrm[c] = gq[c] * prog.sdf_scale;
          if (writer && a.out_grad) { a.out_grad[p * 3] = st.nrm[0]; a.out_grad[p * 3 + 1] = st.nrm[1]; a.out_grad[p * 3 + 2] = st.nrm[2]; }
        } else {  // EPI_PARK: feature block of the last SDF layer -> fp32 scratch slot (read back by the colour stage)
          float* fslot = dscr ? dscr + (size_t)prog.n_hidden * 256 * TCM : nullptr;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            const int n0 = c * 64 + cq * 16;
            float v[16];
            tmem_ld16(t_acc + n0, v);
            const float4 fb0 = lds128(bias_a + c * 256), fb1 = lds128(bias_a + c * 256 + 16), fb2 = lds128(bias_a + c * 256 + 32),
                         fb3 = lds128(bias_a + c * 256 + 48);
            const float fb[16] = {fb0.x, fb0.y, fb0.z, fb0.w, fb1.x, fb1.y, fb1.z, fb1.w, fb2.x, fb2.y, fb2.z, fb2.w, fb3.x, fb3.y, fb3.z, fb3.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float f = fmaf(v[j], prog.feat_inv_scale, fb[j]);
              if (fslot) fslot[(n0 + j) * TCM + row] = f;
              if (valid && a.out_full) a.out_full[p * 257 + 1 + n0 + j] = f;
            }
          }
          if (writer && a.out_full) a.out_full[p * 257] = st.sdf;
        }

        ep.mark(S.epi == EPI_HIDDEN || S.epi == EPI_BWD ? 6 : (S.epi == EPI_PARK ? 7 : 8));  // skip feed | PARK | BWD_LAST body
        if (early) slab_ready(&bar_slab[3], lane);  // nothing below touches the A operand of an "early" step
        ep.mark(9);

        // ---------------------------------------------------------------- narrow layers folded into this epilogue
        if (S.post != POST_NONE) {
          // Combine the partial dot products of the row's four threads.  They sit in four different warps but serve the same
          // TMEM lane, and this step's accumulator has been consumed (the next layer's MMAs write the other half of TMEM): every
          // thread parks its three sums in the first columns of its own section-0 range, the four warps of the lane quarter
          // meet at a 128-thread named barrier, and everyone reads the four triples back -- no global-memory round trip.
          tmem_st4(t_acc + 16 * cq, dot[0], dot[1], dot[2], 0.0f);
          tmem_st_wait();
          tc_fence_before();
          asm volatile("bar.sync %0, 128;" ::"r"(2 + (warp & 3)) : "memory");
          tc_fence_after();
          {
            uint32_t q0[4], q1[4], q2[4], q3[4];
            tmem_ld4_nowait(t_acc, q0);
            tmem_ld4_nowait(t_acc + 16, q1);
            tmem_ld4_nowait(t_acc + 32, q2);
            tmem_ld4_nowait(t_acc + 48, q3);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 3; ++c)
              dot[c] = (__uint_as_float(q0[c]) + __uint_as_float(q1[c])) + (__uint_as_float(q2[c]) + __uint_as_float(q3[c]));
          }
          if (S.post == POST_SDF) {
            st.sdf = (dot[0] + __ldg(packed + S.row_bias_off)) / prog.sdf_scale;
            if (writer && a.out_sdf) a.out_sdf[p] = a.out_sdf_sign * st.sdf;
          } else if (S.post == POST_CG) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              st.cg[c] = dot[c] + __ldg(packed + S.row_bias_off + c);
              if (prog.color_squeeze) st.cg[c] = sigmoidf_(st.cg[c]);
            }
            if (writer && a.out_color) { a.out_color[p * 3] = st.cg[0]; a.out_color[p * 3 + 1] = st.cg[1]; a.out_color[p * 3 + 2] = st.cg[2]; }
          } else {  // POST_DRGB
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float dr = dot[c] + __ldg(packed + S.row_bias_off + c);
              float rel;
              if (prog.relight_inv_sigmoid) {  // sigmoid(inverse_sigmoid(rgb) + drgb), eps 1e-5 (transform.py:317-320)
                const float x = fminf(fmaxf(st.cg[c], 0.0f), 1.0f);
                rel = sigmoidf_(logf(fmaxf(x, 1e-5f) / fmaxf(1.0f - x, 1e-5f)) + dr);
              } else {
                rel = fminf(fmaxf(st.cg[c] + sigmoidf_(dr) - 0.5f, 0.0f), 1.0f);
              }
              if (writer && a.out_drgb) a.out_drgb[p * 3 + c] = dr;
              if (writer && a.out_relit) a.out_relit[p * 3 + c] = rel;
            }
          }
        }

        ep.mark(10);  // narrow-layer exchange + post
        // ---------------------------------------------------------------- stage the A operand of the next step
        bool announced = false;  // the staging below announces its slabs itself
        if (S.prep_next == PREP_SEED) {
          // d sdf / d a_last = W_last[0,:] / scale (.) softplus'(a_last); same (row, column) ownership as the epilogue
          // that stored softplus'; all scratch loads are issued before the first use (L2 latency paid once)
          const uint32_t* D = reinterpret_cast<const uint32_t*>(dscr + (size_t)(prog.n_hidden - 1) * 256 * TCM);
          uint32_t dw[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) dw[i] = D[(((i >> 3) * 64 + cq * 16) / 2 + (i & 7)) * TCM + row];
          const float ssc = BWD_ASCALE / (prog.sdf_scale * 65535.0f);
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              const int nb = sl * 64 + cq * 16 + h8 * 8;
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(packed + prog.seed_row_off + nb));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(packed + prog.seed_row_off + nb) + 1);
              const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t w = dw[sl * 8 + h8 * 4 + (j >> 1)];
                const float q = __uint_as_float(__byte_perm(w, 0x4B000000u, (j & 1) ? 0x7632 : 0x7610)) - 8388608.0f;
                o[j] = (wv[j] * ssc) * q;
              }
              write_a8(a_hi, a_lo, nb >> 6, row, (nb & 63) >> 3, o);
            }
            slab_ready(&bar_slab[sl], lane);
          }
          announced = true;
        } else if (S.prep_next == PREP_COLOR_IN) {
          // colour input = [feature vector (scratch slot)] + small block [pts | PE(view) | normal]; same ownership as
          // the EPI_PARK store, loads issued two slabs ahead
          const float* fslot = dscr + (size_t)prog.n_hidden * 256 * TCM;
          const bool self_announce = prog.n_stages == 3;  // otherwise the small slab is staged after slab 3
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fslot[((half * 2 + (i >> 4)) * 64 + cq * 16 + (i & 15)) * TCM + row];
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {
                const int nb = (half * 2 + sl) * 64 + cq * 16 + h8 * 8;
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = f[sl * 16 + h8 * 8 + j];
                write_a8(a_hi, a_lo, nb >> 6, row, (nb & 63) >> 3, o);
              }
              if (self_announce) slab_ready(&bar_slab[half * 2 + sl], lane);
            }
          }
          announced = self_announce;
          if (prog.n_stages == 2) stage_small(a_hi, a_lo, SMALL_SLAB, row, cq, prog, st, SMALL_COLOR, a.viewdir_mode);
        } else if (S.prep_next == PREP_RELIGHT_IN) {
          if (prog.n_stages == 3) epi_bar_sync();  // slab 0 was just written column-wise by other threads of this row
          stage_small(a_hi, a_lo, prog.n_stages == 3 ? 0 : SMALL_SLAB, row, cq, prog, st, SMALL_RELIGHT, 0);
        } else if (S.prep_next == PREP_CG) {
          if (prog.n_stages == 2) stage_small(a_hi, a_lo, SMALL_SLAB, row, cq, prog, st, SMALL_CG, 0);
        }

        if (!early && !announced && s + 1 < prog.n_steps) slabs_ready_all(bar_slab, lane);
        ep.mark(S.prep_next == PREP_NONE ? 11 : (S.prep_next < PREP_CG ? 11 + S.prep_next : 15));  // 11: announce all (late); 13 seed, 14 colour in, 15 relight in / cg
#ifdef CNEUS_TC_EPI_PROF
        if (ep.on) {  // epilogue cycles (accumulators ready -> step done) by step type, g_tc_prof_type[2 t] cycles / [2 t + 1] count:
                      // 0 softplus + softplus' saved, 1 softplus, 2 gradient chain, 3 ReLU, 4 feature block, 5 encoding adjoint
          const int ty = S.epi == EPI_HIDDEN ? (S.act == TACT_SOFTPLUS ? (S.d_layer >= 0 && dscr ? 0 : 1) : 3)
                                             : (S.epi == EPI_BWD ? 2 : (S.epi == EPI_PARK ? 4 : 5));
          g_tc_prof_type[2 * ty] += (unsigned long long)(clock64() - t_step0);
          g_tc_prof_type[2 * ty + 1] += 1;
        }
#endif
      }
    }
    if (prof) {
      const int o = threadIdx.x == 0 ? 4 : 12;
      g_tc_prof[o] += (unsigned long long)t_wacc; g_tc_prof[o + 1] += (unsigned long long)(clock64() - t_begin);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) {
    cluster_sync_all();  // neither CTA leaves (or frees its TMEM) while the other may still reach into it
    if (warp == TC_EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  } else {
    if (warp == TC_EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

template __global__ void shade_tc_kernel<false>(const __grid_constant__ TcProgram, const float* __restrict__, const __grid_constant__ ShadeArgs,
                                                float* __restrict__);
template __global__ void shade_tc_kernel<true>(const __grid_constant__ TcProgram, const float* __restrict__, const __grid_constant__ ShadeArgs,
                                               float* __restrict__);

}  // namespace cneus

extern "C" int cneus_tc_prof_read_types(unsigned long long* out16, int reset) {
  if (cudaMemcpyFromSymbol(out16, cneus::g_tc_prof_type, 16 * sizeof(unsigned long long)) != cudaSuccess) return CNEUS_ECUDA;
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(cneus::g_tc_prof_type, z, sizeof(z)) != cudaSuccess) return CNEUS_ECUDA;
  }
  return CNEUS_OK;
}
extern "C" int cneus_tc_prof_read(unsigned long long* out32, int reset) {
  if (cudaMemcpyFromSymbol(out32, cneus::g_tc_prof, 32 * sizeof(unsigned long long)) != cudaSuccess) return CNEUS_ECUDA;
  if (reset) {
    unsigned long long z[32] = {0};
    if (cudaMemcpyToSymbol(cneus::g_tc_prof, z, sizeof(z)) != cudaSuccess) return CNEUS_ECUDA;
  }
  return CNEUS_OK;
}
