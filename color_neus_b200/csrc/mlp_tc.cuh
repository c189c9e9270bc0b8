// Shared declarations of the tensor-core point-shading path (mlp_tc_kernel.cu / mlp_tc_host.cu).
#pragma once
#include "common.cuh"

namespace cneus {


constexpr int TCM = 128;

constexpr int SLAB_BYTES = 16384;          // [128 rows][64 halfs], SWIZZLE_128B
constexpr int A_SLABS = 5;                 // 4 main K-blocks + 1 small-input block
constexpr int TC_STAGES = 2;              // 32 KB ring stages next to the A slabs (one-CTA build; a third one reuses the small-input
                                          // slab); the pair build cuts the same 64 KB into 16 KB slots (TC_PAIR_RING_SLOTS + constants)
constexpr int STAGE_BYTES = 2 * SLAB_BYTES;  // hi slab + lo slab of one (K-block, N-half)
constexpr float W_SCALE = 64.0f;
constexpr float BWD_ASCALE = 256.0f;       // scale of the A operand in the gradient chain
constexpr int MAX_TC_STEPS = 28;
// per-CTA fp32 scratch rows of TCM floats behind the softplus' slots: [0,64) encoding part of the skip layer's adjoint,
// [64,96) two exchange areas of the narrow-layer partial sums, [96,160) the tile's positional encoding
constexpr int TC_GXS_ROWS = 160;
constexpr int TC_GXS_XCH = 64, TC_GXS_PE = 96;
constexpr int SMALL_SLAB = 4;

// CTA pairs (cta_group::2) unless built with -DCNEUS_TC_SINGLE
#ifdef CNEUS_TC_SINGLE
constexpr bool kPair = false;
#else
constexpr bool kPair = true;
#endif

// Pair build: the narrow-layer weight rows / rank-update rows of a step ([row_n + n_small][256] fp32, up to 6 KB) are staged
// in shared memory next to the bias and read with broadcast LDS.128 (the ring gives up one of its four 16 KB slots for
// them); otherwise (-DCNEUS_TC_SHFL_CONSTS, and the one-CTA build, whose shared memory is full) every lane holds the
// constants of two columns and the epilogue fetches them with one shuffle per element and row.
#if defined(CNEUS_TC_SINGLE) || defined(CNEUS_TC_SHFL_CONSTS)
constexpr bool kSmemConsts = false;
#else
constexpr bool kSmemConsts = true;
#endif
constexpr int TC_PAIR_RING_SLOTS = kSmemConsts ? 3 : 4;  // 16 KB slots in the ring proper (+ 2: the small-input slabs)

enum { EPI_HIDDEN = 0, EPI_PARK = 1, EPI_BWD = 2, EPI_BWD_LAST = 3, EPI_TAN = 4 };
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
  int8_t n_small;         // rank-n_small fp32 update folded into this epilogue (0, 3: re-injected colour, 6: pts|normal)
  int32_t small_off;      // float offset of its weight rows [n_small][256]
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
  int32_t prof;               // 1: CTA 0 accumulates role cycle counters (cneus_tc_prof_read)
  int32_t tangent;            // 1: tangent-pass program (training instantiation only)
  int32_t n_stages;           // weight ring depth: 3 when the small-input slab is not needed (its 32 KB become a stage)
};


constexpr int TC_EPI_WARPS = 16;                      // four warps per TMEM lane quarter, each owning 64 columns
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
// + one warpgroup for the bulk-copy producer warp and the MMA issuer warp (two idle warps complete it: setmaxnreg is a
// warpgroup-wide instruction); that warpgroup gives registers up, the epilogue warpgroups take them
constexpr int TC_KERNEL_THREADS = TC_EPI_THREADS + 128;
constexpr int TC_REGS_EPI = 112, TC_REGS_OTHER = 32;  // the pool is the launch allocation: 640 x 96 = 512 x 112 + 128 x 32
// A planes + weight ring + barriers (256 B) + the step's bias (1 KB, read by the epilogue with broadcast LDS) + alignment slack
constexpr size_t TC_SMEM_BYTES = 2 * A_SLABS * SLAB_BYTES + TC_STAGES * STAGE_BYTES + 256 + 1024 + 1024;
static_assert(TC_SMEM_BYTES <= 232448, "exceeds the 227 KB a CTA can opt into on sm_100");

template <bool DUMP>
__global__ void shade_tc_kernel(const __grid_constant__ TcProgram prog, const float* __restrict__ packed,
                                const __grid_constant__ ShadeArgs a, float* __restrict__ gxscratch);

}  // namespace cneus
