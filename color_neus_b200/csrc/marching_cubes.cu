// marching cubes on the device, consuming the SDF grid of cneus_sdf_grid in place
// (replaces `mcubes.marching_cubes(u, threshold)`, lib/models/renderers/NeuS.py:35 -- PyMCubes, third party, one CPU core).
//
// Conventions (shared with oracle/mc_oracle.py, which documents why parity with PyMCubes itself is unpinned):
//   * grid u[nx][ny][nz] fp32, x-major (extract_fields, NeuS.py:14-28); corner "inside" <=> (double)u < iso;
//   * classic corner / edge numbering; the per-case triangulation is derived at first use on the host (derive_tables):
//     face segments that cut off the inside corners, chained into loops, triangulated without in-face diagonals;
//   * every grid point owns the three edges leaving it along +x, +y, +z; vertex order = (grid point, axis), triangle
//     order = (cell, table order); vertices in float64 index coordinates, a + (iso - f1) / (f2 - f1) along the axis.
// Memory-bound streaming kernels over the grid (a 512^3 grid is 512 MB; everything here is a few passes over it):
//   count:  classify (1 byte per grid point: 3 cut flags + triangle count of its cell) + per-chunk sums, scan of the sums;
//   emit:   per-chunk scans -> vertex slots (offset << 3 | flags per grid point) and vertices; then triangles, looking the
//           twelve edges' vertex ids up through the owners' slots.
#include <string.h>

#include <array>
#include <vector>

#include "common.cuh"

namespace cneus {

constexpr int MC_THREADS = 256;
constexpr int MC_ITEMS = 8;
constexpr int MC_CHUNK = MC_THREADS * MC_ITEMS;

__constant__ uint8_t c_mc_ntri[256];
__constant__ int8_t c_mc_tri[256 * 16];
// owner offset (dx, dy, dz) and axis of the twelve edges, packed dx | dy << 1 | dz << 2 | axis << 3
__constant__ uint8_t c_mc_owner[12] = {0 | (0 << 3), 1 | (1 << 3), 2 | (0 << 3), 0 | (1 << 3), 4 | (0 << 3), 5 | (1 << 3),
                                       6 | (0 << 3), 4 | (1 << 3), 0 | (2 << 3), 1 | (2 << 3), 3 | (2 << 3), 2 | (2 << 3)};

// ---- host: derivation of the case table ---------------------------------------------------------------------------------
static const int H_EDGE[12][2] = {{0, 1}, {1, 2}, {3, 2}, {0, 3}, {4, 5}, {5, 6}, {7, 6}, {4, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
static const int H_FACE[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {2, 3, 7, 6}, {0, 4, 7, 3}, {1, 2, 6, 5}};  // ccw from outside

static int h_edge_between(int a, int b) {
  for (int e = 0; e < 12; ++e)
    if ((H_EDGE[e][0] == a && H_EDGE[e][1] == b) || (H_EDGE[e][0] == b && H_EDGE[e][1] == a)) return e;
  return -1;
}
static bool h_share_face(int e1, int e2) {
  for (int f = 0; f < 6; ++f) {
    int hit = 0;
    for (int i = 0; i < 4; ++i) {
      const int e = h_edge_between(H_FACE[f][i], H_FACE[f][(i + 1) & 3]);
      hit += (e == e1) + (e == e2);
    }
    if (hit == 2) return true;
  }
  return false;
}

// All triangulations of the polygon with vertices i..j (index triples in increasing order keep the orientation), in the
// order: apex k of the triangle on the chord (i, j) ascending, left part, right part; polygons have at most 7 vertices.
typedef std::array<int, 3> Tri3;
typedef std::vector<Tri3> Triangulation;
static std::vector<Triangulation> h_triangulations(int i, int j) {
  std::vector<Triangulation> out;
  if (j - i < 2) { out.push_back(Triangulation()); return out; }
  for (int k = i + 1; k < j; ++k) {
    const std::vector<Triangulation> left = h_triangulations(i, k), right = h_triangulations(k, j);
    for (const Triangulation& l : left)
      for (const Triangulation& r : right) {
        Triangulation t = l;
        t.push_back(Tri3{i, k, j});
        t.insert(t.end(), r.begin(), r.end());
        out.push_back(t);
      }
  }
  return out;
}

static void derive_tables(uint8_t* ntri, int8_t* tri) {
  for (int c = 0; c < 256; ++c) {
    int nxt[12];
    for (int e = 0; e < 12; ++e) nxt[e] = -1;
    for (int f = 0; f < 6; ++f) {
      int cut_e[4], cut_kind[4], n = 0;  // kind +1: inside -> outside walking ccw, -1: outside -> inside
      for (int i = 0; i < 4; ++i) {
        const int a = H_FACE[f][i], b = H_FACE[f][(i + 1) & 3];
        const int ia = (c >> a) & 1, ib = (c >> b) & 1;
        if (ia != ib) { cut_e[n] = h_edge_between(a, b); cut_kind[n] = ia ? +1 : -1; ++n; }
      }
      // each run of inside corners is cut off by a segment from its leave cut back to its enter cut (inside on the left)
      for (int i = 0; i < n; ++i)
        if (cut_kind[i] == -1) nxt[cut_e[(i + 1) % n]] = cut_e[i];
    }
    bool seen[12] = {false};
    int nt = 0;
    for (int s = 0; s < 12; ++s) {
      if (nxt[s] < 0 || seen[s]) continue;
      int loop[12], n = 0;
      for (int e = s; !seen[e]; e = nxt[e]) { seen[e] = true; loop[n++] = e; }
      // first triangulation (fans preferred) without a diagonal lying in a cube face: such a diagonal could coincide
      // with a face segment of the neighbouring cell
      const Triangulation* best = nullptr;
      bool best_fan = false;
      const std::vector<Triangulation> all = h_triangulations(0, n - 1);
      for (const Triangulation& T : all) {
        bool clean = true;
        for (const Tri3& t : T)
          for (int q = 0; q < 3 && clean; ++q) {
            const int p0 = t[q], p1 = t[(q + 1) % 3];
            const int d = ((p1 - p0) % n + n) % n;
            if (d != 1 && d != n - 1 && h_share_face(loop[p0], loop[p1])) clean = false;
          }
        if (!clean) continue;
        bool fan = false;
        for (int a = 0; a < n && !fan; ++a) {
          bool every = true;
          for (const Tri3& t : T) every = every && (t[0] == a || t[1] == a || t[2] == a);
          fan = every;
        }
        if (!best || (fan && !best_fan)) { best = &T; best_fan = fan; }
        if (fan) break;
      }
      if (!best) best = &all[0];  // cannot happen (every loop of every case has a clean triangulation)
      for (const Tri3& t : *best) {
        for (int q = 0; q < 3; ++q) tri[c * 16 + nt * 3 + q] = (int8_t)loop[t[q]];
        ++nt;
      }
    }
    ntri[c] = (uint8_t)nt;
    for (int q = nt * 3; q < 16; ++q) tri[c * 16 + q] = -1;
  }
}

static uint8_t h_ntri[256];
static int8_t h_tri[256 * 16];
static bool h_tables_ready = false;
static void ensure_tables() {
  if (!h_tables_ready) { derive_tables(h_ntri, h_tri); h_tables_ready = true; }
}

// ---- device -------------------------------------------------------------------------------------------------------------
struct McGrid {
  const float* u;
  int32_t nx, ny, nz;
  int64_t n;  // nx * ny * nz
  double iso;
};

__device__ __forceinline__ bool mc_in(const McGrid& g, int64_t p) { return (double)g.u[p] < g.iso; }

__device__ __forceinline__ int mc_case(const McGrid& g, int64_t p) {
  const int64_t sx = (int64_t)g.ny * g.nz, sy = g.nz;
  return (int)mc_in(g, p) | ((int)mc_in(g, p + sx) << 1) | ((int)mc_in(g, p + sx + sy) << 2) | ((int)mc_in(g, p + sy) << 3) |
         ((int)mc_in(g, p + 1) << 4) | ((int)mc_in(g, p + sx + 1) << 5) | ((int)mc_in(g, p + sx + sy + 1) << 6) |
         ((int)mc_in(g, p + sy + 1) << 7);
}

// exclusive scan of two counters over the block; returns this thread's exclusive prefixes, totals in tot_*
__device__ __forceinline__ void block_scan2(uint32_t& a, uint32_t& b, uint32_t& tot_a, uint32_t& tot_b, uint32_t (*sm)[2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t ia = a, ib = b;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
    if (lane >= o) { ia += ta; ib += tb; }
  }
  if (lane == 31) { sm[warp][0] = ia; sm[warp][1] = ib; }
  __syncthreads();
  uint32_t wa = 0, wb = 0, ta = 0, tb = 0;
  for (int w = 0; w < MC_THREADS / 32; ++w) {
    if (w < warp) { wa += sm[w][0]; wb += sm[w][1]; }
    ta += sm[w][0]; tb += sm[w][1];
  }
  __syncthreads();
  a = wa + ia - a; b = wb + ib - b;
  tot_a = ta; tot_b = tb;
}

// code[p] = cut flags of the owned edges (bits 0-2: +x, +y, +z) | triangle count of the cell at p << 3
__global__ void __launch_bounds__(MC_THREADS) mc_classify_kernel(const __grid_constant__ McGrid g, uint8_t* __restrict__ code,
                                                                 uint2* __restrict__ partial) {
  __shared__ uint32_t sm[MC_THREADS / 32][2];
  const int64_t p0 = ((int64_t)blockIdx.x * MC_THREADS + threadIdx.x) * MC_ITEMS;
  const int64_t sx = (int64_t)g.ny * g.nz, sy = g.nz;
  uint32_t nv = 0, nt = 0;
  uint8_t out[MC_ITEMS];
#pragma unroll
  for (int i = 0; i < MC_ITEMS; ++i) {
    const int64_t p = p0 + i;
    uint8_t c = 0;
    if (p < g.n) {
      const int z = (int)(p % g.nz), y = (int)((p / g.nz) % g.ny), x = (int)(p / sx);
      const bool in0 = mc_in(g, p);
      const bool hx = x + 1 < g.nx, hy = y + 1 < g.ny, hz = z + 1 < g.nz;
      if (hx && mc_in(g, p + sx) != in0) c |= 1;
      if (hy && mc_in(g, p + sy) != in0) c |= 2;
      if (hz && mc_in(g, p + 1) != in0) c |= 4;
      if (hx && hy && hz) {
        const int t = c_mc_ntri[mc_case(g, p)];
        c |= (uint8_t)(t << 3);
        nt += t;
      }
      nv += __popc(c & 7);
    }
    out[i] = c;
  }
  if (p0 + MC_ITEMS <= g.n) {
    *reinterpret_cast<uint2*>(code + p0) = *reinterpret_cast<uint2*>(out);
  } else {
    for (int i = 0; i < MC_ITEMS; ++i)
      if (p0 + i < g.n) code[p0 + i] = out[i];
  }
  uint32_t ta, tb;
  block_scan2(nv, nt, ta, tb, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = make_uint2(ta, tb);
}

// exclusive scan of the per-chunk sums in place (one block); totals -> counts[0..1]
__global__ void __launch_bounds__(1024) mc_scan_partials_kernel(uint2* __restrict__ partial, int64_t n_chunks, int64_t* __restrict__ counts) {
  __shared__ unsigned long long sm[32][2];
  __shared__ unsigned long long carry[2];
  if (threadIdx.x == 0) { carry[0] = 0; carry[1] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n_chunks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const uint2 v = i < n_chunks ? partial[i] : make_uint2(0, 0);
    unsigned long long a = v.x, b = v.y;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) { a += ta; b += tb; }
    }
    if (lane == 31) { sm[warp][0] = a; sm[warp][1] = b; }
    __syncthreads();
    unsigned long long wa = carry[0], wb = carry[1], ta = 0, tb = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) { wa += sm[w][0]; wb += sm[w][1]; }
      ta += sm[w][0]; tb += sm[w][1];
    }
    // offsets beyond 32 bits cannot be represented in the slots; the host checks the totals before emitting
    if (i < n_chunks) partial[i] = make_uint2((uint32_t)(wa + a - v.x), (uint32_t)(wb + b - v.y));
    __syncthreads();
    if (threadIdx.x == 0) { carry[0] += ta; carry[1] += tb; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { counts[0] = (int64_t)carry[0]; counts[1] = (int64_t)carry[1]; }
}

// vertex slots (offset << 3 | flags) for every grid point + the vertices themselves
__global__ void __launch_bounds__(MC_THREADS) mc_vertices_kernel(const __grid_constant__ McGrid g, const uint8_t* __restrict__ code,
                                                                 const uint2* __restrict__ partial, uint32_t* __restrict__ slot,
                                                                 double* __restrict__ vertices) {
  __shared__ uint32_t sm[MC_THREADS / 32][2];
  const int64_t p0 = ((int64_t)blockIdx.x * MC_THREADS + threadIdx.x) * MC_ITEMS;
  const int64_t sx = (int64_t)g.ny * g.nz, sy = g.nz;
  uint8_t c[MC_ITEMS];
  uint32_t nv = 0, dummy = 0;
#pragma unroll
  for (int i = 0; i < MC_ITEMS; ++i) {
    c[i] = (p0 + i < g.n) ? code[p0 + i] : 0;
    nv += __popc(c[i] & 7);
  }
  uint32_t ta, tb;
  block_scan2(nv, dummy, ta, tb, sm);
  uint32_t off = partial[blockIdx.x].x + nv;
#pragma unroll
  for (int i = 0; i < MC_ITEMS; ++i) {
    const int64_t p = p0 + i;
    if (p >= g.n) break;
    const uint32_t fl = c[i] & 7;
    slot[p] = (off << 3) | fl;
    if (fl) {
      const int z = (int)(p % g.nz), y = (int)((p / g.nz) % g.ny), x = (int)(p / sx);
      const double f1 = (double)g.u[p];
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        if (!(fl & (1u << ax))) continue;
        const double f2 = (double)g.u[p + (ax == 0 ? sx : (ax == 1 ? sy : 1))];
        const double t = __ddiv_rn(__dsub_rn(g.iso, f1), __dsub_rn(f2, f1));
        double* v = vertices + (size_t)off * 3;
        v[0] = ax == 0 ? __dadd_rn((double)x, t) : (double)x;
        v[1] = ax == 1 ? __dadd_rn((double)y, t) : (double)y;
        v[2] = ax == 2 ? __dadd_rn((double)z, t) : (double)z;
        ++off;
      }
    }
  }
}

__global__ void __launch_bounds__(MC_THREADS) mc_triangles_kernel(const __grid_constant__ McGrid g, const uint8_t* __restrict__ code,
                                                                  const uint2* __restrict__ partial, const uint32_t* __restrict__ slot,
                                                                  int32_t* __restrict__ triangles) {
  __shared__ uint32_t sm[MC_THREADS / 32][2];
  const int64_t p0 = ((int64_t)blockIdx.x * MC_THREADS + threadIdx.x) * MC_ITEMS;
  const int64_t sx = (int64_t)g.ny * g.nz, sy = g.nz;
  uint8_t c[MC_ITEMS];
  uint32_t nt = 0, dummy = 0;
#pragma unroll
  for (int i = 0; i < MC_ITEMS; ++i) {
    c[i] = (p0 + i < g.n) ? code[p0 + i] : 0;
    nt += c[i] >> 3;
  }
  uint32_t ta, tb;
  block_scan2(dummy, nt, ta, tb, sm);
  uint32_t off = partial[blockIdx.x].y + nt;
#pragma unroll 1
  for (int i = 0; i < MC_ITEMS; ++i) {
    const int n_t = c[i] >> 3;
    if (!n_t) continue;
    const int64_t p = p0 + i;
    const int cs = mc_case(g, p);
    for (int q = 0; q < n_t * 3; ++q) {
      const int e = c_mc_tri[cs * 16 + q];
      const uint32_t o = c_mc_owner[e];
      const int64_t po = p + ((o & 1) ? sx : 0) + ((o & 2) ? sy : 0) + ((o & 4) ? 1 : 0);
      const uint32_t ax = o >> 3, w = slot[po];
      triangles[(size_t)off * 3 + q] = (int32_t)((w >> 3) + __popc(w & ((1u << ax) - 1u)));
    }
    off += n_t;
  }
}

struct McWs {
  uint8_t* code;
  uint32_t* slot;
  uint2* partial;
  int64_t n_chunks;
};
static size_t mc_ws_layout(int64_t n, void* base, McWs* w) {
  const int64_t n_chunks = (n + MC_CHUNK - 1) / MC_CHUNK;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_code = take((size_t)n), o_slot = take((size_t)n * 4), o_part = take((size_t)n_chunks * sizeof(uint2));
  if (w) {
    uint8_t* b = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(base) + 255) & ~(uintptr_t)255);
    w->code = b + o_code; w->slot = reinterpret_cast<uint32_t*>(b + o_slot); w->partial = reinterpret_cast<uint2*>(b + o_part);
    w->n_chunks = n_chunks;
  }
  return off + 256;
}

static int mc_check(const float* u, int32_t nx, int32_t ny, int32_t nz, const void* ws, size_t ws_bytes, const char* what) {
  if (!u || !ws || nx < 1 || ny < 1 || nz < 1) { set_error("%s: bad argument", what); return CNEUS_EINVAL; }
  const int64_t n = (int64_t)nx * ny * nz;
  if ((n + MC_CHUNK - 1) / MC_CHUNK > 0x7fffffff) { set_error("%s: grid too large", what); return CNEUS_EUNSUPPORTED; }
  if (ws_bytes < mc_ws_layout(n, nullptr, nullptr)) { set_error("%s: workspace too small", what); return CNEUS_ENOSPACE; }
  return CNEUS_OK;
}

}  // namespace cneus

extern "C" int cneus_mc_tables(uint8_t* n_tri, int8_t* tri) {
  using namespace cneus;
  if (!n_tri || !tri) { set_error("mc_tables: null pointer"); return CNEUS_EINVAL; }
  ensure_tables();
  memcpy(n_tri, h_ntri, sizeof(h_ntri));
  memcpy(tri, h_tri, sizeof(h_tri));
  return CNEUS_OK;
}

extern "C" size_t cneus_mc_workspace_bytes(int32_t nx, int32_t ny, int32_t nz) {
  if (nx < 1 || ny < 1 || nz < 1) return 0;
  return cneus::mc_ws_layout((int64_t)nx * ny * nz, nullptr, nullptr);
}

extern "C" int cneus_mc_count(const float* u, int32_t nx, int32_t ny, int32_t nz, double iso, void* ws, size_t ws_bytes,
                              int64_t* counts, void* stream) {
  CNEUS_NVTX_RANGE();
  using namespace cneus;
  int rc = mc_check(u, nx, ny, nz, ws, ws_bytes, "mc_count");
  if (rc) return rc;
  if (!counts) { set_error("mc_count: counts is null"); return CNEUS_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  ensure_tables();
  CNEUS_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_mc_ntri, h_ntri, sizeof(h_ntri), 0, cudaMemcpyHostToDevice, st));
  CNEUS_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_mc_tri, h_tri, sizeof(h_tri), 0, cudaMemcpyHostToDevice, st));
  McGrid g{u, nx, ny, nz, (int64_t)nx * ny * nz, iso};
  McWs w;
  mc_ws_layout(g.n, ws, &w);
  mc_classify_kernel<<<(unsigned)w.n_chunks, MC_THREADS, 0, st>>>(g, w.code, w.partial);
  mc_scan_partials_kernel<<<1, 1024, 0, st>>>(w.partial, w.n_chunks, counts);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch(2);
  return CNEUS_OK;
}

extern "C" int cneus_mc_emit(const float* u, int32_t nx, int32_t ny, int32_t nz, double iso, void* ws, size_t ws_bytes,
                             int64_t n_vertices, int64_t n_triangles, double* vertices, int32_t* triangles, void* stream) {
  CNEUS_NVTX_RANGE();
  using namespace cneus;
  int rc = mc_check(u, nx, ny, nz, ws, ws_bytes, "mc_emit");
  if (rc) return rc;
  if (n_vertices < 0 || n_triangles < 0 || (n_vertices > 0 && !vertices) || (n_triangles > 0 && !triangles)) {
    set_error("mc_emit: bad output arguments");
    return CNEUS_EINVAL;
  }
  if (n_vertices >= (1ll << 29) || n_triangles >= (1ll << 31)) { set_error("mc_emit: mesh too large for 32-bit slots"); return CNEUS_EUNSUPPORTED; }
  if (n_vertices == 0) return CNEUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  McGrid g{u, nx, ny, nz, (int64_t)nx * ny * nz, iso};
  McWs w;
  mc_ws_layout(g.n, ws, &w);
  mc_vertices_kernel<<<(unsigned)w.n_chunks, MC_THREADS, 0, st>>>(g, w.code, w.partial, w.slot, vertices);
  if (n_triangles > 0) mc_triangles_kernel<<<(unsigned)w.n_chunks, MC_THREADS, 0, st>>>(g, w.code, w.partial, w.slot, triangles);
  CNEUS_CUDA_CHECK(cudaGetLastError());
  count_launch(n_triangles > 0 ? 2 : 1);
  return CNEUS_OK;
}
