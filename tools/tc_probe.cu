// tcgen05 building-block probe (development tool, not part of the library):
//   D[128 x N] (fp32, TMEM) = A[128 x K] * B[N x K]^T   with fp16 operands, K-major, SWIZZLE_128B slabs in shared
// memory brought in by cp.async.bulk from pre-swizzled global images; 3-pass hi/lo split checked against fp64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row groups 1024 B apart, rows 128 B apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address, 16-byte units
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int M = 128;
constexpr int KB = 64;  // one swizzle atom along K

// element (row, k) of a [rows x 64] fp16 slab -> index (in halfs) inside the SWIZZLE_128B image
__host__ __device__ inline int sw128_index(int row, int k) {
  int chunk = (k >> 3) ^ (row & 7);
  return (row >> 3) * 512 + (row & 7) * 64 + chunk * 8 + (k & 7);
}

// ---- variant 2: B in SWIZZLE_64B slabs of [N rows][32 halfs] (K=32 per slab), one N-wide MMA per k-step -------------
__host__ __device__ inline int sw64_index(int row, int k) {  // k in [0,32)
  int chunk = (k >> 3) ^ ((row >> 1) & 3);
  return (row >> 3) * 256 + (row & 7) * 32 + chunk * 8 + (k & 7);
}
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;   // 8 rows * 64 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;            // SWIZZLE_64B
  return d;
}
template <int N>
__global__ void __launch_bounds__(128) probe64_kernel(const __half* a_hi, const __half* a_lo, const __half* b_hi, const __half* b_lo,
                                                      int kblocks, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __half* sA_hi = (__half*)smem;
  __half* sA_lo = sA_hi + kblocks * M * KB;
  __half* sB_hi = sA_lo + kblocks * M * KB;         // [2*kblocks][N x 32]
  __half* sB_lo = sB_hi + kblocks * N * KB;
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t a_bytes = kblocks * M * KB * 2, b_bytes = kblocks * N * KB * 2;
    mbar_expect_tx(&bar_full, 2 * a_bytes + 2 * b_bytes);
    bulk_g2s(sA_hi, a_hi, a_bytes, &bar_full); bulk_g2s(sA_lo, a_lo, a_bytes, &bar_full);
    bulk_g2s(sB_hi, b_hi, b_bytes, &bar_full); bulk_g2s(sB_lo, b_lo, b_bytes, &bar_full);
    mbar_wait(&bar_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t first = 0;
    for (int kb = 0; kb < kblocks; ++kb)
      for (int k = 0; k < 4; ++k) {
        const int slab = kb * 2 + (k >> 1);              // K=32 slab index
        const uint32_t ka = k * 32, kbo = (k & 1) * 32;
        uint64_t dAh = make_desc_sw128(smem_u32(sA_hi + kb * M * KB) + ka), dAl = make_desc_sw128(smem_u32(sA_lo + kb * M * KB) + ka);
        uint64_t dBh = make_desc_sw64(smem_u32(sB_hi + slab * N * 32) + kbo), dBl = make_desc_sw64(smem_u32(sB_lo + slab * N * 32) + kbo);
        mma_f16(tmem, dAh, dBh, idesc, first);
        mma_f16(tmem + 256, dAl, dBh, idesc, first);
        mma_f16(tmem + 256, dAh, dBl, idesc, 1);
        first = 1;
      }
    mma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32], w[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
#define LD32(arr, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(arr[0]), "=r"(arr[1]), "=r"(arr[2]), "=r"(arr[3]), "=r"(arr[4]), "=r"(arr[5]), "=r"(arr[6]), "=r"(arr[7]), "=r"(arr[8]), "=r"(arr[9]), "=r"(arr[10]), "=r"(arr[11]), "=r"(arr[12]), "=r"(arr[13]), "=r"(arr[14]), "=r"(arr[15]), "=r"(arr[16]), "=r"(arr[17]), "=r"(arr[18]), "=r"(arr[19]), "=r"(arr[20]), "=r"(arr[21]), "=r"(arr[22]), "=r"(arr[23]), "=r"(arr[24]), "=r"(arr[25]), "=r"(arr[26]), "=r"(arr[27]), "=r"(arr[28]), "=r"(arr[29]), "=r"(arr[30]), "=r"(arr[31]) : "r"(addr))
    LD32(v, taddr);
    LD32(w, taddr + 256);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int j = 0; j < 32; ++j) out[row * N + c0 + j] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const __half* a_hi, const __half* a_lo, const __half* b_hi, const __half* b_lo,
                                                    int kblocks, int passes, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // swizzle atoms need 1024-B alignment
  __half* sA_hi = (__half*)smem;                    // [kblocks][128 x 64]
  __half* sA_lo = sA_hi + kblocks * M * KB;
  __half* sB_hi = sA_lo + kblocks * M * KB;         // [kblocks][N x 64]
  __half* sB_lo = sB_hi + kblocks * N * KB;
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (threadIdx.x == 0) {
    const uint32_t a_bytes = kblocks * M * KB * 2, b_bytes = kblocks * N * KB * 2;
    mbar_expect_tx(&bar_full, 2 * a_bytes + 2 * b_bytes);
    bulk_g2s(sA_hi, a_hi, a_bytes, &bar_full);
    bulk_g2s(sA_lo, a_lo, a_bytes, &bar_full);
    bulk_g2s(sB_hi, b_hi, b_bytes, &bar_full);
    bulk_g2s(sB_lo, b_lo, b_bytes, &bar_full);
    mbar_wait(&bar_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // instruction descriptor: D=f32, A=B=f16, both K-major, N, M=128
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t acc = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      for (int k = 0; k < KB / 16; ++k) {
        const uint32_t koff = k * 32;  // 16 halfs = 32 bytes inside the 128-byte swizzled row
        uint64_t dAh = make_desc_sw128(smem_u32(sA_hi + kb * M * KB) + koff);
        uint64_t dAl = make_desc_sw128(smem_u32(sA_lo + kb * M * KB) + koff);
        uint64_t dBh = make_desc_sw128(smem_u32(sB_hi + kb * N * KB) + koff);
        uint64_t dBl = make_desc_sw128(smem_u32(sB_lo + kb * N * KB) + koff);
        mma_f16(tmem, dAh, dBh, idesc, acc); acc = 1;
        if (passes >= 3) { mma_f16(tmem, dAl, dBh, idesc, 1); mma_f16(tmem, dAh, dBl, idesc, 1); }
        if (passes >= 4) mma_f16(tmem, dAl, dBl, idesc, 1);
      }
    }
    mma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // each warp reads its 32 lanes (rows), 32 columns at a time
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
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
    const int row = warp * 32 + lane;
    for (int j = 0; j < 32; ++j) out[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N>
int run(int kblocks, int passes, bool scale_w) {
  const int K = kblocks * KB;
  std::vector<float> A(M * K), B(N * K);
  srand(1234 + N + kblocks);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.0f;            // activation-like, >= 0
  for (auto& x : B) x = ((float)rand() / RAND_MAX - 0.5f) * 0.4f;   // weight-like
  const float wscale = scale_w ? 64.0f : 1.0f;
  std::vector<__half> ah(M * K), al(M * K), bh(N * K), bl(N * K);
  for (int r = 0; r < M; ++r)
    for (int k = 0; k < K; ++k) {
      float x = A[r * K + k];
      __half h = __float2half_rn(x);
      __half l = __float2half_rn(x - __half2float(h));
      int idx = (k / KB) * M * KB + sw128_index(r, k % KB);
      ah[idx] = h; al[idx] = l;
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      float x = B[n * K + k] * wscale;
      __half h = __float2half_rn(x);
      __half l = __float2half_rn(x - __half2float(h));
      int idx = (k / KB) * N * KB + sw128_index(n, k % KB);
      bh[idx] = h; bl[idx] = l;
    }
  __half *dah, *dal, *dbh, *dbl;
  float* dout;
  CK(cudaMalloc(&dah, ah.size() * 2)); CK(cudaMalloc(&dal, al.size() * 2));
  CK(cudaMalloc(&dbh, bh.size() * 2)); CK(cudaMalloc(&dbl, bl.size() * 2));
  CK(cudaMalloc(&dout, M * N * 4));
  CK(cudaMemcpy(dah, ah.data(), ah.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dal, al.data(), al.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbh, bh.data(), bh.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbl, bl.data(), bl.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, M * N * 4));
  size_t smem = (size_t)kblocks * (2 * M + 2 * N) * KB * 2 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<N><<<1, 128, smem>>>(dah, dal, dbh, dbl, kblocks, passes, dout);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> out(M * N);
  CK(cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < M; ++r)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[r * K + k] * (double)B[n * K + k];
      double got = (double)out[r * N + n] / wscale;
      maxerr = fmax(maxerr, fabs(got - ref));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("N=%d K=%d passes=%d wscale=%g : max|err|=%.3e  max|ref|=%.3e  rel=%.3e\n", N, K, passes, wscale, maxerr, maxref,
         maxerr / maxref);
  cudaFree(dah); cudaFree(dal); cudaFree(dbh); cudaFree(dbl); cudaFree(dout);
  return 0;
}

int run64(int kblocks) {
  const int N = 256, K = kblocks * KB;
  std::vector<float> A(M * K), B(N * K);
  srand(99 + kblocks);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.0f;
  for (auto& x : B) x = ((float)rand() / RAND_MAX - 0.5f) * 0.4f;
  const float wscale = 64.0f;
  std::vector<__half> ah(M * K), al(M * K), bh(N * K), bl(N * K);
  for (int r = 0; r < M; ++r)
    for (int k = 0; k < K; ++k) {
      float x = A[r * K + k]; __half h = __float2half_rn(x); __half l = __float2half_rn(x - __half2float(h));
      int idx = (k / KB) * M * KB + sw128_index(r, k % KB); ah[idx] = h; al[idx] = l;
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      float x = B[n * K + k] * wscale; __half h = __float2half_rn(x); __half l = __float2half_rn(x - __half2float(h));
      int idx = (k / 32) * N * 32 + sw64_index(n, k % 32); bh[idx] = h; bl[idx] = l;
    }
  __half *dah, *dal, *dbh, *dbl; float* dout;
  CK(cudaMalloc(&dah, ah.size() * 2)); CK(cudaMalloc(&dal, al.size() * 2)); CK(cudaMalloc(&dbh, bh.size() * 2)); CK(cudaMalloc(&dbl, bl.size() * 2));
  CK(cudaMalloc(&dout, M * N * 4));
  CK(cudaMemcpy(dah, ah.data(), ah.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dal, al.data(), al.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbh, bh.data(), bh.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbl, bl.data(), bl.size() * 2, cudaMemcpyHostToDevice));
  size_t smem = (size_t)kblocks * (2 * M + 2 * N) * KB * 2 + 1024;
  CK(cudaFuncSetAttribute(probe64_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe64_kernel<256><<<1, 128, smem>>>(dah, dal, dbh, dbl, kblocks, dout);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  std::vector<float> out(M * N);
  CK(cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < M; ++r)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[r * K + k] * (double)B[n * K + k];
      maxerr = fmax(maxerr, fabs((double)out[r * N + n] / wscale - ref)); maxref = fmax(maxref, fabs(ref));
    }
  printf("SW64-B N=256 K=%d split-acc : max|err|=%.3e max|ref|=%.3e rel=%.3e\n", K, maxerr, maxref, maxerr / maxref);
  return 0;
}

int main() {
  run64(1);
  run64(2);
  run<256>(1, 1, false);
  run<256>(1, 3, false);
  run<128>(2, 3, false);
  run<256>(2, 3, true);
  run<256>(2, 4, true);
  run<128>(3, 3, true);
  return 0;
}
