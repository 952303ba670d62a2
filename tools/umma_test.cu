// Pilot for the tcgen05 (UMMA) channel-mix tile used by the fused aggregation/mix kernels: validates, in isolation,
//   - the shared-memory tile format  T[row k][32-wide MN panel], 128-byte rows, 128B swizzle (chunk ^= row % 8),
//   - the MN-major SWIZZLE_128B matrix descriptors (LBO = panel stride, SBO = 1024 B per 8 rows of K),
//   - the kind::tf32 instruction descriptor (M=128, N=64, both operands MN-major),
//   - 3xTF32 error compensation (hi*hi + hi*lo + lo*hi), TMEM alloc / commit / mbarrier / tcgen05.ld epilogue
// against an fp64 CPU result.    D[m][n] = sum_k A[m][k] B[n][k],  m < 128, n < 64, k < KD (multiple of 8)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_test tools/umma_test.cu && ./umma_test
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int M = 128, N = 64, KD = 72;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row k, column mn) inside a tile of `rows` K-rows: panel-major, 128 B rows, 128B swizzle
__host__ __device__ __forceinline__ int tile_off_floats(int k, int mn, int rows) {
  const int panel = mn >> 5, c = (mn >> 2) & 7, e = mn & 3;
  return panel * rows * 32 + k * 32 + ((c ^ (k & 7)) << 2) + e;
}

// K-major, no swizzle ("interleave"): 8x16B core matrices, element (r,k) at (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4
__host__ __device__ __forceinline__ int tile_off_kmajor(int k, int r, int kd) {
  return (r >> 3) * (kd / 4) * 32 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ uint64_t make_desc_k_none(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128) umma_test_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                        float* __restrict__ D, int mode, int layout, int* status,
                                                        uint32_t* dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned carve-up
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* a_hi = (float*)base;                       // [4 panels][KD][32]
  float* a_lo = a_hi + 4 * KD * 32;
  float* b_hi = a_lo + 4 * KD * 32;                 // [2 panels][KD][32]
  float* b_lo = b_hi + 2 * KD * 32;
  uint64_t* mbar = (uint64_t*)(b_lo + 2 * KD * 32);
  uint32_t* tmem_slot = (uint32_t*)(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < M * KD; i += 128) {
    const int m = i / KD, k = i - m * KD;
    const float x = A[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const int o = layout == 0 ? tile_off_floats(k, m, KD) : tile_off_kmajor(k, m, KD);
    a_hi[o] = hi;
    a_lo[o] = x - hi;
  }
  for (int i = tid; i < N * KD; i += 128) {
    const int n = i / KD, k = i - n * KD;
    const float x = B[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const int o = layout == 0 ? tile_off_floats(k, n, KD) : tile_off_kmajor(k, n, KD);
    b_hi[o] = hi;
    b_lo[o] = x - hi;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // generic-proxy writes of the operands -> visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  if (tid == 0) dbg[0] = tmem_d;
  if (mode == 0) {
    // TMEM store/load self test: lane L, column c <- 1000*L + c
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)(1000 * tid + j));
    const uint32_t ta = tmem_d + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(ta), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }

  if (tid == 0 && mode != 0) {
    // instruction descriptor: D=F32, A=B=TF32, both MN-major, N=64, M=128
    const uint32_t mn = layout == 0 ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (mn << 15) | (mn << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    uint64_t dah, dal, dbh, dbl;
    uint32_t kstep;
    if (layout == 0) {
      const uint32_t lbo = KD * 128, sbo = 1024;
      dah = make_desc_mn_sw128(smem_u32(a_hi), lbo, sbo); dal = make_desc_mn_sw128(smem_u32(a_lo), lbo, sbo);
      dbh = make_desc_mn_sw128(smem_u32(b_hi), lbo, sbo); dbl = make_desc_mn_sw128(smem_u32(b_lo), lbo, sbo);
      kstep = 1024;
    } else {
      const uint32_t lbo = 128, sbo = (KD / 4) * 128;
      dah = make_desc_k_none(smem_u32(a_hi), lbo, sbo); dal = make_desc_k_none(smem_u32(a_lo), lbo, sbo);
      dbh = make_desc_k_none(smem_u32(b_hi), lbo, sbo); dbl = make_desc_k_none(smem_u32(b_lo), lbo, sbo);
      kstep = 256;
    }
    dbg[1] = idesc; dbg[2] = (uint32_t)dah; dbg[3] = (uint32_t)(dah >> 32); dbg[4] = smem_u32(a_hi);
    uint32_t acc = 0;
    for (int ks = 0; ks < KD / 8; ++ks) {
      const uint64_t adv = (uint64_t)((ks * kstep) >> 4);
      umma_tf32(tmem_d, dah + adv, dbh + adv, idesc, acc);
      acc = 1;
      if (mode == 3) {
        umma_tf32(tmem_d, dah + adv, dbl + adv, idesc, 1);
        umma_tf32(tmem_d, dal + adv, dbh + adv, idesc, 1);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
  }
  // bounded wait on the MMA-complete barrier (phase 0)
  if (mode != 0) {
    uint32_t done = 0;
    for (int it = 0; it < 2000000 && !done; ++it) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(mbar)), "r"(0u)
          : "memory");
    }
    if (!done && tid == 0) *status = 1;   // timed out: report instead of hanging
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows m), 64 columns in two 32-column loads
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 32);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int m = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)(half * 32 + j) * M + m] = __uint_as_float(r[j]);   // D stored [n][m]
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_d) : "memory");
}

int main() {
  float *hA = (float*)malloc(M * KD * 4), *hB = (float*)malloc(N * KD * 4), *hD = (float*)malloc(M * N * 4);
  srand(1);
  for (int i = 0; i < M * KD; ++i) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (int i = 0; i < N * KD; ++i) hB[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  int* dS;
  CK(cudaMalloc(&dA, M * KD * 4));
  CK(cudaMalloc(&dB, N * KD * 4));
  CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, hA, M * KD * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB, N * KD * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(2 * 4 + 2 * 2) * KD * 32 * 4 + 1024 + 64;
  CK(cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  uint32_t* dDbg;
  CK(cudaMalloc(&dDbg, 64));
  const int modes[5][2] = {{0, 0}, {1, 1}, {3, 1}, {1, 0}, {3, 0}};
  for (int t = 0; t < 5; ++t) {
    const int mode = modes[t][0], layout = modes[t][1];
    CK(cudaMemset(dD, 0, M * N * 4));
    CK(cudaMemset(dS, 0, 4));
    CK(cudaMemset(dDbg, 0, 64));
    umma_test_kernel<<<1, 128, smem>>>(dA, dB, dD, mode, layout, dS, dDbg);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int st = 0;
    uint32_t dbg[8];
    CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dbg, dDbg, 32, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost));
    if (mode == 0) {
      int bad = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < 32; ++n)
          if (hD[(size_t)n * M + m] != (float)(1000 * m + n)) ++bad;
      printf("TMEM st/ld self test: %d mismatches of %d   tmem_base=0x%08x  D[n=3][m=5]=%f\n", bad, M * 32, dbg[0],
             hD[(size_t)3 * M + 5]);
      continue;
    }
    double maxerr = 0, maxref = 0;
    int bad_m = -1, bad_n = -1;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < KD; ++k) s += (double)hA[m * KD + k] * (double)hB[n * KD + k];
        double e = fabs(s - (double)hD[(size_t)n * M + m]);
        if (e > maxerr) { maxerr = e; bad_m = m; bad_n = n; }
        if (fabs(s) > maxref) maxref = fabs(s);
      }
    printf("layout %s mode %dxTF32: status %d  max abs err %.3e (m=%d n=%d)  max|ref| %.3f  D00=%f D[5][7]=%f  "
           "tmem=0x%08x idesc=0x%08x adesc=0x%08x%08x smem_a=0x%x\n", layout == 0 ? "MN-SW128" : "K-none  ", mode, st,
           maxerr, bad_m, bad_n, maxref, hD[0], hD[(size_t)7 * M + 5], dbg[0], dbg[1], dbg[3], dbg[2], dbg[4]);
  }
  return 0;
}
