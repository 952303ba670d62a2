// Layout probe for the tcgen05 operand views used by the fused unit kernels (dstd_gcn_b200/csrc/unit_tc.cu).
// One "row image" format (csrc/umma.cuh) is read by the tensor core either as a K-major operand (rows = M/N, K = position)
// or as an MN-major operand (MN = position, K = rows).  Each case below is one contraction of the unit kernels with the
// shapes of the H3.6M encoder (C = 64, K = 22 -> 24, 3 frames per item), run as 3xTF32 with M = 64 accumulators, dumped
// from TMEM lane by lane and compared with an fp64 CPU result; `swap` tries the other LBO/SBO assignment for the MN-major
// descriptors so that one run settles the field semantics.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dstd_gcn_b200/csrc -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "umma.cuh"

using namespace dstd::umma;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

struct Case {
  const char* name;
  int RA, QA, a_mn, a_q0;   // A image: rows, positions; view; position offset (K offset if K-major, M offset if MN-major)
  int RB, QB, b_mn, b_q0;
  int M, N, K;              // instruction M (64 / 128), N, reduction length (multiple of 8)
  int Mvalid;               // rows of D that are checked
};

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ Asrc, const float* __restrict__ Bsrc, Case c,
                                                    int swap, float* __restrict__ dump, int* status) {
  extern __shared__ __align__(16) float smem[];
  const int sboA = img_sbo_f(c.QA), sboB = img_sbo_f(c.QB);
  const int fa = img_floats(c.RA, c.QA), fb = img_floats(c.RB, c.QB);
  float* a_hi = smem;
  float* a_lo = a_hi + fa;
  float* b_hi = a_lo + fa;
  float* b_lo = b_hi + fb;
  float* slack = b_lo + fb;                               // 4096 floats of zeros: out-of-window reads stay finite
  uint64_t* mbar = reinterpret_cast<uint64_t*>(slack + 4096);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * fa + 2 * fb + 4096; i += 128) smem[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < c.RA * c.QA; i += 128) {
    const int r = i / c.QA, q = i - r * c.QA;
    float hi, lo;
    split_tf32(Asrc[i], hi, lo);
    a_hi[img_off(r, q, sboA)] = hi;
    a_lo[img_off(r, q, sboA)] = lo;
  }
  for (int i = tid; i < c.RB * c.QB; i += 128) {
    const int r = i / c.QB, q = i - r * c.QB;
    float hi, lo;
    split_tf32(Bsrc[i], hi, lo);
    b_hi[img_off(r, q, sboB)] = hi;
    b_lo[img_off(r, q, sboB)] = lo;
  }
  if (tid == 0) { mbar_init(mbar, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc(slot, 128);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tm = *slot;
  if (tid == 0) {
    const uint32_t id = idesc_tf32(c.M, c.N, c.a_mn, c.b_mn);
    const uint32_t lbo = IMG_LBO_F * 4;
    auto mk = [&](const float* base, int sbo_f, int mn, int q0) {
      const uint32_t start = smem_u32(base) + (uint32_t)(q0 >> 2) * lbo;
      const uint32_t sbo = (uint32_t)sbo_f * 4;
      if (!mn) return desc(start, lbo, sbo);
      return swap ? desc(start, lbo, sbo) : desc(start, sbo, lbo);   // MN-major: LBO field = 8-row (K) group stride
    };
    const uint64_t dah = mk(a_hi, sboA, c.a_mn, c.a_q0), dal = mk(a_lo, sboA, c.a_mn, c.a_q0);
    const uint64_t dbh = mk(b_hi, sboB, c.b_mn, c.b_q0), dbl = mk(b_lo, sboB, c.b_mn, c.b_q0);
    const uint32_t ka = c.a_mn ? (uint32_t)sboA * 4 : 2 * lbo, kb = c.b_mn ? (uint32_t)sboB * 4 : 2 * lbo;
    for (int ks = 0; ks < c.K / 8; ++ks) {
      const uint64_t ah = desc_add(dah, ks * ka), al = desc_add(dal, ks * ka);
      const uint64_t bh = desc_add(dbh, ks * kb), bl = desc_add(dbl, ks * kb);
      mma_tf32(tm, ah, bh, id, ks > 0);
      mma_tf32(tm, ah, bl, id, 1);
      mma_tf32(tm, al, bh, id, 1);
    }
    commit(mbar);
  }
  if (!mbar_wait(mbar, 0) && tid == 0) *status = 1;
  fence_after();
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t r[8];
    tmem_ld8(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) dump[(warp * 32 + lane) * 256 + c0 + j] = __uint_as_float(r[j]);
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

int main() {
  const Case cases[] = {
      // name                           RA  QA a_mn q0  RB  QB b_mn q0   M    N   K  Mvalid
      {"B1  xf = W x        (A K, B MN)", 64, 64, 0, 0, 64, 72, 1, 0, 64, 72, 64, 64},
      {"B2  h = g xm^T   frame 1 (K, K)", 64, 72, 0, 24, 24, 24, 0, 0, 64, 24, 24, 64},
      {"B3  gxm = xf^T g frame 2 (MN,MN)", 64, 72, 1, 48, 64, 72, 1, 48, 64, 24, 64, 24},
      {"B4  gx = W^T h      (MN, MN)   ", 64, 64, 1, 0, 64, 72, 1, 0, 64, 72, 64, 64},
      {"B5  gW = h x^T      (K, K)     ", 64, 72, 0, 0, 64, 72, 0, 0, 64, 64, 72, 64},
      {"F2  out = xf xm  frame 1 (K,MN)", 64, 72, 0, 24, 24, 24, 1, 0, 64, 24, 24, 64},
      {"B1' M = 128        (A K, B MN) ", 128, 64, 0, 0, 64, 72, 1, 0, 128, 80, 64, 128},
      {"T2  temporal K=40 frame 1 (K,K)", 64, 80, 0, 40, 40, 40, 0, 0, 64, 40, 40, 64},
  };
  float *dA, *dB, *dDump;
  int* dS;
  CK(cudaMalloc(&dA, 128 * 128 * 4));
  CK(cudaMalloc(&dB, 128 * 128 * 4));
  CK(cudaMalloc(&dDump, 128 * 256 * 4));
  CK(cudaMalloc(&dS, 4));
  float* hA = (float*)malloc(128 * 128 * 4);
  float* hB = (float*)malloc(128 * 128 * 4);
  float* hD = (float*)malloc(128 * 256 * 4);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  srand(7);
  for (const Case& c : cases) {
    for (int i = 0; i < c.RA * c.QA; ++i) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (int i = 0; i < c.RB * c.QB; ++i) hB[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    CK(cudaMemcpy(dA, hA, c.RA * c.QA * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB, c.RB * c.QB * 4, cudaMemcpyHostToDevice));
    const int nswap = (c.a_mn || c.b_mn) ? 2 : 1;
    for (int swap = 0; swap < nswap; ++swap) {
      CK(cudaMemset(dDump, 0, 128 * 256 * 4));
      CK(cudaMemset(dS, 0, 4));
      const size_t smem = (size_t)(2 * img_floats(c.RA, c.QA) + 2 * img_floats(c.RB, c.QB) + 4096) * 4 + 64;
      probe_kernel<<<1, 128, smem>>>(dA, dB, c, swap, dDump, dS);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      int st = 0;
      CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hD, dDump, 128 * 256 * 4, cudaMemcpyDeviceToHost));
      // reference; try both row -> lane maps
      double err_m64 = 0, err_lin = 0, maxref = 0;
      for (int m = 0; m < c.Mvalid; ++m)
        for (int n = 0; n < c.N; ++n) {
          double s = 0;
          bool valid = true;
          for (int k = 0; k < c.K; ++k) {
            double a, b;
            if (!c.a_mn) a = hA[m * c.QA + c.a_q0 + k];
            else { if (c.a_q0 + m >= c.QA) { valid = false; break; } a = hA[k * c.QA + c.a_q0 + m]; }
            if (!c.b_mn) b = hB[n * c.QB + c.b_q0 + k];
            else { if (c.b_q0 + n >= c.QB) { valid = false; break; } b = hB[k * c.QB + c.b_q0 + n]; }
            s += a * b;
          }
          if (!valid) continue;
          if (fabs(s) > maxref) maxref = fabs(s);
          const double e1 = fabs(s - hD[m64_lane(m) * 256 + n]), e2 = fabs(s - hD[m * 256 + n]);
          if (e1 > err_m64) err_m64 = e1;
          if (e2 > err_lin) err_lin = e2;
        }
      printf("%s swap=%d status=%d  max|ref| %.3f  err(m64 lanes) %.3e  err(linear lanes) %.3e  -> %s\n", c.name, swap, st,
             maxref, err_m64, err_lin,
             (c.M == 64 ? err_m64 : err_lin) < 1e-4 ? "OK" : "MISMATCH");
    }
  }
  return 0;
}
