// Layout + accuracy probe for the bf16x3 operand scheme of the fused unit kernels (dstd_gcn_b200/csrc/unit_tc.cu).
// tf32 operands can only be K-major without swizzle (tools/umma_probe.cu: every MN-major tf32 case returns zeros; CUTLASS:
// "for mn-major tf32 operands, SW128_32B is the only available smem layout"), and the unit kernels need each activation
// tile in both orientations.  16-bit operands do have a no-swizzle MN-major form, so one 16-bit "row image" can serve
// both views.  fp32 accuracy comes from a three-way bf16 split x = h + m + l and the six products
// hh + hm + mh + mm + hl + lh (relative error ~2^-21, the same as the 3xTF32 scheme; same tensor time: K = 16 per MMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dstd_gcn_b200/csrc -o umma_probe16 tools/umma_probe16.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "umma.cuh"

using namespace dstd::umma;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

struct Case {
  const char* name;
  int RA, QA, a_mn, a_q0;   // A image: rows, positions (multiple of 8); view; position offset (K offset / M offset)
  int RB, QB, b_mn, b_q0;
  int M, N, K;              // instruction M, N; reduction length (multiple of 16; the source matrices are zero beyond Kreal)
  int Kreal, Mvalid;
};

__global__ void __launch_bounds__(128) probe16_kernel(const float* __restrict__ Asrc, const float* __restrict__ Bsrc, Case c,
                                                      int swap, int terms, float* __restrict__ dump, int* status) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int sboA = img16_sbo_b(c.QA), sboB = img16_sbo_b(c.QB);
  const int fa = img16_bytes(c.RA, c.QA), fb = img16_bytes(c.RB, c.QB);
  unsigned char* a_pl[3] = {smem, smem + fa, smem + 2 * fa};
  unsigned char* b_pl[3] = {smem + 3 * fa, smem + 3 * fa + fb, smem + 3 * fa + 2 * fb};
  unsigned char* slack = smem + 3 * fa + 3 * fb;          // 16 KB of zeros: out-of-window reads stay finite
  uint64_t* mbar = reinterpret_cast<uint64_t*>(slack + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (3 * fa + 3 * fb + 16384) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  for (int i = tid; i < c.RA * c.QA; i += 128) {
    const int r = i / c.QA, q = i - r * c.QA;
    uint32_t s[3];
    split_bf16x3(Asrc[i], s[0], s[1], s[2]);
    for (int p = 0; p < 3; ++p) *reinterpret_cast<unsigned short*>(a_pl[p] + img16_off_b(r, q, sboA)) = (unsigned short)(s[p] >> 16);
  }
  for (int i = tid; i < c.RB * c.QB; i += 128) {
    const int r = i / c.QB, q = i - r * c.QB;
    uint32_t s[3];
    split_bf16x3(Bsrc[i], s[0], s[1], s[2]);
    for (int p = 0; p < 3; ++p) *reinterpret_cast<unsigned short*>(b_pl[p] + img16_off_b(r, q, sboB)) = (unsigned short)(s[p] >> 16);
  }
  if (tid == 0) { mbar_init(mbar, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc(slot, 128);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tm = *slot;
  if (tid == 0) {
    const uint32_t id = idesc_bf16(c.M, c.N, c.a_mn, c.b_mn);
    const uint32_t lbo = IMG16_LBO_B;
    auto mk = [&](const unsigned char* base, int sbo, int mn, int q0) {
      const uint32_t start = smem_u32(base) + (uint32_t)(q0 >> 3) * lbo;
      if (!mn) return desc(start, lbo, (uint32_t)sbo);
      return swap ? desc(start, lbo, (uint32_t)sbo) : desc(start, (uint32_t)sbo, lbo);
    };
    uint64_t da[3], db[3];
    for (int p = 0; p < 3; ++p) {
      da[p] = mk(a_pl[p], sboA, c.a_mn, c.a_q0);
      db[p] = mk(b_pl[p], sboB, c.b_mn, c.b_q0);
    }
    const uint32_t ka = c.a_mn ? 2u * sboA : 2 * lbo, kb = c.b_mn ? 2u * sboB : 2 * lbo;
    const int ta[6] = {0, 0, 1, 1, 0, 2}, tb[6] = {0, 1, 0, 1, 2, 0};      // hh hm mh mm hl lh
    for (int ks = 0; ks < c.K / 16; ++ks)
      for (int t = 0; t < terms; ++t)
        mma_f16(tm, desc_add(da[ta[t]], ks * ka), desc_add(db[tb[t]], ks * kb), id, (ks > 0 || t > 0));
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
      // name                            RA  QA a_mn q0  RB  QB b_mn q0   M    N   K Kreal Mvalid
      {"B1  xf = W x        (A K, B MN) ", 64, 64, 0, 0, 64, 96, 1, 0, 64, 96, 64, 64, 64},
      {"B2  h = g xm^T  frame 1 (K, K)  ", 64, 96, 0, 24, 24, 32, 0, 0, 64, 24, 32, 24, 64},
      {"B3  gxm = xf^T g frame 2 (MN,MN)", 64, 96, 1, 48, 64, 96, 1, 48, 64, 24, 64, 64, 24},
      {"B4  gx = W^T h      (MN, MN)    ", 64, 64, 1, 0, 64, 96, 1, 0, 64, 96, 64, 64, 64},
      {"B5  gW = h x^T      (K, K)      ", 64, 96, 0, 0, 64, 96, 0, 0, 64, 64, 96, 96, 64},
      {"F2  out = xf xm frame 1 (K, MN) ", 64, 96, 0, 24, 32, 24, 1, 0, 64, 24, 32, 24, 64},
      {"B1' M = 128         (A K, B MN) ", 128, 64, 0, 0, 64, 96, 1, 0, 128, 96, 64, 64, 128},
      {"B4' M = 128 gx^T = h^T W (MN,MN)", 64, 128, 1, 0, 64, 64, 1, 0, 128, 64, 64, 64, 128},
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
  CK(cudaFuncSetAttribute(probe16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  srand(7);
  for (const Case& c : cases) {
    for (int i = 0; i < c.RA * c.QA; ++i) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (int i = 0; i < c.RB * c.QB; ++i) hB[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    // K-major B of the per-frame cases: the columns beyond Kreal are zero (the A window runs into the next frame there)
    if (!c.b_mn && c.Kreal < c.K)
      for (int r = 0; r < c.RB; ++r)
        for (int q = c.Kreal; q < c.QB; ++q) hB[r * c.QB + q] = 0.f;
    if (c.b_mn && c.Kreal < c.K)         // MN-major B: the K index is the row
      for (int r = c.Kreal; r < c.RB; ++r)
        for (int q = 0; q < c.QB; ++q) hB[r * c.QB + q] = 0.f;
    CK(cudaMemcpy(dA, hA, c.RA * c.QA * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB, c.RB * c.QB * 4, cudaMemcpyHostToDevice));
    const int nswap = (c.a_mn || c.b_mn) ? 2 : 1;
    for (int swap = 0; swap < nswap; ++swap)
      for (int terms = 3; terms <= 6; terms += 3) {
        CK(cudaMemset(dDump, 0, 128 * 256 * 4));
        CK(cudaMemset(dS, 0, 4));
        const size_t smem = (size_t)3 * img16_bytes(c.RA, c.QA) + 3 * img16_bytes(c.RB, c.QB) + 16384 + 64;
        probe16_kernel<<<1, 128, smem>>>(dA, dB, c, swap, terms, dDump, dS);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        int st = 0;
        CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hD, dDump, 128 * 256 * 4, cudaMemcpyDeviceToHost));
        double err_m64 = 0, err_lin = 0, maxref = 0;
        for (int m = 0; m < c.Mvalid; ++m)
          for (int n = 0; n < c.N; ++n) {
            double s = 0;
            bool valid = true;
            for (int k = 0; k < c.K; ++k) {
              double a, b;
              if (!c.a_mn) { if (c.a_q0 + k >= c.QA) { a = 0; } else a = hA[m * c.QA + c.a_q0 + k]; }
              else { if (c.a_q0 + m >= c.QA || k >= c.RA) { valid = false; break; } a = hA[k * c.QA + c.a_q0 + m]; }
              if (!c.b_mn) { if (c.b_q0 + k >= c.QB) { b = 0; } else b = hB[n * c.QB + c.b_q0 + k]; }
              else { if (c.b_q0 + n >= c.QB || k >= c.RB) { valid = false; break; } b = hB[k * c.QB + c.b_q0 + n]; }
              s += a * b;
            }
            if (!valid) continue;
            if (fabs(s) > maxref) maxref = fabs(s);
            const double e1 = fabs(s - hD[m64_lane(m) * 256 + n]), e2 = fabs(s - hD[m * 256 + n]);
            if (e1 > err_m64) err_m64 = e1;
            if (e2 > err_lin) err_lin = e2;
          }
        const double e = c.M == 64 ? err_m64 : err_lin;
        printf("%s swap=%d terms=%d status=%d  max|ref| %.3f  err %.3e (other lane map %.3e) -> %s\n", c.name, swap, terms,
               st, maxref, e, c.M == 64 ? err_lin : err_m64, e < (terms == 6 ? 2e-5 : 2e-3) ? "OK" : "MISMATCH");
      }
  }
  return 0;
}
