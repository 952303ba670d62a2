// tcgen05 (UMMA) / TMEM / mbarrier primitives shared by the fused unit kernels (unit_tc.cu) and the layout probe
// (tools/umma_probe.cu).  Everything here is sm_100a inline PTX; the layouts were validated on a B200 by the probe
// (profiles/r02_umma_probe.txt).
//
// "Row image": the one shared-memory operand format of the unit kernels.  A [rows][Q] fp32 matrix (rows = channels,
// Q = positions, Q % 4 == 0) stored as 8-row x 16-byte core matrices:
//     off(r, q) = (r / 8) * SBO + (q / 4) * LBO + (r % 8) * 4 + (q % 4)          [floats]
// with LBO = 36 floats (144 B: the 16-byte skew makes both the "thread = row" and the "lane = position group"
// 16-byte stores bank-conflict free) and SBO = (Q / 4) * LBO.  The same bytes serve two operand views:
//   K-major view  : operand rows (M or N) = image rows, K = q.      desc(LBO field = LBO, SBO field = SBO);
//                   8 more K = start + 2 * LBO; a frame starting at position q0: start + (q0 / 4) * LBO.
//   MN-major view : operand MN index = q (4 contiguous), K = image rows.  desc(LBO field = SBO, SBO field = LBO);
//                   8 more K = start + SBO; MN offset q0: start + (q0 / 4) * LBO.
// (canonical no-swizzle layouts: cute/atom/mma_traits_sm100.hpp, "make_umma_desc").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dstd {
namespace umma {

constexpr int IMG_LBO_F = 36;                       // floats between adjacent 4-position groups of a row image

__host__ __device__ __forceinline__ int img_sbo_f(int Q) { return (Q >> 2) * IMG_LBO_F; }
__host__ __device__ __forceinline__ int img_off(int r, int q, int sbo_f) {
  return (r >> 3) * sbo_f + (q >> 2) * IMG_LBO_F + ((r & 7) << 2) + (q & 3);
}
__host__ __device__ __forceinline__ int img_floats(int rows, int Q) { return ((rows + 7) >> 3) * img_sbo_f(Q); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle, sm_100 version bits
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// instruction descriptor kind::tf32, fp32 accumulate.  a_mn / b_mn: operand is MN-major
__host__ __device__ __forceinline__ uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(acc)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they retire
__device__ __forceinline__ void commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// bounded wait: a malformed pipeline must fail loudly, not hang the GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* mbar, uint32_t parity) {
  const uint32_t a = smem_u32(mbar);
  uint32_t done = 0;
  for (int it = 0; it < (1 << 22) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
  return done != 0;
}
// ---- TMA bulk copies (cp.async.bulk, SASS: UBLKCP): one thread moves a contiguous, 16-byte aligned block global -> shared;
// completion is counted in bytes on an mbarrier (arm it first with the total of all copies it will see)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 4 / 8 consecutive columns (issue several, then one tmem_ld_wait)
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TF32 truncation split on the full-rate logic pipe: x == hi + lo exactly, |lo| < 2^-10 |x|
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  split_tf32(v.x, hi.x, lo.x);
  split_tf32(v.y, hi.y, lo.y);
  split_tf32(v.z, hi.z, lo.z);
  split_tf32(v.w, hi.w, lo.w);
}


// ---- 16-bit row images (bf16 triple split, see split_bf16x3): [rows][Q] bf16, Q % 8 == 0, 8-row x 16-byte core matrices:
//     off16(r, q) = (r / 8) * SBO + (q / 8) * LBO16 + (r % 8) * 16 + (q % 8) * 2          [bytes]
// K-major view: rows = M/N, K = q (16 per MMA = 2 * LBO16).  MN-major view: MN = q (8 contiguous), K = rows (16 per MMA
// = 2 * SBO); descriptor fields swapped (LBO field = SBO, SBO field = LBO16).
constexpr int IMG16_LBO_B = 144;
__host__ __device__ __forceinline__ int img16_sbo_b(int Q) { return (Q >> 3) * IMG16_LBO_B; }
__host__ __device__ __forceinline__ int img16_off_b(int r, int q, int sbo_b) {
  return (r >> 3) * sbo_b + (q >> 3) * IMG16_LBO_B + ((r & 7) << 4) + ((q & 7) << 1);
}
__host__ __device__ __forceinline__ int img16_bytes(int rows, int Q) { return ((rows + 7) >> 3) * img16_sbo_b(Q); }

// instruction descriptor kind::f16 with bf16 operands, fp32 accumulate
__host__ __device__ __forceinline__ uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(acc)
      : "memory");
}

// The same instruction, to be executed by ALL lanes of a converged warp: one elected lane issues it.  The election lives
// inside the asm block, so the surrounding C++ stays free of divergent branches and ptxas keeps the descriptor
// arithmetic on the uniform datapath (with `if (elected) mma(...)` it sinks that arithmetic into the divergent region,
// computes it in vector registers and pays R2UR moves in front of every MMA).
__device__ __forceinline__ void mma_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(acc)
      : "memory");
}
// descriptors passed as 32-bit halves (the 64-bit values are assembled inside the asm block, so the compiler only ever
// sees 32-bit warp-uniform arithmetic); `leader` != 0 on the one lane that issues
__device__ __forceinline__ void mma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t acc, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %7, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo),
      "r"(b_hi), "r"(idesc), "r"(acc), "r"(leader)
      : "memory");
}
// kind::tf32 form of the same (descriptors as 32-bit halves, `leader` != 0 on the issuing lane)
__device__ __forceinline__ void mma_tf32_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t acc, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %7, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo),
      "r"(b_hi), "r"(idesc), "r"(acc), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* mbar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(mbar))
      : "memory");
}

// x == h + m + l up to 2^-21 |x| with h, m, l exactly representable in bf16 (truncation splits on the logic pipe; bf16
// has the fp32 exponent range, so nothing underflows).  Returned as the upper halves of fp32 bit patterns.
__device__ __forceinline__ void split_bf16x3(float x, uint32_t& h, uint32_t& m, uint32_t& l) {
  h = __float_as_uint(x) & 0xffff0000u;
  const float r1 = x - __uint_as_float(h);
  m = __float_as_uint(r1) & 0xffff0000u;
  const float r2 = r1 - __uint_as_float(m);
  l = __float_as_uint(r2) & 0xffff0000u;
}
// pack the bf16 (upper halves) of two split words: low half = element 0
__device__ __forceinline__ uint32_t pack_bf16(uint32_t e0, uint32_t e1) { return __byte_perm(e0, e1, 0x7632); }

// M = 64 accumulators (cta_group::1) occupy the lower 16 lanes of each 32-lane TMEM quarter:
//   row m -> lane (m % 16) + 32 * (m / 16)          (cute: tmem_frg, "half subpartitions layout atom")
__host__ __device__ __forceinline__ int m64_lane(int m) { return (m & 15) + ((m >> 4) << 5); }

}  // namespace umma
}  // namespace dstd
