// Shape-generic kernels of the DSTD-GC unit for shapes outside the tile limits of the specialised kernels (P or K > 40):
// the stress configuration of BASELINE.json (T = 125 frames, 256 channels: spatial unit P = 125, K = 22; temporal unit
// P = 22, K = 125).  Same maths and the same buffers as dynadj.cu / aggregate.cu (model/dstdgcn.py:82-93 and its
// autograd), CUDA cores, tiled through shared memory, every reduction in a fixed order
// (deterministic).  The pairwise tanh tensor D (2750 KB per sample at the stress shape) is produced and consumed in
// (k', pair) tiles and never leaves the SM; the channel contraction of these shapes runs in bgemm / wgrad (gemm.cu).
//   limits: P <= 128, K <= 128.
#include "kernels.cuh"

namespace dstd {

constexpr int GEN_MAX = 128;

bool generic_supported(int P, int K) { return P >= 1 && K >= 1 && P <= GEN_MAX && K <= GEN_MAX; }

// ------------------------------------------------------------------------------------------ dynamic adjacency, forward
//   pd[n,b,p,e] = brm[p] + sum_{k'=(r,p')} Wrm[p][k'] tanh(m[n,b,r,p',v] - m[n,b,2+r,p',w]),  e = v*K + w
// CTA = (64-pair tile, b, n); thread = pair e, rows p = pg, pg + 4, ...  (pg = tid / 64)
constexpr int DG_TP = 64, DG_KC = 32;

__global__ void __launch_bounds__(256) dynadj_fwd_gen_kernel(DynAdjFwdParams q) {
  __shared__ float Ds[DG_KC][DG_TP];
  __shared__ float Ws[DG_KC][GEN_MAX];
  const int P = q.P, K = q.K, KK = K * K, P2 = 2 * P;
  const int tid = threadIdx.x, el = tid & 63, pg = tid >> 6;
  const int b = blockIdx.y, n = blockIdx.z, e0 = blockIdx.x * DG_TP;
  const float* mb = q.m + (long long)(n * q.nb + b) * 4 * P * K;
  const float* wrm = q.w_rm[b];
  float acc[GEN_MAX / 4];
#pragma unroll
  for (int j = 0; j < GEN_MAX / 4; ++j) acc[j] = 0.f;
  for (int k0 = 0; k0 < P2; k0 += DG_KC) {
    for (int i = tid; i < DG_KC * DG_TP; i += 256) {
      const int kk = i / DG_TP, ee = i - kk * DG_TP, kp = k0 + kk, e = e0 + ee;
      float d = 0.f;
      if (kp < P2 && e < KK) {
        const int r = kp / P, pp = kp - r * P, v = e / K, w = e - v * K;
        d = fast_tanh(__ldg(mb + (r * P + pp) * K + v) - __ldg(mb + ((2 + r) * P + pp) * K + w));
      }
      Ds[kk][ee] = d;
    }
    for (int i = tid; i < DG_KC * P; i += 256) {
      const int p = i / DG_KC, kk = i - p * DG_KC;
      Ws[kk][p] = (k0 + kk < P2) ? __ldg(wrm + (long long)p * P2 + k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < DG_KC; ++kk) {
      const float d = Ds[kk][el];
#pragma unroll
      for (int j = 0; j < GEN_MAX / 4; ++j)
        if (pg + 4 * j < P) acc[j] = fmaf(Ws[kk][pg + 4 * j], d, acc[j]);
    }
    __syncthreads();
  }
  const int e = e0 + el;
  if (e < KK) {
    float* dst = q.pd + (long long)(n * q.nb + b) * P * KK + e;
#pragma unroll
    for (int j = 0; j < GEN_MAX / 4; ++j) {
      const int p = pg + 4 * j;
      if (p < P) dst[(long long)p * KK] = acc[j] + __ldg(q.b_rm[b] + p);
    }
  }
}

int launch_dynadj_fwd_gen(const DynAdjFwdParams& q, cudaStream_t st) {
  DSTD_REQUIRE(generic_supported(q.P, q.K), DSTD_ERR_UNSUPPORTED, "dynadj_fwd_gen: P=%d K=%d outside limits (<= %d)", q.P, q.K, GEN_MAX);
  dim3 grid(cdiv(q.K * q.K, DG_TP), q.nb, q.N);
  dynadj_fwd_gen_kernel<<<grid, 256, 0, st>>>(q);
  count_launch();
  return check_launch("dynadj_fwd_gen");
}

// ------------------------------------------------------------------------------------------ dynamic adjacency, backward
// CTA = (b, n, tile of VT whole rows v with all their w).  With gP = alpha * gxm:
//   gD[k'][e] = sum_p Wrm[p][k'] gP[p][e] ;  gS = gD (1 - D^2)
//   gm1[r,p',v] = sum_w gS[(r,p'),v,w]                                       complete inside the tile -> gm
//   gm2[r,p',w] = - sum_v gS[(r,p'),v,w]                                     partial per tile -> gm2_part, summed below
//   gWrm[p][k'] = sum_e gP[p][e] D[k'][e]   gbrm[p] = sum_e gP[p][e]         -> partial slot (n, tile) (already x alpha)
//   gA_eff[e]   = sum_p gxm[p][e]                                            -> slot n (the tiles own disjoint e ranges)
//   galpha      = sum gxm pd                                                 -> partial slot (n, tile)
constexpr int DB_KC = 16, DB_TP = 128;

int dynadj_gen_tiles(int K) {
  const int VT = DB_TP / K > 0 ? DB_TP / K : 1;
  return (K + VT - 1) / VT;
}

__global__ void __launch_bounds__(256) dynadj_bwd_gen_kernel(DynAdjBwdParams q, float* __restrict__ gm2_part, int NT) {
  extern __shared__ __align__(16) float smem[];
  const int P = q.P, K = q.K, KK = K * K, P2 = 2 * P, P21 = P2 + 1;
  const int VT = max(1, DB_TP / K);                       // K <= 128: at least one row per tile
  constexpr int GLD = DB_TP + 1;
  float* gPs = smem;                         // [P][GLD]
  float* Ds = gPs + P * GLD;                 // [DB_KC][DB_TP]
  float* Gs = Ds + DB_KC * DB_TP;            // [DB_KC][DB_TP]
  float* red = Gs + DB_KC * DB_TP;           // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, n = blockIdx.y, tile = blockIdx.z;
  const int v0 = tile * VT, vt = min(VT, K - v0), tp = vt * K, e0 = v0 * K;
  const float alpha = q.alpha ? __ldg(q.alpha) : 1.0f;
  const float* mb = q.m + (long long)(n * q.nb + b) * 4 * P * K;
  const float* gx = q.gxm + (long long)(n * q.nb + b) * P * KK;
  const float* pdb = q.pd + (long long)(n * q.nb + b) * P * KK;
  const float* wrm = q.w_rm[b];
  float* gmb = q.gm + (long long)(n * q.nb + b) * 4 * P * K;
  const long long slot = ((long long)n * NT + tile) * q.nb + b;
  float* pw = q.part_wrm + slot * P * P21;
  float* pa = q.part_adj + ((long long)n * q.nb + b) * KK;
  float* g2 = gm2_part + (((long long)(n * q.nb + b)) * NT + tile) * P2 * K;
  float ga = 0.f;

  // a. gP tile, alpha gradient
  for (int i = tid; i < P * tp; i += 256) {
    const int p = i / tp, e = i - p * tp;
    const float g = __ldg(gx + (long long)p * KK + e0 + e);
    ga = fmaf(g, __ldg(pdb + (long long)p * KK + e0 + e), ga);
    gPs[p * GLD + e] = alpha * g;
  }
  __syncthreads();
  for (int e = tid; e < tp; e += 256) {          // static-adjacency gradient: the raw sum (alpha may be 0)
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += __ldg(gx + (long long)p * KK + e0 + e);
    pa[e0 + e] = s;
  }
  for (int p = warp; p < P; p += 8) {            // conv_rm bias gradient: row sums (fixed order: lanes then butterfly)
    float s = 0.f;
    for (int e = lane; e < tp; e += 32) s += gPs[p * GLD + e];
    s = warp_sum(s);
    if (lane == 0) pw[(long long)p * P21 + P2] = s;
  }
  for (int k0 = 0; k0 < P2; k0 += DB_KC) {
    // b. D chunk
    for (int i = tid; i < DB_KC * tp; i += 256) {
      const int kk = i / tp, e = i - kk * tp, kp = k0 + kk;
      float d = 0.f;
      if (kp < P2) {
        const int r = kp / P, pp = kp - r * P, vl = e / K, w = e - vl * K;
        d = fast_tanh(__ldg(mb + (r * P + pp) * K + v0 + vl) - __ldg(mb + ((2 + r) * P + pp) * K + w));
      }
      Ds[kk * DB_TP + e] = d;
    }
    __syncthreads();
    // c. gS = (Wrm^T gP) (1 - D^2)
    for (int i = tid; i < DB_KC * tp; i += 256) {
      const int kk = i / tp, e = i - kk * tp, kp = k0 + kk;
      float s = 0.f;
      if (kp < P2) {
        for (int p = 0; p < P; ++p) s = fmaf(__ldg(wrm + (long long)p * P2 + kp), gPs[p * GLD + e], s);
        const float d = Ds[kk * DB_TP + e];
        s *= 1.f - d * d;
      }
      Gs[kk * DB_TP + e] = s;
    }
    // d. gWrm[p][k'] = sum_e gP[p][e] D[k'][e]   (lanes = p: the GLD = 129 pitch keeps the row reads conflict free)
    for (int i = tid; i < DB_KC * P; i += 256) {
      const int kk = i / P, p = i - kk * P, kp = k0 + kk;
      if (kp < P2) {
        float s = 0.f;
        for (int e = 0; e < tp; ++e) s = fmaf(gPs[p * GLD + e], Ds[kk * DB_TP + e], s);
        pw[(long long)p * P21 + kp] = s;
      }
    }
    __syncthreads();
    // e. row / column sums of gS
    for (int i = tid; i < DB_KC * vt; i += 256) {
      const int kk = i / vt, vl = i - kk * vt, kp = k0 + kk;
      if (kp < P2) {
        float s = 0.f;
        for (int w = 0; w < K; ++w) s += Gs[kk * DB_TP + vl * K + w];
        const int r = kp / P, pp = kp - r * P;
        gmb[(r * P + pp) * K + v0 + vl] = s;
      }
    }
    for (int i = tid; i < DB_KC * K; i += 256) {
      const int kk = i / K, w = i - kk * K, kp = k0 + kk;
      if (kp < P2) {
        float s = 0.f;
        for (int vl = 0; vl < vt; ++vl) s += Gs[kk * DB_TP + vl * K + w];
        g2[kp * K + w] = s;
      }
    }
    __syncthreads();
  }
  ga = block_sum(ga, red);
  if (tid == 0) q.part_alpha[slot] = ga;
}

// gm2[n,b,r,p',w] = - sum over the v tiles of the partial column sums (fixed order)
__global__ void gm2_reduce_kernel(const float* __restrict__ gm2_part, float* __restrict__ gm, int NB, int NT, int P, int K) {
  const long long per = (long long)2 * P * K, total = (long long)NB * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long nb_ = i / per, r = i - nb_ * per;
    const float* src = gm2_part + nb_ * NT * per + r;
    float s = 0.f;
    for (int t = 0; t < NT; ++t) s += src[t * per];
    gm[nb_ * 2 * per + per + r] = -s;
  }
}

size_t dynadj_bwd_gen_ws_floats(int N, int nb, int P, int K) { return (size_t)N * nb * dynadj_gen_tiles(K) * 2 * P * K; }

int launch_dynadj_bwd_gen(const DynAdjBwdParams& q, float* gm2_part, cudaStream_t st) {
  DSTD_REQUIRE(generic_supported(q.P, q.K), DSTD_ERR_UNSUPPORTED, "dynadj_bwd_gen: P=%d K=%d outside limits (<= %d)", q.P, q.K, GEN_MAX);
  const int NT = dynadj_gen_tiles(q.K);
  DSTD_REQUIRE(q.S == q.N * NT && gm2_part, DSTD_ERR_BAD_ARG, "dynadj_bwd_gen: one partial slot per (sample, tile) expected");
  const size_t smem = ((size_t)q.P * (DB_TP + 1) + 2 * DB_KC * DB_TP + 32) * sizeof(float);
  ensure_max_smem((const void*)dynadj_bwd_gen_kernel);
  dim3 grid(q.nb, q.N, NT);
  dynadj_bwd_gen_kernel<<<grid, 256, smem, st>>>(q, gm2_part, NT);
  count_launch();
  DSTD_LAUNCH_CHECK("dynadj_bwd_gen");
  const long long total = (long long)q.N * q.nb * 2 * q.P * q.K;
  gm2_reduce_kernel<<<cdiv(total, 256) < 1184 ? cdiv(total, 256) : 1184, 256, 0, st>>>(gm2_part, q.gm, q.N * q.nb, NT, q.P, q.K);
  count_launch();
  return check_launch("gm2_reduce");
}

// ------------------------------------------------------------------------------------------ aggregation, forward
//   xa[n,b,c,p,w] = sum_v x[n,c,p,v] xmu_b[p,v,w]  (c < Cin; row Cin: x := 1)       CTA = (p, n); warp = channel, lanes = w
__global__ void __launch_bounds__(256) aggregate_fwd_gen_kernel(AggParams q) {
  extern __shared__ __align__(16) float smem[];
  const int K = q.K, P = q.P, KK = K * K, Cin = q.Cin, C1 = Cin + 1, KL = K + 1;
  float* xm = smem;                  // [K][KL]   xmu[v][w]
  float* xr = xm + K * KL;           // [8][K]    one input row per warp
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.x, n = blockIdx.y;
  const float alpha = q.alpha ? __ldg(q.alpha) : 1.0f;
  for (int b = 0; b < q.nb; ++b) {
    __syncthreads();
    const float* pdp = q.pd + ((long long)(n * q.nb + b) * P + p) * KK;
    for (int i = tid; i < KK; i += 256) {
      const int r = i / K, c = i - r * K;
      float a = __ldg(q.adj[b] + i);
      if (q.adj_w[b]) a *= __ldg(q.adj_w[b] + i);
      if (q.adj_r[b]) a += __ldg(q.adj_r[b] + i);
      const float val = fmaf(alpha, __ldg(pdp + i), a);
      if (q.adj_t) xm[c * KL + r] = val; else xm[r * KL + c] = val;
    }
    __syncthreads();
    for (int c = warp; c < C1; c += 8) {
      for (int v = lane; v < K; v += 32) xr[warp * K + v] = c < Cin ? __ldg(q.x.p + vix(q.x, n, c, p, v)) : 1.0f;
      __syncwarp();
      float acc[GEN_MAX / 32] = {0.f, 0.f, 0.f, 0.f};
      for (int v = 0; v < K; ++v) {
        const float xv = xr[warp * K + v];
#pragma unroll
        for (int i = 0; i < GEN_MAX / 32; ++i)
          if (lane + 32 * i < K) acc[i] = fmaf(xv, xm[v * KL + lane + 32 * i], acc[i]);
      }
      float* dst = q.xa + (((long long)(n * q.nb + b) * C1 + c) * P + p) * K;
#pragma unroll
      for (int i = 0; i < GEN_MAX / 32; ++i)
        if (lane + 32 * i < K) dst[lane + 32 * i] = acc[i];
      __syncwarp();
    }
  }
}

int launch_aggregate_fwd_gen(const AggParams& q, cudaStream_t st) {
  DSTD_REQUIRE(generic_supported(q.P, q.K), DSTD_ERR_UNSUPPORTED, "aggregate_fwd_gen: P=%d K=%d outside limits", q.P, q.K);
  const size_t smem = ((size_t)q.K * (q.K + 1) + 8 * q.K) * sizeof(float);
  ensure_max_smem((const void*)aggregate_fwd_gen_kernel);
  dim3 grid(q.P, q.N);
  aggregate_fwd_gen_kernel<<<grid, 256, smem, st>>>(q);
  count_launch();
  return check_launch("aggregate_fwd_gen");
}

// ------------------------------------------------------------------------------------------ aggregation, backward
//   gx[n,c,p,v]      = sum_b sum_w gxa_b[c,p,w] xmu_b[p,v,w]
//   gxmu_b[p,v,w]    = sum_{c <= Cin} xaug[c,p,v] gxa_b[c,p,w]          (accumulated in shared memory over channel chunks)
// CTA = (p, n).  gx accumulates over the branches in a shared tile of the CTA's channels chunk by chunk.
constexpr int AB_CC = 32;      // channels per chunk

__global__ void __launch_bounds__(256) aggregate_bwd_gen_kernel(AggParams q) {
  extern __shared__ __align__(16) float smem[];
  const int K = q.K, P = q.P, KK = K * K, Cin = q.Cin, C1 = Cin + 1, KL = K + 1;
  float* xm = smem;                       // [K][KL]     xmu_b[v][w]
  float* gxm = xm + K * KL;               // [K][KL]     gxmu_b[v][w]
  float* xs = gxm + K * KL;               // [AB_CC][K]  x rows (ones row included)
  float* gs = xs + AB_CC * K;             // [AB_CC][K]  gxa_b rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.x, n = blockIdx.y;
  const float alpha = q.alpha ? __ldg(q.alpha) : 1.0f;
  for (int b = 0; b < q.nb; ++b) {
    __syncthreads();
    const float* pdp = q.pd + ((long long)(n * q.nb + b) * P + p) * KK;
    for (int i = tid; i < KK; i += 256) {
      const int r = i / K, c = i - r * K;
      float a = __ldg(q.adj[b] + i);
      if (q.adj_w[b]) a *= __ldg(q.adj_w[b] + i);
      if (q.adj_r[b]) a += __ldg(q.adj_r[b] + i);
      const float val = fmaf(alpha, __ldg(pdp + i), a);
      if (q.adj_t) xm[c * KL + r] = val; else xm[r * KL + c] = val;
    }
    for (int i = tid; i < K * KL; i += 256) gxm[i] = 0.f;
    for (int c0 = 0; c0 < C1; c0 += AB_CC) {
      __syncthreads();
      for (int i = tid; i < AB_CC * K; i += 256) {
        const int cl = i / K, k = i - cl * K, c = c0 + cl;
        xs[i] = c < Cin ? __ldg(q.x.p + vix(q.x, n, c, p, k)) : (c == Cin ? 1.0f : 0.f);
        gs[i] = c < C1 ? __ldg(q.gxa + (((long long)(n * q.nb + b) * C1 + c) * P + p) * K + k) : 0.f;
      }
      __syncthreads();
      // gx rows of this chunk: warp = channel, lanes = v
      for (int cl = warp; cl < AB_CC; cl += 8) {
        const int c = c0 + cl;
        if (c < Cin) {
          float acc[GEN_MAX / 32] = {0.f, 0.f, 0.f, 0.f};
          for (int w = 0; w < K; ++w) {
            const float gv = gs[cl * K + w];
#pragma unroll
            for (int i = 0; i < GEN_MAX / 32; ++i)
              if (lane + 32 * i < K) acc[i] = fmaf(gv, xm[(lane + 32 * i) * KL + w], acc[i]);
          }
#pragma unroll
          for (int i = 0; i < GEN_MAX / 32; ++i) {
            const int v = lane + 32 * i;
            if (v < K) {
              float* d = q.gx.p + vix(q.gx, n, c, p, v);
              *d = (b > 0 ? *d : 0.f) + acc[i];
            }
          }
        }
      }
      // gxmu += xs^T gs over the chunk: thread = (v, w) pairs
      for (int e = tid; e < KK; e += 256) {
        const int v = e / K, w = e - v * K;
        float s = 0.f;
#pragma unroll 8
        for (int cl = 0; cl < AB_CC; ++cl) s = fmaf(xs[cl * K + v], gs[cl * K + w], s);
        gxm[v * KL + w] += s;
      }
    }
    __syncthreads();
    float* dst = q.gxm + ((long long)(n * q.nb + b) * P + p) * KK;
    for (int i = tid; i < KK; i += 256) {
      const int r = i / K, c = i - r * K;
      dst[i] = q.adj_t ? gxm[c * KL + r] : gxm[r * KL + c];
    }
  }
}

int launch_aggregate_bwd_gen(const AggParams& q, cudaStream_t st) {
  DSTD_REQUIRE(generic_supported(q.P, q.K), DSTD_ERR_UNSUPPORTED, "aggregate_bwd_gen: P=%d K=%d outside limits", q.P, q.K);
  const size_t smem = ((size_t)2 * q.K * (q.K + 1) + 2 * AB_CC * q.K) * sizeof(float);
  ensure_max_smem((const void*)aggregate_bwd_gen_kernel);
  dim3 grid(q.P, q.N);
  aggregate_bwd_gen_kernel<<<grid, 256, smem, st>>>(q);
  count_launch();
  return check_launch("aggregate_bwd_gen");
}

}  // namespace dstd
