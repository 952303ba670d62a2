// Adjacency aggregation of one DSTD-GC unit (model/dstdgcn.py:87 / :93, the einsum -> bmm of the reference)
// applied BEFORE the channel mix:
//
//   xm_b[p,v,w] = alpha * pd[n,b,p,v,w] + (adj_b*adj_w_b + adj_r_b)[v,w]         (transposed when adj_t)
//   xa[n,b,c,p,w] = sum_v x[n,c,p,v] xm_b[p,v,w]      c < Cin ;   xa[n,b,Cin,p,w] = sum_v xm_b[p,v,w]
//
// One CTA = (sample n, chunk of PCH "frames" p).  A warp owns one frame; a lane owns two channels (c, c+32):
// it keeps both K-long input rows and both K-long output rows in registers and streams the K x K matrix from
// shared memory with warp-broadcast float4 loads.  Backward: gx with the transposed matrix (same scheme), and the
// K x K outer-product reduction over channels for gxm in 4x4 register tiles.
#include "kernels.cuh"

namespace dstd {

__device__ __forceinline__ float aeff_at(const AggParams& q, int b, int i) {
  float a = __ldg(q.adj[b] + i);
  if (q.adj_w[b]) a *= __ldg(q.adj_w[b] + i);
  if (q.adj_r[b]) a += __ldg(q.adj_r[b] + i);
  return a;
}

static int agg_pch(int K, int P) {
  int pch = 100 / K;
  if (pch < 1) pch = 1;
  if (pch > 8) pch = 8;
  if (pch > P) pch = P;
  return pch;
}
static int agg_kmax(int K) { return K <= 24 ? 24 : K <= 28 ? 28 : K <= 36 ? 36 : 40; }

bool aggregate_supported(int Cin, int P, int K) { return K <= 40 && K >= 1 && P >= 1 && Cin >= 1; }

// ================================================================================= forward
template <int KMAX>
__global__ void __launch_bounds__(256) aggregate_fwd_kernel(AggParams q, int PCH, int RS) {
  extern __shared__ __align__(16) float smem[];
  const int K = q.K, P = q.P, KK = K * K, Cin = q.Cin, C1 = Cin + 1;
  float* xm_s = smem;                               // [nb][PCH][K][KMAX]
  float* xs = xm_s + q.nb * PCH * K * KMAX;         // [Cin][RS]
  float* out_s = xs + Cin * RS;                     // [C1][RS]
  const int tid = threadIdx.x, lane = tid & 31, pl = tid >> 5, nthr = blockDim.x;
  const int n = blockIdx.y, p0 = blockIdx.x * PCH;
  const int pv = min(PCH, P - p0);                  // valid frames in this chunk
  const float alpha = q.alpha ? __ldg(q.alpha) : 1.0f;

  for (int i = tid; i < q.nb * PCH * K * KMAX; i += nthr) {
    int w = i % KMAX, t = i / KMAX;
    int v = t % K;
    t /= K;
    int l = t % PCH, b = t / PCH;
    float val = 0.f;
    if (w < K && l < pv) {
      int e = q.adj_t ? (w * K + v) : (v * K + w);
      val = alpha * __ldg(q.pd + ((long long)(n * q.nb + b) * P + p0 + l) * KK + e) + aeff_at(q, b, e);
    }
    xm_s[i] = val;
  }
  const int J = PCH * K;
  for (int i = tid; i < Cin * J; i += nthr) {
    int c = i / J, j = i - c * J;
    int l = j / K, k = j - l * K;
    xs[c * RS + j] = (l < pv) ? __ldg(q.x.p + vix(q.x, n, c, p0 + l, k)) : 0.f;
  }
  __syncthreads();

  for (int b = 0; b < q.nb; ++b) {
    const float* xmb = xm_s + (b * PCH + pl) * K * KMAX;
    for (int cg = 0; cg < Cin; cg += 64) {
      const int c0 = cg + lane, c1 = c0 + 32;
      float x0[KMAX], x1[KMAX], a0[KMAX], a1[KMAX];
#pragma unroll
      for (int v = 0; v < KMAX; ++v) {
        x0[v] = (v < K && c0 < Cin) ? xs[c0 * RS + pl * K + v] : 0.f;
        x1[v] = (v < K && c1 < Cin) ? xs[c1 * RS + pl * K + v] : 0.f;
        a0[v] = 0.f;
        a1[v] = 0.f;
      }
#pragma unroll
      for (int v = 0; v < KMAX; ++v) {
        if (v < K) {
          const float4* r4 = reinterpret_cast<const float4*>(xmb + v * KMAX);
#pragma unroll
          for (int w4 = 0; w4 < KMAX / 4; ++w4) {
            float4 m = r4[w4];
            a0[w4 * 4 + 0] = fmaf(x0[v], m.x, a0[w4 * 4 + 0]);
            a0[w4 * 4 + 1] = fmaf(x0[v], m.y, a0[w4 * 4 + 1]);
            a0[w4 * 4 + 2] = fmaf(x0[v], m.z, a0[w4 * 4 + 2]);
            a0[w4 * 4 + 3] = fmaf(x0[v], m.w, a0[w4 * 4 + 3]);
            a1[w4 * 4 + 0] = fmaf(x1[v], m.x, a1[w4 * 4 + 0]);
            a1[w4 * 4 + 1] = fmaf(x1[v], m.y, a1[w4 * 4 + 1]);
            a1[w4 * 4 + 2] = fmaf(x1[v], m.z, a1[w4 * 4 + 2]);
            a1[w4 * 4 + 3] = fmaf(x1[v], m.w, a1[w4 * 4 + 3]);
          }
        }
      }
#pragma unroll
      for (int w = 0; w < KMAX; ++w) {
        if (w < K) {
          if (c0 < Cin) out_s[c0 * RS + pl * K + w] = a0[w];
          if (c1 < Cin) out_s[c1 * RS + pl * K + w] = a1[w];
        }
      }
    }
    if (lane < K) {      // ones row: column sums of xm
      float s = 0.f;
      for (int v = 0; v < K; ++v) s += xmb[v * KMAX + lane];
      out_s[Cin * RS + pl * K + lane] = s;
    }
    if (lane + 32 < K) {
      float s = 0.f;
      for (int v = 0; v < K; ++v) s += xmb[v * KMAX + lane + 32];
      out_s[Cin * RS + pl * K + lane + 32] = s;
    }
    __syncthreads();
    float* dst = q.xa + ((long long)(n * q.nb + b) * C1) * P * K + (long long)p0 * K;
    const int Jv = pv * K;
    for (int i = tid; i < C1 * Jv; i += nthr) {
      int c = i / Jv, j = i - c * Jv;
      dst[(long long)c * P * K + j] = out_s[c * RS + j];
    }
    __syncthreads();
  }
}

template <int KMAX>
static int agg_fwd_launch(const AggParams& q, cudaStream_t st) {
  int PCH = agg_pch(q.K, q.P);
  int RS = PCH * q.K;
  if ((RS & 1) == 0) RS += 1;
  size_t smem = ((size_t)q.nb * PCH * q.K * KMAX + (size_t)q.Cin * RS + (size_t)(q.Cin + 1) * RS + 4) * sizeof(float);
  // shrink the chunk until it fits
  while (smem > 200 * 1024 && PCH > 1) {
    --PCH;
    RS = PCH * q.K;
    if ((RS & 1) == 0) RS += 1;
    smem = ((size_t)q.nb * PCH * q.K * KMAX + (size_t)q.Cin * RS + (size_t)(q.Cin + 1) * RS + 4) * sizeof(float);
  }
  DSTD_REQUIRE(smem <= 220 * 1024, DSTD_ERR_UNSUPPORTED, "aggregate_fwd: Cin=%d K=%d needs %zu B shared memory", q.Cin,
               q.K, smem);
  auto kern = aggregate_fwd_kernel<KMAX>;
  if (smem > 48 * 1024) ensure_max_smem((const void*)kern);
  dim3 grid(cdiv(q.P, PCH), q.N);
  kern<<<grid, 32 * PCH, smem, st>>>(q, PCH, RS);
  count_launch();
  return check_launch("aggregate_fwd");
}

int launch_aggregate_fwd(const AggParams& q, cudaStream_t st) {
  DSTD_REQUIRE(aggregate_supported(q.Cin, q.P, q.K), DSTD_ERR_UNSUPPORTED, "aggregate: K=%d outside tile limits (K<=40)",
               q.K);
  switch (agg_kmax(q.K)) {
    case 24: return agg_fwd_launch<24>(q, st);
    case 28: return agg_fwd_launch<28>(q, st);
    case 36: return agg_fwd_launch<36>(q, st);
    default: return agg_fwd_launch<40>(q, st);
  }
}

// ================================================================================= backward
// gx[n,c,p,v]    = sum_b sum_w gxa[n,b,c,p,w] xmu_b[p,v,w]            (c < Cin)
// gxmu_b[p,v,w]  = sum_{c<=Cin} xaug[n,c,p,v] gxa[n,b,c,p,w]          (xaug row Cin == 1)
// where xmu = xm (or its transpose when adj_t); gxm is stored un-transposed.
template <int KMAX>
__global__ void __launch_bounds__(256) aggregate_bwd_kernel(AggParams q, int PCH) {
  extern __shared__ __align__(16) float smem[];
  const int K = q.K, P = q.P, KK = K * K, Cin = q.Cin, C1 = Cin + 1;
  float* xmT = smem;                                 // [nb][PCH][K(w)][KMAX(v)]  = xmu[v][w]
  float* xs = xmT + q.nb * PCH * K * KMAX;           // [C1][PCH][KMAX]   (row Cin = ones) ; reused for the gx tile
  float* gs = xs + C1 * PCH * KMAX;                  // [nb][C1][PCH][KMAX]
  const int tid = threadIdx.x, lane = tid & 31, pl = tid >> 5, nthr = blockDim.x;
  const int n = blockIdx.y, p0 = blockIdx.x * PCH;
  const int pv = min(PCH, P - p0);
  const float alpha = q.alpha ? __ldg(q.alpha) : 1.0f;

  for (int i = tid; i < q.nb * PCH * K * KMAX; i += nthr) {
    int v = i % KMAX, t = i / KMAX;
    int w = t % K;
    t /= K;
    int l = t % PCH, b = t / PCH;
    float val = 0.f;
    if (v < K && l < pv) {
      int e = q.adj_t ? (w * K + v) : (v * K + w);
      val = alpha * __ldg(q.pd + ((long long)(n * q.nb + b) * P + p0 + l) * KK + e) + aeff_at(q, b, e);
    }
    xmT[i] = val;
  }
  for (int i = tid; i < C1 * PCH * KMAX; i += nthr) {
    int k = i % KMAX, t = i / KMAX;
    int l = t % PCH, c = t / PCH;
    float val = 0.f;
    if (k < K && l < pv) val = (c == Cin) ? 1.0f : __ldg(q.x.p + vix(q.x, n, c, p0 + l, k));
    xs[i] = val;
  }
  for (int i = tid; i < q.nb * C1 * PCH * KMAX; i += nthr) {
    int k = i % KMAX, t = i / KMAX;
    int l = t % PCH;
    t /= PCH;
    int c = t % C1, b = t / C1;
    float val = 0.f;
    if (k < K && l < pv)
      val = __ldg(q.gxa + (((long long)(n * q.nb + b) * C1 + c) * P + p0 + l) * K + k);
    gs[i] = val;
  }
  __syncthreads();

  // ---- (B) gxm: 4x4 register tiles over (v,w), reduction over channels
  constexpr int NG = KMAX / 4;
  const int tiles = q.nb * PCH * NG * NG;
  for (int tile = tid; tile < tiles; tile += nthr) {
    int wg = tile % NG, t = tile / NG;
    int vg = t % NG;
    t /= NG;
    int l = t % PCH, b = t / PCH;
    if (l >= pv) continue;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* xp = xs + l * KMAX + vg * 4;
    const float* gp = gs + ((long long)b * C1 * PCH + l) * KMAX + wg * 4;
    for (int c = 0; c < C1; ++c) {
      float4 xv = *reinterpret_cast<const float4*>(xp + (long long)c * PCH * KMAX);
      float4 gv = *reinterpret_cast<const float4*>(gp + (long long)c * PCH * KMAX);
      float xa_[4] = {xv.x, xv.y, xv.z, xv.w};
      float ga_[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa_[i], ga_[j], acc[i][j]);
    }
    float* dst = q.gxm + ((long long)(n * q.nb + b) * P + p0 + l) * KK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int v = vg * 4 + i;
      if (v >= K) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int w = wg * 4 + j;
        if (w < K) dst[q.adj_t ? (w * K + v) : (v * K + w)] = acc[i][j];
      }
    }
  }

  // ---- (A) gx: warp = frame, lane = channel pair, all branches accumulated in registers
  for (int cg = 0; cg < Cin; cg += 64) {
    const int c0 = cg + lane, c1 = c0 + 32;
    float a0[KMAX], a1[KMAX];
#pragma unroll
    for (int v = 0; v < KMAX; ++v) a0[v] = a1[v] = 0.f;
    for (int b = 0; b < q.nb; ++b) {
      const float* xmb = xmT + (b * PCH + pl) * K * KMAX;
      const float* g0p = gs + (((long long)b * C1 + min(c0, Cin)) * PCH + pl) * KMAX;
      const float* g1p = gs + (((long long)b * C1 + min(c1, Cin)) * PCH + pl) * KMAX;
      float g0[KMAX], g1[KMAX];
#pragma unroll
      for (int w4 = 0; w4 < KMAX / 4; ++w4) {
        float4 t0 = *reinterpret_cast<const float4*>(g0p + w4 * 4);
        float4 t1 = *reinterpret_cast<const float4*>(g1p + w4 * 4);
        g0[w4 * 4] = t0.x; g0[w4 * 4 + 1] = t0.y; g0[w4 * 4 + 2] = t0.z; g0[w4 * 4 + 3] = t0.w;
        g1[w4 * 4] = t1.x; g1[w4 * 4 + 1] = t1.y; g1[w4 * 4 + 2] = t1.z; g1[w4 * 4 + 3] = t1.w;
      }
#pragma unroll
      for (int w = 0; w < KMAX; ++w) {
        if (w < K) {
          const float4* r4 = reinterpret_cast<const float4*>(xmb + w * KMAX);
#pragma unroll
          for (int v4 = 0; v4 < KMAX / 4; ++v4) {
            float4 m = r4[v4];
            a0[v4 * 4 + 0] = fmaf(g0[w], m.x, a0[v4 * 4 + 0]);
            a0[v4 * 4 + 1] = fmaf(g0[w], m.y, a0[v4 * 4 + 1]);
            a0[v4 * 4 + 2] = fmaf(g0[w], m.z, a0[v4 * 4 + 2]);
            a0[v4 * 4 + 3] = fmaf(g0[w], m.w, a0[v4 * 4 + 3]);
            a1[v4 * 4 + 0] = fmaf(g1[w], m.x, a1[v4 * 4 + 0]);
            a1[v4 * 4 + 1] = fmaf(g1[w], m.y, a1[v4 * 4 + 1]);
            a1[v4 * 4 + 2] = fmaf(g1[w], m.z, a1[v4 * 4 + 2]);
            a1[v4 * 4 + 3] = fmaf(g1[w], m.w, a1[v4 * 4 + 3]);
          }
        }
      }
    }
    __syncthreads();     // everyone is done reading xs (phase B / previous group) before it is overwritten
    if (pl < pv) {
#pragma unroll
      for (int v4 = 0; v4 < KMAX / 4; ++v4) {
        if (c0 < Cin)
          *reinterpret_cast<float4*>(xs + ((long long)(c0 - cg) * PCH + pl) * KMAX + v4 * 4) =
              make_float4(a0[v4 * 4], a0[v4 * 4 + 1], a0[v4 * 4 + 2], a0[v4 * 4 + 3]);
        if (c1 < Cin)
          *reinterpret_cast<float4*>(xs + ((long long)(c1 - cg) * PCH + pl) * KMAX + v4 * 4) =
              make_float4(a1[v4 * 4], a1[v4 * 4 + 1], a1[v4 * 4 + 2], a1[v4 * 4 + 3]);
      }
    }
    __syncthreads();
    const int cn = min(64, Cin - cg);
    for (int i = tid; i < cn * pv * K; i += nthr) {
      int k = i % K, t = i / K;
      int l = t % pv, c = t / pv;
      q.gx.p[vix(q.gx, n, cg + c, p0 + l, k)] = xs[((long long)c * PCH + l) * KMAX + k];
    }
  }
}

template <int KMAX>
static int agg_bwd_launch(const AggParams& q, cudaStream_t st) {
  int PCH = agg_pch(q.K, q.P);
  auto need = [&](int pch) {
    return ((size_t)q.nb * pch * q.K * KMAX + (size_t)(q.Cin + 1) * pch * KMAX +
            (size_t)q.nb * (q.Cin + 1) * pch * KMAX + 4) * sizeof(float);
  };
  while (need(PCH) > 100 * 1024 && PCH > 1) --PCH;
  size_t smem = need(PCH);
  DSTD_REQUIRE(smem <= 220 * 1024, DSTD_ERR_UNSUPPORTED, "aggregate_bwd: Cin=%d K=%d needs %zu B shared memory", q.Cin,
               q.K, smem);
  auto kern = aggregate_bwd_kernel<KMAX>;
  if (smem > 48 * 1024) ensure_max_smem((const void*)kern);
  dim3 grid(cdiv(q.P, PCH), q.N);
  kern<<<grid, 32 * PCH, smem, st>>>(q, PCH);
  count_launch();
  return check_launch("aggregate_bwd");
}

int launch_aggregate_bwd(const AggParams& q, cudaStream_t st) {
  DSTD_REQUIRE(aggregate_supported(q.Cin, q.P, q.K), DSTD_ERR_UNSUPPORTED, "aggregate: K=%d outside tile limits (K<=40)",
               q.K);
  switch (agg_kmax(q.K)) {
    case 24: return agg_bwd_launch<24>(q, st);
    case 28: return agg_bwd_launch<28>(q, st);
    case 36: return agg_bwd_launch<36>(q, st);
    default: return agg_bwd_launch<40>(q, st);
  }
}

}  // namespace dstd
