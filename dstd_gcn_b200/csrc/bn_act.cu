// Fused BatchNorm(C*V channels, statistics over N,T) + residual + PReLU + dropout mask, forward and backward
// (model/dstdgcn.py:44-50 BatchNorm.forward, :152-154 `x = bn(x); x += r; x = prelu(x)`, :306-308, :283-284).
//
// Work decomposition shared by all four kernels: CTA = (channel c, batch split s); a thread owns up to MAXJ fixed
// (t,v) positions of the T x V plane and walks the samples of its split, so the per-(c,v) BN parameters are
// per-thread constants and every access to a [T x V] plane is one contiguous, coalesced sweep.  When the producer and
// the consumer keep the plane in different memory orders (T-major vs V-major: this op is where the block switches
// between the spatial and the temporal unit layout) the plane is transposed through shared memory.
//
// Statistics use shifted sums (shift = y[0,c,0,v]) accumulated in fp32 per thread and combined in fp64, so the
// E[x^2]-E[x]^2 cancellation does not bite on millimetre-scale poses.
#include <stdlib.h>

#include "kernels.cuh"

namespace dstd {

constexpr int BN_THREADS_MAX = 512;
constexpr int BN_MAXJ = 8;       // positions per thread  -> T*V <= 4096
// samples in flight per thread (memory-level parallelism): 8 for the dataset shapes (T*V <= 1024), 4 beyond
#define BN_U 4
#define BN_UB ((NJ) <= 1 ? 8 : 4)
static int bn_u(int nj) { (void)nj; return 4; }
static int bn_ub(int nj) { return nj <= 1 ? 8 : 4; }

int bn_act_splits(int N, int C) {
  (void)C;
  static const int per_cta = [] {
    const char* e = getenv("DSTD_BN_SAMPLES_PER_CTA");     // tuning knob; default measured best on B200
    int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : v;
  }();
  int s = (N + per_cta - 1) / per_cta;
  if (s > 128) s = 128;
  if (s < 1) s = 1;
  return s;
}

struct BnGeom {
  int threads, nj;
};
static BnGeom bn_geom(int T, int V) {
  int tv = T * V;
  BnGeom g;
  int nj = cdiv(tv, BN_THREADS_MAX);
  g.nj = nj <= 1 ? 1 : nj <= 2 ? 2 : nj <= 4 ? 4 : 8;
  g.threads = cdiv(cdiv(tv, g.nj), 32) * 32;
  return g;
}

// position j (0..TV) in the memory order of view `v` -> logical (t, v)
__device__ __forceinline__ bool t_fastest(const View4& vw) { return !(vw.sk == 1 || vw.sp != 1); }
__device__ __forceinline__ void decode_pos(const View4& vw, int j, int T, int V, int& t, int& v) {
  if (!t_fastest(vw)) {   // V fastest (or generic): j = t*V + v
    t = j / V;
    v = j - t * V;
  } else {                // T fastest: j = v*T + t
    v = j / T;
    t = j - v * T;
  }
}
__device__ __forceinline__ bool same_order(const View4& a, const View4& b) { return t_fastest(a) == t_fastest(b); }
__device__ __forceinline__ int bn_pidx(int c, int v, int C, int V, int vc_order) { return vc_order ? v * C + c : c * V + v; }
__device__ __forceinline__ long long pos_off(const View4& vw, int t, int v) {
  return (long long)t * vw.sp + (long long)v * vw.sk;
}

struct BnFwdP {
  int N, C, T, V, vc_order, training, S;
  float eps, momentum;
  View4 y, r, out;
  const float *gamma, *beta;
  float *running_mean, *running_var;
  long long* nbt;
  const float* prelu;
  const float* mask;
  float *save_mean, *save_invstd;
  float* part;   // [S][C][V][2]
};

// stage BN_U planes of `src` (walked in its own memory order, coalesced) into sh[u][t*V+v]
template <int NJ, int U>
__device__ __forceinline__ void stage_planes(const View4& src, int c, int n, int n1, int T, int V, float* sh, int TV) {
  float tmp[U][NJ];
  int idx[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const int j = threadIdx.x + i * blockDim.x;
    idx[i] = -1;
    if (j < TV) {
      int t, v;
      decode_pos(src, j, T, V, t, v);
      idx[i] = t * V + v;
      const float* p = src.p + (long long)n * src.sn + (long long)c * src.sc + pos_off(src, t, v);
#pragma unroll
      for (int u = 0; u < U; ++u) tmp[u][i] = (n + u < n1) ? __ldg(p + (long long)u * src.sn) : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < NJ; ++i)
    if (idx[i] >= 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) sh[u * TV + idx[i]] = tmp[u][i];
    }
}

// ------------------------------------------------------------------------------------------ forward: statistics
template <int NJ>
__global__ void __launch_bounds__(BN_THREADS_MAX) bn_stats_kernel(BnFwdP q) {
  extern __shared__ float sh[];   // [2][T*V]
  const int c = blockIdx.x, s = blockIdx.y, T = q.T, V = q.V, TV = T * V;
  const int n0 = (int)((long long)q.N * s / q.S), n1 = (int)((long long)q.N * (s + 1) / q.S);
  float a1[NJ], a2[NJ];
  int idx[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    a1[i] = a2[i] = 0.f;
    idx[i] = -1;
    const int j = threadIdx.x + i * blockDim.x;
    if (j < TV) {
      int t, v;
      decode_pos(q.y, j, T, V, t, v);
      idx[i] = t * V + v;
      const float shift = __ldg(q.y.p + vix(q.y, 0, c, 0, v));
      const float* p = q.y.p + (long long)c * q.y.sc + pos_off(q.y, t, v);
      float s1 = 0.f, s2 = 0.f;
      for (int n = n0; n < n1; n += BN_U) {
        float d[BN_U];
#pragma unroll
        for (int u = 0; u < BN_U; ++u) d[u] = (n + u < n1) ? __ldg(p + (long long)(n + u) * q.y.sn) - shift : 0.f;
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
          s1 += d[u];
          s2 = fmaf(d[u], d[u], s2);
        }
      }
      a1[i] = s1;
      a2[i] = s2;
    }
  }
#pragma unroll
  for (int i = 0; i < NJ; ++i)
    if (idx[i] >= 0) {
      sh[idx[i]] = a1[i];
      sh[TV + idx[i]] = a2[i];
    }
  __syncthreads();
  if (threadIdx.x < V) {
    int v = threadIdx.x;
    double s1 = 0., s2 = 0.;
    for (int t = 0; t < T; ++t) {
      s1 += sh[t * V + v];
      s2 += sh[TV + t * V + v];
    }
    float* dst = q.part + (((long long)s * q.C + c) * V + v) * 2;
    dst[0] = (float)s1;
    dst[1] = (float)s2;
  }
}

__global__ void bn_finalize_kernel(BnFwdP q) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0 && q.nbt) *q.nbt += 1;
  if (idx >= q.C * q.V) return;
  const int c = idx / q.V, v = idx - c * q.V;
  double s1 = 0., s2 = 0.;
  for (int s = 0; s < q.S; ++s) {
    const float* p = q.part + (((long long)s * q.C + c) * q.V + v) * 2;
    s1 += p[0];
    s2 += p[1];
  }
  const double cnt = (double)q.N * q.T;
  const double shift = __ldg(q.y.p + vix(q.y, 0, c, 0, v));
  const double dm = s1 / cnt;
  double var = s2 / cnt - dm * dm;
  if (var < 0.) var = 0.;
  const double mean = shift + dm;
  const int pi = bn_pidx(c, v, q.C, q.V, q.vc_order);
  q.save_mean[pi] = (float)mean;
  q.save_invstd[pi] = (float)(1.0 / sqrt(var + (double)q.eps));
  if (q.running_mean) {
    const double unb = cnt > 1. ? var * cnt / (cnt - 1.) : var;
    q.running_mean[pi] = (float)((1.0 - q.momentum) * q.running_mean[pi] + q.momentum * mean);
    q.running_var[pi] = (float)((1.0 - q.momentum) * q.running_var[pi] + q.momentum * unb);
  }
}

__global__ void bn_eval_stats_kernel(BnFwdP q) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= q.C * q.V) return;
  q.save_mean[idx] = q.running_mean[idx];
  q.save_invstd[idx] = rsqrtf(q.running_var[idx] + q.eps);
}

// ------------------------------------------------------------------------------------------ forward: apply
// positions are walked in the memory order of `out`; y / r kept in the other order go through shared memory.
// BN_U samples are in flight per thread.
template <int NJ>
__global__ void __launch_bounds__(BN_THREADS_MAX) bn_apply_kernel(BnFwdP q) {
  extern __shared__ float sh[];   // [2][BN_U][T*V] staging (y, r)
  const int c = blockIdx.x, s = blockIdx.y, T = q.T, V = q.V, TV = T * V;
  const int n0 = (int)((long long)q.N * s / q.S), n1 = (int)((long long)q.N * (s + 1) / q.S);
  const bool y_stage = !same_order(q.y, q.out);
  const bool r_stage = q.r.p && !same_order(q.r, q.out);
  const float slope = q.prelu ? __ldg(q.prelu) : 1.f;
  float* shy = sh;
  float* shr = sh + BN_U * TV;
  int tt[NJ], vv[NJ];
  float sc[NJ], sf[NJ], sm[NJ];   // out = (y - mean) * (gamma * invstd) + beta  (centred first: no cancellation)
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    tt[i] = -1;
    vv[i] = 0;
    sc[i] = sf[i] = sm[i] = 0.f;
    const int j = threadIdx.x + i * blockDim.x;
    if (j < TV) {
      decode_pos(q.out, j, T, V, tt[i], vv[i]);
      const int pi = bn_pidx(c, vv[i], q.C, V, q.vc_order);
      const float g = __ldg(q.gamma + pi), is = q.save_invstd[pi];
      sm[i] = q.save_mean[pi];
      sc[i] = g * is;
      sf[i] = __ldg(q.beta + pi);
    }
  }
  for (int n = n0; n < n1; n += BN_U) {
    if (y_stage || r_stage) {
      __syncthreads();
      if (y_stage) stage_planes<NJ, BN_U>(q.y, c, n, n1, T, V, shy, TV);
      if (r_stage) stage_planes<NJ, BN_U>(q.r, c, n, n1, T, V, shr, TV);
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      if (tt[i] >= 0) {
        const int t = tt[i], v = vv[i], e = t * V + v;
        float yv[BN_U], rv[BN_U], mv[BN_U];
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
          const bool ok = n + u < n1;
          yv[u] = y_stage ? shy[u * TV + e] : (ok ? __ldg(q.y.p + vix(q.y, n + u, c, t, v)) : 0.f);
          rv[u] = !q.r.p ? 0.f : r_stage ? shr[u * TV + e] : (ok ? __ldg(q.r.p + vix(q.r, n + u, c, t, v)) : 0.f);
          mv[u] = (q.mask && ok) ? __ldg(q.mask + (((long long)(n + u) * q.C + c) * T + t) * V + v) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
          if (n + u < n1) {
            const float pre = fmaf(yv[u] - sm[i], sc[i], sf[i]) + rv[u];
            const float a = pre > 0.f ? pre : slope * pre;
            q.out.p[vix(q.out, n + u, c, t, v)] = a * mv[u];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ backward
struct BnBwdP {
  int N, C, T, V, vc_order, training, S;
  View4 y, r, gout, gy, gr;
  const float *gamma, *beta, *prelu, *mask, *save_mean, *save_invstd;
  float *ggamma, *gbeta, *gprelu;
  float* part;    // [S][C][V][2]
  float* part_p;  // [S][C]
};

// d(pre-activation) at one element; also returns xhat and the PReLU-slope contribution
__device__ __forceinline__ float bn_gpre(bool has_prelu, float yv, float rv, float gv, float mv, float mu, float is,
                                         float g, float b, float slope, float& xhat, float& gslope) {
  xhat = (yv - mu) * is;
  gv *= mv;
  gslope = 0.f;
  if (has_prelu) {
    const float pre = fmaf(xhat, g, b) + rv;
    if (pre <= 0.f) {
      gslope = gv * pre;
      gv *= slope;
    }
  }
  return gv;
}

// pass 1: per-(c,v) sums of gpre and gpre*xhat, PReLU slope gradient.  Positions in gout's memory order.
template <int NJ>
__global__ void __launch_bounds__(BN_THREADS_MAX) bn_bwd_reduce_kernel(BnBwdP q) {
  extern __shared__ float sh[];   // [2][BN_UB][T*V]
  __shared__ float red[32];
  const int c = blockIdx.x, s = blockIdx.y, T = q.T, V = q.V, TV = T * V;
  const int n0 = (int)((long long)q.N * s / q.S), n1 = (int)((long long)q.N * (s + 1) / q.S);
  const bool has_prelu = q.prelu != nullptr;
  const bool use_r = q.r.p && has_prelu;
  const bool y_stage = !same_order(q.y, q.gout);
  const bool r_stage = use_r && !same_order(q.r, q.gout);
  const float slope = has_prelu ? __ldg(q.prelu) : 1.f;
  float* shy = sh;
  float* shr = sh + BN_UB * TV;
  int tt[NJ], vv[NJ];
  float mu[NJ], is[NJ], g[NJ], b[NJ], a1[NJ], a2[NJ];
  float gsl = 0.f;
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    tt[i] = -1;
    vv[i] = 0;
    mu[i] = is[i] = g[i] = b[i] = a1[i] = a2[i] = 0.f;
    const int j = threadIdx.x + i * blockDim.x;
    if (j < TV) {
      decode_pos(q.gout, j, T, V, tt[i], vv[i]);
      const int pi = bn_pidx(c, vv[i], q.C, V, q.vc_order);
      mu[i] = __ldg(q.save_mean + pi);
      is[i] = __ldg(q.save_invstd + pi);
      g[i] = __ldg(q.gamma + pi);
      b[i] = __ldg(q.beta + pi);
    }
  }
  for (int n = n0; n < n1; n += BN_UB) {
    if (y_stage || r_stage) {
      __syncthreads();
      if (y_stage) stage_planes<NJ, BN_UB>(q.y, c, n, n1, T, V, shy, TV);
      if (r_stage) stage_planes<NJ, BN_UB>(q.r, c, n, n1, T, V, shr, TV);
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      if (tt[i] >= 0) {
        const int t = tt[i], v = vv[i], e = t * V + v;
        float yv[BN_UB], rv[BN_UB], gv[BN_UB], mv[BN_UB];
#pragma unroll
        for (int u = 0; u < BN_UB; ++u) {
          const bool ok = n + u < n1;
          yv[u] = y_stage ? shy[u * TV + e] : (ok ? __ldg(q.y.p + vix(q.y, n + u, c, t, v)) : 0.f);
          rv[u] = !use_r ? 0.f : r_stage ? shr[u * TV + e] : (ok ? __ldg(q.r.p + vix(q.r, n + u, c, t, v)) : 0.f);
          gv[u] = ok ? __ldg(q.gout.p + vix(q.gout, n + u, c, t, v)) : 0.f;
          mv[u] = (q.mask && ok) ? __ldg(q.mask + (((long long)(n + u) * q.C + c) * T + t) * V + v) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < BN_UB; ++u) {
          if (n + u < n1) {
            float xhat, gs;
            const float gp = bn_gpre(has_prelu, yv[u], rv[u], gv[u], mv[u], mu[i], is[i], g[i], b[i], slope, xhat, gs);
            a1[i] += gp;
            a2[i] = fmaf(gp, xhat, a2[i]);
            gsl += gs;
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    if (tt[i] >= 0) {
      sh[tt[i] * V + vv[i]] = a1[i];
      sh[TV + tt[i] * V + vv[i]] = a2[i];
    }
  }
  __syncthreads();
  if (threadIdx.x < V) {
    int v = threadIdx.x;
    double s1 = 0., s2 = 0.;
    for (int t = 0; t < T; ++t) {
      s1 += sh[t * V + v];
      s2 += sh[TV + t * V + v];
    }
    float* dst = q.part + (((long long)s * q.C + c) * V + v) * 2;
    dst[0] = (float)s1;
    dst[1] = (float)s2;
  }
  float tot = block_sum(gsl, red);
  if (threadIdx.x == 0) q.part_p[(long long)s * q.C + c] = tot;
}

__global__ void bn_bwd_finalize_kernel(BnBwdP q) {
  __shared__ float red[32];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < q.C * q.V) {
    const int c = idx / q.V, v = idx - c * q.V;
    double s1 = 0., s2 = 0.;
    for (int s = 0; s < q.S; ++s) {
      const float* p = q.part + (((long long)s * q.C + c) * q.V + v) * 2;
      s1 += p[0];
      s2 += p[1];
    }
    const int pi = bn_pidx(c, v, q.C, q.V, q.vc_order);
    q.gbeta[pi] = (float)s1;
    q.ggamma[pi] = (float)s2;
  }
  if (blockIdx.x == 0 && q.gprelu) {   // deterministic fixed-order sum of the slope partials
    float a = 0.f;
    for (int i = threadIdx.x; i < q.S * q.C; i += blockDim.x) a += q.part_p[i];
    float tot = block_sum(a, red);
    if (threadIdx.x == 0) q.gprelu[0] = tot;
  }
}

// pass 2: gy (and gr).  Positions in gy's memory order (= y's: gy is allocated like y).
template <int NJ>
__global__ void __launch_bounds__(BN_THREADS_MAX) bn_bwd_apply_kernel(BnBwdP q) {
  extern __shared__ float sh[];   // [3][BN_UB][T*V]: gout, r staging; gr transposition
  const int c = blockIdx.x, s = blockIdx.y, T = q.T, V = q.V, TV = T * V;
  const int n0 = (int)((long long)q.N * s / q.S), n1 = (int)((long long)q.N * (s + 1) / q.S);
  const bool has_prelu = q.prelu != nullptr;
  const bool use_r = q.r.p && has_prelu;
  const bool g_stage = !same_order(q.gout, q.gy);
  const bool r_stage = use_r && !same_order(q.r, q.gy);
  const bool gr_stage = q.gr.p && !same_order(q.gr, q.gy);
  const float slope = has_prelu ? __ldg(q.prelu) : 1.f;
  const float icnt = 1.0f / ((float)q.N * (float)T);
  float* shg = sh;
  float* shr = sh + BN_UB * TV;
  float* sho = sh + 2 * BN_UB * TV;
  int tt[NJ], vv[NJ];
  float mu[NJ], is[NJ], g[NJ], b[NJ], k1[NJ], k2[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    tt[i] = -1;
    vv[i] = 0;
    mu[i] = is[i] = g[i] = b[i] = k1[i] = k2[i] = 0.f;
    const int j = threadIdx.x + i * blockDim.x;
    if (j < TV) {
      decode_pos(q.gy, j, T, V, tt[i], vv[i]);
      const int pi = bn_pidx(c, vv[i], q.C, V, q.vc_order);
      mu[i] = __ldg(q.save_mean + pi);
      is[i] = __ldg(q.save_invstd + pi);
      g[i] = __ldg(q.gamma + pi);
      b[i] = __ldg(q.beta + pi);
      if (q.training) {
        k1[i] = q.gbeta[pi] * icnt;
        k2[i] = q.ggamma[pi] * icnt;
      }
    }
  }
  for (int n = n0; n < n1; n += BN_UB) {
    if (g_stage || r_stage || gr_stage) {
      __syncthreads();
      if (g_stage) stage_planes<NJ, BN_UB>(q.gout, c, n, n1, T, V, shg, TV);
      if (r_stage) stage_planes<NJ, BN_UB>(q.r, c, n, n1, T, V, shr, TV);
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      if (tt[i] >= 0) {
        const int t = tt[i], v = vv[i], e = t * V + v;
        float yv[BN_UB], rv[BN_UB], gv[BN_UB], mv[BN_UB];
#pragma unroll
        for (int u = 0; u < BN_UB; ++u) {
          const bool ok = n + u < n1;
          yv[u] = ok ? __ldg(q.y.p + vix(q.y, n + u, c, t, v)) : 0.f;
          rv[u] = !use_r ? 0.f : r_stage ? shr[u * TV + e] : (ok ? __ldg(q.r.p + vix(q.r, n + u, c, t, v)) : 0.f);
          gv[u] = g_stage ? shg[u * TV + e] : (ok ? __ldg(q.gout.p + vix(q.gout, n + u, c, t, v)) : 0.f);
          mv[u] = (q.mask && ok) ? __ldg(q.mask + (((long long)(n + u) * q.C + c) * T + t) * V + v) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < BN_UB; ++u) {
          if (n + u < n1) {
            float xhat, gs;
            const float gp = bn_gpre(has_prelu, yv[u], rv[u], gv[u], mv[u], mu[i], is[i], g[i], b[i], slope, xhat, gs);
            q.gy.p[vix(q.gy, n + u, c, t, v)] = g[i] * is[i] * (gp - k1[i] - xhat * k2[i]);
            if (q.gr.p) {
              if (gr_stage) sho[u * TV + e] = gp;
              else q.gr.p[vix(q.gr, n + u, c, t, v)] = gp;
            }
          }
        }
      }
    }
    if (gr_stage) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const int j = threadIdx.x + i * blockDim.x;
        if (j < TV) {
          int t, v;
          decode_pos(q.gr, j, T, V, t, v);
#pragma unroll
          for (int u = 0; u < BN_UB; ++u)
            if (n + u < n1) q.gr.p[vix(q.gr, n + u, c, t, v)] = sho[u * TV + t * V + v];
        }
      }
    }
  }
}

#define DSTD_BN_DISPATCH(kern, nj, grid, threads, smem, st, q)          \
  do {                                                                  \
    if ((smem) > 48 * 1024) {                                           \
      ensure_max_smem((const void*)kern<1>);                            \
      ensure_max_smem((const void*)kern<2>);                            \
      ensure_max_smem((const void*)kern<4>);                            \
      ensure_max_smem((const void*)kern<8>);                            \
    }                                                                   \
    switch (nj) {                                                       \
      case 1: kern<1><<<grid, threads, smem, st>>>(q); break;           \
      case 2: kern<2><<<grid, threads, smem, st>>>(q); break;           \
      case 4: kern<4><<<grid, threads, smem, st>>>(q); break;           \
      default: kern<8><<<grid, threads, smem, st>>>(q); break;          \
    }                                                                   \
  } while (0)

}  // namespace dstd

// =========================================================================================== C ABI
using namespace dstd;

extern "C" size_t dstd_bn_act_workspace_bytes(int N, int C, int T, int V) {
  int S = N > 128 ? 128 : N;   // upper bound of bn_act_splits
  if (S < 1) S = 1;
  return arena_need({(size_t)S * C * V * 2 * sizeof(float), (size_t)S * C * sizeof(float)});
}

extern "C" int dstd_bn_act_forward(const dstd_bn_act_fwd_args* a, dstd_stream_t stream) {
  DSTD_REQUIRE(a, DSTD_ERR_BAD_ARG, "bn_act_forward: null args");
  DSTD_REQUIRE(a->N > 0 && a->C > 0 && a->T > 0 && a->V > 0, DSTD_ERR_BAD_ARG, "bn_act_forward: bad dims");
  DSTD_REQUIRE(a->y.ptr && a->out.ptr && a->gamma && a->beta && a->save_mean && a->save_invstd, DSTD_ERR_BAD_ARG,
               "bn_act_forward: null tensor");
  DSTD_REQUIRE(a->T * a->V <= BN_MAXJ * BN_THREADS_MAX && a->V <= 32, DSTD_ERR_UNSUPPORTED,
               "bn_act: T*V=%d (max %d) or V=%d (max 32) outside the compiled limits", a->T * a->V,
               BN_MAXJ * BN_THREADS_MAX, a->V);
  DSTD_REQUIRE(a->training || (a->running_mean && a->running_var), DSTD_ERR_BAD_ARG,
               "bn_act_forward: eval mode needs running statistics");
  DSTD_REQUIRE(a->ws_bytes >= dstd_bn_act_workspace_bytes(a->N, a->C, a->T, a->V) && a->ws, DSTD_ERR_WORKSPACE,
               "bn_act_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  BnFwdP q;
  q.N = a->N; q.C = a->C; q.T = a->T; q.V = a->V;
  q.vc_order = a->vc_order; q.training = a->training;
  q.S = bn_act_splits(a->N, a->C);
  q.eps = a->eps; q.momentum = a->momentum;
  q.y = mk(a->y); q.r = mk(a->r); q.out = mk(a->out);
  q.gamma = a->gamma; q.beta = a->beta;
  q.running_mean = a->running_mean; q.running_var = a->running_var;
  q.nbt = a->num_batches_tracked;
  q.prelu = a->prelu; q.mask = a->mask;
  q.save_mean = a->save_mean; q.save_invstd = a->save_invstd;
  Arena ar(a->ws, a->ws_bytes);
  q.part = ar.take<float>((size_t)q.S * q.C * q.V * 2);
  BnGeom g = bn_geom(q.T, q.V);
  const size_t sm2 = (size_t)2 * q.T * q.V * sizeof(float);
  const size_t smu = sm2 * bn_u(g.nj);
  const int cv = q.C * q.V;
  if (q.training) {
    DSTD_BN_DISPATCH(bn_stats_kernel, g.nj, dim3(q.C, q.S), g.threads, sm2, st, q);
    count_launch();
    DSTD_LAUNCH_CHECK("bn_stats");
    bn_finalize_kernel<<<cdiv(cv, 128), 128, 0, st>>>(q);
    count_launch();
    DSTD_LAUNCH_CHECK("bn_finalize");
  } else {
    bn_eval_stats_kernel<<<cdiv(cv, 128), 128, 0, st>>>(q);
    count_launch();
    DSTD_LAUNCH_CHECK("bn_eval_stats");
  }
  DSTD_BN_DISPATCH(bn_apply_kernel, g.nj, dim3(q.C, q.S), g.threads, smu, st, q);
  count_launch();
  DSTD_LAUNCH_CHECK("bn_apply");
  return DSTD_OK;
}

extern "C" int dstd_bn_act_backward(const dstd_bn_act_bwd_args* a, dstd_stream_t stream) {
  DSTD_REQUIRE(a, DSTD_ERR_BAD_ARG, "bn_act_backward: null args");
  DSTD_REQUIRE(a->N > 0 && a->C > 0 && a->T > 0 && a->V > 0, DSTD_ERR_BAD_ARG, "bn_act_backward: bad dims");
  DSTD_REQUIRE(a->y.ptr && a->gout.ptr && a->gy.ptr && a->gamma && a->beta && a->save_mean && a->save_invstd &&
                   a->ggamma && a->gbeta,
               DSTD_ERR_BAD_ARG, "bn_act_backward: null tensor");
  DSTD_REQUIRE(!a->gr.ptr || a->r.ptr, DSTD_ERR_BAD_ARG, "bn_act_backward: gr requested without r");
  DSTD_REQUIRE(a->T * a->V <= BN_MAXJ * BN_THREADS_MAX && a->V <= 32, DSTD_ERR_UNSUPPORTED,
               "bn_act: T*V=%d or V=%d outside the compiled limits", a->T * a->V, a->V);
  DSTD_REQUIRE(a->ws_bytes >= dstd_bn_act_workspace_bytes(a->N, a->C, a->T, a->V) && a->ws, DSTD_ERR_WORKSPACE,
               "bn_act_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  BnBwdP q;
  q.N = a->N; q.C = a->C; q.T = a->T; q.V = a->V;
  q.vc_order = a->vc_order; q.training = a->training;
  q.S = bn_act_splits(a->N, a->C);
  q.y = mk(a->y); q.r = mk(a->r); q.gout = mk(a->gout); q.gy = mk(a->gy); q.gr = mk(a->gr);
  q.gamma = a->gamma; q.beta = a->beta; q.prelu = a->prelu; q.mask = a->mask;
  q.save_mean = a->save_mean; q.save_invstd = a->save_invstd;
  q.ggamma = a->ggamma; q.gbeta = a->gbeta; q.gprelu = a->prelu ? a->gprelu : nullptr;
  Arena ar(a->ws, a->ws_bytes);
  q.part = ar.take<float>((size_t)q.S * q.C * q.V * 2);
  q.part_p = ar.take<float>((size_t)q.S * q.C);
  BnGeom g = bn_geom(q.T, q.V);
  const size_t tv = (size_t)q.T * q.V * sizeof(float);
  DSTD_BN_DISPATCH(bn_bwd_reduce_kernel, g.nj, dim3(q.C, q.S), g.threads, 2 * bn_ub(g.nj) * tv, st, q);
  count_launch();
  DSTD_LAUNCH_CHECK("bn_bwd_reduce");
  bn_bwd_finalize_kernel<<<cdiv(q.C * q.V, 256), 256, 0, st>>>(q);
  count_launch();
  DSTD_LAUNCH_CHECK("bn_bwd_finalize");
  DSTD_BN_DISPATCH(bn_bwd_apply_kernel, g.nj, dim3(q.C, q.S), g.threads, 3 * bn_ub(g.nj) * tv, st, q);
  count_launch();
  DSTD_LAUNCH_CHECK("bn_bwd_apply");
  return DSTD_OK;
}
