// Fused adjacency aggregation + channel mix of one DSTD-GC unit, forward (model/dstdgcn.py:81 + :87 / :93 for every
// branch, summed as in DSTDGCB.forward :145-150 / :157-161, plus the layer skip :248).
//
// One CTA = (sample n, chunk of PCH frames p).  For each branch b:
//   xm_b[l][v][w] = alpha * pd[n,b,p0+l,v,w] + (adj*adj_w + adj_r)[v,w]             (read transposed when adj_t)
//   xa_b[c][l][w] = sum_v x[n,c,p0+l,v] xm_b[l][v][w]     c < Cin;   xa_b[Cin] = column sums of xm_b (carries the bias)
//   acc[o][l][w] += sum_{c<=Cin} wcat[o][b*(Cin+1)+c] xa_b[c][l][w]
// and finally out = acc (+ skip).  The aggregated tile xa (2x the activation tile for two branches) lives only in
// shared memory; HBM sees x once, pd once and out once.
//
// Thread mapping (256 threads):
//   aggregation : warp = (frame l, half of the w range), lane = channel pair (c, c+32); per v two conflict-free x loads
//                 and WH/4 broadcast float4 loads of the adjacency row feed 2*WH FMAs.
//   channel mix : warp = 8 consecutive output channels, lane = positions (lane + 32 i, i < TN); per reduction index two
//                 broadcast float4 loads of the weights and TN conflict-free loads of xa feed 8*TN FMAs.
#include "kernels.cuh"

namespace dstd {

template <int WH, int TN>
__global__ void __launch_bounds__(256) aggmix_fwd_kernel(AggMixParams q) {
  extern __shared__ __align__(16) float smem[];
  constexpr int KP = 2 * WH;          // padded adjacency row
  constexpr int XA_LD = 32 * TN + 1;  // odd: lane = channel writes and lane = position reads are both conflict free
  const int K = q.K, P = q.P, KK = K * K, Cin = q.Cin, C1 = Cin + 1, Cout = q.Cout, nb = q.nb, PCH = q.PCH;
  const int CoutP = q.CoutP;
  const int npos_max = PCH * K;
  const int XS_LD = npos_max | 1;
  float* xs = smem;                                   // [Cin][XS_LD]
  float* xms = xs + ((Cin * XS_LD + 3) & ~3);         // [nb][PCH][K][KP]
  float* xas = xms + nb * PCH * K * KP;               // [C1][XA_LD]
  float* ws = xas + ((C1 * XA_LD + 3) & ~3);          // [nb*C1][CoutP]
  float* aeff = ws + nb * C1 * CoutP;                 // [nb][K*K]
  // raw dynamic adjacency [nb][PCH][K*K]: dead after staging, so it borrows the xa tile when that is large enough
  float* pdr = (C1 * XA_LD >= nb * PCH * KK) ? xas : aeff + ((nb * KK + 3) & ~3);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.y, p0 = blockIdx.x * PCH;
  const int pv = min(PCH, P - p0);                    // valid frames in this chunk
  const int npos = pv * K;
  const float alpha = q.alpha ? __ldg(q.alpha) : 1.0f;

  // ---- stage with cp.async (every copy of a thread in flight at once): packed weights, input chunk, raw dynamic
  //      adjacency; the static adjacency A*W + R is formed meanwhile, then xm = alpha*pd + A_eff from shared memory
  {
    const int n4 = nb * C1 * CoutP / 4;
    for (int i = tid; i < n4; i += 256) cp_async16(ws + 4 * i, q.wcatT + 4 * i);
    int poff[TN];
#pragma unroll
    for (int i = 0; i < TN; ++i) {
      const int j = lane + 32 * i;
      const int l = j / K, k = j - l * K;
      poff[i] = j < npos ? (int)(l * q.x.sp + k * q.x.sk) : -1;
    }
    const float* xb = q.x.p + (long long)n * q.x.sn + (long long)p0 * q.x.sp;
    for (int c = warp; c < Cin; c += 8) {
#pragma unroll
      for (int i = 0; i < TN; ++i)
        if (poff[i] >= 0) cp_async4(xs + c * XS_LD + lane + 32 * i, xb + (long long)c * q.x.sc + poff[i], true);
    }
    for (int b = 0; b < nb; ++b) {
      const float* pdl = q.pd + ((long long)(n * nb + b) * P + p0) * KK;
      for (int i = tid; i < pv * KK; i += 256) cp_async4(pdr + b * PCH * KK + i, pdl + i, true);
    }
    for (int i = tid; i < nb * KK; i += 256) {
      const int b = i / KK, e = i - b * KK;
      float a = __ldg(q.adj[b] + e);
      if (q.adj_w[b]) a *= __ldg(q.adj_w[b] + e);
      if (q.adj_r[b]) a += __ldg(q.adj_r[b] + e);
      aeff[i] = a;
    }
    cp_async_wait_all();
    __syncthreads();
    for (int i = tid; i < nb * PCH * K * KP; i += 256) {
      int w = i % KP, t = i / KP;
      int v = t % K;
      t /= K;
      int l = t % PCH, b = t / PCH;
      float val = 0.f;
      if (w < K && l < pv) {
        const int e = q.adj_t ? (w * K + v) : (v * K + w);
        val = fmaf(alpha, pdr[(b * PCH + l) * KK + e], aeff[b * KK + e]);
      }
      xms[i] = val;
    }
  }
  __syncthreads();

  const int n_rg = (Cout + 63) / 64;
  for (int rg = 0; rg < n_rg; ++rg) {
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const int o0 = rg * 64 + warp * 8;

    for (int b = 0; b < nb; ++b) {
      if (rg == 0 || nb > 1) {
        // ---- aggregation of branch b into xas (recomputed per row group only when there are several branches)
        for (int item = warp; item < 2 * pv; item += 8) {
          const int l = item >> 1, half = item & 1;
          const float* xm_l = xms + ((b * PCH + l) * K) * KP + half * WH;
          for (int cg = 0; cg < Cin; cg += 64) {
            const int c0 = cg + lane, c1 = c0 + 32;
            const float* x0p = xs + min(c0, Cin - 1) * XS_LD + l * K;
            const float* x1p = xs + min(c1, Cin - 1) * XS_LD + l * K;
            float a0[WH], a1[WH];
#pragma unroll
            for (int j = 0; j < WH; ++j) a0[j] = a1[j] = 0.f;
            for (int v = 0; v < K; ++v) {
              const float x0 = x0p[v], x1 = x1p[v];
              const float4* r4 = reinterpret_cast<const float4*>(xm_l + v * KP);
#pragma unroll
              for (int j4 = 0; j4 < WH / 4; ++j4) {
                const float4 m = r4[j4];
                a0[j4 * 4 + 0] = fmaf(x0, m.x, a0[j4 * 4 + 0]);
                a0[j4 * 4 + 1] = fmaf(x0, m.y, a0[j4 * 4 + 1]);
                a0[j4 * 4 + 2] = fmaf(x0, m.z, a0[j4 * 4 + 2]);
                a0[j4 * 4 + 3] = fmaf(x0, m.w, a0[j4 * 4 + 3]);
                a1[j4 * 4 + 0] = fmaf(x1, m.x, a1[j4 * 4 + 0]);
                a1[j4 * 4 + 1] = fmaf(x1, m.y, a1[j4 * 4 + 1]);
                a1[j4 * 4 + 2] = fmaf(x1, m.z, a1[j4 * 4 + 2]);
                a1[j4 * 4 + 3] = fmaf(x1, m.w, a1[j4 * 4 + 3]);
              }
            }
            float* d0 = xas + c0 * XA_LD + l * K + half * WH;
            float* d1 = xas + c1 * XA_LD + l * K + half * WH;
#pragma unroll
            for (int j = 0; j < WH; ++j) {
              if (half * WH + j < K) {
                if (c0 < Cin) d0[j] = a0[j];
                if (c1 < Cin) d1[j] = a1[j];
              }
            }
          }
          // ones row: column sums of the adjacency (multiplies the conv_f bias)
          if (lane < WH && half * WH + lane < K) {
            float s = 0.f;
            for (int v = 0; v < K; ++v) s += xm_l[v * KP + lane];
            xas[Cin * XA_LD + l * K + half * WH + lane] = s;
          }
        }
        __syncthreads();
        if (q.xa && rg == 0) {   // optional: keep the aggregated tile for a backward pass that does not recompute it
          float* dst = q.xa + ((long long)(n * nb + b) * C1) * P * K + (long long)p0 * K;
          for (int i = tid; i < C1 * npos; i += 256) {
            int c = i / npos, j = i - c * npos;
            dst[(long long)c * P * K + j] = xas[c * XA_LD + j];
          }
        }
      }
      // ---- channel mix: acc[o][pos] += sum_j ws[b*C1+j][o] * xas[j][pos]
      if (o0 < Cout) {
        const float* wrow = ws + (b * C1) * CoutP + o0;
#pragma unroll 2
        for (int j = 0; j < C1; ++j) {
          const float4 wa = *reinterpret_cast<const float4*>(wrow + j * CoutP);
          const float4 wb = *reinterpret_cast<const float4*>(wrow + j * CoutP + 4);
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
          float xv[TN];
#pragma unroll
          for (int i = 0; i < TN; ++i) xv[i] = xas[j * XA_LD + lane + 32 * i];
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < TN; ++i) acc[r][i] = fmaf(wv[r], xv[i], acc[r][i]);
        }
      }
      if (b + 1 < nb || rg + 1 < n_rg) __syncthreads();   // xas is rewritten by the next aggregation
    }

    // ---- epilogue: out = acc (+ skip); lanes run along contiguous positions
    if (o0 < Cout) {
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        const int pos = lane + 32 * i;
        if (pos < npos) {
          const int l = pos / K, k = pos - l * K;
          const long long off_o = (long long)n * q.out.sn + (long long)(p0 + l) * q.out.sp + (long long)k * q.out.sk;
          const long long off_s = q.skip.p ? (long long)n * q.skip.sn + (long long)(p0 + l) * q.skip.sp + (long long)k * q.skip.sk : 0;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int o = o0 + r;
            if (o < Cout) {
              float v = acc[r][i];
              if (q.skip.p) v += __ldg(q.skip.p + off_s + (long long)o * q.skip.sc);
              q.out.p[off_o + (long long)o * q.out.sc] = v;
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ launch
struct AggMixGeom {
  int WH, TN, PCH;
  size_t smem;
};

static size_t aggmix_smem(int Cin, int CoutP, int K, int nb, int WH, int TN, int PCH) {
  const int C1 = Cin + 1, KP = 2 * WH, XS_LD = (PCH * K) | 1, XA_LD = 32 * TN + 1;
  size_t f = (size_t)((Cin * XS_LD + 3) & ~3) + (size_t)nb * PCH * K * KP + (size_t)((C1 * XA_LD + 3) & ~3) +
             (size_t)nb * C1 * CoutP + (size_t)((nb * K * K + 3) & ~3) +
             (C1 * XA_LD >= nb * PCH * K * K ? 0 : (size_t)nb * PCH * K * K);
  return f * sizeof(float);
}

static bool aggmix_geom(int Cin, int Cout, int P, int K, int nb, AggMixGeom& g) {
  if (K > 40 || K < 1) return false;
  g.WH = K <= 24 ? 12 : K <= 32 ? 16 : 20;
  const int CoutP = (Cout + 7) / 8 * 8;
  // the chunk with the best fill of the 32*TN position slots (TN <= 5) and of the 8 aggregation warps that fits
  double best = -1.;
  for (int tn = 3; tn <= 5; ++tn) {
    for (int pch = 1; pch <= 8 && pch <= P; ++pch) {
      if (pch * K > 32 * tn) break;
      size_t sm = aggmix_smem(Cin, CoutP, K, nb, g.WH, tn, pch);
      if (sm > (size_t)MAX_DYN_SMEM - 1024) continue;
      double fill = (double)(pch * K) / (32.0 * tn);
      int items = 2 * pch;
      double agg_eff = (double)items / (8.0 * ((items + 7) / 8));
      int chunks = (P + pch - 1) / pch;
      double tail = (double)P / (chunks * pch);
      double score = fill * (0.6 + 0.4 * agg_eff) * tail * (sm <= 113 * 1024 ? 1.0 : 0.85);
      if (score > best) {
        best = score;
        g.TN = tn;
        g.PCH = pch;
        g.smem = sm;
      }
    }
  }
  return best > 0.;
}

bool aggmix_supported(int Cin, int Cout, int P, int K, int nb) {
  AggMixGeom g;
  return aggmix_geom(Cin, Cout, P, K, nb, g);
}

int launch_aggmix_fwd(AggMixParams q, cudaStream_t st) {
  AggMixGeom g;
  DSTD_REQUIRE(aggmix_geom(q.Cin, q.Cout, q.P, q.K, q.nb, g), DSTD_ERR_UNSUPPORTED,
               "aggmix_fwd: Cin=%d Cout=%d K=%d outside the compiled tile limits", q.Cin, q.Cout, q.K);
  q.PCH = g.PCH;
  q.CoutP = (q.Cout + 7) / 8 * 8;
  dim3 grid(cdiv(q.P, g.PCH), q.N);
#define DSTD_AGGMIX(WH_, TN_)                                        \
  if (g.WH == WH_ && g.TN == TN_) {                                  \
    auto kern = aggmix_fwd_kernel<WH_, TN_>;                         \
    if (g.smem > 48 * 1024) ensure_max_smem((const void*)kern);      \
    kern<<<grid, 256, g.smem, st>>>(q);                              \
  }
  DSTD_AGGMIX(12, 3) DSTD_AGGMIX(12, 4) DSTD_AGGMIX(12, 5)
  DSTD_AGGMIX(16, 3) DSTD_AGGMIX(16, 4) DSTD_AGGMIX(16, 5)
  DSTD_AGGMIX(20, 3) DSTD_AGGMIX(20, 4) DSTD_AGGMIX(20, 5)
#undef DSTD_AGGMIX
  count_launch();
  return check_launch("aggmix_fwd");
}

}  // namespace dstd
