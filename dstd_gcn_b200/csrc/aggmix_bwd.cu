// Fused backward of (adjacency aggregation + channel mix) of one DSTD-GC unit: everything between the unit's output
// gradient and the gradient of the dynamic adjacency (Appendix A of SURVEY.md; reference autograd of
// model/dstdgcn.py:81,87,93 for all branches of DSTDGCB.forward :145-150 / :157-161).
//
// Persistent CTAs walk the work items (sample n, chunk of PCH frames).  Per item and branch b, with the frame chunk
// of x, gout and the adjacency xm_b = alpha*pd + A_eff staged in shared memory (frames padded to KP = 4*ceil(K/4) so
// that every row / frame is float4 aligned):
//   (a) gxa_b[j][pos]   = sum_o wcat[o][b,j] gout[o][pos]                    j <= Cin (row Cin: the conv_f bias)
//   (b) xa_b [j][pos]   = sum_v x[j][l,v] xmu_b[l][v][w]   (recomputed, never stored in HBM)
//   (c) gWf_b[o][j]    += sum_pos gout[o][pos] xa_b[j][pos]                  register accumulators for the whole CTA life
//   (d) gx[c][l,v]     += sum_w gxa_b[c][l,w] xmu_b[l][v][w]
//   (e) gxmu_b[l][v][w] = sum_{c<=Cin} xaug[c][l,v] gxa_b[c][l,w]  -> gxm in HBM (consumed by the dynadj backward)
// HBM traffic per sample: x, gout read once, gx written once, pd read / gxm written once (K/ C of a tile each).
// The per-CTA weight-gradient partials are summed by reduce_segments (deterministic, fixed order).
#include "kernels.cuh"
#include "umma.cuh"

namespace dstd {

// ---- legacy tensor path (mma.sync m16n8k8 TF32, operands in registers) with 3xTF32 error compensation.  The two channel
// GEMMs of this kernel were limited by the shared-memory return path on CUDA cores (every operand value costs a
// 4-byte-per-lane delivery: 11 deliveries per 24 FMAs at the 8x3 register tile the 128-register budget allows); with
// mma.sync each delivered value feeds 64-128 MACs.  tcgen05 would need hi/lo operand images in shared memory that do
// not fit next to the frame-chunk tiles.
__device__ __forceinline__ void split3(float x, uint32_t& hi, uint32_t& lo) {
  // hi = x truncated to TF32 (one full-rate LOP3; cvt.rna.tf32 runs on the quarter-rate conversion pipe and these
  // kernels issue ~20 splits per k-step), lo = x - hi exactly; |lo| < 2^-10 |x|, so hi*hi + hi*lo + lo*hi is good to ~2^-20
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int AMB_NT = 512;   // 16 warps; the two channel GEMMs are split 2-way along their reduction dimension

template <int KP, int TN>
__global__ void __launch_bounds__(AMB_NT, 1) aggmix_bwd_kernel(AggMixBwdParams q) {
  extern __shared__ __align__(16) float smem[];
  constexpr int WH = (KP / 2 + 3) / 4 * 4;   // half of a padded adjacency row, multiple of 4
  constexpr int KP2 = 2 * WH;
  const int K = q.K, P = q.P, KK = K * K, Cin = q.Cin, C1 = Cin + 1, Cout = q.Cout, nb = q.nb, PCH = q.PCH;
  const int LD = q.LD, CinP = q.CinP;
  const int npos_pad = PCH * KP;
  const int C1R8 = (C1 + 7) & ~7;
  float* xs = smem;                         // [C1R8][LD]  (row Cin = ones on valid positions, rows above: zero)
  float* gos = xs + C1R8 * LD;              // [CoutR16][LD]
  const int CoutR8 = (Cout + 7) & ~7, CoutR16 = (Cout + 15) & ~15;
  const int WS = q.WS;                      // weight row stride: >= round16(Cin), == 4 (mod 32)
  float* gxas = gos + CoutR16 * LD;         // [C1R8][LD]  (rows above Cin: zero)
  float* xas = gxas + C1R8 * LD;            // [C1][LD]
  float* gxs = xas + C1 * LD;               // [Cin][LD]
  float* xms = gxs + Cin * LD;              // [nb][PCH][K][KP2]   xmu[l][v][w]
  float* xmT = xms + nb * PCH * K * KP2;    // [nb][PCH][K][KP2]   xmu[l][w][v]
  float* wfB = xmT + nb * PCH * K * KP2;    // [nb][CoutR8][WS]   (zero padded)
  float* bfs = wfB + nb * CoutR8 * WS;      // [nb][Cout]
  float* aeff = bfs + ((nb * Cout + 3) & ~3);   // [nb][K*K]    static adjacency A*W + R
  float* pdr = aeff + ((nb * KK + 3) & ~3);     // [nb][PCH][K*K] raw dynamic adjacency of the current item
  int* rowtab = reinterpret_cast<int*>(pdr + ((nb * PCH * KK + 3) & ~3));   // [nb*PCH*K]  (b*PCH + l) << 8 | v per row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float alpha = q.alpha ? __ldg(q.alpha) : 1.0f;
  const int nchunk = (P + PCH - 1) / PCH;
  const long long nitems = (long long)q.N * nchunk;
  // the raw dynamic adjacency of an item is one contiguous block per branch: when K*K*4 is a multiple of 16 it is
  // staged by the TMA engine (cp.async.bulk, one instruction per branch) instead of one 4-byte cp.async per element
  __shared__ uint64_t pd_bar;
  const bool pd_bulk = (KK & 3) == 0 && (reinterpret_cast<uintptr_t>(q.pd) & 15) == 0;
  uint32_t pd_phase = 0;
  if (tid == 0) {
    umma::mbar_init(&pd_bar, 1);
    umma::mbar_init_fence();
  }

  // ---- once per CTA: weights
  for (int i = tid; i < nb * CoutR8 * WS; i += AMB_NT) {
    int c = i % WS, t = i / WS;
    int o = t % CoutR8, b = t / CoutR8;
    wfB[i] = (c < Cin && o < Cout) ? __ldg(q.w_f[b] + (long long)o * Cin + c) : 0.f;
  }
  // rows / columns that the 8- and 16-wide MMA tiles read beyond the data must be zero (K padding) for the whole kernel
  for (int i = tid; i < CoutR16 * LD; i += AMB_NT) gos[i] = 0.f;
  for (int i = tid; i < C1 * LD; i += AMB_NT) xas[i] = 0.f;
  for (int i = tid; i < (C1R8 - C1) * LD; i += AMB_NT) {
    xs[C1 * LD + i] = 0.f;
    gxas[C1 * LD + i] = 0.f;
  }
  for (int i = tid; i < nb * PCH * K; i += AMB_NT) rowtab[i] = ((i / K) << 8) | (i % K);
  for (int i = tid; i < nb * Cout; i += AMB_NT) bfs[i] = __ldg(q.b_f[i / Cout] + (i % Cout));
  for (int i = tid; i < nb * KK; i += AMB_NT) {
    const int b = i / KK, e = i - b * KK;
    float a = __ldg(q.adj[b] + e);
    if (q.adj_w[b]) a *= __ldg(q.adj_w[b] + e);
    if (q.adj_r[b]) a += __ldg(q.adj_r[b] + e);
    aeff[i] = a;
  }

  // persistent weight-gradient accumulators: warp = 8 output channels, lane = input channels (lane, lane+32)
  // persistent weight-gradient accumulators: warp = one 16-row tile of o x two 8-column tiles of j (mma C fragments)
  float accw[DSTD_MAX_BRANCH][2][4];
  float accb[DSTD_MAX_BRANCH];
#pragma unroll
  for (int b = 0; b < DSTD_MAX_BRANCH; ++b) {
    accb[b] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int r = 0; r < 4; ++r) accw[b][i][r] = 0.f;
  }
  const int fg = lane >> 2, ft = lane & 3;               // mma fragment coordinates
  const int MTo = CoutR16 >> 4, NTj = (Cin + 7) >> 3;    // tiles of the gradient matrix [o][j]
  const int c_mi = warp % MTo, c_np = warp / MTo;        // this warp's tiles in (c): m-tile, pair of n-tiles
  const bool c_act = 2 * c_np < NTj;

  // optional per-phase cycle counters (-DDSTD_PHASE_TIMING, printed by CTA 0): how the optimisation targets were picked
#ifdef DSTD_PHASE_TIMING
  long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#define PH(i) do { __syncthreads(); long long _t = clock64(); tph[i] += _t - tlast; tlast = _t; } while (0)
#else
#define PH(i) do {} while (0)
#endif
  for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int n = (int)(item / nchunk), p0 = (int)(item - (long long)n * nchunk) * PCH;
    const int pv = min(PCH, P - p0);
    __syncthreads();   // previous item fully consumed (also orders the one-time weight staging)

    // ---- stage x (+ ones row), gout and the raw dynamic adjacency with cp.async (all copies of a thread in flight at
    //      once), then build xmu = alpha*pd + A_eff in both orientations from shared memory
    {
      int poff_x[TN], poff_g[TN];
      bool pok[TN];
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        const int pos = lane + 32 * i;
        const int l = pos / KP, k = pos - l * KP;
        pok[i] = pos < npos_pad && l < pv && k < K;
        poff_x[i] = pok[i] ? (int)(l * q.x.sp + k * q.x.sk) : 0;
        poff_g[i] = pok[i] ? (int)(l * q.gout.sp + k * q.gout.sk) : 0;
      }
      const float* xb = q.x.p + (long long)n * q.x.sn + (long long)p0 * q.x.sp;
      for (int c = warp; c < Cin; c += AMB_NT / 32) {
#pragma unroll
        for (int i = 0; i < TN; ++i)
          if (lane + 32 * i < npos_pad) cp_async4(xs + c * LD + lane + 32 * i, xb + (long long)c * q.x.sc + poff_x[i], pok[i]);
      }
      const float* gb = q.gout.p + (long long)n * q.gout.sn + (long long)p0 * q.gout.sp;
      for (int c = warp; c < Cout; c += AMB_NT / 32) {
#pragma unroll
        for (int i = 0; i < TN; ++i)
          if (lane + 32 * i < npos_pad) cp_async4(gos + c * LD + lane + 32 * i, gb + (long long)c * q.gout.sc + poff_g[i], pok[i]);
      }
      if (pd_bulk) {
        if (tid == 0) {
          umma::mbar_expect_tx(&pd_bar, (uint32_t)(nb * pv * KK * 4));
          for (int b = 0; b < nb; ++b)
            umma::bulk_g2s(pdr + b * PCH * KK, q.pd + ((long long)(n * nb + b) * P + p0) * KK, (uint32_t)(pv * KK * 4), &pd_bar);
        }
      } else {
        for (int b = 0; b < nb; ++b) {
          const float* pdl = q.pd + ((long long)(n * nb + b) * P + p0) * KK;
          for (int i = tid; i < pv * KK; i += AMB_NT) cp_async4(pdr + b * PCH * KK + i, pdl + i, true);
        }
      }
      if (warp == 0) {
#pragma unroll
        for (int i = 0; i < TN; ++i)
          if (lane + 32 * i < npos_pad) xs[Cin * LD + lane + 32 * i] = pok[i] ? 1.0f : 0.f;
      }
      cp_async_wait_all();
      if (pd_bulk) {
        umma::mbar_wait(&pd_bar, pd_phase);
        pd_phase ^= 1;
      }
      __syncthreads();
      PH(0);
      // element-parallel (8 independent elements per thread); the row decode comes from a table built once per CTA, so
      // there is no runtime division in the loop (KP2 is a compile-time constant)
      for (int i = tid; i < nb * PCH * K * KP2; i += AMB_NT) {
        const int row = i / KP2, w = i - row * KP2;
        const int info = rowtab[row], bl = info >> 8, v = info & 255;   // bl = b * PCH + l
        const int b = bl >= PCH ? (bl >= 2 * PCH ? 2 : 1) : 0, l = bl - b * PCH;
        float val = 0.f, valT = 0.f;
        if (w < K && l < pv) {
          const int e = q.adj_t ? (w * K + v) : (v * K + w);      // xmu[v][w]
          const int eT = q.adj_t ? (v * K + w) : (w * K + v);     // xmu[w][v]
          const float* pr = pdr + bl * KK;
          val = fmaf(alpha, pr[e], aeff[b * KK + e]);
          valT = fmaf(alpha, pr[eT], aeff[b * KK + eT]);
        }
        xms[i] = val;
        xmT[i] = valT;
      }
    }
    __syncthreads();
    PH(1);
    // The tiles leave no room for a second staging buffer, so the next item's chunk is pulled into L2 instead while
    // this item computes: its cp.async staging then pays L2 instead of HBM latency.
    {
      const long long nxt = item + gridDim.x;
      if (nxt < nitems) {
        const int n2 = (int)(nxt / nchunk), q0 = (int)(nxt - (long long)n2 * nchunk) * PCH;
        const int pv2 = min(PCH, P - q0);
        auto pf_rows = [&](const float* base, long long row_stride, int rows, int run) {   // rows of `run` floats
          const int lines = (run * 4 + 127) / 128 + 1;
          for (int i = tid; i < rows * lines; i += AMB_NT) {
            const int r = i / lines, li = i - r * lines;
            const float* a = base + (long long)r * row_stride + min(li * 32, run - 1);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
        };
        if (q.x.sk == 1 && q.x.sp == K)
          pf_rows(q.x.p + (long long)n2 * q.x.sn + (long long)q0 * q.x.sp, q.x.sc, Cin, pv2 * K);
        if (q.gout.sk == 1 && q.gout.sp == K)
          pf_rows(q.gout.p + (long long)n2 * q.gout.sn + (long long)q0 * q.gout.sp, q.gout.sc, Cout, pv2 * K);
        pf_rows(q.pd + ((long long)n2 * nb * P + q0) * KK, (long long)P * KK, nb, pv2 * KK);
      }
    }

#pragma unroll
    for (int b = 0; b < DSTD_MAX_BRANCH; ++b) {   // unrolled: accw[b] must stay in registers
      if (b >= nb) break;
      // ================= (a) gxa_b = wcat_b^T gout on mma.sync (3xTF32): D[j][pos] = sum_o W[o][j] g[o][pos]
      //                   warp = one 16-row tile of j x TN 8-column tiles of positions, K = o in steps of 8
      {
        const int MT = (Cin + 15) >> 4, NTt = (npos_pad + 7) >> 3, NG = (NTt + TN - 1) / TN;
        const float* wb = wfB + b * CoutR8 * WS;
        for (int grp = warp; grp < MT * NG; grp += AMB_NT / 32) {
          const int m0 = (grp % MT) * 16, n0 = (grp / MT) * TN * 8;
          float acc[TN][4];
#pragma unroll
          for (int i = 0; i < TN; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
          for (int k0 = 0; k0 < CoutR8; k0 += 8) {
            // k-slot permutation (slot t <-> row 2t, slot t+4 <-> row 2t+1, same for A and B): with LD == 4 (mod 8)
            // and WS == 4 (mod 32) the four k rows of a load land 8 banks apart -> conflict-free fragment loads
            const float* wr = wb + (k0 + 2 * ft) * WS + m0 + fg;
            uint32_t ah[4], al[4];
            split3(wr[0], ah[0], al[0]);
            split3(wr[8], ah[1], al[1]);
            split3(wr[WS], ah[2], al[2]);
            split3(wr[WS + 8], ah[3], al[3]);
            const float* gr = gos + (k0 + 2 * ft) * LD + n0 + fg;
#pragma unroll
            for (int i = 0; i < TN; ++i) {
              uint32_t bh[2], bl[2];
              split3(gr[i * 8], bh[0], bl[0]);
              split3(gr[i * 8 + LD], bh[1], bl[1]);
              mma_tf32(acc[i], ah, bh);
              mma_tf32(acc[i], ah, bl);
              mma_tf32(acc[i], al, bh);
            }
          }
#pragma unroll
          for (int i = 0; i < TN; ++i) {
            const int col = n0 + i * 8 + 2 * ft;
            if (col < npos_pad) {
              if (m0 + fg < Cin) *reinterpret_cast<float2*>(gxas + (m0 + fg) * LD + col) = make_float2(acc[i][0], acc[i][1]);
              if (m0 + fg + 8 < Cin)
                *reinterpret_cast<float2*>(gxas + (m0 + fg + 8) * LD + col) = make_float2(acc[i][2], acc[i][3]);
            }
          }
        }
        // bias row: gxa[Cin][pos] = sum_o bf[o] gout[o][pos]   (warp = positions, lanes = output channels)
        // warp = 4 positions x 8 interleaved channel groups (o = 8 i + lane/4): with LD == 4 (mod 8) the 32 lanes hit
        // 32 banks; the 8 partial sums of a position are combined by a fixed butterfly
        for (int p4 = warp * 4; p4 < npos_pad; p4 += (AMB_NT / 32) * 4) {
          const int pos = p4 + (lane & 3), og = lane >> 2;
          float sres = 0.f;
          if (pos < npos_pad)
            for (int o = og; o < Cout; o += 8) sres = fmaf(bfs[b * Cout + o], gos[o * LD + pos], sres);
          sres += __shfl_xor_sync(0xffffffffu, sres, 4);
          sres += __shfl_xor_sync(0xffffffffu, sres, 8);
          sres += __shfl_xor_sync(0xffffffffu, sres, 16);
          if (og == 0 && pos < npos_pad) gxas[Cin * LD + pos] = sres;
        }
      }

      PH(2);
      // ================= (b) xa_b recompute on mma.sync: xa[c][l,w] = sum_v x[c][l,v] xmu_b[l][v][w]
      //                   warp = (frame, 16-channel tile), all w tiles; K = v, last step masked at v >= K
      {
        constexpr int NTW = KP2 / 8;
        const int MTc = (Cin + 15) >> 4;
        for (int grp = warp; grp < pv * MTc; grp += AMB_NT / 32) {
          const int l = grp / MTc, m0 = (grp % MTc) * 16;
          float acc[NTW][4];
#pragma unroll
          for (int i = 0; i < NTW; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
          const float* ar = xs + (m0 + fg) * LD + l * KP + ft;
          const float* br = xms + ((b * PCH + l) * K + ft) * KP2 + fg;
          for (int k0 = 0; k0 < K; k0 += 8) {
            const bool k_lo = k0 + ft < K, k_hi = k0 + ft + 4 < K;
            uint32_t ah[4], al[4];
            split3(k_lo ? ar[k0] : 0.f, ah[0], al[0]);
            split3(k_lo ? ar[k0 + 8 * LD] : 0.f, ah[1], al[1]);
            split3(k_hi ? ar[k0 + 4] : 0.f, ah[2], al[2]);
            split3(k_hi ? ar[k0 + 8 * LD + 4] : 0.f, ah[3], al[3]);
#pragma unroll
            for (int i = 0; i < NTW; ++i) {
              uint32_t bh[2], bl[2];
              split3(k_lo ? br[k0 * KP2 + i * 8] : 0.f, bh[0], bl[0]);
              split3(k_hi ? br[(k0 + 4) * KP2 + i * 8] : 0.f, bh[1], bl[1]);
              mma_tf32(acc[i], ah, bh);
              mma_tf32(acc[i], ah, bl);
              mma_tf32(acc[i], al, bh);
            }
          }
#pragma unroll
          for (int i = 0; i < NTW; ++i) {
            const int w = i * 8 + 2 * ft;
            if (w < KP) {
              if (m0 + fg < Cin)
                *reinterpret_cast<float2*>(xas + (m0 + fg) * LD + l * KP + w) = make_float2(acc[i][0], acc[i][1]);
              if (m0 + fg + 8 < Cin)
                *reinterpret_cast<float2*>(xas + (m0 + fg + 8) * LD + l * KP + w) = make_float2(acc[i][2], acc[i][3]);
            }
          }
        }
        // ones row: column sums of xmu (0 on padded columns); warp = frame, lane = w
        for (int l = warp; l < pv; l += AMB_NT / 32) {
          const float* xm_l = xms + ((b * PCH + l) * K) * KP2;
          for (int w = lane; w < KP; w += 32) {
            float sres = 0.f;
            for (int v = 0; v < K; ++v) sres += xm_l[v * KP2 + w];
            xas[Cin * LD + l * KP + w] = sres;
          }
        }
      }
      // frames beyond pv: xa must read as zero in (c)
      for (int i = tid; i < C1 * (PCH - pv) * KP; i += AMB_NT) {
        int c = i / ((PCH - pv) * KP), j = i - c * ((PCH - pv) * KP);
        xas[c * LD + pv * KP + j] = 0.f;
      }
      __syncthreads();

      PH(3);
      // ================= (c) weight gradient on mma.sync (3xTF32): G[o][j] += sum_pos gout[o][pos] xa_b[j][pos]
      if (c_act) {
        const int m0 = c_mi * 16;
        const float* ar = gos + (m0 + fg) * LD + ft;
        const float* br0 = xas + min((2 * c_np) * 8 + fg, Cin) * LD + ft;
        const float* br1 = xas + min((2 * c_np + 1) * 8 + fg, Cin) * LD + ft;
        const int ksteps = (npos_pad + 7) >> 3;
        for (int ks = 0; ks < ksteps; ++ks) {
          const int k0 = ks * 8;
          uint32_t ah[4], al[4], bh[2], bl[2];
          split3(ar[k0], ah[0], al[0]);
          split3(ar[k0 + 8 * LD], ah[1], al[1]);
          split3(ar[k0 + 4], ah[2], al[2]);
          split3(ar[k0 + 8 * LD + 4], ah[3], al[3]);
          split3(br0[k0], bh[0], bl[0]);
          split3(br0[k0 + 4], bh[1], bl[1]);
          mma_tf32(accw[b][0], ah, bh);
          mma_tf32(accw[b][0], ah, bl);
          mma_tf32(accw[b][0], al, bh);
          split3(br1[k0], bh[0], bl[0]);
          split3(br1[k0 + 4], bh[1], bl[1]);
          mma_tf32(accw[b][1], ah, bh);
          mma_tf32(accw[b][1], ah, bl);
          mma_tf32(accw[b][1], al, bh);
        }
      }
      if (tid < Cout) {   // bias gradient: sum_pos gout[o][pos] * (column sums of xmu)
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;     // four independent chains (the loop is latency bound)
        for (int p4 = 0; p4 < npos_pad; p4 += 4) {
          const float4 a = *reinterpret_cast<const float4*>(gos + tid * LD + p4);
          const float4 c = *reinterpret_cast<const float4*>(xas + Cin * LD + p4);
          s0 = fmaf(a.x, c.x, s0); s1 = fmaf(a.y, c.y, s1); s2 = fmaf(a.z, c.z, s2); s3 = fmaf(a.w, c.w, s3);
        }
        accb[b] += (s0 + s1) + (s2 + s3);
      }

      PH(4);
      // ================= (d) gx += gxa_b xmu_b^T on mma.sync: gx[c][l,v] += sum_w gxa[c][l,w] xmu_b[l][v][w]
      {
        constexpr int NTW = KP2 / 8;
        const int MTc = (Cin + 15) >> 4;
        for (int grp = warp; grp < pv * MTc; grp += AMB_NT / 32) {
          const int l = grp / MTc, m0 = (grp % MTc) * 16;
          float acc[NTW][4];
#pragma unroll
          for (int i = 0; i < NTW; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
          const float* ar = gxas + (m0 + fg) * LD + l * KP + ft;
          const float* br = xmT + ((b * PCH + l) * K + ft) * KP2 + fg;
          for (int k0 = 0; k0 < K; k0 += 8) {
            const bool k_lo = k0 + ft < K, k_hi = k0 + ft + 4 < K;
            uint32_t ah[4], al[4];
            split3(k_lo ? ar[k0] : 0.f, ah[0], al[0]);
            split3(k_lo ? ar[k0 + 8 * LD] : 0.f, ah[1], al[1]);
            split3(k_hi ? ar[k0 + 4] : 0.f, ah[2], al[2]);
            split3(k_hi ? ar[k0 + 8 * LD + 4] : 0.f, ah[3], al[3]);
#pragma unroll
            for (int i = 0; i < NTW; ++i) {
              uint32_t bh[2], bl[2];
              split3(k_lo ? br[k0 * KP2 + i * 8] : 0.f, bh[0], bl[0]);
              split3(k_hi ? br[(k0 + 4) * KP2 + i * 8] : 0.f, bh[1], bl[1]);
              mma_tf32(acc[i], ah, bh);
              mma_tf32(acc[i], ah, bl);
              mma_tf32(acc[i], al, bh);
            }
          }
#pragma unroll
          for (int i = 0; i < NTW; ++i) {
            const int v = i * 8 + 2 * ft;
            if (v < KP) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int c = m0 + fg + 8 * h;
                if (c < Cin) {
                  float2* dp = reinterpret_cast<float2*>(gxs + c * LD + l * KP + v);
                  float2 tv = make_float2(acc[i][2 * h], acc[i][2 * h + 1]);
                  if (b > 0) { const float2 o = *dp; tv.x += o.x; tv.y += o.y; }
                  *dp = tv;
                }
              }
            }
          }
        }
      }

      PH(5);
      // ================= (e) gxmu_b[l][v][w] = sum_{c<=Cin} xaug[c][l,v] gxa_b[c][l,w] on mma.sync -> HBM
      //                   warp = (frame, 16-row tile of v, half of the w tiles); K = c (zero rows above Cin)
      {
        constexpr int NTW = KP2 / 8, NH = (NTW + 1) / 2;
        const int MTv = (K + 15) >> 4;
        for (int grp = warp; grp < pv * MTv * 2; grp += AMB_NT / 32) {
          const int nh = grp & 1, l = (grp >> 1) / MTv, m0 = ((grp >> 1) % MTv) * 16;
          float acc[NH][4];
#pragma unroll
          for (int i = 0; i < NH; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
          const float* ar = xs + 2 * ft * LD + l * KP + m0 + fg;     // same k-slot permutation as in (a)
          const float* br = gxas + 2 * ft * LD + l * KP + nh * NH * 8 + fg;
          for (int k0 = 0; k0 < C1R8; k0 += 8) {
            uint32_t ah[4], al[4];
            split3(ar[k0 * LD], ah[0], al[0]);
            split3(ar[k0 * LD + 8], ah[1], al[1]);
            split3(ar[(k0 + 1) * LD], ah[2], al[2]);
            split3(ar[(k0 + 1) * LD + 8], ah[3], al[3]);
#pragma unroll
            for (int i = 0; i < NH; ++i) {
              uint32_t bh[2], bl[2];
              split3(br[k0 * LD + i * 8], bh[0], bl[0]);
              split3(br[(k0 + 1) * LD + i * 8], bh[1], bl[1]);
              mma_tf32(acc[i], ah, bh);
              mma_tf32(acc[i], ah, bl);
              mma_tf32(acc[i], al, bh);
            }
          }
          float* dst = q.gxm + ((long long)(n * nb + b) * P + p0 + l) * KK;
#pragma unroll
          for (int i = 0; i < NH; ++i) {
            const int w = (nh * NH + i) * 8 + 2 * ft;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int v = m0 + fg + 8 * h;
              if (v < K) {
                if (w < K) dst[q.adj_t ? (w * K + v) : (v * K + w)] = acc[i][2 * h];
                if (w + 1 < K) dst[q.adj_t ? ((w + 1) * K + v) : (v * K + w + 1)] = acc[i][2 * h + 1];
              }
            }
          }
        }
      }
      __syncthreads();   // gxas / xas are rewritten by the next branch
      PH(6);
    }

    // ---- gx chunk -> HBM (coalesced along the contiguous (l,k) run of each channel)
    {
      // warp = channel row, lanes = padded positions (KP is a compile-time constant: no runtime divisions)
      int poff[TN];
      bool pok[TN];
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        const int pos = lane + 32 * i;
        const int l = pos / KP, k = pos - l * KP;
        pok[i] = pos < npos_pad && l < pv && k < K;
        poff[i] = pok[i] ? (int)(l * q.gx.sp + k * q.gx.sk) : 0;
      }
      float* gb = q.gx.p + (long long)n * q.gx.sn + (long long)p0 * q.gx.sp;
      for (int c = warp; c < Cin; c += AMB_NT / 32) {
        const float* src = gxs + c * LD;
        float* dst = gb + (long long)c * q.gx.sc;
#pragma unroll
        for (int i = 0; i < TN; ++i)
          if (pok[i]) dst[poff[i]] = src[lane + 32 * i];
      }
    }
  }

  PH(7);
#ifdef DSTD_PHASE_TIMING
  if (blockIdx.x == 0 && tid == 0)
    printf("aggmix_bwd phases (cycles, CTA 0): stage %lld transform %lld a %lld b %lld c %lld d %lld e %lld tail %lld\n",
           tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], tph[6], tph[7]);
#endif
  // ---- per-CTA partials of the conv_f gradients
  float* pw = q.part_w + (long long)blockIdx.x * nb * Cout * Cin;
  float* pb = q.part_b + (long long)blockIdx.x * nb * Cout;
#pragma unroll
  for (int b = 0; b < DSTD_MAX_BRANCH; ++b) {
    if (b < nb) {
      if (c_act) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int j = (2 * c_np + i) * 8 + 2 * ft;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int o = c_mi * 16 + fg + 8 * h;
            if (o < Cout) {
              if (j < Cin) pw[((long long)b * Cout + o) * Cin + j] = accw[b][i][2 * h];
              if (j + 1 < Cin) pw[((long long)b * Cout + o) * Cin + j + 1] = accw[b][i][2 * h + 1];
            }
          }
        }
      }
      if (tid < Cout) pb[b * Cout + tid] = accb[b];
    }
  }
}

// ------------------------------------------------------------------------------------------ launch
struct AggMixBwdGeom {
  int KP, TN, PCH, LD, CinP, WS;
  size_t smem;
};

static bool aggmix_bwd_geom(int Cin, int Cout, int P, int K, int nb, AggMixBwdGeom& g) {
  if (K > 40 || K < 1 || Cin > 64 || Cout > 64) return false;
  g.KP = K <= 24 ? 24 : K <= 28 ? 28 : K <= 36 ? 36 : 40;
  const int WH = (g.KP / 2 + 3) / 4 * 4, KP2 = 2 * WH, C1 = Cin + 1;
  g.CinP = (Cin + 7) / 8 * 8;
  g.WS = (Cin + 15) / 16 * 16;
  while (g.WS % 32 != 4) g.WS += 4;          // conflict-free A-fragment loads of (a): row stride == 4 (mod 32)
  const int CoutR8 = (Cout + 7) / 8 * 8, CoutR16 = (Cout + 15) / 16 * 16;
  if ((CoutR16 / 16) * ((Cin + 15) / 16) > AMB_NT / 32) return false;   // (c): one warp per (16 x 16) gradient tile pair
  for (int pch = 128 / g.KP; pch >= 1; --pch) {
    if (pch > P && pch > 1) continue;
    const int npad = pch * g.KP;
    int ld = npad + 4;
    if ((ld / 4) % 2 == 0) ld += 4;
    const int C1R8 = (C1 + 7) / 8 * 8;
    size_t f = (size_t)(2 * C1R8 + C1 + CoutR16 + Cin) * ld + (size_t)2 * nb * pch * K * KP2 + (size_t)nb * CoutR8 * g.WS +
               (size_t)nb * Cout + (size_t)nb * K * K * (pch + 1) + (size_t)nb * pch * K + 40;
    if (f * sizeof(float) <= (size_t)MAX_DYN_SMEM - 512) {
      g.PCH = pch;
      g.LD = ld;
      g.TN = npad <= 96 ? 3 : 4;
      g.smem = f * sizeof(float);
      return true;
    }
  }
  return false;
}

bool aggmix_bwd_supported(int Cin, int Cout, int P, int K, int nb) {
  AggMixBwdGeom g;
  return aggmix_bwd_geom(Cin, Cout, P, K, nb, g);
}

int aggmix_bwd_ctas(int N, int P, int K, int Cin, int Cout, int nb) {
  AggMixBwdGeom g;
  if (!aggmix_bwd_geom(Cin, Cout, P, K, nb, g)) return 0;
  long long items = (long long)N * ((P + g.PCH - 1) / g.PCH);
  return (int)(items < num_sms() ? items : num_sms());
}
// per-CTA weight-gradient partials: two (one per warp group)

int launch_aggmix_bwd(AggMixBwdParams q, cudaStream_t st) {
  AggMixBwdGeom g;
  DSTD_REQUIRE(aggmix_bwd_geom(q.Cin, q.Cout, q.P, q.K, q.nb, g), DSTD_ERR_UNSUPPORTED,
               "aggmix_bwd: Cin=%d Cout=%d K=%d outside the compiled tile limits", q.Cin, q.Cout, q.K);
  q.PCH = g.PCH; q.LD = g.LD; q.CinP = g.CinP; q.WS = g.WS;
  const int ctas = aggmix_bwd_ctas(q.N, q.P, q.K, q.Cin, q.Cout, q.nb);
#define DSTD_AMB(KP_, TN_)                                           \
  if (g.KP == KP_ && g.TN == TN_) {                                  \
    auto kern = aggmix_bwd_kernel<KP_, TN_>;                         \
    ensure_max_smem((const void*)kern);                              \
    kern<<<ctas, AMB_NT, g.smem, st>>>(q);                              \
  }
  DSTD_AMB(24, 3) DSTD_AMB(24, 4) DSTD_AMB(28, 3) DSTD_AMB(28, 4)
  DSTD_AMB(36, 3) DSTD_AMB(36, 4) DSTD_AMB(40, 3) DSTD_AMB(40, 4)
#undef DSTD_AMB
  count_launch();
  return check_launch("aggmix_bwd");
}

}  // namespace dstd
