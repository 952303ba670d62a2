// C ABI of libdstd_b200 (include/dstd_b200.h): argument validation, workspace carving and the kernel schedule of
// the DSTD-GC unit (model/dstdgcn.py:80-94 as called from DSTDGCB.forward :145-150 / :157-161) and the 1x1 channel mix.
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <set>
#include <utility>

#include "kernels.cuh"

namespace dstd {

static thread_local char g_err[512] = "";
static std::atomic<int> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DSTD_ERR_CUDA;
  }
  return DSTD_OK;
}

// the 227 KB per-CTA limit covers static + dynamic shared memory
static int max_dyn_for(const void* kernel) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return MAX_DYN_SMEM;
  return MAX_DYN_SMEM - (int)fa.sharedSizeBytes;
}

static int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
  return d;
}

int num_sms() {
  static std::mutex mu;
  static std::map<int, int> cache;
  const int d = current_device();
  if (d < 0) return 148;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(d);
  if (it != cache.end()) return it->second;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n < 1) { cudaGetLastError(); n = 148; }
  cache[d] = n;
  return n;
}

// cudaFuncSetAttribute applies to the current device only: remember (device, kernel) pairs
void ensure_max_smem(const void* kernel) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  const int d = current_device();
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({d, kernel})) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_for(kernel));
  done.insert({d, kernel});
}

// Streaming kernels with modest dynamic shared memory: without a preference the driver picks the smallest carve-out
// that fits ONE CTA (measured: bn_bwd_apply ran 1 CTA/SM), so ask for the largest shared-memory partition once.
void prefer_smem_carveout(const void* kernel, bool need_max) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  const int d = current_device();
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({d, kernel})) return;
  if (need_max) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_for(kernel));
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  done.insert({d, kernel});
}

// ---- device-side error word (host-mapped, so that reading it never synchronises)
static int* g_err_host = nullptr;
static int* g_err_dev = nullptr;
static std::mutex g_err_mu;

int* device_error_word() {
  std::lock_guard<std::mutex> lk(g_err_mu);
  if (g_err_dev) return g_err_dev;
  int* h = nullptr;
  if (cudaHostAlloc((void**)&h, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();     // e.g. called for the first time under stream capture: the kernel then runs without the word
    return nullptr;
  }
  *h = 0;
  int* d = nullptr;
  if (cudaHostGetDevicePointer((void**)&d, h, 0) != cudaSuccess) {
    cudaGetLastError();
    cudaFreeHost(h);
    return nullptr;
  }
  g_err_host = h;
  g_err_dev = d;
  return g_err_dev;
}

int poll_device_error(const char* fn) {
  const int* h = g_err_host;
  if (h && *(volatile const int*)h != 0) {
    set_error("%s: an earlier kernel of this library reported a device-side failure (code %d: tensor-core pipeline "
              "timeout); its outputs are invalid.  dstd_device_error(1) clears the condition", fn, *(volatile const int*)h);
    return DSTD_ERR_CUDA;
  }
  return DSTD_OK;
}

__global__ void raise_device_error_kernel(int* word, int code) {
  *(volatile int*)word = code;
  __threadfence_system();
}

struct GcWs {
  float *wcat, *wm;                         // packed weights
  float *gxa, *gxm, *gm;                    // backward intermediates
  float *p_wcat, *p_wm, *p_wrm, *p_adj, *p_alpha;
  int S1, S2;
};

static size_t gc_fwd_ws(int Cin, int Cout, int nb, int P = 40, int K = 40) {
  return arena_need({(size_t)Cout * nb * (Cin + 1) * 4, (size_t)4 * nb * (Cin + 1) * 4,
                     (size_t)nb * (Cin + 1) * ((Cout + 7) / 8 * 8) * 4, aggmix_tc_ws_floats(Cin, Cout, P, K, nb) * 4,
                     unit_tc_ws_bytes(nb), bgemm_tc_ws_bytes(Cout, nb * (Cin + 1))});
}
// shapes beyond the specialised tiles (P or K > 40) run the shape-generic kernels of generic.cu
static bool use_generic(int Cin, int P, int K) { return !(dynadj_supported(P, K) && aggregate_supported(Cin, P, K)); }
static int dyn_splits(int N, int nb, int Cin, int P, int K) {
  return use_generic(Cin, P, K) ? N * dynadj_gen_tiles(K) : dynadj_bwd_splits(N, nb);
}
// per-CTA partial slots of the persistent kernels (one or two per SM)
static size_t part_slots() { return (size_t)2 * num_sms(); }
static size_t gc_bwd_ws(int N, int Cin, int Cout, int P, int K, int nb) {
  const size_t C1 = Cin + 1, G = (size_t)N * P * K;
  const int S1 = wgrad_splits((long long)G), S2 = dyn_splits(N, nb, Cin, P, K);
  const size_t unf = (!use_generic(Cin, P, K) && (unit_tc_supported(Cin, Cout, P, K, nb) || aggmix_bwd_supported(Cin, Cout, P, K, nb))) ? 0 : 1;   // buffers only the unfused path needs
  const size_t ps = part_slots();
  return arena_need({(size_t)Cout * nb * C1 * 4, (size_t)4 * nb * C1 * 4, unf * G * nb * C1 * 4, G * nb * K * 4,
                     G * nb * 4 * 4, unf * S1 * Cout * nb * C1 * 4, ((size_t)S1 > ps ? (size_t)S1 : ps) * 4 * nb * C1 * 4,
                     (size_t)S2 * nb * P * (2 * P + 1) * 4, (size_t)S2 * nb * K * K * 4, (size_t)S2 * nb * 4,
                     ps * nb * Cout * Cin * 4, ps * nb * Cout * 4, unit_tc_ws_bytes(nb), bgemm_tc_ws_bytes(nb * (int)C1, Cout),
                     use_generic(Cin, P, K) ? dynadj_bwd_gen_ws_floats(N, nb, P, K) * 4 : 0});
}

static int gc_check_common(int N, int Cin, int Cout, int P, int K, int nb, const dstd_branch* br, const char* fn) {
  DSTD_REQUIRE(N > 0 && Cin > 0 && Cout > 0 && P > 0 && K > 0, DSTD_ERR_BAD_ARG, "%s: non-positive dimension", fn);
  DSTD_REQUIRE(nb >= 1 && nb <= DSTD_MAX_BRANCH, DSTD_ERR_BAD_ARG, "%s: nb=%d outside [1,%d]", fn, nb, DSTD_MAX_BRANCH);
  for (int b = 0; b < nb; ++b)
    DSTD_REQUIRE(br[b].w_m1 && br[b].b_m1 && br[b].w_m2 && br[b].b_m2 && br[b].w_rm && br[b].b_rm && br[b].w_f &&
                     br[b].b_f && br[b].adj,
                 DSTD_ERR_BAD_ARG, "%s: branch %d has a null weight", fn, b);
  DSTD_REQUIRE((dynadj_supported(P, K) && aggregate_supported(Cin, P, K)) || generic_supported(P, K), DSTD_ERR_UNSUPPORTED,
               "%s: unit shape P=%d K=%d outside the compiled tile limits (P<=128, K<=128)", fn, P, K);
  return DSTD_OK;
}

static void fill_pack(PackParams& pk, int Cin, int Cout, int nb, const dstd_branch* br, float* wcat, float* wm,
                      float* wcatT = nullptr) {
  pk.wcatT = wcatT;
  pk.Cin = Cin; pk.Cout = Cout; pk.nb = nb;
  for (int b = 0; b < DSTD_MAX_BRANCH; ++b) {
    const dstd_branch& s = br[b < nb ? b : 0];
    pk.w_f[b] = s.w_f; pk.b_f[b] = s.b_f;
    pk.w_m1[b] = s.w_m1; pk.b_m1[b] = s.b_m1;
    pk.w_m2[b] = s.w_m2; pk.b_m2[b] = s.b_m2;
  }
  pk.wcat = wcat; pk.wm = wm;
}

static void fill_agg(AggParams& ag, int N, int Cin, int P, int K, int nb, int flags, const dstd_view& x,
                     const float* pd, const float* alpha, const dstd_branch* br) {
  ag.N = N; ag.Cin = Cin; ag.P = P; ag.K = K; ag.nb = nb;
  ag.adj_t = (flags & DSTD_FLAG_ADJ_T) ? 1 : 0;
  ag.x = mk(x);
  ag.pd = pd;
  ag.alpha = alpha;
  for (int b = 0; b < DSTD_MAX_BRANCH; ++b) {
    const dstd_branch& s = br[b < nb ? b : 0];
    ag.adj[b] = s.adj; ag.adj_w[b] = s.adj_w; ag.adj_r[b] = s.adj_r;
  }
  ag.xa = nullptr; ag.gxa = nullptr; ag.gxm = nullptr;
  ag.gx = View4{nullptr, 0, 0, 0, 0};
}

}  // namespace dstd

using namespace dstd;

extern "C" const char* dstd_last_error(void) { return g_err; }
extern "C" const char* dstd_version(void) { return "dstd_b200 0.1 sm_100a"; }
extern "C" int dstd_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int dstd_device_error(int clear) {
  device_error_word();
  int v = 0;
  if (g_err_host) {
    v = *(volatile int*)g_err_host;
    if (clear) *(volatile int*)g_err_host = 0;
  }
  return v;
}
extern "C" int dstd_debug_raise_device_error(int code, dstd_stream_t stream) {
  int* w = device_error_word();
  DSTD_REQUIRE(w, DSTD_ERR_CUDA, "debug_raise_device_error: the error word is unavailable");
  raise_device_error_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(w, code);
  count_launch();
  return check_launch("raise_device_error");
}

extern "C" int dstd_gc_needs_xa(int Cin, int Cout, int P, int K, int nb) {
  if (use_generic(Cin, P, K)) return 1;
  if (unit_tc_supported(Cin, Cout, P, K, nb)) return 0;
  return (aggmix_supported(Cin, Cout, P, K, nb) && aggmix_bwd_supported(Cin, Cout, P, K, nb)) ? 0 : 1;
}
extern "C" size_t dstd_gc_fwd_workspace_bytes(int N, int Cin, int Cout, int P, int K, int nb) {
  (void)N;
  return gc_fwd_ws(Cin, Cout, nb, P, K);
}
extern "C" size_t dstd_gc_bwd_workspace_bytes(int N, int Cin, int Cout, int P, int K, int nb) {
  return gc_bwd_ws(N, Cin, Cout, P, K, nb);
}

// ---------------------------------------------------------------------------------------------- DSTD-GC forward
extern "C" int dstd_gc_forward(const dstd_gc_fwd_args* a, dstd_stream_t stream) {
  DSTD_REQUIRE(a, DSTD_ERR_BAD_ARG, "gc_forward: null args");
  int rc = poll_device_error("gc_forward");
  if (rc) return rc;
  rc = gc_check_common(a->N, a->Cin, a->Cout, a->P, a->K, a->nb, a->br, "gc_forward");
  if (rc) return rc;
  DSTD_REQUIRE(a->x.ptr && a->out.ptr && a->m && a->pd, DSTD_ERR_BAD_ARG, "gc_forward: null tensor");
  DSTD_REQUIRE(a->ws && a->ws_bytes >= gc_fwd_ws(a->Cin, a->Cout, a->nb, a->P, a->K), DSTD_ERR_WORKSPACE,
               "gc_forward: workspace too small (%zu < %zu)", a->ws_bytes, gc_fwd_ws(a->Cin, a->Cout, a->nb, a->P, a->K));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = a->N, Cin = a->Cin, Cout = a->Cout, P = a->P, K = a->K, nb = a->nb, C1 = Cin + 1;
  const long long G = (long long)N * P * K;
  Arena ar(a->ws, a->ws_bytes);
  float* wcat = ar.take<float>((size_t)Cout * nb * C1);
  float* wm = ar.take<float>((size_t)4 * nb * C1);
  float* wcatT = ar.take<float>((size_t)nb * C1 * ((Cout + 7) / 8 * 8));
  float* wtc = ar.take<float>(aggmix_tc_ws_floats(Cin, Cout, P, K, nb));
  void* wunit = ar.take<char>(unit_tc_ws_bytes(nb));
  void* wbt = ar.take<char>(bgemm_tc_ws_bytes(Cout, nb * C1));               // weight images of the tcgen05 channel GEMM
  const bool use_unit = !a->xa && unit_tc_supported(Cin, Cout, P, K, nb);    // every contraction on tcgen05 (unit_tc.cu)
  const bool use_tc = !a->xa && aggmix_tc_supported(Cin, Cout, P, K, nb);   // tcgen05 channel mix only (aggmix_tc.cu)
  const bool fused = !use_generic(Cin, P, K) && (use_unit || aggmix_supported(Cin, Cout, P, K, nb));

  PackParams pk;
  fill_pack(pk, Cin, Cout, nb, a->br, wcat, wm, fused ? wcatT : nullptr);
  if ((rc = launch_pack(pk, st))) return rc;

  // 1. m = [conv_m1; conv_m2] x + b        (all branches in one streaming pass over x)
  {
    MprojFwdParams mp;
    mp.Cin = Cin; mp.J = 4 * nb; mp.P = P; mp.K = K; mp.G = G;
    mp.x = mk(a->x); mp.wm = wm; mp.m = a->m;
    if ((rc = launch_mproj_fwd(mp, st))) return rc;
  }

  // 2. pd = conv_rm(tanh(m1 - m2))         (pairwise tensor stays on chip)
  DynAdjFwdParams dp;
  dp.N = N; dp.P = P; dp.K = K; dp.nb = nb; dp.m = a->m; dp.pd = a->pd;
  for (int b = 0; b < DSTD_MAX_BRANCH; ++b) {
    dp.w_rm[b] = a->br[b < nb ? b : 0].w_rm;
    dp.b_rm[b] = a->br[b < nb ? b : 0].b_rm;
  }
  const bool generic = use_generic(Cin, P, K);
  if ((rc = generic ? launch_dynadj_fwd_gen(dp, st) : launch_dynadj_fwd(dp, st))) return rc;

  // 3+4 fused: out = wcat ([x;1] (alpha pd + A)) (+ skip); the aggregated tile stays in shared memory
  if (fused) {
    AggMixParams am;
    am.N = N; am.Cin = Cin; am.Cout = Cout; am.P = P; am.K = K; am.nb = nb;
    am.adj_t = (a->flags & DSTD_FLAG_ADJ_T) ? 1 : 0;
    am.PCH = 0; am.CoutP = 0;
    am.x = mk(a->x); am.out = mk(a->out); am.skip = mk(a->skip);
    am.pd = a->pd; am.alpha = a->alpha;
    for (int b = 0; b < DSTD_MAX_BRANCH; ++b) {
      const dstd_branch& s = a->br[b < nb ? b : 0];
      am.adj[b] = s.adj; am.adj_w[b] = s.adj_w; am.adj_r[b] = s.adj_r;
    }
    am.wcatT = wcatT;
    am.wtc = nullptr;
    am.xa = a->xa;
    if (use_unit) return launch_unit_fwd_tc(am, pk, wunit, st);
    if (use_tc) return launch_aggmix_fwd_tc(am, pk, wtc, st);
    return launch_aggmix_fwd(am, st);
  }
  DSTD_REQUIRE(a->xa, DSTD_ERR_BAD_ARG, "gc_forward: the unfused path needs the xa buffer");

  // 3. xa = [x;1] (alpha pd + A)           (aggregation before the channel mix)
  AggParams ag;
  fill_agg(ag, N, Cin, P, K, nb, a->flags, a->x, a->pd, a->alpha, a->br);
  ag.xa = a->xa;
  if ((rc = generic ? launch_aggregate_fwd_gen(ag, st) : launch_aggregate_fwd(ag, st))) return rc;

  // 4. out = [Wf_0 bf_0 | Wf_1 bf_1] xa (+ skip)
  BgemmParams g2;
  g2.M = Cout; g2.Kd = nb * C1; g2.G = G; g2.P = P; g2.K = K;
  g2.w = wcat; g2.wsc = 1; g2.wsi = nb * C1; g2.bias = nullptr;
  g2.in = dense_view(a->xa, nb * C1, P, K); g2.ones_row = -1;
  g2.out = mk(a->out);
  g2.add = mk(a->skip);
  g2.tc_ws = bgemm_tc_ws_bytes(Cout, nb * C1) ? wbt : nullptr;
  return launch_bgemm(g2, st);
}

// ---------------------------------------------------------------------------------------------- DSTD-GC backward
extern "C" int dstd_gc_backward(const dstd_gc_bwd_args* a, dstd_stream_t stream) {
  DSTD_REQUIRE(a, DSTD_ERR_BAD_ARG, "gc_backward: null args");
  int rc = poll_device_error("gc_backward");
  if (rc) return rc;
  rc = gc_check_common(a->N, a->Cin, a->Cout, a->P, a->K, a->nb, a->br, "gc_backward");
  if (rc) return rc;
  DSTD_REQUIRE(a->x.ptr && a->gout.ptr && a->gx.ptr && a->m && a->pd, DSTD_ERR_BAD_ARG, "gc_backward: null tensor");
  const int N = a->N, Cin = a->Cin, Cout = a->Cout, P = a->P, K = a->K, nb = a->nb, C1 = Cin + 1;
  const size_t need = gc_bwd_ws(N, Cin, Cout, P, K, nb);
  DSTD_REQUIRE(a->ws && a->ws_bytes >= need, DSTD_ERR_WORKSPACE, "gc_backward: workspace too small (%zu < %zu)",
               a->ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const long long G = (long long)N * P * K;
  const int S1 = wgrad_splits(G), S2 = dyn_splits(N, nb, Cin, P, K);
  const bool generic = use_generic(Cin, P, K);
  Arena ar(a->ws, a->ws_bytes);
  float* wcat = ar.take<float>((size_t)Cout * nb * C1);
  float* wm = ar.take<float>((size_t)4 * nb * C1);
  const bool use_unit = unit_tc_supported(Cin, Cout, P, K, nb);
  const bool fused = !generic && (use_unit || aggmix_bwd_supported(Cin, Cout, P, K, nb));
  const size_t ps = part_slots();
  float* gxa = ar.take<float>(fused ? 0 : (size_t)G * nb * C1);
  float* gxm = ar.take<float>((size_t)G * nb * K);
  float* gm = ar.take<float>((size_t)G * nb * 4);
  float* p_wcat = ar.take<float>(fused ? 0 : (size_t)S1 * Cout * nb * C1);
  float* p_wm = ar.take<float>(((size_t)S1 > ps ? (size_t)S1 : ps) * 4 * nb * C1);
  float* p_wrm = ar.take<float>((size_t)S2 * nb * P * (2 * P + 1));
  float* p_adj = ar.take<float>((size_t)S2 * nb * K * K);
  float* p_alpha = ar.take<float>((size_t)S2 * nb);
  float* p_wf = ar.take<float>(ps * nb * Cout * Cin);
  float* p_bf = ar.take<float>(ps * nb * Cout);
  void* wunit = ar.take<char>(unit_tc_ws_bytes(nb));
  void* wbt = ar.take<char>(bgemm_tc_ws_bytes(nb * C1, Cout));
  float* gm2_part = ar.take<float>(generic ? dynadj_bwd_gen_ws_floats(N, nb, P, K) : 0);

  PackParams pk;
  fill_pack(pk, Cin, Cout, nb, a->br, wcat, wm);
  if ((rc = launch_pack(pk, st))) return rc;

  int S1a = 0, n_wf = 0;
  if (fused) {
    // 1-3 fused: gxa, xa (recomputed), g(wcat), gx (aggregation part), gxm without leaving the SM
    AggMixBwdParams am;
    am.N = N; am.Cin = Cin; am.Cout = Cout; am.P = P; am.K = K; am.nb = nb;
    am.adj_t = (a->flags & DSTD_FLAG_ADJ_T) ? 1 : 0;
    am.PCH = am.LD = am.CinP = am.WS = 0;
    am.x = mk(a->x); am.gout = mk(a->gout); am.gx = mk(a->gx);
    am.pd = a->pd; am.alpha = a->alpha;
    for (int b = 0; b < DSTD_MAX_BRANCH; ++b) {
      const dstd_branch& s = a->br[b < nb ? b : 0];
      am.adj[b] = s.adj; am.adj_w[b] = s.adj_w; am.adj_r[b] = s.adj_r;
      am.w_f[b] = s.w_f; am.b_f[b] = s.b_f;
    }
    am.gxm = gxm; am.part_w = p_wf; am.part_b = p_bf;
    if (use_unit) {
      if ((rc = launch_unit_bwd_tc(am, pk, wunit, st))) return rc;
      n_wf = unit_bwd_tc_ctas(N, Cin, Cout, P, K, nb);
    } else {
      if ((rc = launch_aggmix_bwd(am, st))) return rc;
      n_wf = aggmix_bwd_ctas(N, P, K, Cin, Cout, nb);
    }
  } else {
  DSTD_REQUIRE(a->xa, DSTD_ERR_BAD_ARG, "gc_backward: the unfused path needs the saved xa buffer");

  // 1. gxa = wcat^T gout
  BgemmParams g1;
  g1.M = nb * C1; g1.Kd = Cout; g1.G = G; g1.P = P; g1.K = K;
  g1.w = wcat; g1.wsc = nb * C1; g1.wsi = 1; g1.bias = nullptr;
  g1.in = mk(a->gout); g1.ones_row = -1;
  g1.out = dense_view(gxa, nb * C1, P, K);
  g1.add = View4{nullptr, 0, 0, 0, 0};
  g1.tc_ws = bgemm_tc_ws_bytes(nb * C1, Cout) ? wbt : nullptr;
  if ((rc = launch_bgemm(g1, st))) return rc;

  // 2. g(wcat)[o, j] = sum gout[o] xa[j]
  WgradParams w1;
  w1.M = Cout; w1.Cd = nb * C1; w1.G = G; w1.P = P; w1.K = K;
  w1.a = mk(a->gout);
  w1.b = dense_view(const_cast<float*>(a->xa), nb * C1, P, K);
  w1.b_ones_row = -1;
  w1.partial = p_wcat;
  if ((rc = launch_wgrad(w1, st))) return rc;
  S1a = w1.S;

  // 3. aggregation backward: gx (first contribution), gxm
  AggParams ag;
  fill_agg(ag, N, Cin, P, K, nb, a->flags, a->x, a->pd, a->alpha, a->br);
  ag.gxa = gxa;
  ag.gx = mk(a->gx);
  ag.gxm = gxm;
  if ((rc = generic ? launch_aggregate_bwd_gen(ag, st) : launch_aggregate_bwd(ag, st))) return rc;
  }

  // 4. dynamic adjacency backward: gm, partial gWrm/gbrm, gA_eff, galpha
  DynAdjBwdParams db;
  db.N = N; db.P = P; db.K = K; db.nb = nb;
  db.m = a->m; db.pd = a->pd; db.gxm = gxm; db.alpha = a->alpha; db.gm = gm;
  for (int b = 0; b < DSTD_MAX_BRANCH; ++b) {
    db.w_rm[b] = a->br[b < nb ? b : 0].w_rm;
    db.b_rm[b] = a->br[b < nb ? b : 0].b_rm;
  }
  db.S = S2; db.part_wrm = p_wrm; db.part_adj = p_adj; db.part_alpha = p_alpha;
  if ((rc = generic ? launch_dynadj_bwd_gen(db, gm2_part, st) : launch_dynadj_bwd(db, st))) return rc;

  // 5+6. gx += wm^T gm ; g(wm)[j, c] = sum gm[j] [x;1][c]   (one pass over x)
  int S1b = 0;
  if (mproj_bwd_supported(Cin, 4 * nb)) {
    MprojBwdParams mp;
    mp.Cin = Cin; mp.J = 4 * nb; mp.P = P; mp.K = K; mp.G = G;
    mp.x = mk(a->x); mp.gx = mk(a->gx); mp.gx_add = mk(a->gx_add); mp.gm = gm; mp.wm = wm; mp.partial = p_wm;
    if ((rc = launch_mproj_bwd(mp, st))) return rc;
    S1b = mproj_bwd_ctas(G);
  } else {
    DSTD_REQUIRE(!a->gx_add.ptr, DSTD_ERR_UNSUPPORTED, "gc_backward: gx_add needs the fused m-projection backward (Cin <= 320)");
    BgemmParams g2;
    g2.M = Cin; g2.Kd = 4 * nb; g2.G = G; g2.P = P; g2.K = K;
    g2.w = wm; g2.wsc = C1; g2.wsi = 1; g2.bias = nullptr;
    g2.in = dense_view(gm, 4 * nb, P, K); g2.ones_row = -1;
    g2.out = mk(a->gx);
    g2.add = mk(a->gx);
    if ((rc = launch_bgemm(g2, st))) return rc;
    WgradParams w2;
    w2.M = 4 * nb; w2.Cd = C1; w2.G = G; w2.P = P; w2.K = K;
    w2.a = dense_view(gm, 4 * nb, P, K);
    w2.b = mk(a->x);
    w2.b_ones_row = Cin;
    w2.partial = p_wm;
    if ((rc = launch_wgrad(w2, st))) return rc;
    S1b = w2.S;
  }

  // 7. reduce the split partials and scatter them into the parameter gradients
  ReduceParams rp;
  rp.nseg = 0;
  auto seg = [&](const float* src, int S, long long sstride, int rows, int cols, int src_ld, float* dst, int dst_ld,
                 const float* mul = nullptr, float* dst2 = nullptr) {
    if (!dst && !dst2) return;
    ReduceSeg& s = rp.seg[rp.nseg++];
    s.src = src; s.S = S; s.sstride = sstride; s.rows = rows; s.cols = cols; s.src_ld = src_ld;
    s.dst = dst; s.dst_ld = dst_ld; s.mul = mul; s.dst2 = dst2; s.scale = 1.0f;
  };
  const long long st_wcat = (long long)Cout * nb * C1, st_wm = (long long)4 * nb * C1;
  const int P21 = 2 * P + 1;
  for (int b = 0; b < nb; ++b) {
    const dstd_branch_grad& g = a->gbr[b];
    if (fused) {
      seg(p_wf + (long long)b * Cout * Cin, n_wf, (long long)nb * Cout * Cin, Cout, Cin, Cin, g.w_f, Cin);
      seg(p_bf + (long long)b * Cout, n_wf, (long long)nb * Cout, Cout, 1, 1, g.b_f, 1);
    } else {
      seg(p_wcat + b * C1, S1a, st_wcat, Cout, Cin, nb * C1, g.w_f, Cin);
      seg(p_wcat + b * C1 + Cin, S1a, st_wcat, Cout, 1, nb * C1, g.b_f, 1);
    }
    seg(p_wm + (b * 4 + 0) * C1, S1b, st_wm, 2, Cin, C1, g.w_m1, Cin);
    seg(p_wm + (b * 4 + 0) * C1 + Cin, S1b, st_wm, 2, 1, C1, g.b_m1, 1);
    seg(p_wm + (b * 4 + 2) * C1, S1b, st_wm, 2, Cin, C1, g.w_m2, Cin);
    seg(p_wm + (b * 4 + 2) * C1 + Cin, S1b, st_wm, 2, 1, C1, g.b_m2, 1);
    seg(p_wrm + (long long)b * P * P21, S2, (long long)nb * P * P21, P, 2 * P, P21, g.w_rm, 2 * P);
    seg(p_wrm + (long long)b * P * P21 + 2 * P, S2, (long long)nb * P * P21, P, 1, P21, g.b_rm, 1);
    float* gadjw = a->br[b].adj_w ? g.adj_w : nullptr;
    seg(p_adj + (long long)b * K * K, generic ? N : S2, (long long)nb * K * K, K, K, K, g.adj_eff, K, a->br[b].adj, gadjw);
  }
  seg(p_alpha, S2 * nb, 1, 1, 1, 1, a->alpha ? a->galpha : nullptr, 1);
  return launch_reduce(rp, st);
}

// ---------------------------------------------------------------------------------------------- 1x1 channel mix
extern "C" size_t dstd_chmix_bwd_workspace_bytes(int N, int Cin, int Cout, int P, int K) {
  const long long G = (long long)N * P * K;
  return arena_need({(size_t)wgrad_splits(G) * Cout * (Cin + 1) * 4});
}

extern "C" int dstd_chmix_forward(const dstd_chmix_fwd_args* a, dstd_stream_t stream) {
  DSTD_REQUIRE(a && a->x.ptr && a->out.ptr && a->w, DSTD_ERR_BAD_ARG, "chmix_forward: null argument");
  DSTD_REQUIRE(a->N > 0 && a->Cin > 0 && a->Cout > 0 && a->P > 0 && a->K > 0, DSTD_ERR_BAD_ARG,
               "chmix_forward: non-positive dimension");
  BgemmParams g;
  g.M = a->Cout; g.Kd = a->Cin; g.G = (long long)a->N * a->P * a->K; g.P = a->P; g.K = a->K;
  g.w = a->w; g.wsc = 1; g.wsi = a->Cin; g.bias = a->b;
  g.in = mk(a->x); g.ones_row = -1;
  g.out = mk(a->out);
  g.add = View4{nullptr, 0, 0, 0, 0};
  return launch_bgemm(g, (cudaStream_t)stream);
}

extern "C" int dstd_chmix_backward(const dstd_chmix_bwd_args* a, dstd_stream_t stream) {
  DSTD_REQUIRE(a && a->x.ptr && a->gout.ptr && a->w, DSTD_ERR_BAD_ARG, "chmix_backward: null argument");
  DSTD_REQUIRE(a->N > 0 && a->Cin > 0 && a->Cout > 0 && a->P > 0 && a->K > 0, DSTD_ERR_BAD_ARG,
               "chmix_backward: non-positive dimension");
  cudaStream_t st = (cudaStream_t)stream;
  const int Cin = a->Cin, Cout = a->Cout, C1 = Cin + 1;
  const long long G = (long long)a->N * a->P * a->K;
  int rc;
  if (a->gx.ptr) {
    BgemmParams g;
    g.M = Cin; g.Kd = Cout; g.G = G; g.P = a->P; g.K = a->K;
    g.w = a->w; g.wsc = Cin; g.wsi = 1; g.bias = nullptr;
    g.in = mk(a->gout); g.ones_row = -1;
    g.out = mk(a->gx);
    g.add = View4{nullptr, 0, 0, 0, 0};
    if ((rc = launch_bgemm(g, st))) return rc;
  }
  if (a->gw || a->gb) {
    DSTD_REQUIRE(a->ws && a->ws_bytes >= dstd_chmix_bwd_workspace_bytes(a->N, Cin, Cout, a->P, a->K),
                 DSTD_ERR_WORKSPACE, "chmix_backward: workspace too small");
    Arena ar(a->ws, a->ws_bytes);
    const int S = wgrad_splits(G);
    float* part = ar.take<float>((size_t)S * Cout * C1);
    WgradParams w;
    w.M = Cout; w.Cd = C1; w.G = G; w.P = a->P; w.K = a->K;
    w.a = mk(a->gout);
    w.b = mk(a->x);
    w.b_ones_row = Cin;
    w.partial = part;
    if ((rc = launch_wgrad(w, st))) return rc;
    const int Sa = w.S;
    ReduceParams rp;
    rp.nseg = 0;
    if (a->gw) {
      ReduceSeg& s = rp.seg[rp.nseg++];
      s.src = part; s.S = Sa; s.sstride = (long long)Cout * C1; s.rows = Cout; s.cols = Cin; s.src_ld = C1;
      s.dst = a->gw; s.dst_ld = Cin; s.mul = nullptr; s.dst2 = nullptr; s.scale = 1.0f;
    }
    if (a->gb) {
      ReduceSeg& s = rp.seg[rp.nseg++];
      s.src = part + Cin; s.S = Sa; s.sstride = (long long)Cout * C1; s.rows = Cout; s.cols = 1; s.src_ld = C1;
      s.dst = a->gb; s.dst_ld = 1; s.mul = nullptr; s.dst2 = nullptr; s.scale = 1.0f;
    }
    return launch_reduce(rp, st);
  }
  return DSTD_OK;
}
