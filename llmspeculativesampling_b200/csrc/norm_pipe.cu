// Kernel 1, persistent warp-specialised version of the fast top-k path (the headline configuration:
// 0 < top_k <= 128, any top_p, dense fp32 probabilities out).
//
// Why a second kernel: in norm.cu a CTA loads its slice, selects, writes — the HBM pipe idles while the
// (latency-bound) selection runs, and with only ~2.6 waves of CTAs the phases never de-synchronise.  Here
// every SM runs ONE persistent CTA made of
//     1 memory warp      issues the 1-D TMA bulk load of work item i+G (slice of a logits row) into the
//                        shared-memory buffer that compute group (i+G) % G just released, then writes the
//                        zeros of that item's output slice (st.global.cs.v4, 512 B per instruction)
//     G compute groups   of 4 warps; group g owns buffer g and work items g, g+G, ...: thread maxima ->
//                        exact pivot (k-th largest thread maximum) -> re-scan of the few "hot" threads ->
//                        candidates pushed to the peer CTAs of the cluster through distributed shared
//                        memory (st.shared::cluster + remote mbarrier arrive, no cluster-wide barrier) ->
//                        merge, rank sort, top-k / top-p / softmax on ~2k candidates -> scatter of the
//                        non-zero probabilities (-> optional inverse-CDF sample)
// so the selection latency of G items is overlapped with the memory traffic of the next ones.  Buffers are
// handed back and forth with mbarriers (full / zero-filled / empty).  Rows whose candidate lists overflow
// (massive ties, adversarial layouts) are remembered in a shared-memory bit mask and re-run by the same cluster on
// the general path (norm_row) after its pipeline has drained, so results never depend on the fast path's luck.
//
// Replaces the same reference code as norm.cu: /root/reference/sampling/utils.py:152-210 (+213-233).
#include "norm_pipe_kernel.cuh"

#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace sd {

// one translation unit per logits dtype instantiates the kernels (norm_pipe_f32.cu, norm_pipe_bf16.cu, norm_pipe_f16.cu)
cudaError_t pipe_dispatch_f32(const NormParams& p, int rows, cudaStream_t st, int* q);
cudaError_t pipe_dispatch_bf16(const NormParams& p, int rows, cudaStream_t st, int* q);
cudaError_t pipe_dispatch_f16(const NormParams& p, int rows, cudaStream_t st, int* q);

static cudaError_t pipe_dispatch_dtype(const NormParams& p, int dtype, int rows, cudaStream_t st, int* q) {
  switch (dtype) {
    case kF32: return pipe_dispatch_f32(p, rows, st, q);
    case kBF16: return pipe_dispatch_bf16(p, rows, st, q);
    case kF16: return pipe_dispatch_f16(p, rows, st, q);
    default: return cudaErrorInvalidValue;
  }
}

// Decides whether the pipelined kernel applies and with which geometry: among the cluster sizes whose slices leave
// room for >= 2 buffers, take the one that keeps the most SMs busy (clusters of 3..16 CTAs cannot use every SM).
// The choice is cached per (dtype, V, top_k bucket).  Returns false if the kernel is not applicable.
struct PipePlan { long long V; int dtype, kcap, cluster, groups, buffers, cap, max_clusters; };
static PipePlan g_plans[32];
static int g_n_plans = 0;
static std::mutex g_plan_mutex;          // the library may be called from several host threads (one stream each)

bool plan_pipe(NormParams& p, int dtype, int rows, int tune_cluster) {
  if (p.sched == nullptr) return false;
  const size_t es = dtype == kF32 ? 4 : 2;
  if (p.top_k <= 0 || p.top_k > kPipeFastK || p.force_general || rows < 2) return false;
  const long long row_bytes = p.V * static_cast<long long>(es);
  const bool aligned_in = (reinterpret_cast<uintptr_t>(p.logits) % 16 == 0) && ((p.ld_in * es) % 16 == 0) && (row_bytes % 16 == 0);
  if (!aligned_in) return false;
  const int kcap = p.top_k + 8;                              // per-rank receive region must hold k + slack
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  const PipePlan* plan = nullptr;
  for (int i = 0; i < g_n_plans; ++i)
    if (g_plans[i].V == p.V && g_plans[i].dtype == dtype && g_plans[i].kcap == kcap && tune_cluster == 0) plan = &g_plans[i];
  PipePlan fresh = {p.V, dtype, kcap, 0, 0, 0, 0, 0};
  if (plan == nullptr) {
    double best_score = 0.0;
    for (int C = 1; C <= kMaxPortableCluster; ++C) {
      if (tune_cluster > 0 && C != tune_cluster) continue;
      const long long slice = ((p.V + C - 1) / C + 127) & ~127LL;
      if (C > 1 && slice * (C - 1) >= p.V) continue;         // last rank would be empty
      if (slice / kPipeGroupThreads > 60000) continue;
      const size_t slice_bytes = (static_cast<size_t>(slice) * es + 127) & ~static_cast<size_t>(127);
      const PipeVariant* var = nullptr;
      for (int vi = 0; vi < kNumPipeVariants && var == nullptr; ++vi)
        if (pipe_fits(kPipeVariants[vi], slice_bytes, C, kcap)) var = &kPipeVariants[vi];
      if (var == nullptr) continue;
      NormParams q = p;
      q.cluster = C; q.slice_elems = static_cast<int>(slice); q.slice_smem_bytes = static_cast<int>(slice_bytes);
      q.pipe_groups = var->ng; q.pipe_buffers = var->nb; q.pipe_cap = var->cap;
      int n = 0;
      if (pipe_dispatch_dtype(q, dtype, rows, nullptr, &n) != cudaSuccess || n < 1) { (void)cudaGetLastError(); continue; }
      // busy SMs, preferring three buffers (deeper prefetch), more groups than buffers, and small clusters
      const double score = static_cast<double>(n) * C * (var->nb == 3 ? 1.0 : 0.93) * (var->ng > var->nb ? 1.0 : 0.96) * (1.0 - 0.01 * C);
      if (getenv("SD_DEBUG") != nullptr)
        fprintf(stderr, "[specdec] pipe plan V=%lld es=%zu C=%d groups=%d buffers=%d CAP=%d slice=%zuB: %d clusters, score %.1f\n",
                p.V, es, C, var->ng, var->nb, var->cap, slice_bytes, n, score);
      if (score > best_score) {
        best_score = score; fresh.cluster = C; fresh.groups = var->ng; fresh.buffers = var->nb; fresh.cap = var->cap; fresh.max_clusters = n;
      }
      const int G = var->nb;
      if (tune_cluster == 0 && n * C >= 148 && G == 3) break;          // cannot do better than all SMs with three buffers
    }
    if (tune_cluster == 0 && g_n_plans < 32) { g_plans[g_n_plans] = fresh; plan = &g_plans[g_n_plans++]; }
    else plan = &fresh;
  }
  if (plan->cluster == 0) return false;
  const int C = plan->cluster;
  const long long slice = ((p.V + C - 1) / C + 127) & ~127LL;
  const int n_clusters = plan->max_clusters < rows ? plan->max_clusters : rows;
  p.cluster = C;
  p.slice_elems = static_cast<int>(slice);
  p.slice_smem_bytes = static_cast<int>((static_cast<size_t>(slice) * es + 127) & ~static_cast<size_t>(127));
  p.pipe_groups = plan->groups;
  p.pipe_buffers = plan->buffers;
  p.pipe_cap = plan->cap;
  p.pipe_clusters = n_clusters;
  p.use_tma = 1;
  p.vec_out = (p.probs != nullptr && reinterpret_cast<uintptr_t>(p.probs) % 16 == 0 && p.ld_out % 4 == 0) ? 1 : 0;
  p.rows = rows;
  return true;
}

cudaError_t launch_norm_pipe(const NormParams& p, int dtype, int rows, cudaStream_t st) {
  return pipe_dispatch_dtype(p, dtype, rows, st, nullptr);
}

}  // namespace sd
