// Kernel 1: fused temperature -> top-k -> top-p -> softmax (-> optional inverse-CDF sample) over the
// vocabulary, one thread-block cluster per logits row.
//
// Replaces, per row, the reference's ~25-30 ATen launches and 3 host syncs:
//   /root/reference/sampling/utils.py:152-179 (top_k_top_p_filter), :182-210 (norm_logits) and,
//   when a uniform is supplied, :213-233 (sample, with multinomial -> inverse CDF, see
//   oracle/ref_ops.py).  Callers: sampling/kvcache_model.py:166-168, 235-236, 283.
//
// Data movement (the whole point): each logits element is read from HBM exactly once — a 1-D TMA
// bulk copy (cp.async.bulk + mbarrier complete_tx) lands the CTA's slice of the row in shared
// memory — and each fp32 probability is written exactly once with 16-byte streaming stores.
// A row is split over a cluster of C CTAs (C = 1, 2, 4, 8) so that several CTAs fit per SM and the
// hardware overlaps one CTA's loads with another's selection work; CTAs of a cluster exchange a
// few scalars / candidate lists through distributed shared memory.
//
// Paths (chosen per launch, with a data-dependent fall-back from the first to the third):
//   fast top-k  (0 < top_k <= 128)   thread maxima -> pivot that is guaranteed to keep >= k elements
//                                   -> <= ~100 candidates per CTA -> exact selection, top-p and
//                                   softmax on the candidate list; the rest of the row is zeros.
//   dense       (top_k = 0, top_p = 0)  max / sum / exp passes over the staged slice.
//   general     (anything else, or candidate overflow because of massive ties)  sort-free 16-way
//                                   threshold search on order-preserving keys (count for top-k,
//                                   probability mass for top-p, index for ties).
#include "norm_row.cuh"

namespace sd {

// One cluster per row (grid = rows * C), or — as the follow-up of the pipelined kernel — a small persistent grid
// that walks all rows and only processes the flagged ones (row_filter).
template <typename T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) norm_probs_kernel(const NormParams p) {
  const int C = p.cluster;
  const int cid = blockIdx.x / C, n_cl = gridDim.x / C;
  bool first = true;
  for (int row = cid; row < p.rows; row += n_cl) {
    if (p.row_filter != nullptr && p.row_filter[row] == 0) continue;     // uniform across the cluster
    if (p.use_defer_bitmap && ((p.defer_bitmap[row >> 5] >> (row & 31)) & 1u) == 0u) continue;   // (uniform as well)
    if (!first) __syncthreads();                                          // shared memory / mbarriers are reused
    first = false;
    norm_row<T, THREADS>(p, row);
    if (p.use_defer_bitmap) {                                             // served: clear the flag for the next launch
      if (C > 1) cg::this_cluster().sync(); else __syncthreads();         // (every CTA of the cluster has read the bit)
      if (threadIdx.x == 0 && blockIdx.x % C == 0) atomicAnd(p.defer_bitmap + (row >> 5), ~(1u << (row & 31)));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
static int g_tune_cluster = 0, g_tune_threads = 0;
static long long* g_prof = nullptr;
static int g_pdl = 1;
void set_pdl(int enable) { g_pdl = enable ? 1 : 0; }
int pdl_enabled() { return g_pdl; }

void set_norm_prof(long long* ptr) { g_prof = ptr; }
long long* get_norm_prof() { return g_prof; }

void set_norm_tuning(int cluster, int threads) { g_tune_cluster = cluster; g_tune_threads = threads; }

template <typename T, int THREADS, int MINB>
static cudaError_t launch_cfg(const NormParams& p, size_t smem, int rows, cudaStream_t st) {
  auto kern = norm_probs_kernel<T, THREADS, MINB>;
  static bool attr_set_dev[64] = {};            // per device: the attribute belongs to the device's copy of the kernel
  int dev_id = 0;
  (void)cudaGetDevice(&dev_id);
  bool& attr_set = attr_set_dev[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, device_max_smem_optin());
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  int n_cl = rows;
  if (p.row_filter != nullptr || p.use_defer_bitmap) n_cl = min(rows, max(1, device_sm_count() / p.cluster));   // follow-up mode: walk the rows
  cfg.gridDim = dim3(static_cast<unsigned>(n_cl) * p.cluster);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

template <typename T>
static cudaError_t launch_typed(NormParams p, int rows, cudaStream_t st) {
  const size_t es = sizeof(T);
  const long long row_bytes = p.V * static_cast<long long>(es);
  int C = 1;
  while (C < kMaxPortableCluster && (row_bytes + C - 1) / C > 64 * 1024) C <<= 1;
  while (C < kMaxPortableCluster && static_cast<long long>(rows) * C < device_sm_count() && row_bytes / (2 * C) >= 8192) C <<= 1;
  {
    // every CTA publishes at least min(k, slice) candidates: keep the per-rank share of the merged list above that
    const int kc = p.top_k > 0 ? p.top_k : (p.top_p > 0.f ? kTopPCandidates : 0);
    while (C > 1 && kc > 0 && kc <= kFastK && kCapTotal / C < kc + kc / 2 + 16 && (row_bytes + C / 2 - 1) / (C / 2) <= 200 * 1024) C >>= 1;
  }
  if (g_tune_cluster > 0) C = g_tune_cluster;
  // slice: multiple of 128 elements so that every slice start is 16-byte aligned in both dtypes
  long long slice = ((p.V + C - 1) / C + 127) & ~127LL;
  while (C > 1 && slice * (C - 1) >= p.V) { C >>= 1; slice = ((p.V + C - 1) / C + 127) & ~127LL; }
  p.cluster = C;
  p.rows = rows;
  p.prof = g_prof;
  p.slice_elems = static_cast<int>(slice);
  const size_t slice_bytes = (static_cast<size_t>(slice) * es + 127) & ~static_cast<size_t>(127);
  p.slice_smem_bytes = static_cast<int>(slice_bytes);
  const bool aligned_in = (reinterpret_cast<uintptr_t>(p.logits) % 16 == 0) && ((p.ld_in * es) % 16 == 0) && (row_bytes % 16 == 0);
  p.use_tma = aligned_in ? 1 : 0;
  p.vec_out = (p.probs != nullptr && reinterpret_cast<uintptr_t>(p.probs) % 16 == 0 && p.ld_out % 4 == 0) ? 1 : 0;
  int threads = g_tune_threads > 0 ? g_tune_threads : (slice_bytes > 96 * 1024 ? 512 : 256);
  if (p.row_filter != nullptr && g_tune_cluster > 0) { /* tuning applies to the main kernel only */ }
  size_t smem;
  switch (threads) {
    case 512:
      smem = slice_bytes + sizeof(NormShared<512>);
      if (smem > device_max_smem_optin()) return cudaErrorInvalidValue;
      return launch_cfg<T, 512, 1>(p, smem, rows, st);
    case 1024:
      smem = slice_bytes + sizeof(NormShared<1024>);
      if (smem > device_max_smem_optin()) return cudaErrorInvalidValue;
      return launch_cfg<T, 1024, 1>(p, smem, rows, st);
    default:
      smem = slice_bytes + sizeof(NormShared<256>);
      if (smem > device_max_smem_optin()) return cudaErrorInvalidValue;
      return launch_cfg<T, 256, 3>(p, smem, rows, st);
  }
}

static cudaError_t launch_classic(const NormParams& p, int dtype, int rows, cudaStream_t st);

cudaError_t launch_norm(const NormParams& pin, int dtype, int rows, cudaStream_t st) {
  NormParams p = pin;
  p.prof = g_prof;
  p.row_filter = nullptr;
  p.fv_rows = 0;
  // ring kernel (one persistent CTA per SM, whole rows through a shared-memory ring) where a row fits one CTA;
  // else the persistent cluster pipeline where it applies (both fall back to the general path per row by themselves)
  if (g_tune_threads == 0 && g_tune_cluster == 0 && plan_ring(p, dtype, rows)) {
    cudaError_t e = launch_norm_ring(p, dtype, st);
    if (e != cudaSuccess || !(p.ring_long && p.ring_mode == 0)) return e;
    // top-k rows longer than the ring: the rows the ring kernel flagged (massive ties) are re-run on the general path by a
    // small persistent grid of the one-cluster-per-row kernel that walks the bitmap (no host round trip: always launched)
    NormParams f = pin;
    f.prof = nullptr;
    f.row_filter = nullptr;
    f.fv_rows = 0;
    f.use_defer_bitmap = 1;
    f.force_general = 1;
    return launch_classic(f, dtype, rows, st);
  }
  if (g_tune_threads == 0 && !p.no_pipeline && plan_pipe(p, dtype, rows, g_tune_cluster)) return launch_norm_pipe(p, dtype, rows, st);
  return launch_classic(p, dtype, rows, st);
}

cudaError_t launch_norm_verify(const NormParams& pin, int dtype, int rows, cudaStream_t st) {
  NormParams p = pin;
  p.prof = g_prof;
  p.row_filter = nullptr;
  // one launch: needs the pipelined kernel and compact lists on both sides (the in-kernel verify is the sparse one)
  const bool lists = p.fv.q != nullptr && p.fv.pc.cnt != nullptr && p.fv.qc.cnt != nullptr && p.fv_cnt != nullptr && p.fv_rows > 0;
  if (lists && g_tune_threads == 0 && !p.no_pipeline && plan_pipe(p, dtype, rows, g_tune_cluster)) return launch_norm_pipe(p, dtype, rows, st);
  p.fv_rows = 0;
  cudaError_t e = launch_norm(p, dtype, rows, st);
  if (e != cudaSuccess) return e;
  return launch_verify(pin.fv, st);
}

static cudaError_t launch_classic(const NormParams& p, int dtype, int rows, cudaStream_t st) {
  switch (dtype) {
    case kF32: return launch_typed<float>(p, rows, st);
    case kBF16: return launch_typed<__nv_bfloat16>(p, rows, st);
    case kF16: return launch_typed<__half>(p, rows, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace sd
