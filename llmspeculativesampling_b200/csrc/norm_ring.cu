// Kernel 1, ring version — planning and launch (the kernel is in norm_ring_kernel.cuh, instantiated per logits dtype in
// norm_ring_f32.cu / _bf16.cu / _f16.cu).
//
// The ring kernel serves the two settings the reference actually runs with (evaluation.py uses top_k = 20,
// top_p = 0.9; the API default is top_k = 0, top_p = 0, /root/reference/sampling/speculative_sampling.py:1879-1880):
//     TOPK   0 < top_k <= 128, any top_p
//     DENSE  top_k = 0, top_p = 0
// whenever a row fits the shared-memory ring of one CTA (V * sizeof(logit) <= slots * 16 KB: every vocabulary of the
// named models: top-k 12 chunks = fp32 up to V = 49152, dense 13 chunks = fp32 up to V = 53248; bf16 / fp16 twice that), rows are 16-byte aligned and a scheduler
// workspace is given.  Everything else keeps using the cluster kernels (norm_pipe.cu, norm.cu).
#include "norm_ring_kernel.cuh"

#include <cstdlib>

namespace sd {

cudaError_t ring_dispatch_f32(const NormParams& p, cudaStream_t st);
cudaError_t ring_dispatch_bf16(const NormParams& p, cudaStream_t st);
cudaError_t ring_dispatch_f16(const NormParams& p, cudaStream_t st);

namespace {
struct DevInfo { int sms, smem_optin; bool valid; };
DevInfo g_dev[64] = {};
const DevInfo& dev_info() {
  int dev = 0;
  (void)cudaGetDevice(&dev);
  DevInfo& d = g_dev[dev & 63];
  if (!d.valid) {                          // (benign race: every thread writes the same values)
    int sms = 0, smem = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    if (cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || smem <= 0) smem = 227 * 1024;
    d.sms = sms; d.smem_optin = smem; d.valid = true;
  }
  return d;
}
}  // namespace

int device_sm_count() { return dev_info().sms; }
int device_max_smem_optin() { return dev_info().smem_optin; }

bool plan_ring(NormParams& p, int dtype, int rows) {
  if (p.sched == nullptr || p.no_ring || p.no_pipeline || p.force_general || rows < 1) return false;
  const size_t es = dtype == kF32 ? 4 : 2;
  int mode;
  if (p.top_k > 0 && p.top_k <= kFastK) mode = kRingTopK;
  else if (p.top_k <= 0 && !(p.top_p > 0.f)) mode = kRingDense;
  else return false;
  // the dense sampler re-reads the probabilities it needs from the row the compute warps just wrote
  if (mode == kRingDense && p.u != nullptr && p.probs == nullptr) return false;
  const long long row_bytes = p.V * static_cast<long long>(es);
  const bool aligned_in = (reinterpret_cast<uintptr_t>(p.logits) % 16 == 0) && ((p.ld_in * es) % 16 == 0) && (row_bytes % 16 == 0);
  const bool aligned_out = p.probs == nullptr || (reinterpret_cast<uintptr_t>(p.probs) % 16 == 0 && p.ld_out % 4 == 0 && p.V % 4 == 0);
  if (!aligned_in || !aligned_out) return false;
  const int smem_max = device_max_smem_optin();
  const size_t fixed = ((mode == kRingDense ? kRingSharedDenseBytes : sizeof(RingShared)) + 127) & ~static_cast<size_t>(127);
  const int extra = mode == kRingTopK ? 1 : 0;               // TOPK: one more chunk holds the zeros of the output rows' zero fill
  // SD_RING_SPARE=<bytes> leaves part of the SM's shared memory unallocated, so that a small kernel of another stream (the
  // sparse verify: 12.7 KB static, 128 threads) can run BESIDE the ring CTAs instead of waiting for an SM to drain.
  // Measured on the software-pipelined bench step: no gain (31.4 vs 31.0 us) — off by default.
  size_t spare = 0;
  { const char* ev = getenv("SD_RING_SPARE"); if (ev != nullptr) spare = static_cast<size_t>(atoi(ev)); }
  int slots = static_cast<int>((static_cast<size_t>(smem_max) - fixed - spare) / kRingChunkBytes) - extra;
  if (slots > kRingMaxSlots) slots = kRingMaxSlots;
  const int n_chunks = static_cast<int>((row_bytes + kRingChunkBytes - 1) / kRingChunkBytes);
  // DENSE rows longer than the ring are streamed through it twice (max / sum pass, write pass; the second read is an L2 hit)
  // TOPK rows longer than the ring: pass 1 releases every chunk at once, pass 2 fetches a thread's one or two hot vectors
  // from global memory (L2); rows the fast selection cannot serve are flagged in the workspace bitmap and re-run by a
  // follow-up launch of the one-cluster-per-row kernel (launch_norm) — needs one bit per row in the workspace
  // (measured at 576 rows: up to 16 chunks the cluster pipeline is as fast or faster — 45.5 vs 53.5 us at fp32 V = 50272,
  //  82.7 vs 85.4 us at bf16 V = 131072 —, from 32 chunks on the ring wins: 113 vs 151 us at fp32 V = 131072, 212 vs 654 us
  //  at fp32 V = 262144)
  const bool long_rows = n_chunks > slots && (mode == kRingDense || (rows <= kDeferBitmapRows && n_chunks >= 24 && rows * 4 >= device_sm_count()));   // (a handful of rows: one CTA per row starves, clusters split them)
  if (slots < 2 || (n_chunks > slots && !long_rows) || n_chunks > kRingMaxLongChunks) return false;
  p.ring_long = long_rows ? 1 : 0;
  const size_t shared_off = static_cast<size_t>(slots + extra) * kRingChunkBytes;
  // in-kernel general-path fallback (norm_row with the whole row in this CTA) must fit in front of the deferred-row list
  const long long slice = (p.V + 127) & ~127LL;
  const size_t slice_bytes = (static_cast<size_t>(slice) * es + 127) & ~static_cast<size_t>(127);
  if (mode == kRingTopK && !long_rows && slice_bytes + sizeof(NormShared<kRingThreads>) > shared_off + offsetof(RingShared, fail_rows)) return false;
  // pivot taken while the row's last chunks are still in flight: pays for rows of >= 7 chunks (measured: 8 chunks 31.9 -> 31.2 us,
  // 7 chunks 36.1 -> 35.0 us at 576 rows; rows of 4 chunks lose: the early pivot is too weak there)
  { const char* ev = getenv("SD_RING_EARLY"); p.ring_early = ev != nullptr ? atoi(ev) : (n_chunks >= 7 ? 3 : 0); }
  { const char* ev = getenv("SD_RING_TRIGGER"); p.ring_trigger = ev != nullptr ? atoi(ev) : 0; }
  p.ring_mode = mode;
  p.ring_slots = slots;
  p.ring_shared_off = static_cast<int>(shared_off);
  p.ring_smem_bytes = static_cast<int>(shared_off + fixed);
  p.ring_ctas = rows < device_sm_count() ? rows : device_sm_count();
  p.ring_row_elems = static_cast<int>(slice);
  p.ring_row_smem_bytes = static_cast<int>(slice_bytes);
  p.use_tma = 1;
  p.vec_out = 1;
  p.rows = rows;
  return true;
}

cudaError_t launch_norm_ring(const NormParams& p, int dtype, cudaStream_t st) {
  switch (dtype) {
    case kF32: return ring_dispatch_f32(p, st);
    case kBF16: return ring_dispatch_bf16(p, st);
    case kF16: return ring_dispatch_f16(p, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace sd
