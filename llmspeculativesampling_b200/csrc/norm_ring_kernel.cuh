// Kernel 1, ring version: ONE persistent CTA per SM streams whole logits rows through a shared-memory ring of 16 KB
// chunks (1-D TMA bulk copies), no thread-block cluster — every SM is usable whatever the vocabulary, and all 16
// compute warps of the CTA work on the same row, so the latency between "last byte landed" and "row finished" is that
// of one short selection / normalisation phase instead of a 4-warp group's.
//
//   warp 16  (loader)   draws rows (static first row, then a global ticket counter), publishes them, and issues the TMA
//                       load of every chunk into the next ring slot as soon as the 16 compute warps have released it
//   warp 17  (aux)      TOPK : writes the zeros of each output row (st.global.cs.v4) ahead of the compute warps
//                       DENSE: the sampler — picks the token of a finished row from the exact per-piece weight sums
//   warps 0..15         TOPK : pass 1 (three largest vector maxima per thread) runs chunk by chunk AS THE CHUNKS LAND;
//                              pivot = k-th largest of the 512 thread maxima; re-scan of the hot threads' vectors; the
//                              ring slots are released; merge, rank sort, top-k / top-p / softmax / sample on the ~k
//                              candidates; scatter over the zero-filled row
//                       DENSE: pass A = per-thread online (max, sum of exp) chunk by chunk as the chunks land, one
//                              CTA-wide combine; pass B = exp2 of every element, 16-byte streaming stores, ring slots
//                              released chunk by chunk; the sampler's 64-bit fixed-point weights are accumulated as two
//                              20-bit limbs with round-down float adds (no float -> u64 conversion per element)
//
// Rows the fast top-k selection cannot serve (massive ties) are deferred to the general path (norm_row) and re-run by
// the same CTA after its ring has drained, exactly as in the cluster pipeline (norm_pipe_kernel.cuh).
//
// Replaces /root/reference/sampling/utils.py:152-210 (+ :213-233 when a uniform is supplied), called per row from
// sampling/kvcache_model.py:166-168, 235-236, 280-283.
#pragma once
#include "norm_row.cuh"

namespace sd {

constexpr int kRingComputeWarps = 16;
constexpr int kRingComputeThreads = kRingComputeWarps * 32;
constexpr int kRingThreads = kRingComputeThreads + 64;      // + loader warp + aux warp
constexpr int kRingChunkBytes = 16384;
constexpr int kRingVecPerChunk = kRingChunkBytes / 16;      // 16-byte vectors per chunk
constexpr int kRingMaxSlots = 14;
constexpr int kRingItemRing = 32;                           // row descriptors in flight (loader lead <= slots + 1)
constexpr int kRingCap = 256;                               // merged candidates per row
constexpr int kRingWarpCap = 48;                            // candidates one warp may collect per row
constexpr int kRingMaxFail = 192;                           // rows per round that may be deferred to the general path
constexpr int kRingFailSlack = 40;
constexpr int kRingMaxLongChunks = 64;                      // DENSE rows longer than the ring (streamed twice): up to 1 MB per row
constexpr int kRingMaxPieces = kRingMaxLongChunks * kRingComputeWarps;   // (chunk, warp) pieces of 64 vectors
constexpr int kRingEndDone = -1, kRingEndPause = -2;
constexpr uint32_t kRingTieUlps = 8;

enum : int { kRingTopK = 0, kRingDense = 1, kRingDenseT1 = 2, kRingTopKLong = 3 };
// (DenseT1: temperature == 1, no logit / T arithmetic compiled in; TopKLong: top-k rows longer than the ring — a separate
//  instantiation: its branches in pass 1 / pass 2 cost the resident-row kernel 4 % when they were run-time)

// scratch of the two modes (a union inside RingShared: a dense launch needs 5 KB next to the ring instead of 19 KB, which
// is one more 16 KB slot — fp32 rows of V = 50272 are 13 chunks)
struct alignas(16) RingTopKScratch {
  alignas(16) float tm_in[128];        // maxima of the 128 thread quads
  float tm[128];                       // the same, sorted per warp (4 lists of 32)
  uint2 w_pair[kRingComputeWarps][kRingWarpCap];     // slow path only: a warp's candidates before they are allocated
  float tau;
  int cand_cnt, cand_over;             // candidates allocated for the current row, per-warp overflow flag
  int n_sorted[2];                     // length of the sorted list handed to the aux warp (-1: row deferred)
  unsigned long long a_key[kRingCap];  // the row's candidates as sort keys (value key << 32 | ~index)
  float s_val[2][kRingCap];            // sorted candidates (logit / T, descending), double buffered by item parity (compute -> aux)
  int s_idx[2][kRingCap];
  float f_val[kRingCap];               // ... and the final probabilities of the row it is finishing
};
struct alignas(16) RingDenseScratch {
  uint64_t comb[2];                    // the 16 warp results of a row are in place (16 compute warps -> each other), by item parity
  float wm[2][kRingComputeWarps];      // (double buffered by item parity: a warp may be a whole pass ahead of the slowest reader)
  double ws[2][kRingComputeWarps];
  unsigned long long piece[2][kRingMaxPieces];
  float info_c2[2];
};

struct alignas(16) RingShared {
  uint64_t full[kRingMaxSlots];        // chunk landed                       (TMA -> compute warps)
  uint64_t empty[kRingMaxSlots];       // slot may be overwritten            (16 compute warps -> loader)
  uint64_t rowfull[kRingItemRing];     // row of item it published           (loader -> everyone)
  uint64_t sorted[2];                  // TOPK: candidates of item it complete (16 compute warps -> aux), by item parity
  uint64_t sfree[2];                   // TOPK: that candidate array may be reused    (aux -> compute)
  uint64_t row_done[2];                // DENSE: row written, piece sums ready (16 compute warps -> aux), by item parity
  uint64_t tfree[2];                   // DENSE: piece table may be reused   (aux -> compute)
  int row_of[kRingItemRing];
  int n_fail, end_reason;              // (read into registers before the general path re-purposes the memory)
  union {
    RingTopKScratch k;
    RingDenseScratch d;
  };
  // ---- kept LAST (TOPK only; survives norm_row, which re-purposes everything in front of it)
  int fail_rows[kRingMaxFail];
};
// bytes of RingShared a dense launch touches
constexpr size_t kRingSharedDenseBytes = offsetof(RingShared, d) + sizeof(RingDenseScratch);

__device__ __forceinline__ void ring_named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void ring_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float ex2_ftz(float x) {         // MUFU.EX2 (results below 2^-126 flush to 0: far under the 2^-40 sampling resolution)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// round-down float add (full-rate FADD.RM): with c = 2^23 the low mantissa bits of the result are floor(a), 0 <= a < 2^23
__device__ __forceinline__ float fadd_rd(float a, float b) { return __fadd_rd(a, b); }

// debug timeline (tools/ring_prof.py): prof[(cta * 8 + item) * 16 + slot] = clock64() for the first 8 items of every CTA
// Compiled in only with -DSD_RING_PROF (python tools/ring_prof.py builds its own copy of the library): carrying the probes
// costs the production kernels registers and issue slots.
#ifdef SD_RING_PROF
#define RING_PROF(slot) do { if (p.prof != nullptr && tid == 0 && it < 8) p.prof[(static_cast<long long>(blockIdx.x) * 8 + it) * 16 + (slot)] = clock64(); } while (0)
#define RING_PROF_AUX(slot) do { if (p.prof != nullptr && lane == 0 && it < 8) p.prof[(static_cast<long long>(blockIdx.x) * 8 + it) * 16 + (slot)] = clock64(); } while (0)
#define RING_PROF_CTA(i, expr) do { if (prof_cta != nullptr && tid == 0) prof_cta[i] = static_cast<long long>(expr); } while (0)
#else
#define RING_PROF(slot) do { } while (0)
#define RING_PROF_AUX(slot) do { } while (0)
#define RING_PROF_CTA(i, expr) do { } while (0)
#endif

template <typename T, int MODE>
__global__ void __launch_bounds__(kRingThreads, 1) norm_ring_kernel(const NormParams p) {
  constexpr int PV = Elem<T>::kPerVec;
  constexpr int CT = kRingComputeThreads;
  constexpr int CW = kRingComputeWarps;
  constexpr float kLog2e = 1.4426950408889634f;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int NS = p.ring_slots;
  // layout: NS ring slots | (TOPK: one chunk of zeros, the source of the output rows' zero fill) | RingShared
  unsigned char* zbuf = smem_raw + static_cast<size_t>(NS) * kRingChunkBytes;
  RingShared& sh = *reinterpret_cast<RingShared*>(smem_raw + p.ring_shared_off);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = static_cast<int>(p.V);
  const int n_vec = V / PV;                                   // (TMA path: V * sizeof(T) is a multiple of 16)
  const int NCH = (n_vec + kRingVecPerChunk - 1) / kRingVecPerChunk;
  const uint32_t row_bytes = static_cast<uint32_t>(n_vec) * 16u;
  const float temp = p.temperature;
  const float r_temp = 1.0f / temp;
  constexpr bool kTopK = MODE == kRingTopK || MODE == kRingTopKLong;
  const bool t1 = MODE == kRingDenseT1 ? true : (MODE == kRingDense ? false : temp == 1.0f);
  const int k_eff = min(p.top_k, V);
  const bool want_probs = p.probs != nullptr;
  const int static_rows = static_cast<int>(gridDim.x);         // first item of CTA b is row b
  auto slot_ptr = [&](int slot) { return smem_raw + static_cast<size_t>(slot) * kRingChunkBytes; };
  // debug timeline, per CTA (after the per-item area): globaltimer at entry / after the dependency wait / at exit, clock64 at entry
#ifdef SD_RING_PROF
  long long* prof_cta = p.prof != nullptr ? p.prof + (static_cast<long long>(gridDim.x) * 8 * 16 + static_cast<long long>(blockIdx.x) * 4) : nullptr;
#endif
  RING_PROF_CTA(0, globaltimer_ns());
  RING_PROF_CTA(3, clock64());

  if constexpr (kTopK) {
    for (int i = tid; i < kRingChunkBytes / 16; i += kRingThreads) reinterpret_cast<uint4*>(zbuf)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();                                       // generic writes above -> bulk-copy (async proxy) reads below
  }

  for (int round = 0;; ++round) {                              // (a second round only after a kRingEndPause)
    if (tid == 0) {
      for (int s = 0; s < NS; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], CW); }
      for (int i = 0; i < kRingItemRing; ++i) mbar_init(&sh.rowfull[i], 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&sh.row_done[i], CW); mbar_init(&sh.tfree[i], 1); mbar_init(&sh.sorted[i], 1); mbar_init(&sh.sfree[i], 1);
      }
      sh.n_fail = 0;
      sh.end_reason = kRingEndDone;
      if constexpr (kTopK) { sh.k.cand_cnt = 0; sh.k.cand_over = 0; }   // (a dense launch does not allocate the top-k scratch)
      else { mbar_init(&sh.d.comb[0], CW); mbar_init(&sh.d.comb[1], CW); }
      fence_barrier_init();
    }
    if (round > 0) {
      // the general path used the ring (and the zero chunk) through the generic proxy: restore the zeros, then order
      // everything before the bulk copies of the new round
      if constexpr (kTopK)
        for (int i = tid; i < kRingChunkBytes / 16; i += kRingThreads) reinterpret_cast<uint4*>(zbuf)[i] = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async();
    }
    __syncthreads();
    if (round == 0) pdl_wait();                                // everything above overlapped the previous kernel's tail
    // SD_RING_TRIGGER=1: dependents may be scheduled from here on (the next kernel's CTAs then take over every SM as soon as
    // this kernel's CTA on it exits and run their prologue up to their dependency wait).  Measured: 0.5 us per boundary in
    // eager back-to-back launches, nothing inside CUDA graphs, and early-launched dependents can hold SMs that a concurrent
    // independent launch on another stream could use — off by default (the trigger then comes after all work of the CTA).
    if (round == 0 && p.ring_trigger) pdl_launch_dependents();
    if (round == 0) RING_PROF_CTA(1, globaltimer_ns());

    if (warp == CW) {
      // =========================================================================== loader
      if (lane == 0) {
        auto draw = [&]() {
          if (sh.n_fail >= kRingMaxFail - kRingFailSlack) return kRingEndPause;
          const int t = static_cast<int>(atomicAdd(p.sched, 1u)) + static_rows;
          return t < p.rows ? t : kRingEndDone;
        };
        int row = round == 0 ? (static_cast<int>(blockIdx.x) < p.rows ? static_cast<int>(blockIdx.x) : kRingEndDone) : draw();
        int slot = 0, wraps = 0;
        for (int it = 0;; ++it) {
          sh.row_of[it % kRingItemRing] = row;
          ring_arrive(&sh.rowfull[it % kRingItemRing]);
          if (row < 0) { sh.end_reason = row; break; }
          const int next = draw();                             // (latency hidden behind this row's loads)
          const unsigned char* src = reinterpret_cast<const unsigned char*>(p.logits) + static_cast<size_t>(row) * p.ld_in * sizeof(T);
          // (a DENSE row longer than the ring is streamed twice: once for the max / sum pass, once for the write pass — the
          //  second read of the 148 rows in flight comes from the 126 MB L2)
          for (int cc = 0; cc < ((!kTopK && p.ring_long) ? 2 * NCH : NCH); ++cc) {
            const int c = cc >= NCH ? cc - NCH : cc;
            if (wraps > 0) mbar_wait(&sh.empty[slot], static_cast<uint32_t>(wraps - 1) & 1u);
            const uint32_t off = static_cast<uint32_t>(c) * kRingChunkBytes;
            const uint32_t bytes = min(static_cast<uint32_t>(kRingChunkBytes), row_bytes - off);
            mbar_expect_tx(&sh.full[slot], bytes);
            tma_load_1d(slot_ptr(slot), src + off, bytes, &sh.full[slot]);
            if (++slot == NS) { slot = 0; ++wraps; }
          }
          row = next;
        }
      }
    } else if (warp == CW + 1) {
      // =========================================================================== aux warp
      if constexpr (kTopK) {
        // The finisher.  For every row: (1) its zero fill — bulk copies (TMA) of the zero chunk, issued by one thread as soon
        // as the row is known, so no store instruction of the SM is spent on the ~V zeros of a top-k filtered row; (2) when
        // the compute warps hand over the sorted candidate list: top-k cut (ties kept), top-p cut, softmax, optional
        // inverse-CDF sample, compact list; (3) the scatter of the few non-zeros over the zero-filled row.  The compute
        // warps are already selecting the next row meanwhile.
        for (int it = 0;; ++it) {
          mbar_wait(&sh.rowfull[it % kRingItemRing], static_cast<uint32_t>(it / kRingItemRing) & 1u);
          const int row = *reinterpret_cast<volatile int*>(&sh.row_of[it % kRingItemRing]);
          if (row < 0) break;
          float* orow = want_probs ? p.probs + static_cast<long long>(row) * p.ld_out : nullptr;
          if (want_probs && lane == 0) {
            unsigned char* o = reinterpret_cast<unsigned char*>(orow);
            const uint32_t out_bytes = static_cast<uint32_t>(V) * 4u;
            for (uint32_t off = 0; off < out_bytes; off += kRingChunkBytes)
              tma_store_1d(o + off, zbuf, min(static_cast<uint32_t>(kRingChunkBytes), out_bytes - off));
            tma_store_commit();
          }
          const float u_row = p.u != nullptr ? __ldg(p.u + row) : -1.f;        // (requested now, needed after the softmax)
          const int par = it & 1;
          mbar_wait(&sh.sorted[par], static_cast<uint32_t>(it >> 1) & 1u);
          const int n_tot = *reinterpret_cast<volatile int*>(&sh.k.n_sorted[par]);
          RING_PROF_AUX(8);
          if (n_tot >= 0) {                                        // (< 0: the row goes to the general path)
            const float* sv = sh.k.s_val[par];
            const int* si = sh.k.s_idx[par];
            int np_out = 0;
            int nk = 0;
            const float kth = sv[k_eff - 1];
            for (int base = 0; base < n_tot; base += 32) {
              const int i = base + lane;
              const unsigned ge = __ballot_sync(0xffffffffu, i < n_tot && sv[i] >= kth);
              nk += __popc(ge);
              if (ge != 0xffffffffu) break;
            }
            if (nk <= 32) {
              const bool in_k = lane < nk;
              const float x = in_k ? sv[lane] : -INFINITY;
              const int id = in_k ? si[lane] : 0x7fffffff;
              const float M = __shfl_sync(0xffffffffu, x, 0);
              const float e = in_k ? expf(x - M) : 0.f;
              const double zs = warp_sum(static_cast<double>(e));
              int np = nk;
              if (p.top_p > 0.f) {
                const float sp = e * (1.0f / static_cast<float>(zs));
                const double cum = warp_scan_incl(static_cast<double>(sp), lane);
                const unsigned ball = __ballot_sync(0xffffffffu, in_k && static_cast<float>(cum) > p.top_p);
                if (ball) np = min(nk, __ffs(ball));
              }
              const bool in_p = lane < np;
              const double z2 = warp_sum(in_p ? static_cast<double>(e) : 0.0);
              const float logz = logf(static_cast<float>(z2));
              const float pr = in_p ? expf((x - M) - logz) : 0.f;
              if (in_p && (!(pr >= 0.f) || isinf(pr))) atomicOr(p.err_flag, kErrNanLogit);
              if (in_p) sh.k.f_val[lane] = pr;
              np_out = np;
              if (p.cmp.cnt != nullptr) {
                const long long cr = static_cast<long long>(row) * p.cmp.row_stride;
                if (np <= p.cmp.cap) {
                  if (in_p) { p.cmp.idx[cr * p.cmp.cap + lane] = id; p.cmp.val[cr * p.cmp.cap + lane] = pr; }
                  if (lane == 0) p.cmp.cnt[cr] = np;
                } else if (lane == 0) p.cmp.cnt[cr] = -1;
              }
              if (p.u != nullptr && u_row >= 0.f) {
                const int e2 = frexp_exp(__shfl_sync(0xffffffffu, pr, 0));
                const unsigned long long wi = weight_of(pr, e2);
                const unsigned long long tot = warp_sum(wi);
                if (tot == 0ull) {
                  if (lane == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.tok_out[row] = 0; }
                } else {
                  const unsigned long long target = scale_target(tot, u_to_int(u_row));
                  unsigned long long before = 0ull;
                  for (int j = 0; j < np; ++j) {
                    const int idj = __shfl_sync(0xffffffffu, id, j);
                    const unsigned long long wj = __shfl_sync(0xffffffffu, wi, j);
                    before += idj < id ? wj : 0ull;
                  }
                  const int top_id = __shfl_sync(0xffffffffu, id, 0);
                  if (in_p && wi > 0ull && target >= before && target < before + wi)
                    p.tok_out[row] = (pr < kProbGuard) ? top_id : id;
                }
              }
            } else {
              const float M = sv[0];
              double zs = 0.0;
              for (int i = lane; i < nk; i += 32) zs += static_cast<double>(expf(sv[i] - M));
              zs = warp_sum(zs);
              int np = nk;
              if (p.top_p > 0.f) {
                const float rz = 1.0f / static_cast<float>(zs);
                double run = 0.0;
                for (int base = 0; base < nk; base += 32) {
                  const int i = base + lane;
                  const float sp = i < nk ? expf(sv[i] - M) * rz : 0.f;
                  const double cum = warp_scan_incl(static_cast<double>(sp), lane) + run;
                  const unsigned ball = __ballot_sync(0xffffffffu, i < nk && static_cast<float>(cum) > p.top_p);
                  if (ball) { np = min(nk, base + __ffs(ball)); break; }
                  run = __shfl_sync(0xffffffffu, cum, 31);
                }
              }
              double z2 = 0.0;
              for (int i = lane; i < np; i += 32) z2 += static_cast<double>(expf(sv[i] - M));
              z2 = warp_sum(z2);
              const float logz = logf(static_cast<float>(z2));
              bool badp = false;
              for (int i = lane; i < np; i += 32) {
                const float pr = expf((sv[i] - M) - logz);
                badp |= !(pr >= 0.f) || isinf(pr);
                sh.k.f_val[i] = pr;
              }
              if (badp) atomicOr(p.err_flag, kErrNanLogit);
              np_out = np;
              __syncwarp();
              if (p.cmp.cnt != nullptr) {
                const long long cr = static_cast<long long>(row) * p.cmp.row_stride;
                if (np <= p.cmp.cap) {
                  for (int i = lane; i < np; i += 32) { p.cmp.idx[cr * p.cmp.cap + i] = si[i]; p.cmp.val[cr * p.cmp.cap + i] = sh.k.f_val[i]; }
                  if (lane == 0) p.cmp.cnt[cr] = np;
                } else if (lane == 0) p.cmp.cnt[cr] = -1;
              }
              if (p.u != nullptr && u_row >= 0.f) {
                const int e = frexp_exp(sh.k.f_val[0]);
                unsigned long long tot = 0ull;
                for (int i = lane; i < np; i += 32) tot += weight_of(sh.k.f_val[i], e);
                tot = warp_sum(tot);
                if (tot == 0ull) {
                  if (lane == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.tok_out[row] = 0; }
                } else {
                  const unsigned long long target = scale_target(tot, u_to_int(u_row));
                  for (int i = lane; i < np; i += 32) {
                    const int id = si[i];
                    const unsigned long long wi = weight_of(sh.k.f_val[i], e);
                    unsigned long long before = 0ull;
                    for (int j = 0; j < np; ++j) before += (si[j] < id) ? weight_of(sh.k.f_val[j], e) : 0ull;
                    if (wi > 0ull && target >= before && target < before + wi)
                      p.tok_out[row] = (sh.k.f_val[i] < kProbGuard) ? si[0] : id;
                  }
                }
              }
            }
            __syncwarp();
            RING_PROF_AUX(9);
            if (want_probs) {                                     // scatter the non-zeros over the zero-filled row
              if (lane == 0) tma_store_wait_all();                // the zeros of this row are in place
              __syncwarp();
              for (int i = lane; i < np_out; i += 32) orow[si[i]] = sh.k.f_val[i];
            }
          } else if (lane == 0) {
            if (MODE == kRingTopKLong) {
              // a row longer than the ring cannot be re-run by this CTA: flag it in the caller's workspace; the follow-up
              // launch of the one-cluster-per-row kernel (general path) serves the flagged rows
              atomicOr(p.defer_bitmap + (row >> 5), 1u << (row & 31));
            } else {
              sh.fail_rows[atomicAdd(&sh.n_fail, 1)] = row;       // deferred to the general path (runs after the ring has drained)
            }
            if (want_probs) tma_store_wait_all();                 // its zero fill must not land after the general path's writes
          }
          __syncwarp();
          RING_PROF_AUX(10);
          if (lane == 0) ring_arrive(&sh.sfree[par]);
        }
      } else {
        // ---- DENSE sampler: token of row `row` from the exact per-piece weight sums the compute warps left behind
        const int n_pieces = NCH * CW;
        for (int it = 0;; ++it) {
          mbar_wait(&sh.rowfull[it % kRingItemRing], static_cast<uint32_t>(it / kRingItemRing) & 1u);
          const int row = *reinterpret_cast<volatile int*>(&sh.row_of[it % kRingItemRing]);
          if (row < 0) break;
          const int par = it & 1;
          mbar_wait(&sh.row_done[par], static_cast<uint32_t>(it >> 1) & 1u);
          const bool do_sample = p.u != nullptr && p.u[row] >= 0.f;
          if (do_sample) {
            const volatile unsigned long long* tab = sh.d.piece[par];
            const float c2 = *reinterpret_cast<volatile float*>(&sh.d.info_c2[par]);
            const int e = frexp_exp(ex2_ftz(c2));
            const int ppl = (n_pieces + 31) / 32;               // pieces per lane (contiguous range)
            unsigned long long mine = 0ull;
            for (int j = lane * ppl; j < min(n_pieces, (lane + 1) * ppl); ++j) mine += tab[j];
            const unsigned long long incl = warp_scan_incl(mine, lane);
            const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
            const float* orow = p.probs + static_cast<long long>(row) * p.ld_out;
            if (total == 0ull) {
              if (lane == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.tok_out[row] = 0; }
            } else {
              const unsigned long long target = scale_target(total, u_to_int(p.u[row]));
              const unsigned ball = __ballot_sync(0xffffffffu, incl > target);
              const int owner = __ffs(ball) - 1;                // first lane whose range crosses the target
              int pc = 0;
              unsigned long long base = 0ull;
              if (lane == owner) {
                base = incl - mine;
                pc = lane * ppl;
                while (base + tab[pc] <= target) { base += tab[pc]; ++pc; }
              }
              pc = __shfl_sync(0xffffffffu, pc, owner);
              base = __shfl_sync(0xffffffffu, base, owner);
              // piece pc = 64 input vectors = 64 * PV elements, contiguous in the vocabulary; walk it 128 elements at a time
              const int c = pc / CW, w = pc - c * CW;
              const int e0 = (c * kRingVecPerChunk + w * 64) * PV;
              int found = -1;
              float pfound = 1.f;
              unsigned long long run = base;
              for (int blk = 0; blk < 64 * PV && found < 0; blk += 128) {
                const int i0 = e0 + blk + lane * 4;
                float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i0 < V) pr = __ldcg(reinterpret_cast<const float4*>(orow + i0));
                const float pv[4] = {pr.x, pr.y, pr.z, pr.w};
                unsigned long long wv[4], vs = 0ull;
#pragma unroll
                for (int j = 0; j < 4; ++j) { wv[j] = weight_of(pv[j], e); vs += wv[j]; }
                const unsigned long long inc2 = warp_scan_incl(vs, lane) + run;
                const unsigned b2 = __ballot_sync(0xffffffffu, inc2 > target);
                if (b2) {
                  const int ow = __ffs(b2) - 1;
                  int f = -1;
                  float pf = 1.f;
                  if (lane == ow) {
                    unsigned long long cc = inc2 - vs;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { cc += wv[j]; if (f < 0 && cc > target) { f = i0 + j; pf = pv[j]; } }
                  }
                  found = __shfl_sync(0xffffffffu, f, ow);
                  pfound = __shfl_sync(0xffffffffu, pf, ow);
                }
                run = __shfl_sync(0xffffffffu, inc2, 31);
              }
              if (found >= 0 && pfound < kProbGuard) {
                // utils.py:228-230: the reference falls back to argmax(probs) — rare (probability < ~1e-9 * V per
                // row): one warp scans the row it just wrote
                unsigned long long best = 0ull;
                for (int i0 = lane * 4; i0 < V; i0 += 128) {
                  const float4 pr = __ldcg(reinterpret_cast<const float4*>(orow + i0));
                  const float pv[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const unsigned long long pk = (static_cast<unsigned long long>(f2key(pv[j])) << 32) | (0xffffffffu - static_cast<uint32_t>(i0 + j));
                    best = pk > best ? pk : best;
                  }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
                found = static_cast<int>(0xffffffffu - static_cast<uint32_t>(best & 0xffffffffull));
              }
              if (lane == 0) {
                if (found >= 0) p.tok_out[row] = found;
                else { atomicOr(p.err_flag, kErrEmptyRow); p.tok_out[row] = 0; }
              }
            }
          }
          __syncwarp();
          if (lane == 0) ring_arrive(&sh.tfree[par]);
        }
      }
    } else {
      // =========================================================================== compute warps
      int slot0 = 0, wraps0 = 0;                                // ring position of the current row's first chunk
      // DENSE: pass A of row it + 1 runs inside pass B of row it (see below); its per-thread results are carried over
      int cur_slot = 0, cur_wraps = 0;                          // DENSE long rows: ring cursor (chunks are consumed strictly in ring order)
      // rows longer than the ring (DENSE: streamed twice, run-time; TOPK: pass 2 from L2, compile-time)
      const bool long_rows = MODE == kRingTopKLong ? true : (MODE == kRingTopK ? false : p.ring_long != 0);
      float car_m = -INFINITY, car_nan = -INFINITY, car_u = -1.f;
      f32x2 car_s2 = 0ull;
      bool carried = false;
      for (int it = 0;; ++it) {
        mbar_wait(&sh.rowfull[it % kRingItemRing], static_cast<uint32_t>(it / kRingItemRing) & 1u);
        const int row = *reinterpret_cast<volatile int*>(&sh.row_of[it % kRingItemRing]);
        if (row < 0) break;
        float* orow = want_probs ? p.probs + static_cast<long long>(row) * p.ld_out : nullptr;
        auto slot_of = [&](int c) { const int s = slot0 + c; return s >= NS ? s - NS : s; };
        auto wait_chunk = [&](int c) {
          const int s = slot0 + c;
          const bool wrapped = s >= NS;
          mbar_wait(&sh.full[wrapped ? s - NS : s], static_cast<uint32_t>(wraps0 + (wrapped ? 1 : 0)) & 1u);
        };

        RING_PROF(0);
        if constexpr (kTopK) {
          // ---- pass 1, chunk by chunk as the chunks land: the three largest VECTOR maxima of the thread (tmax >= m2 >= m3)
          //      and the vector indices of the first two.  Thread t owns vectors c * 1024 + t and c * 1024 + 512 + t.
          float tmax = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
          int i1 = 0, i2 = 0;
          float nan_acc = -INFINITY;
          auto scan_chunk = [&](int c) {
            int sl;
            if (long_rows) {                                     // chunks are consumed strictly in ring order
              sl = cur_slot;
              mbar_wait(&sh.full[sl], static_cast<uint32_t>(cur_wraps) & 1u);
              if (++cur_slot == NS) { cur_slot = 0; ++cur_wraps; }
            } else {
              wait_chunk(c);
              sl = slot_of(c);
            }
            const uint4* s4 = reinterpret_cast<const uint4*>(slot_ptr(sl));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int vl = h * CT + tid;
              const int v = c * kRingVecPerChunk + vl;
              float vm = -INFINITY;
              if (v < n_vec) vm = vec_max_nan<T>(-INFINITY, s4[vl]);
              asm("max.NaN.f32 %0, %0, %1;" : "+f"(nan_acc) : "f"(vm));
              const float lo1 = fminf(tmax, vm);
              const bool c1 = vm > tmax;
              tmax = fmaxf(tmax, vm);
              const bool c2 = lo1 > m2;
              m3 = fmaxf(m3, fminf(m2, lo1));
              m2 = fmaxf(m2, lo1);
              i2 = c1 ? i1 : (c2 ? v : i2);
              i1 = c1 ? v : i1;
            }
            if (long_rows) {                                     // the row does not stay resident: pass 2 re-reads its few hot
              __syncwarp();                                      // vectors from global memory (L2)
              if (lane == 0) ring_arrive(&sh.empty[sl]);
            }
          };
          // The pivot is taken EARLY, from the chunks that have landed when about a quarter of the row is still in flight:
          // the k-th largest quad maximum of any part of the row is a value that at least k elements reach, so it is a valid
          // (slightly less selective) pivot — and finding it overlaps the wait for the row's last chunks instead of sitting
          // on the critical path behind them.
          const int c_piv = NCH - 1 - min(p.ring_early, NCH - 1);   // (ring_early chunks are scanned while warps 0-3 find the pivot)
          for (int c = 0; c <= c_piv; ++c) scan_chunk(c);
          RING_PROF(1);

          // ---- pivot = k-th largest of the 128 QUAD maxima (max over four neighbouring threads): at least k elements of the
          //      row reach it (one per quad).  Warps 0-3 sort the 128 values (one bitonic sort of 32 each) and rank the heads
          //      of the four lists by binary search; the other twelve warps go straight on with the row's remaining chunks.
          {
            float q = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
            q = fmaxf(q, __shfl_xor_sync(0xffffffffu, q, 2));
            if ((lane & 3) == 0) sh.k.tm_in[warp * 8 + (lane >> 2)] = q;
          }
          ring_named_bar(1, CT);
          // (every thread has read the previous row's count by now.  Done by a thread of a warp that does NOT sort below: a
          //  single-lane branch in front of the shuffles made warp 0 diverge there — measured 5k cycles for a 0.5k sort)
          if (tid == CT - 1) { sh.k.cand_cnt = 0; sh.k.cand_over = 0; }
          if (warp < 4) {
            const float sv = warp_sort_desc(sh.k.tm_in[warp * 32 + lane], lane);
            sh.k.tm[warp * 32 + lane] = sv;
            ring_named_bar(2, 128);
            const int kk = min(k_eff, 32);
            if (tid < 4 * kk) {                                  // element (list ew, position ej), one per thread
              const int ew = tid / kk, ej = tid - ew * kk;
              const float ev = sh.k.tm[ew * 32 + ej];
              int lo[4], hi[4];
#pragma unroll
              for (int w = 0; w < 4; ++w) { lo[w] = 0; hi[w] = 32; }
#pragma unroll
              for (int st = 0; st < 6; ++st) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  const int mid = (lo[w] + hi[w]) >> 1;
                  const float y = sh.k.tm[w * 32 + min(mid, 31)];
                  const bool before = mid < 32 && ((y > ev) || (y == ev && w < ew));
                  lo[w] = before ? mid + 1 : lo[w];
                  hi[w] = before ? hi[w] : mid;
                }
              }
              int rank = ej;
#pragma unroll
              for (int w = 0; w < 4; ++w) rank += (w == ew) ? 0 : lo[w];
              if (rank == k_eff - 1) sh.k.tau = ev;                // (exactly one value has this rank: k_eff <= 128)
            }
          }
          RING_PROF(4);
          for (int c = c_piv + 1; c < NCH; ++c) scan_chunk(c);    // the rest of pass 1 while the pivot settles
          if (nan_acc != nan_acc || tmax == INFINITY) { atomicOr(p.err_flag, kErrNanLogit); tmax = INFINITY; }
          ring_named_bar(1, CT);
          const float tau = float_down(sh.k.tau, t1 ? 0u : kRingTieUlps);
          RING_PROF(2);

          // ---- pass 2: every element >= pivot of the threads whose maximum reaches it goes straight into the row's candidate
          //      array (sort keys: value key << 32 | ~index); a warp allocates its slots with one shared-memory atomic
          const int par = it & 1;
          const uint4* grow = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.logits) +
                                                             static_cast<size_t>(row) * p.ld_in * sizeof(T));
          auto vec_at = [&](int v) {
            return long_rows ? ld_nc_v4(grow + v)
                             : reinterpret_cast<const uint4*>(slot_ptr(slot_of(v >> 10)))[v & (kRingVecPerChunk - 1)];
          };
          static_assert(kRingVecPerChunk == 1024, "vec_at assumes 1024 vectors per chunk");
          auto key_of = [&](float logit, int idx) {
            const float xv = __fdiv_rn(logit, temp) + 0.0f;        // logit / T;  -0 -> +0: equal values tie on the index
            return (static_cast<unsigned long long>(f2key(xv)) << 32) | (0xffffffffu - static_cast<uint32_t>(idx));
          };
          {
            const bool hot = tmax >= tau;
            const bool slow = hot && m3 >= tau;
            const bool quick = hot && !slow;
            const bool two = quick && m2 >= tau;
            float o1[PV], o2[PV];
            int c = 0;
            if (quick) {
              Elem<T>::unpack(vec_at(i1), o1);
#pragma unroll
              for (int j = 0; j < PV; ++j) c += (o1[j] >= tau) ? 1 : 0;
              if (two) {
                Elem<T>::unpack(vec_at(i2), o2);
#pragma unroll
                for (int j = 0; j < PV; ++j) c += (o2[j] >= tau) ? 1 : 0;
              }
            }
            const int incl = warp_scan_incl(c, lane);
            const int wq = __shfl_sync(0xffffffffu, incl, 31);    // candidates of the quick threads
            // a thread whose THIRD vector also reaches the pivot (rare): the warp walks all of that thread's vectors and stages
            // what it finds in its private region
            int ws = 0;
            unsigned hm = __ballot_sync(0xffffffffu, slow);
            while (hm) {
              const int t = warp * 32 + (__ffs(hm) - 1);
              hm &= hm - 1;
              for (int i0 = 0; i0 < 2 * NCH; i0 += 32) {
                const int i = i0 + lane;                         // i-th vector of thread t: chunk i / 2, half i % 2
                const int v = (i >> 1) * kRingVecPerChunk + (i & 1) * CT + t;
                const bool inb = i < 2 * NCH && v < n_vec;
                uint4 raw = make_uint4(0u, 0u, 0u, 0u);
                if (inb) raw = vec_at(v);
                const float vmax = inb ? vec_max_nan<T>(-INFINITY, raw) : -INFINITY;
                unsigned vmk = __ballot_sync(0xffffffffu, vmax >= tau);
                float o[PV];
                if (vmk) Elem<T>::unpack(raw, o);
                while (vmk) {
                  const int src = __ffs(vmk) - 1;
                  vmk &= vmk - 1;
#pragma unroll
                  for (int j = 0; j < PV; ++j) {
                    const float val = __shfl_sync(0xffffffffu, o[j], src);
                    const int idx = __shfl_sync(0xffffffffu, v, src) * PV + j;
                    if (val >= tau) {
                      if (lane == 0 && ws < kRingWarpCap) sh.k.w_pair[warp][ws] = make_uint2(__float_as_uint(val), static_cast<uint32_t>(idx));
                      ++ws;
                    }
                  }
                }
              }
            }
            __syncwarp();
            int base = 0;
            if (lane == 0 && wq + ws > 0) {
              base = atomicAdd(&sh.k.cand_cnt, wq + ws);
              if (ws > kRingWarpCap) sh.k.cand_over = 1;
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            if (c > 0) {
              int w = base + incl - c;
#pragma unroll
              for (int j = 0; j < PV; ++j)
                if (o1[j] >= tau) { if (w < kRingCap) sh.k.a_key[w] = key_of(o1[j], i1 * PV + j); ++w; }
              if (two) {
#pragma unroll
                for (int j = 0; j < PV; ++j)
                  if (o2[j] >= tau) { if (w < kRingCap) sh.k.a_key[w] = key_of(o2[j], i2 * PV + j); ++w; }
              }
            }
            for (int i = lane; i < min(ws, kRingWarpCap); i += 32)
              if (base + wq + i < kRingCap) sh.k.a_key[base + wq + i] = key_of(__uint_as_float(sh.k.w_pair[warp][i].x), static_cast<int>(sh.k.w_pair[warp][i].y));
          }
          // the row's logits are no longer needed: hand the ring slots back to the loader
          __syncwarp();
          if (lane == 0 && !long_rows)
            for (int c = 0; c < NCH; ++c) ring_arrive(&sh.empty[slot_of(c)]);
          // the aux warp must be done with the sorted list of item it - 2 before the rank sort below overwrites it
          if (it >= 2) mbar_wait(&sh.sfree[par], (static_cast<uint32_t>(it >> 1) - 1u) & 1u);
          ring_named_bar(1, CT);
          RING_PROF(12);
          const int n_tot = sh.k.cand_cnt;
          // self-check: every element >= pivot collected (no overflow), at least k of them; else the general path
          const bool ok = sh.k.cand_over == 0 && n_tot <= kRingCap && n_tot >= k_eff;
          if (ok) {
            // ---- rank sort on the 64-bit keys (value descending, then vocabulary index ascending): four threads per candidate
            for (int b0 = 0; b0 < n_tot; b0 += CT / 4) {
              const int i = b0 + (tid >> 2);
              const bool live = i < n_tot;
              const unsigned long long ki = live ? sh.k.a_key[i] : 0ull;
              int r = 0;
              if (live) {
#pragma unroll 4
                for (int j = tid & 3; j < n_tot; j += 4) r += sh.k.a_key[j] > ki ? 1 : 0;
              }
              r += __shfl_xor_sync(0xffffffffu, r, 1);
              r += __shfl_xor_sync(0xffffffffu, r, 2);
              if (live && (tid & 3) == 0) {
                sh.k.s_val[par][r] = key2f(static_cast<uint32_t>(ki >> 32));
                sh.k.s_idx[par][r] = static_cast<int>(0xffffffffu - static_cast<uint32_t>(ki & 0xffffffffull));
              }
            }
          }
          ring_named_bar(1, CT);                                // sorted list complete
          if (tid == 0) {
            sh.k.n_sorted[par] = ok ? n_tot : -1;
            ring_arrive(&sh.sorted[par]);                        // hand the row over to the aux warp and go on
          }
          RING_PROF(7);
        } else {
          // =================================================================== DENSE (top_k = 0, top_p = 0)
          // thread -> vectors: warp w owns vectors c * 1024 + w * 64 + {lane, 32 + lane} of chunk c: one contiguous
          // "piece" of 64 vectors per (chunk, warp), the unit of the sampler's exact prefix sums.  A thread handles its
          // 2 * PV elements of a chunk TOGETHER (one maximum, one rescale, independent exponentials) and the arithmetic
          // runs on packed fp32 pairs (FFMA2 / FADD2 / FMUL2): the two passes are issue- and latency-lean enough for the
          // row period to be set by HBM, not by the SM.
          constexpr int E = 2 * PV;                              // elements of one thread per chunk
          // logit / T, correctly rounded; -inf (a masked token) is clamped to a huge negative number first so that no
          // inf - inf can appear (its probability is exactly 0 either way)
          const f32x2 rt2 = pack2(r_temp, r_temp), nt2 = pack2(-temp, -temp);
          auto xof = [&](float l) {
            l = fmaxf(l, -1.0e30f);
            const float q0 = l * r_temp;
            return fmaf(fmaf(-q0, temp, l), r_temp, q0);
          };
          auto xof2 = [&](float a, float b) {                    // the same on a pair
            const f32x2 l2 = pack2(fmaxf(a, -1.0e30f), fmaxf(b, -1.0e30f));
            const f32x2 q0 = mul2(l2, rt2);
            return fma2(fma2(q0, nt2, l2), rt2, q0);
          };
          const f32x2 l2e2 = pack2(kLog2e, kLog2e);
          const int vl0 = warp * 64 + lane;
          // ---- pass A: per-thread online (max, sum of exp2) in the log2 domain, one chunk of a row whose first chunk sits in
          //      ring slot s0 (wrap count w0).  For the CTA's first row it runs on its own, as the chunks land; for every
          //      later row it is interleaved, chunk by chunk, with pass B of the row before (below) — the SM then reads and
          //      writes HBM at the same time instead of in alternating phases.
          float m_t = -INFINITY, nan_acc = -INFINITY;
          f32x2 s2 = pack2(0.f, 0.f);
          auto pass_a = [&](int c, int s0, int w0, bool release) {
            int sl = s0 + c;
            const bool wrapped = sl >= NS;
            if (wrapped) sl -= NS;
            mbar_wait(&sh.full[sl], static_cast<uint32_t>(w0 + (wrapped ? 1 : 0)) & 1u);
            const uint4* s4 = reinterpret_cast<const uint4*>(slot_ptr(sl));
            uint4 r0 = s4[vl0], r1 = s4[vl0 + 32];
            if (c == NCH - 1) {                                  // the row's last chunk may be partial
              constexpr uint32_t ninf = Elem<T>::kNegInfWord;
              if (c * kRingVecPerChunk + vl0 >= n_vec) r0 = make_uint4(ninf, ninf, ninf, ninf);
              if (c * kRingVecPerChunk + vl0 + 32 >= n_vec) r1 = make_uint4(ninf, ninf, ninf, ninf);
            }
            // maximum of the raw logits (logit -> logit / T is monotone), NaN-propagating
            float vm_raw = vec_max_nan<T>(vec_max_nan<T>(-INFINITY, r0), r1);
            asm("max.NaN.f32 %0, %0, %1;" : "+f"(nan_acc) : "f"(vm_raw));
            const float vm = t1 ? vm_raw : xof(vm_raw);
            const float m_new = fmaxf(m_t, vm);                  // (a NaN leaves m_t as it is; the row is flagged below)
            const float mm = m_new == -INFINITY ? 0.f : m_new;
            // (x - mm first, then * log2(e): the difference is exact near the maximum whatever the magnitude of the logits — a
            //  fused x * log2(e) - mm * log2(e) loses |mm| * 2^-24 in the exponent, fatal for the -1e30 clamp of masked tokens)
            const float resc = ex2_ftz((m_t - mm) * kLog2e);     // m_t = -inf: 0 (and s2 is 0)
            m_t = m_new;
            const f32x2 nm2 = pack2(-mm, -mm);
            float o[E];
            {
              float a[PV], b[PV];
              Elem<T>::unpack(r0, a);
              Elem<T>::unpack(r1, b);
#pragma unroll
              for (int j = 0; j < PV; ++j) { o[j] = a[j]; o[PV + j] = b[j]; }
            }
            f32x2 acc[E / 2];
#pragma unroll
            for (int j = 0; j < E; j += 2) {
              const f32x2 x2 = t1 ? pack2(o[j], o[j + 1]) : xof2(o[j], o[j + 1]);
              float a0, a1;
              unpack2(mul2(add2(x2, nm2), l2e2), a0, a1);
              acc[j / 2] = pack2(ex2_ftz(a0), ex2_ftz(a1));
            }
#pragma unroll
            for (int w = E / 4; w >= 1; w >>= 1) {
#pragma unroll
              for (int j = 0; j < w; ++j) acc[j] = add2(acc[j], acc[j + w]);
            }
            s2 = fma2(s2, pack2(resc, resc), acc[0]);
            if (release) {                                       // long rows: the chunk comes back for pass B in a later slot
              __syncwarp();
              if (lane == 0) ring_arrive(&sh.empty[sl]);
            }
          };
          float u_row;
          if (carried) {                                         // scanned while the previous row was written out
            m_t = car_m; nan_acc = car_nan; s2 = car_s2; u_row = car_u;
          } else {
            u_row = p.u != nullptr ? __ldg(p.u + row) : -1.f;    // (requested now, needed after pass A)
            if (long_rows) {
              for (int c = 0; c < NCH; ++c) {
                pass_a(c, cur_slot - c, cur_wraps, true);        // (slot = cursor, no wrap arithmetic inside)
                if (++cur_slot == NS) { cur_slot = 0; ++cur_wraps; }
              }
            } else {
              for (int c = 0; c < NCH; ++c) pass_a(c, slot0, wraps0, false);
            }
          }
          if (nan_acc != nan_acc || nan_acc == INFINITY) atomicOr(p.err_flag, kErrNanLogit);
          float s_t;
          { float sa, sb; unpack2(s2, sa, sb); s_t = sa + sb; }
          RING_PROF(1);
          // ---- combine: warp, then CTA.  The 16 warp results meet on an mbarrier (arrive now, wait later): between the two a
          //      warp already scans the first chunks of the NEXT row (they sit in the ring's spare slots), so the serial
          //      reduction chain of the combine does not leave the SM's issue slots empty
          const int par = it & 1;
          {
            const float Mw = warp_max(m_t);
            const float sc = (m_t == -INFINITY) ? 0.f : s_t * ex2_ftz((m_t - Mw) * kLog2e);
            const double Sw = warp_sum(static_cast<double>(sc));
            if (lane == 0) { sh.d.wm[par][warp] = Mw; sh.d.ws[par][warp] = Sw; }
            __syncwarp();
            if (lane == 0) ring_arrive(&sh.d.comb[par]);
          }
          // the row after this one (the loader published it while it issued this row's loads): its pass A is interleaved below
          int next_row = -1;
          if (!long_rows) {
            mbar_wait(&sh.rowfull[(it + 1) % kRingItemRing], static_cast<uint32_t>((it + 1) / kRingItemRing) & 1u);
            next_row = *reinterpret_cast<volatile int*>(&sh.row_of[(it + 1) % kRingItemRing]);
          }
          int slot0n = slot0 + NCH, wraps0n = wraps0;
          if (slot0n >= NS) { slot0n -= NS; ++wraps0n; }
          m_t = -INFINITY; nan_acc = -INFINITY; s2 = pack2(0.f, 0.f);
          if (next_row >= 0) car_u = p.u != nullptr ? __ldg(p.u + next_row) : -1.f;
          int na = 0;                                            // chunks of the next row already scanned
          if (next_row >= 0) {
            // (only chunks that can land without a release from this row; 16-bit rows only: they are bound by the SM —
            //  measured -6 % at 576 rows —, fp32 rows by HBM, where delaying pass B's stores and slot releases costs 6 %)
            const int pre = sizeof(T) == 2 ? min(min(2, NS - NCH), NCH) : 0;
            for (; na < pre; ++na) pass_a(na, slot0n, wraps0n, false);
          }
          mbar_wait(&sh.d.comb[par], static_cast<uint32_t>(it >> 1) & 1u);
          // (lane w of every warp folds warp w's pair; the warp-wide butterflies have the same order in all warps, so M and z
          //  are bit-identical across the CTA)
          const float mw = lane < CW ? sh.d.wm[par][lane] : -INFINITY;
          const float M = warp_max(mw);
          const double zw = (lane < CW && mw > -INFINITY) ? sh.d.ws[par][lane] * static_cast<double>(ex2_ftz((mw - M) * kLog2e)) : 0.0;
          const double z = warp_sum(zw);
          const float logz = logf(static_cast<float>(z));
          if (!(z > 0.0) || isinf(logz) || logz != logz) { if (tid == 0) atomicOr(p.err_flag, kErrNanLogit); }   // (uniform condition)
          const float c2 = -logz * kLog2e;
          const bool do_sample = u_row >= 0.f;                   // row-uniform
          if (it >= 2) mbar_wait(&sh.tfree[par], (static_cast<uint32_t>(it >> 1) - 1u) & 1u);   // sampler is done with item it - 2
          if (warp == 0) sh.d.info_c2[par] = c2;                   // (whole warp, same value: no single-lane branch in front of the warp reductions below)
          // sampler weights  w = floor(p * 2^(40 - e)),  e = frexp exponent of the row maximum exp2(c2)
          const float scale = ldexpf(1.0f, kScaleBits - frexp_exp(ex2_ftz(c2)));
          const float scale_hi = scale * 9.5367431640625e-07f;              // 2^-20 * scale (exact)
          const f32x2 nM2 = pack2(-M, -M), c22 = pack2(c2, c2), sh2 = pack2(scale_hi, scale_hi);
          const f32x2 k23 = pack2(8388608.0f, 8388608.0f), m1 = pack2(-1.0f, -1.0f), k20 = pack2(1048576.0f, 1048576.0f);
          RING_PROF(2);
          // ---- pass B: probabilities out (16-byte streaming stores), exact weight sums per piece, slots released.
          //      With fewer than 3 spare ring slots next to a row (fp32 V = 50272: none) the next row's chunk c can only be
          //      requested once this row's chunk c + NCH - NS has been released: its pass A then lags `lag` chunks behind
          //      pass B, so that a load has `lag` chunk-steps to land instead of being waited for right after its issue
          const int lag = min(NCH, max(0, 3 - (NS - NCH)));
          const bool out32 = want_probs && (reinterpret_cast<uintptr_t>(orow) & 31) == 0;
          for (int c = 0; c < NCH; ++c) {
            int slot;
            if (long_rows) {                                     // second trip of the chunk through the ring
              slot = cur_slot;
              mbar_wait(&sh.full[slot], static_cast<uint32_t>(cur_wraps) & 1u);
              if (++cur_slot == NS) { cur_slot = 0; ++cur_wraps; }
            } else {
              slot = slot_of(c);
            }
            const uint4* s4 = reinterpret_cast<const uint4*>(slot_ptr(slot));
            const int v0 = c * kRingVecPerChunk + vl0;
            bool ok0 = true, ok1 = true;
            uint4 r0 = s4[vl0], r1 = s4[vl0 + 32];
            if (c == NCH - 1) {                                  // the row's last chunk may be partial: -inf -> probability 0, weight 0
              constexpr uint32_t ninf = Elem<T>::kNegInfWord;
              ok0 = v0 < n_vec; ok1 = v0 + 32 < n_vec;
              if (!ok0) r0 = make_uint4(ninf, ninf, ninf, ninf);
              if (!ok1) r1 = make_uint4(ninf, ninf, ninf, ninf);
            }
            float o[E];
            {
              float a[PV], b[PV];
              Elem<T>::unpack(r0, a);
              Elem<T>::unpack(r1, b);
#pragma unroll
              for (int j = 0; j < PV; ++j) { o[j] = a[j]; o[PV + j] = b[j]; }
            }
            uint32_t acc_hi = 0u, acc_lo = 0u;
#pragma unroll
            for (int j = 0; j < E; j += 2) {
              const f32x2 x2 = t1 ? pack2(o[j], o[j + 1]) : xof2(o[j], o[j + 1]);
              float a0, a1;
              unpack2(fma2(add2(x2, nM2), l2e2, c22), a0, a1);   // exp((x - M) - logZ), utils.py:199
              o[j] = ex2_ftz(a0);
              o[j + 1] = ex2_ftz(a1);
              if (do_sample) {
                // floor(W), W = p * scale < 2^40, as two 20-bit limbs.  A round-down add of 2^23 leaves floor() of the
                // other operand in the low mantissa bits; every operation below is exact (kernel 1)
                const f32x2 p2 = pack2(o[j], o[j + 1]);
                const f32x2 t1f = fma2_rd(p2, sh2, k23);         // 2^23 + H,  H = floor(W / 2^20)
                const f32x2 nh = fma2(t1f, m1, k23);             // -H
                const f32x2 fr = fma2(p2, sh2, nh);              // W / 2^20 - H  in [0, 1)  (exact)
                const f32x2 t2f = fma2_rd(fr, k20, k23);         // 2^23 + floor(W - 2^20 H)
                float h0, h1, l0, l1;
                unpack2(t1f, h0, h1);
                unpack2(t2f, l0, l1);
                acc_hi += __float_as_uint(h0) + __float_as_uint(h1);
                acc_lo += __float_as_uint(l0) + __float_as_uint(l1);
              }
            }
            if (want_probs) {
              if (PV == 8 && out32) {                            // 16-bit logits: 8 floats per vector = one 32-byte sector per thread
                if (ok0) st_cs_v8(orow + static_cast<long long>(v0) * PV, o);
                if (ok1) st_cs_v8(orow + static_cast<long long>(v0 + 32) * PV, o + PV);
              } else {
#pragma unroll
                for (int j = 0; j < PV; j += 4) {
                  if (ok0) st_cs_v4(orow + static_cast<long long>(v0) * PV + j, o[j], o[j + 1], o[j + 2], o[j + 3]);
                  if (ok1) st_cs_v4(orow + static_cast<long long>(v0 + 32) * PV + j, o[PV + j], o[PV + j + 1], o[PV + j + 2], o[PV + j + 3]);
                }
              }
            }
            if (do_sample) {
              // (every element added 2^23's bit pattern to both limb sums: E * 0x4B000000 per thread, modulo 2^32)
              acc_hi -= static_cast<uint32_t>(E) * 0x4B000000u;
              acc_lo -= static_cast<uint32_t>(E) * 0x4B000000u;
              const uint32_t hs = __reduce_add_sync(0xffffffffu, acc_hi), ls = __reduce_add_sync(0xffffffffu, acc_lo);
              if (lane == 0) sh.d.piece[par][c * CW + warp] = (static_cast<unsigned long long>(hs) << 20) + ls;
            }
            __syncwarp();
            if (lane == 0) ring_arrive(&sh.empty[slot]);
            if (next_row >= 0 && c >= lag && na < NCH) { pass_a(na, slot0n, wraps0n, false); ++na; }
          }
          if (next_row >= 0)
            for (; na < NCH; ++na) pass_a(na, slot0n, wraps0n, false);
          __syncwarp();
          if (lane == 0) ring_arrive(&sh.row_done[par]);       // (release: this warp's stores and piece sums first)
          carried = next_row >= 0;
          car_m = m_t; car_nan = nan_acc; car_s2 = s2;
          RING_PROF(7);
          if (p.cmp.cnt != nullptr && tid == 0) p.cmp.cnt[static_cast<long long>(row) * p.cmp.row_stride] = -1;   // no compact list
        }
        slot0 += NCH;
        if (slot0 >= NS) { slot0 -= NS; ++wraps0; }
      }
    }

    // =============================================================================== drained: general path for deferred rows
    __syncthreads();
    const int reason = sh.end_reason;
    const int n_fail = sh.n_fail;
    if (n_fail > 0 || reason != kRingEndDone) {
      if (tid == 0) {                                           // retire the mbarriers before their memory is re-purposed / re-initialised
        for (int s = 0; s < NS; ++s) { mbar_inval(&sh.full[s]); mbar_inval(&sh.empty[s]); }
        for (int i = 0; i < kRingItemRing; ++i) mbar_inval(&sh.rowfull[i]);
        for (int i = 0; i < 2; ++i) { mbar_inval(&sh.row_done[i]); mbar_inval(&sh.tfree[i]); mbar_inval(&sh.sorted[i]); mbar_inval(&sh.sfree[i]); }
        if constexpr (!kTopK) { mbar_inval(&sh.d.comb[0]); mbar_inval(&sh.d.comb[1]); }
      }
      __syncthreads();
    }
    if (n_fail > 0) {
      NormParams p2 = p;
      p2.force_general = 1;
      p2.prof = nullptr;
      p2.cluster = 1;
      p2.slice_elems = p.ring_row_elems;
      p2.slice_smem_bytes = p.ring_row_smem_bytes;
      for (int i = 0; i < n_fail; ++i) {
        norm_row<T, kRingThreads>(p2, sh.fail_rows[i]);
        __syncthreads();
      }
    }
    if (reason == kRingEndDone) break;
  }  // rounds

  if (!p.ring_trigger) pdl_launch_dependents();
  RING_PROF_CTA(2, globaltimer_ns());
  // the last CTA to finish re-arms the row counter for the next launch that uses this scheduler block
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(p.sched + 1, 1u) == gridDim.x - 1u) {
      p.sched[0] = 0u;
      p.sched[1] = 0u;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <typename T, int MODE>
static cudaError_t ring_launch(const NormParams& p, cudaStream_t st) {
  auto kern = norm_ring_kernel<T, MODE>;
  static bool attr_set_dev[64] = {};
  int dev_id = 0;
  (void)cudaGetDevice(&dev_id);
  bool& attr_set = attr_set_dev[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.ring_smem_bytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kRingThreads);
  cfg.gridDim = dim3(static_cast<unsigned>(p.ring_ctas));
  cfg.dynamicSmemBytes = static_cast<size_t>(p.ring_smem_bytes);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

template <typename T>
static cudaError_t ring_dispatch(const NormParams& p, cudaStream_t st) {
  if (p.ring_mode != kRingDense) return p.ring_long ? ring_launch<T, kRingTopKLong>(p, st) : ring_launch<T, kRingTopK>(p, st);
  return p.temperature == 1.0f ? ring_launch<T, kRingDenseT1>(p, st) : ring_launch<T, kRingDense>(p, st);
}

}  // namespace sd
