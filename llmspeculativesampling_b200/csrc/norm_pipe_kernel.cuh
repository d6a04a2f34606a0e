// Persistent top-k kernel (see norm_pipe.cu for the description): shared-memory layout, the kernel template, the table
// of compiled variants and the per-dtype launcher.  Included by norm_pipe.cu (planning) and by one translation unit per
// logits dtype (norm_pipe_f32.cu / _bf16.cu / _f16.cu) so that the 36 kernel instantiations compile in parallel.
#pragma once
#include "norm_row.cuh"
#include "verify_sparse.cuh"

namespace sd {


constexpr int kPipeGroupWarps = 4;
constexpr int kPipeGroupThreads = kPipeGroupWarps * 32;
constexpr int kPipeMaxGroups = 3;
constexpr int kPipeFastK = 128;
constexpr uint32_t kPipeTieUlps = 8;
constexpr int kPipeWarpCap = 48;       // candidates one warp may collect per work item
constexpr int kPipeMaxFail = 256;      // rows per cluster and round that may be deferred to the general path
constexpr int kPipeFailSlack = 24;     // ... minus the items that can still be in flight when the leader stops taking rows
constexpr int kPipeMaxPend = 64;       // requests per cluster and round whose fused verify needs the dense scan (served at the end)
constexpr int kPipeRowRing = 32;       // row indices of the items in flight (the leader is < 16 items ahead of a peer)

template <int CAP>
struct alignas(16) PipeGroupShared {
  float tm[kPipeGroupThreads];                       // per-warp sorted thread maxima
  uint2 w_pair[kPipeGroupWarps][kPipeWarpCap];       // candidates found by each warp (logit/T bits, index)
  int w_cnt[kPipeGroupWarps];
  int n_keep_p;
  int fv_req;                                        // fused verify: request completed by this group's last item, or -1
  float tau;
  uint2 recv_cnt2[2][kMaxCluster];                   // [item parity][cluster rank].x = candidate count (-1: general path)
  uint2 r_pair[2][CAP];                         // receive regions (logit/T bits, index), double buffered by item parity
  unsigned long long a_key[CAP];                // merged list as sort keys (value key << 32 | ~index)
  float a_val[CAP];                             // final probabilities (sorted order)
  float s_val[CAP]; int s_idx[CAP];        // sorted list
};

template <int NG, int NB, int CAP>
struct alignas(16) PipeShared {
  // per slice buffer:
  uint64_t empty[NB];                 // buffer may be overwritten            (group that re-scanned it -> memory warp)
  // per compute group (ONE waiter per barrier, so a waiter is never more than one phase behind: with NG > NB the
  // consumers of one buffer alternate between groups and a per-buffer "full" barrier would alias phases):
  uint64_t full[NG];                  // TMA bytes of the group's next item landed   (memory warp -> group)
  uint64_t zeroed[NG];                // that item's output slice is zero-filled     (memory warp -> group)
  uint64_t taken[NG];                 // group has observed full + zeroed            (group -> memory warp)
  uint64_t xbar[NG][2];               // peers' candidates landed             (remote groups -> group), by item parity
  uint64_t rowbar[kPipeRowRing];      // row index of item it (slot it % ring) landed  (leader CTA's memory warp -> peer's)
  uint2 row_in[kPipeRowRing];         // landing slots of those pushes (slot = dynamic item index % ring)
  uint2 item_row[kPipeRowRing];       // .x = row of item it (slot it % ring), written by this CTA's memory warp for its
                                      // groups; < 0: no more items (kPipeEndDone / kPipeEndPause)
  PipeGroupShared<CAP> g[NG];
  // kept LAST (survives norm_row, which re-purposes everything in front of it):
  int n_fail;                         // rows deferred to the general path in this round ...
  int end_reason;                     // ... and why the round ended
  int fail_rows[kPipeMaxFail];
  int n_pend, fv_req_cta, fv_na;      // fused verify: requests waiting for the dense scan; end-phase hand-over slots
  int2 pend[kPipeMaxPend];            // (request, accepted tokens)
};
constexpr int kPipeEndDone = -1;       // every row has been handed out
constexpr int kPipeEndPause = -2;      // the deferred-row list is nearly full: run the general path, then resume

__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive (release at cluster scope) on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, int rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// 8-byte store into the shared memory of CTA `rank` that completes (complete_tx, 8 bytes) on that CTA's mbarrier:
// data and signal travel together through the async proxy, so the sender needs NO release fence — a fence or a
// release-arrive would have to drain this SM's queue of in-flight zero-fill stores first (measured: ~14k cycles)
__device__ __forceinline__ void st_async_remote_v2(void* local_addr, uint32_t a, uint32_t b, uint64_t* local_bar, int rank) {
  uint32_t raddr, rbar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_addr)), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(local_bar)), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "r"(a), "r"(b), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  } while (!ok);
}

// debug timeline: prof[(cta * 32 + item) * 16 + slot] = clock64()   (items >= 32 are not recorded)
#ifdef SD_DEBUG_HANG
#define PIPE_PROF(item, slot, cond) do { } while (0)
#define PIPE_DBG(slot, val, cond) do { if (p.prof != nullptr && (cond)) { \
    *reinterpret_cast<volatile long long*>(p.prof + static_cast<long long>(blockIdx.x) * 16 + (slot)) = (val); __threadfence_system(); } } while (0)
#else
#define PIPE_DBG(slot, val, cond) do { } while (0)
#define PIPE_PROF(item, slot, cond) do { if constexpr (PROF) if (p.prof != nullptr && (cond) && (item) < 32) \
    p.prof[(static_cast<long long>(blockIdx.x) * 32 + (item)) * 16 + (slot)] = clock64(); } while (0)
#endif

// NG compute groups share NB slice buffers: work item i uses buffer i % NB and is processed by group i % NG.  A buffer
// is only held from the TMA issue to the end of the re-scan (~40 % of an item's latency), so NG > NB groups keep the
// buffers — i.e. the HBM pipe — busier than one group per buffer would.
// FV: with the in-kernel verify (its code costs the bf16 variants 3-10 %); PROF: with the clock64 timeline (3 % for fp32)
template <typename T, int NG, int NB, int CAP, bool FV, bool PROF>
__global__ void __launch_bounds__(32 + NG * kPipeGroupThreads, 1) norm_topk_pipe_kernel(const NormParams p) {
  constexpr int PV = Elem<T>::kPerVec;
  constexpr int GT = kPipeGroupThreads;
  constexpr int GW = kPipeGroupWarps;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PipeShared<NG, NB, CAP>& sh = *reinterpret_cast<PipeShared<NG, NB, CAP>*>(smem_raw + static_cast<size_t>(NB) * p.slice_smem_bytes);
  cg::cluster_group cluster = cg::this_cluster();
  const int C = p.cluster;
  const int crank = C > 1 ? static_cast<int>(cluster.block_rank()) : 0;
  const int cid = blockIdx.x / C, n_clusters = gridDim.x / C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = static_cast<int>(p.V);
  const long long start = static_cast<long long>(crank) * p.slice_elems;
  const int n = max(0, min(p.slice_elems, V - static_cast<int>(start)));
  const int n_vec = (n + PV - 1) / PV;
  const float temp = p.temperature;
  const int k_eff = min(p.top_k, V);
  const bool want_probs = p.probs != nullptr;
  // Rows are handed out dynamically: the first NB items of a cluster are rows cid, cid + n_clusters, ... (so that the
  // loads start without any exchange), every later item is the next ticket of a global counter, drawn by the leader
  // CTA's memory warp one item ahead and pushed to its peers.  SMs run at visibly different speeds under full HBM load
  // (lifetimes of CTAs with identical work differ by +-20 %), a static split would wait for the slowest.
  const int static_rows = NB * n_clusters;                  // rows covered by the static prefix
  if (p.prof != nullptr && tid == 0) {
    unsigned long long gt0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
    p.prof[(static_cast<long long>(blockIdx.x) * 32 + 0) * 16 + 15] = clock64();
    p.prof[(static_cast<long long>(blockIdx.x) * 32 + 2) * 16 + 15] = static_cast<long long>(gt0);
  }
  bool shook = (C == 1);

  for (int round = 0;; ++round) {                            // (a second round only after a kPipeEndPause)
  if (tid == 0) {
    for (int b = 0; b < NB; ++b) mbar_init(&sh.empty[b], GW);
    for (int g = 0; g < NG; ++g) {
      mbar_init(&sh.full[g], 1);
      mbar_init(&sh.zeroed[g], 1);
      mbar_init(&sh.taken[g], 1);
      mbar_init(&sh.xbar[g][0], 1);
      mbar_init(&sh.xbar[g][1], 1);
    }
    for (int i = 0; i < kPipeRowRing; ++i) mbar_init(&sh.rowbar[i], 1);
    sh.n_fail = 0;
    sh.n_pend = 0;
    sh.end_reason = kPipeEndDone;
    fence_barrier_init();
  }
  // later rounds: the general path just used this shared memory through the generic proxy; order those accesses
  // before the TMA (async proxy) writes of the new round
  if (round > 0) fence_proxy_async();
  __syncthreads();
  if (round == 0) {
    pdl_wait();                                        // everything above overlapped the previous kernel's tail
    // "my mbarriers are initialised": peers may only push to me (st.async) after they waited on this; the wait is
    // deferred to the first push so that the loads start immediately
    if (C > 1) cluster.barrier_arrive();
  } else if (C > 1) {
    cluster.sync();
    shook = true;
  }

  // =============================================================================== memory warp
  if (warp == NG * GW) {
    auto issue_load = [&](int it, int row) {
      const int g = it % NG;                                   // consuming group
      const int b = it % NB;                                   // slice buffer
      if (lane == 0 && n > 0) {
        const T* src = reinterpret_cast<const T*>(p.logits) + static_cast<long long>(row) * p.ld_in + start;
        const uint32_t bytes = static_cast<uint32_t>(n) * sizeof(T);
        mbar_expect_tx(&sh.full[g], bytes);
        unsigned char* dst = smem_raw + static_cast<size_t>(b) * p.slice_smem_bytes;
        for (uint32_t off = 0; off < bytes; off += 32768u)
          tma_load_1d(dst + off, reinterpret_cast<const unsigned char*>(src) + off, min(32768u, bytes - off), &sh.full[g]);
      } else if (lane == 0) {
        mbar_arrive_local(&sh.full[g]);
      }
    };
    auto zero_fill = [&](int it, int row) {
      const int g = it % NG;
      if (want_probs) {                                                     // zeros of this item's output slice
        float* o = p.probs + static_cast<long long>(row) * p.ld_out + start;
        if (p.vec_out) {
          const int nv4 = n >> 2;
          for (int v = lane; v < nv4; v += 32) st_cs_v4(o + 4 * v, 0.f, 0.f, 0.f, 0.f);
          for (int i = (nv4 << 2) + lane; i < n; i += 32) o[i] = 0.f;
        } else {
          for (int i = lane; i < n; i += 32) o[i] = 0.f;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&sh.zeroed[g]);
    };
    // row of item `it` in the static prefix (round 0 only)
    auto static_row = [&](int it) { const int r = cid + it * n_clusters; return r < p.rows ? r : kPipeEndDone; };
    const int n_static = round == 0 ? NB : 0;
    int row = n_static > 0 ? static_row(0) : 0;             // (dynamic first item: drawn below)
    bool have_row = n_static > 0;
    for (int it = 0;; ++it) {
      const int slot = it % kPipeRowRing;
      if (!have_row) {                                        // first item of a later round: nothing was prefetched
        if (crank == 0) {
          int t = 0;
          if (lane == 0) {
            t = static_cast<int>(atomicAdd(p.sched, 1u)) + static_rows;
            t = t < p.rows ? t : kPipeEndDone;
            for (int r = 1; r < C; ++r) st_async_remote_v2(&sh.row_in[0], static_cast<uint32_t>(t), 0u, &sh.rowbar[0], r);
          }
          row = __shfl_sync(0xffffffffu, t, 0);
        } else {
          if (lane == 0) { mbar_expect_tx(&sh.rowbar[0], 8u); mbar_wait_cluster(&sh.rowbar[0], 0u); }
          __syncwarp();
          row = static_cast<int>(*reinterpret_cast<volatile uint32_t*>(&sh.row_in[0].x));
        }
        have_row = true;
      }
      if (row < 0) {
        // no more items: tell each group (its next NG item slots), in item order
        if (lane == 0) {
          sh.end_reason = row;
          PIPE_DBG(0, it * 100 + 5, true);
          for (int j = 0; j < NG; ++j) {
            const int e = it + j;
            if (e >= NG) mbar_wait(&sh.taken[e % NG], (static_cast<uint32_t>(e / NG) - 1) & 1);
            sh.item_row[e % kPipeRowRing] = make_uint2(static_cast<uint32_t>(row), 0u);
            mbar_arrive_local(&sh.full[e % NG]);
          }
        }
        break;
      }
      PIPE_PROF(it, 0, lane == 0);
      PIPE_DBG(0, it * 100 + 1, lane == 0);
      if (it >= NB) mbar_wait(&sh.empty[it % NB], (static_cast<uint32_t>(it / NB) - 1) & 1);   // the re-scan of item it - NB is done
      PIPE_DBG(0, it * 100 + 2, lane == 0);
      if (it >= NG) mbar_wait(&sh.taken[it % NG], (static_cast<uint32_t>(it / NG) - 1) & 1);   // its group saw item it - NG
      PIPE_DBG(0, it * 100 + 3, lane == 0);
      PIPE_PROF(it, 1, lane == 0);
      if (lane == 0) sh.item_row[slot] = make_uint2(static_cast<uint32_t>(row), 0u);          // (released by the arrive on full[g])
      issue_load(it, row);
      // draw the next item's row while this item's zeros are written
      const int nd = it + 1 - n_static;                      // dynamic index of the next item: landing slot and phase
      const int nslot = nd % kPipeRowRing;
      const uint32_t nphase = static_cast<uint32_t>(nd / kPipeRowRing) & 1u;
      int next = 0;
      const bool next_static = it + 1 < n_static;
      if (next_static) next = static_row(it + 1);
      else if (crank == 0 && lane == 0) {
        if (sh.n_fail >= kPipeMaxFail - kPipeFailSlack || sh.n_pend >= kPipeMaxPend - kPipeFailSlack) next = kPipeEndPause;
        else {
          next = static_cast<int>(atomicAdd(p.sched, 1u)) + static_rows;
          next = next < p.rows ? next : kPipeEndDone;
        }
      }
      zero_fill(it, row);
      if (!next_static) {
        if (crank == 0) {
          if (C > 1) {
            if (!shook) { cluster.barrier_wait(); shook = true; }
            if (lane == 0)
              for (int r = 1; r < C; ++r) st_async_remote_v2(&sh.row_in[nslot], static_cast<uint32_t>(next), 0u, &sh.rowbar[nslot], r);
          }
          next = __shfl_sync(0xffffffffu, next, 0);
        } else {
          if (lane == 0) { mbar_expect_tx(&sh.rowbar[nslot], 8u); mbar_wait_cluster(&sh.rowbar[nslot], nphase); }
          __syncwarp();
          next = static_cast<int>(*reinterpret_cast<volatile uint32_t*>(&sh.row_in[nslot].x));
        }
      }
      PIPE_PROF(it, 2, lane == 0);
      PIPE_DBG(0, it * 100 + 4, lane == 0);
      row = next;
    }
  } else {
  // =============================================================================== compute groups
  const int g = warp / GW;                   // group == buffer
  const int gt = tid - g * GT;               // thread index inside the group
  const int gw = warp - g * GW;              // warp index inside the group
  PipeGroupShared<CAP>& gs = sh.g[g];

  const int bar_id = 1 + g;
  const int cap = CAP / C;
  const int vpt = (n_vec + GT - 1) / GT;

  for (int it = g;; it += NG) {
    const uint32_t use = static_cast<uint32_t>(it / NG);      // how often this group's scratch / xbar have been used
    const int par = use & 1;
    const int buf = it % NB;
    const uint4* s4 = reinterpret_cast<const uint4*>(smem_raw + static_cast<size_t>(buf) * p.slice_smem_bytes);

    // ---- pass 1: thread maxima (NaN-propagating)
    PIPE_PROF(it, 3, gt == 0);
    PIPE_DBG(1 + g, it * 100 + 1, gt == 0);
    mbar_wait(&sh.full[g], use & 1);
    const int row = static_cast<int>(*reinterpret_cast<volatile uint32_t*>(&sh.item_row[it % kPipeRowRing].x));
    if (row < 0) break;                                       // no more items (every group gets its own end marker)
    float* orow = want_probs ? p.probs + static_cast<long long>(row) * p.ld_out : nullptr;
    PIPE_PROF(it, 4, gt == 0);
    // Per thread: the three largest VECTOR maxima (tmax >= m2 >= m3) and the rounds i1, i2 of the first two — the
    // re-scan below then only has to touch vectors i1 / i2 of a thread unless its third vector also reaches the pivot.
    float tmax = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    int i1 = 0, i2 = 0;
    bool bad = false;
    // ownership: in round i thread t owns vector i*GT + ((t + i) mod GT): consecutive lanes read consecutive vectors
    // here, and a warp that later re-reads ONE thread's vectors (lane = round) also walks consecutive banks
    float nan_acc = -INFINITY;                                // NaN-propagating running maximum: NaN / +inf detection only
    constexpr int kBatch = 8;                                 // loads of a batch are issued before their results are used
    for (int i0 = 0; i0 < vpt; i0 += kBatch) {
      uint4 raw[kBatch];
      bool valid[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int i = i0 + j;
        const int v = i * GT + ((gt + i) & (GT - 1));
        valid[j] = i < vpt && v < n_vec;
        raw[j] = make_uint4(0u, 0u, 0u, 0u);
        if (valid[j]) raw[j] = s4[v];
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int i = i0 + j;
        float vm = vec_max_nan<T>(-INFINITY, raw[j]);
        vm = valid[j] ? vm : -INFINITY;
        asm("max.NaN.f32 %0, %0, %1;" : "+f"(nan_acc) : "f"(vm));
        const float lo1 = fminf(tmax, vm);                    // (fminf / fmaxf / > ignore a NaN operand: flagged through nan_acc)
        const bool c1 = vm > tmax;
        tmax = fmaxf(tmax, vm);
        const bool c2 = lo1 > m2;
        m3 = fmaxf(m3, fminf(m2, lo1));
        m2 = fmaxf(m2, lo1);
        i2 = c1 ? i1 : (c2 ? i : i2);
        i1 = c1 ? i : i1;
      }
    }
    bad = nan_acc != nan_acc;
    if (bad || tmax == INFINITY) { atomicOr(p.err_flag, kErrNanLogit); tmax = INFINITY; }

    PIPE_PROF(it, 5, gt == 0);
    // ---- pivot = k-th largest of the GT thread maxima (see norm.cu)
    const float sv = warp_sort_desc(tmax, lane);
    gs.tm[gw * 32 + lane] = sv;
    if (gt == 0) gs.tau = -INFINITY;
    named_bar(bar_id, GT);
    const int kk = min(k_eff, 32);
    if (gt < GW * kk) {                      // element (list ew, position ej), one per thread, packed into few warps
      const int ew = gt / kk, ej = gt - ew * kk;
      const float ev = gs.tm[ew * 32 + ej];
      int lo[GW], hi[GW];
#pragma unroll
      for (int w = 0; w < GW; ++w) { lo[w] = 0; hi[w] = 32; }
#pragma unroll
      for (int s = 0; s < 6; ++s) {
#pragma unroll
        for (int w = 0; w < GW; ++w) {
          const int mid = (lo[w] + hi[w]) >> 1;
          const float y = gs.tm[w * 32 + min(mid, 31)];
          const bool before = mid < 32 && ((y > ev) || (y == ev && w < ew));
          lo[w] = before ? mid + 1 : lo[w];
          hi[w] = before ? hi[w] : mid;
        }
      }
      int rank = ej;
#pragma unroll
      for (int w = 0; w < GW; ++w) rank += (w == ew) ? 0 : lo[w];
      if (rank == k_eff - 1) gs.tau = ev;
    }
    named_bar(bar_id, GT);
    PIPE_PROF(it, 6, gt == 0);
    const float tau = float_down(gs.tau, temp == 1.0f ? 0u : kPipeTieUlps);

    // ---- pass 2: collect every element >= pivot of the threads whose maximum reaches it, into a per-warp candidate
    //      region (no shared-memory atomics, no hot list, no barrier).  Usual case: the thread's candidates sit in its
    //      vectors i1 (and i2), which it re-reads itself; a thread whose THIRD vector also reaches the pivot is
    //      re-scanned completely by its warp (lane = round index: bank-conflict free thanks to the swizzled ownership).
    uint2* my_pair = gs.r_pair[par] + crank * cap;
    int wc = 0;                                               // candidates found by this warp (warp-uniform)
    {
      const bool hot = tmax >= tau;
      const bool slow = hot && m3 >= tau;
      const bool quick = hot && !slow;
      const int v1 = i1 * GT + ((gt + i1) & (GT - 1)), v2 = i2 * GT + ((gt + i2) & (GT - 1));
      const bool two = quick && m2 >= tau && v2 < n_vec;
      const int g1 = static_cast<int>(start) + v1 * PV, g2 = static_cast<int>(start) + v2 * PV;
      float o1[PV], o2[PV];
      int c = 0;
      if (quick && v1 < n_vec) {
        Elem<T>::unpack(s4[v1], o1);
#pragma unroll
        for (int j = 0; j < PV; ++j) c += (o1[j] >= tau && g1 + j < V) ? 1 : 0;
        if (two) {
          Elem<T>::unpack(s4[v2], o2);
#pragma unroll
          for (int j = 0; j < PV; ++j) c += (o2[j] >= tau && g2 + j < V) ? 1 : 0;
        }
      }
      const int incl = warp_scan_incl(c, lane);
      wc = __shfl_sync(0xffffffffu, incl, 31);
      if (c > 0) {
        int w = incl - c;
#pragma unroll
        for (int j = 0; j < PV; ++j)
          if (o1[j] >= tau && g1 + j < V) {
            if (w < kPipeWarpCap) gs.w_pair[gw][w] = make_uint2(__float_as_uint(o1[j]), static_cast<uint32_t>(g1 + j));
            ++w;
          }
        if (two) {
#pragma unroll
          for (int j = 0; j < PV; ++j)
            if (o2[j] >= tau && g2 + j < V) {
              if (w < kPipeWarpCap) gs.w_pair[gw][w] = make_uint2(__float_as_uint(o2[j]), static_cast<uint32_t>(g2 + j));
              ++w;
            }
        }
      }
      unsigned hm = __ballot_sync(0xffffffffu, slow);
      while (hm) {
        const int t = gw * 32 + (__ffs(hm) - 1);
        hm &= hm - 1;
        for (int i0 = 0; i0 < vpt; i0 += 32) {
          const int v = (i0 + lane) * GT + ((t + i0 + lane) & (GT - 1));
          const bool inb = i0 + lane < vpt && v < n_vec;
          uint4 raw = make_uint4(0u, 0u, 0u, 0u);
          if (inb) raw = s4[v];
          const float vmax = inb ? vec_max_nan<T>(-INFINITY, raw) : -INFINITY;
          unsigned vm = __ballot_sync(0xffffffffu, vmax >= tau);      // vectors holding at least one candidate
          float o[PV];
          if (vm) Elem<T>::unpack(raw, o);                            // warp-uniform
          const int gi = static_cast<int>(start) + v * PV;
          while (vm) {
            const int src = __ffs(vm) - 1;
            vm &= vm - 1;
#pragma unroll
            for (int j = 0; j < PV; ++j) {
              const float val = __shfl_sync(0xffffffffu, o[j], src);
              const int idx = __shfl_sync(0xffffffffu, gi, src) + j;
              if (val >= tau && idx < V) {                             // warp-uniform
                if (lane == 0 && wc < kPipeWarpCap) gs.w_pair[gw][wc] = make_uint2(__float_as_uint(val), static_cast<uint32_t>(idx));
                ++wc;
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) gs.w_cnt[gw] = wc;
    }
    PIPE_PROF(it, 14, gt == 0);
    // the zero-fill of this item must be observed before the buffer is handed back (see header), then release it
    PIPE_PROF(it, 7, gt == 0);
    PIPE_DBG(1 + g, it * 100 + 3, gt == 0);
    mbar_wait(&sh.zeroed[g], use & 1);
    __syncwarp();
    if (lane == 0) mbar_arrive_local(&sh.empty[buf]);
    named_bar(bar_id, GT);
    if (gt == 0) mbar_arrive_local(&sh.taken[g]);              // every warp of the group has observed full + zeroed

    PIPE_PROF(it, 8, gt == 0);
    PIPE_DBG(1 + g, it * 100 + 4, gt == 0);
    // ---- publish: push my candidates into every peer's receive region, then signal its mbarrier
    int wofs[GW + 1];
    bool w_over = false;
    wofs[0] = 0;
#pragma unroll
    for (int w = 0; w < GW; ++w) { const int c = gs.w_cnt[w]; w_over |= c > kPipeWarpCap; wofs[w + 1] = wofs[w] + c; }
    const int c_mine = wofs[GW];
    const int mine = (w_over || c_mine > cap || c_mine < min(k_eff, n)) ? -1 : c_mine;     // self-check, see norm.cu
    if (gt == 0) gs.recv_cnt2[par][crank] = make_uint2(static_cast<uint32_t>(mine), 0u);
    uint2 my_e = make_uint2(0u, 0u);                          // entry gt of my (concatenated) candidate list
    for (int i = gt; i < mine; i += GT) {
      int w = 0;
#pragma unroll
      for (int q = 1; q < GW; ++q) w += i >= wofs[q] ? 1 : 0;
      const uint2 e = gs.w_pair[w][i - wofs[w]];
      my_pair[i] = e;
      if (i == gt) my_e = e;
    }
    if (C > 1) {
      if (!shook) { cluster.barrier_wait(); shook = true; }   // every peer's mbarriers are initialised
      // fixed-size records (cap entries + the count) so that the receiver can arm its mbarrier with a known byte count
      if (gt == 0) mbar_expect_tx(&sh.xbar[g][par], static_cast<uint32_t>(C - 1) * (static_cast<uint32_t>(cap) * 8u + 8u));
      for (int r = 0; r < C; ++r) {
        if (r == crank) continue;
        if (gt < cap) st_async_remote_v2(&gs.r_pair[par][crank * cap + gt], my_e.x, my_e.y, &sh.xbar[g][par], r);
        if (gt == 0) st_async_remote_v2(&gs.recv_cnt2[par][crank], static_cast<uint32_t>(mine), 0u, &sh.xbar[g][par], r);
      }
      mbar_wait_cluster(&sh.xbar[g][par], (use >> 1) & 1);    // every peer's candidates have landed here
    }
    named_bar(bar_id, GT);

    PIPE_PROF(it, 9, gt == 0);
    PIPE_DBG(1 + g, it * 100 + 6, gt == 0);
    // ---- merge (identical in every CTA of the cluster): concatenate the C receive regions as 64-bit sort keys
    int n_tot = 0;
    bool ok = true;
    for (int r = 0; r < C; ++r) {
      const int c = static_cast<int>(gs.recv_cnt2[par][r].x);
      ok &= c >= 0;
      for (int i = gt; i < c; i += GT) {
        const uint2 e = gs.r_pair[par][r * cap + i];
        const float xv = __fdiv_rn(__uint_as_float(e.x), temp) + 0.0f;    // logit / T;  -0 -> +0: equal values tie on the index
        gs.a_key[n_tot + i] = (static_cast<unsigned long long>(f2key(xv)) << 32) | (0xffffffffu - e.y);
      }
      n_tot += max(c, 0);
    }
    ok &= n_tot >= k_eff;
    if (!ok) {                               // (keys written above are simply discarded)
      if (gt == 0) sh.fail_rows[atomicAdd(&sh.n_fail, 1)] = row;        // same decision in every CTA of the cluster
      named_bar(bar_id, GT);
      continue;
    }
    named_bar(bar_id, GT);
    PIPE_PROF(it, 12, gt == 0);

    // ---- rank sort on 64-bit keys (value descending, then vocabulary index ascending): four threads per candidate,
    //      branch-free inner loop
    for (int base = 0; base < n_tot; base += GT / 4) {
      const int i = base + (gt >> 2);
      const bool live = i < n_tot;
      const unsigned long long ki = live ? gs.a_key[i] : 0ull;
      int r = 0;
      if (live) {
#pragma unroll 4
        for (int j = gt & 3; j < n_tot; j += 4) r += gs.a_key[j] > ki ? 1 : 0;
      }
      r += __shfl_xor_sync(0xffffffffu, r, 1);
      r += __shfl_xor_sync(0xffffffffu, r, 2);
      if (live && (gt & 3) == 0) {
        gs.s_val[r] = key2f(static_cast<uint32_t>(ki >> 32));
        gs.s_idx[r] = static_cast<int>(0xffffffffu - static_cast<uint32_t>(ki & 0xffffffffull));
      }
    }
    PIPE_PROF(it, 13, gt == 0);
    named_bar(bar_id, GT);
    PIPE_PROF(it, 10, gt == 0);
    int nk = 0;
    if (gw == 0) {                           // entries >= the k-th value form a prefix of the sorted list (ties kept)
      const float kth = gs.s_val[k_eff - 1];
      for (int base = 0; base < n_tot; base += 32) {
        const int i = base + lane;
        const unsigned ge = __ballot_sync(0xffffffffu, i < n_tot && gs.s_val[i] >= kth);
        nk += __popc(ge);
        if (ge != 0xffffffffu) break;
      }
    }

    // ---- top-p cut, softmax, optional sample: first warp of the group
    if (gw == 0 && nk <= 32) {
      const bool in_k = lane < nk;
      const float x = in_k ? gs.s_val[lane] : -INFINITY;
      const int id = in_k ? gs.s_idx[lane] : 0x7fffffff;
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
      if (in_p) gs.a_val[lane] = pr;
      if (lane == 0) gs.n_keep_p = np;
      if (p.cmp.cnt != nullptr && crank == 0) {                   // compact form of the row for the sparse verify path
        const long long cr = static_cast<long long>(row) * p.cmp.row_stride;
        if (np <= p.cmp.cap) {
          if (in_p) { p.cmp.idx[cr * p.cmp.cap + lane] = id; p.cmp.val[cr * p.cmp.cap + lane] = pr; }
          if (lane == 0) p.cmp.cnt[cr] = np;
        } else if (lane == 0) p.cmp.cnt[cr] = -1;
      }
      if (p.u != nullptr && crank == 0 && p.u[row] >= 0.f) {
        const int e2 = frexp_exp(__shfl_sync(0xffffffffu, pr, 0));
        const unsigned long long wi = weight_of(pr, e2);
        const unsigned long long tot = warp_sum(wi);
        if (tot == 0ull) {
          if (lane == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.tok_out[row] = 0; }
        } else {
          const unsigned long long target = scale_target(tot, u_to_int(p.u[row]));
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
    } else if (gw == 0) {
      const float M = gs.s_val[0];
      double zs = 0.0;
      for (int i = lane; i < nk; i += 32) zs += static_cast<double>(expf(gs.s_val[i] - M));
      zs = warp_sum(zs);
      int np = nk;
      if (p.top_p > 0.f) {
        const float rz = 1.0f / static_cast<float>(zs);
        double run = 0.0;
        for (int base = 0; base < nk; base += 32) {
          const int i = base + lane;
          const float sp = i < nk ? expf(gs.s_val[i] - M) * rz : 0.f;
          const double cum = warp_scan_incl(static_cast<double>(sp), lane) + run;
          const unsigned ball = __ballot_sync(0xffffffffu, i < nk && static_cast<float>(cum) > p.top_p);
          if (ball) { np = min(nk, base + __ffs(ball)); break; }
          run = __shfl_sync(0xffffffffu, cum, 31);
        }
      }
      double z2 = 0.0;
      for (int i = lane; i < np; i += 32) z2 += static_cast<double>(expf(gs.s_val[i] - M));
      z2 = warp_sum(z2);
      const float logz = logf(static_cast<float>(z2));
      bool badp = false;
      for (int i = lane; i < np; i += 32) {
        const float pr = expf((gs.s_val[i] - M) - logz);
        badp |= !(pr >= 0.f) || isinf(pr);
        gs.a_val[i] = pr;
      }
      if (badp) atomicOr(p.err_flag, kErrNanLogit);
      if (lane == 0) gs.n_keep_p = np;
      __syncwarp();
      if (p.cmp.cnt != nullptr && crank == 0) {
        const long long cr = static_cast<long long>(row) * p.cmp.row_stride;
        if (np <= p.cmp.cap) {
          for (int i = lane; i < np; i += 32) { p.cmp.idx[cr * p.cmp.cap + i] = gs.s_idx[i]; p.cmp.val[cr * p.cmp.cap + i] = gs.a_val[i]; }
          if (lane == 0) p.cmp.cnt[cr] = np;
        } else if (lane == 0) p.cmp.cnt[cr] = -1;
      }
      if (p.u != nullptr && crank == 0 && p.u[row] >= 0.f) {
        const int e = frexp_exp(gs.a_val[0]);
        unsigned long long tot = 0ull;
        for (int i = lane; i < np; i += 32) tot += weight_of(gs.a_val[i], e);
        tot = warp_sum(tot);
        if (tot == 0ull) {
          if (lane == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.tok_out[row] = 0; }
        } else {
          const unsigned long long target = scale_target(tot, u_to_int(p.u[row]));
          for (int i = lane; i < np; i += 32) {
            const int id = gs.s_idx[i];
            const unsigned long long wi = weight_of(gs.a_val[i], e);
            unsigned long long before = 0ull;
            for (int j = 0; j < np; ++j) before += (gs.s_idx[j] < id) ? weight_of(gs.a_val[j], e) : 0ull;
            if (wi > 0ull && target >= before && target < before + wi)
              p.tok_out[row] = (gs.a_val[i] < kProbGuard) ? gs.s_idx[0] : id;
          }
        }
      }
    }
    named_bar(bar_id, GT);
    if (want_probs) {                        // scatter the non-zeros over the zero-filled slice
      const int np = gs.n_keep_p;
      for (int i = gt; i < np; i += GT) {
        const int id = gs.s_idx[i];
        if (id >= start && id < start + n) orow[id] = gs.a_val[i];
      }
    }
    named_bar(bar_id, GT);                   // group scratch is reused by the next item
    PIPE_PROF(it, 11, gt == 0);
    PIPE_DBG(1 + g, it * 100 + 7, gt == 0);
    // ---- fused verify: the group that finishes the last row of a request verifies it right away (one warp, from the
    //      compact lists the rows' clusters just wrote; a request that needs the dense scan is left for the end)
    if constexpr (FV) if (p.fv_rows > 0 && crank == 0) {
      if (gt == 0) {
        __threadfence();                                       // this row's outputs (ordered by the barrier above) first
        const int b = row / p.fv_rows;
        int done_req = -1;
        if (atomicAdd(p.fv_cnt + b, 1) == p.fv_rows - 1) { p.fv_cnt[b] = 0; __threadfence(); done_req = b; }
        gs.fv_req = done_req;
      }
      named_bar(bar_id, GT);
      const int b = gs.fv_req;
      if (b >= 0 && gw == 0) {
        const long long t_v0 = clock64();
        const int na = sparse_verify_warp(p.fv, b, lane, reinterpret_cast<SparseVerifyScratch*>(gs.a_key));
        if (p.prof != nullptr && lane == 0) {                  // debug timeline: duration and end time of the in-kernel verify
          p.prof[(static_cast<long long>(blockIdx.x) * 32 + 31) * 16 + g] = clock64() - t_v0;
          p.prof[(static_cast<long long>(blockIdx.x) * 32 + 31) * 16 + 4 + g] = clock64();
        }
        if (na >= 0 && lane == 0) sh.pend[atomicAdd(&sh.n_pend, 1)] = make_int2(b, na);
      }
    }
  }
  }  // compute groups

  // =============================================================================== drained: general path for deferred rows
  PIPE_DBG(5 + (warp == NG * GW ? 4 : warp / GW), 9, lane == 0 && (warp % GW == 0 || warp == NG * GW));
  if (!shook) { cluster.barrier_wait(); shook = true; }
  __syncthreads();
  PIPE_DBG(10, 1, tid == 0);
  const int reason = sh.end_reason;
  if (round == 0 && p.prof != nullptr && tid == 0) {
    unsigned long long gt1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
    p.prof[(static_cast<long long>(blockIdx.x) * 32 + 1) * 16 + 15] = clock64();
    p.prof[(static_cast<long long>(blockIdx.x) * 32 + 3) * 16 + 15] = static_cast<long long>(gt1);
  }
  const int n_fail = sh.n_fail;                 // identical in every CTA of the cluster (same rows, same decisions) ...
  if (n_fail > 0 || reason != kPipeEndDone) {
    // the pipeline's mbarriers are retired before the general path re-purposes their memory (and before a later round
    // initialises them again): nothing can touch them any more — every warp of this CTA is here, and a peer only signals
    // barriers of items this CTA has finished
    if (C > 1) cluster.sync();
    if (tid == 0) {
      for (int b = 0; b < NB; ++b) mbar_inval(&sh.empty[b]);
      for (int g = 0; g < NG; ++g) {
        mbar_inval(&sh.full[g]); mbar_inval(&sh.zeroed[g]); mbar_inval(&sh.taken[g]);
        mbar_inval(&sh.xbar[g][0]); mbar_inval(&sh.xbar[g][1]);
      }
      for (int i = 0; i < kPipeRowRing; ++i) mbar_inval(&sh.rowbar[i]);
    }
    __syncthreads();
  }
  if (n_fail > 0) {
    // ... but appended in racing order: sort, so that the CTAs of a cluster walk the rows in lock step
    int mine = 0, rank = 0;
    if (tid < n_fail) {
      mine = sh.fail_rows[tid];
      for (int j = 0; j < n_fail; ++j) rank += sh.fail_rows[j] < mine ? 1 : 0;
    }
    __syncthreads();
    if (tid < n_fail) sh.fail_rows[rank] = mine;
    __syncthreads();                            // (peers have drained too — cluster.sync above — so shared memory may be re-purposed)
    NormParams p2 = p;
    p2.force_general = 1;
    p2.prof = nullptr;
#ifdef SD_DEBUG_HANG
    if (p.prof != nullptr) p2.prof = p.prof + 256 * 16;
#endif
    for (int i = 0; i < n_fail; ++i) {
      const int frow = sh.fail_rows[i];
      PIPE_DBG(10, 1000 + i * 10 + 3, tid == 0);
      PIPE_DBG(11, frow, tid == 0);
      norm_row<T, 32 + NG * kPipeGroupThreads>(p2, frow);
      __syncthreads();
      PIPE_DBG(10, 1000 + i * 10 + 4, tid == 0);
      if constexpr (FV) if (p.fv_rows > 0 && crank == 0) {   // fused verify: count the row; the leader CTA verifies a completed request
        constexpr int TH = 32 + NG * kPipeGroupThreads;
        if (tid == 0) {
          __threadfence();
          const int b = frow / p.fv_rows;
          int done_req = -1;
          if (atomicAdd(p.fv_cnt + b, 1) == p.fv_rows - 1) { p.fv_cnt[b] = 0; __threadfence(); done_req = b; }
          sh.fv_req_cta = done_req;
        }
        __syncthreads();
        const int b = sh.fv_req_cta;
        if (b >= 0) {
          RowScratch<TH>* rs_cta = reinterpret_cast<RowScratch<TH>*>(smem_raw);
          SparseVerifyScratch* sc_cta = reinterpret_cast<SparseVerifyScratch*>(smem_raw + ((sizeof(RowScratch<TH>) + 15) & ~size_t(15)));
          if (warp == 0) {
            const int na = sparse_verify_warp(p.fv, b, lane, sc_cta);
            if (lane == 0) sh.fv_na = na;
          }
          __syncthreads();
          if (sh.fv_na >= 0) dense_verify_cta<TH>(p.fv, b, sh.fv_na, rs_cta);
          __syncthreads();
        }
      }
    }
  }
  if constexpr (FV) if (p.fv_rows > 0 && crank == 0 && sh.n_pend > 0) {   // requests whose lists were unavailable: dense scan by the CTA
    constexpr int TH = 32 + NG * kPipeGroupThreads;
    const int n_pend = sh.n_pend;
    for (int i = 0; i < n_pend; ++i) {
      dense_verify_cta<TH>(p.fv, sh.pend[i].x, sh.pend[i].y, reinterpret_cast<RowScratch<TH>*>(smem_raw));
      __syncthreads();
    }
  }
  if (reason == kPipeEndDone) break;
  }  // rounds

  // Dependents are only triggered here, after ALL work of the CTA: an earlier griddepcontrol.launch_dependents (before
  // the general-path phase) hung the kernel on B200 when that phase was long (measured; the trigger buys nothing anyway,
  // the dependent's prologue overlaps the slowest CTA's tail either way).
  PIPE_DBG(10, 6, tid == 0);
  pdl_launch_dependents();
  // the last cluster to finish re-arms the row counter for the next launch that uses this scheduler block
  if (crank == 0 && tid == 0) {
    __threadfence();
    if (atomicAdd(p.sched + 1, 1u) == static_cast<unsigned>(n_clusters) - 1u) {
      p.sched[0] = 0u;
      p.sched[1] = 0u;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The kernel variants that are compiled: (compute groups, slice buffers, merged-candidate capacity)
struct PipeVariant { int ng, nb, cap; size_t fixed, mask_off, row_need; };
#define SD_PIPE_VARIANT(NG, NB, CAP) \
  {NG, NB, CAP, sizeof(PipeShared<NG, NB, CAP>), offsetof(PipeSharedAlias##NG##NB##CAP, n_fail), sizeof(NormShared<32 + NG * kPipeGroupThreads>)}
using PipeSharedAlias43128 = PipeShared<4, 3, 128>;
using PipeSharedAlias33256 = PipeShared<3, 3, 256>;
using PipeSharedAlias32128 = PipeShared<3, 2, 128>;
using PipeSharedAlias22256 = PipeShared<2, 2, 256>;
static const PipeVariant kPipeVariants[] = {      // in order of preference
    SD_PIPE_VARIANT(4, 3, 128), SD_PIPE_VARIANT(3, 3, 256), SD_PIPE_VARIANT(3, 2, 128), SD_PIPE_VARIANT(2, 2, 256)};
constexpr int kNumPipeVariants = 4;

// does variant v fit: NB slice buffers + its scratch within 227 KB, the in-kernel fallback (norm_row) in front of the
// failed-item mask, and k + slack candidates per cluster rank
static bool pipe_fits(const PipeVariant& v, size_t slice_bytes, int C, int kcap) {
  if (kcap > v.cap / C) return false;
  if (static_cast<size_t>(v.nb) * slice_bytes + v.fixed > device_max_smem_optin()) return false;
  return slice_bytes + v.row_need <= static_cast<size_t>(v.nb) * slice_bytes + v.mask_off;
}

template <typename T, int NG, int NB, int CAP, bool FV, bool PROF>
static cudaError_t pipe_launch_or_query_fv(const NormParams& p, int rows, cudaStream_t st, int* query_max_clusters) {
  auto kern = norm_topk_pipe_kernel<T, NG, NB, CAP, FV, PROF>;
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
  cfg.blockDim = dim3(32 + NG * kPipeGroupThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(NB) * p.slice_smem_bytes + sizeof(PipeShared<NG, NB, CAP>);
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (pdl_enabled() && query_max_clusters == nullptr) ? 2 : 1;
  if (query_max_clusters != nullptr) {
    // a persistent grid must be co-resident: clusters cannot use every SM (GPC boundaries), ask the driver
    cfg.gridDim = dim3(static_cast<unsigned>(p.cluster * device_sm_count()));
    return cudaOccupancyMaxActiveClusters(query_max_clusters, kern, &cfg);
  }
  cfg.gridDim = dim3(static_cast<unsigned>(p.pipe_clusters) * p.cluster);
  return cudaLaunchKernelEx(&cfg, kern, p);
}

template <typename T, int NG, int NB, int CAP>
static cudaError_t pipe_launch_or_query(const NormParams& p, int rows, cudaStream_t st, int* query_max_clusters) {
  if (p.fv_rows > 0) return pipe_launch_or_query_fv<T, NG, NB, CAP, true, false>(p, rows, st, query_max_clusters);
  if (p.prof != nullptr) return pipe_launch_or_query_fv<T, NG, NB, CAP, false, true>(p, rows, st, query_max_clusters);
  return pipe_launch_or_query_fv<T, NG, NB, CAP, false, false>(p, rows, st, query_max_clusters);
}

template <typename T>
static cudaError_t pipe_dispatch(const NormParams& p, int rows, cudaStream_t st, int* q) {
  switch (p.pipe_groups * 100 + p.pipe_buffers * 10 + (p.pipe_cap == 128 ? 1 : 2)) {
    case 431: return pipe_launch_or_query<T, 4, 3, 128>(p, rows, st, q);
    case 332: return pipe_launch_or_query<T, 3, 3, 256>(p, rows, st, q);
    case 321: return pipe_launch_or_query<T, 3, 2, 128>(p, rows, st, q);
    case 222: return pipe_launch_or_query<T, 2, 2, 256>(p, rows, st, q);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace sd
