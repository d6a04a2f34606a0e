// Per-row implementation of kernel 1 (fast top-k path, dense path, general path) shared by the
// one-cluster-per-row kernel (norm.cu) and the in-kernel fallback of the persistent pipeline (norm_pipe.cu).
#pragma once

#include "rowops.cuh"
#include "specdec_internal.h"

#include <type_traits>

namespace sd {

constexpr int kCapTotal = 512;     // merged candidates per row
constexpr int kFastK = 128;        // largest top_k served by the fast path
constexpr int kTopPCandidates = 64;   // top-p-only rows: candidates tried before falling back to the general path
constexpr int kMaxChunks = 4;      // TMA chunks per slice (pass 1 starts when the first lands)
constexpr uint32_t kTieUlps = 8;   // pivot slack so that logits that tie AFTER the division by T are kept

template <int THREADS>
struct alignas(16) NormShared {
  RowScratch<THREADS> rs;
  uint64_t bar[kMaxChunks];
  int cand_cnt, hot_cnt, n_keep_k, n_keep_p;
  float tau;                                      // pivot: k-th largest thread maximum of this CTA
  int recv_cnt[kMaxCluster];                      // candidates pushed by cluster rank r (-1: that slice needs the general path)
  unsigned short hot[THREADS];                    // threads whose maximum reaches the pivot
  // receive buffer: rank r of the cluster pushes its candidates (logit / T, vocabulary index) to slots
  // [r * cap, r * cap + recv_cnt[r]) of EVERY peer (distributed shared memory stores); after the merge it holds the
  // row's candidate list sorted by (value desc, index asc)
  float r_val[kCapTotal]; int r_idx[kCapTotal];
  union alignas(8) {
    struct { float a_val[kCapTotal]; int a_idx[kCapTotal]; };   // merged list as 64-bit sort keys, later (a_val) final probabilities
    float tm[THREADS];                                          // per-warp sorted thread maxima (pivot phase only)
  };
};

template <typename T, int THREADS, class F>
__device__ __forceinline__ void for_each_elem(const T* slice, int n_vec, long long slice_start, int tid, F f) {
  constexpr int PV = Elem<T>::kPerVec;
  const uint4* s4 = reinterpret_cast<const uint4*>(slice);
  for (int v = tid; v < n_vec; v += THREADS) {
    float o[PV];
    Elem<T>::unpack(s4[v], o);
    const int g = static_cast<int>(slice_start) + v * PV;
#pragma unroll
    for (int j = 0; j < PV; ++j) f(o[j], g + j);
  }
}

__device__ __forceinline__ float max_nan(float a, float b) {   // NaN-propagating maximum
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

// logit / T, correctly rounded in all but pathological cases, in 3 FMA-class instructions (T != 1)
__device__ __forceinline__ float div_fast(float l, float temp, float r_temp) {
  const float q0 = l * r_temp;
  const float x = fmaf(fmaf(-q0, temp, l), r_temp, q0);
  return (fabsf(q0) == INFINITY) ? q0 : x;
}

#ifdef SD_DEBUG_HANG
#define SD_PROF(slot) do { if (p.prof != nullptr && tid == 0) { *reinterpret_cast<volatile long long*>(p.prof + static_cast<long long>(blockIdx.x) * 16 + (slot)) = clock64(); __threadfence_system(); } } while (0)
#else
#define SD_PROF(slot) do { if (p.prof != nullptr && tid == 0) p.prof[static_cast<long long>(blockIdx.x) * 16 + (slot)] = clock64(); } while (0)
#endif

template <typename T, int THREADS>
__device__ __forceinline__ void norm_row(const NormParams& p, const int row) {
  constexpr int PV = Elem<T>::kPerVec;
  constexpr int W = THREADS / 32;
  constexpr float kLog2e = 1.4426950408889634f;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* slice = reinterpret_cast<T*>(smem_raw);
  NormShared<THREADS>& sh = *reinterpret_cast<NormShared<THREADS>*>(smem_raw + p.slice_smem_bytes);
  RowCtx<THREADS> cx(&sh.rs, p.cluster);
  const int tid = cx.tid, lane = cx.lane, warp = cx.warp, C = cx.C;
  const int V = static_cast<int>(p.V);
  const long long start = static_cast<long long>(cx.crank) * p.slice_elems;
  const int n = max(0, min(p.slice_elems, V - static_cast<int>(start)));
  const int n_vec = (n + PV - 1) / PV;
  const T* grow = reinterpret_cast<const T*>(p.logits) + static_cast<long long>(row) * p.ld_in + start;
  const float temp = p.temperature;
  const int k_eff = p.top_k > 0 ? min(p.top_k, V) : 0;
  const bool want_probs = p.probs != nullptr;
  float* orow = want_probs ? p.probs + static_cast<long long>(row) * p.ld_out : nullptr;
  // top-p without top-k: try the candidate machinery on the kTopPCandidates largest logits — the nucleus of a peaked
  // distribution is far smaller than that; if the cumulative mass of the candidates never crosses top_p the row
  // goes to the general path (exactness is never at stake: the crossing must be FOUND among complete candidates)
  const bool topp_only = k_eff == 0 && p.top_p > 0.f && !p.force_general;
  const int k_sel = topp_only ? min(kTopPCandidates, V) : k_eff;
  const bool fast = k_sel > 0 && k_sel <= kFastK && !p.force_general;
  if (fast && C > 1) cx.cluster.barrier_arrive();   // matched by the wait right before candidates are pushed to peers
  SD_PROF(0);

  // ------------------------------------------------------------------ stage the slice
  int n_chunks = 1, chunk_vecs = n_vec;
  if (p.use_tma) {
    const uint32_t bytes = static_cast<uint32_t>(n) * sizeof(T);
    if (bytes >= 16384u) { n_chunks = kMaxChunks; chunk_vecs = ((n_vec + n_chunks - 1) / n_chunks + THREADS - 1) / THREADS * THREADS; }
    if (tid == 0) {
      for (int c = 0; c < n_chunks; ++c) mbar_init(&sh.bar[c], 1);
      fence_barrier_init();
      for (int c = 0; c < n_chunks; ++c) {
        const int v0 = c * chunk_vecs, v1 = min(n_vec, v0 + chunk_vecs);
        if (v1 > v0) {
          const uint32_t cb = static_cast<uint32_t>(v1 - v0) * 16u;
          mbar_expect_tx(&sh.bar[c], cb);
          tma_load_1d(reinterpret_cast<unsigned char*>(slice) + static_cast<size_t>(v0) * 16,
                      reinterpret_cast<const unsigned char*>(grow) + static_cast<size_t>(v0) * 16, cb, &sh.bar[c]);
        }
      }
    }
  } else {
    for (int i = tid; i < n_vec * PV; i += THREADS) slice[i] = i < n ? grow[i] : Elem<T>::neg_inf();
  }
  if (fast && want_probs) {
    // The output row is zero except for <= top_k (+ties) entries: write the zeros NOW, while the logits are still in
    // flight (the threads would only spin on the mbarrier otherwise).  The few non-zeros are scattered at the end,
    // ordered after these stores by the CTA barriers in between.
    float* o = orow + start;
    if (p.vec_out) {
      const int nv4 = n >> 2;
      for (int v = tid; v < nv4; v += THREADS) st_cs_v4(o + 4 * v, 0.f, 0.f, 0.f, 0.f);
      for (int i = (nv4 << 2) + tid; i < n; i += THREADS) o[i] = 0.f;
    } else {
      for (int i = tid; i < n; i += THREADS) o[i] = 0.f;
    }
  }
  __syncthreads();     // mbarrier init (TMA) / staged slice (plain loads) visible to every thread
  SD_PROF(1);

  // ------------------------------------------------------------------ pass 1: thread maxima (NaN-propagating)
  float tmax = -INFINITY;
  {
    const uint4* s4 = reinterpret_cast<const uint4*>(slice);
    for (int c = 0; c < n_chunks; ++c) {
      const int v0 = c * chunk_vecs, v1 = min(n_vec, v0 + chunk_vecs);
      if (v1 <= v0) break;
      if (p.use_tma) mbar_wait(&sh.bar[c], 0);
      if (c == 0) SD_PROF(2);
      for (int v = v0 + tid; v < v1; v += THREADS) tmax = vec_max_nan<T>(tmax, s4[v]);
    }
  }
  if (tmax != tmax || tmax == INFINITY) { atomicOr(p.err_flag, kErrNanLogit); tmax = INFINITY; }
  if (p.use_tma) {
    // Every thread is past its waits: retire the mbarriers.  The next row of this CTA initialises them again, and an
    // mbarrier.init on a still-valid object is undefined (observed on B200: now and then the second row's TMA
    // completion never showed up on the re-initialised barrier and the CTA hung).
    __syncthreads();
    if (tid == 0) for (int c = 0; c < n_chunks; ++c) mbar_inval(&sh.bar[c]);
  }
  SD_PROF(3);
  bool done = false;

  // ================================================================== fast top-k path
  if (fast) {
    // pivot = k-th largest of the THREADS thread maxima: at least k elements of the slice are >= it (one per
    // thread) and, for data without pathological layout, only a handful more.  Each warp sorts its 32 maxima
    // (bitonic, shuffles); lane j < k then ranks its value against the other warps' sorted lists by binary search.
    const float sv = warp_sort_desc(tmax, lane);
    sh.tm[warp * 32 + lane] = sv;
    if (tid == 0) { sh.cand_cnt = 0; sh.hot_cnt = 0; sh.tau = -INFINITY; }
    __syncthreads();
    SD_PROF(11);
    if (lane < min(k_sel, 32)) {
      // W-1 independent binary searches, interleaved so that their shared-memory latencies overlap
      int lo[W], hi[W];
#pragma unroll
      for (int w = 0; w < W; ++w) { lo[w] = 0; hi[w] = 32; }
#pragma unroll
      for (int it = 0; it < 6; ++it) {                      // 33 possible counts (0..32)
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const int mid = (lo[w] + hi[w]) >> 1;
          const float y = sh.tm[w * 32 + min(mid, 31)];
          const bool before = mid < 32 && ((y > sv) || (y == sv && w < warp));
          lo[w] = before ? mid + 1 : lo[w];
          hi[w] = before ? hi[w] : mid;
        }
      }
      int rank = lane;
#pragma unroll
      for (int w = 0; w < W; ++w) rank += (w == warp) ? 0 : lo[w];
      if (rank == k_sel - 1) sh.tau = sv;
    }
    __syncthreads();
    SD_PROF(12);
    const float tau = float_down(sh.tau, temp == 1.0f ? 0u : kTieUlps);

    // pass 2: only threads whose maximum reaches the pivot own candidates; their vectors are re-scanned by all threads
    const int cap = kCapTotal / C;                            // receive slots per cluster rank
    float* my_val = sh.r_val + cx.crank * cap;
    int* my_idx = sh.r_idx + cx.crank * cap;
    if (tmax >= tau && n_vec > tid) { const int h = atomicAdd(&sh.hot_cnt, 1); sh.hot[h] = static_cast<unsigned short>(tid); }
    __syncthreads();
    SD_PROF(13);
    {
      const int H = sh.hot_cnt;
      const int vpt = (n_vec + THREADS - 1) / THREADS;
      const uint4* s4 = reinterpret_cast<const uint4*>(slice);
      for (int item = tid; item < H * vpt; item += THREADS) {
        const int v = sh.hot[item / vpt] + (item % vpt) * THREADS;
        if (v < n_vec) {
          float o[PV];
          Elem<T>::unpack(s4[v], o);
          const int g = static_cast<int>(start) + v * PV;
#pragma unroll
          for (int j = 0; j < PV; ++j) {
            if (o[j] >= tau && g + j < V) {
              const int pos = atomicAdd(&sh.cand_cnt, 1);
              if (pos < cap) { my_val[pos] = __fdiv_rn(o[j], temp); my_idx[pos] = g + j; }
            }
          }
        }
      }
    }
    __syncthreads();
    SD_PROF(4);
    {
      // self-check: a slice that collected every element >= its pivot and at least min(k, n) of them cannot miss a
      // member of the row's top-k, whatever the pivot was; anything else sends the row to the general path
      const int c = sh.cand_cnt;
      const int mine = (c > cap || c < min(k_sel, n)) ? -1 : c;
      if (C > 1) {
        cx.cluster.barrier_wait();                            // (arrive at kernel start) every peer is running
        for (int r = 0; r < C; ++r) {
          if (r == cx.crank) continue;
          NormShared<THREADS>* ps = cx.cluster.map_shared_rank(&sh, r);
          for (int i = tid; i < max(mine, 0); i += THREADS) {
            ps->r_val[cx.crank * cap + i] = my_val[i];
            ps->r_idx[cx.crank * cap + i] = my_idx[i];
          }
          if (tid == 0) ps->recv_cnt[cx.crank] = mine;
        }
      }
      if (tid == 0) sh.recv_cnt[cx.crank] = mine;
    }
    if (C > 1) cx.cluster.sync(); else __syncthreads();       // pushes visible; no remote access after this point
    SD_PROF(5);

    // merge: concatenate the C receive regions as 64-bit sort keys (value key << 32 | ~index) — identical in every
    // CTA of the cluster; the key array overlays a_val/a_idx, which are only needed again after the sort
    unsigned long long* a_key = reinterpret_cast<unsigned long long*>(sh.a_val);
    int n_tot = 0;
    bool ok = true;
    for (int r = 0; r < C; ++r) {
      const int c = sh.recv_cnt[r];
      ok &= c >= 0;
      for (int i = tid; i < c; i += THREADS) {
        const float xv = sh.r_val[r * cap + i] + 0.0f;                       // -0 -> +0: equal values tie on the index
        a_key[n_tot + i] = (static_cast<unsigned long long>(f2key(xv)) << 32) | (0xffffffffu - static_cast<uint32_t>(sh.r_idx[r * cap + i]));
      }
      n_tot += max(c, 0);
    }
    ok &= n_tot >= k_sel;
    if (p.prof != nullptr && tid == 0) p.prof[static_cast<long long>(blockIdx.x) * 16 + 14] = n_tot;
    __syncthreads();

    if (ok) {
      // rank sort (value descending, vocabulary index ascending): four threads per candidate, branch-free inner loop
      for (int base = 0; base < n_tot; base += THREADS / 4) {
        const int i = base + (tid >> 2);
        const bool live = i < n_tot;
        const unsigned long long ki = live ? a_key[i] : 0ull;
        int r = 0;
        if (live) {
#pragma unroll 4
          for (int j = tid & 3; j < n_tot; j += 4) r += a_key[j] > ki ? 1 : 0;
        }
        r += __shfl_xor_sync(0xffffffffu, r, 1);
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        if (live && (tid & 3) == 0) {
          sh.r_val[r] = key2f(static_cast<uint32_t>(ki >> 32));
          sh.r_idx[r] = static_cast<int>(0xffffffffu - static_cast<uint32_t>(ki & 0xffffffffull));
        }
      }
      __syncthreads();
      SD_PROF(6);
      const float kth = sh.r_val[k_sel - 1];
      for (int i = tid; i < n_tot; i += THREADS)
        if (sh.r_val[i] >= kth && (i + 1 == n_tot || sh.r_val[i + 1] < kth)) sh.n_keep_k = i + 1;
      __syncthreads();
      const int nk = sh.n_keep_k;
      double zfull = 0.0;              // top-p only: softmax denominator of the WHOLE row (reference utils.py:172)
      if (topp_only && nk < V) {       // cluster-uniform
        const float Mx = sh.r_val[0];
        const float r_temp = 1.0f / temp;
        float acc = 0.f;
        for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int) {
          const float x = temp == 1.0f ? l : div_fast(l, temp, r_temp);
          acc += exp2f((x - Mx) * 1.4426950408889634f); });            // padding is -inf -> 0
        zfull = cx.allreduce_sum(static_cast<double>(acc));
      }

      if (warp == 0 && nk <= 32) {   // usual case: the whole kept list lives in one warp's registers
        const bool in_k = lane < nk;
        const float x = in_k ? sh.r_val[lane] : -INFINITY;
        const int id = in_k ? sh.r_idx[lane] : 0x7fffffff;
        const float M = __shfl_sync(0xffffffffu, x, 0);
        const float e = in_k ? expf(x - M) : 0.f;
        const double zs = warp_sum(static_cast<double>(e));
        int np = nk;
        if (p.top_p > 0.f) {
          // reference: cum = cumsum(softmax(sorted)) (fp64 accumulate); drop entry r>0 iff cum[r-1] > top_p (utils.py:171-176)
          const float sp = e * (1.0f / static_cast<float>(zs));
          const double cum = warp_scan_incl(static_cast<double>(sp), lane);
          const unsigned ball = __ballot_sync(0xffffffffu, in_k && static_cast<float>(cum) > p.top_p);
          if (ball) np = min(nk, __ffs(ball));                           // crossing entry itself is kept
        }
        const bool in_p = lane < np;
        const double z2 = warp_sum(in_p ? static_cast<double>(e) : 0.0);
        const float logz = logf(static_cast<float>(z2));                 // reference: exp(log_softmax), utils.py:199
        const float pr = in_p ? expf((x - M) - logz) : 0.f;
        if (in_p && (!(pr >= 0.f) || isinf(pr))) atomicOr(p.err_flag, kErrNanLogit);
        if (in_p) sh.a_val[lane] = pr;
        if (lane == 0) sh.n_keep_p = np;
        if (p.cmp.cnt != nullptr && cx.crank == 0) {              // compact form of the row for the sparse verify path
          const long long cr = static_cast<long long>(row) * p.cmp.row_stride;
          if (np <= p.cmp.cap) {
            if (in_p) { p.cmp.idx[cr * p.cmp.cap + lane] = id; p.cmp.val[cr * p.cmp.cap + lane] = pr; }
            if (lane == 0) p.cmp.cnt[cr] = np;
          } else if (lane == 0) p.cmp.cnt[cr] = -1;
        }
        if (p.u != nullptr && cx.crank == 0 && p.u[row] >= 0.f) {   // inverse-CDF sample in vocabulary order over the kept list
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
              p.tok_out[row] = (pr < kProbGuard) ? top_id : id;          // utils.py:228-230 guard
          }
        }
      } else if (warp == 0) {   // long kept list (large top_k, many ties, top-p-only candidates): strided over the list
        const float M = sh.r_val[0];
        double zs = 0.0;
        for (int i = lane; i < nk; i += 32) zs += static_cast<double>(expf(sh.r_val[i] - M));
        zs = warp_sum(zs);
        if (topp_only && nk < V) zs = zfull;
        int np = nk;
        bool crossed = false;
        if (p.top_p > 0.f) {
          const float rz = 1.0f / static_cast<float>(zs);
          double run = 0.0;
          for (int base = 0; base < nk; base += 32) {
            const int i = base + lane;
            const float sp = i < nk ? expf(sh.r_val[i] - M) * rz : 0.f;
            const double cum = warp_scan_incl(static_cast<double>(sp), lane) + run;
            const bool over = i < nk && static_cast<float>(cum) > p.top_p;
            const unsigned ball = __ballot_sync(0xffffffffu, over);
            if (ball) { np = min(nk, base + __ffs(ball)); crossed = true; break; }
            run = __shfl_sync(0xffffffffu, cum, 31);
          }
        }
        if (topp_only && nk < V && !crossed) np = -1;      // nucleus larger than the candidate list: general path
        double z2 = 0.0;
        for (int i = lane; i < np; i += 32) z2 += static_cast<double>(expf(sh.r_val[i] - M));
        z2 = warp_sum(z2);
        const float logz = np > 0 ? logf(static_cast<float>(z2)) : 0.f;
        bool badp = false;
        for (int i = lane; i < np; i += 32) {
          const float pr = expf((sh.r_val[i] - M) - logz);
          badp |= !(pr >= 0.f) || isinf(pr);
          sh.a_val[i] = pr;
        }
        if (badp) atomicOr(p.err_flag, kErrNanLogit);
        if (lane == 0) sh.n_keep_p = np;
        __syncwarp();
        if (np >= 0 && p.cmp.cnt != nullptr && cx.crank == 0) {
          const long long cr = static_cast<long long>(row) * p.cmp.row_stride;
          if (np <= p.cmp.cap) {
            for (int i = lane; i < np; i += 32) { p.cmp.idx[cr * p.cmp.cap + i] = sh.r_idx[i]; p.cmp.val[cr * p.cmp.cap + i] = sh.a_val[i]; }
            if (lane == 0) p.cmp.cnt[cr] = np;
          } else if (lane == 0) p.cmp.cnt[cr] = -1;
        }
        if (np >= 0 && p.u != nullptr && cx.crank == 0 && p.u[row] >= 0.f) {
          const int e = frexp_exp(sh.a_val[0]);
          unsigned long long tot = 0ull;
          for (int i = lane; i < np; i += 32) tot += weight_of(sh.a_val[i], e);
          tot = warp_sum(tot);
          if (tot == 0ull) {
            if (lane == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.tok_out[row] = 0; }
          } else {
            const unsigned long long target = scale_target(tot, u_to_int(p.u[row]));
            for (int i = lane; i < np; i += 32) {
              const int id = sh.r_idx[i];
              const unsigned long long wi = weight_of(sh.a_val[i], e);
              unsigned long long before = 0ull;
              for (int j = 0; j < np; ++j) before += (sh.r_idx[j] < id) ? weight_of(sh.a_val[j], e) : 0ull;
              if (wi > 0ull && target >= before && target < before + wi)
                p.tok_out[row] = (sh.a_val[i] < kProbGuard) ? sh.r_idx[0] : id;
            }
          }
        }
      }
      __syncthreads();
      if (sh.n_keep_p >= 0) {            // (< 0: top-p-only nucleus larger than the candidate list -> general path below)
        if (want_probs) {
          const int np = sh.n_keep_p;
          for (int i = tid; i < np; i += THREADS) {
            const int id = sh.r_idx[i];
            if (id >= start && id < start + n) orow[id] = sh.a_val[i];
          }
        }
        done = true;
      }
    }
    SD_PROF(7);
    if (!done) __syncthreads();
  }

  if (!done && p.cmp.cnt != nullptr && cx.crank == 0 && tid == 0)
    p.cmp.cnt[static_cast<long long>(row) * p.cmp.row_stride] = -1;        // dense / general path: no compact list

  // ================================================================== dense path (no filter): max / sum / exp
  if (!done && k_eff == 0 && !(p.top_p > 0.f) && !p.force_general) {
    const float r_temp = 1.0f / temp;
    auto xof = [&](float l) { return temp == 1.0f ? l : div_fast(l, temp, r_temp); };
    // one cluster exchange only: every CTA reduces (max, sum of exp relative to ITS max) locally, peers combine
    float lmax = warp_max(tmax);
    if (lane == 0) sh.rs.wf[warp] = lmax;
    __syncthreads();
    lmax = sh.rs.wf[0];
#pragma unroll
    for (int w = 1; w < W; ++w) lmax = fmaxf(lmax, sh.rs.wf[w]);
    const float Mc = temp == 1.0f ? lmax : __fdiv_rn(lmax, temp);        // -inf for an empty / fully masked slice
    const bool do_sample = p.u != nullptr && p.u[row] >= 0.f;              // row-uniform
    float acc = 0.f;
    int amin = 0x7fffffff;                                                // first index holding this CTA's maximum
    if (Mc > -INFINITY) {
      if (do_sample)
        for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int g) {
          acc += exp2f((xof(l) - Mc) * kLog2e);
          if (l == lmax && g < V) amin = min(amin, g); });
      else
        for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int) { acc += exp2f((xof(l) - Mc) * kLog2e); });
    }
    double lsum = warp_sum(static_cast<double>(acc));                     // padding is -inf -> contributes 0
    if (do_sample) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amin = min(amin, __shfl_xor_sync(0xffffffffu, amin, o));
    }
    if (lane == 0) { sh.rs.wd[warp] = lsum; sh.rs.wi[warp] = amin; }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0;
      int am = 0x7fffffff;
      for (int w = 0; w < W; ++w) { a += sh.rs.wd[w]; am = min(am, sh.rs.wi[w]); }
      sh.rs.xf[cx.parity][0] = lmax;
      sh.rs.xd[cx.parity] = a;
      sh.rs.xi[cx.parity][0] = am;
    }
    cx.xchg_sync();
    float Ml = -INFINITY;
    for (int r = 0; r < C; ++r) Ml = fmaxf(Ml, cx.peer(r)->xf[cx.parity][0]);
    const float M = temp == 1.0f ? Ml : __fdiv_rn(Ml, temp);
    double z = 0.0;
    int argmax = 0x7fffffff;                                              // first index holding the ROW maximum
    for (int r = 0; r < C; ++r) {
      const float mr_l = cx.peer(r)->xf[cx.parity][0];
      const float mr = temp == 1.0f ? mr_l : __fdiv_rn(mr_l, temp);
      if (mr > -INFINITY) z += cx.peer(r)->xd[cx.parity] * static_cast<double>(exp2f((mr - M) * kLog2e));
      if (mr_l == Ml) argmax = min(argmax, cx.peer(r)->xi[cx.parity][0]);
    }
    cx.parity ^= 1;
    bool pending_wait = false;
    if (C > 1 && !do_sample) { cx.cluster.barrier_arrive(); pending_wait = true; }   // done with peers' memory
    const float logz = logf(static_cast<float>(z));
    if (!(z > 0.0) || isinf(logz) || logz != logz) atomicOr(p.err_flag, kErrNanLogit);
    const float c2 = -logz * kLog2e;
    auto vec_probs = [&](int v, float (&pr)[PV]) {
      float o[PV];
      Elem<T>::unpack(reinterpret_cast<const uint4*>(slice)[v], o);
#pragma unroll
      for (int j = 0; j < PV; ++j) pr[j] = exp2f(fmaf(xof(o[j]) - M, kLog2e, c2));   // exp((x - M) - logZ), utils.py:199
    };
    SD_PROF(8);
    if (!do_sample) {
      if (want_probs) {
        float* o = orow + start;
        const int nfull = p.vec_out ? n / PV : 0;
        for (int v = tid; v < nfull; v += THREADS) {
          float pr[PV];
          vec_probs(v, pr);
          st_cs_vec<PV>(o + v * PV, pr);
        }
        for (int v = nfull + tid; v < n_vec; v += THREADS) {
          float pr[PV];
          vec_probs(v, pr);
          for (int j = 0; j < PV; ++j) if (v * PV + j < n) o[v * PV + j] = pr[j];
        }
      }
    } else {
      // write pass and the sampler's weight sums in one sweep, in the sampler's segment order (one contiguous segment of
      // vectors per warp, lanes strided: still 512 contiguous bytes per warp store)
      const float pmax = exp2f(c2);
      const int e = frexp_exp(pmax);
      const int vpw = (n_vec + W - 1) / W;
      const int v_begin = warp * vpw, v_end = min(n_vec, v_begin + vpw);
      const int nfull = (want_probs && p.vec_out) ? n / PV : 0;
      float* o = want_probs ? orow + start : nullptr;
      unsigned long long lane_w = 0ull;
      for (int v = v_begin + lane; v < v_end; v += 32) {
        float pr[PV];
        vec_probs(v, pr);
#pragma unroll
        for (int j = 0; j < PV; ++j) lane_w += weight_of(pr[j], e);
        if (v < nfull) {
          st_cs_vec<PV>(o + v * PV, pr);
        } else if (want_probs) {
          for (int j = 0; j < PV; ++j) if (v * PV + j < n) o[v * PV + j] = pr[j];
        }
      }
      unsigned long long total = 0ull;
      float psel = 1.f;
      const int tok = cluster_icdf<PV, THREADS>(cx, n_vec, start, pmax, p.u[row], vec_probs, &total, &psel, &lane_w);
      if (total == 0ull) {
        if (tid == 0 && cx.crank == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.tok_out[row] = 0; }
      } else if (tok >= 0) {
        p.tok_out[row] = psel < kProbGuard ? argmax : tok;
      }
    }
    SD_PROF(9);
    if (pending_wait) cx.cluster.barrier_wait();
    else if (C > 1) cx.cluster.sync();
    done = true;
  }

  // ================================================================== general path (sort-free threshold search)
  if (!done) {
    auto xof = [&](float l) { return temp == 1.0f ? l : __fdiv_rn(l, temp); };
    const float Ml = cx.allreduce_max(tmax);
    const float M = xof(Ml);

    uint32_t Kk = 0u;                         // keep keys >= Kk (top-k, ties kept: utils.py:169)
    if (k_eff > 0 && k_eff < V) {
      Kk = search16<true, THREADS, int>(cx, 0u, 32,
          [&](auto f) { for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int g) {
              f(g < V ? 1 : 0, f2key(xof(l))); }); },
          [&](int cnt) { return cnt >= k_eff; });
    }
    uint32_t Kp = Kk;                         // top-p: keep key > Kp, or key == Kp and index <= tie_last
    int tie_last = 0x7fffffff;
    if (p.top_p > 0.f) {
      double zl = 0.0;
      {
        float acc = 0.f;
        for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int g) {
          const float x = xof(l);
          acc += (g < V && f2key(x) >= Kk) ? expf(x - M) : 0.f; });
        zl = cx.allreduce_sum(static_cast<double>(acc));
      }
      const float rz = 1.0f / static_cast<float>(zl);
      auto mass_each = [&](auto f) { for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int g) {
          const float x = xof(l);
          const uint32_t key = f2key(x);
          f((g < V && key >= Kk) ? expf(x - M) * rz : 0.f, key); }); };
      // largest key K with mass(key >= K) > top_p : the value at which the sorted cumsum crosses top_p
      // (if even the whole kept mass is <= top_p nothing is cut: every pivot fails, the search returns 0 and the
      //  total-mass test below leaves Kp = Kk)
      const uint32_t Kc = search16<true, THREADS, float>(cx, 0u, 32, mass_each, [&](float m) { return m > p.top_p; });
      float acc_gt = 0.f, acc_all = 0.f; int ties = 0;
      for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int g) {
        const float x = xof(l);
        const uint32_t key = f2key(x);
        const float m = (g < V && key >= Kk) ? expf(x - M) * rz : 0.f;
        acc_all += m;
        acc_gt += key > Kc ? m : 0.f;
        ties += (g < V && key == Kc) ? 1 : 0; });
      const double g_all = cx.allreduce_sum(static_cast<double>(acc_all));
      const double g_gt = cx.allreduce_sum(static_cast<double>(acc_gt));
      const int n_ties = cx.allreduce_sum(ties);
      if (static_cast<float>(g_all) > p.top_p) {
        Kp = Kc;
        // number of tied entries (ascending index) needed for the running sum to exceed top_p
        const double spv = static_cast<double>(expf(key2f(Kc) - M) * rz);
        long long c = 1;
        if (spv > 0.0) {
          c = static_cast<long long>(floor((static_cast<double>(p.top_p) - g_gt) / spv)) - 1;
          if (c < 1) c = 1;
          while (c < n_ties && !(static_cast<float>(g_gt + static_cast<double>(c) * spv) > p.top_p)) ++c;
        }
        if (c < n_ties) {
          const int need = static_cast<int>(c);
          // largest I with #(ties with index < I) < need  ==  index of the need-th tie
          tie_last = static_cast<int>(search16<false, THREADS, int>(cx, 0u, 24,
              [&](auto f) { for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int g) {
                  f((g < V && f2key(xof(l)) == Kc) ? 1 : 0, static_cast<uint32_t>(g)); }); },
              [&](int cnt) { return cnt < need; }));
        }
      }
    }
    auto kept = [&](uint32_t key, int g) {
      return g < V && key >= Kk && (key > Kp || (key == Kp && g <= tie_last));
    };
    float acc = 0.f;
    for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int g) {
      const float x = xof(l);
      acc += kept(f2key(x), g) ? expf(x - M) : 0.f; });
    const double z2 = cx.allreduce_sum(static_cast<double>(acc));
    const float logz = logf(static_cast<float>(z2));
    if (!(z2 > 0.0) || isinf(logz) || logz != logz) atomicOr(p.err_flag, kErrNanLogit);

    auto vec_probs = [&](int v, float (&pr)[PV]) {
      float o[PV];
      Elem<T>::unpack(reinterpret_cast<const uint4*>(slice)[v], o);
      const int g = static_cast<int>(start) + v * PV;
#pragma unroll
      for (int j = 0; j < PV; ++j) {
        const float x = xof(o[j]);
        pr[j] = kept(f2key(x), g + j) ? expf((x - M) - logz) : 0.f;
      }
    };
    if (want_probs) {
      float* o = orow + start;
      const int nfull = p.vec_out ? n / PV : 0;
      for (int v = tid; v < nfull; v += THREADS) {
        float pr[PV];
        vec_probs(v, pr);
        st_cs_vec<PV>(o + v * PV, pr);
      }
      for (int v = nfull + tid; v < n_vec; v += THREADS) {
        float pr[PV];
        vec_probs(v, pr);
        for (int j = 0; j < PV; ++j) if (v * PV + j < n) o[v * PV + j] = pr[j];
      }
    }
    if (p.u != nullptr && p.u[row] >= 0.f) {
      // argmax index for the < 1e-9 guard (first index holding the row maximum)
      int amin = 0x7fffffff;
      for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int g) { if (l == Ml && g < V) amin = min(amin, g); });
      const int argmax = cx.allreduce_min(amin);
      unsigned long long total = 0ull;
      float psel = 1.f;
      const int tok = cluster_icdf<PV, THREADS>(cx, n_vec, start, expf(-logz), p.u[row], vec_probs, &total, &psel);
      if (total == 0ull) {
        if (tid == 0 && cx.crank == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.tok_out[row] = 0; }
      } else if (tok >= 0) {
        p.tok_out[row] = psel < kProbGuard ? argmax : tok;
      }
    }
    if (C > 1) cx.cluster.sync();   // keep this CTA's shared memory alive until every peer finished reading it
  }
  SD_PROF(10);
}


}  // namespace sd
