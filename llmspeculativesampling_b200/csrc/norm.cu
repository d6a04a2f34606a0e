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
#include "rowops.cuh"
#include "specdec_internal.h"

#include <type_traits>

namespace sd {

constexpr int kCapLocal = 192;     // candidates one CTA may collect / publish
constexpr int kCapTotal = 512;     // merged candidates per row
constexpr int kFastK = 128;        // largest top_k served by the fast path
constexpr int kMaxChunks = 4;      // TMA chunks per slice (pass 1 starts when the first lands)
constexpr uint32_t kTieUlps = 8;   // pivot slack so that logits that tie AFTER the division by T are kept

template <int THREADS>
struct alignas(16) NormShared {
  RowScratch<THREADS> rs;
  uint64_t bar[kMaxChunks];
  int cand_cnt, n_pub, overflow, n_keep_k, n_keep_p, token;
  float l_val[kCapLocal]; int l_idx[kCapLocal];   // local candidates sorted (value desc, index asc); read by peers
  float a_val[kCapTotal]; int a_idx[kCapTotal];   // local candidates (unsorted), later the row's sorted list
  float m_val[kCapTotal]; int m_idx[kCapTotal];   // merged list, later (m_val) the final probabilities
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

// rank sort of n (value, index) pairs: value descending, index ascending
template <int THREADS>
__device__ __forceinline__ void rank_sort(const float* __restrict__ iv, const int* __restrict__ ii, int n,
                                          float* __restrict__ ov, int* __restrict__ oi, int tid) {
  for (int i = tid; i < n; i += THREADS) {
    const float x = iv[i];
    const int id = ii[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float y = iv[j];
      rank += (y > x || (y == x && ii[j] < id)) ? 1 : 0;
    }
    ov[rank] = x;
    oi[rank] = id;
  }
}

template <typename T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) norm_probs_kernel(const NormParams p) {
  constexpr int PV = Elem<T>::kPerVec;
  constexpr int W = THREADS / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* slice = reinterpret_cast<T*>(smem_raw);
  NormShared<THREADS>& sh = *reinterpret_cast<NormShared<THREADS>*>(smem_raw + p.slice_smem_bytes);
  RowCtx<THREADS> cx(&sh.rs, p.cluster);
  const int tid = cx.tid, lane = cx.lane, warp = cx.warp, C = cx.C;
  const int row = blockIdx.x / C;
  const int V = static_cast<int>(p.V);
  const long long start = static_cast<long long>(cx.crank) * p.slice_elems;
  const int n = max(0, min(p.slice_elems, V - static_cast<int>(start)));
  const int n_vec = (n + PV - 1) / PV;
  const T* grow = reinterpret_cast<const T*>(p.logits) + static_cast<long long>(row) * p.ld_in + start;
  const float temp = p.temperature;

  // ------------------------------------------------------------------ stage the slice
  int n_chunks = 1, chunk_vecs = n_vec;
  if (p.use_tma) {
    const uint32_t bytes = static_cast<uint32_t>(n) * sizeof(T);
    if (bytes >= 16384u) { n_chunks = kMaxChunks; chunk_vecs = ((n_vec + n_chunks - 1) / n_chunks + 63) & ~63; }
    if (tid == 0) {
      for (int c = 0; c < n_chunks; ++c) mbar_init(&sh.bar[c], 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
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
    __syncthreads();
  }

  // ------------------------------------------------------------------ pass 1: thread maxima
  float tmax = -INFINITY;
  bool bad = false;
  {
    const uint4* s4 = reinterpret_cast<const uint4*>(slice);
    for (int c = 0; c < n_chunks; ++c) {
      const int v0 = c * chunk_vecs, v1 = min(n_vec, v0 + chunk_vecs);
      if (v1 <= v0) break;
      if (p.use_tma) mbar_wait(&sh.bar[c], 0);
      for (int v = v0 + tid; v < v1; v += THREADS) {
        float o[PV];
        Elem<T>::unpack(s4[v], o);
#pragma unroll
        for (int j = 0; j < PV; ++j) { tmax = fmaxf(tmax, o[j]); bad |= (o[j] != o[j]); }
      }
    }
  }
  if (bad || tmax == INFINITY) atomicOr(p.err_flag, kErrNanLogit);

  const int k_eff = p.top_k > 0 ? min(p.top_k, V) : 0;
  const bool want_probs = p.probs != nullptr;
  float* orow = want_probs ? p.probs + static_cast<long long>(row) * p.ld_out : nullptr;
  bool done = false;

  // ================================================================== fast top-k path
  if (k_eff > 0 && k_eff <= kFastK && !p.force_general) {
    // pivot: every warp has >= a thread maxima >= tau_min, so the slice has >= a*W >= k elements >= tau_min;
    // alternatively one warp alone has k thread maxima >= its k-th  (k <= 32).  Take the tighter bound.
    const float sv = warp_sort_desc(tmax, lane);
    const int a = (k_eff + W - 1) / W;
    const float va = __shfl_sync(0xffffffffu, sv, a - 1);
    const float vk = __shfl_sync(0xffffffffu, sv, min(k_eff, 32) - 1);
    if (lane == 0) { sh.rs.wf[warp] = va; sh.rs.w15f[warp][0] = (k_eff <= 32) ? vk : -INFINITY; }
    if (tid == 0) { sh.cand_cnt = 0; sh.n_pub = 0; sh.overflow = 0; }
    __syncthreads();
    float tau_min = INFINITY, tau_alt = -INFINITY;
#pragma unroll
    for (int w = 0; w < W; ++w) { tau_min = fminf(tau_min, sh.rs.wf[w]); tau_alt = fmaxf(tau_alt, sh.rs.w15f[w][0]); }
    const float tau = float_down(fmaxf(tau_min, tau_alt), temp == 1.0f ? 0u : kTieUlps);

    // pass 2: collect candidates (logit / T computed only for them)
    for_each_elem<T, THREADS>(slice, n_vec, start, tid, [&](float l, int g) {
      if (l >= tau && g < V) {
        const int pos = atomicAdd(&sh.cand_cnt, 1);
        if (pos < kCapLocal) { sh.a_val[pos] = __fdiv_rn(l, temp); sh.a_idx[pos] = g; }
      }
    });
    __syncthreads();
    const int nl = min(sh.cand_cnt, kCapLocal);
    const bool local_over = sh.cand_cnt > kCapLocal;
    rank_sort<THREADS>(sh.a_val, sh.a_idx, nl, sh.l_val, sh.l_idx, tid);
    __syncthreads();
    if (nl > 0) {     // publish the local top-k plus everything tied with the local k-th
      const float kth_l = sh.l_val[min(k_eff, nl) - 1];
      for (int i = tid; i < nl; i += THREADS)
        if (sh.l_val[i] >= kth_l && (i + 1 == nl || sh.l_val[i + 1] < kth_l)) sh.n_pub = i + 1;
    }
    if (tid == 0 && local_over) sh.overflow = 1;
    if (C > 1) cx.cluster.sync(); else __syncthreads();

    // merge the published lists of all CTAs of the cluster (every CTA builds the same merged list)
    int n_tot = 0, any_over = 0, my_off = 0;
    int offs[kMaxCluster];
    for (int r = 0; r < C; ++r) {
      const NormShared<THREADS>* ps = C > 1 ? cx.cluster.map_shared_rank(&sh, r) : &sh;
      offs[r] = n_tot;
      n_tot += ps->n_pub;
      any_over |= ps->overflow;
    }
    (void)my_off;
    const bool ok = !any_over && n_tot <= kCapTotal && n_tot >= k_eff;
    if (ok) {
      for (int r = 0; r < C; ++r) {
        const NormShared<THREADS>* ps = C > 1 ? cx.cluster.map_shared_rank(&sh, r) : &sh;
        const int cnt = (r + 1 < C ? offs[r + 1] : n_tot) - offs[r];
        for (int i = tid; i < cnt; i += THREADS) { sh.m_val[offs[r] + i] = ps->l_val[i]; sh.m_idx[offs[r] + i] = ps->l_idx[i]; }
      }
    }
    if (C > 1) cx.cluster.sync(); else __syncthreads();   // peers are done reading my l_val / l_idx

    if (ok) {
      rank_sort<THREADS>(sh.m_val, sh.m_idx, n_tot, sh.a_val, sh.a_idx, tid);
      __syncthreads();
      const float kth = sh.a_val[k_eff - 1];
      for (int i = tid; i < n_tot; i += THREADS)
        if (sh.a_val[i] >= kth && (i + 1 == n_tot || sh.a_val[i + 1] < kth)) sh.n_keep_k = i + 1;
      __syncthreads();
      const int nk = sh.n_keep_k;

      if (warp == 0) {   // top-p cut and final softmax on the (short) sorted list
        const float M = sh.a_val[0];
        double zs = 0.0;
        for (int i = lane; i < nk; i += 32) zs += static_cast<double>(expf(sh.a_val[i] - M));
        zs = warp_sum(zs);
        int np = nk;
        if (p.top_p > 0.f) {
          // reference: cum = cumsum(softmax(sorted)); drop entry r>0 iff cum[r-1] > top_p  (utils.py:171-176)
          const float rz = 1.0f / static_cast<float>(zs);
          double run = 0.0;
          for (int base = 0; base < nk; base += 32) {
            const int i = base + lane;
            const float sp = i < nk ? expf(sh.a_val[i] - M) * rz : 0.f;
            const double cum = warp_scan_incl(static_cast<double>(sp), lane) + run;
            const bool over = i < nk && static_cast<float>(cum) > p.top_p;
            const unsigned ball = __ballot_sync(0xffffffffu, over);
            if (ball) { np = min(nk, base + __ffs(ball)); break; }     // crossing entry itself is kept
            run = __shfl_sync(0xffffffffu, cum, 31);
          }
        }
        double z2 = 0.0;
        for (int i = lane; i < np; i += 32) z2 += static_cast<double>(expf(sh.a_val[i] - M));
        z2 = warp_sum(z2);
        const float logz = logf(static_cast<float>(z2));                 // reference: exp(log_softmax), utils.py:199
        bool badp = false;
        for (int i = lane; i < np; i += 32) {
          const float pr = expf((sh.a_val[i] - M) - logz);
          badp |= !(pr >= 0.f) || isinf(pr);
          sh.m_val[i] = pr;
        }
        if (badp) atomicOr(p.err_flag, kErrNanLogit);
        if (lane == 0) sh.n_keep_p = np;
        __syncwarp();

        if (p.u != nullptr && cx.crank == 0) {   // inverse-CDF sample in vocabulary order over the kept list
          const int e = frexp_exp(sh.m_val[0]);
          unsigned long long tot = 0ull;
          for (int i = lane; i < np; i += 32) tot += weight_of(sh.m_val[i], e);
          tot = warp_sum(tot);
          if (tot == 0ull) {
            if (lane == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.tok_out[row] = 0; }
          } else {
            const unsigned long long target = scale_target(tot, u_to_int(p.u[row]));
            for (int i = lane; i < np; i += 32) {
              const int id = sh.a_idx[i];
              const unsigned long long wi = weight_of(sh.m_val[i], e);
              unsigned long long before = 0ull;
              for (int j = 0; j < np; ++j) before += (sh.a_idx[j] < id) ? weight_of(sh.m_val[j], e) : 0ull;
              if (wi > 0ull && target >= before && target < before + wi)
                p.tok_out[row] = (sh.m_val[i] < kProbGuard) ? sh.a_idx[0] : id;   // utils.py:228-230 guard
            }
          }
        }
      }
      __syncthreads();
      if (want_probs) {
        const int np = sh.n_keep_p;
        float* o = orow + start;
        if (p.vec_out) {
          const int nv4 = n >> 2;
          for (int v = tid; v < nv4; v += THREADS) st_cs_v4(o + 4 * v, 0.f, 0.f, 0.f, 0.f);
          for (int i = (nv4 << 2) + tid; i < n; i += THREADS) o[i] = 0.f;
        } else {
          for (int i = tid; i < n; i += THREADS) o[i] = 0.f;
        }
        __syncthreads();
        for (int i = tid; i < np; i += THREADS) {
          const int id = sh.a_idx[i];
          if (id >= start && id < start + n) orow[id] = sh.m_val[i];
        }
      }
      done = true;
    }
  }

  // ================================================================== dense / general path
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
      if (p.vec_out && PV == 4) {
        for (int v = tid; v < (n >> 2); v += THREADS) {
          float pr[PV];
          vec_probs(v, pr);
          st_cs_v4(o + 4 * v, pr[0], pr[1], pr[2], pr[3]);
        }
        for (int v = (n >> 2) + tid; v < n_vec; v += THREADS) {       // ragged last vector
          float pr[PV];
          vec_probs(v, pr);
          for (int j = 0; j < PV; ++j) if (v * PV + j < n) o[v * PV + j] = pr[j];
        }
      } else if (p.vec_out && PV == 8) {
        for (int v = tid; v < (n >> 3); v += THREADS) {
          float pr[PV];
          vec_probs(v, pr);
          st_cs_v4(o + 8 * v, pr[0], pr[1], pr[2], pr[3]);
          st_cs_v4(o + 8 * v + 4, pr[4 % PV], pr[5 % PV], pr[6 % PV], pr[7 % PV]);
        }
        for (int v = (n >> 3) + tid; v < n_vec; v += THREADS) {
          float pr[PV];
          vec_probs(v, pr);
          for (int j = 0; j < PV; ++j) if (v * PV + j < n) o[v * PV + j] = pr[j];
        }
      } else {
        for (int v = tid; v < n_vec; v += THREADS) {
          float pr[PV];
          vec_probs(v, pr);
          for (int j = 0; j < PV; ++j) if (v * PV + j < n) o[v * PV + j] = pr[j];
        }
      }
    }
    if (p.u != nullptr) {
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
}

// ------------------------------------------------------------------------------------------------
// host side
static int g_tune_cluster = 0, g_tune_threads = 0;

void set_norm_tuning(int cluster, int threads) { g_tune_cluster = cluster; g_tune_threads = threads; }

template <typename T, int THREADS, int MINB>
static cudaError_t launch_cfg(const NormParams& p, size_t smem, int rows, cudaStream_t st) {
  auto kern = norm_probs_kernel<T, THREADS, MINB>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(rows) * p.cluster);
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
  while (C < kMaxCluster && (row_bytes + C - 1) / C > 64 * 1024) C <<= 1;
  while (C < kMaxCluster && static_cast<long long>(rows) * C < 148 && row_bytes / (2 * C) >= 8192) C <<= 1;
  if (g_tune_cluster > 0) C = g_tune_cluster;
  // slice: multiple of 128 elements so that every slice start is 16-byte aligned in both dtypes
  long long slice = ((p.V + C - 1) / C + 127) & ~127LL;
  while (C > 1 && slice * (C - 1) >= p.V) { C >>= 1; slice = ((p.V + C - 1) / C + 127) & ~127LL; }
  p.cluster = C;
  p.slice_elems = static_cast<int>(slice);
  const size_t slice_bytes = (static_cast<size_t>(slice) * es + 127) & ~static_cast<size_t>(127);
  p.slice_smem_bytes = static_cast<int>(slice_bytes);
  const bool aligned_in = (reinterpret_cast<uintptr_t>(p.logits) % 16 == 0) && ((p.ld_in * es) % 16 == 0) && (row_bytes % 16 == 0);
  p.use_tma = aligned_in ? 1 : 0;
  p.vec_out = (p.probs != nullptr && reinterpret_cast<uintptr_t>(p.probs) % 16 == 0 && p.ld_out % 4 == 0) ? 1 : 0;
  int threads = g_tune_threads > 0 ? g_tune_threads : (slice_bytes > 96 * 1024 ? 512 : 256);
  size_t smem;
  switch (threads) {
    case 512:
      smem = slice_bytes + sizeof(NormShared<512>);
      if (smem > 227 * 1024) return cudaErrorInvalidValue;
      return launch_cfg<T, 512, 1>(p, smem, rows, st);
    case 1024:
      smem = slice_bytes + sizeof(NormShared<1024>);
      if (smem > 227 * 1024) return cudaErrorInvalidValue;
      return launch_cfg<T, 1024, 1>(p, smem, rows, st);
    default:
      smem = slice_bytes + sizeof(NormShared<256>);
      if (smem > 227 * 1024) return cudaErrorInvalidValue;
      return launch_cfg<T, 256, 3>(p, smem, rows, st);
  }
}

cudaError_t launch_norm(const NormParams& p, int dtype, int rows, cudaStream_t st) {
  switch (dtype) {
    case kF32: return launch_typed<float>(p, rows, st);
    case kBF16: return launch_typed<__nv_bfloat16>(p, rows, st);
    case kF16: return launch_typed<__half>(p, rows, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace sd
