// Row-level building blocks shared by the norm (filter+softmax) and verify kernels:
// block / cluster reductions over distributed shared memory, the 16-way threshold search used by
// the general (sort-free) top-k / top-p path, and the vocabulary-order inverse-CDF sampler on
// 64-bit fixed-point weights.
#pragma once

#include "common.cuh"

#include <type_traits>

namespace sd {

constexpr int kMaxPortableCluster = 8;
constexpr int kMaxCluster = 16;    // 16 needs cudaFuncAttributeNonPortableClusterSizeAllowed (pipelined kernel only)

// Scratch every CTA of a cluster keeps at the same shared-memory offset, so that peers can read it
// through distributed shared memory (cluster.map_shared_rank).
template <int THREADS>
struct alignas(16) RowScratch {
  static constexpr int W = THREADS / 32;
  float wf[W];                    // per-warp partials
  double wd[W];
  unsigned long long wu[W];
  int wi[W];
  float w15f[W][16];              // per-warp partials of the 15 pivots of one search round
  int w15i[W][16];
  float r15f[16];                 // cluster-wide totals of one search round
  int r15i[16];
  // cluster exchange slots, double buffered (see RowCtx::xchg)
  float xf[2][16];
  int xi[2][16];
  double xd[2];
  unsigned long long xu[2];
  float bf;                       // block-level broadcast values
  double bd;
  unsigned long long bu;
  int bi;
};

template <int THREADS>
struct RowCtx {
  static constexpr int W = THREADS / 32;
  RowScratch<THREADS>* s;
  cg::cluster_group cluster;
  int C, crank, tid, lane, warp;
  int parity;                      // exchange-slot parity, advanced identically by every CTA

  __device__ RowCtx(RowScratch<THREADS>* scratch, int cluster_size)
      : s(scratch), cluster(cg::this_cluster()), C(cluster_size) {
    tid = threadIdx.x;
    lane = tid & 31;
    warp = tid >> 5;
    crank = C > 1 ? static_cast<int>(cluster.block_rank()) : 0;
    parity = 0;
  }
  __device__ __forceinline__ const RowScratch<THREADS>* peer(int r) const {
    return C > 1 ? cluster.map_shared_rank(s, r) : s;
  }
  // One cluster barrier per exchange is enough because the slots are double buffered: a CTA can
  // only overwrite slot[parity] two exchanges later, i.e. after every peer passed the barrier of
  // the exchange in between and therefore finished reading.
  __device__ __forceinline__ void xchg_sync() {
    if (C > 1) cluster.sync(); else __syncthreads();
  }

  // ---- all-reduce (block, then cluster in rank order => deterministic) ---------------------
  __device__ float allreduce_max(float v) {
    v = warp_max(v);
    if (lane == 0) s->wf[warp] = v;
    __syncthreads();
    if (tid == 0) {
      float m = s->wf[0];
      for (int w = 1; w < W; ++w) m = fmaxf(m, s->wf[w]);
      s->xf[parity][0] = m;
    }
    xchg_sync();
    float m = peer(0)->xf[parity][0];
    for (int r = 1; r < C; ++r) m = fmaxf(m, peer(r)->xf[parity][0]);
    parity ^= 1;
    return m;
  }
  __device__ double allreduce_sum(double v) {
    v = warp_sum(v);
    if (lane == 0) s->wd[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double a = 0.0;
      for (int w = 0; w < W; ++w) a += s->wd[w];
      s->xd[parity] = a;
    }
    xchg_sync();
    double a = 0.0;
    for (int r = 0; r < C; ++r) a += peer(r)->xd[parity];
    parity ^= 1;
    return a;
  }
  __device__ int allreduce_sum(int v) {
    v = warp_sum(v);
    if (lane == 0) s->wi[warp] = v;
    __syncthreads();
    if (tid == 0) {
      int a = 0;
      for (int w = 0; w < W; ++w) a += s->wi[w];
      s->xi[parity][0] = a;
    }
    xchg_sync();
    int a = 0;
    for (int r = 0; r < C; ++r) a += peer(r)->xi[parity][0];
    parity ^= 1;
    return a;
  }
  __device__ int allreduce_min(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) s->wi[warp] = v;
    __syncthreads();
    if (tid == 0) {
      int a = s->wi[0];
      for (int w = 1; w < W; ++w) a = min(a, s->wi[w]);
      s->xi[parity][0] = a;
    }
    xchg_sync();
    int a = peer(0)->xi[parity][0];
    for (int r = 1; r < C; ++r) a = min(a, peer(r)->xi[parity][0]);
    parity ^= 1;
    return a;
  }
  __device__ unsigned long long allreduce_max(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
      v = t > v ? t : v;
    }
    if (lane == 0) s->wu[warp] = v;
    __syncthreads();
    if (tid == 0) {
      unsigned long long a = s->wu[0];
      for (int w = 1; w < W; ++w) a = s->wu[w] > a ? s->wu[w] : a;
      s->xu[parity] = a;
    }
    xchg_sync();
    unsigned long long a = peer(0)->xu[parity];
    for (int r = 1; r < C; ++r) { unsigned long long t = peer(r)->xu[parity]; a = t > a ? t : a; }
    parity ^= 1;
    return a;
  }
  // 15 pivots at once.  After the call every thread can read totals from s->r15f / s->r15i.
  __device__ void allreduce15(float (&acc)[15]) {
#pragma unroll
    for (int j = 0; j < 15; ++j) {
      float v = warp_sum(acc[j]);
      if (lane == 0) s->w15f[warp][j] = v;
    }
    __syncthreads();
    if (tid < 15) {
      float a = 0.f;
      for (int w = 0; w < W; ++w) a += s->w15f[w][tid];
      s->xf[parity][tid] = a;
    }
    xchg_sync();
    if (tid < 15) {
      float a = 0.f;
      for (int r = 0; r < C; ++r) a += peer(r)->xf[parity][tid];
      s->r15f[tid] = a;
    }
    parity ^= 1;
    __syncthreads();
  }
  __device__ void allreduce15(int (&acc)[15]) {
#pragma unroll
    for (int j = 0; j < 15; ++j) {
      int v = warp_sum(acc[j]);
      if (lane == 0) s->w15i[warp][j] = v;
    }
    __syncthreads();
    if (tid < 15) {
      int a = 0;
      for (int w = 0; w < W; ++w) a += s->w15i[w][tid];
      s->xi[parity][tid] = a;
    }
    xchg_sync();
    if (tid < 15) {
      int a = 0;
      for (int r = 0; r < C; ++r) a += peer(r)->xi[parity][tid];
      s->r15i[tid] = a;
    }
    parity ^= 1;
    __syncthreads();
  }
};

// ------------------------------------------------------------------------------------------------
// 16-way search:  largest P in [lo, lo + 2^top_bit) (top_bit multiple of 4) such that pred(stat(P))
// holds, where stat(P) = sum over elements of weight(e) * [pos(e) >= P]   (GE = true)
//                     or sum over elements of weight(e) * [pos(e) <  P]   (GE = false)
// and pred(stat(.)) is monotone true -> false in P with pred(stat(lo)) true.
// `each(f)` must call f(weight, pos) for every element this thread owns (weight 0 = excluded).
template <bool GE, int THREADS, typename Acc, class Each, class Pred>
__device__ uint32_t search16(RowCtx<THREADS>& cx, uint32_t lo, int top_bit, Each each, Pred pred) {
  for (int shift = top_bit - 4; shift >= 0; shift -= 4) {
    Acc acc[15];
#pragma unroll
    for (int j = 0; j < 15; ++j) acc[j] = Acc(0);
    each([&](Acc w, uint32_t pos) {
      int jlo, jhi;
      if (GE) {        // contributes to pivots lo + (j << shift) <= pos
        jlo = 1;
        jhi = pos >= lo ? static_cast<int>(min((pos - lo) >> shift, 15u)) : 0;
      } else {         // contributes to pivots lo + (j << shift) > pos
        jhi = 15;
        jlo = pos >= lo ? static_cast<int>(min(((pos - lo) >> shift) + 1u, 16u)) : 1;
      }
#pragma unroll
      for (int j = 1; j <= 15; ++j) acc[j - 1] += (j >= jlo && j <= jhi) ? w : Acc(0);
    });
    cx.allreduce15(acc);
    int best = 0;
    if constexpr (std::is_same<Acc, float>::value) {
#pragma unroll
      for (int j = 1; j <= 15; ++j)
        if (pred(cx.s->r15f[j - 1])) best = j;
    } else {
#pragma unroll
      for (int j = 1; j <= 15; ++j)
        if (pred(cx.s->r15i[j - 1])) best = j;
    }
    lo += static_cast<uint32_t>(best) << shift;
    __syncthreads();   // r15 is rewritten by the next round
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------
// Inverse-CDF sample in vocabulary order over a row that is distributed over the cluster.
//   vec_probs(v, out[PV]) yields the PV (>= 0) weights of 16-byte vector v of this CTA's slice,
//   already zeroed for out-of-range / filtered entries.  pmax = row maximum (identical in all CTAs).
// Returns the selected global index in the CTA+warp+lane that owns it through *hit (others: -1);
// `total_out` receives the cluster-wide weight total (0 => nothing to sample from).
template <int PV, int THREADS, class VecProbs>
__device__ int cluster_icdf(RowCtx<THREADS>& cx, int n_vec, long long slice_start, float pmax, float u,
                            VecProbs vec_probs, unsigned long long* total_out, float* prob_out,
                            const unsigned long long* lane_weights = nullptr) {
  constexpr int W = THREADS / 32;
  const int e = frexp_exp(pmax);
  const int vpw = (n_vec + W - 1) / W;                       // vectors per warp segment
  const int v_begin = cx.warp * vpw, v_end = min(n_vec, v_begin + vpw);
  unsigned long long mine = 0ull;
  if (lane_weights != nullptr) {
    mine = *lane_weights;          // caller already summed weight_of(p, e) over this lane's vectors of this warp's segment
  } else {
    for (int v = v_begin + cx.lane; v < v_end; v += 32) {
      float pr[PV];
      vec_probs(v, pr);
#pragma unroll
      for (int j = 0; j < PV; ++j) mine += weight_of(pr[j], e);
    }
  }
  mine = warp_sum(mine);
  if (cx.lane == 0) cx.s->wu[cx.warp] = mine;
  __syncthreads();
  if (cx.tid == 0) {
    unsigned long long a = 0ull;
    for (int w = 0; w < W; ++w) a += cx.s->wu[w];
    cx.s->xu[cx.parity] = a;
  }
  cx.xchg_sync();
  unsigned long long base = 0ull, total = 0ull;
  for (int r = 0; r < cx.C; ++r) {
    unsigned long long tr = cx.peer(r)->xu[cx.parity];
    if (r < cx.crank) base += tr;
    total += tr;
  }
  cx.parity ^= 1;
  *total_out = total;
  if (total == 0ull) return -1;
  const unsigned long long target = scale_target(total, u_to_int(u));
  for (int w = 0; w < cx.warp; ++w) base += cx.s->wu[w];
  const unsigned long long seg = cx.s->wu[cx.warp];
  if (!(target >= base && target < base + seg)) return -1;   // warp-uniform
  // the owning warp walks its segment, 32 vectors per step, in vocabulary order
  unsigned long long run = base;
  for (int v0 = v_begin; v0 < v_end; v0 += 32) {
    const int v = v0 + cx.lane;
    float pr[PV];
    unsigned long long wv[PV], vs = 0ull;
#pragma unroll
    for (int j = 0; j < PV; ++j) { pr[j] = 0.f; wv[j] = 0ull; }
    if (v < v_end) {
      vec_probs(v, pr);
#pragma unroll
      for (int j = 0; j < PV; ++j) { wv[j] = weight_of(pr[j], e); vs += wv[j]; }
    }
    const unsigned long long incl = warp_scan_incl(vs, cx.lane) + run;
    const bool crossed = incl > target;
    const unsigned ball = __ballot_sync(0xffffffffu, crossed);
    if (ball) {
      const int owner = __ffs(ball) - 1;
      int found = -1;
      if (cx.lane == owner) {
        unsigned long long c = incl - vs;
#pragma unroll
        for (int j = 0; j < PV; ++j) {
          c += wv[j];
          if (found < 0 && c > target) { found = static_cast<int>(slice_start) + v * PV + j; *prob_out = pr[j]; }
        }
      }
      return found;                                           // >= 0 only in the owning lane
    }
    run = __shfl_sync(0xffffffffu, incl, 31);
  }
  return -1;
}

}  // namespace sd
