// Kernel 2: fused verify — accept test, first rejection (warp ballot), residual max(0, p - q) with
// its normaliser, inverse-CDF draw from a pre-drawn uniform, token append and length update.
// One thread-block cluster per request; the two probability rows that matter (p_n and q_n) are
// staged once through 1-D TMA bulk copies and never re-read from HBM.
//
// Replaces /root/reference/sampling/speculative_sampling.py:1966-2027 (accept loop with ~7 host
// syncs per drafted token, max_fn + sample, rollback bookkeeping, torch.cat append) and, with
// strict = 1, the accept rule of speculative_sampling_v2 (:2152-2181).  With q == nullptr the
// kernel is the drop-in for sampling/utils.py:213-233 (sample) on rows of p.
#include <algorithm>

#include "rowops.cuh"
#include "specdec_internal.h"
#include "verify_sparse.cuh"
#include "verify_row.cuh"

namespace sd {

template <int THREADS>
struct alignas(16) VerifyShared {
  RowScratch<THREADS> rs;
  uint64_t bar;
  int n_acc;
};

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) verify_kernel(const VerifyParams p) {
  constexpr int PV = 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int slice_bytes = p.slice_elems * 4;
  float* ps = reinterpret_cast<float*>(smem_raw);   // p_n slice, turned in place into max(0, p_n - q_n)
  VerifyShared<THREADS>& sh = *reinterpret_cast<VerifyShared<THREADS>*>(smem_raw + slice_bytes);
  RowCtx<THREADS> cx(&sh.rs, p.cluster);
  const int tid = cx.tid, lane = cx.lane, warp = cx.warp, C = cx.C;
  const int b = blockIdx.x / C;
  if (p.active != nullptr && p.active[b] == 0) return;       // whole cluster leaves together
  const int V = static_cast<int>(p.V);
  const int gamma = p.gamma;
  const bool has_q = p.q != nullptr;

  if (tid == 0 && p.use_tma) { mbar_init(&sh.bar, 1); fence_barrier_init(); }

  // ---- accept scan: lane i tests drafted token i (speculative_sampling.py:1975-1990)
  if (warp == 0) {
    int n_acc = 0;
    if (has_q) {
      bool acc = true, tie = false;
      float ratio = 0.f;
      if (lane < gamma) {
        long long tok = p.draft[b * p.draft_stride + lane];
        if (tok < 0 || tok >= V) { atomicOr(p.err_flag, kErrBadToken); tok = 0; }
        const float pv = p.p[b * p.p_req_stride + lane * p.p_row_stride + tok];
        const float qv = p.q[b * p.q_req_stride + lane * p.q_row_stride + tok];
        if (qv == 0.f) atomicOr(p.err_flag, kErrZeroQ);
        ratio = __fdiv_rn(pv, qv);
        const float u = p.u_acc[b * p.u_acc_stride + lane];
        const float thr = p.strict ? fminf(1.0f, ratio) : ratio;
        acc = p.strict ? (u < thr) : !(u > thr);
        tie = (u == thr);
        if (cx.crank == 0 && p.ratios != nullptr) p.ratios[b * gamma + lane] = ratio;
      }
      const unsigned rej = __ballot_sync(0xffffffffu, !acc);
      n_acc = rej ? (__ffs(rej) - 1) : gamma;
      if (cx.crank == 0 && p.tie_count != nullptr && tie && lane < gamma && lane <= n_acc) atomicAdd(p.tie_count, 1);
    }
    if (lane == 0) sh.n_acc = n_acc;
  }
  __syncthreads();
  const int n_acc = sh.n_acc;
  bool use_q = has_q && n_acc < gamma;

  // ---- stage this CTA's slice of p_n (and q_n)
  const long long start = static_cast<long long>(cx.crank) * p.slice_elems;
  const int n = max(0, min(p.slice_elems, V - static_cast<int>(start)));
  const int n_vec = (n + PV - 1) / PV;
  const float* prow = p.p + b * p.p_req_stride + (has_q ? n_acc : 0) * p.p_row_stride + start;
  const float* qrow = use_q ? p.q + b * p.q_req_stride + n_acc * p.q_row_stride + start : nullptr;
  auto stage_p = [&](bool tma) {
    if (tma) {
      if (tid == 0 && n > 0) {
        const uint32_t bytes = static_cast<uint32_t>(n) * 4u;
        mbar_expect_tx(&sh.bar, bytes);
        tma_load_1d(ps, prow, bytes, &sh.bar);
      }
      if (n > 0) mbar_wait(&sh.bar, 0);
    } else {
      for (int i = tid; i < n_vec * PV; i += THREADS) ps[i] = i < n ? prow[i] : 0.f;
      __syncthreads();
    }
  };
  stage_p(p.use_tma != 0);
  // residual max(0, p - q) (utils.py:240) built in place: q_n is read from HBM exactly once, straight into registers
  auto fold_q = [&]() {
    if (p.use_tma) {
      for (int v = tid; v < n_vec; v += THREADS) {
        float4 a = reinterpret_cast<float4*>(ps)[v];
        const uint4 cu = ld_nc_v4(reinterpret_cast<const uint4*>(qrow) + v);
        a.x = fmaxf(a.x - __uint_as_float(cu.x), 0.f); a.y = fmaxf(a.y - __uint_as_float(cu.y), 0.f);
        a.z = fmaxf(a.z - __uint_as_float(cu.z), 0.f); a.w = fmaxf(a.w - __uint_as_float(cu.w), 0.f);
        reinterpret_cast<float4*>(ps)[v] = a;
      }
    } else {
      for (int i = tid; i < n; i += THREADS) ps[i] = fmaxf(ps[i] - qrow[i], 0.f);
    }
    __syncthreads();
  };
  if (use_q) fold_q();

  auto vecw = [&](int v, float (&w)[PV]) {
    const float4 a = reinterpret_cast<const float4*>(ps)[v];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
  };
  unsigned long long best = 0ull;
  float rmax = 0.f;
  int argmax = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    unsigned long long mine = 0ull;
    bool bad = false;
    for (int v = tid; v < n_vec; v += THREADS) {
      float w[PV];
      vecw(v, w);
#pragma unroll
      for (int j = 0; j < PV; ++j) {
        const int g = static_cast<int>(start) + v * PV + j;
        bad |= !(w[j] >= 0.f) || isinf(w[j]);
        if (w[j] > 0.f) {
          const unsigned long long pk = (static_cast<unsigned long long>(f2key(w[j])) << 32) | (0xffffffffu - static_cast<uint32_t>(g));
          mine = pk > mine ? pk : mine;
        }
      }
    }
    if (bad) atomicOr(p.err_flag, kErrEmptyRow);               // negative / NaN / inf weights: 'prob error'
    best = cx.allreduce_max(mine);
    if (best != 0ull || !use_q || p.strict) break;
    use_q = false;                                             // empty residual: resample from p_n (:2009-2010)
    __syncthreads();
    stage_p(false);
  }
  if (best == 0ull) {
    if (tid == 0 && cx.crank == 0) {
      atomicOr(p.err_flag, kErrEmptyRow);
      p.next_tok[b] = 0;
      if (p.n_accepted != nullptr) p.n_accepted[b] = n_acc;
    }
  } else {
    rmax = key2f(static_cast<uint32_t>(best >> 32));
    argmax = static_cast<int>(0xffffffffu - static_cast<uint32_t>(best & 0xffffffffu));
    unsigned long long total = 0ull;
    float psel = 1.f;
    const int tok = cluster_icdf<PV, THREADS>(cx, n_vec, start, rmax, p.u_final[b], vecw, &total, &psel);
    if (tok >= 0) {
      float guard_val = psel;
      if (use_q) {                                             // sample(max_fn(p - q)): guard sees the normalised value
        const float s = ldexpf(__ull2float_rn(total), frexp_exp(rmax) - kScaleBits);
        guard_val = __fdiv_rn(psel, s + 1e-6f);
      }
      const long long out = guard_val < kProbGuard ? argmax : tok;
      p.next_tok[b] = out;
      if (p.n_accepted != nullptr) p.n_accepted[b] = n_acc;
      if (p.stats != nullptr) { atomicAdd(&p.stats[0], static_cast<unsigned long long>(n_acc)); atomicAdd(&p.stats[1], 1ull); }
      if (p.tokens != nullptr) {                               // append + "rollback": one counter write
        const int L = p.seq_len[b];
        p.tokens[b * p.tokens_stride + L + n_acc] = out;
        p.seq_len[b] = L + n_acc + 1;
      }
    }
  }
  if (C > 1) cx.cluster.sync();
}

// ------------------------------------------------------------------------------------------------
// Dense verify, one CTA per request, row staged in shared memory (verify_row.cuh) — used whenever a row of V fp32
// probabilities fits one CTA (V <= ~57k); larger vocabularies use the cluster kernel above.
__global__ void __launch_bounds__(kRowThreads, 1) verify_row_kernel(const VerifyParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int V = static_cast<int>(p.V), gamma = p.gamma;
  float* row = reinterpret_cast<float*>(smem_raw);
  const size_t row_al = (static_cast<size_t>(V) * 4 + 127) & ~static_cast<size_t>(127);
  unsigned char* qring = smem_raw + row_al;                      // p.q_slots x 16 KB behind the staged row
  RowSampleShared& sh = *reinterpret_cast<RowSampleShared*>(smem_raw + row_al + static_cast<size_t>(p.q_slots) * kRowQChunkBytes);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();                                                   // (launched with programmatic stream serialization)
  long long* prof = p.prof != nullptr ? p.prof + static_cast<long long>(b) * 8 : nullptr;
  if (prof != nullptr && tid == 0) prof[0] = clock64();
  if (p.active != nullptr && p.active[b] == 0) return;
  const bool has_q = p.q != nullptr;
  const bool vec_rows = (V & 3) == 0;                            // (the sampler's TMA path; misaligned pointers re-check inside)
  if (tid == 32 && vec_rows) row_sample_init_barriers(sh, p.q_slots);   // while warp 0 runs the accept scan
  // ---- accept scan: lane i tests drafted token i (speculative_sampling.py:1975-1990)
  if (warp == 0) {
    int n_acc = 0;
    if (has_q) {
      bool acc = true, tie = false;
      if (lane < gamma) {
        long long tok = p.draft[b * p.draft_stride + lane];
        if (tok < 0 || tok >= V) { atomicOr(p.err_flag, kErrBadToken); tok = 0; }
        const float pv = p.p[b * p.p_req_stride + lane * p.p_row_stride + tok];
        const float qv = p.q[b * p.q_req_stride + lane * p.q_row_stride + tok];
        if (qv == 0.f) atomicOr(p.err_flag, kErrZeroQ);
        const float ratio = __fdiv_rn(pv, qv);
        const float u = p.u_acc[b * p.u_acc_stride + lane];
        const float thr = p.strict ? fminf(1.0f, ratio) : ratio;
        acc = p.strict ? (u < thr) : !(u > thr);
        tie = (u == thr);
        if (p.ratios != nullptr) p.ratios[b * gamma + lane] = ratio;
      }
      const unsigned rej = __ballot_sync(0xffffffffu, !acc);
      n_acc = rej ? (__ffs(rej) - 1) : gamma;
      if (p.tie_count != nullptr && tie && lane < gamma && lane <= n_acc) atomicAdd(p.tie_count, 1);
    }
    if (lane == 0) sh.n_acc = n_acc;
  }
  __syncthreads();
  if (prof != nullptr && tid == 0) prof[1] = clock64();          // accept scan done
  const int n_acc = sh.n_acc;
  const bool use_q = has_q && n_acc < gamma;
  const float* prow = p.p + b * p.p_req_stride + (has_q ? n_acc : 0) * p.p_row_stride;
  const float* qrow = use_q ? p.q + b * p.q_req_stride + n_acc * p.q_row_stride : nullptr;
  const long long tok = row_residual_sample(prow, qrow, V, p.u_final[b], !p.strict, row, sh, p.err_flag, prof, qring, p.q_slots, vec_rows);
  if (prof != nullptr && tok >= 0) prof[5] = clock64();          // token found (the one thread that holds it)
  if (tok == -2) {
    if (tid == 0) { p.next_tok[b] = 0; if (p.n_accepted != nullptr) p.n_accepted[b] = n_acc; }
  } else if (tok >= 0) {
    verify_commit(p, b, n_acc, tok);
  }
}

// ------------------------------------------------------------------------------------------------
// The same with TWO CTAs per request (a cluster of 2) for batches that leave half of the SMs idle: one SM streams about
// 25 bytes per clock through TMA, so the 2 * V * 4 bytes of a request are the bulk of its critical path; each CTA stages
// one half of p_n and streams the matching half of q_n, the maximum and the exact weight totals of the halves are
// exchanged through distributed shared memory (verify_row.cuh), the CTA whose half holds the target commits.
__global__ void __launch_bounds__(kRowThreads, 1) verify_row2_kernel(const VerifyParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int V = static_cast<int>(p.V), gamma = p.gamma;
  const int b = blockIdx.x >> 1, crank = blockIdx.x & 1;          // (1-D cluster of 2: rank = blockIdx.x % 2)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // split in units of one q chunk (1024 vectors): rank 0 takes the first ceil(chunks / 2) chunks
  const int n_vec = V >> 2;
  const int n_qc = (n_vec + 1023) >> 10;
  const int vec0 = min(n_vec, ((n_qc + 1) >> 1) << 10);
  const int v_first = crank ? vec0 : 0, v_cnt = crank ? n_vec - vec0 : vec0;
  float* row = reinterpret_cast<float*>(smem_raw);
  const size_t row_al = (static_cast<size_t>(vec0) * 16 + 127) & ~static_cast<size_t>(127);
  unsigned char* qring = smem_raw + row_al;
  RowSampleShared& sh = *reinterpret_cast<RowSampleShared*>(smem_raw + row_al + static_cast<size_t>(p.q_slots) * kRowQChunkBytes);
  pdl_wait();
  if (p.active != nullptr && p.active[b] == 0) return;           // (both CTAs of the cluster: no barrier is left waiting)
  if (tid == 32) row_sample_init_barriers(sh, p.q_slots);
  if (warp == 0) {                                               // accept scan, redundantly in both CTAs; rank 0 publishes the statistics
    bool acc = true, tie = false;
    if (lane < gamma) {
      long long tok = p.draft[b * p.draft_stride + lane];
      if (tok < 0 || tok >= V) { if (crank == 0) atomicOr(p.err_flag, kErrBadToken); tok = 0; }
      const float pv = p.p[b * p.p_req_stride + lane * p.p_row_stride + tok];
      const float qv = p.q[b * p.q_req_stride + lane * p.q_row_stride + tok];
      if (qv == 0.f && crank == 0) atomicOr(p.err_flag, kErrZeroQ);
      const float ratio = __fdiv_rn(pv, qv);
      const float u = p.u_acc[b * p.u_acc_stride + lane];
      const float thr = p.strict ? fminf(1.0f, ratio) : ratio;
      acc = p.strict ? (u < thr) : !(u > thr);
      tie = (u == thr);
      if (p.ratios != nullptr && crank == 0) p.ratios[b * gamma + lane] = ratio;
    }
    const unsigned rej = __ballot_sync(0xffffffffu, !acc);
    const int n_acc = rej ? (__ffs(rej) - 1) : gamma;
    if (p.tie_count != nullptr && crank == 0 && tie && lane < gamma && lane <= n_acc) atomicAdd(p.tie_count, 1);
    if (lane == 0) sh.n_acc = n_acc;
  }
  __syncthreads();
  const int n_acc = sh.n_acc;
  const bool use_q = n_acc < gamma;
  const float* prow = p.p + b * p.p_req_stride + n_acc * p.p_row_stride + static_cast<long long>(v_first) * 4;
  const float* qrow = use_q ? p.q + b * p.q_req_stride + n_acc * p.q_row_stride + static_cast<long long>(v_first) * 4 : nullptr;
  const long long tok = row_residual_sample(prow, qrow, v_cnt * 4, p.u_final[b], !p.strict, row, sh, p.err_flag, nullptr, qring, p.q_slots,
                                            true, crank, 2, v_first * 4);
  if (tok == -2) {
    if (tid == 0 && crank == 0) { p.next_tok[b] = 0; if (p.n_accepted != nullptr) p.n_accepted[b] = n_acc; }
  } else if (tok >= 0) {
    verify_commit(p, b, n_acc, tok);
  }
}

// Static shared memory of the row kernels (scratch + a few words) comes out of the same per-block budget as the staged row.
constexpr int kRowStaticReserve = 4096;
static bool row_kernel_fits(long long V) {
  return row_sample_smem(V) + kRowStaticReserve <= static_cast<size_t>(device_max_smem_optin());
}

template <class K>
static cudaError_t set_row_smem(K kern, bool* flags) {
  int dev_id = 0;
  (void)cudaGetDevice(&dev_id);
  if (!flags[dev_id & 63]) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    if (fa.sharedSizeBytes > static_cast<size_t>(kRowStaticReserve)) return cudaErrorInvalidValue;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, device_max_smem_optin() - kRowStaticReserve);
    if (e != cudaSuccess) return e;
    flags[dev_id & 63] = true;
  }
  return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// Sparse verify: when kernel 1 also emitted the compact (index, probability) lists of its top-k filtered rows, the
// residual max(0, p_n - q_n) lives on the <= ~2k entries of p_n's support, so one warp per request does the whole
// verify step from a few hundred bytes instead of two dense vocabulary rows.  Arithmetic (fp32 subtraction, 64-bit
// fixed-point weights, integer prefix sums in vocabulary order) is the dense kernel's, so tokens are identical.
// Requests whose lists are unavailable (count -1: the row was served by the dense / general path) fall back, inside
// this kernel, to a single-CTA dense scan straight from HBM/L2.
constexpr int kSparseThreads = 128;

__global__ void __launch_bounds__(kSparseThreads) verify_sparse_kernel(const VerifyParams p) {
  __shared__ RowScratch<kSparseThreads> rs;
  __shared__ int s_n_acc, s_dense;
  constexpr int kPre = 9;                                     // rows per tensor prefetched into shared memory
  __shared__ uint2 p_pre[kPre][kSparseCap], q_pre[kPre][kSparseCap];
  __shared__ int p_cnt_pre[kPre], q_cnt_pre[kPre];
  __shared__ int q_idx[kSparseCap];
  __shared__ float q_val[kSparseCap];
  __shared__ int e_idx[kSparseCap];
  __shared__ unsigned long long e_w[kSparseCap];
  const int b = blockIdx.x;
  pdl_wait();
  if (p.active != nullptr && p.active[b] == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = static_cast<int>(p.V), gamma = p.gamma;
  const int n_pre_p = min(gamma + 1, kPre), n_pre_q = min(gamma, kPre);

  // warp 0: everything the accept scan needs that does not depend on another load is requested up front
  long long tok = 0;
  float u_a = 0.f, u_f = 0.f;
  int L_pre = -1;
  if (warp == 0) {
    if (p.tokens != nullptr) L_pre = p.seq_len[b];
    if (lane < gamma) {
      tok = p.draft[b * p.draft_stride + lane];
      u_a = p.u_acc[b * p.u_acc_stride + lane];
    }
    u_f = p.u_final[b];
  } else {
    // warps 1..3 pull the (tiny) compact lists of every row into shared memory: p[token] / q[token] of the accept
    // scan and the residual row are then looked up there — no dependent chain of global loads
    for (int i = tid - 32; i < (n_pre_p + n_pre_q) * kSparseCap; i += kSparseThreads - 32) {
      const int rr = i / kSparseCap, j = i - rr * kSparseCap;
      if (rr < n_pre_p) {
        const long long cr = b * p.pc_req_stride + rr * p.pc.row_stride;
        if (j < p.pc.cap) p_pre[rr][j] = make_uint2(static_cast<uint32_t>(p.pc.idx[cr * p.pc.cap + j]), __float_as_uint(p.pc.val[cr * p.pc.cap + j]));
        if (j == 0) p_cnt_pre[rr] = p.pc.cnt[cr];
      } else {
        const int r2 = rr - n_pre_p;
        const long long cr = b * p.qc_req_stride + r2 * p.qc.row_stride;
        if (j < p.qc.cap) q_pre[r2][j] = make_uint2(static_cast<uint32_t>(p.qc.idx[cr * p.qc.cap + j]), __float_as_uint(p.qc.val[cr * p.qc.cap + j]));
        if (j == 0) q_cnt_pre[r2] = p.qc.cnt[cr];
      }
    }
  }
  __syncthreads();                                            // lists are in shared memory  
  int n_acc_w0 = 0;
  bool tie_w0 = false;
  if (warp == 0) {
    // ---- accept scan: p[token], q[token] from the compact lists (an index that is not listed has probability 0);
    //      rows without a list (count -1 / too long) gather the single elements from the dense rows
    bool acc = true, tie = false;
    float ratio = 0.f;
    if (lane < gamma) {
      if (tok < 0 || tok >= V) { atomicOr(p.err_flag, kErrBadToken); tok = 0; }
      float pv = 0.f, qv = 0.f;
      const int cpl = lane < kPre ? p_cnt_pre[lane] : -1, cql = lane < kPre ? q_cnt_pre[lane] : -1;
      if (cpl >= 0 && cpl <= kSparseCap && cpl <= p.pc.cap) {
        for (int t = 0; t < cpl; ++t) pv = static_cast<long long>(p_pre[lane][t].x) == tok ? __uint_as_float(p_pre[lane][t].y) : pv;
      } else {
        pv = p.p[b * p.p_req_stride + lane * p.p_row_stride + tok];
      }
      if (cql >= 0 && cql <= kSparseCap && cql <= p.qc.cap) {
        for (int t = 0; t < cql; ++t) qv = static_cast<long long>(q_pre[lane][t].x) == tok ? __uint_as_float(q_pre[lane][t].y) : qv;
      } else {
        qv = p.q[b * p.q_req_stride + lane * p.q_row_stride + tok];
      }
      if (qv == 0.f) atomicOr(p.err_flag, kErrZeroQ);
      ratio = __fdiv_rn(pv, qv);
      const float thr = p.strict ? fminf(1.0f, ratio) : ratio;
      acc = p.strict ? (u_a < thr) : !(u_a > thr);
      tie = (u_a == thr);
      if (p.ratios != nullptr) p.ratios[b * gamma + lane] = ratio;
    }
    const unsigned rej = __ballot_sync(0xffffffffu, !acc);
    n_acc_w0 = rej ? (__ffs(rej) - 1) : gamma;
    tie_w0 = tie && lane < gamma && lane <= n_acc_w0;
    if (p.tie_count != nullptr && tie_w0) atomicAdd(p.tie_count, 1);
  }
  if (warp == 0) {
    const int n_acc = n_acc_w0;
    bool use_q = n_acc < gamma;
    // ---- compact lists of the two rows that matter
    const long long pcr = b * p.pc_req_stride + n_acc * p.pc.row_stride;
    const long long qcr = b * p.qc_req_stride + n_acc * p.qc.row_stride;
    const bool pre = n_acc < kPre;
    const int cp = pre ? p_cnt_pre[n_acc] : p.pc.cnt[pcr];
    const int cq = use_q ? (pre ? q_cnt_pre[n_acc] : p.qc.cnt[qcr]) : 0;
    const bool sparse_ok = cp >= 0 && cp <= kSparseCap && cp <= p.pc.cap && cq >= 0 && cq <= kSparseCap && cq <= p.qc.cap;
    if (lane == 0) { s_n_acc = n_acc; s_dense = sparse_ok ? 0 : 1; }
    if (sparse_ok) {
      for (int t = lane; t < cq; t += 32) {
        if (pre) { q_idx[t] = static_cast<int>(q_pre[n_acc][t].x); q_val[t] = __uint_as_float(q_pre[n_acc][t].y); }
        else { q_idx[t] = p.qc.idx[qcr * p.qc.cap + t]; q_val[t] = p.qc.val[qcr * p.qc.cap + t]; }
      }
      int id[2]; float pv[2], r[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        if (j < cp) {
          id[h] = pre ? static_cast<int>(p_pre[n_acc][j].x) : p.pc.idx[pcr * p.pc.cap + j];
          pv[h] = pre ? __uint_as_float(p_pre[n_acc][j].y) : p.pc.val[pcr * p.pc.cap + j];
        } else { id[h] = 0x7fffffff; pv[h] = 0.f; }
      }
      __syncwarp();
      for (int attempt = 0; attempt < 2; ++attempt) {
        unsigned long long best = 0ull;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float qv = 0.f;
          if (use_q) for (int t = 0; t < cq; ++t) qv = (q_idx[t] == id[h]) ? q_val[t] : qv;
          r[h] = use_q ? fmaxf(pv[h] - qv, 0.f) : pv[h];
          if (r[h] > 0.f) {
            const unsigned long long pk = (static_cast<unsigned long long>(f2key(r[h])) << 32) | (0xffffffffu - static_cast<uint32_t>(id[h]));
            best = pk > best ? pk : best;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
        if (best == 0ull) {
          if (use_q && !p.strict && attempt == 0) { use_q = false; continue; }     // empty residual: resample from p_n
          if (lane == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.next_tok[b] = 0; if (p.n_accepted != nullptr) p.n_accepted[b] = n_acc; }
          break;
        }
        const float rmax = key2f(static_cast<uint32_t>(best >> 32));
        const int argmax = static_cast<int>(0xffffffffu - static_cast<uint32_t>(best & 0xffffffffu));
        const int e = frexp_exp(rmax);
        unsigned long long w[2], tot = 0ull;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          w[h] = weight_of(r[h], e);
          tot += w[h];
          const int j = lane + 32 * h;
          if (j < kSparseCap) { e_idx[j] = id[h]; e_w[j] = w[h]; }
        }
        tot = warp_sum(tot);
        __syncwarp();
        const unsigned long long target = scale_target(tot, u_to_int(u_f));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          unsigned long long before = 0ull;
          for (int t = 0; t < cp; ++t) before += e_idx[t] < id[h] ? e_w[t] : 0ull;
          if (w[h] > 0ull && target >= before && target < before + w[h]) {
            float guard_val = r[h];
            if (use_q) guard_val = __fdiv_rn(r[h], ldexpf(__ull2float_rn(tot), e - kScaleBits) + 1e-6f);
            verify_commit(p, b, n_acc, guard_val < kProbGuard ? argmax : id[h], L_pre);
          }
        }
        break;
      }
    }
  }
  __syncthreads();
  if (s_dense == 0) return;

  // ---- dense fallback for this request: one CTA scans p_n (and q_n) straight from global memory
  dense_verify_cta<kSparseThreads>(p, b, s_n_acc, &rs);
}

// ------------------------------------------------------------------------------------------------
// Multi-draft verify (reference multi_speculative_sampling(strategy='iid'), speculative_sampling.py:1612-1667): W drafts
// of gamma tokens per request.  Warp 0 walks the drafts in order — lane i tests token i of draft w with the uniform at
// the request's running offset (the reference draws its uniforms lazily: a draft consumes one per tested token and
// stops at its first reject, :1616-1634) — and keeps the first draft with the longest accepted run; an all-accepted
// draft ends the scan.  The whole CTA then samples the residual max(0, p_n - q_n) of the winning draft (or its bonus
// row) exactly like the dense path of kernel 2.
constexpr int kMultiThreads = 256;

struct VerifyMultiParams {
  VerifyParams v;                       // p/q/draft describe draft 0 of every request; *_req_stride spans all W drafts
  long long p_draft_stride, q_draft_stride, draft_draft_stride;
  int width;
  int* choice;
};

// ROW: the winning draft's residual row is staged in shared memory (verify_row.cuh, kRowThreads threads); otherwise the
// CTA scans it straight from global memory (vocabularies too large for one CTA's shared memory)
template <bool ROW>
__global__ void __launch_bounds__(ROW ? kRowThreads : kMultiThreads) verify_multi_kernel(const VerifyMultiParams mp) {
  constexpr int THREADS = ROW ? kRowThreads : kMultiThreads;
  __shared__ RowScratch<kMultiThreads> rs;
  __shared__ int s_choice, s_n_acc;
  const VerifyParams& p = mp.v;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = static_cast<int>(p.V), gamma = p.gamma, W = mp.width;
  if (p.active != nullptr && p.active[b] == 0) return;
  if (p.ratios != nullptr) {                                               // statistics: p/q of EVERY drafted token (:1600-1609)
    for (int idx = tid; idx < W * gamma; idx += THREADS) {
      const int w = idx / gamma, i = idx - w * gamma;
      long long tok = p.draft[b * p.draft_stride + w * mp.draft_draft_stride + i];
      if (tok < 0 || tok >= V) tok = 0;
      p.ratios[static_cast<long long>(b) * W * gamma + idx] =
          __fdiv_rn(p.p[b * p.p_req_stride + w * mp.p_draft_stride + i * p.p_row_stride + tok],
                    p.q[b * p.q_req_stride + w * mp.q_draft_stride + i * p.q_row_stride + tok]);
    }
  }
  if (warp == 0) {
    int off = 0, max_l = 0, choice = 0;
    for (int w = 0; w < W; ++w) {
      bool acc = true;
      if (lane < gamma) {
        long long tok = p.draft[b * p.draft_stride + w * mp.draft_draft_stride + lane];
        if (tok < 0 || tok >= V) { atomicOr(p.err_flag, kErrBadToken); tok = 0; }
        const float pv = p.p[b * p.p_req_stride + w * mp.p_draft_stride + lane * p.p_row_stride + tok];
        const float qv = p.q[b * p.q_req_stride + w * mp.q_draft_stride + lane * p.q_row_stride + tok];
        const float ratio = __fdiv_rn(pv, qv);
        const float r = p.u_acc[b * p.u_acc_stride + off + lane];        // (only read where the reference would draw it)
        acc = (ratio == ratio) && (r < fminf(1.0f, ratio));               // torch.min(1, NaN) is NaN: rejects
      }
      const unsigned rej = __ballot_sync(0xffffffffu, !acc);
      const int cur_l = rej ? (__ffs(rej) - 1) : gamma;
      off += rej ? cur_l + 1 : gamma;
      if (cur_l > max_l) {
        max_l = cur_l;
        choice = w;
        if (!rej) break;                                                   // warp-uniform
      }
    }
    if (lane == 0) { s_choice = choice; s_n_acc = max_l; if (mp.choice != nullptr) mp.choice[b] = choice; }
  }
  __syncthreads();
  VerifyParams vp = p;                                                     // the winning draft's rows, non-strict residual rule
  vp.p = p.p + s_choice * mp.p_draft_stride;
  vp.q = p.q + s_choice * mp.q_draft_stride;
  vp.strict = 0;
  if constexpr (ROW) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* row = reinterpret_cast<float*>(smem_raw);
    RowSampleShared& sh = *reinterpret_cast<RowSampleShared*>(smem_raw + ((static_cast<size_t>(V) * 4 + 127) & ~static_cast<size_t>(127)));
    const int n_acc = s_n_acc;
    const float* prow = vp.p + b * vp.p_req_stride + n_acc * vp.p_row_stride;
    const float* qrow = n_acc < gamma ? vp.q + b * vp.q_req_stride + n_acc * vp.q_row_stride : nullptr;
    const long long tok = row_residual_sample(prow, qrow, V, vp.u_final[b], true, row, sh, vp.err_flag);
    if (tok == -2) { if (tid == 0) { vp.next_tok[b] = 0; if (vp.n_accepted != nullptr) vp.n_accepted[b] = n_acc; } }
    else if (tok >= 0) verify_commit(vp, b, n_acc, tok);
  } else {
    dense_verify_cta<kMultiThreads>(vp, b, s_n_acc, &rs);
  }
}

cudaError_t launch_verify_multi(const VerifyParams& v, long long p_draft_stride, long long q_draft_stride,
                                long long draft_draft_stride, int width, int* choice, cudaStream_t st) {
  if (v.gamma < 1 || v.gamma > 32 || width < 1) return cudaErrorInvalidValue;
  VerifyMultiParams mp = {v, p_draft_stride, q_draft_stride, draft_draft_stride, width, choice};
  if (row_kernel_fits(v.V)) {
    static bool attr_dev[64] = {};
    cudaError_t e = set_row_smem(verify_multi_kernel<true>, attr_dev);
    if (e != cudaSuccess) return e;
    verify_multi_kernel<true><<<static_cast<unsigned>(v.B), kRowThreads, row_sample_smem(v.V), st>>>(mp);
  } else {
    verify_multi_kernel<false><<<static_cast<unsigned>(v.B), kMultiThreads, 0, st>>>(mp);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// BiLD check (reference BiLD_sampling, speculative_sampling.py:1793-1813): the target keeps unchecked draft tokens while
// -log p[token] <= rollback_thres, then ALWAYS samples its own next token from its distribution at the first
// position it did not keep (plain sample of a p row — there is no residual in BiLD).
template <bool ROW>
__global__ void __launch_bounds__(ROW ? kRowThreads : kMultiThreads) verify_bild_kernel(const VerifyParams p, const int* n_check, const float fallback_thres,
                                                                                         const float rollback_thres, const int* limit, int* n_drafted, const long long eos) {
  constexpr int THREADS = ROW ? kRowThreads : kMultiThreads;
  __shared__ RowScratch<THREADS> rs;
  __shared__ int s_n;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int V = static_cast<int>(p.V);
  if (p.active != nullptr && p.active[b] == 0) return;
  int nc = n_check != nullptr ? min(n_check[b], p.gamma) : p.gamma;
  if (p.q != nullptr) {
    // engine mode: gamma tokens were drafted up front; the reference would have stopped drafting at the first token whose
    // distribution was unsure (max q < fallback_thres, speculative_sampling.py:1784) — tokens after it do not exist for it
    RowCtx<THREADS> cx(&rs, 1);
    nc = p.gamma;
    for (int i = 0; i < p.gamma; ++i) {
      const float* qrow = p.q + b * p.q_req_stride + i * p.q_row_stride;
      float m = 0.f;
      const int qcnt = p.qc.cnt != nullptr ? p.qc.cnt[(b * p.qc_req_stride + i) * p.qc.row_stride] : -1;
      if (qcnt > 0 && qcnt <= p.qc.cap) {
        // kernel 1 left the row's non-zeros in its compact list: the maximum of <= cap values instead of a scan of V
        const long long cr = (b * p.qc_req_stride + i) * p.qc.row_stride * p.qc.cap;
        for (int j = tid; j < qcnt; j += THREADS) m = fmaxf(m, p.qc.val[cr + j]);
      } else if ((V & 3) == 0 && (reinterpret_cast<uintptr_t>(qrow) & 15) == 0) {
        for (int j = tid; j < (V >> 2); j += THREADS) {
          const uint4 qv = ld_nc_v4(reinterpret_cast<const uint4*>(qrow) + j);
          m = fmaxf(fmaxf(m, fmaxf(__uint_as_float(qv.x), __uint_as_float(qv.y))), fmaxf(__uint_as_float(qv.z), __uint_as_float(qv.w)));
        }
      } else {
        for (int j = tid; j < V; j += THREADS) m = fmaxf(m, qrow[j]);
      }
      m = cx.allreduce_max(m);
      if (m < fallback_thres) { nc = i + 1; break; }                       // block-uniform
    }
  }
  if (n_drafted != nullptr && tid == 0) n_drafted[b] = nc;
  if (eos >= 0 && p.q != nullptr && p.seq_len != nullptr) {
    // the reference tests for EOS after EVERY draft token (speculative_sampling.py:1826-1841): an EOS drafted before the
    // token that triggers the check ends generation right there, with the drafted tokens kept unchecked
    int first = nc;
    for (int i = 0; i < nc - 1; ++i)
      if (p.draft[b * p.draft_stride + i] == eos) { first = i; break; }
    if (first < nc - 1) {
      const int keep = limit != nullptr ? min(first + 1, max(limit[b] - p.seq_len[b], 0)) : first + 1;
      if (tid == 0) {
        p.n_accepted[b] = -1 - keep;
        p.next_tok[b] = -1;
        p.seq_len[b] += keep;
        if (n_drafted != nullptr) n_drafted[b] = keep;
      }
      return;
    }
  }
  if (limit != nullptr && p.seq_len != nullptr) {
    // the reference tests its length limit before every draft token (:1764): with fewer than nc tokens of room it leaves
    // the loop with the drafted tokens unchecked and no target token
    const int room = limit[b] - p.seq_len[b];
    if (room < nc) {
      if (tid == 0) {
        const int keep = max(room, 0);
        p.n_accepted[b] = -1 - keep;                                       // (< 0: no check happened; -1 - kept tokens)
        p.next_tok[b] = -1;
        p.seq_len[b] += keep;
        if (n_drafted != nullptr) n_drafted[b] = keep;
      }
      return;
    }
  }
  if (tid < 32) {
    bool fail = false;
    if (lane < nc) {
      long long tok = p.draft[b * p.draft_stride + lane];
      if (tok < 0 || tok >= V) { atomicOr(p.err_flag, kErrBadToken); tok = 0; }
      const float nll = -logf(p.p[b * p.p_req_stride + lane * p.p_row_stride + tok]);
      fail = nll > rollback_thres;                                         // (-log 0 = inf fails; NaN keeps, as in the reference)
      if (p.ratios != nullptr) p.ratios[b * p.gamma + lane] = nll;
    }
    const unsigned bad = __ballot_sync(0xffffffffu, fail);
    if (lane == 0) s_n = bad ? (__ffs(bad) - 1) : nc;
  }
  __syncthreads();
  VerifyParams vp = p;
  vp.gamma = 0;                                                            // never a residual: plain sample of p row n
  vp.strict = 0;
  vp.q = nullptr;
  if constexpr (ROW) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* row = reinterpret_cast<float*>(smem_raw);
    RowSampleShared& sh = *reinterpret_cast<RowSampleShared*>(smem_raw + ((static_cast<size_t>(V) * 4 + 127) & ~static_cast<size_t>(127)));
    const int n = s_n;
    const long long tok = row_residual_sample(vp.p + b * vp.p_req_stride + n * vp.p_row_stride, nullptr, V, vp.u_final[b], false, row, sh, vp.err_flag);
    if (tok == -2) { if (tid == 0) { vp.next_tok[b] = 0; if (vp.n_accepted != nullptr) vp.n_accepted[b] = n; } }
    else if (tok >= 0) verify_commit(vp, b, n, tok);
  } else {
    dense_verify_cta<kMultiThreads>(vp, b, s_n, &rs);
  }
}

cudaError_t launch_verify_bild(const VerifyParams& v, const int* n_check, float fallback_thres, float rollback_thres,
                               const int* limit, int* n_drafted, long long eos, cudaStream_t st) {
  if (v.gamma < 1 || v.gamma > 32) return cudaErrorInvalidValue;
  if (row_kernel_fits(v.V)) {
    static bool attr_dev[64] = {};
    cudaError_t e = set_row_smem(verify_bild_kernel<true>, attr_dev);
    if (e != cudaSuccess) return e;
    verify_bild_kernel<true><<<static_cast<unsigned>(v.B), kRowThreads, row_sample_smem(v.V), st>>>(v, n_check, fallback_thres, rollback_thres, limit, n_drafted, eos);
  } else {
    verify_bild_kernel<false><<<static_cast<unsigned>(v.B), kMultiThreads, 0, st>>>(v, n_check, fallback_thres, rollback_thres, limit, n_drafted, eos);
  }
  return cudaGetLastError();
}

static int g_verify_cluster = 0;
void set_verify_tuning(int cluster) { g_verify_cluster = cluster; }

cudaError_t launch_verify(const VerifyParams& pin, cudaStream_t st) {
  VerifyParams p = pin;
  constexpr int THREADS = 256;
  if (p.gamma > 32 || (p.q != nullptr && p.gamma < 1)) return cudaErrorInvalidValue;
  if (p.q != nullptr && p.pc.cnt != nullptr && p.qc.cnt != nullptr) {      // compact lists available: sparse path
    cudaLaunchConfig_t scfg = {};
    scfg.gridDim = dim3(static_cast<unsigned>(p.B));
    scfg.blockDim = dim3(kSparseThreads);
    scfg.stream = st;
    cudaLaunchAttribute sat[1];
    sat[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    sat[0].val.programmaticStreamSerializationAllowed = 1;
    scfg.attrs = sat; scfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&scfg, verify_sparse_kernel, p);
  }
  // two CTAs per request while that still fits the machine (aligned rows with a q side; see verify_row2_kernel)
  if (g_verify_cluster == 0 && p.q != nullptr && row_kernel_fits(p.V) && 2 * p.B <= device_sm_count() && (p.V & 3) == 0 && p.V >= 8192 &&
      reinterpret_cast<uintptr_t>(p.p) % 16 == 0 && reinterpret_cast<uintptr_t>(p.q) % 16 == 0 && p.p_req_stride % 4 == 0 &&
      p.p_row_stride % 4 == 0 && p.q_req_stride % 4 == 0 && p.q_row_stride % 4 == 0) {
    static bool attr_dev2[64] = {};
    cudaError_t e = set_row_smem(verify_row2_kernel, attr_dev2);
    if (e != cudaSuccess) return e;
    const long long n_vec = p.V >> 2, n_qc = (n_vec + 1023) >> 10;
    const long long vec0 = std::min<long long>(n_vec, ((n_qc + 1) >> 1) << 10);
    const size_t row_al = (static_cast<size_t>(vec0) * 16 + 127) & ~static_cast<size_t>(127);
    const long long room = static_cast<long long>(device_max_smem_optin()) - kRowStaticReserve - static_cast<long long>(row_al + sizeof(RowSampleShared));
    const int qs = static_cast<int>(room / kRowQChunkBytes);
    p.q_slots = qs > kRowQMaxSlots ? kRowQMaxSlots : qs;       // (half a row leaves room for the full ring)
    cudaLaunchConfig_t ccfg = {};
    ccfg.gridDim = dim3(static_cast<unsigned>(2 * p.B));
    ccfg.blockDim = dim3(kRowThreads);
    ccfg.dynamicSmemBytes = row_al + static_cast<size_t>(p.q_slots) * kRowQChunkBytes + sizeof(RowSampleShared);
    ccfg.stream = st;
    cudaLaunchAttribute cat[2];
    cat[0].id = cudaLaunchAttributeClusterDimension;
    cat[0].val.clusterDim.x = 2; cat[0].val.clusterDim.y = 1; cat[0].val.clusterDim.z = 1;
    cat[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    cat[1].val.programmaticStreamSerializationAllowed = 1;
    ccfg.attrs = cat; ccfg.numAttrs = pdl_enabled() ? 2 : 1;
    p.prof = nullptr;
    return cudaLaunchKernelEx(&ccfg, verify_row2_kernel, p);
  }
  if (g_verify_cluster == 0 && row_kernel_fits(p.V)) {                     // one CTA per request, row staged in shared memory
    static bool attr_dev[64] = {};
    cudaError_t e = set_row_smem(verify_row_kernel, attr_dev);
    if (e != cudaSuccess) return e;
    // q ring: as many 16 KB slots (up to 4) as fit behind the staged p row; fewer than 2: q through per-thread loads
    {
      const long long room = static_cast<long long>(device_max_smem_optin()) - kRowStaticReserve - static_cast<long long>(row_sample_smem(p.V));
      const int qs = p.q != nullptr ? static_cast<int>(room / kRowQChunkBytes) : 0;
      p.q_slots = qs >= 2 ? (qs > kRowQMaxSlots ? kRowQMaxSlots : qs) : 0;
    }
    cudaLaunchConfig_t rcfg = {};
    rcfg.gridDim = dim3(static_cast<unsigned>(p.B));
    rcfg.blockDim = dim3(kRowThreads);
    rcfg.dynamicSmemBytes = row_sample_smem(p.V, p.q_slots);
    rcfg.stream = st;
    cudaLaunchAttribute rat[1];
    rat[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    rat[0].val.programmaticStreamSerializationAllowed = 1;
    rcfg.attrs = rat; rcfg.numAttrs = pdl_enabled() ? 1 : 0;
    p.prof = get_norm_prof();
    return cudaLaunchKernelEx(&rcfg, verify_row_kernel, p);
  }
  const long long row_bytes = p.V * 4;
  int C = 1;
  while (C < kMaxPortableCluster && (row_bytes + C - 1) / C > 48 * 1024) C <<= 1;
  while (C < kMaxPortableCluster && static_cast<long long>(p.B) * C < device_sm_count() && row_bytes / (2 * C) >= 4096) C <<= 1;
  if (g_verify_cluster > 0) C = g_verify_cluster;
  long long slice = ((p.V + C - 1) / C + 127) & ~127LL;
  while (C > 1 && slice * (C - 1) >= p.V) { C >>= 1; slice = ((p.V + C - 1) / C + 127) & ~127LL; }
  p.cluster = C;
  p.slice_elems = static_cast<int>(slice);
  auto al = [](const void* ptr, long long s1, long long s2) {
    return reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && s1 % 4 == 0 && s2 % 4 == 0;
  };
  p.use_tma = (al(p.p, p.p_req_stride, p.p_row_stride) && (p.q == nullptr || al(p.q, p.q_req_stride, p.q_row_stride)) &&
               p.V % 4 == 0) ? 1 : 0;
  const size_t smem = static_cast<size_t>(slice) * 4 + sizeof(VerifyShared<THREADS>);
  if (smem > device_max_smem_optin()) return cudaErrorInvalidValue;
  auto kern = verify_kernel<THREADS, 3>;
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
  cfg.gridDim = dim3(static_cast<unsigned>(p.B) * C);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

}  // namespace sd
