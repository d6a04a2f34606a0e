// Verify building blocks shared by the stand-alone verify kernels (verify.cu) and the norm kernel that verifies a
// request as soon as its last row is normalised (norm_pipe.cu):  reference
// /root/reference/sampling/speculative_sampling.py:1966-2027 (strict = 0) and :2147-2185 (strict = 1).
#pragma once

#include "rowops.cuh"
#include "specdec_internal.h"

namespace sd {

constexpr int kSparseCap = 64;     // longest compact list the sparse verify path handles

__device__ __forceinline__ void verify_commit(const VerifyParams& p, int b, int n_acc, long long out, int L_pre = -1) {
  p.next_tok[b] = out;
  if (p.n_accepted != nullptr) p.n_accepted[b] = n_acc;
  if (p.stats != nullptr) { atomicAdd(&p.stats[0], static_cast<unsigned long long>(n_acc)); atomicAdd(&p.stats[1], 1ull); }
  if (p.tokens != nullptr) {
    const int L = L_pre >= 0 ? L_pre : p.seq_len[b];
    p.tokens[b * p.tokens_stride + L + n_acc] = out;
    p.seq_len[b] = L + n_acc + 1;
  }
}

// Dense residual / resample for request b after n_acc accepted tokens: all THREADS threads of the CTA scan p_n (and
// q_n) straight from global memory.  `rs` is CTA-wide scratch.
template <int THREADS>
__device__ void dense_verify_cta(const VerifyParams& p, int b, int n_acc, RowScratch<THREADS>* rs) {
  RowCtx<THREADS> cx(rs, 1);
  const int tid = threadIdx.x;
  const int V = static_cast<int>(p.V), gamma = p.gamma;
  bool use_q = n_acc < gamma;
  const float* prow = p.p + b * p.p_req_stride + n_acc * p.p_row_stride;
  const float* qrow = p.q + b * p.q_req_stride + n_acc * p.q_row_stride;
  const int n_vec = (V + 3) / 4;
  auto vecw = [&](int v, float (&w)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = v * 4 + j;
      float a = i < V ? prow[i] : 0.f;
      if (use_q && i < V) a = fmaxf(a - qrow[i], 0.f);
      w[j] = a;
    }
  };
  unsigned long long best = 0ull;
  for (int attempt = 0; attempt < 2; ++attempt) {
    unsigned long long mine = 0ull;
    bool bad = false;
    for (int v = tid; v < n_vec; v += THREADS) {
      float w[4];
      vecw(v, w);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bad |= !(w[j] >= 0.f) || isinf(w[j]);
        if (w[j] > 0.f) {
          const unsigned long long pk = (static_cast<unsigned long long>(f2key(w[j])) << 32) | (0xffffffffu - static_cast<uint32_t>(v * 4 + j));
          mine = pk > mine ? pk : mine;
        }
      }
    }
    if (bad) atomicOr(p.err_flag, kErrEmptyRow);
    best = cx.allreduce_max(mine);
    if (best != 0ull || !use_q || p.strict) break;
    use_q = false;
  }
  if (best == 0ull) {
    if (tid == 0) { atomicOr(p.err_flag, kErrEmptyRow); p.next_tok[b] = 0; if (p.n_accepted != nullptr) p.n_accepted[b] = n_acc; }
    return;
  }
  const float rmax = key2f(static_cast<uint32_t>(best >> 32));
  const int argmax = static_cast<int>(0xffffffffu - static_cast<uint32_t>(best & 0xffffffffu));
  unsigned long long total = 0ull;
  float psel = 1.f;
  const int tok_d = cluster_icdf<4, THREADS>(cx, n_vec, 0, rmax, p.u_final[b], vecw, &total, &psel);
  if (tok_d >= 0) {
    float guard_val = psel;
    if (use_q) guard_val = __fdiv_rn(psel, ldexpf(__ull2float_rn(total), frexp_exp(rmax) - kScaleBits) + 1e-6f);
    verify_commit(p, b, n_acc, guard_val < kProbGuard ? argmax : tok_d);
  }
}

// Scratch of sparse_verify_warp: 1280 bytes, 8-byte aligned.
struct SparseVerifyScratch {
  unsigned long long e_w[kSparseCap];
  int e_idx[kSparseCap];
  int q_idx[kSparseCap];
  float q_val[kSparseCap];
};

// The whole verify step of request b by ONE warp, from the compact lists in global memory (read through L2 with
// ld.cg: inside the norm kernel they were just written by other SMs).  Same arithmetic as verify_sparse_kernel.
// Returns -1 when the request is done, otherwise n_acc of a request that needs dense_verify_cta (a list is missing).
__device__ __forceinline__ int sparse_verify_warp(const VerifyParams& p, int b, int lane, SparseVerifyScratch* sc) {
  const int V = static_cast<int>(p.V), gamma = p.gamma;
  if (p.active != nullptr && __ldcg(p.active + b) == 0) return -1;
  int L_pre = -1;
  if (p.tokens != nullptr) L_pre = __ldcg(p.seq_len + b);
  const float u_f = __ldcg(p.u_final + b);
  // accept scan: lane i owns drafted token i.  Its p / q values come from the compact lists of row i, which the WHOLE
  // warp fetches (two entries per lane and list, four rows per round, every load of a round issued before the first
  // use) — a per-lane walk over a list would be a chain of ~2 x cnt dependent L2 round trips.
  long long tok = 0;
  float u_a = 0.f, pv = 0.f, qv = 0.f;
  int cpl = -1, cql = -1;
  if (lane < gamma) {
    tok = __ldcg(p.draft + b * p.draft_stride + lane);
    u_a = __ldcg(p.u_acc + b * p.u_acc_stride + lane);
    cpl = __ldcg(p.pc.cnt + b * p.pc_req_stride + lane * p.pc.row_stride);
    cql = __ldcg(p.qc.cnt + b * p.qc_req_stride + lane * p.qc.row_stride);
    if (tok < 0 || tok >= V) { atomicOr(p.err_flag, kErrBadToken); tok = 0; }
    if (!(cpl >= 0 && cpl <= kSparseCap && cpl <= p.pc.cap)) { cpl = -1; pv = __ldcg(p.p + b * p.p_req_stride + lane * p.p_row_stride + tok); }
    if (!(cql >= 0 && cql <= kSparseCap && cql <= p.qc.cap)) { cql = -1; qv = __ldcg(p.q + b * p.q_req_stride + lane * p.q_row_stride + tok); }
  }
  for (int i0 = 0; i0 < gamma; i0 += 4) {
    // eight lanes per row, four rows per round: lane sub = lane % 8 walks entries sub, sub + 8, ... of both lists
    const int i = i0 + (lane >> 3), sub = lane & 7, ic = min(i, gamma - 1);
    const long long tok_i = __shfl_sync(0xffffffffu, tok, ic);
    const int cp_i = i < gamma ? __shfl_sync(0xffffffffu, cpl, ic) : (__shfl_sync(0xffffffffu, cpl, ic), -1);
    const int cq_i = i < gamma ? __shfl_sync(0xffffffffu, cql, ic) : (__shfl_sync(0xffffffffu, cql, ic), -1);
    const long long pcr = (b * p.pc_req_stride + ic * p.pc.row_stride) * p.pc.cap, qcr = (b * p.qc_req_stride + ic * p.qc.row_stride) * p.qc.cap;
    float fp = 0.f, fq = 0.f;                                 // listed probabilities are > 0 and indices are unique: max = lookup
#pragma unroll 4
    for (int j = sub; j < cp_i; j += 8) {
      const int id = __ldcg(p.pc.idx + pcr + j);
      const float v = __ldcg(p.pc.val + pcr + j);
      fp = id == tok_i ? v : fp;
    }
#pragma unroll 4
    for (int j = sub; j < cq_i; j += 8) {
      const int id = __ldcg(p.qc.idx + qcr + j);
      const float v = __ldcg(p.qc.val + qcr + j);
      fq = id == tok_i ? v : fq;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      fp = fmaxf(fp, __shfl_xor_sync(0xffffffffu, fp, o));
      fq = fmaxf(fq, __shfl_xor_sync(0xffffffffu, fq, o));
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float a = __shfl_sync(0xffffffffu, fp, 8 * r), c = __shfl_sync(0xffffffffu, fq, 8 * r);
      if (lane == i0 + r) { if (cpl >= 0) pv = a; if (cql >= 0) qv = c; }
    }
  }
  bool acc = true, tie = false;
  if (lane < gamma) {
    if (qv == 0.f) atomicOr(p.err_flag, kErrZeroQ);
    const float ratio = __fdiv_rn(pv, qv);
    const float thr = p.strict ? fminf(1.0f, ratio) : ratio;
    acc = p.strict ? (u_a < thr) : !(u_a > thr);
    tie = (u_a == thr);
    if (p.ratios != nullptr) p.ratios[b * gamma + lane] = ratio;
  }
  const unsigned rej = __ballot_sync(0xffffffffu, !acc);
  const int n_acc = rej ? (__ffs(rej) - 1) : gamma;
  if (p.tie_count != nullptr && tie && lane < gamma && lane <= n_acc) atomicAdd(p.tie_count, 1);

  bool use_q = n_acc < gamma;
  const long long pcr = b * p.pc_req_stride + n_acc * p.pc.row_stride;
  const long long qcr = b * p.qc_req_stride + n_acc * p.qc.row_stride;
  const int cp = __ldcg(p.pc.cnt + pcr);
  const int cq = use_q ? __ldcg(p.qc.cnt + qcr) : 0;
  const bool sparse_ok = cp >= 0 && cp <= kSparseCap && cp <= p.pc.cap && cq >= 0 && cq <= kSparseCap && cq <= p.qc.cap;
  if (!sparse_ok) return n_acc;
  for (int t = lane; t < cq; t += 32) {
    sc->q_idx[t] = __ldcg(p.qc.idx + qcr * p.qc.cap + t);
    sc->q_val[t] = __ldcg(p.qc.val + qcr * p.qc.cap + t);
  }
  int id[2]; float pn[2], r[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    if (j < cp) { id[h] = __ldcg(p.pc.idx + pcr * p.pc.cap + j); pn[h] = __ldcg(p.pc.val + pcr * p.pc.cap + j); }
    else { id[h] = 0x7fffffff; pn[h] = 0.f; }
  }
  __syncwarp();
  for (int attempt = 0; attempt < 2; ++attempt) {
    unsigned long long best = 0ull;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float qv = 0.f;
      if (use_q) for (int t = 0; t < cq; ++t) qv = (sc->q_idx[t] == id[h]) ? sc->q_val[t] : qv;
      r[h] = use_q ? fmaxf(pn[h] - qv, 0.f) : pn[h];
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
      if (j < kSparseCap) { sc->e_idx[j] = id[h]; sc->e_w[j] = w[h]; }
    }
    tot = warp_sum(tot);
    __syncwarp();
    const unsigned long long target = scale_target(tot, u_to_int(u_f));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      unsigned long long before = 0ull;
      for (int t = 0; t < cp; ++t) before += sc->e_idx[t] < id[h] ? sc->e_w[t] : 0ull;
      if (w[h] > 0ull && target >= before && target < before + w[h]) {
        float guard_val = r[h];
        if (use_q) guard_val = __fdiv_rn(r[h], ldexpf(__ull2float_rn(tot), e - kScaleBits) + 1e-6f);
        verify_commit(p, b, n_acc, guard_val < kProbGuard ? argmax : id[h], L_pre);
      }
    }
    break;
  }
  __syncwarp();
  return -1;
}

}  // namespace sd
