// Dense residual / resample step of kernel 2 for ONE request by ONE CTA, with the row staged in shared memory:
//   p_n is brought in by 1-D TMA bulk copies (32 KB chunks, one mbarrier each), q_n is read from HBM exactly once
//   straight into registers, the residual max(0, p_n - q_n) replaces p_n in place, and the two passes that follow
//   (row maximum -> fixed-point scale; exact weight sums -> inverse-CDF position) never leave the SM.
// Every warp owns one CONTIGUOUS range of the vocabulary, so the exact prefix sums of the sampler are a warp scan over
// per-warp totals plus a walk of the owning warp's range; per-element weights are accumulated as two 20-bit limbs with
// round-down float adds (no float -> u64 conversion per element, see norm_ring_kernel.cuh).
//
// Replaces /root/reference/sampling/speculative_sampling.py:2005-2023 (sample(max_fn(p - q)) with its fall-back to
// sample(p), bonus sample) and sampling/utils.py:213-245; shared by sd_verify (dense path), sd_sample, sd_verify_multi
// and sd_verify_bild whenever a row fits one CTA's shared memory (V <= ~57k fp32 probabilities).
#pragma once

#include "rowops.cuh"
#include "specdec_internal.h"
#include "verify_sparse.cuh"

namespace sd {

constexpr int kRowThreads = 512;
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kRowChunkBytes = 32768;
constexpr int kRowMaxChunks = 8;

constexpr int kRowQChunkBytes = 16384;                          // q row: streamed through a small ring of TMA chunks
constexpr int kRowQMaxSlots = 4;

struct alignas(16) RowSampleShared {
  uint64_t bar[kRowMaxChunks];
  uint64_t qfull[kRowQMaxSlots], qempty[kRowQMaxSlots];
  unsigned long long wbest[kRowWarps];
  unsigned long long wsum[kRowWarps];
  unsigned long long xbest[2][2];      // cluster of 2 (one request on two SMs): maxima of the two halves, per attempt
  unsigned long long xtot[2];          // ... and their exact weight totals
  int n_acc, aux;
};

// bytes of dynamic shared memory a CTA needs for a row of V probabilities
static inline size_t row_sample_smem(long long V, int q_slots = 0) {
  return ((static_cast<size_t>(V) * 4 + 127) & ~static_cast<size_t>(127)) + static_cast<size_t>(q_slots) * kRowQChunkBytes + sizeof(RowSampleShared);
}

// The barriers of row_residual_sample, initialised by ONE thread ahead of time (e.g. while another warp runs the accept scan);
// a CTA-wide barrier must follow before row_residual_sample(..., barriers_ready = true) is called.
__device__ __forceinline__ void row_sample_init_barriers(RowSampleShared& sh, int q_slots) {
  for (int c = 0; c < kRowMaxChunks; ++c) mbar_init(&sh.bar[c], 1);
  for (int c = 0; c < q_slots; ++c) { mbar_init(&sh.qfull[c], 1); mbar_init(&sh.qempty[c], kRowWarps); }
  fence_barrier_init();
}

// Samples request b's next token from max(0, prow - qrow) (qrow == nullptr: from prow), falling back to prow when the
// residual is empty and `fallback` is set (reference :2009-2010).  All kRowThreads threads of the CTA must call it;
// `row` is the CTA's staging area (>= V floats, 16-byte aligned), `sh` its scratch (barriers NOT yet initialised).
// On success exactly one thread gets a token >= 0 back (with the guard value for the < 1e-9 rule applied); the others -1.
// Returns -2 in every thread when there is nothing to sample from (err_flag is set).
__device__ __forceinline__ long long row_residual_sample(const float* __restrict__ prow, const float* __restrict__ qrow, int V,
                                                         float u_final, bool fallback, float* row, RowSampleShared& sh,
                                                         int* err_flag, long long* prof = nullptr, unsigned char* qring = nullptr,
                                                         int q_slots = 0, bool barriers_ready = false, int crank = 0, int csize = 1,
                                                         int idx_off = 0) {
  // csize == 2: the CTA is one of a cluster of two that share the row — prow / qrow / V describe THIS CTA's part, idx_off is
  // the vocabulary index of its first element; the maximum and the weight total are exchanged through distributed shared
  // memory (two cluster barriers), the CTA whose part holds the target position returns the token
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_vec = (V + 3) >> 2;
  const bool vec_ok = (V & 3) == 0 && (reinterpret_cast<uintptr_t>(prow) & 15) == 0 &&
                      (qrow == nullptr || (reinterpret_cast<uintptr_t>(qrow) & 15) == 0);
  const uint32_t row_bytes = static_cast<uint32_t>(V) * 4u;
  const int n_chunks = vec_ok ? static_cast<int>((row_bytes + kRowChunkBytes - 1) / kRowChunkBytes) : 0;
  const int n_qchunks = static_cast<int>((row_bytes + kRowQChunkBytes - 1) / kRowQChunkBytes);
  const bool q_ring = vec_ok && qrow != nullptr && qring != nullptr && q_slots >= 2;   // q streamed by TMA instead of per-thread loads
  // warp w owns vectors [w * vpw, (w + 1) * vpw), lanes strided inside
  const int vpw = (n_vec + kRowWarps - 1) / kRowWarps;
  const int v_begin = warp * vpw, v_end = min(n_vec, v_begin + vpw);
  float4* row4 = reinterpret_cast<float4*>(row);

  if (vec_ok) {
    if (tid == 0) {
      if (!barriers_ready) {                                    // (else: row_sample_init_barriers ran before a CTA-wide barrier)
        for (int c = 0; c < n_chunks; ++c) mbar_init(&sh.bar[c], 1);
        if (q_ring) for (int c = 0; c < q_slots; ++c) { mbar_init(&sh.qfull[c], 1); mbar_init(&sh.qempty[c], kRowWarps); }
        fence_barrier_init();
      }
    }
    if (!barriers_ready) __syncthreads();                       // barriers initialised before anyone waits on them
    // one bulk copy per thread (a single thread needs ~200 cycles per issue: 8 copies in a row were 1.7k cycles of the
    // request's critical path): lanes 0.. of warp 0 the p chunks, lanes 0.. of warp 1 the first q chunks
    if (tid < n_chunks) {
      const uint32_t off = static_cast<uint32_t>(tid) * kRowChunkBytes;
      const uint32_t bytes = min(static_cast<uint32_t>(kRowChunkBytes), row_bytes - off);
      mbar_expect_tx(&sh.bar[tid], bytes);
      tma_load_1d(reinterpret_cast<unsigned char*>(row) + off, reinterpret_cast<const unsigned char*>(prow) + off, bytes, &sh.bar[tid]);
    }
    if (q_ring && tid >= 32 && tid - 32 < min(q_slots, n_qchunks)) {
      const int cq = tid - 32;
      const uint32_t qoff = static_cast<uint32_t>(cq) * kRowQChunkBytes;
      const uint32_t qb = min(static_cast<uint32_t>(kRowQChunkBytes), row_bytes - qoff);
      mbar_expect_tx(&sh.qfull[cq], qb);
      tma_load_1d(qring + static_cast<size_t>(cq) * kRowQChunkBytes, reinterpret_cast<const unsigned char*>(qrow) + qoff, qb, &sh.qfull[cq]);
    }
  } else {
    for (int i = tid; i < n_vec * 4; i += kRowThreads) row[i] = i < V ? prow[i] : 0.f;
    __syncthreads();
  }

  if (prof != nullptr && threadIdx.x == 0) prof[2] = clock64();   // loads issued
  bool use_q = qrow != nullptr;
  unsigned long long best = 0ull;
  for (int attempt = 0; attempt < 2; ++attempt) {
    // ---- pass 1: residual in place (own vectors only), row maximum as a packed (value key, ~index)
    unsigned long long mine = 0ull;
    bool bad = false;
    if (attempt == 0 && q_ring) {
      // q arrives chunk by chunk in a small ring (the SM keeps p: V * 4 bytes + q: q_slots * 16 KB in flight through TMA);
      // chunk c holds vectors c * 1024 ..: thread t takes vectors t and t + 512 of it
      int slot = 0;
      uint32_t par = 0u;
      for (int c = 0; c < n_qchunks; ++c) {
        mbar_wait(&sh.qfull[slot], par);
        mbar_wait(&sh.bar[(c * kRowQChunkBytes) / kRowChunkBytes], 0);
        const uint4* q4 = reinterpret_cast<const uint4*>(qring + static_cast<size_t>(slot) * kRowQChunkBytes);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int vl = h * kRowThreads + tid, v = c * (kRowQChunkBytes / 16) + vl;
          if (v < n_vec) {
            const float4 a4 = row4[v];
            const uint4 qv = q4[vl];
            float w[4] = {fmaxf(a4.x - __uint_as_float(qv.x), 0.f), fmaxf(a4.y - __uint_as_float(qv.y), 0.f),
                          fmaxf(a4.z - __uint_as_float(qv.z), 0.f), fmaxf(a4.w - __uint_as_float(qv.w), 0.f)};   // utils.py:240
            row4[v] = make_float4(w[0], w[1], w[2], w[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              bad |= !(w[j] >= 0.f) || isinf(w[j]);
              const unsigned long long pk = (static_cast<unsigned long long>(f2key(w[j])) << 32) | (0xffffffffu - static_cast<uint32_t>(idx_off + v * 4 + j));
              mine = (w[j] > 0.f && pk > mine) ? pk : mine;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.qempty[slot]);
        if (tid == 0 && c + q_slots < n_qchunks) {               // refill the slot with chunk c + q_slots once all 16 warps released it
          mbar_wait(&sh.qempty[slot], par);
          const uint32_t qoff = static_cast<uint32_t>(c + q_slots) * kRowQChunkBytes;
          const uint32_t qb = min(static_cast<uint32_t>(kRowQChunkBytes), row_bytes - qoff);
          mbar_expect_tx(&sh.qfull[slot], qb);
          tma_load_1d(qring + static_cast<size_t>(slot) * kRowQChunkBytes, reinterpret_cast<const unsigned char*>(qrow) + qoff, qb, &sh.qfull[slot]);
        }
        if (++slot == q_slots) { slot = 0; par ^= 1u; }
      }
    } else {
    constexpr int kAhead = 8;                                   // q vectors requested ahead of their use
    uint4 qbuf[kAhead];
    if (use_q && vec_ok) {
#pragma unroll
      for (int a = 0; a < kAhead; ++a) {
        const int v = v_begin + lane + 32 * a;
        qbuf[a] = v < v_end ? ld_nc_v4(reinterpret_cast<const uint4*>(qrow) + v) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    for (int v0 = v_begin; v0 < v_end; v0 += 32 * kAhead) {
#pragma unroll
      for (int a = 0; a < kAhead; ++a) {
        const int v = v0 + 32 * a + lane;
        uint4 qv = make_uint4(0u, 0u, 0u, 0u);
        if (use_q && vec_ok) {
          qv = qbuf[a];
          const int vn = v + 32 * kAhead;
          qbuf[a] = vn < v_end ? ld_nc_v4(reinterpret_cast<const uint4*>(qrow) + vn) : make_uint4(0u, 0u, 0u, 0u);
        }
        if (v < v_end) {
          if (attempt == 0 && vec_ok) mbar_wait(&sh.bar[(v * 16) / kRowChunkBytes], 0);
          float4 a4 = row4[v];
          float w[4] = {a4.x, a4.y, a4.z, a4.w};
          if (use_q) {
            float qq[4];
            if (vec_ok) { qq[0] = __uint_as_float(qv.x); qq[1] = __uint_as_float(qv.y); qq[2] = __uint_as_float(qv.z); qq[3] = __uint_as_float(qv.w); }
            else {
#pragma unroll
              for (int j = 0; j < 4; ++j) qq[j] = v * 4 + j < V ? qrow[v * 4 + j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = fmaxf(w[j] - qq[j], 0.f);           // utils.py:240
            row4[v] = make_float4(w[0], w[1], w[2], w[3]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            bad |= !(w[j] >= 0.f) || isinf(w[j]);
            const unsigned long long pk = (static_cast<unsigned long long>(f2key(w[j])) << 32) | (0xffffffffu - static_cast<uint32_t>(idx_off + v * 4 + j));
            mine = (w[j] > 0.f && pk > mine) ? pk : mine;
          }
        }
      }
    }
    }
    if (bad) atomicOr(err_flag, kErrEmptyRow);                  // negative / NaN / inf weights: 'prob error'
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, mine, o); mine = t > mine ? t : mine; }
    __syncthreads();                                            // (wbest of a previous attempt has been read by everyone)
    if (lane == 0) sh.wbest[warp] = mine;
    __syncthreads();
    best = lane < kRowWarps ? sh.wbest[lane] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
    if (csize > 1) {                                            // the other half's maximum
      cg::cluster_group cl = cg::this_cluster();
      if (tid == 0) {
        sh.xbest[attempt][crank] = best;
        cl.map_shared_rank(&sh, crank ^ 1)->xbest[attempt][crank] = best;
      }
      cl.sync();
      const unsigned long long b0 = sh.xbest[attempt][0], b1 = sh.xbest[attempt][1];
      best = b0 > b1 ? b0 : b1;
    }
    if (best != 0ull || !use_q || !fallback) break;
    // empty residual: resample from p_n itself (reference :2009-2010) — bring the row back (rare: plain loads)
    use_q = false;
    __syncthreads();
    for (int i = tid; i < n_vec * 4; i += kRowThreads) row[i] = i < V ? prow[i] : 0.f;
    __syncthreads();
  }
  if (best == 0ull) {
    if (tid == 0) atomicOr(err_flag, kErrEmptyRow);
    return -2;
  }
  if (prof != nullptr && threadIdx.x == 0) prof[3] = clock64();   // pass 1 done, maximum known
  const float rmax = key2f(static_cast<uint32_t>(best >> 32));
  const int argmax = static_cast<int>(0xffffffffu - static_cast<uint32_t>(best & 0xffffffffu));
  const int e = frexp_exp(rmax);

  // ---- pass 2: exact weight sums, w = floor(r * 2^(40 - e)) as two 20-bit limbs (packed fp32, round-down adds: no
  //      float -> u64 conversion per element).  One "block" = the 32 vectors a warp reads in one step; lane k keeps the
  //      sum of the warp's block k, so that the walk below is two warp scans instead of one per block.
  const float scale = ldexpf(1.0f, kScaleBits - e);
  const float scale_hi = scale * 9.5367431640625e-07f;          // 2^-20 * scale (exact)
  const f32x2 sh2 = pack2(scale_hi, scale_hi), k23 = pack2(8388608.0f, 8388608.0f), m1 = pack2(-1.0f, -1.0f), k20 = pack2(1048576.0f, 1048576.0f);
  unsigned long long blk = 0ull, wsum = 0ull;
  {
    // four blocks per trip: their loads and limb arithmetic are independent, and the eight warp reductions that follow
    // are issued back to back instead of one dependent pair per block
    constexpr int kU = 4;
    for (int v0 = v_begin, k = 0; v0 < v_end; v0 += 32 * kU, k += kU) {
      uint32_t acc_hi[kU], acc_lo[kU];
#pragma unroll
      for (int t = 0; t < kU; ++t) {
        const int v = v0 + 32 * t + lane;
        acc_hi[t] = 0u; acc_lo[t] = 0u;
        if (v < v_end) {
          const float4 a4 = row4[v];
          const f32x2 pr[2] = {pack2(a4.x, a4.y), pack2(a4.z, a4.w)};
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const f32x2 t1f = fma2_rd(pr[j], sh2, k23);         // 2^23 + H,  H = floor(W / 2^20),  W = r * scale < 2^40
            const f32x2 nh = fma2(t1f, m1, k23);                // -H
            const f32x2 fr = fma2(pr[j], sh2, nh);              // W / 2^20 - H  in [0, 1)  (exact)
            const f32x2 t2f = fma2_rd(fr, k20, k23);            // 2^23 + floor(W - 2^20 H)
            float h0, h1, l0, l1;
            unpack2(t1f, h0, h1);
            unpack2(t2f, l0, l1);
            acc_hi[t] += (__float_as_uint(h0) - 0x4B000000u) + (__float_as_uint(h1) - 0x4B000000u);
            acc_lo[t] += (__float_as_uint(l0) - 0x4B000000u) + (__float_as_uint(l1) - 0x4B000000u);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < kU; ++t) {
        const uint32_t hs = __reduce_add_sync(0xffffffffu, acc_hi[t]), ls = __reduce_add_sync(0xffffffffu, acc_lo[t]);
        const unsigned long long bs = (static_cast<unsigned long long>(hs) << 20) + ls;
        if (lane == k + t) blk = bs;                            // (at most 29 blocks per warp: the row fits shared memory)
        wsum += bs;
      }
    }
  }
  if (lane == 0) sh.wsum[warp] = wsum;
  __syncthreads();
  if (prof != nullptr && threadIdx.x == 0) prof[4] = clock64();   // pass 2 done
  const unsigned long long mine_w = lane < kRowWarps ? sh.wsum[lane] : 0ull;
  unsigned long long incl = warp_scan_incl(mine_w, lane);
  unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long cta_base = 0ull;
  if (csize > 1) {                                              // totals of the two halves, in vocabulary order
    cg::cluster_group cl = cg::this_cluster();
    if (tid == 0) {
      sh.xtot[crank] = total;
      cl.map_shared_rank(&sh, crank ^ 1)->xtot[crank] = total;
    }
    cl.sync();
    cta_base = crank == 1 ? sh.xtot[0] : 0ull;
    total = sh.xtot[0] + sh.xtot[1];
    incl += cta_base;
  }
  if (total == 0ull) {
    if (tid == 0 && crank == 0) atomicOr(err_flag, kErrEmptyRow);
    return -2;
  }
  const unsigned long long target = scale_target(total, u_to_int(u_final));
  if (target < cta_base) return -1;                             // (the target lies in the other CTA's half)
  const unsigned ball = __ballot_sync(0xffffffffu, incl > target);
  const int owner = __ffs(ball) - 1;                            // first warp whose range crosses the target
  if (warp != owner) return -1;
  const unsigned long long base = __shfl_sync(0xffffffffu, incl - mine_w, owner);
  // the owning warp: block that crosses the target, then the element inside it (vocabulary order)
  const unsigned long long bincl = warp_scan_incl(blk, lane) + base;
  const unsigned bb = __ballot_sync(0xffffffffu, bincl > target);
  const int kb = __ffs(bb) - 1;
  const unsigned long long run = __shfl_sync(0xffffffffu, bincl - blk, kb);
  {
    const int v = v_begin + kb * 32 + lane;
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned long long wv[4], vs = 0ull;
    if (v < v_end) { const float4 a4 = row4[v]; w[0] = a4.x; w[1] = a4.y; w[2] = a4.z; w[3] = a4.w; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { wv[j] = weight_of(w[j], e); vs += wv[j]; }
    const unsigned long long inc2 = warp_scan_incl(vs, lane) + run;
    const unsigned b2 = __ballot_sync(0xffffffffu, inc2 > target);
    if (b2 == 0u || lane != __ffs(b2) - 1) return -1;
    unsigned long long c = inc2 - vs;
    long long found = -1;
    float psel = 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c += wv[j]; if (found < 0 && c > target) { found = idx_off + v * 4 + j; psel = w[j]; } }
    float guard_val = psel;
    if (use_q) guard_val = __fdiv_rn(psel, ldexpf(__ull2float_rn(total), e - kScaleBits) + 1e-6f);   // sample(max_fn(p - q)) sees the normalised value
    return guard_val < kProbGuard ? argmax : found;
  }
  return -1;
}

}  // namespace sd
