/* specdec_b200.h — C ABI of libspecdec_b200.so (sm_100a).
 *
 * Drop-in boundary for the speculative-decoding draft-and-verify hot path of
 * ZongyueQin/LLMSpeculativeSampling.  The reference has NO native/FFI boundary of its own (it is pure
 * Python; its boundary is the `sampling` package), so every entry point below cites the
 * reference Python code it replaces, as /root/reference/<file>:<lines>.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (torch tensors in the Python host code);
 *    the library allocates nothing and frees nothing.  Process-wide state it does keep: the sd_set_tuning /
 *    sd_set_pdl / sd_debug_set_prof knobs (plain ints / a pointer, meant to be set once before use, not synchronised),
 *    a mutex-protected cache of launch plans per (dtype, V, top_k), and per-device "attribute set" flags and device
 *    properties (SM count, shared-memory limit) read from the driver on first use;
 *  - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it, never synchronise,
 *    and are CUDA-graph capturable;
 *  - return value: SD_OK (0), SD_EINVAL (-1, argument error) or a positive cudaError_t;
 *    sd_last_error() gives the text (thread local);
 *  - numerical faults do not fail the call: kernels OR bits into the caller's device word `err_flag`
 *    (SD_ERR_*), which the host checks once per public-API call and turns into the reference's
 *    exceptions (RuntimeError('norm logits error') utils.py:207, RuntimeError('prob error')
 *    utils.py:224, RuntimeError('s') speculative_sampling.py:2046);
 *  - strides / leading dimensions are in ELEMENTS;
 *  - inverse-CDF rule for every sampled token (replaces torch.multinomial, utils.py:221):
 *        e = frexp exponent of the row maximum, w_i = floor(p_i * 2^(40-e)), m = floor(u * 2^24),
 *        t = (sum_i w_i * m) >> 24, token = first i with w_0 + .. + w_i > t,
 *        then the reference's guard: if p[token] < 1e-9 take argmax(p) (utils.py:228-230).
 */
#ifndef SPECDEC_B200_H_
#define SPECDEC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_VERSION 210
#define SD_OK 0
#define SD_EINVAL (-1)

/* logits dtypes */
#define SD_F32 0
#define SD_BF16 1
#define SD_F16 2

/* bits OR-ed into *err_flag by the kernels */
#define SD_ERR_NORM_LOGITS 1 /* NaN / +inf logit or non-finite probability   -> 'norm logits error' */
#define SD_ERR_PROB 2        /* no positive weight to sample from, or p < 0   -> 'prob error'        */
#define SD_ERR_ZERO_Q 4      /* q[drafted token] == 0 (ZeroDivisionError)     -> 's'                 */
#define SD_ERR_BAD_TOKEN 8   /* drafted token id outside [0, V)                                       */

/* Optional compact form of top-k filtered probability rows (all pointers device memory, caller-owned).
 * Logical row r owns compact row cr = r * row_stride:  cnt[cr] = number of kept (non-zero) entries, or -1 when the
 * row has no compact form (it was served by the dense / general path or has more than `cap` entries);
 * idx[cr * cap + j], val[cr * cap + j], j < cnt[cr], are the vocabulary indices and probabilities (any order).
 * Kernel 1 writes it next to the dense row; kernel 2 can then verify from a few hundred bytes per request instead of
 * two dense vocabulary rows.  The reference has no counterpart: it always materialises dense rows
 * (sampling/kvcache_model.py:246). */
typedef struct sd_compact {
  int32_t* cnt;
  int32_t* idx;
  float* val;
  int32_t cap;
  int64_t row_stride;
} sd_compact_t;

int sd_version(void);
const char* sd_last_error(void);

/* Benchmark / test knobs: cluster size (1,2,4,8; 0 = heuristic) and threads per CTA (256/512/1024;
 * 0 = heuristic) of the norm kernel, cluster size of the verify kernel. */
void sd_set_tuning(int norm_cluster, int norm_threads, int verify_cluster);

/* Programmatic dependent launch (on by default): when enabled, the pipelined norm kernel and the sparse verify
 * kernel are launched with cudaLaunchAttributeProgrammaticStreamSerialization, so their CTA scheduling and
 * shared-memory / mbarrier set-up overlap the tail of the previous kernel in the stream; both execute
 * griddepcontrol.wait before their first global-memory access, so results are unchanged. */
void sd_set_pdl(int enable);

/* Debug: when non-NULL, every CTA of the norm kernel writes clock64() phase timestamps into
 * device_buf[cta * 16 + slot] (tools/microbench.py --prof).  NULL switches it off. */
void sd_debug_set_prof(int64_t* device_buf);

/* Kernel 1 — fused  logits / T -> top-k -> top-p -> softmax  for `rows` rows of V logits.
 * Replaces sampling/utils.py:152-179 (top_k_top_p_filter) + :182-210 (norm_logits), called once per row
 * from sampling/kvcache_model.py:166-168 and :235-236.
 *   logits  (rows, V) of dtype SD_F32/SD_BF16/SD_F16, row stride ld_in; read once (TMA bulk copy when
 *           base and row stride are 16-byte aligned, plain loads otherwise); not modified
 *   top_k   <= 0 disables; ties with the k-th value are kept (utils.py:169)
 *   top_p   <= 0 disables; keeps the smallest prefix of the descending order whose mass exceeds top_p,
 *           crossing entry included; equal values are ordered by ascending index (utils.py:171-178)
 *   probs   (rows, V) fp32, row stride ld_out: exp(log_softmax(filtered)) (utils.py:199), 0 outside
 * Sets SD_ERR_NORM_LOGITS where the reference raises 'norm logits error' (utils.py:203-207). */
int sd_norm_probs(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld_in, float temperature, int top_k,
                  float top_p, float* probs, int64_t ld_out, const sd_compact_t* compact, int* err_flag, int flags,
                  void* workspace, void* stream);

/* `workspace` of the two norm entry points: SD_NORM_WORKSPACE_BYTES of device memory, zeroed ONCE by the caller.  The
 * persistent kernels hand rows to their CTAs through a ticket counter that lives in its first 16 bytes (SMs run at
 * visibly different speeds under full HBM load; a static split waits for the slowest); the rest is one bit per row
 * (up to 65536 rows per call) in which the ring kernel flags the top-k rows LONGER than its shared-memory ring that its
 * fast selection cannot serve (massive ties) for the follow-up launch of the general path.  Every launch leaves the
 * block zeroed when it ends, so the same block serves any number of launches that are ordered after one another (one
 * stream, one CUDA graph).  Launches that may run CONCURRENTLY must use different blocks.  NULL is always safe: it
 * selects the one-cluster-per-row kernel. */
#define SD_NORM_WORKSPACE_BYTES (16 + 65536 / 8)

/* `flags` of the two norm entry points.  By default, with a `workspace` and 16-byte aligned rows, persistent kernels run:
 *  - the ring kernel (one CTA per SM streams whole rows through a shared-memory ring of 16 KB TMA chunks) for
 *    0 < top_k <= 128 (rows that fit the ring, or of >= 24 chunks) and for the dense default top_k = 0, top_p = 0 (rows up
 *    to 1 MB);
 *  - the cluster pipeline (one cluster per row slice, DSMEM candidate exchange) for the top-k rows in between.
 * Rows their fast selection cannot serve — massive ties — are re-run on the general path inside the same call.  Every other
 * case (top-p only, top_k > 128, unaligned rows, no workspace) uses the one-cluster-per-row kernel. */
#define SD_NORM_DEFAULT 0
#define SD_NORM_NO_PIPELINE 1   /* one-cluster-per-row kernel even where the pipeline applies                 */
#define SD_NORM_NO_RING 4       /* skip the ring kernel (one CTA per SM, whole rows through a shared-memory ring): use the
                                   cluster pipeline / one-cluster-per-row kernels even where a row fits one CTA           */
#define SD_NORM_FORCE_GENERAL 2 /* test hook: sort-free threshold-search path (normally top_k = 0, top_k > 128
                                   or rows with massive ties only); implies SD_NORM_NO_PIPELINE                */

/* Kernel 1b — same, plus one inverse-CDF sample per row from the uniform u[row] (draft step):
 * replaces norm_logits + sample(q) of sampling/kvcache_model.py:280-283 and
 * sampling/autoregressive_sampling.py:41-44.  `probs` may be NULL (token only). */
int sd_norm_sample(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld_in, float temperature, int top_k,
                   float top_p, float* probs, int64_t ld_out, const float* u, int64_t* tok_out,
                   const sd_compact_t* compact, int* err_flag, int flags, void* workspace, void* stream);

/* sample — one inverse-CDF draw per row of non-negative weights.  Replaces sampling/utils.py:213-233. */
int sd_sample(const float* probs, int64_t rows, int64_t V, int64_t ld, const float* u, int64_t* tok_out, int* err_flag,
              void* stream);

/* Kernel 2 — fused verify for B requests.  Replaces sampling/speculative_sampling.py:1966-2027 (accept
 * loop :1975-1990, residual sample(max_fn(p-q)) :2005-2015 with its fall-back to sample(p) :2009-2010,
 * bonus sample :2016-2023, append :2027) and, with strict = 1, speculative_sampling_v2 :2152-2181.
 *   p_probs  target probabilities, request b row i at p_probs + b*p_req_stride + i*p_row_stride, i = 0..gamma
 *   q_probs  draft probabilities, rows i = 0..gamma-1
 *   draft_tok (B, gamma) int64 drafted ids, row stride draft_stride
 *   u_acc    (B, gamma) accept uniforms (row stride u_acc_stride), u_final (B,) resample/bonus uniform
 *   strict   0: accept iff !(u > p/q)  (:1981)      1: accept iff u < min(1, p/q)  (:2156)
 *   n_accepted (B,) int32, next_tok (B,) int64
 *   ratios   optional (B, gamma) fp32 p/q of every drafted token (acc_rate statistic, :1966-1971)
 *   tie_count optional device int32, += number of tested positions with u == threshold exactly
 *   tokens / seq_len  optional fused append: tokens[b, seq_len[b] + n_acc] = next_tok;
 *            seq_len[b] += n_acc + 1  — with static caches this counter write IS the rollback of
 *            sampling/kvcache_model.py:360-431 (approx rollback(n+1) :2000, target :2015/:2023)
 *   active   optional (B,) int32; requests with active[b] == 0 are skipped entirely.
 *   p_compact / q_compact  optional compact lists of the same rows (request b, row i -> logical row
 *            b * p_cmp_req_stride + i, scaled by the struct's row_stride); when both are given one warp per request
 *            works from the lists (sparse path) and only requests whose lists are unavailable read dense rows.
 *   stats    optional device uint64[2]: [0] += accepted tokens, [1] += requests verified (acc_len / call counters of
 *            sampling/speculative_sampling.py:1991, 2062-2073 without a host round trip). */
int sd_verify(const float* p_probs, int64_t p_req_stride, int64_t p_row_stride, const float* q_probs,
              int64_t q_req_stride, int64_t q_row_stride, const int64_t* draft_tok, int64_t draft_stride,
              const float* u_acc, int64_t u_acc_stride, const float* u_final, int B, int gamma, int64_t V, int strict,
              int32_t* n_accepted, int64_t* next_tok, float* ratios, int32_t* tie_count, int64_t* tokens,
              int64_t tokens_stride, int32_t* seq_len, const int32_t* active, const sd_compact_t* p_compact,
              int64_t p_cmp_req_stride, const sd_compact_t* q_compact, int64_t q_cmp_req_stride, uint64_t* stats,
              int* err_flag, void* stream);

/* Kernels 1 + 2 in ONE launch.  The arguments of sd_verify as a struct, then: sd_norm_sample over `rows` rows where
 * request b owns the `rows_per_request` consecutive rows b * rows_per_request .. (rows = verify->B * rows_per_request);
 * as soon as the last of them is normalised the request is verified exactly as sd_verify would (same arithmetic, same
 * outputs) by the thread group that finished that row — the accept / resample step of
 * sampling/speculative_sampling.py:1966-2027 no longer costs a second kernel launch and its drain / ramp.
 * The verify arguments usually point into this call's own outputs (probs, tok_out, compact lists); rows of a request
 * that are not part of this call (e.g. draft rows normalised earlier) must be complete before the launch.
 *   request_counters  (verify->B,) int32 device scratch, zeroed once by the caller, left zeroed by every launch
 * Needs compact lists on both sides and the persistent kernel (workspace given, 0 < top_k <= 128, aligned rows);
 * otherwise the library runs sd_norm_sample followed by sd_verify on the same stream — the result is the same. */
typedef struct sd_verify_args {
  const float* p_probs; int64_t p_req_stride, p_row_stride;
  const float* q_probs; int64_t q_req_stride, q_row_stride;
  const int64_t* draft_tok; int64_t draft_stride;
  const float* u_acc; int64_t u_acc_stride;
  const float* u_final;
  int32_t B, gamma; int64_t V; int32_t strict;
  int32_t* n_accepted; int64_t* next_tok; float* ratios; int32_t* tie_count;
  int64_t* tokens; int64_t tokens_stride; int32_t* seq_len; const int32_t* active;
  const sd_compact_t* p_compact; int64_t p_cmp_req_stride;
  const sd_compact_t* q_compact; int64_t q_cmp_req_stride;
  uint64_t* stats;
} sd_verify_args_t;
int sd_norm_sample_verify(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld_in, float temperature,
                          int top_k, float top_p, float* probs, int64_t ld_out, const float* u, int64_t* tok_out,
                          const sd_compact_t* compact, const sd_verify_args_t* verify, int rows_per_request,
                          int32_t* request_counters, int* err_flag, int flags, void* workspace, void* stream);

/* Kernel 2, multi-draft variant.  Replaces the accept loop and resample of multi_speculative_sampling(strategy='iid'),
 * sampling/speculative_sampling.py:1612-1667: `width` drafts of gamma tokens per request.
 *   p_probs   draft w, row i of request b at p_probs + b*p_req_stride + w*p_draft_stride + i*p_row_stride (i = 0..gamma)
 *   q_probs   likewise, rows i = 0..gamma-1;   draft_tok (B, width, gamma) int64 with the two strides given
 *   u_acc     per request the accept uniforms IN DRAWING ORDER (row stride u_acc_stride >= width*gamma): the reference
 *             draws lazily — draft 0 one per tested token until its first reject, then draft 1, ... (:1616-1634)
 *   accept iff u < min(1, p/q) (a NaN ratio rejects); the first draft with the longest accepted run wins (:1636-1640),
 *   an all-accepted draft ends the scan; then the residual max(0, p_n - q_n) of the winning draft (empty: p_n) or its
 *   bonus row is sampled with u_final as in sd_verify.
 *   choice (B,) int32 winning draft, n_accepted (B,) its accepted run, next_tok (B,), ratios optional (B, width, gamma). */
int sd_verify_multi(const float* p_probs, int64_t p_req_stride, int64_t p_draft_stride, int64_t p_row_stride,
                    const float* q_probs, int64_t q_req_stride, int64_t q_draft_stride, int64_t q_row_stride,
                    const int64_t* draft_tok, int64_t draft_req_stride, int64_t draft_draft_stride, const float* u_acc,
                    int64_t u_acc_stride, const float* u_final, int B, int width, int gamma, int64_t V, int32_t* choice,
                    int32_t* n_accepted, int64_t* next_tok, float* ratios, int* err_flag, void* stream);

/* Kernel 2, BiLD variant.  Replaces the target's check of BiLD_sampling, sampling/speculative_sampling.py:1793-1813:
 * request b has n_check[b] (<= max_check; NULL: max_check) unchecked draft tokens draft_tok[b, i] whose target
 * distributions are rows i of p_probs; the target keeps tokens while -log p[i][token] <= rollback_thres (:1800), then
 * always samples its own token from row n = number of kept tokens (:1812) with u_final (row max_check must exist).
 *   n_accepted (B,) kept tokens, next_tok (B,), nll optional (B, max_check) the tested -log p values.
 * Engine mode (q_probs != NULL, max_check tokens drafted up front in a fixed-shape graph): the number of tokens the
 * reference would have drafted is derived first — it stops at the first token whose distribution is unsure,
 * max q[i] < fallback_thres (:1784) — and written to n_drafted[b]; tokens / seq_len (optional, together) get the fused
 * append tokens[b, seq_len + n] = next_tok, seq_len += n + 1; with limit[b] (optional total-length limit) a request with
 * less room than drafted tokens ends like the reference's loop (:1764): the drafted tokens stay unchecked, no target
 * token, n_accepted = -1 - kept tokens, next_tok = -1.  active optional (B,).
 * q_compact (optional, engine mode): compact lists of the q rows as written by kernel 1 (request b, row i -> logical row
 * b * q_cmp_req_stride + i): max q is then the maximum of the <= cap listed values instead of a scan of the dense row.
 * eos_token_id (engine mode, < 0: none): the reference tests for EOS after every draft token (:1826-1841) — an EOS drafted
 * before the token that triggers the check ends the request with the drafted tokens up to it kept unchecked
 * (n_accepted = -1 - kept tokens, next_tok = -1). */
int sd_verify_bild(const float* p_probs, int64_t p_req_stride, int64_t p_row_stride, const float* q_probs,
                   int64_t q_req_stride, int64_t q_row_stride, const int64_t* draft_tok, int64_t draft_stride,
                   const int32_t* n_check, int max_check, float fallback_thres, float rollback_thres, const float* u_final,
                   int B, int64_t V, int32_t* n_accepted, int64_t* next_tok, float* nll, int32_t* n_drafted, int64_t* tokens,
                   int64_t tokens_stride, int32_t* seq_len, const int32_t* limit, const int32_t* active,
                   const sd_compact_t* q_compact, int64_t q_cmp_req_stride, int64_t eos_token_id, int* err_flag, void* stream);

/* max_fn — out = max(x,0) / (sum(max(x,0)) + 1e-6) per row.  Replaces sampling/utils.py:236-245. */
int sd_max_fn(const float* x, int64_t rows, int64_t V, int64_t ld, float* out, int64_t ld_out, void* stream);

/* Kernel 3a — static KV-cache append at per-request offsets: for every request b, head h, new row j:
 *   cache[b, h, write_pos[b] + j, :] = new[b, h, j, :]   for both K and V  (caches are (B, H, S, D) contiguous,
 *   new tensors have element strides stride_b/h/q and a contiguous last dim).  Replaces the torch.cat growth of
 *   the legacy cache (sampling/models/modeling_llama.py:337-338) and makes KVCacheModel.rollback
 *   (sampling/kvcache_model.py:379-384) a no-op on data: stale rows are overwritten by the next append. */
int sd_kv_append(const void* k_new, const void* v_new, int64_t stride_b, int64_t stride_h, int64_t stride_q,
                 void* k_cache, void* v_cache, const int32_t* write_pos, int B, int H, int q, int D, int S,
                 int elem_size, void* stream);

/* Kernel 3a', multi-draft rollback — replaces rollback(end_pos, choice) of sampling/kvcache_model.py:390-396 (keep draft
 * `choice`; the next forward expands it to all rows again, :180-200) on static caches that hold W rows per request
 * (row b*W + w): positions start[b*start_stride] .. + count[b] - 1 of row choice[b] are copied over the request's other
 * rows (everything before is the shared prefix).  max_count bounds count; requests with active[b*active_stride] == 0 are
 * skipped (active may be NULL). */
int sd_kv_select(void* k_cache, void* v_cache, int B, int W, int H, int S, int D, int elem_size, int max_count,
                 const int32_t* choice, const int32_t* start, int start_stride, const int32_t* count, const int32_t* active,
                 int active_stride, void* stream);

/* The same for EVERY layer of a model in one launch: k_caches / v_caches are DEVICE arrays of n_layers cache pointers (all
 * layers share the geometry).  Replaces the per-layer loop of sampling/kvcache_model.py:390-396. */
int sd_kv_select_layers(void* const* k_caches, void* const* v_caches, int n_layers, int B, int W, int H, int S, int D,
                        int elem_size, int max_count, const int32_t* choice, const int32_t* start, int start_stride,
                        const int32_t* count, const int32_t* active, int active_stride, void* stream);

/* Token append of the multi-draft loop (sampling/speculative_sampling.py:1644, :1677): every row b*W + w of request b
 * becomes prefix + the winning draft's n_acc[b] accepted tokens + next_tok[b]; seq_len of all W rows advances by
 * n_acc[b] + 1.  tokens (B*W, >= S) int64, seq_len (B*W,) int32, active optional (B*W,). */
int sd_multi_commit(int64_t* tokens, int64_t tokens_stride, int32_t* seq_len, int B, int W, const int32_t* choice,
                    const int32_t* n_acc, const int64_t* next_tok, const int32_t* active, int S, void* stream);

/* Kernel 3b — per-step input builder for the graph-captured draft/target steps.  For request b the step
 * consumes the q tokens at positions start..start+q-1, start = seq_len[b] + offset; prev_tok (optional)
 * is stored at the last of them first (token append of sampling/kvcache_model.py:293).  Emits input_ids
 * (B,q), position_ids (B,q), write_pos (B,) and mask (B,1,q,S) uint8: key s visible to query j iff
 * s <= start + j.  Replaces `input_ids[:, cached_len:]` (kvcache_model.py:175,206). */
int sd_build_step(int64_t* tokens, int64_t tokens_stride, const int32_t* seq_len, int offset, int q,
                  const int64_t* prev_tok, int B, int S, int64_t* input_ids, int64_t* position_ids, int32_t* write_pos,
                  uint8_t* mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPECDEC_B200_H_ */
