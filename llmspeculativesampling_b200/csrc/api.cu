// C ABI of libspecdec_b200.so — see include/specdec_b200.h for the contract of every entry point.
// All pointers are device pointers owned by the caller; nothing is allocated, freed or synchronised
// here, every call is asynchronous on the caller's stream and CUDA-graph capturable.
#include "specdec_internal.h"
#include "../../include/specdec_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace {
thread_local char g_err[256] = "";
int fail(int code, const char* what, cudaError_t ce = cudaSuccess) {
  if (ce != cudaSuccess)
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(ce));
  else
    snprintf(g_err, sizeof(g_err), "%s", what);
  return code;
}
int done(const char* what, cudaError_t ce) {
  if (ce == cudaSuccess) return SD_OK;
  return fail(static_cast<int>(ce), what, ce);
}
}  // namespace

extern "C" {

int sd_version(void) { return SD_VERSION; }
const char* sd_last_error(void) { return g_err; }

void sd_set_pdl(int enable) { sd::set_pdl(enable); }

void sd_set_tuning(int norm_cluster, int norm_threads, int verify_cluster) {
  sd::set_norm_tuning(norm_cluster, norm_threads);
  sd::set_verify_tuning(verify_cluster);
}

void sd_debug_set_prof(int64_t* device_buf) { sd::set_norm_prof(reinterpret_cast<long long*>(device_buf)); }

static sd::Compact to_compact(const sd_compact_t* c) {
  sd::Compact o = {};
  if (c != nullptr && c->cnt != nullptr && c->idx != nullptr && c->val != nullptr && c->cap > 0) {
    o.cnt = c->cnt; o.idx = c->idx; o.val = c->val; o.cap = c->cap; o.row_stride = c->row_stride;
  }
  return o;
}

static int fill_norm(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld_in, float temperature,
                     int top_k, float top_p, float* probs, int64_t ld_out, const float* u, int64_t* tok_out,
                     const sd_compact_t* compact, int* err_flag, int flags, void* workspace, sd::NormParams& p) {
  (void)dtype;
  if (logits == nullptr || err_flag == nullptr || rows < 0 || V <= 0 || ld_in < V) return fail(SD_EINVAL, "sd_norm: bad logits/shape");
  if (!(temperature > 0.f) || std::isinf(temperature)) return fail(SD_EINVAL, "sd_norm: temperature must be finite and > 0");
  if (probs != nullptr && ld_out < V) return fail(SD_EINVAL, "sd_norm: ld_out < V");
  if ((u == nullptr) != (tok_out == nullptr)) return fail(SD_EINVAL, "sd_norm: u and tok_out go together");
  if (probs == nullptr && u == nullptr) return fail(SD_EINVAL, "sd_norm: nothing to produce");
  if (V >= (1LL << 24) || rows >= (1LL << 28)) return fail(SD_EINVAL, "sd_norm: shape too large");
  p = {};
  p.logits = logits; p.ld_in = ld_in; p.V = V;
  p.temperature = temperature; p.top_k = top_k < 0 ? 0 : top_k; p.top_p = top_p;
  p.probs = probs; p.ld_out = ld_out;
  p.u = u; p.tok_out = reinterpret_cast<long long*>(tok_out);
  p.err_flag = err_flag;
  p.cmp = to_compact(compact);
  p.force_general = (flags & SD_NORM_FORCE_GENERAL) ? 1 : 0;
  p.no_pipeline = ((flags & SD_NORM_NO_PIPELINE) || workspace == nullptr) ? 1 : 0;
  p.no_ring = (flags & SD_NORM_NO_RING) ? 1 : 0;
  p.sched = static_cast<unsigned int*>(workspace);
  p.defer_bitmap = workspace != nullptr ? static_cast<unsigned int*>(workspace) + 4 : nullptr;   // (SD_NORM_WORKSPACE_BYTES = 16 + bitmap)
  p.use_defer_bitmap = 0;
  return SD_OK;
}

static int norm_common(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld_in, float temperature,
                       int top_k, float top_p, float* probs, int64_t ld_out, const float* u, int64_t* tok_out,
                       const sd_compact_t* compact, int* err_flag, int flags, void* workspace, void* stream) {
  if (rows == 0) return SD_OK;
  sd::NormParams p;
  const int rc = fill_norm(logits, dtype, rows, V, ld_in, temperature, top_k, top_p, probs, ld_out, u, tok_out, compact,
                           err_flag, flags, workspace, p);
  if (rc != SD_OK) return rc;
  return done("sd_norm launch", sd::launch_norm(p, dtype, static_cast<int>(rows), static_cast<cudaStream_t>(stream)));
}

int sd_norm_probs(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld_in, float temperature, int top_k,
                  float top_p, float* probs, int64_t ld_out, const sd_compact_t* compact, int* err_flag, int flags,
                  void* workspace, void* stream) {
  if (probs == nullptr) return fail(SD_EINVAL, "sd_norm_probs: probs is null");
  return norm_common(logits, dtype, rows, V, ld_in, temperature, top_k, top_p, probs, ld_out, nullptr, nullptr,
                     compact, err_flag, flags, workspace, stream);
}

int sd_norm_sample(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld_in, float temperature, int top_k,
                   float top_p, float* probs, int64_t ld_out, const float* u, int64_t* tok_out,
                   const sd_compact_t* compact, int* err_flag, int flags, void* workspace, void* stream) {
  if (u == nullptr || tok_out == nullptr) return fail(SD_EINVAL, "sd_norm_sample: u/tok_out is null");
  return norm_common(logits, dtype, rows, V, ld_in, temperature, top_k, top_p, probs, ld_out, u, tok_out, compact,
                     err_flag, flags, workspace, stream);
}

int sd_sample(const float* probs, int64_t rows, int64_t V, int64_t ld, const float* u, int64_t* tok_out, int* err_flag,
              void* stream) {
  if (rows == 0) return SD_OK;
  if (probs == nullptr || u == nullptr || tok_out == nullptr || err_flag == nullptr || V <= 0 || ld < V)
    return fail(SD_EINVAL, "sd_sample: bad argument");
  sd::VerifyParams p = {};
  p.p = probs; p.p_req_stride = ld; p.p_row_stride = 0;
  p.u_final = u; p.B = static_cast<int>(rows); p.gamma = 0; p.V = V;
  p.next_tok = reinterpret_cast<long long*>(tok_out); p.err_flag = err_flag;
  return done("sd_sample launch", sd::launch_verify(p, static_cast<cudaStream_t>(stream)));
}

static int fill_verify(const sd_verify_args_t& a, int* err_flag, sd::VerifyParams& p) {
  if (!a.p_probs || !a.q_probs || !a.draft_tok || !a.u_acc || !a.u_final || !a.n_accepted || !a.next_tok || !err_flag)
    return fail(SD_EINVAL, "sd_verify: null argument");
  if (a.B < 0 || a.gamma < 1 || a.gamma > 32 || a.V <= 0 || a.V >= (1LL << 24)) return fail(SD_EINVAL, "sd_verify: bad shape (1 <= gamma <= 32)");
  if ((a.tokens == nullptr) != (a.seq_len == nullptr)) return fail(SD_EINVAL, "sd_verify: tokens and seq_len go together");
  p = {};
  p.p = a.p_probs; p.p_req_stride = a.p_req_stride; p.p_row_stride = a.p_row_stride;
  p.q = a.q_probs; p.q_req_stride = a.q_req_stride; p.q_row_stride = a.q_row_stride;
  p.draft = reinterpret_cast<const long long*>(a.draft_tok); p.draft_stride = a.draft_stride;
  p.u_acc = a.u_acc; p.u_acc_stride = a.u_acc_stride; p.u_final = a.u_final;
  p.B = a.B; p.gamma = a.gamma; p.V = a.V; p.strict = a.strict;
  p.n_accepted = a.n_accepted; p.next_tok = reinterpret_cast<long long*>(a.next_tok); p.ratios = a.ratios;
  p.tie_count = a.tie_count; p.err_flag = err_flag;
  p.tokens = reinterpret_cast<long long*>(a.tokens); p.tokens_stride = a.tokens_stride; p.seq_len = a.seq_len; p.active = a.active;
  p.stats = reinterpret_cast<unsigned long long*>(a.stats);
  p.pc = to_compact(a.p_compact); p.qc = to_compact(a.q_compact);
  p.pc_req_stride = a.p_cmp_req_stride * p.pc.row_stride; p.qc_req_stride = a.q_cmp_req_stride * p.qc.row_stride;
  return SD_OK;
}

int sd_verify(const float* p_probs, int64_t p_req_stride, int64_t p_row_stride, const float* q_probs,
              int64_t q_req_stride, int64_t q_row_stride, const int64_t* draft_tok, int64_t draft_stride,
              const float* u_acc, int64_t u_acc_stride, const float* u_final, int B, int gamma, int64_t V, int strict,
              int32_t* n_accepted, int64_t* next_tok, float* ratios, int32_t* tie_count, int64_t* tokens,
              int64_t tokens_stride, int32_t* seq_len, const int32_t* active, const sd_compact_t* p_compact,
              int64_t p_cmp_req_stride, const sd_compact_t* q_compact, int64_t q_cmp_req_stride, uint64_t* stats,
              int* err_flag, void* stream) {
  if (B == 0) return SD_OK;
  const sd_verify_args_t a = {p_probs, p_req_stride, p_row_stride, q_probs, q_req_stride, q_row_stride, draft_tok, draft_stride,
                              u_acc, u_acc_stride, u_final, B, gamma, V, strict, n_accepted, next_tok, ratios, tie_count,
                              tokens, tokens_stride, seq_len, active, p_compact, p_cmp_req_stride, q_compact,
                              q_cmp_req_stride, stats};
  sd::VerifyParams p;
  const int rc = fill_verify(a, err_flag, p);
  if (rc != SD_OK) return rc;
  return done("sd_verify launch", sd::launch_verify(p, static_cast<cudaStream_t>(stream)));
}

int sd_norm_sample_verify(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld_in, float temperature,
                          int top_k, float top_p, float* probs, int64_t ld_out, const float* u, int64_t* tok_out,
                          const sd_compact_t* compact, const sd_verify_args_t* verify, int rows_per_request,
                          int32_t* request_counters, int* err_flag, int flags, void* workspace, void* stream) {
  if (verify == nullptr || request_counters == nullptr || rows_per_request < 1)
    return fail(SD_EINVAL, "sd_norm_sample_verify: verify / request_counters / rows_per_request");
  if (rows != static_cast<int64_t>(verify->B) * rows_per_request) return fail(SD_EINVAL, "sd_norm_sample_verify: rows != B * rows_per_request");
  if (rows == 0) return SD_OK;
  sd::NormParams p;
  const int rc = fill_norm(logits, dtype, rows, V, ld_in, temperature, top_k, top_p, probs, ld_out, u, tok_out, compact,
                           err_flag, flags, workspace, p);
  if (rc != SD_OK) return rc;
  const int rc2 = fill_verify(*verify, err_flag, p.fv);
  if (rc2 != SD_OK) return rc2;
  p.fv_rows = rows_per_request;
  p.fv_cnt = request_counters;
  return done("sd_norm_sample_verify launch", sd::launch_norm_verify(p, dtype, static_cast<int>(rows), static_cast<cudaStream_t>(stream)));
}

int sd_verify_multi(const float* p_probs, int64_t p_req_stride, int64_t p_draft_stride, int64_t p_row_stride,
                    const float* q_probs, int64_t q_req_stride, int64_t q_draft_stride, int64_t q_row_stride,
                    const int64_t* draft_tok, int64_t draft_req_stride, int64_t draft_draft_stride, const float* u_acc,
                    int64_t u_acc_stride, const float* u_final, int B, int width, int gamma, int64_t V, int32_t* choice,
                    int32_t* n_accepted, int64_t* next_tok, float* ratios, int* err_flag, void* stream) {
  if (B == 0) return SD_OK;
  if (width < 1 || u_acc_stride < static_cast<int64_t>(width) * gamma) return fail(SD_EINVAL, "sd_verify_multi: width / u_acc_stride");
  const sd_verify_args_t a = {p_probs, p_req_stride, p_row_stride, q_probs, q_req_stride, q_row_stride, draft_tok,
                              draft_req_stride, u_acc, u_acc_stride, u_final, B, gamma, V, 1, n_accepted, next_tok, ratios,
                              nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr};
  sd::VerifyParams p;
  const int rc = fill_verify(a, err_flag, p);
  if (rc != SD_OK) return rc;
  return done("sd_verify_multi launch", sd::launch_verify_multi(p, p_draft_stride, q_draft_stride, draft_draft_stride, width,
                                                                choice, static_cast<cudaStream_t>(stream)));
}

int sd_verify_bild(const float* p_probs, int64_t p_req_stride, int64_t p_row_stride, const float* q_probs,
                   int64_t q_req_stride, int64_t q_row_stride, const int64_t* draft_tok, int64_t draft_stride,
                   const int32_t* n_check, int max_check, float fallback_thres, float rollback_thres, const float* u_final,
                   int B, int64_t V, int32_t* n_accepted, int64_t* next_tok, float* nll, int32_t* n_drafted, int64_t* tokens,
                   int64_t tokens_stride, int32_t* seq_len, const int32_t* limit, const int32_t* active,
                   const sd_compact_t* q_compact, int64_t q_cmp_req_stride, int64_t eos_token_id, int* err_flag, void* stream) {
  if (B == 0) return SD_OK;
  if (!p_probs || !draft_tok || !u_final || !n_accepted || !next_tok || !err_flag) return fail(SD_EINVAL, "sd_verify_bild: null argument");
  if (B < 0 || max_check < 1 || max_check > 32 || V <= 0 || V >= (1LL << 24)) return fail(SD_EINVAL, "sd_verify_bild: bad shape (1 <= max_check <= 32)");
  if ((tokens == nullptr) != (seq_len == nullptr)) return fail(SD_EINVAL, "sd_verify_bild: tokens and seq_len go together");
  sd::VerifyParams p = {};
  p.p = p_probs; p.p_req_stride = p_req_stride; p.p_row_stride = p_row_stride;
  p.q = q_probs; p.q_req_stride = q_req_stride; p.q_row_stride = q_row_stride;
  p.draft = reinterpret_cast<const long long*>(draft_tok); p.draft_stride = draft_stride;
  p.u_final = u_final; p.B = B; p.gamma = max_check; p.V = V;
  p.n_accepted = n_accepted; p.next_tok = reinterpret_cast<long long*>(next_tok); p.ratios = nll; p.err_flag = err_flag;
  p.tokens = reinterpret_cast<long long*>(tokens); p.tokens_stride = tokens_stride; p.seq_len = seq_len; p.active = active;
  p.qc = to_compact(q_compact); p.qc_req_stride = q_cmp_req_stride;
  return done("sd_verify_bild launch", sd::launch_verify_bild(p, n_check, fallback_thres, rollback_thres, limit, n_drafted,
                                                              static_cast<long long>(eos_token_id), static_cast<cudaStream_t>(stream)));
}

int sd_max_fn(const float* x, int64_t rows, int64_t V, int64_t ld, float* out, int64_t ld_out, void* stream) {
  if (rows == 0) return SD_OK;
  if (!x || !out || V <= 0 || ld < V || ld_out < V) return fail(SD_EINVAL, "sd_max_fn: bad argument");
  return done("sd_max_fn launch", sd::launch_max_fn(x, rows, V, ld, out, ld_out, static_cast<cudaStream_t>(stream)));
}

int sd_kv_append(const void* k_new, const void* v_new, int64_t stride_b, int64_t stride_h, int64_t stride_q,
                 void* k_cache, void* v_cache, const int32_t* write_pos, int B, int H, int q, int D, int S,
                 int elem_size, void* stream) {
  if (!k_new || !v_new || !k_cache || !v_cache || !write_pos) return fail(SD_EINVAL, "sd_kv_append: null argument");
  if (elem_size != 2 && elem_size != 4) return fail(SD_EINVAL, "sd_kv_append: elem_size must be 2 or 4");
  if ((D * elem_size) % 16 != 0 || (stride_b * elem_size) % 16 || (stride_h * elem_size) % 16 || (stride_q * elem_size) % 16)
    return fail(SD_EINVAL, "sd_kv_append: rows must be 16-byte aligned");
  return done("sd_kv_append launch",
              sd::launch_kv_append(k_new, v_new, stride_b, stride_h, stride_q, k_cache, v_cache, write_pos, B, H, q, D,
                                   S, elem_size, static_cast<cudaStream_t>(stream)));
}

int sd_kv_select(void* k_cache, void* v_cache, int B, int W, int H, int S, int D, int elem_size, int max_count,
                 const int32_t* choice, const int32_t* start, int start_stride, const int32_t* count, const int32_t* active,
                 int active_stride, void* stream) {
  if (!k_cache || !v_cache || !choice || !start || !count) return fail(SD_EINVAL, "sd_kv_select: null argument");
  return done("sd_kv_select launch", sd::launch_kv_select(k_cache, v_cache, B, W, H, S, D, elem_size, max_count, choice, start,
                                                          start_stride, count, active, active_stride, static_cast<cudaStream_t>(stream)));
}

int sd_kv_select_layers(void* const* k_caches, void* const* v_caches, int n_layers, int B, int W, int H, int S, int D,
                        int elem_size, int max_count, const int32_t* choice, const int32_t* start, int start_stride,
                        const int32_t* count, const int32_t* active, int active_stride, void* stream) {
  if (!k_caches || !v_caches || !choice || !start || !count) return fail(SD_EINVAL, "sd_kv_select_layers: null argument");
  return done("sd_kv_select_layers launch",
              sd::launch_kv_select_layers(k_caches, v_caches, n_layers, B, W, H, S, D, elem_size, max_count, choice, start,
                                          start_stride, count, active, active_stride, static_cast<cudaStream_t>(stream)));
}

int sd_multi_commit(int64_t* tokens, int64_t tokens_stride, int32_t* seq_len, int B, int W, const int32_t* choice,
                    const int32_t* n_acc, const int64_t* next_tok, const int32_t* active, int S, void* stream) {
  if (!tokens || !seq_len || !choice || !n_acc || !next_tok) return fail(SD_EINVAL, "sd_multi_commit: null argument");
  return done("sd_multi_commit launch", sd::launch_multi_commit(reinterpret_cast<long long*>(tokens), tokens_stride, seq_len, B, W,
                                                                choice, n_acc, reinterpret_cast<const long long*>(next_tok), active, S,
                                                                static_cast<cudaStream_t>(stream)));
}

int sd_build_step(int64_t* tokens, int64_t tokens_stride, const int32_t* seq_len, int offset, int q,
                  const int64_t* prev_tok, int B, int S, int64_t* input_ids, int64_t* position_ids, int32_t* write_pos,
                  uint8_t* mask, void* stream) {
  if (!tokens || !seq_len || !input_ids || !position_ids || !write_pos) return fail(SD_EINVAL, "sd_build_step: null argument");
  return done("sd_build_step launch",
              sd::launch_build_step(reinterpret_cast<long long*>(tokens), tokens_stride, seq_len, offset, q,
                                    reinterpret_cast<const long long*>(prev_tok), B, S,
                                    reinterpret_cast<long long*>(input_ids), reinterpret_cast<long long*>(position_ids),
                                    write_pos, mask, static_cast<cudaStream_t>(stream)));
}

}  // extern "C"
