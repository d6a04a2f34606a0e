// Internal launch interfaces between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sd {

// Optional compact form of top-k filtered rows: row r of a norm call owns compact row r * row_stride:
//   cnt[cr] = number of kept entries (or -1: not available, use the dense row), idx/val[cr * cap + j] its entries.
struct Compact {
  int* cnt; int* idx; float* val;
  int cap; long long row_stride;
};

struct VerifyParams {
  const float* p; long long p_req_stride, p_row_stride;     // target probs  (B, gamma+1, V)
  const float* q; long long q_req_stride, q_row_stride;     // draft probs   (B, gamma,   V); null => plain sample of p row 0
  const long long* draft; long long draft_stride;           // drafted ids   (B, gamma)
  const float* u_acc; long long u_acc_stride;               // (B, gamma)
  const float* u_final;                                     // (B,)
  int B, gamma; long long V; int strict;
  int* n_accepted; long long* next_tok; float* ratios;      // outputs (ratios nullable, (B, gamma))
  int* tie_count; int* err_flag;
  unsigned long long* stats;                                // optional: [0] += accepted tokens, [1] += requests verified
  // optional fused token append / length update ("rollback" of the static caches is this counter write)
  long long* tokens; long long tokens_stride; int* seq_len; const int* active;
  // optional compact lists of the p and q rows (request b, row i -> compact row b * req_stride + i * row_stride)
  Compact pc, qc; long long pc_req_stride, qc_req_stride;
  // filled by the launcher
  int cluster, slice_elems, use_tma;
  long long* prof;                                          // debug: clock64 timestamps, 8 per request (verify_row_kernel), nullable
  int q_slots;                                              // verify_row_kernel: 16 KB slots of the q ring behind the staged p row (0: per-thread loads)
};
struct NormParams {
  const void* logits; long long ld_in; long long V;
  float temperature; int top_k; float top_p;
  float* probs; long long ld_out;          // nullable: no dense write (sample only)
  const float* u; long long* tok_out;      // nullable pair: per-row uniform -> sampled token
  int* err_flag;
  Compact cmp;                             // cmp.cnt == nullptr: disabled
  // filled by the launcher
  int cluster, slice_elems, slice_smem_bytes, use_tma, vec_out;
  int force_general;                       // test knob: skip the fast top-k path
  int no_pipeline;                         // use the one-cluster-per-row kernel even where the persistent one applies
  long long* prof;                         // debug: per-CTA clock64 timestamps (16 slots per CTA), nullable
  int rows;                                // number of logits rows
  unsigned int* sched;                     // pipelined kernel: {next row ticket, finished clusters}; zero between launches
  int pipe_groups;                         // persistent kernel: compute groups (= slice buffers) per CTA
  int pipe_buffers, pipe_cap, pipe_clusters;             // ... merged-candidate capacity, clusters in the persistent grid
  // ring kernel (norm_ring.cu): one persistent CTA per SM, whole rows streamed through a ring of 16 KB chunks
  int no_ring;                             // caller asked for the older kernels (SD_NORM_NO_RING)
  int ring_mode, ring_slots, ring_smem_bytes, ring_ctas, ring_shared_off, ring_early, ring_trigger, ring_long;
  int ring_row_elems, ring_row_smem_bytes; // geometry of the in-kernel general-path fallback (whole row in one CTA)
  const int* row_filter;                   // norm_probs_kernel: process only rows with row_filter[row] != 0
  unsigned int* defer_bitmap;              // workspace + 16: one bit per row the ring kernel could not serve (rows longer than the ring);
  int use_defer_bitmap;                    // norm_probs_kernel: process (and clear) only the flagged rows
  // pipelined kernel, optional: verify request b as soon as its fv_rows rows (b * fv_rows ..) are all normalised
  int fv_rows;                             // 0: disabled
  int* fv_cnt;                             // (B,) finished-row counters, zero between launches
  VerifyParams fv;
};
cudaError_t launch_norm(const NormParams& p, int dtype, int rows, cudaStream_t st);
// kernel 1 + kernel 2 in one launch where the pipelined kernel applies, else two launches
cudaError_t launch_norm_verify(const NormParams& p, int dtype, int rows, cudaStream_t st);
void set_norm_tuning(int cluster, int threads);
void set_norm_prof(long long* ptr);
long long* get_norm_prof();
void set_pdl(int enable);
int pdl_enabled();
bool plan_pipe(NormParams& p, int dtype, int rows, int tune_cluster);
constexpr int kDeferBitmapRows = 65536;  // rows the workspace bitmap covers (SD_NORM_WORKSPACE_BYTES = 16 + kDeferBitmapRows / 8)
bool plan_ring(NormParams& p, int dtype, int rows);
cudaError_t launch_norm_ring(const NormParams& p, int dtype, cudaStream_t st);
int device_sm_count();                    // multiprocessors of the current device (cached per device)
int device_max_smem_optin();              // largest opt-in dynamic shared memory per block of the current device
cudaError_t launch_norm_pipe(const NormParams& p, int dtype, int rows, cudaStream_t st);

cudaError_t launch_verify(const VerifyParams& p, cudaStream_t st);
cudaError_t launch_verify_bild(const VerifyParams& v, const int* n_check, float fallback_thres, float rollback_thres,
                               const int* limit, int* n_drafted, long long eos, cudaStream_t st);
cudaError_t launch_verify_multi(const VerifyParams& v, long long p_draft_stride, long long q_draft_stride,
                                long long draft_draft_stride, int width, int* choice, cudaStream_t st);
void set_verify_tuning(int cluster);

cudaError_t launch_max_fn(const float* x, long long rows, long long V, long long ld, float* out, long long ld_out,
                          cudaStream_t st);
cudaError_t launch_kv_append(const void* k_new, const void* v_new, long long sb, long long sh, long long sq,
                             void* k_cache, void* v_cache, const int* pos, int B, int H, int q, int D, int S,
                             int elem_size, cudaStream_t st);
cudaError_t launch_kv_select(void* k_cache, void* v_cache, int B, int W, int H, int S, int D, int elem_size, int max_count,
                             const int* choice, const int* start, int start_stride, const int* count, const int* active,
                             int active_stride, cudaStream_t st);
cudaError_t launch_kv_select_layers(void* const* k_caches, void* const* v_caches, int n_layers, int B, int W, int H, int S, int D,
                                    int elem_size, int max_count, const int* choice, const int* start, int start_stride,
                                    const int* count, const int* active, int active_stride, cudaStream_t st);
cudaError_t launch_multi_commit(long long* tokens, long long tokens_stride, int* seq_len, int B, int W, const int* choice,
                                const int* n_acc, const long long* next_tok, const int* active, int S, cudaStream_t st);
cudaError_t launch_build_step(long long* tokens, long long tokens_stride, const int* seq_len, int offset, int q,
                              const long long* prev_tok, int B, int S, long long* input_ids, long long* position_ids,
                              int* write_pos, unsigned char* mask, cudaStream_t st);

}  // namespace sd
