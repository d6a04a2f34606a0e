// Small kernels around the two fused ones: max_fn (drop-in utility), static KV-cache append at
// per-request offsets, and the per-step input builder (ids / positions / write offsets / mask)
// that lets the gamma-step draft loop run inside one CUDA graph with ragged request lengths.
#include "common.cuh"
#include "specdec_internal.h"

namespace sd {

// ---------------------------------------------------------------------------------------------
// max_fn: out = max(x, 0) / (sum(max(x, 0)) + 1e-6)     (/root/reference/sampling/utils.py:236-245)
__global__ void __launch_bounds__(1024) max_fn_kernel(const float* __restrict__ x, long long V, long long ld,
                                                      float* __restrict__ out, long long ld_out) {
  __shared__ double wsum[32];
  const float* xr = x + blockIdx.x * ld;
  float* orow = out + blockIdx.x * ld_out;
  double acc = 0.0;
  for (long long i = threadIdx.x; i < V; i += blockDim.x) acc += static_cast<double>(fmaxf(xr[i], 0.f));
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) tot += wsum[w];
  const float denom = static_cast<float>(tot) + 1e-6f;
  for (long long i = threadIdx.x; i < V; i += blockDim.x) orow[i] = __fdiv_rn(fmaxf(xr[i], 0.f), denom);
}

cudaError_t launch_max_fn(const float* x, long long rows, long long V, long long ld, float* out, long long ld_out,
                          cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  max_fn_kernel<<<static_cast<unsigned>(rows), 1024, 0, st>>>(x, V, ld, out, ld_out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// KV append: write the q new key/value rows of every (request, head) at cache position
// pos[b] + j of the static (B, H, S, D) caches.  Replaces the torch.cat growth of the legacy tuple
// cache (reference sampling/models/modeling_llama.py:337-338) — rollback is then just a smaller
// pos[b] on the next step (reference kvcache_model.py:381-382 slices and re-materialises).
// One warp per (b, h, j) row, 16-byte copies; D * elem_size must be a multiple of 16.
__global__ void kv_append_kernel(const unsigned char* __restrict__ kn, const unsigned char* __restrict__ vn,
                                 long long sb, long long sh, long long sq,           // strides of the new tensors, bytes
                                 unsigned char* __restrict__ kc, unsigned char* __restrict__ vc,
                                 const int* __restrict__ pos, int B, int H, int q, int S, int row_bytes) {
  const int warps_per_block = blockDim.x >> 5;
  const long long r = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  if (r >= static_cast<long long>(B) * H * q) return;
  const int j = static_cast<int>(r % q);
  const int h = static_cast<int>((r / q) % H);
  const int b = static_cast<int>(r / (static_cast<long long>(q) * H));
  const int dst_pos = pos[b] + j;
  if (dst_pos < 0 || dst_pos >= S) return;
  const long long src = b * sb + h * sh + j * sq;
  const long long dst = ((static_cast<long long>(b) * H + h) * S + dst_pos) * row_bytes;
  for (int o = (threadIdx.x & 31) * 16; o < row_bytes; o += 32 * 16) {
    *reinterpret_cast<uint4*>(kc + dst + o) = *reinterpret_cast<const uint4*>(kn + src + o);
    *reinterpret_cast<uint4*>(vc + dst + o) = *reinterpret_cast<const uint4*>(vn + src + o);
  }
}

cudaError_t launch_kv_append(const void* k_new, const void* v_new, long long sb, long long sh, long long sq,
                             void* k_cache, void* v_cache, const int* pos, int B, int H, int q, int D, int S,
                             int elem_size, cudaStream_t st) {
  const int row_bytes = D * elem_size;
  if (row_bytes % 16 != 0) return cudaErrorInvalidValue;
  const long long rows = static_cast<long long>(B) * H * q;
  if (rows <= 0) return cudaSuccess;
  const int wpb = 8;
  kv_append_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(
      static_cast<const unsigned char*>(k_new), static_cast<const unsigned char*>(v_new), sb * elem_size,
      sh * elem_size, sq * elem_size, static_cast<unsigned char*>(k_cache), static_cast<unsigned char*>(v_cache), pos,
      B, H, q, S, row_bytes);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Multi-draft rollback (reference kvcache_model.py:390-396: rollback(end_pos, choice) keeps draft `choice` only; the next
// forward expands it to all rows again, :180-200).  On static caches with W rows per request: the kept positions
// [start, start + count) of the winning row are copied over the other W - 1 rows (everything before `start` is the
// shared prefix and identical already).  One warp per (request, draft != choice, head, position).
// With layer tables (kcs / vcs != nullptr: device arrays of per-layer cache pointers) blockIdx.y is the layer: ONE launch
// rolls back every layer of a model.
__global__ void kv_select_kernel(unsigned char* __restrict__ kc, unsigned char* __restrict__ vc, unsigned char* const* __restrict__ kcs,
                                 unsigned char* const* __restrict__ vcs, int B, int W, int H, int S,
                                 int row_bytes, int max_count, const int* __restrict__ choice, const int* __restrict__ start,
                                 int start_stride, const int* __restrict__ count, const int* __restrict__ active, int active_stride) {
  if (kcs != nullptr) { kc = kcs[blockIdx.y]; vc = vcs[blockIdx.y]; }
  const int warps_per_block = blockDim.x >> 5;
  const long long r = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  if (r >= static_cast<long long>(B) * W * H * max_count) return;
  const int j = static_cast<int>(r % max_count);
  const int h = static_cast<int>((r / max_count) % H);
  const int w = static_cast<int>((r / (static_cast<long long>(max_count) * H)) % W);
  const int b = static_cast<int>(r / (static_cast<long long>(max_count) * H * W));
  if (active != nullptr && active[b * active_stride] == 0) return;
  const int c = choice[b];
  if (w == c || j >= count[b]) return;
  const int pos = start[b * start_stride] + j;
  if (pos < 0 || pos >= S) return;
  const long long src = ((static_cast<long long>(b * W + c) * H + h) * S + pos) * row_bytes;
  const long long dst = ((static_cast<long long>(b * W + w) * H + h) * S + pos) * row_bytes;
  for (int o = (threadIdx.x & 31) * 16; o < row_bytes; o += 32 * 16) {
    *reinterpret_cast<uint4*>(kc + dst + o) = *reinterpret_cast<const uint4*>(kc + src + o);
    *reinterpret_cast<uint4*>(vc + dst + o) = *reinterpret_cast<const uint4*>(vc + src + o);
  }
}

cudaError_t launch_kv_select(void* k_cache, void* v_cache, int B, int W, int H, int S, int D, int elem_size, int max_count,
                             const int* choice, const int* start, int start_stride, const int* count, const int* active,
                             int active_stride, cudaStream_t st) {
  const int row_bytes = D * elem_size;
  if (row_bytes % 16 != 0 || max_count < 1) return cudaErrorInvalidValue;
  const long long rows = static_cast<long long>(B) * W * H * max_count;
  if (rows <= 0) return cudaSuccess;
  const int wpb = 8;
  kv_select_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(
      static_cast<unsigned char*>(k_cache), static_cast<unsigned char*>(v_cache), nullptr, nullptr, B, W, H, S, row_bytes, max_count,
      choice, start, start_stride, count, active, active_stride);
  return cudaGetLastError();
}

cudaError_t launch_kv_select_layers(void* const* k_caches, void* const* v_caches, int n_layers, int B, int W, int H, int S, int D,
                                    int elem_size, int max_count, const int* choice, const int* start, int start_stride,
                                    const int* count, const int* active, int active_stride, cudaStream_t st) {
  const int row_bytes = D * elem_size;
  if (row_bytes % 16 != 0 || max_count < 1 || n_layers < 1 || n_layers > 65535) return cudaErrorInvalidValue;
  const long long rows = static_cast<long long>(B) * W * H * max_count;
  if (rows <= 0) return cudaSuccess;
  const int wpb = 8;
  dim3 grid(static_cast<unsigned>((rows + wpb - 1) / wpb), static_cast<unsigned>(n_layers));
  kv_select_kernel<<<grid, wpb * 32, 0, st>>>(nullptr, nullptr, reinterpret_cast<unsigned char* const*>(k_caches),
                                              reinterpret_cast<unsigned char* const*>(v_caches), B, W, H, S, row_bytes, max_count,
                                              choice, start, start_stride, count, active, active_stride);
  return cudaGetLastError();
}

// Token append of the multi-draft loop (reference speculative_sampling.py:1644, :1677): every row of request b becomes
// prefix + the winning draft's accepted tokens + the target's token, and all W lengths advance together.
__global__ void multi_commit_kernel(long long* __restrict__ tokens, long long tokens_stride, int* __restrict__ seq_len, int W,
                                    const int* __restrict__ choice, const int* __restrict__ n_acc,
                                    const long long* __restrict__ next_tok, const int* __restrict__ active, int S) {
  const int b = blockIdx.x;
  if (active != nullptr && active[b * W] == 0) return;
  const int c = choice[b], n = n_acc[b];
  const int L = seq_len[b * W];
  const long long* src = tokens + static_cast<long long>(b * W + c) * tokens_stride;
  for (int i = threadIdx.x; i < W * (n + 1); i += blockDim.x) {
    const int w = i / (n + 1), j = i - w * (n + 1);
    if (L + j >= S) continue;
    long long* dst = tokens + static_cast<long long>(b * W + w) * tokens_stride;
    if (j == n) dst[L + j] = next_tok[b];
    else if (w != c) dst[L + j] = src[L + j];
  }
  __syncthreads();
  if (threadIdx.x < W) seq_len[b * W + threadIdx.x] = L + n + 1;
}

cudaError_t launch_multi_commit(long long* tokens, long long tokens_stride, int* seq_len, int B, int W, const int* choice,
                                const int* n_acc, const long long* next_tok, const int* active, int S, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  if (W < 1 || W > 128) return cudaErrorInvalidValue;
  multi_commit_kernel<<<static_cast<unsigned>(B), 128, 0, st>>>(tokens, tokens_stride, seq_len, W, choice, n_acc, next_tok, active, S);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Step builder.  For request b with current length L = seq_len[b] the step consumes the q tokens at
// positions start .. start+q-1, start = L + offset.  If prev_tok is given it is first stored at the
// last of these positions (the token the previous draft step sampled).  Emits input ids, position
// ids, the cache write offset and the boolean attention mask (key s visible to query j iff
// s <= start + j), so that nothing in the gamma-step loop depends on host-side lengths.
__global__ void build_step_kernel(long long* __restrict__ tokens, long long tokens_stride,
                                  const int* __restrict__ seq_len, int offset, int q,
                                  const long long* __restrict__ prev_tok, int S,
                                  long long* __restrict__ input_ids, long long* __restrict__ position_ids,
                                  int* __restrict__ write_pos, unsigned char* __restrict__ mask) {
  const int b = blockIdx.x;
  const int start = max(seq_len[b] + offset, 0);
  long long* trow = tokens + b * tokens_stride;
  if (threadIdx.x < q) {
    const int j = threadIdx.x;
    // A request that has finished keeps being stepped while others in the batch run (its results are discarded): its
    // positions may reach past the token buffer.  Reads are clamped, the append is skipped — never touch row b + 1.
    const bool inb = start + j < S;
    long long tok = trow[inb ? start + j : S - 1];
    if (prev_tok != nullptr && j == q - 1) { tok = prev_tok[b]; if (inb) trow[start + j] = tok; }
    input_ids[b * q + j] = tok;
    position_ids[b * q + j] = min(start + j, S - 1);
  }
  if (threadIdx.x == 0) write_pos[b] = start;
  if (mask != nullptr) {
    unsigned char* mrow = mask + static_cast<long long>(b) * q * S;
    for (int i = threadIdx.x; i < q * S; i += blockDim.x) {
      const int j = i / S, s = i - j * S;
      mrow[i] = s <= start + j ? 1 : 0;
    }
  }
}

cudaError_t launch_build_step(long long* tokens, long long tokens_stride, const int* seq_len, int offset, int q,
                              const long long* prev_tok, int B, int S, long long* input_ids, long long* position_ids,
                              int* write_pos, unsigned char* mask, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  if (q < 1 || q > 256) return cudaErrorInvalidValue;
  build_step_kernel<<<B, 256, 0, st>>>(tokens, tokens_stride, seq_len, offset, q, prev_tok, S, input_ids,
                                       position_ids, write_pos, mask);
  return cudaGetLastError();
}

}  // namespace sd
