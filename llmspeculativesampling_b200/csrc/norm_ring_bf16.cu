// Instantiation unit of the ring kernel for __nv_bfloat16 logits (one unit per dtype so that the variants compile in parallel).
#include "norm_ring_kernel.cuh"

namespace sd {
cudaError_t ring_dispatch_bf16(const NormParams& p, cudaStream_t st) { return ring_dispatch<__nv_bfloat16>(p, st); }
}  // namespace sd
