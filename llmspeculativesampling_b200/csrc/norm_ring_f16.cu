// Instantiation unit of the ring kernel for __half logits (one unit per dtype so that the variants compile in parallel).
#include "norm_ring_kernel.cuh"

namespace sd {
cudaError_t ring_dispatch_f16(const NormParams& p, cudaStream_t st) { return ring_dispatch<__half>(p, st); }
}  // namespace sd
