// Instantiation unit of the ring kernel for float logits (one unit per dtype so that the variants compile in parallel).
#include "norm_ring_kernel.cuh"

namespace sd {
cudaError_t ring_dispatch_f32(const NormParams& p, cudaStream_t st) { return ring_dispatch<float>(p, st); }
}  // namespace sd
