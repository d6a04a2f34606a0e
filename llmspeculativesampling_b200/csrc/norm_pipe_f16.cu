// Persistent top-k kernel, __half logits: the kernel instantiations of this dtype (see norm_pipe_kernel.cuh).
#include "norm_pipe_kernel.cuh"

namespace sd {

cudaError_t pipe_dispatch_f16(const NormParams& p, int rows, cudaStream_t st, int* q) {
  return pipe_dispatch<__half>(p, rows, st, q);
}

}  // namespace sd
