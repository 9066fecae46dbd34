// rb200_elementwise_nd2.cu — instantiation of the fused elementwise kernel for iteration rank 2
// (one translation unit per rank so that they compile in parallel).
// N-d ops use handler set 2 (direct views instead of staged ones)
#define RB200_HANDLER_SET 2
#include "rb200_elementwise.cuh"
namespace rb200 {
cudaError_t launch_vm_elementwise_nd2(const KParams& P, unsigned blocks, size_t smem, cudaStream_t stream) {
  return launch_vm_elementwise_nd<kV, 2>(P, blocks, smem, stream);
}
}  // namespace rb200
