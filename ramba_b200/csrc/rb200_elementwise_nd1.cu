// rb200_elementwise_nd1.cu — instantiation of the fused elementwise kernel for iteration rank 1
// (one translation unit per rank so that they compile in parallel).
#include "rb200_elementwise.cuh"
namespace rb200 {
cudaError_t launch_vm_elementwise_nd1(const KParams& P, unsigned blocks, size_t smem, cudaStream_t stream) {
  return launch_vm_elementwise_nd<kV1, 1>(P, blocks, smem, stream);
}
}  // namespace rb200
