// rb200_elementwise_ax1d.cu — the 1-D kernel in axis-reduction mode (column sums of row-major boxes
// whose row length is a multiple of the tile): staging, handlers and 8 elements per thread like the
// plain 1-D kernel, V column accumulators per thread.
#include "rb200_elementwise.cuh"
namespace rb200 {
cudaError_t launch_vm_elementwise_ax1d(const KParams& P, unsigned blocks, size_t smem, cudaStream_t stream) {
  return launch_vm_elementwise_nd<kV1, 1, true>(P, blocks, smem, stream);
}
}  // namespace rb200
