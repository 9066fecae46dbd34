#pragma once
#include "rb200_vm.cuh"

namespace rb200 {
// launchers (one translation unit per kernel so that they compile in parallel)
cudaError_t launch_vm_elementwise(const KParams& P, unsigned blocks, size_t smem, cudaStream_t stream);
cudaError_t launch_vm_axis_reduce(const KParams& P, unsigned blocks, size_t smem, cudaStream_t stream);
}  // namespace rb200
