// rb200_axis.cu — K6: axis reduction (column form).
#include "rb200_interp.cuh"
#include "rb200_launch.h"
namespace rb200 {
// ---------------------------------------------------------------------------------------------
// K6: axis reduction, column form. Iteration dims are ordered [reduced..., kept...]; a work item
// is (split, kept row, chunk): it walks its slice of the reduced range sequentially running the
// whole op list per element (like the pndindex(itershape2) x ndindex(itershape3) nest of
// ramba/ramba.py:8235-8244) and keeps V accumulators per reduction slot in registers; partials
// go to part[slot][split][kept_linear] as raw 64-bit values of the accumulator class.
template <int V> __global__ void __launch_bounds__(kThreads) vm_axis_reduce_kernel(const __grid_constant__ KParams P) {
  extern __shared__ unsigned long long regfile[];
  Ctx<V> cx(P, regfile);
  const int nk0 = P.red_ndim;  // first kept dim
  const long long inner = P.shape[P.ndim - 1];
  long long kept_rows = 1;
#pragma unroll
  for (int d = 0; d < kMaxD; ++d)
    if (d >= nk0 && d < P.ndim - 1) kept_rows *= P.shape[d];
  const long long kept_work = kept_rows * P.n_chunks;
  const long long kept_elems = kept_rows * inner;
  const long long stride_w = (long long)gridDim.x * kThreads;
  for (long long w = (long long)blockIdx.x * kThreads + threadIdx.x; w < P.total_work; w += stride_w) {
    const long long split = w / kept_work;
    const long long kw = w - split * kept_work;
    const long long row = kw / P.n_chunks;
    const long long chunk = kw - row * P.n_chunks;
    const long long j0 = chunk * V;
#pragma unroll
    for (int d = 0; d < kMaxD; ++d) cx.idx[d] = 0;
    {
      long long rem = row;
#pragma unroll
      for (int d = kMaxD - 2; d >= 0; --d) {
        if (d >= nk0 && d < P.ndim - 1) {
          long long sd = P.shape[d];
          long long q = rem / sd;
          cx.idx[d] = rem - q * sd;
          rem = q;
        }
      }
    }
#pragma unroll
    for (int d = 0; d < kMaxD; ++d)
      if (d == P.ndim - 1) cx.idx[d] = j0;
    long long left = inner - j0;
    cx.nvalid = left < V ? (int)left : V;

    Val racc[RB200_MAX_REDS][V];
#pragma unroll
    for (int s = 0; s < RB200_MAX_REDS; ++s)
#pragma unroll
      for (int k = 0; k < V; ++k) racc[s][k] = red_identity(s < P.n_reds ? P.reds[s].op : 0, s < P.n_reds ? P.reds[s].ctype : 0);

    const long long r0 = split * P.red_split;
    long long r1 = r0 + P.red_split;
    if (r1 > P.red_len) r1 = P.red_len;
    for (long long r = r0; r < r1; ++r) {
      // decode r into the leading reduced dims
      long long rem = r;
#pragma unroll
      for (int d = kMaxD - 1; d >= 0; --d) {
        if (d < nk0) {
          if (d == 0) cx.idx[0] = rem;
          else {
            long long sd = P.shape[d];
            long long q = rem / sd;
            cx.idx[d] = rem - q * sd;
            rem = q;
          }
        }
      }
      run_program<V, true>(cx, racc);
    }
    for (int s = 0; s < P.n_reds; ++s) {
#pragma unroll
      for (int q = 0; q < RB200_MAX_REDS; ++q)
        if (q == s) {
#pragma unroll
          for (int k = 0; k < V; ++k)
            if (k < cx.nvalid) P.red_partials[((long long)s * P.n_split + split) * kept_elems + row * inner + j0 + k] = racc[q][k].u;
        }
    }
  }
}


cudaError_t launch_vm_axis_reduce(const KParams& P, unsigned blocks, size_t smem, cudaStream_t stream) {
  constexpr int V = 4;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(vm_axis_reduce_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  vm_axis_reduce_kernel<V><<<blocks, kThreads, smem, stream>>>(P);
  return cudaGetLastError();
}
}  // namespace rb200
