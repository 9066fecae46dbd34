// rb200_axis.cu — K6: axis reduction (column form).
// generic decode path only (the specialised handlers with an N-d context blow up build time)
#define RB200_NO_FAST_HANDLERS 1
#include "rb200_interp.cuh"
#include "rb200_launch.h"
namespace rb200 {

// Iteration dims are ordered [reduced..., kept...]; a work item is (split, V kept elements): it
// walks its slice of the reduced range sequentially running the whole op list per element (like the
// pndindex(itershape2) x ndindex(itershape3) nest of ramba/ramba.py:8235-8244) and keeps V
// accumulators per reduction slot in registers; partials go to
// part[(slot*n_split + split)*kept_elems + kept_linear] as raw 64-bit values of the accumulator class.
// Consecutive threads own consecutive kept elements (coalesced along the innermost kept dim).
template <int V> __global__ void __launch_bounds__(kThreads, 2) vm_axis_reduce_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int ND = kMaxD;
  constexpr int TILE = kThreads * V;
  Ctx<V, ND> cx(P);
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);
  cx.regfile_s = smem_s + threadIdx.x * 8u;
  cx.pf_s = 0;
  cx.tid = threadIdx.x;
  cx.ocls_s = 0;
  const int nk0 = P.red_ndim;       // first kept dim
  const long long kept = P.total;   // kept elements
  const long long tiles_per_split = (kept + TILE - 1) / TILE;
#pragma unroll 1
  for (long long t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
    const long long split = t / tiles_per_split;
    const long long kt = t - split * tiles_per_split;
    const long long e0 = kt * TILE + threadIdx.x;
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const long long e = e0 + (long long)k * kThreads;
#pragma unroll
      for (int d = 0; d < ND; ++d) cx.idx[k][d] = 0;
      if (e < kept) {
        valid |= (1u << k);
        decode_index<ND>(P, e, nk0, cx.idx[k]);
      }
    }
    cx.valid = valid;
    u64 racc[RB200_MAX_REDS][V];
#pragma unroll
    for (int s = 0; s < RB200_MAX_REDS; ++s)
#pragma unroll
      for (int k = 0; k < V; ++k) racc[s][k] = red_identity_bits(s < P.n_reds ? P.reds[s].op : 0, s < P.n_reds ? P.reds[s].ctype : 0);

    const long long r0 = split * P.red_split;
    long long r1 = r0 + P.red_split;
    if (r1 > P.red_len) r1 = P.red_len;
#pragma unroll 1
    for (long long r = r0; r < r1; ++r) {
      // decode r into the leading reduced dims (same for all V elements)
      long long rem = r;
#pragma unroll
      for (int d = ND - 1; d >= 0; --d) {
        if (d < nk0) {
          long long v;
          if (d == 0) v = rem;
          else {
            const long long sd = P.shape[d];
            const long long q = rem / sd;
            v = rem - q * sd;
            rem = q;
          }
#pragma unroll
          for (int k = 0; k < V; ++k) cx.idx[k][d] = v;
        }
      }
      run_program<V, true, RB200_MAX_REDS>(cx, racc);
    }
    for (int s = 0; s < P.n_reds; ++s) {
#pragma unroll
      for (int q = 0; q < RB200_MAX_REDS; ++q)
        if (q == s) {
#pragma unroll
          for (int k = 0; k < V; ++k)
            if ((valid >> k) & 1u) P.red_partials[((long long)s * P.n_split + split) * kept + e0 + (long long)k * kThreads] = racc[q][k];
        }
    }
  }
}

cudaError_t launch_vm_axis_reduce(const KParams& P, unsigned blocks, size_t smem, cudaStream_t stream) {
  if (smem + 2048 > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(vm_axis_reduce_kernel<kV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  vm_axis_reduce_kernel<kV><<<blocks, kThreads, smem, stream>>>(P);
  return cudaGetLastError();
}
}  // namespace rb200
