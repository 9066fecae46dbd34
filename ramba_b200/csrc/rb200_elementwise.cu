// rb200_elementwise.cu — K1-K5: fused elementwise kernel with optional global reductions.
#include "rb200_interp.cuh"
#include "rb200_launch.h"
namespace rb200 {
// ---------------------------------------------------------------------------------------------
// K1/K2 (+K3/K4/K5): fused elementwise kernel with optional global reductions.
// grid-stride over work items (row, chunk-of-V along the innermost dim); consecutive threads take
// consecutive chunks so that a warp touches 32*V contiguous elements of every contiguous view.
template <int V> __global__ void __launch_bounds__(kThreads) vm_elementwise_kernel(const __grid_constant__ KParams P) {
  extern __shared__ unsigned long long regfile[];
  Ctx<V> cx(P, regfile);
  Val racc[RB200_MAX_REDS][1];
#pragma unroll
  for (int s = 0; s < RB200_MAX_REDS; ++s) racc[s][0] = red_identity(s < P.n_reds ? P.reds[s].op : 0, s < P.n_reds ? P.reds[s].ctype : 0);

  const long long inner = P.shape[P.ndim - 1];
  const long long stride_w = (long long)gridDim.x * kThreads;
  for (long long w = (long long)blockIdx.x * kThreads + threadIdx.x; w < P.total_work; w += stride_w) {
    long long row, chunk;
    if (P.ndim == 1) {
      row = 0;
      chunk = w;
    } else if (P.total_work < 0x7fffffffll) {
      unsigned uw = (unsigned)w, nc = (unsigned)P.n_chunks;
      row = uw / nc;
      chunk = uw - (unsigned)row * nc;
    } else {
      row = w / P.n_chunks;
      chunk = w - row * P.n_chunks;
    }
    const long long j0 = chunk * V;
#pragma unroll
    for (int d = 0; d < kMaxD; ++d) cx.idx[d] = 0;
    // decode row into outer indices (dims 0..ndim-2), last outer dim fastest
    if (P.ndim > 1) {
      long long rem = row;
#pragma unroll
      for (int d = kMaxD - 2; d >= 0; --d) {
        if (d < P.ndim - 1) {
          if (d == 0) {
            cx.idx[0] = rem;
          } else {
            long long sd = P.shape[d];
            long long q;
            if (rem < 0x7fffffffll && sd < 0x7fffffffll) q = (unsigned)rem / (unsigned)sd;
            else q = rem / sd;
            cx.idx[d] = rem - q * sd;
            rem = q;
          }
        }
      }
    }
#pragma unroll
    for (int d = 0; d < kMaxD; ++d)
      if (d == P.ndim - 1) cx.idx[d] = j0;
    long long left = inner - j0;
    cx.nvalid = left < V ? (int)left : V;
    run_program<V, false>(cx, racc);
  }

  // ---- global reductions: thread -> warp shuffle -> block -> per-block partial -> last block
  if (P.n_reds > 0) {
    __shared__ unsigned long long wpart[RB200_MAX_REDS][kThreads / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = 0; s < P.n_reds; ++s) {
      const int op = P.reds[s].op, ct = P.reds[s].ctype;
      Val v = racc[0][0];
#pragma unroll
      for (int q = 0; q < RB200_MAX_REDS; ++q)
        if (q == s) v = racc[q][0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Val t;
        t.u = __shfl_down_sync(0xffffffffu, v.u, o);
        v = red_combine_val(op, ct, v, t);
      }
      if (lane == 0) wpart[s][warp] = v.u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int s = 0; s < P.n_reds; ++s) {
        const int op = P.reds[s].op, ct = P.reds[s].ctype;
        Val v;
        v.u = wpart[s][0];
        for (int q = 1; q < kThreads / 32; ++q) {
          Val t;
          t.u = wpart[s][q];
          v = red_combine_val(op, ct, v, t);
        }
        P.red_partials[(long long)s * gridDim.x + blockIdx.x] = v.u;
      }
      __threadfence();
      unsigned prev = atomicAdd(P.red_counter, 1u);
      is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      for (int s = 0; s < P.n_reds; ++s) {
        const int op = P.reds[s].op, ct = P.reds[s].ctype;
        // fixed order: thread t folds partials t, t+256, ...; then the same tree as above
        Val v = red_identity(op, ct);
        for (unsigned b = threadIdx.x; b < gridDim.x; b += kThreads) {
          Val t;
          t.u = __ldcg(&P.red_partials[(long long)s * gridDim.x + b]);
          v = red_combine_val(op, ct, v, t);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          Val t;
          t.u = __shfl_down_sync(0xffffffffu, v.u, o);
          v = red_combine_val(op, ct, v, t);
        }
        __syncthreads();
        if (lane == 0) wpart[s][warp] = v.u;
        __syncthreads();
        if (threadIdx.x == 0) {
          v.u = wpart[s][0];
          for (int q = 1; q < kThreads / 32; ++q) {
            Val t;
            t.u = wpart[s][q];
            v = red_combine_val(op, ct, v, t);
          }
          // red[0,..] = red[0,..] (op) acc  (ramba/ramba.py:5805-5806), rounded to the partial
          // array's dtype on store
          void* out = P.reds[s].out;
          switch (P.reds[s].out_dtype) {
            case RB200_F64: { double* o = (double*)out; *o = red_combine<double>(op, *o, ct == RB200_T_F64 ? v.d : (double)v.i); } break;
            case RB200_F32: { float* o = (float*)out; double cur = (double)*o; *o = (float)red_combine<double>(op, cur, ct == RB200_T_F64 ? v.d : (double)v.i); } break;
            case RB200_I64: { long long* o = (long long*)out; *o = red_combine<long long>(op, *o, ct == RB200_T_F64 ? (long long)v.d : v.i); } break;
            case RB200_I32: { int* o = (int*)out; *o = (int)red_combine<long long>(op, (long long)*o, ct == RB200_T_F64 ? (long long)v.d : v.i); } break;
            case RB200_BOOL: { unsigned char* o = (unsigned char*)out; long long t2 = red_combine<long long>(op, (long long)*o, ct == RB200_T_F64 ? (long long)v.d : v.i); *o = t2 != 0 ? 1 : 0; } break;
            case RB200_U8: { unsigned char* o = (unsigned char*)out; *o = (unsigned char)red_combine<long long>(op, (long long)*o, ct == RB200_T_F64 ? (long long)v.d : v.i); } break;
            case RB200_I8: { signed char* o = (signed char*)out; *o = (signed char)red_combine<long long>(op, (long long)*o, ct == RB200_T_F64 ? (long long)v.d : v.i); } break;
            case RB200_I16: { short* o = (short*)out; *o = (short)red_combine<long long>(op, (long long)*o, ct == RB200_T_F64 ? (long long)v.d : v.i); } break;
            case RB200_U16: { unsigned short* o = (unsigned short*)out; *o = (unsigned short)red_combine<long long>(op, (long long)*o, ct == RB200_T_F64 ? (long long)v.d : v.i); } break;
            case RB200_U32: { unsigned int* o = (unsigned int*)out; *o = (unsigned int)red_combine<long long>(op, (long long)*o, ct == RB200_T_F64 ? (long long)v.d : v.i); } break;
            default: break;
          }
        }
      }
      if (threadIdx.x == 0) *P.red_counter = 0u;  // leave scratch ready for the next launch
    }
  }
}


cudaError_t launch_vm_elementwise(const KParams& P, unsigned blocks, size_t smem, cudaStream_t stream) {
  constexpr int V = 4;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(vm_elementwise_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  vm_elementwise_kernel<V><<<blocks, kThreads, smem, stream>>>(P);
  return cudaGetLastError();
}
}  // namespace rb200
