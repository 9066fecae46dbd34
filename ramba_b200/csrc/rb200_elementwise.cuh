#pragma once
// rb200_elementwise.cuh — K1-K5: fused elementwise kernel with optional global reductions,
// instantiated per iteration-space rank in rb200_elementwise_nd*.cu (parallel compilation).
#include "rb200_interp.cuh"
#include "rb200_launch.h"

namespace rb200 {

// Persistent-style grid (a multiple of the SM count): CTA b walks tiles b, b+grid, ...
// Shared memory layout (dynamic): [register file: n_regs*V*256*8 B][prefetch: 2 stages * n_pf*V*256*8 B]
template <int V, int ND> __global__ void __launch_bounds__(kThreads, (ND == 1) ? 3 : 2) vm_elementwise_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int TILE = kThreads * V;
  Ctx<V, ND> cx(P);
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);
  cx.regfile_s = smem_s + threadIdx.x * 8u;
  const unsigned pf_base = smem_s + (unsigned)(P.n_regs * V * kThreads * 8) + threadIdx.x * 8u;
  const unsigned pf_stage_bytes = (unsigned)(P.n_pf * V * kThreads * 8);
  cx.pf_s = pf_base;

  u64 racc[RB200_MAX_REDS][1];
#pragma unroll
  for (int s = 0; s < RB200_MAX_REDS; ++s) racc[s][0] = red_identity_bits(s < P.n_reds ? P.reds[s].op : 0, s < P.n_reds ? P.reds[s].ctype : 0);

  const int n_pf = (ND == 1) ? P.n_pf : 0;
  // stage the inputs of tile `t` (ND == 1 only): one cp.async per element per staged view
  auto issue_prefetch = [&](long long t, unsigned stage_s) {
    const long long e0 = t * TILE + threadIdx.x;
#pragma unroll 1
    for (int j = 0; j < n_pf; ++j) {
      const KView& vw = P.views[P.pf_view[j]];
      const long long s = vw.stride[0];
      const int es = (vw.dtype == RB200_F64 || vw.dtype == RB200_I64) ? 8 : 4;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const long long e = e0 + (long long)k * kThreads;
        const bool ok = e < P.total;
        const char* src = vw.base + (ok ? e * s * es : 0);
        const unsigned dst = stage_s + (unsigned)((j * V + k) * kThreads * 8);
        if (es == 8) cp_async8(dst, src, ok);
        else cp_async4(dst, src, ok);
      }
    }
  };

  long long tile = blockIdx.x;
  unsigned stage = 0;
  if (n_pf > 0) {
    if (tile < P.n_tiles) issue_prefetch(tile, pf_base);
    cp_async_commit();
  }
#pragma unroll 1
  for (; tile < P.n_tiles; tile += gridDim.x) {
    if (n_pf > 0) {
      const long long nxt = tile + gridDim.x;
      if (nxt < P.n_tiles) issue_prefetch(nxt, pf_base + (stage ^ 1u) * pf_stage_bytes);
      cp_async_commit();
      cp_async_wait<1>();  // everything but the group just committed has landed: this tile's inputs
      cx.pf_s = pf_base + stage * pf_stage_bytes;
      stage ^= 1u;
    }
    const long long e0 = tile * TILE + threadIdx.x;
    unsigned valid = 0;
    if constexpr (ND == 1) {
      cx.e0 = e0;
#pragma unroll
      for (int k = 0; k < V; ++k)
        if (e0 + (long long)k * kThreads < P.total) valid |= (1u << k);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const long long e = e0 + (long long)k * kThreads;
        if (e < P.total) {
          valid |= (1u << k);
          decode_index<ND>(P, e, 0, cx.idx[k]);
        } else {
#pragma unroll
          for (int d = 0; d < ND; ++d) cx.idx[k][d] = 0;
        }
      }
    }
    cx.valid = valid;
    run_program<V, false>(cx, racc);
  }
  if (n_pf > 0) cp_async_wait<0>();

  // ---- global reductions: thread -> warp shuffle -> block -> per-block partial -> last block
  if (P.n_reds > 0) {
    __shared__ u64 wpart[RB200_MAX_REDS][kThreads / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = 0; s < P.n_reds; ++s) {
      const int op = P.reds[s].op, ct = P.reds[s].ctype;
      u64 v = racc[0][0];
#pragma unroll
      for (int q = 0; q < RB200_MAX_REDS; ++q)
        if (q == s) v = racc[q][0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = red_combine_bits(op, ct, v, __shfl_down_sync(0xffffffffu, v, o));
      if (lane == 0) wpart[s][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int s = 0; s < P.n_reds; ++s) {
        const int op = P.reds[s].op, ct = P.reds[s].ctype;
        u64 v = wpart[s][0];
        for (int q = 1; q < kThreads / 32; ++q) v = red_combine_bits(op, ct, v, wpart[s][q]);
        P.red_partials[(long long)s * gridDim.x + blockIdx.x] = v;
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
        u64 v = red_identity_bits(op, ct);
        for (unsigned b = threadIdx.x; b < gridDim.x; b += kThreads) v = red_combine_bits(op, ct, v, __ldcg(&P.red_partials[(long long)s * gridDim.x + b]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = red_combine_bits(op, ct, v, __shfl_down_sync(0xffffffffu, v, o));
        __syncthreads();
        if (lane == 0) wpart[s][warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
          v = wpart[s][0];
          for (int q = 1; q < kThreads / 32; ++q) v = red_combine_bits(op, ct, v, wpart[s][q]);
          // red[0,..] = red[0,..] (op) acc  (ramba/ramba.py:5805-5806), rounded to the partial
          // array's dtype on store
          void* out = P.reds[s].out;
          const double vd = (ct == RB200_T_F64) ? CT<double>::get(v) : (double)(long long)v;
          const long long vi = (ct == RB200_T_F64) ? (long long)CT<double>::get(v) : (long long)v;
          switch (P.reds[s].out_dtype) {
            case RB200_F64: { double* o = (double*)out; *o = red_combine<double>(op, *o, vd); } break;
            case RB200_F32: { float* o = (float*)out; *o = (float)red_combine<double>(op, (double)*o, vd); } break;
            case RB200_I64: { long long* o = (long long*)out; *o = red_combine<long long>(op, *o, vi); } break;
            case RB200_I32: { int* o = (int*)out; *o = (int)red_combine<long long>(op, (long long)*o, vi); } break;
            case RB200_BOOL: { unsigned char* o = (unsigned char*)out; *o = red_combine<long long>(op, (long long)*o, vi) != 0 ? 1 : 0; } break;
            case RB200_U8: { unsigned char* o = (unsigned char*)out; *o = (unsigned char)red_combine<long long>(op, (long long)*o, vi); } break;
            case RB200_I8: { signed char* o = (signed char*)out; *o = (signed char)red_combine<long long>(op, (long long)*o, vi); } break;
            case RB200_I16: { short* o = (short*)out; *o = (short)red_combine<long long>(op, (long long)*o, vi); } break;
            case RB200_U16: { unsigned short* o = (unsigned short*)out; *o = (unsigned short)red_combine<long long>(op, (long long)*o, vi); } break;
            case RB200_U32: { unsigned int* o = (unsigned int*)out; *o = (unsigned int)red_combine<long long>(op, (long long)*o, vi); } break;
            default: break;
          }
        }
      }
      if (threadIdx.x == 0) *P.red_counter = 0u;  // leave scratch ready for the next launch
    }
  }
}

template <int V, int ND> cudaError_t launch_vm_elementwise_nd(const KParams& P, unsigned blocks, size_t smem, cudaStream_t stream) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(vm_elementwise_kernel<V, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  vm_elementwise_kernel<V, ND><<<blocks, kThreads, smem, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace rb200
