// rb200_mapred.cu — K5/K6 for the map + reduce form: `red = red (+) f(X[i])` over ONE contiguous source (sm_100a).
//
// What it stands for in the reference: the generated loop of a reduction whose operand is an elementwise map of one
// array with scalars - `(X*2.0 + 1.0).sum()` stage 1 (ramba/ramba.py:5798-5807, body 8247-8255) - and the axis-reduction
// loop nest over a row-split matrix with a row-broadcast operand, `(M + v).sum(axis=0)` stage 1
// (ramba/ramba.py:5809-5814, 8231-8244).  At 4 bytes per element these must not pay ANY per-element interpretation:
//   * the source is read with 128-bit loads straight into registers, 4 loads in flight per thread (no staging needed:
//     nothing is reused), 16 float / 8 double elements per thread and iteration;
//   * the map is the op list's own chain of scalar operations, decoded ONCE per 16 elements by warp-uniform branches
//     (the scalar chain is data-independent), applied in the op list's order and classes, one rounding each;
//   * global form: float64 accumulator per thread -> warp shuffle -> CTA -> last CTA (fixed order); column form: every
//     thread owns 4 (2) consecutive columns and walks its rows, partials[split][column] like the general kernels.
// The form is recognised on the TERM list (rb200_terms.h), so bit-identical results to every other path are kept.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "rb200_launch.h"
#include "rb200_lean.cuh"
#include "rb200_terms.h"
#include "rb200_mapred.h"

namespace rb200 {

enum MrCode { M_ADDW = 0, M_SUBW, M_RSUBW, M_MULW, M_NEG, M_ROUND32, M_ADDV, M_SUBV, M_RSUBV, M_MULV };

template <class F, int N> __device__ __forceinline__ void mr_apply(int code, F w, const F (&v)[N], F (&a)[N]) {
  switch (code) {
    case M_ADDW:
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = l_add<F>(a[k], w);
      break;
    case M_SUBW:
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = l_sub<F>(a[k], w);
      break;
    case M_RSUBW:
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = l_sub<F>(w, a[k]);
      break;
    case M_MULW:
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = l_mul<F>(a[k], w);
      break;
    case M_NEG:
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = -a[k];
      break;
    case M_ROUND32:
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = (F)(float)a[k];
      break;
    case M_ADDV:
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = l_add<F>(a[k], v[k]);
      break;
    case M_SUBV:
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = l_sub<F>(a[k], v[k]);
      break;
    case M_RSUBV:
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = l_sub<F>(v[k], a[k]);
      break;
    default:  // M_MULV
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = l_mul<F>(a[k], v[k]);
  }
}

// the map of N elements: float32 phase (source float32), promotion, float64 phase.  vf / vd: the broadcast operand of these
// N elements in both classes.
template <int N> __device__ __forceinline__ void mr_map32(const MrParams& P, const float (&x)[N], const float (&vf)[N], const double (&vd)[N], double (&out)[N]) {
  float a[N];
#pragma unroll
  for (int k = 0; k < N; ++k) a[k] = x[k];
#pragma unroll 1
  for (int j = 0; j < P.n32; ++j) mr_apply<float, N>(P.code[j], (float)P.w[j], vf, a);
#pragma unroll
  for (int k = 0; k < N; ++k) out[k] = (double)a[k];
#pragma unroll 1
  for (int j = P.n32; j < P.n32 + P.n64; ++j) mr_apply<double, N>(P.code[j], P.w[j], vd, out);
}
template <int N> __device__ __forceinline__ void mr_map64(const MrParams& P, const double (&vd)[N], double (&a)[N]) {
#pragma unroll 1
  for (int j = 0; j < P.n64; ++j) mr_apply<double, N>(P.code[j], P.w[j], vd, a);
}

__device__ __forceinline__ float4 ldg128f(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ldg128d(const double* p) {
  double2 v;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

constexpr int kMrU = 4;  // 128-bit loads in flight per thread

// ---- mode 0: global reduction over `total` contiguous elements
template <class TE> __global__ void __launch_bounds__(kThreads) mapred_global_kernel(const __grid_constant__ MrParams P) {
  constexpr int VEC = 16 / (int)sizeof(TE);  // elements per 128-bit load
  constexpr int N = VEC * kMrU;              // elements per thread and iteration
  const TE* src = reinterpret_cast<const TE*>(P.src);
  const int rop = P.redop;
  double acc = CT<double>::get(red_identity_bits(rop, RB200_T_F64));
  const long long n_full = P.total / (long long)(N * kThreads);  // whole blocks of N * 256 elements
  const long long n_blk = n_full * kThreads;                     // (block, thread) groups
  const float zf[N] = {};
  const double zd[N] = {};
  for (long long g = (long long)blockIdx.x * kThreads + threadIdx.x; g < n_blk; g += (long long)gridDim.x * kThreads) {
    // group g: VEC-element vectors g, g + n_blk, ... would stride badly; instead a group is N CONSECUTIVE elements of
    // one warp-interleaved block: thread-contiguous 128-bit pieces, a warp reads 512 consecutive bytes per load
    const long long blk = g / kThreads, t = g % kThreads;
    const TE* base = src + blk * (long long)(N * kThreads) + t * VEC;
    double m[N];
    if constexpr (sizeof(TE) == 4) {
      float x[N];
#pragma unroll
      for (int u = 0; u < kMrU; ++u) {
        const float4 v = ldg128f(base + (long long)u * (VEC * kThreads));
        x[u * 4 + 0] = v.x; x[u * 4 + 1] = v.y; x[u * 4 + 2] = v.z; x[u * 4 + 3] = v.w;
      }
      mr_map32<N>(P, x, zf, zd, m);
    } else {
#pragma unroll
      for (int u = 0; u < kMrU; ++u) {
        const double2 v = ldg128d(base + (long long)u * (VEC * kThreads));
        m[u * 2 + 0] = v.x; m[u * 2 + 1] = v.y;
      }
      mr_map64<N>(P, zd, m);
    }
#pragma unroll
    for (int w = N / 2; w > 0; w /= 2) {
#pragma unroll
      for (int k = 0; k < w; ++k) m[k] = red_combine<double>(rop, m[k], m[k + w]);
    }
    acc = red_combine<double>(rop, acc, m[0]);
  }
  // ragged tail: one element per thread
  for (long long e = n_full * (long long)(N * kThreads) + (long long)blockIdx.x * kThreads + threadIdx.x; e < P.total; e += (long long)gridDim.x * kThreads) {
    double m1[1];
    const float zf1[1] = {0.f};
    const double zd1[1] = {0.0};
    if constexpr (sizeof(TE) == 4) {
      const float x1[1] = {ldg<float>(reinterpret_cast<const float*>(src) + e)};
      mr_map32<1>(P, x1, zf1, zd1, m1);
    } else {
      m1[0] = ldg<double>(reinterpret_cast<const double*>(src) + e);
      mr_map64<1>(P, zd1, m1);
    }
    acc = red_combine<double>(rop, acc, m1[0]);
  }
  // thread -> warp -> CTA -> per-CTA partial -> last CTA (fixed order)
  __shared__ u64 wpart[kThreads / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 v = CT<double>::bits(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_combine_bits(rop, RB200_T_F64, v, __shfl_down_sync(0xffffffffu, v, o));
  if (lane == 0) wpart[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    v = wpart[0];
    for (int q = 1; q < kThreads / 32; ++q) v = red_combine_bits(rop, RB200_T_F64, v, wpart[q]);
    P.red_partials[blockIdx.x] = v;
    __threadfence();
    const unsigned prev = atomicAdd(P.red_counter, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    v = red_identity_bits(rop, RB200_T_F64);
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kThreads) v = red_combine_bits(rop, RB200_T_F64, v, __ldcg(&P.red_partials[b]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_combine_bits(rop, RB200_T_F64, v, __shfl_down_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) wpart[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      v = wpart[0];
      for (int q = 1; q < kThreads / 32; ++q) v = red_combine_bits(rop, RB200_T_F64, v, wpart[q]);
      // red[0,..] = red[0,..] (op) acc  (ramba/ramba.py:5805-5806), rounded to the partial array's dtype on store
      const double vd = CT<double>::get(v);
      if (P.red.out_dtype == RB200_F64) {
        double* o = (double*)P.red.out;
        *o = red_combine<double>(rop, *o, vd);
      } else {
        float* o = (float*)P.red.out;
        *o = (float)red_combine<double>(rop, (double)*o, vd);
      }
      *P.red_counter = 0u;
    }
  }
}

// ---- mode 1: column form.  A CTA owns kThreads * VEC consecutive columns and rows [r0, r1); thread t owns columns
// col0 + t*VEC .. +VEC-1.
template <class TE, class TV> __global__ void __launch_bounds__(kThreads) mapred_columns_kernel(const __grid_constant__ MrParams P) {
  constexpr int VEC = 16 / (int)sizeof(TE);
  const long long split = blockIdx.x / (unsigned)P.n_chunks;
  const long long chunk = blockIdx.x - split * P.n_chunks;
  const long long col = chunk * (long long)(kThreads * VEC) + (long long)threadIdx.x * VEC;
  const long long r0 = split * P.rows_per_split;
  long long r1 = r0 + P.rows_per_split;
  if (r1 > P.R) r1 = P.R;
  const int rop = P.redop;
  double cacc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) cacc[k] = CT<double>::get(red_identity_bits(rop, RB200_T_F64));
  // the row-broadcast operand of my columns, in both classes
  float vf[VEC];
  double vd[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    vf[k] = 0.f;
    vd[k] = 0.0;
  }
  if (P.vsrc) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const TV x = reinterpret_cast<const TV*>(P.vsrc)[col + k];
      vf[k] = (float)x;
      vd[k] = (double)x;
    }
  }
  const TE* src = reinterpret_cast<const TE*>(P.src) + col;
  long long r = r0;
  for (; r + kMrU <= r1; r += kMrU) {
    double m[kMrU][VEC];
    if constexpr (sizeof(TE) == 4) {
      float x[kMrU][VEC];
#pragma unroll
      for (int u = 0; u < kMrU; ++u) {
        const float4 v = ldg128f(src + (r + u) * P.C);
        x[u][0] = v.x; x[u][1] = v.y; x[u][2] = v.z; x[u][3] = v.w;
      }
#pragma unroll
      for (int u = 0; u < kMrU; ++u) mr_map32<VEC>(P, x[u], vf, vd, m[u]);
    } else {
#pragma unroll
      for (int u = 0; u < kMrU; ++u) {
        const double2 v = ldg128d(src + (r + u) * P.C);
        m[u][0] = v.x; m[u][1] = v.y;
      }
#pragma unroll
      for (int u = 0; u < kMrU; ++u) mr_map64<VEC>(P, vd, m[u]);
    }
#pragma unroll
    for (int u = 0; u < kMrU; ++u) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) cacc[k] = red_combine<double>(rop, cacc[k], m[u][k]);
    }
  }
  for (; r < r1; ++r) {
    double m[VEC];
    if constexpr (sizeof(TE) == 4) {
      const float4 v = ldg128f(src + r * P.C);
      const float x[VEC] = {v.x, v.y, v.z, v.w};
      mr_map32<VEC>(P, x, vf, vd, m);
    } else {
      const double2 v = ldg128d(src + r * P.C);
      m[0] = v.x; m[1] = v.y;
      mr_map64<VEC>(P, vd, m);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) cacc[k] = red_combine<double>(rop, cacc[k], m[k]);
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) P.red_partials[split * P.C + col + k] = CT<double>::bits(cacc[k]);
}

// =============================================================================================
// host side: recognise the form on the term list the streaming planner built

// terms: SET x [* w], then scalar / broadcast-operand operations, exactly one RED at the very end.  `src_of(t)` describes
// the operand of a view term (nullptr base: not usable).
int mapred_try(int mode, const TermStep* terms, int n_terms, int n32, const u64* scal, MrSource (*src_of)(void*, const TermStep&), void* ctx, MrParams* out) {
  MrParams P;
  memset(&P, 0, sizeof(P));
  P.mode = mode;
  if (n_terms < 2 || terms[n_terms - 1].kind != TK_RED) return 1;
  const TermStep& first = terms[0];
  if (first.kind != TK_SET || first.xkind == X_NONE) return 1;
  const MrSource s0 = src_of(ctx, first);
  if (!s0.base || s0.row_broadcast || (((uintptr_t)s0.base) & 15u) != 0) return 1;
  P.src = s0.base;
  P.src_f32 = s0.f32;
  if (n32 > 0 && !s0.f32) return 1;  // a float32 phase over a float64 source narrows on fetch: not this form
  int no = 0;
  auto scalar_of = [&](const TermStep& t, bool f32cls) -> double {
    const u64 bits = scal[t.sidx];
    if (f32cls) {
      float f;
      const unsigned u = (unsigned)bits;
      memcpy(&f, &u, 4);
      return (double)f;
    }
    double d;
    memcpy(&d, &bits, 8);
    return d;
  };
  auto push = [&](int code, double w) -> bool {
    if (no >= kMrMaxOps) return false;
    P.code[no] = code;
    P.w[no] = w;
    ++no;
    return true;
  };
  const bool first_f32cls = n32 > 0;
  int n32_ops = 0;
  if (first.flags & TF_W) {
    if (!push(M_MULW, scalar_of(first, first_f32cls))) return 1;
  }
  if (first_f32cls) n32_ops = no;
  for (int i = 1; i < n_terms - 1; ++i) {
    const TermStep& t = terms[i];
    const bool f32cls = i < n32;
    int code = -1;
    double w = 0.0;
    if (t.kind == TK_NEG) {
      code = M_NEG;
    } else if (t.kind == TK_ROUND32) {
      code = M_ROUND32;
    } else if (t.kind == TK_ADD || t.kind == TK_MUL) {
      if (t.xkind == X_NONE) {
        w = scalar_of(t, f32cls);
        code = t.kind == TK_MUL ? M_MULW : (t.flags & TF_NEGP) ? M_SUBW : (t.flags & TF_NEGACC) ? M_RSUBW : M_ADDW;
      } else {
        if (t.flags & TF_W) return 1;  // (weighted broadcast operand: not covered)
        const MrSource sv = src_of(ctx, t);
        if (mode != 1 || !sv.base || !sv.row_broadcast) return 1;
        if (P.vsrc && P.vsrc != sv.base) return 1;  // one broadcast operand
        P.vsrc = sv.base;
        P.v_f32 = sv.f32;
        code = t.kind == TK_MUL ? M_MULV : (t.flags & TF_NEGP) ? M_SUBV : (t.flags & TF_NEGACC) ? M_RSUBV : M_ADDV;
      }
    } else {
      return 1;  // stores, further sets, reductions in the middle
    }
    if (!push(code, w)) return 1;
    if (f32cls) n32_ops = no;
  }
  P.n32 = n32_ops;
  P.n64 = no - n32_ops;
  P.redop = terms[n_terms - 1].dzl;
  *out = P;
  return 0;
}

cudaError_t mapred_launch(const MrParams& P, unsigned blocks, cudaStream_t stream) {
  if (P.mode == 0) {
    if (P.src_f32) mapred_global_kernel<float><<<blocks, kThreads, 0, stream>>>(P);
    else mapred_global_kernel<double><<<blocks, kThreads, 0, stream>>>(P);
  } else {
    if (P.src_f32) {
      if (P.v_f32) mapred_columns_kernel<float, float><<<blocks, kThreads, 0, stream>>>(P);
      else mapred_columns_kernel<float, double><<<blocks, kThreads, 0, stream>>>(P);
    } else {
      if (P.v_f32) mapred_columns_kernel<double, float><<<blocks, kThreads, 0, stream>>>(P);
      else mapred_columns_kernel<double, double><<<blocks, kThreads, 0, stream>>>(P);
    }
  }
  return cudaGetLastError();
}

}  // namespace rb200
