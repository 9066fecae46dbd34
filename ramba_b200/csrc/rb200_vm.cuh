// rb200_vm.cuh — device side of the op-list accumulator machine (sm_100a).
//
// The reference's worker executes Python source generated per fused op and JIT-compiled by
// Numba (ramba/ramba.py:8247-8265, 3758-3780).  Here the same loop body is an op list that a
// hand-written CUDA kernel walks: every thread owns V consecutive elements of the innermost
// iteration dim, keeps the running value in registers (the accumulator), spills to a shared
// memory register file only when the host-side allocator says a value is needed later, and
// touches HBM only for live array views (16-byte vector loads/stores when the view is
// contiguous and aligned).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ramba_b200.h"

namespace rb200 {

constexpr int kThreads = 256;
constexpr int kMaxD = RB200_MAX_DIMS;

struct KView {
  char* base;
  long long stride[kMaxD];  // elements
  int dtype;
  int vec;  // 1: innermost stride 1 and base/rows 16B aligned -> vector path
};

struct KRed {
  int op, ctype;
  void* out;
  int out_dtype;
  int pad;
};

struct KParams {
  int ndim, n_insns, n_views, n_regs, n_reds, axis_mode;
  long long shape[kMaxD];
  long long gstart[kMaxD];
  long long n_chunks;     // chunks of V along the innermost dim
  long long total_work;   // rows * n_chunks
  // axis reduction (column form): kept work items x splits of the reduced range
  long long red_len;      // product of reduced dims (walked sequentially)
  long long red_split;    // elements of the reduced range per split
  int n_split;
  int red_ndim;           // number of leading reduced dims folded into red_len
  long long red_shape[kMaxD];
  KView views[RB200_MAX_VIEWS];
  unsigned long long scalars[RB200_MAX_SCALARS];
  rb200_insn insns[RB200_MAX_INSNS];
  KRed reds[RB200_MAX_REDS];
  unsigned long long* red_partials;  // [n_reds][grid]
  unsigned int* red_counter;
};

union Val {
  double d;
  float f;
  long long i;
  unsigned long long u;
};

// ---------------------------------------------------------------------------------------------
// type traits for the three compute classes
template <class T> struct CT;
template <> struct CT<double> {
  static __device__ __forceinline__ double get(const Val& v) { return v.d; }
  static __device__ __forceinline__ void set(Val& v, double x) { v.d = x; }
  static __device__ __forceinline__ double scal(unsigned long long u) { return __longlong_as_double((long long)u); }
};
template <> struct CT<float> {
  static __device__ __forceinline__ float get(const Val& v) { return v.f; }
  static __device__ __forceinline__ void set(Val& v, float x) { v.u = 0; v.f = x; }
  static __device__ __forceinline__ float scal(unsigned long long u) { return __uint_as_float((unsigned)u); }
};
template <> struct CT<long long> {
  static __device__ __forceinline__ long long get(const Val& v) { return v.i; }
  static __device__ __forceinline__ void set(Val& v, long long x) { v.i = x; }
  static __device__ __forceinline__ long long scal(unsigned long long u) { return (long long)u; }
};

// conversions between storage values and compute classes (C semantics == Numba/LLVM casts:
// float->int truncates toward zero (fptosi), int->float rounds to nearest, f64->f32 rn)
template <class T, class S> __device__ __forceinline__ T cvt(S x) { return (T)x; }

// ---------------------------------------------------------------------------------------------
// global loads / stores. Streams are touched once per launch: bypass L1 allocation for the
// vector path (ld.global.nc / st.global with L1::no_allocate), default caching for scalar
// (possibly re-used: broadcast, shifted stencil) accesses.
__device__ __forceinline__ int4 ldg_stream16(const void* p) {
  int4 r;
  asm volatile("ld.global.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream16(void* p, int4 v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

template <class T, class S, int V>
__device__ __forceinline__ void load_typed(const KView& vw, long long off, long long istride, int nvalid, bool vec,
                                           T (&out)[V]) {
  const S* p = reinterpret_cast<const S*>(vw.base) + off;
  if (vec && nvalid == V) {
    constexpr int BYTES = V * (int)sizeof(S);
    if constexpr (BYTES % 16 == 0) {
      S tmp[V];
#pragma unroll
      for (int k = 0; k < BYTES / 16; ++k) reinterpret_cast<int4*>(tmp)[k] = ldg_stream16(reinterpret_cast<const int4*>(p) + k);
#pragma unroll
      for (int k = 0; k < V; ++k) out[k] = cvt<T, S>(tmp[k]);
      return;
    } else if constexpr (BYTES == 8) {
      S tmp[V];
      *reinterpret_cast<int2*>(tmp) = *reinterpret_cast<const int2*>(p);
#pragma unroll
      for (int k = 0; k < V; ++k) out[k] = cvt<T, S>(tmp[k]);
      return;
    } else if constexpr (BYTES == 4) {
      S tmp[V];
      *reinterpret_cast<int*>(tmp) = *reinterpret_cast<const int*>(p);
#pragma unroll
      for (int k = 0; k < V; ++k) out[k] = cvt<T, S>(tmp[k]);
      return;
    }
  }
  if (istride == 0) {
    T x = cvt<T, S>(*p);
#pragma unroll
    for (int k = 0; k < V; ++k) out[k] = x;
    return;
  }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    if (k < nvalid) out[k] = cvt<T, S>(p[(long long)k * istride]);
    else out[k] = T(0);
  }
}

// narrow integer dtypes are off the hot path: one out-of-line copy per compute class
template <class T, int V> struct Pack { T v[V]; };
template <class T, int V>
__device__ __noinline__ Pack<T, V> load_view_narrow(const KView& vw, long long off, long long istride, int nvalid);

template <class T, int V>
__device__ __forceinline__ void load_view(const KView& vw, long long off, long long istride, int nvalid, T (&out)[V]) {
  const bool vec = vw.vec != 0;
  switch (vw.dtype) {
    case RB200_F64: load_typed<T, double, V>(vw, off, istride, nvalid, vec, out); break;
    case RB200_F32: load_typed<T, float, V>(vw, off, istride, nvalid, vec, out); break;
    case RB200_I64: load_typed<T, long long, V>(vw, off, istride, nvalid, vec, out); break;
    case RB200_I32: load_typed<T, int, V>(vw, off, istride, nvalid, vec, out); break;
    default: {
      Pack<T, V> p = load_view_narrow<T, V>(vw, off, istride, nvalid);
#pragma unroll
      for (int k = 0; k < V; ++k) out[k] = p.v[k];
    }
  }
}

template <class T, int V>
__device__ __noinline__ Pack<T, V> load_view_narrow(const KView& vw, long long off, long long istride, int nvalid) {
  const bool vec = vw.vec != 0;
  Pack<T, V> pk;
  T (&out)[V] = pk.v;
  switch (vw.dtype) {
    case RB200_BOOL:
    case RB200_U8: load_typed<T, unsigned char, V>(vw, off, istride, nvalid, vec, out); break;
    case RB200_I8: load_typed<T, signed char, V>(vw, off, istride, nvalid, vec, out); break;
    case RB200_I16: load_typed<T, short, V>(vw, off, istride, nvalid, vec, out); break;
    case RB200_U16: load_typed<T, unsigned short, V>(vw, off, istride, nvalid, vec, out); break;
    case RB200_U32: load_typed<T, unsigned int, V>(vw, off, istride, nvalid, vec, out); break;
    default:
#pragma unroll
      for (int k = 0; k < V; ++k) out[k] = T(0);
  }
  return pk;
}

template <class T, class S> __device__ __forceinline__ S store_cvt(T x) { return (S)x; }
// bool stores: any non-zero -> 1 (numpy bool_ cast)
template <class T> __device__ __forceinline__ unsigned char store_bool(T x) { return x != T(0) ? 1 : 0; }

template <class T, class S, int V, bool IS_BOOL>
__device__ __forceinline__ void store_typed(const KView& vw, long long off, long long istride, int nvalid,
                                            const T (&val)[V], unsigned mask) {
  S* p = reinterpret_cast<S*>(vw.base) + off;
  S tmp[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    if constexpr (IS_BOOL) tmp[k] = (S)store_bool<T>(val[k]);
    else tmp[k] = store_cvt<T, S>(val[k]);
  }
  constexpr int BYTES = V * (int)sizeof(S);
  constexpr unsigned FULL = (1u << V) - 1u;
  if (vw.vec != 0 && nvalid == V && mask == FULL) {
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
      for (int k = 0; k < BYTES / 16; ++k) stg_stream16(reinterpret_cast<int4*>(p) + k, reinterpret_cast<int4*>(tmp)[k]);
      return;
    } else if constexpr (BYTES == 8) {
      *reinterpret_cast<int2*>(p) = *reinterpret_cast<int2*>(tmp);
      return;
    } else if constexpr (BYTES == 4) {
      *reinterpret_cast<int*>(p) = *reinterpret_cast<int*>(tmp);
      return;
    }
  }
  if (istride == 0) {
    // broadcast target (reduction accumulators written through a stride-0 view): last valid wins
    int last = -1;
#pragma unroll
    for (int k = 0; k < V; ++k)
      if (k < nvalid && ((mask >> k) & 1u)) last = k;
    if (last >= 0) *p = tmp[last];
    return;
  }
#pragma unroll
  for (int k = 0; k < V; ++k)
    if (k < nvalid && ((mask >> k) & 1u)) p[(long long)k * istride] = tmp[k];
}

template <class T, int V>
__device__ __noinline__ void store_view_narrow(const KView& vw, long long off, long long istride, int nvalid,
                                               Pack<T, V> pk, unsigned mask);

template <class T, int V>
__device__ __forceinline__ void store_view(const KView& vw, long long off, long long istride, int nvalid,
                                           const T (&val)[V], unsigned mask) {
  switch (vw.dtype) {
    case RB200_F64: store_typed<T, double, V, false>(vw, off, istride, nvalid, val, mask); break;
    case RB200_F32: store_typed<T, float, V, false>(vw, off, istride, nvalid, val, mask); break;
    case RB200_I64: store_typed<T, long long, V, false>(vw, off, istride, nvalid, val, mask); break;
    case RB200_I32: store_typed<T, int, V, false>(vw, off, istride, nvalid, val, mask); break;
    default: {
      Pack<T, V> pk;
#pragma unroll
      for (int k = 0; k < V; ++k) pk.v[k] = val[k];
      store_view_narrow<T, V>(vw, off, istride, nvalid, pk, mask);
    }
  }
}

template <class T, int V>
__device__ __noinline__ void store_view_narrow(const KView& vw, long long off, long long istride, int nvalid,
                                               Pack<T, V> pk, unsigned mask) {
  const T (&val)[V] = pk.v;
  switch (vw.dtype) {
    case RB200_BOOL: store_typed<T, unsigned char, V, true>(vw, off, istride, nvalid, val, mask); break;
    case RB200_U8: store_typed<T, unsigned char, V, false>(vw, off, istride, nvalid, val, mask); break;
    case RB200_I8: store_typed<T, signed char, V, false>(vw, off, istride, nvalid, val, mask); break;
    case RB200_I16: store_typed<T, short, V, false>(vw, off, istride, nvalid, val, mask); break;
    case RB200_U16: store_typed<T, unsigned short, V, false>(vw, off, istride, nvalid, val, mask); break;
    case RB200_U32: store_typed<T, unsigned int, V, false>(vw, off, istride, nvalid, val, mask); break;
    default: break;
  }
}

// ---------------------------------------------------------------------------------------------
// scalar op semantics

// Python floor division / modulo (what Numba emits for `//` and `%`)
__device__ __forceinline__ long long py_floordiv(long long a, long long b) {
  if (b == 0) return 0;
  long long q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
__device__ __forceinline__ long long py_mod(long long a, long long b) {
  if (b == 0) return 0;
  long long r = a % b;
  if (r != 0 && ((r < 0) != (b < 0))) r += b;
  return r;
}
template <class F> __device__ __forceinline__ F py_fmod(F a, F b) {
  F r = fmod(a, b);
  if (r != F(0)) {
    if ((b < F(0)) != (r < F(0))) r += b;
  } else {
    r = copysign(F(0), b);
  }
  return r;
}
template <class F> __device__ __forceinline__ F py_ffloordiv(F a, F b) {
  // CPython float_floor_div / Numba real_floordiv
  F mod = fmod(a, b);
  F div = (a - mod) / b;
  if (mod != F(0) && ((b < F(0)) != (mod < F(0)))) div -= F(1);
  if (div != F(0)) {
    F fl = floor(div);
    if (div - fl > F(0.5)) fl += F(1);
    return fl;
  }
  return copysign(F(0), a / b);
}

// x ** n for integer n: Numba's int_power_impl (exponentiation by squaring, r starts at 1)
template <class F> __device__ __forceinline__ F powi(F a, long long b) {
  bool invert = b < 0;
  unsigned long long e = invert ? (unsigned long long)(-b) : (unsigned long long)b;
  if (e > 0x10000ull) return (F)pow((double)a, (double)b);
  F r = F(1);
  while (e != 0) {
    if (e & 1ull) r *= a;
    e >>= 1;
    a *= a;
  }
  return invert ? F(1) / r : r;
}
__device__ __forceinline__ long long ipowi(long long a, long long b) {
  if (b < 0) return (a == 1) ? 1 : ((a == -1) ? ((b & 1) ? -1 : 1) : 0);
  long long r = 1;
  unsigned long long e = (unsigned long long)b;
  while (e != 0) {
    if (e & 1ull) r *= a;
    e >>= 1;
    a *= a;
  }
  return r;
}

}  // namespace rb200
