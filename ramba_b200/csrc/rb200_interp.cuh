#pragma once
// rb200_interp.cuh — the op-list interpreter shared by the elementwise and axis-reduction kernels.
//
// Reference behaviour restated (not translated): the generated Numba loop of
// ramba/ramba.py:8247-8265 executed by RemoteState.run_deferred_ops (ramba/ramba.py:3758-3780).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>

#include "rb200_vm.cuh"

namespace rb200 {

// ---------------------------------------------------------------------------------------------
// reduction combine in the accumulator class
template <class T> __device__ __forceinline__ T red_combine(int op, T a, T b) {
  switch (op) {
    case RB200_RED_ADD: return a + b;
    case RB200_RED_MUL: return a * b;
    case RB200_RED_MIN: return (b < a) ? b : a;
    default: return (b > a) ? b : a;
  }
}
__device__ __forceinline__ Val red_combine_val(int op, int ctype, Val a, Val b) {
  Val r;
  if (ctype == RB200_T_F64) r.d = red_combine<double>(op, a.d, b.d);
  else r.i = red_combine<long long>(op, a.i, b.i);
  return r;
}
__device__ __forceinline__ Val red_identity(int op, int ctype) {
  Val r;
  if (ctype == RB200_T_F64) {
    r.d = (op == RB200_RED_ADD) ? 0.0 : (op == RB200_RED_MUL) ? 1.0 : (op == RB200_RED_MIN) ? INFINITY : -INFINITY;
  } else {
    r.i = (op == RB200_RED_ADD) ? 0ll : (op == RB200_RED_MUL) ? 1ll : (op == RB200_RED_MIN) ? 0x7fffffffffffffffll
                                                                                             : (long long)0x8000000000000000ull;
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// per-thread interpreter state
template <int V> struct Ctx {
  const KParams& P;
  unsigned long long* regfile;  // [reg][k][thread]
  long long idx[kMaxD];         // current index (innermost = first element of this thread's chunk)
  int nvalid;
  Val acc[V];
  int tid;
  __device__ __forceinline__ Ctx(const KParams& p, unsigned long long* rf) : P(p), regfile(rf) { tid = threadIdx.x; }

  __device__ __forceinline__ long long view_off(const KView& vw) const {
    long long off = 0;
#pragma unroll
    for (int d = 0; d < kMaxD; ++d)
      if (d < P.ndim) off += idx[d] * vw.stride[d];
    return off;
  }
  __device__ __forceinline__ long long inner_stride(const KView& vw) const {
    long long s = 0;
#pragma unroll
    for (int d = 0; d < kMaxD; ++d)
      if (d == P.ndim - 1) s = vw.stride[d];
    return s;
  }

  template <class T> __device__ __forceinline__ void fetch(int kind, int i, T (&out)[V]) {
    switch (kind) {
      case RB200_K_ACC:
#pragma unroll
        for (int k = 0; k < V; ++k) out[k] = CT<T>::get(acc[k]);
        break;
      case RB200_K_REG:
#pragma unroll
        for (int k = 0; k < V; ++k) {
          Val v;
          v.u = regfile[(i * V + k) * kThreads + tid];
          out[k] = CT<T>::get(v);
        }
        break;
      case RB200_K_VIEW: {
        const KView& vw = P.views[i];
        load_view<T, V>(vw, view_off(vw), inner_stride(vw), nvalid, out);
      } break;
      case RB200_K_SCAL: {
        T s = CT<T>::scal(P.scalars[i]);
#pragma unroll
        for (int k = 0; k < V; ++k) out[k] = s;
      } break;
      case RB200_K_IOTA: {
        long long base = 0;
        bool inner = (i == P.ndim - 1);
#pragma unroll
        for (int d = 0; d < kMaxD; ++d)
          if (d == i) base = idx[d] + P.gstart[d];
#pragma unroll
        for (int k = 0; k < V; ++k) out[k] = (T)(base + (inner ? k : 0));
      } break;
      default:
#pragma unroll
        for (int k = 0; k < V; ++k) out[k] = T(0);
    }
  }

  __device__ __forceinline__ unsigned store_mask(const rb200_insn& I) const {
    if (I.mask_reg == RB200_NOSTORE) return (1u << V) - 1u;
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < V; ++k)
      if (regfile[(I.mask_reg * V + k) * kThreads + tid] != 0ull) m |= (1u << k);
    return m;
  }

  template <class R> __device__ __forceinline__ void finish(const rb200_insn& I, const R (&r)[V]) {
#pragma unroll
    for (int k = 0; k < V; ++k) CT<R>::set(acc[k], r[k]);
    if (I.st_reg != RB200_NOSTORE) {
#pragma unroll
      for (int k = 0; k < V; ++k) regfile[(I.st_reg * V + k) * kThreads + tid] = acc[k].u;
    }
    if (I.st_view != RB200_NOSTORE) {
      const KView& vw = P.views[I.st_view];
      store_view<R, V>(vw, view_off(vw), inner_stride(vw), nvalid, r, store_mask(I));
    }
  }
};

// ---------------------------------------------------------------------------------------------
// floating-point instruction set (F = double | float)
template <class F, int V, class C> __device__ __forceinline__ bool exec_float(C& cx, const rb200_insn& I) {
  F a[V], b[V], c[V], r[V];
  long long p[V];
  const int op = I.op;
  // operands are fetched once, up front (three inlined fetch sites per compute class)
  cx.template fetch<F>(I.a_kind, I.a_idx, a);
  if (I.b_kind != RB200_K_NONE && op != RB200_OP_POWI) cx.template fetch<F>(I.b_kind, I.b_idx, b);
  if (I.c_kind != RB200_K_NONE) cx.template fetch<F>(I.c_kind, I.c_idx, c);
  switch (op) {
    // ---- binary arithmetic
    case RB200_OP_ADD:
    case RB200_OP_SUB:
    case RB200_OP_MUL:
    case RB200_OP_DIV:
    case RB200_OP_MIN:
    case RB200_OP_MAX: {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        F x = a[k], y = b[k];
        // __d/f*_rn intrinsics: no FMA contraction across op-list instructions, every op rounds
        // once like the reference's un-fused scalar statements
        if constexpr (sizeof(F) == 8) {
          r[k] = op == RB200_OP_ADD   ? __dadd_rn(x, y)
                 : op == RB200_OP_SUB ? __dsub_rn(x, y)
                 : op == RB200_OP_MUL ? __dmul_rn(x, y)
                 : op == RB200_OP_DIV ? __ddiv_rn(x, y)
                 : op == RB200_OP_MIN ? ((y < x) ? y : x)
                                      : ((y > x) ? y : x);
        } else {
          r[k] = op == RB200_OP_ADD   ? __fadd_rn(x, y)
                 : op == RB200_OP_SUB ? __fsub_rn(x, y)
                 : op == RB200_OP_MUL ? __fmul_rn(x, y)
                 : op == RB200_OP_DIV ? __fdiv_rn(x, y)
                 : op == RB200_OP_MIN ? ((y < x) ? y : x)
                                      : ((y > x) ? y : x);
        }
      }
      cx.template finish<F>(I, r);
      return true;
    }
    case RB200_OP_FLOORDIV:
    case RB200_OP_MOD:
    case RB200_OP_POW: {
#pragma unroll
      for (int k = 0; k < V; ++k)
        r[k] = op == RB200_OP_FLOORDIV ? py_ffloordiv<F>(a[k], b[k]) : op == RB200_OP_MOD ? py_fmod<F>(a[k], b[k]) : (F)pow(a[k], b[k]);
      cx.template finish<F>(I, r);
      return true;
    }
    case RB200_OP_POWI: {
      long long e[V];
      cx.template fetch<long long>(I.b_kind, I.b_idx, e);
      if (e[0] == 2) {  // uniform scalar exponent in practice; x**2 == x*x exactly (int_power)
#pragma unroll
        for (int k = 0; k < V; ++k) r[k] = (e[k] == 2) ? a[k] * a[k] : powi<F>(a[k], e[k]);
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) r[k] = powi<F>(a[k], e[k]);
      }
      cx.template finish<F>(I, r);
      return true;
    }
    // ---- comparisons -> bool (I64 class 0/1)
    case RB200_OP_GT:
    case RB200_OP_LT:
    case RB200_OP_GE:
    case RB200_OP_LE:
    case RB200_OP_EQ:
    case RB200_OP_NE: {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        F x = a[k], y = b[k];
        bool t = op == RB200_OP_GT ? x > y : op == RB200_OP_LT ? x < y : op == RB200_OP_GE ? x >= y : op == RB200_OP_LE ? x <= y : op == RB200_OP_EQ ? x == y : x != y;
        p[k] = t ? 1 : 0;
      }
      cx.template finish<long long>(I, p);
      return true;
    }
    case RB200_OP_LAND:
    case RB200_OP_LOR:
    case RB200_OP_LXOR: {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        bool x = a[k] != F(0), y = b[k] != F(0);
        p[k] = (op == RB200_OP_LAND ? (x && y) : op == RB200_OP_LOR ? (x || y) : (x != y)) ? 1 : 0;
      }
      cx.template finish<long long>(I, p);
      return true;
    }
    case RB200_OP_ISFINITE:
    case RB200_OP_ISINF:
    case RB200_OP_ISNAN:
    case RB200_OP_ISNEGINF:
    case RB200_OP_ISPOSINF:
    case RB200_OP_LNOT: {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        F x = a[k];
        bool t = op == RB200_OP_ISFINITE ? isfinite(x)
                 : op == RB200_OP_ISINF  ? isinf(x)
                 : op == RB200_OP_ISNAN  ? isnan(x)
                 : op == RB200_OP_ISNEGINF ? (isinf(x) && x < F(0))
                 : op == RB200_OP_ISPOSINF ? (isinf(x) && x > F(0))
                                           : (x == F(0));
        p[k] = t ? 1 : 0;
      }
      cx.template finish<long long>(I, p);
      return true;
    }
    // ---- unary
    case RB200_OP_MOV:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = a[k];
      cx.template finish<F>(I, r);
      return true;
    case RB200_OP_ABS:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = fabs(a[k]);
      cx.template finish<F>(I, r);
      return true;
    case RB200_OP_NEG:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = -a[k];
      cx.template finish<F>(I, r);
      return true;
    case RB200_OP_SQUARE:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = a[k] * a[k];
      cx.template finish<F>(I, r);
      return true;
    case RB200_OP_SQRT:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = sqrt(a[k]);
      cx.template finish<F>(I, r);
      return true;
    case RB200_OP_SIN:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = sin(a[k]);
      cx.template finish<F>(I, r);
      return true;
    case RB200_OP_COS:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = cos(a[k]);
      cx.template finish<F>(I, r);
      return true;
    case RB200_OP_SINCOS: {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        F sn, cs;
        sincos(a[k], &sn, &cs);
        r[k] = I.imm ? cs : sn;  // imm 1: accumulator half is cos, parked half is sin
        c[k] = I.imm ? sn : cs;
      }
#pragma unroll
      for (int k = 0; k < V; ++k) {
        Val v;
        CT<F>::set(v, c[k]);
        cx.regfile[(I.st2 * V + k) * kThreads + cx.tid] = v.u;
      }
      cx.template finish<F>(I, r);
      return true;
    }
    case RB200_OP_TAN:
    case RB200_OP_SINH:
    case RB200_OP_COSH:
    case RB200_OP_TANH:
    case RB200_OP_ASIN:
    case RB200_OP_ACOS:
    case RB200_OP_ATAN:
    case RB200_OP_EXP:
    case RB200_OP_LOG:
    case RB200_OP_CBRT: {
      // rarely on the hot path: one element at a time keeps the code small
      for (int k = 0; k < V; ++k) {
        F x = a[k];
        r[k] = op == RB200_OP_TAN    ? tan(x)
               : op == RB200_OP_SINH ? sinh(x)
               : op == RB200_OP_COSH ? cosh(x)
               : op == RB200_OP_TANH ? tanh(x)
               : op == RB200_OP_ASIN ? asin(x)
               : op == RB200_OP_ACOS ? acos(x)
               : op == RB200_OP_ATAN ? atan(x)
               : op == RB200_OP_EXP  ? exp(x)
               : op == RB200_OP_LOG  ? log(x)
                                     : cbrt(x);
      }
      cx.template finish<F>(I, r);
      return true;
    }
    case RB200_OP_WHERE: {
      // condition arrives in `a`, already converted to the compute class (non-zero = true)
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = (a[k] != F(0)) ? b[k] : c[k];
      cx.template finish<F>(I, r);
      return true;
    }
    default: return false;
  }
}

// integer instruction set (all integer arithmetic is int64, like Numba's intp promotion)
template <int V, class C> __device__ __forceinline__ bool exec_int(C& cx, const rb200_insn& I) {
  long long a[V], b[V], c[V], r[V];
  cx.template fetch<long long>(I.a_kind, I.a_idx, a);
  if (I.b_kind != RB200_K_NONE) cx.template fetch<long long>(I.b_kind, I.b_idx, b);
  if (I.c_kind != RB200_K_NONE) cx.template fetch<long long>(I.c_kind, I.c_idx, c);
  const int op = I.op;
  switch (op) {
    case RB200_OP_MOV:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = a[k];
      break;
    case RB200_OP_ADD:
    case RB200_OP_SUB:
    case RB200_OP_MUL:
    case RB200_OP_MIN:
    case RB200_OP_MAX:
    case RB200_OP_BAND:
    case RB200_OP_BOR:
    case RB200_OP_BXOR:
    case RB200_OP_SHL:
    case RB200_OP_SHR:
#pragma unroll
      for (int k = 0; k < V; ++k) {
        long long x = a[k], y = b[k];
        r[k] = op == RB200_OP_ADD   ? x + y
               : op == RB200_OP_SUB ? x - y
               : op == RB200_OP_MUL ? x * y
               : op == RB200_OP_MIN ? ((y < x) ? y : x)
               : op == RB200_OP_MAX ? ((y > x) ? y : x)
               : op == RB200_OP_BAND ? (x & y)
               : op == RB200_OP_BOR  ? (x | y)
               : op == RB200_OP_BXOR ? (x ^ y)
               : op == RB200_OP_SHL  ? (long long)((unsigned long long)x << (y & 63))
                                     : (x >> (y & 63));
      }
      break;
    case RB200_OP_FLOORDIV:
      for (int k = 0; k < V; ++k) r[k] = py_floordiv(a[k], b[k]);
      break;
    case RB200_OP_MOD:
      for (int k = 0; k < V; ++k) r[k] = py_mod(a[k], b[k]);
      break;
    case RB200_OP_POWI:
    case RB200_OP_POW:
      for (int k = 0; k < V; ++k) r[k] = ipowi(a[k], b[k]);
      break;
    case RB200_OP_GT:
    case RB200_OP_LT:
    case RB200_OP_GE:
    case RB200_OP_LE:
    case RB200_OP_EQ:
    case RB200_OP_NE:
    case RB200_OP_LAND:
    case RB200_OP_LOR:
    case RB200_OP_LXOR:
#pragma unroll
      for (int k = 0; k < V; ++k) {
        long long x = a[k], y = b[k];
        bool t = op == RB200_OP_GT   ? x > y
                 : op == RB200_OP_LT ? x < y
                 : op == RB200_OP_GE ? x >= y
                 : op == RB200_OP_LE ? x <= y
                 : op == RB200_OP_EQ ? x == y
                 : op == RB200_OP_NE ? x != y
                 : op == RB200_OP_LAND ? (x != 0 && y != 0)
                 : op == RB200_OP_LOR  ? (x != 0 || y != 0)
                                       : ((x != 0) != (y != 0));
        r[k] = t ? 1 : 0;
      }
      break;
    case RB200_OP_ABS:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = a[k] < 0 ? -a[k] : a[k];
      break;
    case RB200_OP_NEG:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = -a[k];
      break;
    case RB200_OP_SQUARE:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = a[k] * a[k];
      break;
    case RB200_OP_INVERT:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = (I.imm == 1) ? (a[k] == 0 ? 1 : 0) : ~a[k];  // imm 1: bool operand
      break;
    case RB200_OP_LNOT:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = a[k] == 0 ? 1 : 0;
      break;
    case RB200_OP_ISFINITE:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = 1;
      break;
    case RB200_OP_ISINF:
    case RB200_OP_ISNAN:
    case RB200_OP_ISNEGINF:
    case RB200_OP_ISPOSINF:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = 0;
      break;
    case RB200_OP_WHERE:
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = (a[k] != 0) ? b[k] : c[k];
      break;
    default: return false;
  }
  cx.template finish<long long>(I, r);
  return true;
}

// value a store + reload through storage dtype `dt` would give (narrowing round / wrap)
template <class T> __device__ __forceinline__ T through_storage(T x, int dt) {
  switch (dt) {
    case RB200_F64: return (T)(double)x;
    case RB200_F32: return (T)(float)x;
    case RB200_I64: return (T)(long long)x;
    case RB200_I32: return (T)(int)x;
    case RB200_BOOL: return (T)(x != T(0) ? 1 : 0);
    case RB200_U8: return (T)(unsigned char)(long long)x;
    case RB200_I8: return (T)(signed char)(long long)x;
    case RB200_I16: return (T)(short)(long long)x;
    case RB200_U16: return (T)(unsigned short)(long long)x;
    case RB200_U32: return (T)(unsigned int)(long long)x;
    default: return x;
  }
}

template <class S, int V, class C> __device__ __forceinline__ void exec_cvt_from(C& cx, const rb200_insn& I) {
  S a[V];
  cx.template fetch<S>(I.a_kind, I.a_idx, a);
  const int through = (int)(I.imm >> 8);
  if (through != 0) {
#pragma unroll
    for (int k = 0; k < V; ++k) a[k] = through_storage<S>(a[k], through - 1);
  }
  switch (I.ctype) {
    case RB200_T_F64: {
      double r[V];
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = (double)a[k];
      cx.template finish<double>(I, r);
    } break;
    case RB200_T_F32: {
      float r[V];
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = (float)a[k];
      cx.template finish<float>(I, r);
    } break;
    default: {
      long long r[V];
#pragma unroll
      for (int k = 0; k < V; ++k) r[k] = (long long)a[k];
      cx.template finish<long long>(I, r);
    }
  }
}

// one interpreter pass over the op list for the current V elements.
// RACC: reduction accumulators, [slot] (global mode, AX=false) or [slot][k] (axis mode)
template <int V, bool AX, class C> __device__ __forceinline__ void run_program(C& cx, Val (&racc)[RB200_MAX_REDS][AX ? V : 1]) {
  const KParams& P = cx.P;
  for (int pc = 0; pc < P.n_insns; ++pc) {
    const rb200_insn I = P.insns[pc];
    if (I.op == RB200_OP_CVT) {
      switch (I.imm & 0xff) {
        case RB200_T_F64: exec_cvt_from<double, V>(cx, I); break;
        case RB200_T_F32: exec_cvt_from<float, V>(cx, I); break;
        default: exec_cvt_from<long long, V>(cx, I);
      }
      continue;
    }
    if (I.op == RB200_OP_RED) {
      const int slot = I.b_idx;
      const int rop = (int)I.imm;
      if (I.ctype == RB200_T_F64) {
        double a[V];
        cx.template fetch<double>(I.a_kind, I.a_idx, a);
#pragma unroll
        for (int s = 0; s < RB200_MAX_REDS; ++s)
          if (s == slot) {
#pragma unroll
            for (int k = 0; k < V; ++k)
              if (k < cx.nvalid) {
                Val& t = racc[s][AX ? k : 0];
                t.d = red_combine<double>(rop, t.d, a[k]);
              }
          }
      } else {
        long long a[V];
        cx.template fetch<long long>(I.a_kind, I.a_idx, a);
#pragma unroll
        for (int s = 0; s < RB200_MAX_REDS; ++s)
          if (s == slot) {
#pragma unroll
            for (int k = 0; k < V; ++k)
              if (k < cx.nvalid) {
                Val& t = racc[s][AX ? k : 0];
                t.i = red_combine<long long>(rop, t.i, a[k]);
              }
          }
      }
      continue;
    }
    switch (I.ctype) {
      case RB200_T_F64: exec_float<double, V>(cx, I); break;
      case RB200_T_F32: exec_float<float, V>(cx, I); break;
      default: exec_int<V>(cx, I);
    }
  }
}

}  // namespace rb200
