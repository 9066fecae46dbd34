// rb200_terms.h — the TERM form of a lean op list, shared by the stencil kernel (rb200_tile.cu) and the streaming kernel
// (rb200_stream.cu).
//
// The reference's generated loop body for the programs that matter at 4 bytes per element is one running value:
//   acc = U[o1+i] + U[o2+i] + ... ; V[i] = acc - c*U[..]                      (stencils, ramba/ramba.py:8146-8188)
//   tmp = X[i]*2.0 + 1.0 ; red = red + tmp                                    (map + reduce, ramba/ramba.py:5798-5807)
//   red[j] = red[j] + (M[i,j] + v[j])                                         (axis reduction, ramba/ramba.py:8231-8244)
// Such an op list - ONE running value updated by views and scalars, at most one float32 -> float64 promotion - is
// flattened on the host into TERMS.  Every arithmetic term is   p = x*w | x | w   (x: element of a staged or direct view,
// w: scalar; the product is rounded on its own) followed by   acc = p,  acc = (+-acc) + (+-p),  acc = acc * p   or
// acc = -acc;  the streaming kernel adds  acc = (double)(float)acc  (the value a float32 temporary would hold),
// "store acc to a view" and "fold acc into a reduction slot".  The kernels walk terms with no dispatch tree and the running
// value never leaves its registers.  Same operations, same order, same classes, one rounding each as in the op list
// (a - b is a + (-b) exactly), so the results are bit-identical to the general interpreter and the oracle.
#pragma once
#include <string.h>

#include "rb200_lean.cuh"

namespace rb200 {

enum TermKind { TK_SET = 0, TK_ADD = 1, TK_MUL = 2, TK_NEG = 3, TK_ROUND32 = 4, TK_STORE = 5, TK_RED = 6 };
enum TermX { X_NONE = 0, X_STAGED = 1, X_DIRECT = 2 };
enum TermFlags { TF_W = 1, TF_NEGP = 2, TF_NEGACC = 4 };
struct TermStep {  // 8 bytes, one constant-bank load
  unsigned char kind;   // TermKind
  unsigned char xkind;  // TermX
  unsigned char xidx;   // staged / direct index (TK_STORE: direct index of the destination)
  unsigned char sidx;   // scalar index (TF_W); TK_RED: reduction slot
  unsigned short off;   // stencil kernel: byte offset of a staged operand inside a plane
  unsigned char dzl;    // stencil kernel: plane of the ring; TK_RED: rb200_redop
  unsigned char flags;  // TermFlags
};
constexpr int kMaxTerms = 48;

struct TermBuild {
  int n_regs;
  const LDirect* direct;
  bool stream;  // streaming kernel: rounding steps, stores anywhere, reductions
  bool (*staged_fill)(void* ctx, int arg, int cls_f32, TermStep* t);  // kernel-specific fields of a staged operand
  void* ctx;
};

// Flatten translated lean instructions into terms.  false: not of the term form.  *out_view: direct index of the final
// store (stencil kernel; -1 when the op list ends otherwise).
static inline bool build_terms(const TermBuild& B, const LInsn* L, int n, TermStep* T, int max_terms, int* n_terms, int* n32_out, int* out_view) {
  if (B.n_regs != 0 || n < 1) return false;
  int nt = 0, cls = -1, n32 = -1;
  auto is_view = [](int k) { return k == L_STAGED || k == L_DIRECT; };
  auto raw = [&](const TermStep& t) -> bool {
    if (nt >= max_terms) return false;
    T[nt++] = t;
    return true;
  };
  // one term: p from (view operand, scalar operand) - either may be absent (kind L_NONE) but not both
  auto push = [&](int kind, int vkind, int varg, int skind, int sarg, int flags) -> bool {
    TermStep t;
    memset(&t, 0, sizeof(t));
    t.kind = (unsigned char)kind;
    t.flags = (unsigned char)flags;
    if (kind != TK_NEG) {
      if (vkind == L_STAGED) {
        t.xkind = X_STAGED;
        t.xidx = (unsigned char)varg;
        if (!B.staged_fill(B.ctx, varg, cls == 1, &t)) return false;
      } else if (vkind == L_DIRECT) {
        t.xkind = X_DIRECT;
        t.xidx = (unsigned char)varg;
      } else if (vkind != L_NONE) {
        return false;
      }
      if (skind == L_SCAL) {
        t.flags |= TF_W;
        t.sidx = (unsigned char)sarg;
      } else if (skind != L_NONE) {
        return false;
      }
      if (t.xkind == X_NONE && !(t.flags & TF_W)) return false;
    }
    return raw(t);
  };
  // p = one operand (view or scalar)
  auto one = [&](int kind, int okind, int oarg, int flags) -> bool {
    if (is_view(okind)) return push(kind, okind, oarg, L_NONE, 0, flags);
    if (okind == L_SCAL) return push(kind, L_NONE, 0, L_SCAL, oarg, flags);
    return false;
  };
  int last_store = -1;
  for (int i = 0; i < n; ++i) {
    const LInsn& I = L[i];
    const int lop = I.handler >> 2, f32 = (I.handler >> 1) & 1;
    const bool aacc = (I.handler & 1) != 0;
    if (I.st_reg != RB200_NOSTORE) return false;
    if (I.st_view != RB200_NOSTORE && i != n - 1 && !B.stream) return false;
    if (lop == LO_CVT) {
      if (!aacc || nt == 0) return false;
      if (f32 == 0) {  // float32 -> float64
        if (cls == 1 && n32 < 0) {  // the one promotion
          n32 = nt;
          cls = 0;
        } else {
          return false;
        }
      } else {  // float64 -> float32
        if (cls != 0) return false;
        // followed by the conversion back: the value a float32 temporary holds, still in the float64 phase
        if (B.stream && i + 1 < n && (L[i + 1].handler >> 2) == LO_CVT && ((L[i + 1].handler >> 1) & 1) == 0 && (L[i + 1].handler & 1) &&
            L[i + 1].st_reg == RB200_NOSTORE) {
          TermStep t;
          memset(&t, 0, sizeof(t));
          t.kind = TK_ROUND32;
          if (!raw(t)) return false;
          // stores of either instruction see the rounded value: float32 views get it exactly, float64 views too
          for (int q = i; q <= i + 1; ++q)
            if (L[q].st_view != RB200_NOSTORE) {
              TermStep st;
              memset(&st, 0, sizeof(st));
              st.kind = TK_STORE;
              st.xidx = L[q].st_view;
              if (!raw(st)) return false;
              last_store = L[q].st_view;
            }
          ++i;
          continue;
        }
        // otherwise only as the last instruction into a float32 view (the store converts)
        if (i != n - 1 || I.st_view == RB200_NOSTORE || B.direct[I.st_view].dtype != RB200_F32) return false;
        if (B.stream) {
          TermStep st;
          memset(&st, 0, sizeof(st));
          st.kind = TK_STORE;
          st.xidx = I.st_view;
          if (!raw(st)) return false;
        }
        last_store = I.st_view;
      }
      continue;
    }
    if (lop == LO_RED) {
      if (!B.stream || !aacc || nt == 0 || f32 != 0 || cls != 0 && !(cls == 1 && false)) return false;
      TermStep t;
      memset(&t, 0, sizeof(t));
      t.kind = TK_RED;
      t.sidx = I.b_arg;
      t.dzl = I.red_op;
      if (!raw(t)) return false;
      continue;
    }
    if (cls < 0) cls = f32;
    else if (cls != f32) return false;
    const bool fresh = nt == 0;  // no running value yet: the instruction may start from its own operands
    switch (lop) {
      case LO_MOV:
        if (!aacc && !(fresh && one(TK_SET, I.a_kind, I.a_arg, 0))) return false;
        break;
      case LO_NEG:
        if (!aacc && !(fresh && one(TK_SET, I.a_kind, I.a_arg, 0))) return false;
        if (!push(TK_NEG, L_NONE, 0, L_NONE, 0, 0)) return false;
        break;
      case LO_ADD:
      case LO_SUB:
      case LO_RSUB:
      case LO_MUL: {
        if (!aacc && !(fresh && one(TK_SET, I.a_kind, I.a_arg, 0))) return false;
        const int kind = lop == LO_MUL ? TK_MUL : TK_ADD;
        const int fl = lop == LO_SUB ? TF_NEGP : lop == LO_RSUB ? TF_NEGACC : 0;
        if (!one(kind, I.b_kind, I.b_arg, fl)) return false;
      } break;
      case LO_MULADD:
      case LO_MULSUB:
      case LO_MULRSUB: {
        // r = a + p, a - p, p - a  with p = b*c rounded first
        const int fl = lop == LO_MULADD ? 0 : lop == LO_MULSUB ? TF_NEGP : TF_NEGACC;
        if (aacc) {
          if (is_view(I.b_kind) && I.c_kind == L_SCAL) {
            if (!push(TK_ADD, I.b_kind, I.b_arg, L_SCAL, I.c_arg, fl)) return false;
          } else if (is_view(I.c_kind) && I.b_kind == L_SCAL) {
            if (!push(TK_ADD, I.c_kind, I.c_arg, L_SCAL, I.b_arg, fl)) return false;
          } else {
            return false;
          }
          break;
        }
        // the running value is (or becomes) the product; then `a` is folded in: a + p, a - p (= -p + a), p - a
        if (I.b_kind == L_ACC || I.c_kind == L_ACC) {
          const int ok = I.b_kind == L_ACC ? I.c_kind : I.b_kind, oa = I.b_kind == L_ACC ? I.c_arg : I.b_arg;
          if (!one(TK_MUL, ok, oa, 0)) return false;
        } else {
          if (!fresh) return false;
          if (is_view(I.b_kind) && I.c_kind == L_SCAL) {
            if (!push(TK_SET, I.b_kind, I.b_arg, L_SCAL, I.c_arg, 0)) return false;
          } else if (is_view(I.c_kind) && I.b_kind == L_SCAL) {
            if (!push(TK_SET, I.c_kind, I.c_arg, L_SCAL, I.b_arg, 0)) return false;
          } else {
            if (!one(TK_SET, I.b_kind, I.b_arg, 0) || !one(TK_MUL, I.c_kind, I.c_arg, 0)) return false;
          }
        }
        // now acc = p;  MULADD: a + p -> acc + a;  MULSUB: a - p -> (-acc) + a;  MULRSUB: p - a -> acc - a
        const int fl2 = lop == LO_MULADD ? 0 : lop == LO_MULSUB ? TF_NEGACC : TF_NEGP;
        if (!one(TK_ADD, I.a_kind, I.a_arg, fl2)) return false;
      } break;
      default: return false;
    }
    if (I.st_view != RB200_NOSTORE) {
      last_store = I.st_view;
      if (B.stream) {
        // (a float64 result stored to a float32 view is converted by the store; the running value keeps its class)
        TermStep st;
        memset(&st, 0, sizeof(st));
        st.kind = TK_STORE;
        st.xidx = I.st_view;
        if (!raw(st)) return false;
      }
    }
  }
  if (nt == 0) return false;
  if (!B.stream && (L[n - 1].st_view == RB200_NOSTORE)) return false;
  *out_view = last_store;
  *n_terms = nt;
  *n32_out = n32 >= 0 ? n32 : (cls == 1 ? nt : 0);
  return true;
}

}  // namespace rb200
