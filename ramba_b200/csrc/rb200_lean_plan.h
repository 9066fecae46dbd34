// rb200_lean_plan.h — host side of the lean op-list machine: which op lists qualify, and their translation into
// LInsn records (rb200_lean.cuh).  The translation is 1:1 - same operation order, same compute classes - so a lean
// kernel and the general interpreter produce identical bits.
#pragma once
#include <string.h>

#include "rb200_lean.cuh"

namespace rb200 {

// lean opcode of an op-list instruction (-1: not in the lean vocabulary)
static inline int lean_opcode(const rb200_fused_op* op, const rb200_insn& I) {
  switch (I.op) {
    case RB200_OP_MOV: return LO_MOV;
    case RB200_OP_ADD: return LO_ADD;
    case RB200_OP_SUB: return LO_SUB;
    case RB200_OP_MUL: return LO_MUL;
    case RB200_OP_DIV: return LO_DIV;
    case RB200_OP_NEG: return LO_NEG;
    case RB200_OP_ABS: return LO_ABS;
    case RB200_OP_SQUARE: return LO_SQUARE;
    case RB200_OP_MIN: return LO_MIN;
    case RB200_OP_MAX: return LO_MAX;
    case RB200_OP_MULADD: return LO_MULADD;
    case RB200_OP_MULSUB: return LO_MULSUB;
    case RB200_OP_MULRSUB: return LO_MULRSUB;
    case RB200_OP_RED: return LO_RED;
    case RB200_OP_POWI:  // x ** 2 with a scalar exponent is exactly x * x (Numba int_power)
      if (I.b_kind == RB200_K_SCAL && (long long)op->scalars[I.b_idx] == 2) return LO_SQUARE;
      return -1;
    case RB200_OP_CVT: {
      const int src = (int)(I.imm & 0xff);
      if ((I.imm >> 8) != 0) return -1;
      if (!(src == RB200_T_F64 || src == RB200_T_F32) || src == (int)I.ctype) return -1;
      return LO_CVT;
    }
    default: return -1;
  }
}

// float-arithmetic-only op list over float32/float64 views, no masks, no index operands
static inline bool lean_eligible(const rb200_fused_op* op, bool allow_red) {
  if (op->n_insns < 1) return false;
  for (int i = 0; i < op->n_views; ++i)
    if (op->views[i].dtype != RB200_F32 && op->views[i].dtype != RB200_F64) return false;
  for (int i = 0; i < op->n_insns; ++i) {
    const rb200_insn& I = op->insns[i];
    if (I.ctype != RB200_T_F64 && I.ctype != RB200_T_F32) return false;
    if (I.mask_reg != RB200_NOSTORE) return false;
    const int lop = lean_opcode(op, I);
    if (lop < 0) return false;
    if (lop == LO_RED && (!allow_red || I.ctype != RB200_T_F64)) return false;
    const uint8_t kinds[3] = {I.a_kind, I.b_kind, I.c_kind};
    for (int q = 0; q < 3; ++q) {
      if (lop == LO_RED && q == 1) continue;
      if (kinds[q] == RB200_K_IOTA) return false;
    }
    if (I.a_kind == RB200_K_NONE) return false;
  }
  return true;
}

// view_kind[v] / view_arg[v]: how operand reads of view v are served (L_STAGED + staged-operand index, or L_DIRECT +
// direct-view index); stores always use dview_of_store[v] (index into the direct-view table)
static inline void lean_translate(const rb200_fused_op* op, const int* view_kind, const int* view_arg, const int* store_arg, LInsn* out) {
  for (int i = 0; i < op->n_insns; ++i) {
    const rb200_insn& I = op->insns[i];
    LInsn L;
    memset(&L, 0, sizeof(L));
    int lop = lean_opcode(op, I);
    uint8_t kinds[3] = {I.a_kind, I.b_kind, I.c_kind};
    uint8_t idxs[3] = {I.a_idx, I.b_idx, I.c_idx};
    if (lop == LO_SQUARE || lop == LO_RED) kinds[1] = RB200_K_NONE;  // (POWI's exponent / RED's slot are not operands)
    if ((lop == LO_ADD || lop == LO_MUL) && kinds[0] != RB200_K_ACC && kinds[1] == RB200_K_ACC) {
      kinds[1] = kinds[0]; idxs[1] = idxs[0];
      kinds[0] = RB200_K_ACC; idxs[0] = 0;
    } else if (lop == LO_SUB && kinds[0] != RB200_K_ACC && kinds[1] == RB200_K_ACC) {
      lop = LO_RSUB;  // b - a with a = the accumulator
      kinds[1] = kinds[0]; idxs[1] = idxs[0];
      kinds[0] = RB200_K_ACC; idxs[0] = 0;
    }
    unsigned char lk[3], la[3];
    for (int q = 0; q < 3; ++q) {
      switch (kinds[q]) {
        case RB200_K_ACC: lk[q] = L_ACC; la[q] = 0; break;
        case RB200_K_REG: lk[q] = L_REG; la[q] = idxs[q]; break;
        case RB200_K_SCAL: lk[q] = L_SCAL; la[q] = idxs[q]; break;
        case RB200_K_VIEW: lk[q] = (unsigned char)view_kind[idxs[q]]; la[q] = (unsigned char)view_arg[idxs[q]]; break;
        default: lk[q] = L_NONE; la[q] = 0;
      }
    }
    L.a_kind = lk[0]; L.a_arg = la[0];
    L.b_kind = lk[1]; L.b_arg = la[1];
    L.c_kind = lk[2]; L.c_arg = la[2];
    if (lop == LO_RED) {
      L.b_arg = I.b_idx;  // reduction slot
      L.red_op = (unsigned char)I.imm;
    }
    L.st_reg = I.st_reg;
    L.st_view = I.st_view == RB200_NOSTORE ? (unsigned char)RB200_NOSTORE : (unsigned char)store_arg[I.st_view];
    L.handler = (unsigned char)(lop * 4 + (I.ctype == RB200_T_F32 ? 2 : 0) + (L.a_kind == L_ACC ? 1 : 0));
    out[i] = L;
  }
}

// Fuse runs of add / sub / mul instructions whose right operand is a STAGED view and whose left operand is the running
// value into LO_CHAIN instructions (rb200_lean.cuh).  `insns` is rewritten in place; returns the new instruction count
// and fills `chain` (at most max_chain steps).
static inline int lean_fuse_chains(LInsn* insns, int n, LChainStep* chain, int max_chain, int* n_chain_out) {
  int out = 0, nc = 0;
  int i = 0;
  auto chain_op = [](const LInsn& L) -> int {
    const int lop = L.handler >> 2;
    return lop == LO_ADD ? LC_ADD : lop == LO_SUB ? LC_SUB : lop == LO_RSUB ? LC_RSUB : lop == LO_MUL ? LC_MUL : -1;
  };
  while (i < n) {
    const LInsn first = insns[i];
    int j = i;
    if (chain_op(first) >= 0 && first.b_kind == L_STAGED) {
      const int cls = (first.handler >> 1) & 1;
      j = i + 1;
      while (j < n && insns[j - 1].st_reg == RB200_NOSTORE && insns[j - 1].st_view == RB200_NOSTORE && chain_op(insns[j]) >= 0 &&
             insns[j].b_kind == L_STAGED && insns[j].a_kind == L_ACC && ((insns[j].handler >> 1) & 1) == cls && nc + (j - i) + 1 <= max_chain && (j - i) < 200)
        ++j;
      if (j - i >= 2) {
        LInsn L = first;
        L.handler = (unsigned char)(LO_CHAIN * 4 + cls * 2 + (first.a_kind == L_ACC ? 1 : 0));
        L.b_kind = L_NONE;
        L.b_arg = (unsigned char)nc;
        L.c_arg = (unsigned char)(j - i);
        L.st_reg = insns[j - 1].st_reg;
        L.st_view = insns[j - 1].st_view;
        for (int q = i; q < j; ++q) {
          chain[nc].op = (unsigned char)chain_op(insns[q]);
          chain[nc].staged = insns[q].b_arg;
          ++nc;
        }
        insns[out++] = L;
        i = j;
        continue;
      }
    }
    insns[out++] = first;
    ++i;
  }
  *n_chain_out = nc;
  return out;
}

}  // namespace rb200
