/*
 * ramba_b200.h — C-ABI of libramba_b200.so (sm_100a), the execution seam of the
 * fused elementwise / reduction / shifted-slice hot path.
 *
 * What it replaces in the reference (Python-for-HPC/ramba, all file:line relative to
 * the reference tree):
 *
 *   rb200_run_deferred_ops   <- RemoteState.run_deferred_ops kernel launch,
 *                               ramba/ramba.py:3758-3780: one call runs one fused
 *                               op over ONE iteration range of ONE worker, with the
 *                               generated-kernel signature
 *                               f(global_start, itershape, worker_num, num_workers,
 *                                 *array_views, *scalars)   (ramba/ramba.py:8262, 8265).
 *                               The Python-source kernel body (ramba/ramba.py:8247-8255)
 *                               becomes the op-list `insns`; the per-view NumPy views
 *                               (shardview.array_to_view, ramba/shardview_array.py:557-614)
 *                               become `rb200_view` stride descriptors; pickled scalars
 *                               (ramba/ramba.py:3661-3666) become `scalars`; pre/postcode
 *                               of global reductions (ramba/ramba.py:5798-5807) and the
 *                               axis-reduction loop nest (ramba/ramba.py:8231-8244)
 *                               become the `red_*` fields.
 *   rb200_reduce_partials    <- stage 2 of an axis reduction over partial slices,
 *                               ndarray.internal_reduction2_executor, ramba/ramba.py:5818-5849.
 *   rb200_cumulative         <- RemoteState.scumulative_worker, ramba/ramba.py:3378-3437 (cumsum / scumulative).
 *   rb200_last_error         <- worker exception -> ("ERROR", worker, traceback) reply,
 *                               ramba/ramba.py:3875-3881.
 *
 * Ownership: every device pointer is BORROWED for the duration of one call (the
 * Python side owns shards as torch tensors keyed by gid, like the worker's
 * numpy_map, ramba/ramba.py:1898, 2005). Launches are asynchronous on `stream`.
 * All entry points return 0 on success, non-zero on error (see rb200_last_error).
 * There is no CPU fallback: without a CUDA device every launch fails.
 */
#ifndef RAMBA_B200_H
#define RAMBA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB200_ABI_VERSION 4

#define RB200_MAX_DIMS 5     /* iteration dims after host-side collapsing            */
#define RB200_MAX_VIEWS 16   /* distinct array views per fused op                    */
#define RB200_MAX_SCALARS 32 /* scalar table entries                                 */
#define RB200_MAX_INSNS 96   /* op-list length                                       */
#define RB200_MAX_REGS 12    /* spill registers of the accumulator machine           */
#define RB200_MAX_REDS 4     /* global reductions fused into one op                  */

/* storage dtypes of array views */
enum rb200_dtype {
  RB200_F64 = 0,
  RB200_F32 = 1,
  RB200_I64 = 2,
  RB200_I32 = 3,
  RB200_BOOL = 4, /* 1 byte, 0/1 */
  RB200_U8 = 5,
  RB200_I8 = 6,
  RB200_I16 = 7,
  RB200_U16 = 8,
  RB200_U32 = 9,
  RB200_NUM_DTYPES = 10
};

/* compute classes of the op-list machine (what Numba's scalar typing gives the
 * reference's generated loop body) */
enum rb200_ctype { RB200_T_F64 = 0, RB200_T_F32 = 1, RB200_T_I64 = 2 };

/* operand kinds */
enum rb200_kind {
  RB200_K_NONE = 0,
  RB200_K_ACC = 1,  /* result of the previous instruction                           */
  RB200_K_REG = 2,  /* spill register idx                                           */
  RB200_K_VIEW = 3, /* element of views[idx] at the current index                   */
  RB200_K_SCAL = 4, /* scalars[idx] (already in the instruction's compute class)    */
  RB200_K_IOTA = 5  /* index[idx] + global_start[idx]  (ramba/ramba.py:8955-8960)   */
};

/* opcodes; vocabulary = the reference's op tables, ramba/ramba.py:7893-7993 */
enum rb200_op {
  RB200_OP_MOV = 0,
  RB200_OP_ADD = 1,
  RB200_OP_SUB = 2,
  RB200_OP_MUL = 3,
  RB200_OP_DIV = 4,      /* true division                                         */
  RB200_OP_FLOORDIV = 5, /* Python semantics                                      */
  RB200_OP_MOD = 6,      /* Python semantics                                      */
  RB200_OP_POW = 7,      /* float ** float                                        */
  RB200_OP_POWI = 8,     /* x ** int64 by repeated squaring (Numba int_power)     */
  RB200_OP_MIN = 9,      /* builtins.min(a,b)                                     */
  RB200_OP_MAX = 10,     /* builtins.max(a,b)                                     */
  RB200_OP_GT = 11,
  RB200_OP_LT = 12,
  RB200_OP_GE = 13,
  RB200_OP_LE = 14,
  RB200_OP_EQ = 15,
  RB200_OP_NE = 16,
  RB200_OP_LAND = 17,
  RB200_OP_LOR = 18,
  RB200_OP_LXOR = 19,
  RB200_OP_BAND = 20,
  RB200_OP_BOR = 21,
  RB200_OP_BXOR = 22,
  RB200_OP_SHL = 23,
  RB200_OP_SHR = 24,
  RB200_OP_ABS = 25,
  RB200_OP_SQUARE = 26,
  RB200_OP_SQRT = 27,
  RB200_OP_SIN = 28,
  RB200_OP_COS = 29,
  RB200_OP_TAN = 30,
  RB200_OP_SINH = 31,
  RB200_OP_COSH = 32,
  RB200_OP_TANH = 33,
  RB200_OP_ASIN = 34,
  RB200_OP_ACOS = 35,
  RB200_OP_ATAN = 36,
  RB200_OP_NEG = 37,
  RB200_OP_EXP = 38,
  RB200_OP_LOG = 39,
  RB200_OP_ISFINITE = 40,
  RB200_OP_ISINF = 41,
  RB200_OP_ISNAN = 42,
  RB200_OP_ISNEGINF = 43,
  RB200_OP_ISPOSINF = 44,
  RB200_OP_LNOT = 45,
  RB200_OP_INVERT = 46,
  RB200_OP_WHERE = 47,  /* a ? b : c  (a is tested != 0 in the compute class)      */
  RB200_OP_CVT = 48,    /* convert a from class `imm & 0xff` to `ctype`; if
                           (imm >> 8) != 0 first wrap/round through storage dtype
                           ((imm >> 8) - 1), i.e. the value a store + reload yields */
  RB200_OP_SINCOS = 49, /* acc = sin(a), regs[st2] = cos(a) (imm 1: swapped); one range
                           reduction for both.  If c_kind == RB200_K_VIEW the parked
                           half is also stored to views[c_idx]                       */
  RB200_OP_RED = 50,    /* red[b_idx] = red[b_idx] (+,*,min,max by imm) a          */
  RB200_OP_CBRT = 51,
  /* three-operand forms of `a (+/-) b*c`: the product and the sum are rounded SEPARATELY (no FMA), exactly like the
     two statements of the reference's loop body they stand for; they exist so that a weighted term of a stencil
     (`... - 6.0*U[...]`, ramba/ramba.py:8146-8188) does not need a spill register                                */
  RB200_OP_MULADD = 52,  /* a + b*c                                                */
  RB200_OP_MULSUB = 53,  /* a - b*c                                                */
  RB200_OP_MULRSUB = 54, /* b*c - a                                                */
  RB200_NUM_OPS = 55
};

enum rb200_redop { RB200_RED_ADD = 0, RB200_RED_MUL = 1, RB200_RED_MIN = 2, RB200_RED_MAX = 3 };

#define RB200_NOSTORE 0xff

/* one op-list instruction, 16 bytes */
typedef struct rb200_insn {
  uint8_t op;      /* rb200_op                                                     */
  uint8_t ctype;   /* compute class operands are fetched in / op is evaluated in   */
  uint8_t a_kind, a_idx;
  uint8_t b_kind, b_idx;
  uint8_t c_kind, c_idx;
  uint8_t st_reg;  /* also copy the result into spill register (RB200_NOSTORE: no) */
  uint8_t st_view; /* also store the result into views[st_view] (converted to its
                      dtype)                                                       */
  uint8_t st2;     /* second result register (SINCOS)                              */
  uint8_t mask_reg;/* RB200_NOSTORE, or spill register holding the write mask of a
                      masked store  (`if mask[index]:` guard, ramba/ramba.py:8476) */
  uint32_t imm;
} rb200_insn;

/* one array view bound to one iteration range: element (i0..ik) of the range lives
 * at base + sum(i_d * stride[d]) elements.  stride 0 = broadcast axis
 * (axis_map == -1, ramba/shardview_array.py:36).                                  */
typedef struct rb200_view {
  void* base;
  int64_t stride[RB200_MAX_DIMS];
  int32_t dtype; /* rb200_dtype */
  int32_t flags; /* bit0: written by this op                                       */
  /* [alloc_lo, alloc_hi): the device buffer `base` points into (the worker's shard, LocalNdarray.bcontainer,
   * ramba/ramba.py:1208-1214).  Optional (both NULL = unknown).  The stencil kernel stages a tile plus its halo
   * with whole-box TMA copies only when the box lies inside these bounds.                                       */
  const void* alloc_lo;
  const void* alloc_hi;
} rb200_view;

typedef struct rb200_red {
  int32_t op;    /* rb200_redop                                                    */
  int32_t ctype; /* accumulator class: RB200_T_F64 or RB200_T_I64                  */
  void* out;     /* device pointer to this worker's element of the partial array
                    (red[0,..] = red[0,..] (op) acc, ramba/ramba.py:5805-5806)       */
  int32_t out_dtype;
  int32_t pad;
} rb200_red;

/* One fused op over one iteration range of one worker. */
typedef struct rb200_fused_op {
  int32_t abi_version; /* RB200_ABI_VERSION                                        */
  int32_t ndim;        /* 1..RB200_MAX_DIMS (collapsed iteration space)            */
  int64_t itershape[RB200_MAX_DIMS];
  int64_t global_start[RB200_MAX_DIMS]; /* added to IOTA operands                   */
  int32_t iota_dim[RB200_MAX_DIMS];     /* unused (reserved)                        */
  int32_t worker_num, num_workers;
  int32_t n_views, n_scalars, n_insns, n_regs, n_reds;
  /* axis reduction (ramba/ramba.py:8231-8244): the host orders the iteration dims
   * [reduced..., kept...]; the first n_axis_red_dims dims are walked sequentially per
   * output element (split into axis_nsplit slices for parallelism) and every RED slot
   * s leaves raw 64-bit partials of its accumulator class in
   *   red_scratch[(s*axis_nsplit + split)*kept_elems + kept_linear_index].
   * 0 = not an axis reduction (RED slots are global reductions written to reds[].out). */
  int32_t n_axis_red_dims;
  int32_t axis_nsplit;
  rb200_view views[RB200_MAX_VIEWS];
  uint64_t scalars[RB200_MAX_SCALARS]; /* raw bits: double / float(low 32) / int64 */
  rb200_insn insns[RB200_MAX_INSNS];
  rb200_red reds[RB200_MAX_REDS];
  /* scratch for cross-block reduction: >= rb200_red_scratch_bytes() bytes, zeroed
   * once at allocation (the kernel leaves its counters zero on exit)               */
  void* red_scratch;
} rb200_fused_op;

/* Launch one fused op on `stream` (a cudaStream_t, may be NULL = legacy default).   */
int rb200_run_deferred_ops(const rb200_fused_op* op, void* stream);

/* Bytes of zero-initialised device scratch a launch with reductions needs.          */
int64_t rb200_red_scratch_bytes(void);

/* out[j] = reduce_k partial[k*stride_k + j], j < n  (stage 2 of an axis reduction). */
int rb200_reduce_partials(void* out, const void* partials, int64_t n, int64_t k, int64_t stride_k,
                          int32_t dtype, int32_t redop, void* stream);

/* Inclusive cumulative scan (cumsum / scumulative with +, *, min, max) of one worker's block, replacing
 * RemoteState.scumulative_worker (ramba/ramba.py:3378-3437): the block is [n_outer][len][n_inner] elements in C order,
 * the scan runs along `len`.  dtype: RB200_F64 / F32 (float64 accumulation) or RB200_I64 / I32 (int64 accumulation).
 * carry_in (optional): one accumulator-class value per sequence (n_outer * n_inner), the total of the blocks that
 * precede this one along the axis; totals_out (optional): each sequence's inclusive total.  `scratch` must hold
 * rb200_cumulative_scratch_bytes() bytes.  src == dst is allowed.  One read and one write of every element.         */
int64_t rb200_cumulative_scratch_bytes(int64_t n_outer, int64_t len, int64_t n_inner);
int rb200_cumulative(const void* src, void* dst, int32_t dtype, int64_t n_outer, int64_t len, int64_t n_inner,
                     int32_t redop, const void* carry_in, void* totals_out, void* scratch, void* stream);

/* Which kernel rb200_run_deferred_ops would run `op` on and how (staged views, halos, TMA or cp.async loader, ring depth,
 * lean instructions, CTAs), as one text line in out[0..cap).  Needs no device and touches no pointer: the counterpart
 * of RAMBA_SHOW_CODE printing the generated kernel (ramba/ramba.py:8266-8284).                                       */
int rb200_describe_plan(const rb200_fused_op* op, char* out, int64_t cap);

/* Thread-local description of the last error returned on this thread.               */
const char* rb200_last_error(void);

/* RB200_ABI_VERSION the library was built with.                                     */
int rb200_abi_version(void);

/* Number of kernels this library has launched since load / last reset (bench.py's
 * gpu_launches claim).                                                              */
int64_t rb200_launch_count(void);
void rb200_reset_launch_count(void);

/* Device properties used for grid sizing: returns SM count of the current device,
 * or -1 when no CUDA device is usable.                                              */
int rb200_device_sm_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RAMBA_B200_H */
