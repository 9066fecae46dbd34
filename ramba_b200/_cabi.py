"""ctypes binding of libramba_b200.so (include/ramba_b200.h).

This is the thin C-ABI the deferred-op fuser dispatches through.  The reference ships Python
source to its workers and lets Numba compile it (ramba/ramba.py:3526-3545, 249-438); here the
worker side is a prebuilt sm_100a library and the "kernel source" is an op list.

There is no fallback: if the library is missing or no CUDA device is usable, loading or launching
raises.
"""
import ctypes as C
import os

ABI_VERSION = 4
MAX_DIMS = 5
MAX_VIEWS = 16
MAX_SCALARS = 32
MAX_INSNS = 96
MAX_REGS = 12
MAX_REDS = 4
NOSTORE = 0xFF

# storage dtypes
F64, F32, I64, I32, BOOL, U8, I8, I16, U16, U32 = range(10)
# compute classes
T_F64, T_F32, T_I64 = range(3)
# operand kinds
K_NONE, K_ACC, K_REG, K_VIEW, K_SCAL, K_IOTA = range(6)

OPS = [
    "MOV", "ADD", "SUB", "MUL", "DIV", "FLOORDIV", "MOD", "POW", "POWI", "MIN", "MAX",
    "GT", "LT", "GE", "LE", "EQ", "NE", "LAND", "LOR", "LXOR", "BAND", "BOR", "BXOR", "SHL", "SHR",
    "ABS", "SQUARE", "SQRT", "SIN", "COS", "TAN", "SINH", "COSH", "TANH", "ASIN", "ACOS", "ATAN",
    "NEG", "EXP", "LOG", "ISFINITE", "ISINF", "ISNAN", "ISNEGINF", "ISPOSINF", "LNOT", "INVERT",
    "WHERE", "CVT", "SINCOS", "RED", "CBRT", "MULADD", "MULSUB", "MULRSUB",
]
OP = {name: i for i, name in enumerate(OPS)}
RED_ADD, RED_MUL, RED_MIN, RED_MAX = range(4)


class Insn(C.Structure):
    _fields_ = [
        ("op", C.c_uint8), ("ctype", C.c_uint8),
        ("a_kind", C.c_uint8), ("a_idx", C.c_uint8),
        ("b_kind", C.c_uint8), ("b_idx", C.c_uint8),
        ("c_kind", C.c_uint8), ("c_idx", C.c_uint8),
        ("st_reg", C.c_uint8), ("st_view", C.c_uint8),
        ("st2", C.c_uint8), ("mask_reg", C.c_uint8),
        ("imm", C.c_uint32),
    ]


class View(C.Structure):
    _fields_ = [
        ("base", C.c_void_p),
        ("stride", C.c_int64 * MAX_DIMS),
        ("dtype", C.c_int32),
        ("flags", C.c_int32),
        ("alloc_lo", C.c_void_p),
        ("alloc_hi", C.c_void_p),
    ]


class Red(C.Structure):
    _fields_ = [
        ("op", C.c_int32), ("ctype", C.c_int32),
        ("out", C.c_void_p),
        ("out_dtype", C.c_int32), ("pad", C.c_int32),
    ]


class FusedOp(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("ndim", C.c_int32),
        ("itershape", C.c_int64 * MAX_DIMS),
        ("global_start", C.c_int64 * MAX_DIMS),
        ("iota_dim", C.c_int32 * MAX_DIMS),
        ("worker_num", C.c_int32), ("num_workers", C.c_int32),
        ("n_views", C.c_int32), ("n_scalars", C.c_int32), ("n_insns", C.c_int32),
        ("n_regs", C.c_int32), ("n_reds", C.c_int32),
        ("n_axis_red_dims", C.c_int32), ("axis_nsplit", C.c_int32),
        ("views", View * MAX_VIEWS),
        ("scalars", C.c_uint64 * MAX_SCALARS),
        ("insns", Insn * MAX_INSNS),
        ("reds", Red * MAX_REDS),
        ("red_scratch", C.c_void_p),
    ]


assert C.sizeof(Insn) == 16

# every symbol include/ramba_b200.h declares
EXPORTS = [
    "rb200_run_deferred_ops",
    "rb200_red_scratch_bytes",
    "rb200_reduce_partials",
    "rb200_cumulative",
    "rb200_cumulative_scratch_bytes",
    "rb200_describe_plan",
    "rb200_last_error",
    "rb200_abi_version",
    "rb200_launch_count",
    "rb200_reset_launch_count",
    "rb200_device_sm_count",
]

_LIB = None


def lib_path():
    # RAMBA_B200_LIB: A/B-test another build of the same ABI (development aid)
    return os.environ.get("RAMBA_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libramba_b200.so")


class CabiError(RuntimeError):
    pass


def load():
    """Load libramba_b200.so; raises (loudly) if it was not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise CabiError(
            "libramba_b200.so not found at %s — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % path
        )
    lib = C.CDLL(path)
    lib.rb200_run_deferred_ops.argtypes = [C.POINTER(FusedOp), C.c_void_p]
    lib.rb200_run_deferred_ops.restype = C.c_int
    lib.rb200_red_scratch_bytes.argtypes = []
    lib.rb200_red_scratch_bytes.restype = C.c_int64
    lib.rb200_reduce_partials.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.rb200_reduce_partials.restype = C.c_int
    lib.rb200_cumulative.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    lib.rb200_cumulative.restype = C.c_int
    lib.rb200_cumulative_scratch_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int64]
    lib.rb200_cumulative_scratch_bytes.restype = C.c_int64
    lib.rb200_describe_plan.argtypes = [C.POINTER(FusedOp), C.c_char_p, C.c_int64]
    lib.rb200_describe_plan.restype = C.c_int
    lib.rb200_last_error.argtypes = []
    lib.rb200_last_error.restype = C.c_char_p
    lib.rb200_abi_version.argtypes = []
    lib.rb200_abi_version.restype = C.c_int
    lib.rb200_launch_count.argtypes = []
    lib.rb200_launch_count.restype = C.c_int64
    lib.rb200_reset_launch_count.argtypes = []
    lib.rb200_reset_launch_count.restype = None
    lib.rb200_device_sm_count.argtypes = []
    lib.rb200_device_sm_count.restype = C.c_int
    if lib.rb200_abi_version() != ABI_VERSION:
        raise CabiError("libramba_b200.so ABI %d != binding ABI %d: rebuild" % (lib.rb200_abi_version(), ABI_VERSION))
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        raise CabiError("libramba_b200: " + load().rb200_last_error().decode("utf-8", "replace"))


def run_deferred_ops(fop, stream=None):
    """Launch one fused op over one range (FusedOp struct) on a cudaStream_t handle (int or None)."""
    lib = load()
    check(lib.rb200_run_deferred_ops(C.byref(fop), C.c_void_p(stream) if stream else None))


def reduce_partials(out_ptr, part_ptr, n, k, stride_k, dtype, redop, stream=None):
    lib = load()
    check(lib.rb200_reduce_partials(C.c_void_p(out_ptr), C.c_void_p(part_ptr), n, k, stride_k, dtype, redop,
                                    C.c_void_p(stream) if stream else None))


def cumulative(src, dst, dtype, n_outer, length, n_inner, redop, carry_in=None, totals_out=None, scratch=None, stream=None):
    """Inclusive scan of one block [n_outer][length][n_inner] along `length` (device pointers as ints)."""
    lib = load()
    check(lib.rb200_cumulative(C.c_void_p(src), C.c_void_p(dst), dtype, n_outer, length, n_inner, redop,
                               C.c_void_p(carry_in) if carry_in else None, C.c_void_p(totals_out) if totals_out else None,
                               C.c_void_p(scratch) if scratch else None, C.c_void_p(stream) if stream else None))


def cumulative_scratch_bytes(n_outer, length, n_inner):
    return int(load().rb200_cumulative_scratch_bytes(n_outer, length, n_inner))


def describe_plan(fop):
    """One text line: the kernel the library would run this fused op on, and how (no device needed)."""
    lib = load()
    buf = C.create_string_buffer(600)
    check(lib.rb200_describe_plan(C.byref(fop), buf, 600))
    return buf.value.decode()


def red_scratch_bytes():
    return int(load().rb200_red_scratch_bytes())


def launch_count():
    return int(load().rb200_launch_count())


def reset_launch_count():
    load().rb200_reset_launch_count()
