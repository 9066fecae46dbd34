"""The C-ABI library loads (no GPU needed) and exports every symbol include/ramba_b200.h declares;
without a CUDA device launches fail loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def test_library_exports_header_symbols():
    from ramba_b200 import _cabi

    hdr = open(os.path.join(ROOT, "include", "ramba_b200.h")).read()
    declared = set(re.findall(r"\b(rb200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_cabi.EXPORTS), (declared, set(_cabi.EXPORTS))
    lib = ctypes.CDLL(_cabi.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert _cabi.load().rb200_abi_version() == _cabi.ABI_VERSION


def test_struct_layout_matches_header():
    from ramba_b200 import _cabi

    assert ctypes.sizeof(_cabi.Insn) == 16
    assert ctypes.sizeof(_cabi.View) == 8 + 8 * _cabi.MAX_DIMS + 8
    assert _cabi.FusedOp.views.offset % 8 == 0 and _cabi.FusedOp.insns.offset % 8 == 0


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from ramba_b200 import _cabi

    fop = _cabi.FusedOp()
    fop.abi_version = _cabi.ABI_VERSION
    fop.ndim = 1
    fop.itershape[0] = 16
    fop.n_insns = 1
    with pytest.raises(_cabi.CabiError):
        _cabi.run_deferred_ops(fop)
    from ramba_b200.runtime import Runtime

    with pytest.raises(RuntimeError):
        Runtime().device


def _one_insn_op():
    """A minimal well-formed fused op: views[0][0:10] = scalars[0] (float64)."""
    from ramba_b200 import _cabi

    f = _cabi.FusedOp()
    f.abi_version = _cabi.ABI_VERSION
    f.ndim, f.n_views, f.n_insns, f.n_scalars, f.num_workers = 1, 1, 1, 1, 1
    f.itershape[0] = 10
    f.views[0].base = 0x1000  # never dereferenced: every case below is rejected (or empty) before a launch
    f.views[0].stride[0] = 1
    f.views[0].dtype = _cabi.F64
    i = f.insns[0]
    i.op, i.ctype, i.a_kind, i.a_idx = _cabi.OP["MOV"], _cabi.T_F64, _cabi.K_SCAL, 0
    i.st_reg = i.st2 = i.mask_reg = _cabi.NOSTORE
    i.st_view = 0
    return f


def test_malformed_op_lists_are_rejected_with_a_reason():
    """Error convention of the boundary (include/ramba_b200.h): nonzero status + thread-local message; the op
    list is validated before any device work, so this needs no GPU."""
    import ctypes as C

    from ramba_b200 import _cabi

    lib = _cabi.load()

    def run(mutate):
        f = _one_insn_op()
        mutate(f)
        rc = lib.rb200_run_deferred_ops(C.byref(f), None)
        return rc, lib.rb200_last_error().decode()

    cases = [
        (lambda f: setattr(f, "abi_version", 99), "ABI version"),
        (lambda f: setattr(f, "ndim", 9), "ndim"),
        (lambda f: setattr(f, "n_insns", 500), "too many instructions"),
        (lambda f: f.itershape.__setitem__(0, -1), "negative itershape"),
        (lambda f: setattr(f.insns[0], "op", 200), "bad opcode"),
        (lambda f: setattr(f.insns[0], "ctype", 7), "bad compute class"),
        (lambda f: setattr(f.insns[0], "a_kind", 9), "bad operand kind"),
        (lambda f: setattr(f.insns[0], "a_idx", 5), "scalar index out of range"),
        (lambda f: setattr(f.insns[0], "st_view", 3), "st_view out of range"),
        (lambda f: setattr(f.insns[0], "st_reg", 2), "st_reg out of range"),
        (lambda f: setattr(f.insns[0], "mask_reg", 1), "mask_reg out of range"),
        (lambda f: setattr(f.views[0], "dtype", 55), "bad view dtype"),
        (lambda f: setattr(f.views[0], "base", 0), "null view base pointer"),
    ]
    for mutate, reason in cases:
        rc, msg = run(mutate)
        assert rc != 0 and reason in msg, (reason, rc, msg)
    # an empty iteration space is a successful no-op
    rc, _ = run(lambda f: f.itershape.__setitem__(0, 0))
    assert rc == 0
    with pytest.raises(_cabi.CabiError):
        bad = _one_insn_op()
        bad.insns[0].op = 200
        _cabi.run_deferred_ops(bad)
