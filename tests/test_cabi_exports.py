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
