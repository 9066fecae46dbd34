"""-m gpu leg of the DAG fuzzer (tests/_dag_fuzz.py) through the CUDA library.  Collected last on purpose: it was added
after the round's GPU budget was spent, so its first run on a B200 is the driver's; under `pytest -x` it must not stand
in front of the tests that have been seen green.  Elementwise results are compared exactly; reductions may be
summed in another order on the GPU, so everything is compared with a tolerance far below what a mis-ordered or dropped
statement would change (values are small multiples of powers of two)."""
import numpy as onp
import pytest

pytestmark = pytest.mark.gpu


def _close(got, exp, name):
    assert len(got) == len(exp), name
    for i, (g, e) in enumerate(zip(got, exp)):
        g, e = onp.asarray(g), onp.asarray(e)
        assert g.shape == e.shape and onp.allclose(g, e, rtol=1e-12, atol=1e-9), "%s[%d]" % (name, i)


@pytest.mark.parametrize("chunk", range(4))
def test_fuzzed_programs_cuda(gpu_engine, chunk):
    import _dag_fuzz
    import ramba_b200 as rb
    from ramba_b200 import _cabi
    from ramba_b200.runtime import RT

    before = _cabi.launch_count()
    for f in _dag_fuzz.CASES[chunk * 15:(chunk + 1) * 15]:
        _close(f(rb), f(onp), f.__name__)
    assert RT.is_cuda and _cabi.launch_count() > before


@pytest.mark.parametrize("chunk", range(2))
def test_programs_at_the_fusers_table_sizes_cuda(gpu_engine, chunk):
    """tests/_limit_fuzz.py through the CUDA library: includes op lists that load a view into a register, store to the same
    view and use the OLD value afterwards (`t = a*2; a += 1; r = t - b` fused) - the pattern whose oracle evaluation was
    wrong until this round (oracle/vm.py handed out an alias of the host buffer on loads)."""
    import _limit_fuzz
    import ramba_b200 as rb
    from ramba_b200 import _cabi
    from ramba_b200.runtime import RT

    before = _cabi.launch_count()
    for f in _limit_fuzz.CASES[chunk * 12:(chunk + 1) * 12]:
        _close(f(rb), f(onp), f.__name__)
    assert RT.is_cuda and _cabi.launch_count() > before


def test_reshape_copy_cuda(gpu_engine):
    """tests/test_reshape_copy.py's programs through the CUDA library (strided copies on the op-list kernels)."""
    import ramba_b200 as rb
    import test_reshape_copy
    from ramba_b200 import _cabi
    from ramba_b200.runtime import RT

    before = _cabi.launch_count()
    got, exp = test_reshape_copy.reshape_programs(rb), test_reshape_copy.reshape_programs(onp)
    assert len(got) == len(exp)
    for i, (g, e) in enumerate(zip(got, exp)):
        assert g.shape == e.shape and g.dtype == e.dtype and onp.array_equal(g, e), i
    assert RT.is_cuda and _cabi.launch_count() > before


def test_late_api_cases_cuda(gpu_engine):
    """The API cases of tests/test_api_parity.py that were added after the GPU budget was spent."""
    import ramba_b200 as rb
    import test_api_parity
    from ramba_b200 import _cabi
    from ramba_b200.runtime import RT

    for f in test_api_parity.CASES:
        if f.__name__ in test_api_parity.FIRST_GPU_RUN_IS_THE_DRIVERS:
            before = _cabi.launch_count()
            got = f(rb)
            assert RT.is_cuda and _cabi.launch_count() > before
            test_api_parity._compare(got, f(onp), f.__name__)


@pytest.mark.parametrize("chunk", range(2))
def test_mixed_statement_forms_cuda(gpu_engine, chunk):
    """tests/_expr_fuzz.py through the CUDA library (integer-valued data: results are exact; compared with the file's
    tolerance all the same)."""
    import _expr_fuzz
    import ramba_b200 as rb
    from ramba_b200 import _cabi
    from ramba_b200.runtime import RT

    before = _cabi.launch_count()
    for f in _expr_fuzz.CASES[chunk * 15:(chunk + 1) * 15]:
        _close(f(rb), f(onp), f.__name__)
    assert RT.is_cuda and _cabi.launch_count() > before


@pytest.mark.parametrize("which", ["VIEW_CASES", "API_CASES", "TRIG_CASES", "SHAPE_CASES", "SKELETON_CASES", "REDUCTION_CASES", "MIXED_CASES"])
def test_more_fuzzed_families_cuda(gpu_engine, which):
    """The view / library-call / trig-and-mask / odd-shape / skeleton / reduction programs of tests/_expr_fuzz.py through the CUDA library (15 seeds each; the
    trig family with a tolerance for libdevice vs NumPy transcendentals and the summation order of their sums)."""
    import _expr_fuzz
    import ramba_b200 as rb
    from ramba_b200 import _cabi
    from ramba_b200.runtime import RT

    before = _cabi.launch_count()
    for f in getattr(_expr_fuzz, which)[:15]:
        got, exp = f(rb), f(onp)
        assert len(got) == len(exp), f.__name__
        for i, (g, e) in enumerate(zip(got, exp)):
            g, e = onp.asarray(g), onp.asarray(e)
            tol = 1e-10 if which in ("TRIG_CASES", "MIXED_CASES") else 1e-12
            assert g.shape == e.shape and g.dtype == e.dtype and onp.allclose(g, e, rtol=tol, atol=tol), "%s[%d]" % (f.__name__, i)
    assert RT.is_cuda and _cabi.launch_count() > before


def test_fuzz_programs_match_the_real_reference_cuda(gpu_engine):
    """The fuzz programs the real reference runs correctly (tests/golden/fuzz_golden.npz, 70 programs), through the CUDA library
    against the reference's own outputs."""
    import json
    import os

    import ramba_b200 as rb
    import test_golden

    z = onp.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuzz_golden.npz"))
    status = json.loads(str(z["__status__"]))
    n = 0
    for name, st in status.items():
        if st == "ok" and not name.startswith("typing"):
            fn, seed = test_golden._fuzz_fn(name)
            got = fn(rb, seed)
            keys = sorted((k for k in z.files if k.startswith(name + "__")), key=lambda k: int(k.rsplit("o", 1)[1]))
            assert len(keys) == len(got), name
            for i, k in enumerate(keys):
                g, e = onp.asarray(got[i]), z[k]
                assert g.shape == e.shape and onp.allclose(g, e, rtol=1e-12, atol=1e-9), "%s[%d]" % (name, i)
            n += 1
    assert n >= 60


def test_typing_and_rounding_follow_the_real_reference_cuda(gpu_engine):
    """The typing programs the real reference runs (float32 / float64 / int64 with scalars, inexact values), through the CUDA
    library against the reference's bits: float32 results exact, float64 results within one rounding of a product (the
    reference may fuse `x*y + z`; see tests/test_golden.py)."""
    import json
    import os

    import ramba_b200 as rb
    import test_golden

    z = onp.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuzz_golden.npz"))
    status = json.loads(str(z["__status__"]))
    n = 0
    for name, st in status.items():
        if st == "ok" and name.startswith("typing"):
            fn, seed = test_golden._fuzz_fn(name)
            test_golden._compare_fuzz(name, fn(rb, seed), z, one_ulp_f64=True, f32_ulps=1)
            n += 1
    assert n >= 30
