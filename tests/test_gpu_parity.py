"""-m gpu: the CUDA path (through the C-ABI) against the oracle on the same seeded programs."""
import numpy as onp
import pytest

import _programs

pytestmark = pytest.mark.gpu


def _run_oracle(prog):
    import _oracle_backend
    import ramba_b200 as rb
    from ramba_b200 import ramba
    from ramba_b200.runtime import RT

    ramba.deferred_op.ramba_deferred_ops = None
    RT.reset()
    _oracle_backend.install()
    try:
        return prog(rb)
    finally:
        ramba.deferred_op.ramba_deferred_ops = None
        RT.reset()


@pytest.mark.parametrize("prog", _programs.ALL, ids=lambda p: p.__name__)
def test_cuda_matches_oracle(gpu_engine, prog):
    import ramba_b200 as rb
    from ramba_b200 import _cabi
    from ramba_b200.runtime import RT

    before = _cabi.launch_count()
    got = prog(rb)
    assert RT.is_cuda and _cabi.launch_count() > before, "the CUDA library did not run"
    exp = _run_oracle(prog)
    for i, (g, e) in enumerate(zip(got, exp)):
        g, e = onp.asarray(g), onp.asarray(e)
        assert g.shape == e.shape and g.dtype == e.dtype
        if e.dtype.kind == "f" and prog.__name__ in ("chain",):
            # fp64 transcendentals: CUDA libdevice vs glibc, stated tolerance
            assert onp.allclose(g, e, rtol=1e-13, atol=1e-15), "%s[%d]" % (prog.__name__, i)
        elif e.dtype.kind == "f" and prog.__name__ == "arith_float":
            assert onp.allclose(g, e, rtol=1e-15, atol=0), "%s[%d]" % (prog.__name__, i)
        else:
            assert onp.array_equal(g, e), "%s[%d]" % (prog.__name__, i)


def test_chain_large(gpu_engine):
    """config-2 program at 2^26 elements: size-independent properties."""
    import ramba_b200 as rb

    N = 1 << 26
    A = rb.arange(N) / 1000.0
    B = rb.sin(A)
    C = rb.cos(A)
    D = B * B + C ** 2
    a, d = A.asarray(), D.asarray()
    assert onp.array_equal(a, onp.arange(N) * 0.001)
    assert onp.max(onp.abs(d - 1.0)) <= 4 * onp.finfo(onp.float64).eps
    b = B.asarray()
    idx = onp.arange(0, N, 4099)
    assert onp.allclose(b[idx], onp.sin(a[idx]), rtol=1e-13, atol=1e-15)


def _stencil3d(np, n=6, m=5, l=900):
    # config-4 shape of op (float32 field, float64 scalar), small outer dims and a long innermost dim
    u = np.fromfunction(lambda i, j, k: (i + 2 * j + 3 * k) % 64, (n, m, l), dtype=onp.float32)
    v = np.zeros((n, m, l), dtype=onp.float32)
    v[1:-1, 1:-1, 1:-1] = (u[:-2, 1:-1, 1:-1] + u[2:, 1:-1, 1:-1] + u[1:-1, :-2, 1:-1] + u[1:-1, 2:, 1:-1]
                           + u[1:-1, 1:-1, :-2] + u[1:-1, 1:-1, 2:] - 6.0 * u[1:-1, 1:-1, 1:-1])
    return [onp.asarray(v.asarray()), onp.asarray(v.sum())]


# N-d ops whose innermost dim fills whole tiles take the kernels' row mode (outer indices decoded per tile,
# not per element); the default sizes of tests/_programs.py all take the flat mode
_ROW_MODE_CASES = [
    ("stencil2d_1026", _programs.stencil2d, dict(n=12, m=1026)),
    ("stencil2d_1000", _programs.stencil2d, dict(n=9, m=1000)),
    ("stencil2d_2040", _programs.stencil2d, dict(n=5, m=2040)),
    ("reductions_2000", _programs.reductions, dict(n=7, m=2000)),
    ("minmax_900", _programs.reductions_minmax, dict(n=5, m=900)),
    ("broadcast_1000", _programs.broadcast_axis_sum, dict(n=6, m=1000)),
    ("transpose_900", _programs.transpose, dict(n=40, m=900)),
    ("stencil3d_900", _stencil3d, dict()),
    ("stencil3d_1024", _stencil3d, dict(n=4, m=4, l=1026)),
]


@pytest.mark.parametrize("case", _ROW_MODE_CASES, ids=lambda c: c[0])
def test_row_mode_matches_oracle(gpu_engine, case):
    import functools

    import ramba_b200 as rb

    _, fn, kw = case
    prog = functools.partial(fn, **kw)
    got = prog(rb)
    exp = _run_oracle(prog)
    for i, (g, e) in enumerate(zip(got, exp)):
        g, e = onp.asarray(g), onp.asarray(e)
        assert g.shape == e.shape and g.dtype == e.dtype
        assert onp.array_equal(g, e), "%s[%d]" % (case[0], i)
