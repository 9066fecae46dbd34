"""Host logic (API -> fuser -> lowering -> range binding) checked on CPU: op lists run through the
NumPy oracle instead of the CUDA kernels; results must equal plain NumPy running the same program
(the reference's run_both pattern, ramba/tests/test_distributed_array.py:255-259)."""
import numpy as onp
import pytest

import _programs


def _check(got, exp, name):
    assert len(got) == len(exp)
    for i, (g, e) in enumerate(zip(got, exp)):
        g, e = onp.asarray(g), onp.asarray(e)
        assert g.shape == e.shape, "%s[%d] shape %s vs %s" % (name, i, g.shape, e.shape)
        if e.dtype.kind == "f":
            assert onp.allclose(g, e, rtol=1e-13, atol=1e-15), "%s[%d]" % (name, i)
        else:
            assert onp.array_equal(g, e), "%s[%d]" % (name, i)


@pytest.mark.parametrize("prog", _programs.ALL, ids=lambda p: p.__name__)
def test_program_matches_numpy(oracle_engine, prog):
    import ramba_b200 as rb

    got = prog(rb)
    exp = prog(onp)
    _check(got, exp, prog.__name__)
