"""Seeded random programs (tests/_random_programs.py) on the oracle executor and through the CUDA library."""
import numpy as onp
import pytest

import _random_programs


def _check(got, exp, name):
    assert len(got) == len(exp), name
    for i, (g, e) in enumerate(zip(got, exp)):
        g, e = onp.asarray(g), onp.asarray(e)
        assert g.shape == e.shape, "%s[%d]: shape %s vs %s" % (name, i, g.shape, e.shape)
        assert onp.allclose(g, e, rtol=1e-13, atol=1e-12), "%s[%d]" % (name, i)


@pytest.mark.parametrize("f", _random_programs.CASES, ids=lambda f: f.__name__)
def test_random_program_oracle(oracle_engine, f):
    import ramba_b200 as rb

    _check(f(rb), f(onp), f.__name__)


@pytest.mark.gpu
@pytest.mark.parametrize("f", _random_programs.CASES, ids=lambda f: f.__name__)
def test_random_program_cuda(gpu_engine, f):
    import ramba_b200 as rb
    from ramba_b200 import _cabi
    from ramba_b200.runtime import RT

    before, before_rt = _cabi.launch_count(), RT.launches
    got = f(rb)
    assert RT.is_cuda
    # a seed may draw only empty slices: then there is nothing to launch (RT.launches counts the op lists the
    # engine handed to its executor); whenever the engine did launch, the CUDA library must have run them
    if RT.launches > before_rt:
        assert _cabi.launch_count() > before, "the CUDA library did not run"
    _check(got, f(onp), f.__name__)
