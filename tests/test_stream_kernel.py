"""The streaming kernel of the lean machine (ramba_b200/csrc/rb200_stream.cu) against the oracle: float-arithmetic op
lists over 1-D spaces - full and ragged tiles, 16-byte aligned and unaligned bases (staged vs direct operands),
strided views, float32/float64 mixes, spill registers, global reductions, and the column form of axis reductions
(rows x multiples of 2048 columns, row-broadcast operands).  Elementwise results are compared BIT-exactly; reductions
use exactly representable data, so every summation order gives the same bits."""
import numpy as onp
import pytest


def _h(x):
    return x.asarray() if hasattr(x, "asarray") else onp.asarray(x)


def _data(n, dtype, seed):
    rng = onp.random.RandomState(seed)
    return (rng.rand(n) * 8 - 4).astype(dtype)


def elementwise(np, n, dtype, seed=0):
    ah, bh = _data(n, dtype, seed), _data(n, dtype, seed + 1)
    a, b = (ah, bh) if np is onp else (np.fromarray(ah), np.fromarray(bh))
    c = a * 2.0 + b                      # python float: computed in float64 for float32 arrays, rounded on store
    d = (a - b) * (a + b) - 0.5 * c
    e = abs(d) + np.minimum(a, b) * np.maximum(a, 0.25)
    f = a[3:-5] * b[8:] + c[3:-5]        # odd element offsets: unaligned bases -> direct operands next to staged ones
    g = a[::3] - b[::3] * 0.125          # strided
    a2 = a * 1.0
    a2 += b
    a2 *= 0.5
    return [_h(c), _h(d), _h(e), _h(f), _h(g), _h(a2), _h(b ** 2 - a * a)]


def many_temporaries(np, n, dtype, seed=2):
    ah = _data(n, dtype, seed)
    a = ah if np is onp else np.fromarray(ah)
    ts = [a * float(i + 2) for i in range(5)]
    acc = ts[0]
    for t in ts[1:]:
        acc = acc * 0.5 + t
    for t in ts:
        acc = acc - t * 0.25
    del ts, t
    return [_h(acc)]


def global_sums(np, n, dtype, seed=3):
    rng = onp.random.RandomState(seed)
    xh = rng.randint(0, 4, size=n).astype(dtype)
    x = xh if np is onp else np.fromarray(xh)
    out = [onp.asarray((x * 2.0 + 1.0).sum()), onp.asarray(x.sum()), onp.asarray((x * x).sum()), onp.asarray((x - 1.0).min()),
           onp.asarray((x * 0.5).max())]
    y = x * 3.0
    out += [onp.asarray(y.sum()), _h(y)]
    return out


def column_sums(np, r, c, dtype, seed=4):
    rng = onp.random.RandomState(seed)
    mh = rng.randint(0, 8, size=(r, c)).astype(dtype)
    vh = rng.randint(0, 8, size=c).astype(dtype)
    m, v = (mh, vh) if np is onp else (np.fromarray(mh), np.fromarray(vh))
    return [_h((m + v).sum(axis=0)), _h(m.sum(axis=0)), _h((m * 2.0 + v).sum(axis=0)), _h((m * v).max(axis=0))]


CASES = [
    ("elementwise_f32_full", lambda np: elementwise(np, 8 * 2048, onp.float32)),
    ("elementwise_f32_ragged", lambda np: elementwise(np, 5 * 2048 + 777, onp.float32)),
    ("elementwise_f64_ragged", lambda np: elementwise(np, 3 * 2048 + 5, onp.float64)),
    ("elementwise_f32_small", lambda np: elementwise(np, 300, onp.float32)),
    ("elementwise_f64_big", lambda np: elementwise(np, 700_001, onp.float64)),
    ("many_temporaries_f32", lambda np: many_temporaries(np, 40_000, onp.float32)),
    ("many_temporaries_f64", lambda np: many_temporaries(np, 9_999, onp.float64)),
    ("global_sums_f32", lambda np: global_sums(np, 1_000_003, onp.float32)),
    ("global_sums_f64", lambda np: global_sums(np, 300_000, onp.float64)),
    ("global_sums_f32_tiny", lambda np: global_sums(np, 150, onp.float32)),
    ("column_sums_f32", lambda np: column_sums(np, 300, 4096, onp.float32)),
    ("column_sums_f64", lambda np: column_sums(np, 77, 2048, onp.float64)),
    ("column_sums_f32_wide", lambda np: column_sums(np, 40, 8 * 2048, onp.float32)),
]


def _check(got, exp, name, exact):
    assert len(got) == len(exp)
    for i, (g, e) in enumerate(zip(got, exp)):
        g, e = onp.asarray(g), onp.asarray(e)
        assert g.shape == e.shape and g.dtype == e.dtype, "%s[%d]: %s %s vs %s %s" % (name, i, g.shape, g.dtype, e.shape, e.dtype)
        if exact:
            assert onp.array_equal(g, e), "%s[%d]: %d elements differ" % (name, i, int((g != e).sum()))
        else:
            assert onp.allclose(g, e, rtol=1e-5 if e.dtype == onp.float32 else 1e-12, atol=1e-5 if e.dtype == onp.float32 else 1e-12), "%s[%d]" % (name, i)


@pytest.mark.parametrize("name,prog", CASES, ids=[c[0] for c in CASES])
def test_stream_oracle_vs_numpy(oracle_engine, name, prog):
    import ramba_b200 as rb

    # float32 arrays with Python-float scalars are computed in float64 by the op list (Numba's typing): tolerance
    # against NumPy here, bit-exact against the oracle on the GPU leg
    _check(prog(rb), prog(onp), name, exact=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name,prog", CASES, ids=[c[0] for c in CASES])
def test_stream_cuda_matches_oracle(gpu_engine, name, prog):
    import _oracle_backend
    import ramba_b200 as rb
    from ramba_b200 import _cabi, ramba
    from ramba_b200.runtime import RT

    before = _cabi.launch_count()
    got = prog(rb)
    assert RT.is_cuda and _cabi.launch_count() > before, "the CUDA library did not run"
    ramba.deferred_op.ramba_deferred_ops = None
    RT.reset()
    _oracle_backend.install()
    try:
        exp = prog(rb)
    finally:
        ramba.deferred_op.ramba_deferred_ops = None
        RT.reset()
    _check(got, exp, name, exact=True)
