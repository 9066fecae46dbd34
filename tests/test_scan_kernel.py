"""cumsum / scumulative on the single-pass scan kernel (ramba_b200/csrc/rb200_scan.cu): the reference's own shapes
(ramba/tests/test_distributed_array.py:1368-1386) plus sizes that need many tiles of the decoupled look-back (the chain
across tiles, across sequences, ragged last tiles), the column form (scan axis not the fastest) and the user-function
front end.  Integer and exactly representable data: bit-exact against NumPy."""
import numpy as onp
import pytest


def _h(x):
    return x.asarray() if hasattr(x, "asarray") else onp.asarray(x)


def scans(np, big):
    n1 = 5_000_003 if big else 10_007
    rng = onp.random.RandomState(3)
    out = []
    ah = rng.randint(-5, 6, size=n1)
    a = ah if np is onp else np.fromarray(ah)
    out.append(_h(np.cumsum(a)))
    fh = rng.randint(0, 4, size=n1 // 3).astype(onp.float32)  # partial sums stay < 2^24: exact in float32
    f = fh if np is onp else np.fromarray(fh)
    out.append(_h(np.cumsum(f)))
    dh = rng.randint(-8, 9, size=n1 // 2) * 0.25
    d = dh if np is onp else np.fromarray(dh)
    out.append(_h(np.cumsum(d)))
    shapes = [(37, 4099), (4099, 37), (13, 700, 9)] if big else [(5, 300), (300, 5), (3, 70, 4)]
    for shp in shapes:
        mh = rng.randint(-3, 4, size=shp)
        m = mh if np is onp else np.fromarray(mh)
        for axis in range(len(shp)):
            out.append(_h(np.cumsum(m, axis=axis)))
    out.append(_h(np.cumsum(a[100:-50:3])))  # a strided view is materialised first
    if np is onp:
        out += [onp.maximum.accumulate(ah), onp.minimum.accumulate(dh), onp.cumprod(onp.where(ah[:40] == 0, 1, ah[:40]))]
    else:
        out.append(_h(np.scumulative(lambda p, q: np.maximum(p, q) if hasattr(p, "shape") else max(p, q), None, a, axis=0)))
        out.append(_h(np.scumulative(lambda p, q: np.minimum(p, q) if hasattr(p, "shape") else min(p, q), None, d, axis=0)))
        nz = np.where(a[:40] == 0, 1, a[:40])
        out.append(_h(np.scumulative(lambda p, q: p * q, None, nz, axis=0)))
    return out


def _check(got, exp):
    assert len(got) == len(exp)
    for i, (g, e) in enumerate(zip(got, exp)):
        g, e = onp.asarray(g), onp.asarray(e)
        assert g.shape == e.shape and g.dtype == e.dtype, (i, g.shape, g.dtype, e.shape, e.dtype)
        assert onp.array_equal(g, e), "result %d: %d elements differ" % (i, int((g != e).sum()))


def test_scan_oracle(oracle_engine):
    import ramba_b200 as rb

    _check(scans(rb, False), scans(onp, False))


@pytest.mark.gpu
@pytest.mark.parametrize("big", [False, True], ids=["small", "many_tiles"])
def test_scan_cuda(gpu_engine, big):
    import ramba_b200 as rb
    from ramba_b200 import _cabi
    from ramba_b200.runtime import RT

    before = _cabi.launch_count()
    got = scans(rb, big)
    assert RT.is_cuda and _cabi.launch_count() > before
    _check(got, scans(onp, big))
