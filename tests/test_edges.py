"""Edge cases the domain has (ragged and empty inputs, sizes around the tile, every storage dtype,
> 2^31 elements).  Oracle-executor versions run on CPU; the same cases run through CUDA under -m gpu."""
import numpy as onp
import pytest

SIZES = [1, 2, 31, 99, 100, 101, 255, 256, 257, 2047, 2048, 2049, 4097, 10007]


def _h(x):
    return x.asarray() if hasattr(x, "asarray") else onp.asarray(x)


def _ragged(np):
    out = []
    for n in SIZES:
        a = np.arange(n) * 0.5 + 1.0
        b = np.sqrt(a) + a * a
        out.append(_h(b))
        out.append(onp.asarray((np.arange(n) % 3).sum()))
        s = np.arange(n)[n // 3:]
        out.append(_h((s * 2 + 1)))
    return out


def _empty(np):
    a = np.arange(0)
    z = np.zeros((0, 5))
    return [_h(a), _h(a + 1), _h(z), _h(np.arange(10)[5:5])]


def _dtypes(np):
    out = []
    base = np.arange(300)
    for dt in (onp.int32, onp.int16, onp.int8, onp.uint8, onp.uint16, onp.uint32, onp.float32, onp.float64, onp.int64):
        x = (base % 100).astype(dt)
        out.append(_h(x))
        out.append(_h((x + x)))
        out.append(_h((x * 2)))
        out.append(_h((x.astype(onp.float64) * 0.5)))
    m = (base % 2 == 0)
    out.append(_h(m))
    out.append(_h(np.logical_and(m, base > 100)))
    out.append(_h((~m)))
    out.append(onp.asarray(m.astype(onp.int64).sum()))  # (bool.sum() keeps dtype bool in the reference, array_unaryop ramba.py:5893-5894)
    return out


def _compare(got, exp):
    assert len(got) == len(exp)
    for i, (g, e) in enumerate(zip(got, exp)):
        g, e = onp.asarray(g), onp.asarray(e)
        assert g.shape == e.shape, (i, g.shape, e.shape)
        assert g.dtype == e.dtype or (g.dtype.kind == e.dtype.kind and g.dtype.itemsize == e.dtype.itemsize), (i, g.dtype, e.dtype)
        if e.dtype.kind == "f":
            assert onp.allclose(g, e, rtol=1e-15 if e.dtype == onp.float64 else 1e-6, atol=0), i
        else:
            assert onp.array_equal(g, e), (i, g.reshape(-1)[:6], e.reshape(-1)[:6])


@pytest.mark.parametrize("f", [_ragged, _empty, _dtypes], ids=lambda f: f.__name__)
def test_edges_oracle(oracle_engine, f):
    import ramba_b200 as rb

    _compare(f(rb), f(onp))


@pytest.mark.gpu
@pytest.mark.parametrize("f", [_ragged, _empty, _dtypes], ids=lambda f: f.__name__)
def test_edges_cuda(gpu_engine, f):
    import ramba_b200 as rb

    _compare(f(rb), f(onp))


@pytest.mark.gpu
def test_more_than_2_31_elements(gpu_engine):
    """64-bit indexing: 2.2e9-element int64 arange (17.6 GB) — values near the end and the sum."""
    import ramba_b200 as rb

    N = 2_200_000_000
    a = rb.arange(N)
    tail = a[N - 5:].asarray()
    assert onp.array_equal(tail, onp.arange(N - 5, N))
    mid = a[(1 << 31) - 2:(1 << 31) + 3].asarray()
    assert onp.array_equal(mid, onp.arange((1 << 31) - 2, (1 << 31) + 3))
    assert int(a.sum()) == N * (N - 1) // 2
    b = (a % 7).astype(onp.uint8)
    del a
    total = int(sum((N - r + 6) // 7 * r for r in range(7)))
    # like the reference, a reduction's result keeps the array dtype (array_unaryop, ramba/ramba.py:5893-5894):
    # the exact int64 accumulator is cast to uint8 at the end
    assert int(b.sum()) == total % 256


@pytest.mark.gpu
def test_unaligned_and_strided_views_full_speed_path(gpu_engine):
    """Views whose base is not 16-byte aligned (odd slice starts) and strided views take the cp.async
    staging path instead of bulk copies: results must be identical."""
    import ramba_b200 as rb

    N = 1 << 20
    a = rb.arange(N) * 1.0
    ref = onp.arange(N) * 1.0
    assert onp.array_equal((a[1:] + a[:-1]).asarray(), ref[1:] + ref[:-1])
    assert onp.array_equal((a[3::2] * 2.0).asarray(), ref[3::2] * 2.0)
    assert onp.array_equal((a[::-1] - 1.0).asarray(), ref[::-1] - 1.0)
    f = a.astype(onp.float32)
    assert onp.array_equal((f[5:] + 1.0).asarray(), (ref.astype(onp.float32)[5:] + 1.0).astype(onp.float32))
