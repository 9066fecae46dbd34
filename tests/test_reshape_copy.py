"""reshape_copy (ramba/ramba.py:9241-9277): the C-order redistribution behind a general reshape, as linear-run pieces packed
per peer (ramba_b200/redistribute.py).  Checked against NumPy's reshape for merges, splits, unrelated factorizations,
views as sources, every dtype class, empty and 0-d arrays; the run algebra against brute force."""
import numpy as onp
import pytest

PAIRS = [((360,), (24, 15)), ((24, 15), (360,)), ((24, 15), (15, 24)), ((12, 10, 3), (36, 10)), ((36, 10), (4, 9, 10)),
         ((7, 11, 13), (1001,)), ((1001,), (7, 143)), ((4, 9), (6, 6)), ((2, 3, 4, 5), (5, 4, 3, 2)), ((128, 3), (3, 128)),
         ((500, 2), (-1,)), ((6, 70), (3, -1, 7)), ((1, 300), (300, 1)), ((17, 19), (19, 17))]


def reshape_programs(np):
    out = []
    h = (lambda x: x.asarray()) if np is not onp else (lambda x: onp.asarray(x))
    rc = (lambda a, s: a.reshape_copy(s)) if np is not onp else (lambda a, s: onp.reshape(a, s).copy())
    for i, (a, b) in enumerate(PAIRS):
        n = int(onp.prod(a))
        x = np.arange(n).reshape(a) if np is onp else np.fromarray(onp.arange(n).reshape(a))
        r = rc(x * 2 + 1, b)
        out.append(h(r))
        out.append(h(r + 1))                       # the result is an ordinary operand of the fused path
    # sources that are views: slices, steps, transposes
    base = onp.arange(40 * 30, dtype=onp.float32).reshape(40, 30) * 0.5
    X = base if np is onp else np.fromarray(base)
    out.append(h(rc(X[3:33, 5:25], (20, 30))))
    out.append(h(rc(X.T, (40, 30))))
    out.append(h(rc(X[::2, ::-1], (30, 20))))
    out.append(h(rc(X > 100.0, (1200,))))         # bool
    out.append(h(rc((X * 3).astype(onp.int32), (24, 50))))
    return out


def test_reshape_copy_matches_numpy(oracle_engine):
    import ramba_b200 as rb

    got, exp = reshape_programs(rb), reshape_programs(onp)
    assert len(got) == len(exp)
    for i, (g, e) in enumerate(zip(got, exp)):
        assert g.shape == e.shape and g.dtype == e.dtype and onp.array_equal(g, e), i


def test_edge_cases_and_the_reshape_switch(oracle_engine, monkeypatch):
    import ramba_b200 as rb
    from ramba_b200 import common, ramba

    x = rb.fromarray(onp.arange(120.0))
    with pytest.raises(ramba.ReshapeError):
        x.reshape(10, 12)                          # like the reference: an in-place reshape cannot be distributed
    monkeypatch.setattr(common, "reshape_forwarding", True)   # RAMBA_RESHAPE_COPY=1
    assert onp.array_equal(x.reshape(10, 12).asarray(), onp.arange(120.0).reshape(10, 12))
    assert onp.array_equal(rb.reshape(x, (2, -1)).asarray(), onp.arange(120.0).reshape(2, 60))
    with pytest.raises(ValueError):
        x.reshape_copy((7, 11))
    with pytest.raises(ValueError):
        x.reshape_copy((-1, -1))
    assert rb.zeros((0, 5)).reshape_copy((5, 0)).asarray().shape == (5, 0)
    assert onp.array_equal(rb.array(3.5).reshape_copy((1, 1)).asarray(), onp.full((1, 1), 3.5))
    # unit-dim reshapes stay views
    y = x.reshape(1, 120, 1)
    assert y.base is not None and onp.array_equal(y.asarray(), onp.arange(120.0).reshape(1, 120, 1))


def test_regular_reshapes_are_a_handful_of_strided_copies(oracle_engine):
    """A reshape between row-partitioned layouts is one 2-D strided copy per run family, not one launch per row."""
    import ramba_b200 as rb
    from ramba_b200.runtime import RT

    x = rb.fromarray(onp.arange(4096 * 8, dtype=onp.float32).reshape(4096, 8))
    rb.sync()
    l0 = RT.launches
    y = x.reshape_copy((8, 4096))
    assert RT.launches - l0 <= 2
    assert onp.array_equal(y.asarray(), onp.arange(4096 * 8, dtype=onp.float32).reshape(8, 4096))


def test_run_algebra_against_brute_force():
    from ramba_b200 import redistribute as R

    rng = onp.random.RandomState(0)
    for _ in range(300):
        k = int(rng.randint(1, 4))
        shape = [int(rng.randint(1, 7)) for _ in range(k)]
        start = [int(rng.randint(0, s)) for s in shape]
        size = [int(rng.randint(1, s - st + 1)) for s, st in zip(shape, start)]
        if rng.rand() < 0.4:
            for j in range(k - 1, 0, -1):
                start[j], size[j] = 0, shape[j]
                if rng.rand() < 0.5:
                    break
        lin = onp.arange(int(onp.prod(shape))).reshape(shape)
        blk = lin[tuple(slice(a, a + n) for a, n in zip(start, size))].reshape(-1)
        for tail in (True, False):
            st, ln, m = R.block_runs(shape, start, size, tail)
            assert onp.array_equal((st[:, None] + onp.arange(ln)[None, :]).reshape(-1), blk)
            lstr = [1] * k
            for j in range(k - 2, -1, -1):
                lstr[j] = lstr[j + 1] * size[j + 1]
            off = R.run_local_offsets(size, m, lstr, 0)
            assert onp.array_equal((off[:, None] + onp.arange(ln)[None, :]).reshape(-1), onp.arange(len(blk)))
    for _ in range(200):
        la, lb = int(rng.randint(1, 9)), int(rng.randint(1, 9))
        a = onp.cumsum(rng.randint(la, la + 6, size=int(rng.randint(0, 8)))).astype(onp.int64)
        b = onp.cumsum(rng.randint(lb, lb + 6, size=int(rng.randint(0, 8)))).astype(onp.int64)
        ia, ib, s, n = R.intersect_runs(a, la, b, lb)
        got = sorted((int(x), int(y)) for x, y in zip(s, n))
        exp = []
        for x in a:
            for y in b:
                lo, hi = max(int(x), int(y)), min(int(x) + la, int(y) + lb)
                if hi > lo:
                    exp.append((lo, hi - lo))
        assert got == sorted(exp)
