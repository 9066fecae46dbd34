"""The lazy DAG in front of the fuser (ramba_b200/ramba.py::DAG; reference: ramba/ramba.py:4387-5293).

What a user of the reference's default (DAG) mode can observe, checked on the oracle executor:
results nobody can observe are never computed, materialising one array runs only what it depends on, interleaved chains
over different shapes fuse per shape, and every reordering respects read-after-write / write-after-read /
write-after-write on the arrays (views included).  RAMBA_NO_DAG behaviour (statements go straight to the fuser) must give
the same values."""
import numpy as onp
import pytest

import _random_programs


@pytest.fixture
def eng(oracle_engine, monkeypatch):
    import ramba_b200 as rb
    from ramba_b200 import ramba
    from ramba_b200.runtime import RT

    monkeypatch.setattr(ramba, "NO_DAG", False)  # (whatever RAMBA_NO_DAG says: these tests are about the DAG)
    return rb, ramba, RT


def test_statements_wait_in_the_dag_until_something_is_read(eng):
    rb, ramba, RT = eng
    l0 = RT.launches
    a = rb.arange(1000) * 2.0
    b = a + 1.0
    assert len(ramba.DAG.pending) == 3 and ramba.deferred_op.ramba_deferred_ops is None and RT.launches == l0
    assert onp.array_equal(b.asarray(), onp.arange(1000) * 2.0 + 1.0)
    assert not ramba.DAG.pending and RT.launches == l0 + 1


def test_unobservable_results_are_never_computed(eng):
    rb, ramba, RT = eng
    x = rb.fromarray(onp.arange(5000, dtype=onp.float64))
    y = rb.fromarray(onp.arange(300, dtype=onp.float64))
    rb.sync()
    l0, p0 = RT.launches, ramba.DAG.pruned_count
    rb.sin(y)                 # result dropped at once
    t = y * 3.0
    u = t + 1.0               # t is held by u's node only
    del t, u                  # -> both nodes go, cascading through the operand the second one held
    z = x * 2.0
    assert len(ramba.DAG.pending) == 1 and ramba.DAG.pruned_count - p0 == 3
    assert onp.array_equal(z.asarray(), onp.arange(5000) * 2.0)
    assert RT.launches - l0 == 1   # (no launch over the 300-element shape)


def test_materialising_one_array_runs_only_its_dependencies(eng):
    rb, ramba, RT = eng
    x = rb.fromarray(onp.arange(4000, dtype=onp.float64))
    y = rb.fromarray(onp.arange(700, dtype=onp.float64))
    rb.sync()
    a = x + 1.0
    b = y * 2.0
    c = a * a
    l0 = RT.launches
    assert onp.array_equal(c.asarray(), (onp.arange(4000) + 1.0) ** 2)
    assert RT.launches - l0 == 1 and len(ramba.DAG.pending) == 1   # b's statement is still waiting
    assert onp.array_equal(b.asarray(), onp.arange(700) * 2.0)
    assert not ramba.DAG.pending


def test_interleaved_chains_fuse_per_shape(eng, monkeypatch):
    rb, ramba, RT = eng
    xs, ys = onp.arange(3000, dtype=onp.float64), onp.arange(900, dtype=onp.float64)

    def program():
        x, y = rb.fromarray(xs), rb.fromarray(ys)
        rb.sync()
        l0 = RT.launches
        outs = []
        for i in range(4):   # alternate between the two shapes
            x = x * 1.5 + float(i)
            y = y - float(i)
            outs += [x, y]
        rb.sync()
        return RT.launches - l0, [o.asarray() for o in outs]

    n_dag, got = program()
    monkeypatch.setattr(ramba, "NO_DAG", True)
    n_nodag, exp = program()
    assert n_dag == 2 and n_nodag == 8   # one fused op per shape vs a flush at every change of shape
    for g, e in zip(got, exp):
        assert onp.array_equal(g, e)
    ex, ey = xs.copy(), ys.copy()
    for i in range(4):
        ex = ex * 1.5 + float(i)
        ey = ey - float(i)
    assert onp.array_equal(got[-2], ex) and onp.array_equal(got[-1], ey)


def test_reordering_respects_hazards_on_arrays_and_views(eng):
    rb, ramba, RT = eng

    def program(np, arr):
        a = arr(onp.arange(100, dtype=onp.float64))
        b = arr(onp.arange(50, dtype=onp.float64))
        c = a + 1.0              # reads a
        d = b * 2.0
        a += 10.0                # write after read: c must see the old a
        e = a + c                # read after write: the new a
        b[10:20] = 7.0           # view write
        f = b + d
        v = a[5:55]              # view of a, same shape as b
        g = v * b
        a[5:55] = g + 1.0        # write through another view after the read through v
        h = a.sum()
        b *= 0.5
        i = (b + f).sum()
        return [x if isinstance(x, (float, onp.floating)) else onp.asarray(x.asarray() if hasattr(x, "asarray") else x)
                for x in (c, d, e, f, g, a, b, h, i)]

    got = program(rb, rb.fromarray)
    exp = program(onp, lambda x: x.copy())
    for k, (g, e) in enumerate(zip(got, exp)):
        assert onp.array_equal(onp.asarray(g), onp.asarray(e)), k


def test_in_place_updates_of_a_dead_handle_still_reach_the_views_base(eng):
    rb, ramba, RT = eng
    a = rb.zeros(200)
    v = a[50:150]
    v += 3.0
    del v                        # the view handle is gone, the statement is not: `a` observes it
    a[0:10] = 1.0
    e = onp.zeros(200)
    e[50:150] += 3.0
    e[0:10] = 1.0
    assert onp.array_equal(a.asarray(), e)


def test_reduction_temporary_is_elided_only_inside_one_fused_op(eng):
    """`(X*2.0 + 1.0).sum()` never stores its temporary - unless the statement limit cuts the fused op between the
    temporary and the reduction (then it must be stored: the second op reads it back)."""
    rb, ramba, RT = eng
    x = onp.arange(1000, dtype=onp.float32) % 7
    expect = float((x * 2.0 + 1.0).sum())
    for nodag in (False, True):
        ramba.NO_DAG = nodag
        try:
            for nfill in range(34, 42):
                X = rb.fromarray(x)
                keep = [X * float(i) for i in range(nfill)]
                assert float((X * 2.0 + 1.0).sum()) == expect, (nodag, nfill)
                del keep
        finally:
            ramba.NO_DAG = False


def test_long_dependency_chains_need_no_recursion(eng):
    rb, ramba, RT = eng
    a = rb.zeros(256)
    for _ in range(3000):
        a = a + 1.0
    assert onp.array_equal(a.asarray(), onp.full(256, 3000.0))
    assert not ramba.DAG.pending


def test_pending_graph_is_bounded(eng, monkeypatch):
    rb, ramba, RT = eng
    monkeypatch.setattr(ramba.DAG, "max_pending", 64)
    keep = [rb.arange(128) * float(i) for i in range(200)]
    assert len(ramba.DAG.pending) < 2 * 64
    for i in (0, 63, 64, 199):
        assert onp.array_equal(keep[i].asarray(), onp.arange(128) * float(i))


@pytest.mark.parametrize("case", _random_programs.CASES[:20], ids=lambda f: f.__name__)
def test_same_values_with_and_without_the_dag(eng, case, monkeypatch):
    rb, ramba, RT = eng
    got = case(rb)
    monkeypatch.setattr(ramba, "NO_DAG", True)
    exp = case(rb)
    for g, e in zip(got, exp):
        assert onp.array_equal(onp.asarray(g), onp.asarray(e), equal_nan=True)


def test_baseline_programs_lower_to_the_same_op_lists_with_and_without_the_dag(eng, monkeypatch, capsys):
    """The measured kernels are chosen from the op list: the DAG must hand the fuser the BASELINE programs (bench.py's
    configs 2-5, here at reduced size) in an order that lowers to the identical op lists and kernel plans."""
    import types

    import _oracle_backend
    import bench
    from ramba_b200 import common

    rb, ramba, RT = eng
    monkeypatch.setattr(common, "debug_showcode", True)
    scales = {2: 1, 3: 1 / 64, 4: 1 / 16, 5: 1 / 512}

    def run_all():
        capsys.readouterr()
        _oracle_backend.PLANS.clear()
        for cfg in (2, 3, 4, 5):
            wl = bench.CLASSES[cfg](rb, 1, 1 << 14, types.SimpleNamespace(scale=scales[cfg]))
            wl.step()
            wl.step()
            assert wl.check(), cfg
        code = capsys.readouterr().out
        # gids differ from run to run: compare the op lists, not the array numbering
        code = "\n".join(line for line in code.splitlines() if ": gid " not in line)
        return code, list(_oracle_backend.PLANS)

    code_dag, plans_dag = run_all()
    monkeypatch.setattr(ramba, "NO_DAG", True)
    code_nodag, plans_nodag = run_all()
    assert plans_dag == plans_nodag
    assert code_dag == code_nodag and "SINCOS" in code_dag


def _same(got, exp, name):
    assert len(got) == len(exp), name
    for i, (g, e) in enumerate(zip(got, exp)):
        g, e = onp.asarray(g), onp.asarray(e)
        assert g.shape == e.shape and onp.array_equal(g, e), "%s[%d]" % (name, i)


@pytest.mark.parametrize("chunk", range(10))
def test_fuzzed_programs_match_numpy(eng, chunk, monkeypatch):
    """tests/_dag_fuzz.py: long stretches of pending statements over three shapes with in-place updates, live views,
    rebinding, del and partial reads - bit-identical to NumPy in program order, with the DAG and without."""
    import _dag_fuzz

    rb, ramba, RT = eng
    for f in _dag_fuzz.CASES[chunk * 30:(chunk + 1) * 30]:
        exp = f(onp)
        _same(f(rb), exp, f.__name__)
        assert not ramba.DAG.in_evaluate
    monkeypatch.setattr(ramba, "NO_DAG", True)
    for f in _dag_fuzz.CASES[chunk * 30:chunk * 30 + 6]:
        _same(f(rb), f(onp), f.__name__ + " (NO_DAG)")


@pytest.mark.parametrize("chunk", range(8))
def test_programs_at_the_fusers_table_sizes(eng, chunk, monkeypatch):
    """tests/_limit_fuzz.py: long same-shaped runs with many live arrays, dying temporaries and reductions of temporaries at
    every distance from the 40-statement / 16-view limits - values as in NumPy, with the DAG and with RAMBA_NO_DAG."""
    import _limit_fuzz

    rb, ramba, RT = eng
    cases = _limit_fuzz.CASES[chunk * 20:(chunk + 1) * 20]
    for f in cases:
        _same(f(rb), f(onp), f.__name__)
    monkeypatch.setattr(ramba, "NO_DAG", True)
    for f in cases:
        _same(f(rb), f(onp), f.__name__ + " (NO_DAG)")


def test_instantiate_all_and_flags(eng):
    rb, ramba, RT = eng
    a = rb.arange(500) * 1.0
    b = a + 1.0
    c = rb.arange(200) * 3.0
    rb.instantiate_all(a, b, 5, "x")
    assert len(ramba.DAG.pending) == 2          # c's two statements are still waiting
    assert b.flags.writeable and b.flags["WRITEABLE"]
    b.flags.writeable = False
    with pytest.raises(ValueError):
        b += 1.0
    with pytest.raises(ValueError):
        b[3:9] = 0.0
    v = b[10:20]
    assert not v.flags.writeable
    with pytest.raises(ValueError):
        v.flags.writeable = True
    b.flags.writeable = True
    b += 1.0
    assert onp.array_equal(b.asarray(), onp.arange(500) + 2.0) and onp.array_equal(c.asarray(), onp.arange(200) * 3.0)


@pytest.mark.parametrize("chunk", range(5))
def test_mixed_statement_forms_match_numpy(eng, chunk, monkeypatch):
    """tests/_expr_fuzz.py: integer / float / bool arrays, where, comparisons, floor division, astype, in-place updates, strided /
    reversed / transposed / broadcast operands, sliced assignments (possibly overlapping), temporaries consumed statements
    later, reductions - exact against NumPy, with the DAG and without."""
    import _expr_fuzz

    rb, ramba, RT = eng
    cases = _expr_fuzz.CASES[chunk * 40:(chunk + 1) * 40]
    for f in cases:
        got, exp = f(rb), f(onp)
        _same(got, exp, f.__name__)
        assert all(onp.asarray(g).dtype == onp.asarray(e).dtype for g, e in zip(got, exp)), f.__name__
    monkeypatch.setattr(ramba, "NO_DAG", True)
    for f in cases[:10]:
        _same(f(rb), f(onp), f.__name__ + " (NO_DAG)")


def test_fuzzer_finds_of_this_round_stay_fixed(eng):
    """Seeds beyond the committed ranges that exposed the lowering's store-order bugs."""
    import _dag_fuzz

    rb, ramba, RT = eng
    for seed in (592, 1660):
        _same(_dag_fuzz.dag_program(rb, seed), _dag_fuzz.dag_program(onp, seed), "dag_program_%d" % seed)


@pytest.mark.parametrize("chunk", range(4))
def test_trig_mask_and_float32_programs(eng, chunk, monkeypatch):
    """tests/_expr_fuzz.py::trig_mask_program: sin / cos pairs around in-place updates, boolean-mask assignment (also into
    arrays that live in registers, with masks materialised by an earlier flush), masked sums, float32 next to float64."""
    import _expr_fuzz

    rb, ramba, RT = eng

    def close(got, exp, name):
        assert len(got) == len(exp), name
        for i, (g, e) in enumerate(zip(got, exp)):
            assert g.shape == e.shape and g.dtype == e.dtype and onp.allclose(g, e, rtol=1e-12, atol=1e-12), "%s[%d]" % (name, i)

    cases = _expr_fuzz.TRIG_CASES[chunk * 30:(chunk + 1) * 30]
    for f in cases:
        close(f(rb), f(onp), f.__name__)
    monkeypatch.setattr(ramba, "NO_DAG", True)
    for f in cases[:8]:
        close(f(rb), f(onp), f.__name__ + " (NO_DAG)")


@pytest.mark.parametrize("nodag", [False, True])
def test_masked_assignment_corner_cases(eng, nodag, monkeypatch):
    """Found by the trig / mask fuzzer: (1) a mask that is a STORED bool array (materialised by an earlier flush) used to fail
    in the lowering ('mask not in a register'); (2) a masked assignment into an array that never leaves the registers
    overwrote it whole (`t = a + b; t[m] = 0.5; r = cos(t)` with t dead gave cos(0.5) everywhere)."""
    rb, ramba, RT = eng
    monkeypatch.setattr(ramba, "NO_DAG", nodag)
    a = onp.arange(300.0) - 150.0
    x = rb.fromarray(a)
    m = x > 0.0
    rb.sync()
    x[m] = 0.5
    e = a.copy()
    e[a > 0] = 0.5
    assert onp.array_equal(x.asarray(), e)
    x, y = rb.fromarray(a), rb.fromarray(a * 0.5)
    rb.sync()
    t = x + y
    t[t > 30.0] = 0.5
    r = rb.cos(t)
    del t
    et = a + a * 0.5
    et[et > 30.0] = 0.5
    assert onp.allclose(r.asarray(), onp.cos(et), rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("chunk", range(3))
def test_writes_through_views_match_numpy(eng, chunk, monkeypatch):
    """tests/_expr_fuzz.py::view_program: strided / reversed / transposed targets, row and column broadcast assignment,
    shifted windows of one array (stencil form, also in place), slices of slices, axis sums fed back - exact."""
    import _expr_fuzz

    rb, ramba, RT = eng
    cases = _expr_fuzz.VIEW_CASES[chunk * 40:(chunk + 1) * 40]
    for f in cases:
        _same(f(rb), f(onp), f.__name__)
    monkeypatch.setattr(ramba, "NO_DAG", True)
    for f in cases[:8]:
        _same(f(rb), f(onp), f.__name__ + " (NO_DAG)")


@pytest.mark.parametrize("nodag", [False, True])
def test_zero_d_operands_are_read_when_the_statement_is_written(eng, nodag, monkeypatch):
    """0-d arrays keep their value on the host: a statement uses the value the array has when the statement is WRITTEN,
    not the one it has when the fused op finally runs."""
    rb, ramba, RT = eng
    monkeypatch.setattr(ramba, "NO_DAG", nodag)
    x = rb.fromarray(onp.arange(200.0))
    z = rb.array(3.0)
    y = x * z
    w = rb.where(x > 100.0, x, x) + z
    z[()] = 5.0
    v = x - z
    assert onp.array_equal(y.asarray(), onp.arange(200.0) * 3.0) and onp.array_equal(w.asarray(), onp.arange(200.0) + 3.0)
    assert onp.array_equal(v.asarray(), onp.arange(200.0) - 5.0)


@pytest.mark.parametrize("chunk", range(2))
def test_library_calls_inside_pending_stretches(eng, chunk, monkeypatch):
    """tests/_expr_fuzz.py::api_program: cumsum / concatenate / stack / pad / reshape_copy / clip / where / transposes /
    unit-dim reshapes / broadcast_to / astype with the source updated in place right after the call - exact."""
    import _expr_fuzz

    rb, ramba, RT = eng
    cases = _expr_fuzz.API_CASES[chunk * 50:(chunk + 1) * 50]
    for f in cases:
        _same(f(rb), f(onp), f.__name__)
    monkeypatch.setattr(ramba, "NO_DAG", True)
    for f in cases[:8]:
        _same(f(rb), f(onp), f.__name__ + " (NO_DAG)")


def test_shapes_around_the_distribution_threshold(eng, monkeypatch):
    """tests/_expr_fuzz.py::shape_program: single-owner arrays (< 100 elements), single elements, empty arrays, 3-D / 4-D,
    unit dims; int64 / float64 / bool / int32 - exact."""
    import _expr_fuzz

    rb, ramba, RT = eng
    for f in _expr_fuzz.SHAPE_CASES[:60]:
        _same(f(rb), f(onp), f.__name__)
    monkeypatch.setattr(ramba, "NO_DAG", True)
    for f in _expr_fuzz.SHAPE_CASES[60:70]:
        _same(f(rb), f(onp), f.__name__ + " (NO_DAG)")


def test_skeletons_inside_pending_stretches(eng, monkeypatch):
    """tests/_expr_fuzz.py::skeleton_program: smap / smap_index / sreduce / sstencil / stencil / scumulative / fromfunction /
    triu with their sources updated right after the call - exact."""
    import _expr_fuzz

    rb, ramba, RT = eng
    for f in _expr_fuzz.SKELETON_CASES[:50]:
        _same(f(rb), f(onp), f.__name__)
    monkeypatch.setattr(ramba, "NO_DAG", True)
    for f in _expr_fuzz.SKELETON_CASES[50:58]:
        _same(f(rb), f(onp), f.__name__ + " (NO_DAG)")


def test_operands_with_different_partitions(eng):
    """tests/_expr_fuzz.py::partition_program on one rank (its point is the multi-rank worker: row-split, column-split and
    default-split operands in one statement; tests/test_multirank_gloo.py runs it on worlds 2, 3, 4 and 8)."""
    import _expr_fuzz

    rb, ramba, RT = eng
    for f in _expr_fuzz.PARTITION_CASES[:20]:
        _same(f(rb), f(onp), f.__name__)


def test_reductions_in_every_form(eng, monkeypatch):
    """tests/_expr_fuzz.py::reduction_program: sum / min / max / prod / all / any over all axes, one, several, keepdims, of
    strided / reversed / transposed views and windows, of dying temporaries, chained, next to in-place updates - exact."""
    import _expr_fuzz

    rb, ramba, RT = eng
    for f in _expr_fuzz.REDUCTION_CASES[:50]:
        _same(f(rb), f(onp), f.__name__)
    monkeypatch.setattr(ramba, "NO_DAG", True)
    for f in _expr_fuzz.REDUCTION_CASES[50:58]:
        _same(f(rb), f(onp), f.__name__ + " (NO_DAG)")


@pytest.mark.parametrize("chunk", range(2))
def test_everything_on_one_pool(eng, chunk, monkeypatch):
    """tests/_expr_fuzz.py::mixed_program: three partitions + padded shards, arithmetic / in-place / windows / view targets /
    masks / sin-cos / held temporaries / library calls / skeletons / reductions fed back / rebinding, del, sync, partial
    reads - all on shared arrays."""
    import _expr_fuzz

    rb, ramba, RT = eng

    def close(got, exp, name):
        assert len(got) == len(exp), name
        for i, (g, e) in enumerate(zip(got, exp)):
            assert g.shape == e.shape and g.dtype == e.dtype and onp.allclose(g, e, rtol=1e-11, atol=1e-11), "%s[%d]" % (name, i)

    cases = _expr_fuzz.MIXED_CASES[chunk * 40:(chunk + 1) * 40]
    for f in cases:
        close(f(rb), f(onp), f.__name__)
    monkeypatch.setattr(ramba, "NO_DAG", True)
    for f in cases[:6]:
        close(f(rb), f(onp), f.__name__ + " (NO_DAG)")
