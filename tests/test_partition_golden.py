"""Partition algebra against the reference's own pure functions (fixtures produced by
tests/golden/make_golden.py from ramba.common / ramba.shardview_array)."""
import json
import os

import numpy as onp
import pytest

from ramba_b200 import partition, shardview as sv

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "partition_golden.json")


@pytest.fixture(scope="module")
def G():
    with open(GOLD) as f:
        return json.load(f)


def dist_for(W, shape):
    return sv.divisions_to_distribution(partition.compute_regular_schedule(W, tuple(shape)))


def same_part(mine, ref, what):
    """Non-empty parts must agree in every field; empty parts only in emptiness."""
    m = mine.to_lists()
    ref_empty = any(x == 0 for x in ref[0])
    assert sv.is_empty(mine) == ref_empty, "%s: emptiness %r vs reference %r" % (what, m, ref)
    if not ref_empty:
        assert m == ref, "%s: %r vs reference %r" % (what, m, ref)


def test_work_division(G):
    """Block boundaries for every (W, shape), incl. the BASELINE shapes."""
    bad = []
    for case in G["schedule"]:
        W, shape = case["W"], tuple(case["shape"])
        mine = partition.compute_regular_schedule(W, shape)
        if mine.tolist() != case["divisions"]:
            bad.append((W, shape))
    assert not bad, "work division differs from the reference for %r" % (bad,)


def _sl(x):
    return tuple(slice(a, b, c) for a, b, c in x)


def test_slice_distribution(G):
    for case in G["slice"]:
        W, shape = case["W"], tuple(case["shape"])
        D = dist_for(W, shape)
        S = sv.slice_distribution(_sl(case["slices"]), D)
        if "slices2" in case:
            S = sv.slice_distribution(_sl(case["slices2"]), S)
        for i in range(W):
            same_part(S[i], case["dist"][i], "slice %r of %r W=%d worker %d" % (case["slices"], shape, W, i))


def test_broadcast(G):
    for case in G["broadcast"]:
        W, shape = case["W"], tuple(case["shape"])
        B = sv.broadcast(dist_for(W, shape), case["bdims"], tuple(case["size"]))
        for i in range(W):
            same_part(B[i], case["dist"][i], "broadcast %r W=%d worker %d" % (shape, W, i))


def test_intersect_and_compat(G):
    for case in G["intersect"]:
        W, shape = case["W"], tuple(case["shape"])
        big = (7, shape[0])
        B = sv.broadcast(dist_for(W, shape), [True, False], big)
        D2 = dist_for(W, big)
        part = sv.intersect(B[case["i"]], D2[case["j"]])
        ref = case["part"]
        ref_empty = any(x == 0 for x in ref[0])
        assert sv.is_empty(part) == ref_empty
        if not ref_empty:
            assert part.to_lists()[:2] == ref[:2]
        assert sv.is_compat(sv.clean_range(D2[case["j"]]), B[case["j"]]) == case["compat"]


def test_remap_axis(G):
    for case in G["remap"]:
        W, shape = case["W"], tuple(case["shape"])
        ns, R = sv.remap_axis(shape, dist_for(W, shape), case["perm"])
        assert list(ns) == case["new_shape"]
        for i in range(W):
            same_part(R[i], case["dist"][i], "transpose %r W=%d worker %d" % (shape, W, i))


def test_reduce_axes(G):
    for case in G["reduce"]:
        W, shape = case["W"], tuple(case["shape"])
        D = dist_for(W, shape)
        if any(sv.is_empty(d) for d in D):
            continue  # the reference's treatment of empty workers here is a quirk we do not follow
        rsz, rdist, bdist = sv.reduce_axes(shape, D, case["axes"])
        assert [int(x) for x in rsz] == case["rsz"]
        for i in range(W):
            # which partial slot a division gets along a reduced axis is `list(set(starts))` order
            # inside a Numba-compiled function in the reference (ramba/shardview_array.py:1059) —
            # arbitrary but consistent; we number divisions in ascending order instead.  Everything
            # else must agree, and our numbering must be a bijection onto range(rsz[j]).
            ref = [list(x) for x in case["rdist"][i]]
            mine = rdist[i].copy()
            for j in case["axes"]:
                ref[1][j] = int(mine.start[j])
            same_part(mine, ref, "rdist %r axes %r W=%d worker %d" % (shape, case["axes"], W, i))
            same_part(bdist[i], case["bdist"][i], "bdist %r axes %r W=%d worker %d" % (shape, case["axes"], W, i))
        for j in case["axes"]:
            assert sorted(set(int(r.start[j]) for r in rdist)) == list(range(int(rsz[j])))


def test_range_splits(G):
    for case in G["splits"]:
        W, shape = case["W"], tuple(case["shape"])
        D = dist_for(W, shape)
        sp = sv.get_range_splits_list([sv.clean_range(d) for d in D])
        mine = sorted([s.to_lists()[:2] for s in sp])
        ref = sorted([x for x in case["splits"] if all(v > 0 for v in x[0])])
        mine = [x for x in mine if all(v > 0 for v in x[0])]
        assert mine == ref, "range splits of %r W=%d" % (shape, W)
