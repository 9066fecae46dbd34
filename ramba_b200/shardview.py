"""Partition algebra: which part of every array view each worker (GPU) holds, and where.

Fresh implementation of the model of ramba/shardview_array.py (function names kept so that the
rest of the engine and the parity tests read like the reference):

A *shardview* describes, for ONE worker, a k-dim box of a view:
    size[k]        extent of the box (all zeros = this worker holds nothing),
    start[k]       first index of the box in the VIEW's own global coordinates,
    axis_map[k]    which axis of the worker-local buffer view-axis d walks (-1: broadcast axis),
    steps[k]       buffer step per index step (negative for reversed views),
    base_offset[m] lowest local-buffer coordinate the box touches, per buffer axis
(ramba/shardview_array.py:32-70).  Element `start + i` of the box sits at buffer coordinate
`base_offset[a] + i*step` (step > 0) or `base_offset[a] + (size-1-i)*|step|` (step < 0) along
a = axis_map[d] (ramba/shardview_array.py:220-233).  A *distribution* is one shardview per worker.

Unlike the reference (packed int32/int64 matrices + Numba), a shardview here is a tiny object with
int64 NumPy vectors; indices are always 64-bit.
"""
import numpy as np

from . import common, partition

I64 = np.int64


class ShardView:
    __slots__ = ("size", "start", "axis_map", "steps", "base_offset", "_k")

    def __init__(self, size, start=None, base_offset=None, axis_map=None, steps=None):
        size = np.asarray(size, dtype=I64)
        k = len(size)
        # an empty part is all-zero in size (ramba/shardview_array.py:51-52)
        self.size = size.copy() if np.all(size > 0) else np.zeros(k, dtype=I64)
        self.start = np.zeros(k, dtype=I64) if start is None else np.asarray(start, dtype=I64).copy()
        self.axis_map = np.arange(k, dtype=I64) if axis_map is None else np.asarray(axis_map, dtype=I64).copy()
        self.steps = np.ones(k, dtype=I64) if steps is None else np.asarray(steps, dtype=I64).copy()
        self.base_offset = np.zeros(k, dtype=I64) if base_offset is None else np.asarray(base_offset, dtype=I64).copy()
        self._k = None

    def key(self):
        """Hashable identity (box, addressing); cached — shardviews are not mutated once in use."""
        if self._k is None:
            self._k = (tuple(self.size.tolist()), tuple(self.start.tolist()), tuple(self.axis_map.tolist()),
                       tuple(self.steps.tolist()), tuple(self.base_offset.tolist()))
        return self._k

    def copy(self):
        return ShardView(self.size, self.start, self.base_offset, self.axis_map, self.steps)

    @property
    def stop(self):
        return self.start + self.size

    def to_lists(self):
        return [self.size.tolist(), self.start.tolist(), self.axis_map.tolist(), self.steps.tolist(),
                self.base_offset.tolist()]

    def __repr__(self):
        return "SV(size=%s,start=%s,map=%s,steps=%s,bo=%s)" % tuple(self.to_lists())


def shardview(size, index_start=None, base_offset=None, axis_map=None, steps=None):
    return ShardView(size, index_start, base_offset, axis_map, steps)


# ---- accessors (reference spelling)
def _size(sv):
    return sv.size


def _index_start(sv):
    return sv.start


_start = _index_start
get_start = _index_start
get_size = _size


def _stop(sv):
    return sv.start + sv.size


def _axis_map(sv):
    return sv.axis_map


def _steps(sv):
    return sv.steps


def _base_offset(sv):
    return sv.base_offset


def len_size(sv):
    return len(sv.size)


def len_base_offset(sv):
    return len(sv.base_offset)


# ---- predicates
def is_eq(a, b):
    return a is b or a.key() == b.key()


def is_empty(sv):
    return 0 in sv.key()[0]


def is_compat(a, b):
    """Same box (size and start) — the operand is aligned with the iteration range
    (ramba/shardview_array.py:183-186)."""
    ka, kb = a.key(), b.key()
    return ka[0] == kb[0] and ka[1] == kb[1]


def overlaps(a, b):
    s1, e1, s2, e2 = a.start, _stop(a), b.start, _stop(b)
    return bool(np.all(((s1 <= s2) & (s2 < e1)) | ((s2 <= s1) & (s1 < e2))))


_MEMO_MAX = 1 << 14


def _memo2(fn):
    """Memoise a pure function of two shardviews on their keys (shardviews are immutable once in use; a flush asks the
    same geometric questions every iteration of a loop)."""
    cache = {}

    def wrapped(a, b):
        k = (a.key(), b.key())
        r = cache.get(k, cache)
        if r is cache:
            if len(cache) >= _MEMO_MAX:
                cache.clear()
            r = cache[k] = fn(a, b)
        return r

    wrapped.__name__, wrapped.__doc__, wrapped.cache = fn.__name__, fn.__doc__, cache
    return wrapped


@_memo2
def contains(a, b):
    return (not is_empty(b)) and bool(np.all((a.start <= b.start) & (_stop(b) <= _stop(a))))


def has_index(sv, index):
    index = np.asarray(index, dtype=I64)
    return bool(np.all((sv.start <= index) & (index < _stop(sv))))


_IDENT = {}


def _is_clean(sv):
    k = sv.key()
    n = len(k[0])
    ident = _IDENT.get(n)
    if ident is None:
        ident = _IDENT[n] = (tuple(range(n)), (1,) * n, (0,) * n)
    return k[2] == ident[0] and k[3] == ident[1] and k[4] == ident[2]


_clean_cache = {}


def clean_range(sv):
    """Just the box: drop offset, axis map and steps (ramba/shardview_array.py:207-208).
    An already clean shardview is returned as is (shardviews are immutable once in use)."""
    if _is_clean(sv):
        return sv
    k = sv.key()
    r = _clean_cache.get(k)
    if r is None:
        if len(_clean_cache) >= _MEMO_MAX:
            _clean_cache.clear()
        r = _clean_cache[k] = ShardView(sv.size, sv.start)
    return r


# ---- index -> buffer coordinates
def index_to_base(sv, index, end=False):
    """Buffer coordinate (per buffer axis) of view index `index` (ramba/shardview_array.py:221-241).
    Buffer axes no view axis maps to keep base_offset (+1 if `end`)."""
    index = np.asarray(index, dtype=I64)
    off = index - sv.start
    neg = sv.steps < 0
    off = np.where(neg, np.maximum(-1, sv.size - 1 - off), off) * np.abs(sv.steps)
    out = sv.base_offset + (1 if end else 0)
    out = out.copy()
    for d, a in enumerate(sv.axis_map):
        if a >= 0:
            out[a] = sv.base_offset[a] + off[d]
    return out


def get_base_steps(sv):
    st = np.ones(len(sv.base_offset), dtype=I64)
    for d, a in enumerate(sv.axis_map):
        if a >= 0:
            st[a] = sv.steps[d]
    return st


def to_slice(sv):
    """Global-index slices of the box (for scattering shards into a host array,
    ramba/shardview_array.py:328-335).  Only meaningful for clean (step 1) boxes here."""
    return tuple(slice(int(s), int(s + n)) for s, n in zip(sv.start, sv.size))


def to_base_slice(sv):
    """Slices into the worker-local buffer covering the box (ramba/shardview_array.py:347-369)."""
    m = len(sv.base_offset)
    lo = sv.base_offset.copy()
    ext = np.ones(m, dtype=I64)
    st = np.ones(m, dtype=I64)
    for d, a in enumerate(sv.axis_map):
        if a >= 0:
            ext[a] = (sv.size[d] - 1) * abs(sv.steps[d]) + 1
            st[a] = sv.steps[d]
    out = []
    for j in range(m):
        if st[j] > 0:
            out.append(slice(int(lo[j]), int(lo[j] + ext[j]), int(st[j])))
        else:
            hi = int(lo[j] + ext[j] - 1)
            stop = int(lo[j] - 1)
            out.append(slice(hi, stop if stop >= 0 else None, int(st[j])))
    return tuple(out)


# ---- slicing
def _map_one(a, b, c, s, e):
    """Indices a, a+c, a+2c, ... (towards b, exclusive) of the parent that fall in [s, e).
    Returns (first_parent, last_parent, count, first_new_index)."""
    if c > 0:
        n_total = max(0, -(-(b - a) // c))
        lo = max(a, s)
        n0 = -(-(lo - a) // c)
        hi = min(b, e)
        n1 = -(-(hi - a) // c)  # first n with p >= hi
        n1 = min(n1, n_total)
        cnt = max(0, n1 - n0)
        return a + n0 * c, a + (n0 + cnt - 1) * c, cnt, n0
    cc = -c
    n_total = max(0, -(-(a - b) // cc))
    # p = a - n*cc must satisfy s <= p < e
    hi = min(a, e - 1)
    n0 = -(-(a - hi) // cc)
    lo = max(b + 1, s)
    n1 = (a - lo) // cc + 1  # one past the last n with p >= lo
    n1 = min(n1, n_total)
    cnt = max(0, n1 - n0)
    return a - n0 * cc, a - (n0 + cnt - 1) * cc, cnt, n0


def mapslice(sv, sl):
    """Part of `sv` selected by the canonical slices `sl` (start/stop/step resolved against the
    view's shape; stop may be -1 for reversed slices running to index 0), re-indexed in the sliced
    view's coordinates (ramba/shardview_array.py:474-492)."""
    k = len(sv.size)
    size = np.zeros(k, dtype=I64)
    start = np.zeros(k, dtype=I64)
    steps = np.zeros(k, dtype=I64)
    first = np.zeros(k, dtype=I64)
    last = np.zeros(k, dtype=I64)
    s, e = sv.start, _stop(sv)
    for d in range(k):
        a, b, c = sl[d].start, sl[d].stop, sl[d].step
        c = 1 if c is None else c
        p0, p1, cnt, n0 = _map_one(int(a), int(b), int(c), int(s[d]), int(e[d]))
        size[d], start[d], steps[d] = cnt, n0, sv.steps[d] * c
        first[d], last[d] = p0, p1
    if np.any(size <= 0) or is_empty(sv):
        return ShardView(np.zeros(k, dtype=I64), start, sv.base_offset, sv.axis_map, np.where(steps == 0, 1, steps))
    b0 = index_to_base(sv, first)
    b1 = index_to_base(sv, last)
    return ShardView(size, start, np.minimum(b0, b1), sv.axis_map, steps)


def mapsv(sv, box):
    """Part of `sv` inside the clean box `box` (same coordinates), keeping sv's addressing
    (ramba/shardview_array.py:496-525 restricted to step-1 boxes)."""
    return mapslice_keep(sv, box.start, _stop(box))


def mapslice_keep(sv, lo, hi):
    """sv restricted to global indices [lo, hi) per dim, WITHOUT re-basing the coordinates."""
    key = (sv.key(), tuple(np.asarray(lo).tolist()), tuple(np.asarray(hi).tolist()))
    r = _keep_cache.get(key)
    if r is None:
        if len(_keep_cache) >= _MEMO_MAX:
            _keep_cache.clear()
        r = _keep_cache[key] = _mapslice_keep(sv, lo, hi)
    return r


_keep_cache = {}


def _mapslice_keep(sv, lo, hi):
    k = len(sv.size)
    s = np.maximum(sv.start, lo)
    e = np.minimum(_stop(sv), hi)
    if np.any(e <= s) or is_empty(sv):
        return ShardView(np.zeros(k, dtype=I64), s, sv.base_offset, sv.axis_map, sv.steps)
    b0 = index_to_base(sv, s)
    b1 = index_to_base(sv, e - 1)
    return ShardView(e - s, s, np.minimum(b0, b1), sv.axis_map, sv.steps)


@_memo2
def intersect(a, b):
    """Box of `a` clipped to the box of `b` (ramba/shardview_array.py:530-538)."""
    s = np.minimum(np.maximum(b.start, a.start), _stop(a))
    e = np.minimum(np.maximum(_stop(b), a.start), _stop(a))
    return ShardView(e - s, s, np.zeros(len(a.base_offset), dtype=I64), a.axis_map)


def union(a, b):
    s = np.minimum(a.start, b.start)
    e = np.maximum(_stop(a), _stop(b))
    return ShardView(e - s, s, np.zeros(len(a.base_offset), dtype=I64), a.axis_map)


def as_base(sv, part):
    """The buffer-coordinate box (clean shardview over buffer axes) of `part`, a clean box in
    sv's view coordinates (ramba/shardview_array.py:293-303)."""
    s = index_to_base(sv, part.start)
    e = index_to_base(sv, _stop(part) - 1)
    lo = np.minimum(s, e)
    hi = np.maximum(s, e)
    return ShardView(hi - lo + 1, lo)


# ---- distributions
def clean_dist(dist):
    return [clean_range(s) for s in dist]


def compatible_distributions(d1, d2):
    if d1 is d2:
        return True
    if len(d1) != len(d2):
        return False
    for a, b in zip(d1, d2):
        if is_empty(a) and is_empty(b):
            continue
        if not is_compat(a, b):
            return False
    return True


def dist_is_eq(d1, d2):
    if d1 is d2:
        return True
    if len(d1) != len(d2):
        return False
    for a, b in zip(d1, d2):
        if a is not b and a.key() != b.key():
            return False
    return True


def dist_has_neg_step(dist):
    d0 = dist[0]
    return bool(np.any((d0.axis_map >= 0) & (d0.steps < 0)))


_dist_fn_cache = {}


def _dist_key(dist):
    return tuple([s.key() for s in dist])


def _memo_dist(tag, dist, args, compute):
    """Memoise a pure function of (a distribution, hashable arguments) that returns distributions: shardviews are
    immutable once in use, so the cached objects are shared; the LIST is copied (callers may own and re-point it).
    An iterative program slices / broadcasts / reduces fresh arrays with the same partitions every step."""
    try:
        key = (tag, _dist_key(dist), args)
        hit = _dist_fn_cache.get(key)
    except TypeError:
        return compute()
    if hit is None:
        if len(_dist_fn_cache) >= _MEMO_MAX:
            _dist_fn_cache.clear()
        hit = _dist_fn_cache[key] = compute()
    return hit


def slice_distribution(sl, dist):
    args = tuple([(x.start, x.stop, x.step) for x in sl])
    return list(_memo_dist("slice", dist, args, lambda: tuple([mapslice(s, sl) for s in dist])))


def find_index(dist, index):
    for i, s in enumerate(dist):
        if has_index(s, index):
            return i
    return None


def get_overlaps(k, dist1, dist2):
    return [i for i in range(len(dist1)) if overlaps(dist1[i], dist2[k]) or overlaps(dist1[k], dist2[i])]


def divisions_to_distribution(divs, base_offset=None, axis_map=None):
    out = []
    for i in range(divs.shape[0]):
        out.append(ShardView(divs[i, 1] - divs[i, 0] + 1, divs[i, 0], None if base_offset is None else base_offset[i], axis_map))
    return out


def distribution_to_divisions(dist):
    k = len(dist[0].size)
    out = np.empty((len(dist), 2, k), dtype=I64)
    for i, s in enumerate(dist):
        out[i, 0] = s.start
        out[i, 1] = _stop(s) - 1
    return out


_dist_cache = {}


def default_distribution(size, dims_do_not_distribute=(), dist_dims=None, num_workers=None):
    """Block distribution of an array of shape `size` over all workers
    (ramba/shardview_array.py:908-935; cached like dist_cache there)."""
    W = common.num_workers if num_workers is None else num_workers
    size = tuple(int(s) for s in size)
    ck = (size, tuple(dims_do_not_distribute or ()), tuple(dist_dims) if isinstance(dist_dims, (list, tuple)) else dist_dims, W)
    hit = _dist_cache.get(ck)
    if hit is not None:
        return list(hit)
    out = _default_distribution(size, dims_do_not_distribute, dist_dims, W)
    _dist_cache[ck] = tuple(out)
    return out


_plain_dist_cache = {}


def default_distribution_of(shape):
    """default_distribution(shape) for a shape that is a tuple of ints already and no options (every result array)."""
    ck = (shape, common.num_workers)
    hit = _plain_dist_cache.get(ck)
    if hit is None:
        if len(_plain_dist_cache) >= 4096:
            _plain_dist_cache.clear()
        hit = _plain_dist_cache[ck] = tuple(default_distribution(shape))
    return list(hit)


def _default_distribution(size, dims_do_not_distribute, dist_dims, W):
    k = len(size)
    if isinstance(dist_dims, int):
        dist_dims = [dist_dims]
    if isinstance(dist_dims, (list, tuple)):
        dims_do_not_distribute = [i for i in range(k) if i not in dist_dims]
    if dist_dims is None and common.do_not_distribute(size):
        divs = partition.make_uni_divisions(W, size, 0)
    else:
        divs = partition.compute_regular_schedule(W, size, tuple(dims_do_not_distribute or ()))
    return divisions_to_distribution(divs)


def make_uni_dist(size, node=0, num_workers=None):
    W = common.num_workers if num_workers is None else num_workers
    return divisions_to_distribution(partition.make_uni_divisions(W, tuple(size), node))


def broadcast(distribution, broadcasted_dims, size):
    """Distribution of `distribution`'s array viewed at shape `size`: broadcast dims get
    axis_map -1 and every worker 'holds' their full extent (ramba/shardview_array.py:978-1017)."""
    args = (tuple([bool(b) for b in broadcasted_dims]), tuple([int(x) for x in size]))
    return list(_memo_dist("bcast", distribution, args, lambda: tuple(_broadcast(distribution, broadcasted_dims, size))))


def _broadcast(distribution, broadcasted_dims, size):
    old_k = len(distribution[0].size)
    new_dims = len(size) - old_k
    k = len(size)
    d0 = distribution[0]
    amap = np.array([-1 if broadcasted_dims[j] else d0.axis_map[j - new_dims] for j in range(k)], dtype=I64)
    out = []
    for sv in distribution:
        nsz = np.array([(size[j] if (j < new_dims or sv.size[j - new_dims] > 0) else 0) if broadcasted_dims[j]
                        else sv.size[j - new_dims] for j in range(k)], dtype=I64)
        nst = np.array([0 if broadcasted_dims[j] else sv.start[j - new_dims] for j in range(k)], dtype=I64)
        nstep = np.array([1 if broadcasted_dims[j] else sv.steps[j - new_dims] for j in range(k)], dtype=I64)
        if is_empty(sv):
            nsz = np.zeros(k, dtype=I64)
        out.append(ShardView(nsz, nst, sv.base_offset, amap, nstep))
    return out


def expand_unit_dims(size, distribution, axes):
    """Distribution of the same array viewed with unit dims inserted at positions `axes` of the NEW shape
    (no storage behind them: axis_map -1, like a broadcast dim of extent 1).  The reference gets there through
    reshape (ramba/ramba.py:9438-9453, 9125-9238)."""
    args = (tuple([int(x) for x in size]), tuple([int(a) for a in axes]))
    new_size, out = _memo_dist("expand", distribution, args, lambda: _expand_unit_dims(size, distribution, axes))
    return new_size, list(out)


def _expand_unit_dims(size, distribution, axes):
    k = len(size) + len(axes)
    old_pos = [j for j in range(k) if j not in axes]
    d0 = distribution[0]
    amap = np.full(k, -1, dtype=I64)
    amap[old_pos] = d0.axis_map
    new_size = [1] * k
    for j, o in zip(old_pos, range(len(size))):
        new_size[j] = size[o]
    out = []
    for sv in distribution:
        nsz = np.ones(k, dtype=I64)
        nst = np.zeros(k, dtype=I64)
        nstep = np.ones(k, dtype=I64)
        nsz[old_pos], nst[old_pos], nstep[old_pos] = sv.size, sv.start, sv.steps
        if is_empty(sv):
            nsz = np.zeros(k, dtype=I64)
        out.append(ShardView(nsz, nst, sv.base_offset, amap, nstep))
    return tuple(new_size), out


def remap_axis(size, distribution, newmap):
    """Re-order / drop axes (transpose family, ramba/shardview_array.py:1024-1042)."""
    args = (tuple([int(x) for x in size]), tuple([int(a) for a in newmap]))
    new_size, out = _memo_dist("remap", distribution, args, lambda: _remap_axis(size, distribution, newmap))
    return new_size, list(out)


def _remap_axis(size, distribution, newmap):
    old = distribution[0].axis_map
    amap = np.array([old[i] for i in newmap], dtype=I64)
    new_size = tuple(size[i] for i in newmap)
    out = []
    for sv in distribution:
        nm = list(newmap)
        out.append(ShardView(sv.size[nm], sv.start[nm], sv.base_offset, amap, sv.steps[nm]))
    return new_size, out


def reduce_axes(size, dist, axes):
    """Distributions for stage 1 of a reduction over `axes`
    (ramba/shardview_array.py:1046-1066): the partial array keeps ONE element per division along
    every reduced axis (shape rsz, distribution rdist); bdist views it back at the source's shape
    with the reduced axes broadcast."""
    args = (tuple([int(x) for x in size]), tuple([int(a) for a in axes]))
    rsz, rdist, bdist = _memo_dist("reduce", dist, args, lambda: _reduce_axes(size, dist, axes))
    return rsz, list(rdist), list(bdist)


def _reduce_axes(size, dist, axes):
    rdist = [ShardView(s.size, s.start) for s in dist]  # fresh objects: they are edited below
    bdist = [ShardView(s.size, s.start) for s in dist]
    rsz = list(size)
    for j in axes:
        divs = sorted(set(int(dist[i].start[j]) for i in range(len(dist)) if not is_empty(dist[i])))
        for i in range(len(rdist)):
            if is_empty(dist[i]):
                continue
            rdist[i].start[j] = divs.index(int(rdist[i].start[j]))
            rdist[i].size[j] = 1
            bdist[i].axis_map[j] = -1
        for i in range(len(bdist)):
            bdist[i].axis_map[j] = -1
        rsz[j] = len(divs)
    for sv in rdist + bdist:
        sv._k = None
        if (sv.size == 0).any():
            sv.size[:] = 0
    return tuple(rsz), rdist, bdist


def reduce_all_axes(size, dist):
    return reduce_axes(size, dist, tuple(range(len(size))))


def get_range_splits_list(svl):
    """Cut the space covered by the boxes in `svl` along every box boundary: the cartesian product
    of per-axis intervals between consecutive boundaries (ramba/shardview_array.py:697-720).
    Inside one resulting range every operand piece is either wholly present or absent."""
    key = tuple([s.key()[:2] for s in svl])
    hit = _splits_cache.get(key)
    if hit is not None:
        return list(hit)
    out = _range_splits(svl)
    if len(_splits_cache) >= 4096:
        _splits_cache.clear()
    _splits_cache[key] = tuple(out)
    return out


_splits_cache = {}


def _range_splits(svl):
    k = len(svl[0].size)
    cuts = []
    for d in range(k):
        pts = sorted(set(int(x) for s in svl if not is_empty(s) for x in (s.start[d], s.start[d] + s.size[d])))
        cuts.append(pts)
    out = []
    if any(len(c) < 2 for c in cuts):
        return out
    import itertools

    for combo in itertools.product(*[range(len(c) - 1) for c in cuts]):
        s = np.array([cuts[d][combo[d]] for d in range(k)], dtype=I64)
        e = np.array([cuts[d][combo[d] + 1] for d in range(k)], dtype=I64)
        out.append(ShardView(e - s, s))
    return out


def compute_from_border(size, distribution, border):
    """Which inclusive [start],[end] regions every worker must fetch from / send to every other
    worker for a halo of width `border` (ramba/shardview_array.py:1069-1136)."""
    divs = distribution_to_divisions(distribution)
    W, k = divs.shape[0], len(size)
    need = np.zeros_like(divs)
    for i in range(W):
        need[i, 0] = np.maximum(0, divs[i, 0] - border)
        need[i, 1] = np.minimum(np.asarray(size, dtype=I64), divs[i, 1] + border)
    from_ret = [{} for _ in range(W)]
    to_ret = [{} for _ in range(W)]
    for i in range(W):  # owner
        for j in range(W):  # needer
            if i == j:
                continue
            lo = np.maximum(divs[i, 0], need[j, 0])
            hi = np.minimum(divs[i, 1], need[j, 1])
            if np.all(lo <= hi):
                reg = np.stack([lo, hi])
                from_ret[j][i] = reg
                to_ret[i][j] = reg
    return from_ret, to_ret
