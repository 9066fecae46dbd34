"""API-level parity in the style of the reference's own suite
(ramba/tests/test_distributed_array.py: `run_both(func)` runs the same function under ramba and
NumPy and compares, exact by default, :240-259).  Every case runs twice here: on the oracle executor
(CPU, always) and through the CUDA library (-m gpu).  Array sizes straddle distribute_min_size = 100
like the reference's (:… 100, 120, 200)."""
import numpy as onp
import pytest

CASES = []
FIRST_GPU_RUN_IS_THE_DRIVERS = {"ref_reshape_copy"}  # -m gpu legs of these cases live in tests/test_zz_dag_fuzz_gpu.py


def case(f):
    CASES.append(f)
    return f


def _h(x):
    if hasattr(x, "asarray"):
        return x.asarray()
    return onp.asarray(x)


# ---- creation / fillers (TestBasic)
@case
def arange_plain(np):
    return [_h(np.arange(120)), _h(np.arange(5, 125)), _h(np.arange(3, 300, 7))]


@case
def fills(np):
    return [_h(np.zeros(150)), _h(np.ones((20, 10))), _h(np.full((7, 30), 2.5)), _h(np.zeros(130, dtype=onp.int32)),
            _h(np.ones(50)), _h(np.full((3,), 7))]


@case
def linspace_like(np):
    return [_h(np.arange(101) * (5.0 / 100) + 1.0)]


@case
def fromfunction_2d(np):
    return [_h(np.fromfunction(lambda i, j: i * 10 + j, (40, 30))), _h(np.fromfunction(lambda i, j: (i + j) % 3, (12, 11), dtype=onp.int64))]


@case
def eye_like(np):
    return [_h(np.fromfunction(lambda i, j: i == j, (30, 30)).astype(onp.float64))]


# ---- arithmetic over {array, scalar} x operand order (TestOps)
@case
def ops_array_array(np):
    a = np.arange(200) * 1.0
    b = np.arange(200) * 2.0 + 1.0
    return [_h(a + b), _h(a - b), _h(a * b), _h(a // b), _h(b % (a + 1.0))]


@case
def ops_array_scalar(np):
    a = np.arange(120) * 1.0
    return [_h(a + 7), _h(7 + a), _h(a - 7), _h(7 - a), _h(a * 7), _h(7 * a), _h(a // 7), _h(a % 7), _h(a * 13.0), _h(13.0 * a),
            _h(a * onp.float64(1.5)), _h(a + onp.int64(3))]


@case
def ops_small_nondistributed(np):
    a = np.arange(50) * 1.0  # < 100 elements: lives on worker 0 only
    b = np.arange(50) + 10
    return [_h(a + b), _h(a * 2), _h(b - a), onp.asarray((a + b).sum())]


@case
def ops_int(np):
    a = np.arange(300)
    return [_h(a + a), _h(a * 3 - 1), _h(a // 4), _h(a % 4), _h(-a), _h(a ** 2), _h(a & 5), _h(a | 8), _h(a ^ 3), _h(a << 2), _h(a >> 1)]


@case
def ops_inplace(np):
    a = np.ones(150)
    a += 4
    a *= 3.0
    a -= 1
    b = np.arange(150)
    b += 2
    b *= b
    return [_h(a), _h(b)]


@case
def ops_mixed_dtypes(np):
    f32 = (np.arange(128) % 16).astype(onp.float32)
    i64 = np.arange(128)
    return [_h(f32 + f32), _h(f32 * 2), _h(f32 * 2.5), _h(f32 + i64), _h(i64 * 1.5), _h((f32 * f32).astype(onp.float64)), _h(i64.astype(onp.int32) + 1)]


@case
def unary_math(np):
    a = np.arange(200) * 0.01 + 0.5
    return [_h(np.sqrt(a)), _h(np.exp(a)), _h(np.log(a)), _h(np.tanh(a)), _h(np.arctan(a)), _h(np.square(a)), _h(abs(a - 1.2)), _h(-a)]


@case
def predicates(np):
    a = np.arange(100) * 1.0 - 50.0
    return [_h(a > 0), _h(a <= -10), _h(a == 0), _h(a != 3), _h(np.logical_and(a > -5, a < 5)), _h(np.logical_not(a > 0)),
            _h(np.isnan(a)), _h(np.isfinite(a))]


# ---- where / clip / minimum / maximum
@case
def where_clip(np):
    a = np.arange(160) * 1.0
    b = 160.0 - a
    return [_h(np.where(a > b, a, b)), _h(np.where(a > 80, a, 0.0)), _h(a.clip(10.0, 100.0)), _h(np.minimum(a, b)), _h(np.maximum(a, 30.0))]


# ---- views: slices, steps, negative steps, setitem (TestBasic test_slice*, test_setitem)
@case
def slices_1d(np):
    a = np.arange(200)
    return [_h(a[10:50]), _h(a[:-20]), _h(a[::2]), _h(a[5::3]), _h(a[::-1]), _h(a[150:20:-3]), _h(a[-30:]), _h(a[10:50][5:20]), _h(a[::2][::3])]


@case
def slices_2d(np):
    a = np.fromfunction(lambda i, j: i * 100 + j, (30, 40))
    return [_h(a[2:10, 5:15]), _h(a[:, 3]), _h(a[7]), _h(a[::2, ::3]), _h(a[::-1, :]), _h(a[1:-1, 1:-1] + a[:-2, 1:-1])]


@case
def setitem_slices(np):
    a = np.zeros(200)
    a[10:20] = 5.0
    a[::50] = 1.0
    b = np.arange(200) * 1.0
    a[100:150] = b[0:50]
    c = np.zeros((20, 30))
    c[2:5, :] = 3.0
    c[:, 4] = 7.0
    c[10:, 10:] = c[:10, :20] + 1.0
    return [_h(a), _h(c)]


@case
def view_aliasing(np):
    # writes through a view are visible through the base and other views (views never copy)
    a = np.arange(150) * 1.0
    v = a[50:100]
    v += 1000.0
    w = a[::2]
    return [_h(a), _h(v), _h(w)]


@case
def shifted_update_hazard(np):
    # read-after-write through shifted views must not fuse wrongly (tests/…:23-45 of the reference)
    a = np.arange(120) * 1.0
    a[1:] = a[:-1] + a[1:]
    b = np.arange(120) * 1.0
    b[:-1] = b[1:] * 2.0
    return [_h(a), _h(b)]


# ---- broadcast (TestBroadcast)
@case
def broadcast_ops(np):
    m = np.fromfunction(lambda i, j: i + j, (20, 30))
    v = np.arange(30) * 1.0
    c = np.arange(20) * 1.0
    return [_h(m + v), _h(m * v), _h(v + m), _h(m - 3.0 * v), _h((m.T + c).T)]


# ---- reductions (TestReduction)
@case
def reductions_full(np):
    a = np.fromfunction(lambda i, j: (i * 7 + j) % 5, (40, 25))
    i = np.arange(300)
    return [onp.asarray(a.sum()), onp.asarray(i.sum()), onp.asarray((i % 7).prod() == 0), onp.asarray(a.min()), onp.asarray(a.max()),
            onp.asarray((a > 10).any()), onp.asarray((a >= 0).all()), onp.asarray(a.mean())]


@case
def reductions_axis(np):
    a = np.fromfunction(lambda i, j, k: (i + 2 * j + 3 * k) % 11, (10, 12, 14))
    return [_h(a.sum(axis=0)), _h(a.sum(axis=1)), _h(a.sum(axis=2)), _h(a.sum(axis=(0, 2))), _h(a.max(axis=1)), _h(a.min(axis=0))]


@case
def reductions_of_views(np):
    a = np.fromfunction(lambda i, j: i * 3 + j, (30, 20))
    return [onp.asarray(a.T.sum()), _h(a.T.sum(axis=0)), _h(a[5:25, 2:18].sum(axis=1)), onp.asarray(a[::2, ::-1].sum())]


@case
def reduction_fusion_equalities(np):
    # sum(axis=0).sum() == sum(axis=1).sum() == sum() on exactly representable data (TestStencil::test_reduction_fusion)
    a = np.fromfunction(lambda i, j: (i + j) % 9, (50, 60))
    s0 = a.sum(axis=0).sum()
    s1 = a.sum(axis=1).sum()
    s = a.sum()
    return [onp.asarray(s0), onp.asarray(s1), onp.asarray(s), onp.asarray(float(s0) == float(s) == float(s1))]


# ---- stencils through slice views (TestStencil)
@case
def stencil_weighted_2d(np):
    a = np.fromfunction(lambda i, j: (i * 5 + j * 3) % 32, (40, 50))
    b = np.zeros((40, 50))
    b[1:-1, 1:-1] = 0.25 * (a[:-2, 1:-1] + a[2:, 1:-1] + a[1:-1, :-2] + a[1:-1, 2:]) - a[1:-1, 1:-1]
    return [_h(b), onp.asarray(b.sum())]


@case
def stencil_3d_laplacian(np):
    u = np.fromfunction(lambda i, j, k: (i + 2 * j + 3 * k) % 64, (16, 18, 20), dtype=onp.float32)
    v = np.zeros((16, 18, 20), dtype=onp.float32)
    v[1:-1, 1:-1, 1:-1] = (u[:-2, 1:-1, 1:-1] + u[2:, 1:-1, 1:-1] + u[1:-1, :-2, 1:-1] + u[1:-1, 2:, 1:-1]
                           + u[1:-1, 1:-1, :-2] + u[1:-1, 1:-1, 2:] - 6.0 * u[1:-1, 1:-1, 1:-1])
    return [_h(v)]


@case
def sstencil_skeleton(np):
    # TestStencil::test1-style: a relative-index stencil function applied to a distributed array
    def star(a):
        return 0.25 * (a[-1, 0] + a[1, 0] + a[0, -1] + a[0, 1]) + a[0, 0]

    def wide(a, b):
        return a[0, -2] + a[0, 2] - 2.0 * b[0, 0]

    n, m = 30, 40
    kw = {} if np is onp else {"local_border": 2}  # the reference's sstencil wants padded shards (ramba/ramba.py:10010-10015)
    x = np.fromfunction(lambda i, j: (i * 7 + j * 3) % 16, (n, m), **kw)
    y = np.fromfunction(lambda i, j: (i + j) % 5, (n, m), **kw)
    if np is onp:
        r1 = onp.zeros((n, m))
        r1[1:-1, 1:-1] = 0.25 * (x[:-2, 1:-1] + x[2:, 1:-1] + x[1:-1, :-2] + x[1:-1, 2:]) + x[1:-1, 1:-1]
        r2 = onp.zeros((n, m))
        r2[:, 2:-2] = x[:, :-4] + x[:, 4:] - 2.0 * y[:, 2:-2]
        return [r1, r2]
    return [_h(np.sstencil(np.stencil(star), x)), _h(np.stencil(wide)(x, y))]


# (module level: the reference's StencilMetadata.compile reads the source and wants an unindented `def`, ramba/ramba.py:468-474)
def _st_star(a):
    return 0.25 * (a[-1, 0] + a[1, 0] + a[0, -1] + a[0, 1]) + a[0, 0]


def _st_wide(a, b):
    return a[0, -2] + a[0, 2] - 2.0 * b[0, 0]


def _st_diag(a):
    return a[-2, -2] + a[0, 0] + a[2, 2]


@case
def sstencil_local_border(np):
    # the reference's sstencil form (ramba/tests/test_distributed_array.py:17-22, README.md:271-299): arrays created with
    # local_border, a relative-index stencil function, the result allocated like the first argument or given as out=
    star, wide, diag = _st_star, _st_wide, _st_diag
    n, m = 130, 140
    xh = ((onp.arange(n)[:, None] * 7 + onp.arange(m)[None, :] * 3) % 16).astype(onp.float64)
    yh = ((onp.arange(n)[:, None] + onp.arange(m)[None, :]) % 5).astype(onp.float64)
    if np is onp:
        r1 = onp.zeros((n, m))
        r1[1:-1, 1:-1] = 0.25 * (xh[:-2, 1:-1] + xh[2:, 1:-1] + xh[1:-1, :-2] + xh[1:-1, 2:]) + xh[1:-1, 1:-1]
        r2 = onp.zeros((n, m))
        r2[:, 2:-2] = xh[:, :-4] + xh[:, 4:] - 2.0 * yh[:, 2:-2]
        r3 = onp.zeros((n, m))
        r3[2:-2, 2:-2] = xh[:-4, :-4] + xh[2:-2, 2:-2] + xh[4:, 4:]
        r4 = onp.zeros((n, m))
        r4[1:-1, 1:-1] = 0.25 * (r1[:-2, 1:-1] + r1[2:, 1:-1] + r1[1:-1, :-2] + r1[1:-1, 2:]) + r1[1:-1, 1:-1]
        return [r1, r2, r3, r4]
    x = np.fromarray(xh, local_border=2)
    y = np.fromarray(yh, local_border=2)
    s1 = np.sstencil(np.stencil(star), x)
    out = np.zeros((n, m), local_border=2)
    np.sstencil(np.stencil(star), s1, out=out)  # the result of one sstencil (padded like its argument) feeds the next
    return [_h(s1), _h(np.sstencil(np.stencil(wide), x, y)), _h(np.sstencil(np.stencil(diag), x)), _h(out)]


# ---- apps (TestApps): pi integration, manual matmul via broadcast + axis sum
@case
def pi_integration(np):
    n = 100000
    x = (np.arange(n) + 0.5) * (1.0 / n)
    return [onp.asarray(int((4.0 / (1.0 + x * x)).sum() * (1.0 / n) * 1e8))]


@case
def manual_matmul(np):
    a = np.fromfunction(lambda i, j: (i + j) % 4, (12, 15))
    b = np.fromfunction(lambda i, j: (i * 2 + j) % 3, (15, 10))
    # c[i, k] = sum_j a[i, j] * b[j, k] through broadcast + transposes + axis sum
    prods = [_h((a * b.T[k]).sum(axis=1)) for k in range(3)]
    return prods


# ---- masks (test_masked*)
@case
def masked_assign(np):
    a = np.arange(150) * 1.0
    a[a > 100.0] = -1.0
    b = np.arange(150) * 1.0
    m = b % 2 == 0
    b[m] += 1000.0
    return [_h(a), _h(b)]


# ---- TestOps (:262-432): the five arithmetic operators over every pairing of distributed / small / 0-d ramba
# arrays, NumPy arrays and Python / NumPy scalars, in both operand orders; TestBroadcast (:202-237)
@case
def ref_ops_matrix(np):
    out = []
    pairs = [
        lambda: (np.ones((100, 100)), np.ones((100, 100))), lambda: (np.ones((5, 5)), np.ones((5, 5))), lambda: (np.ones(()), np.ones(())),
        lambda: (np.ones((100, 100)), onp.ones((100, 100))), lambda: (onp.ones((100, 100)), np.ones((100, 100))),
        lambda: (np.ones((5, 5)), onp.ones((5, 5))), lambda: (onp.ones((5, 5)), np.ones((5, 5))),
        lambda: (np.ones((100, 100)) * 3, 7.0), lambda: (7.0, np.ones((100, 100)) * 3), lambda: (np.ones((5, 5)) * 3, 7), lambda: (7, np.ones((5, 5)) * 3),
        lambda: (np.array(13), 7), lambda: (7, np.array(13)), lambda: (np.array(13), 7.0), lambda: (7.0, np.array(13)),
        lambda: (np.ones((100, 100)) * 3, onp.ones(1)[0] * 7), lambda: (onp.ones(1)[0] * 7, np.ones((100, 100)) * 3),
        lambda: (np.array(13), onp.ones(1)[0]), lambda: (onp.ones(1)[0], np.array(13)),
    ]
    for mk in pairs:
        for op in ("+", "-", "*", "/", "//"):
            a, b = mk()
            out.append(_h(eval("a" + op + "b")))
    return out


@case
def ref_broadcast(np):
    out = []
    for N in (10, 100):
        a = np.arange(N)
        X = np.fromfunction(lambda x, y: x + y, (N, 1))
        out += [_h(a), _h(X), _h(a + X)]
    return out


@case
def ref_init_array(np):  # TestBasic::test7-12 (:646-694): fillers receive the global index as one tuple
    if np is onp:
        return [onp.arange(120) * 100.0, onp.fromfunction(lambda i, j: (i + j) * 5, (120, 100)), onp.fromfunction(lambda i, j: i * j + 7, (120, 100))]
    return [_h(np.init_array(120, lambda index: index[0] * 100)), _h(np.init_array((120, 100), lambda index: (index[0] + index[1]) * 5)),
            _h(np.init_array((120, 100), lambda x: (x[0] * x[1]) + 7))]


# ---- TestStencil / TestApps / view tests of the reference, as written there (:23-73, :761-872, :1277-1306)
@case
def ref_weighted_subarrays(np):
    def weighted(update_all):
        A = np.ones(100, dtype=float)
        B = np.zeros(100, dtype=float)
        for _ in range(10):
            B[2:-2] += 0.2 * A[:-4] - 0.5 * A[1:-3] + 0.4 * A[2:-2] - 0.5 * A[3:-1] + 0.2 * A[4:]
            if update_all:
                A *= 1.1
            else:
                A[2:-2] *= 1.1   # same size as the stencil: needs read-after-write detection, no fusion
        return onp.asarray(int(abs(B).sum() * 1e8))

    return [weighted(True), weighted(False)]


@case
def ref_reduction_fusion(np):
    A = np.ones((50, 5))
    e = 0.2 * A[:-2] + 0.5 * A[1:-1] + 0.3 * A[2:]
    v = (0.2 * A[:-2] + 0.5 * A[1:-1] + 0.3 * A[2:]).sum(axis=0).sum()
    h = (0.2 * A[:-2] + 0.5 * A[1:-1] + 0.3 * A[2:]).sum(axis=1).sum()
    z = (0.2 * A[:-2] + 0.5 * A[1:-1] + 0.3 * A[2:]).sum()
    return [onp.asarray(v), onp.asarray(h), onp.asarray(z), onp.asarray(bool(v == h and h == z)), _h(e)]


@case
def ref_matmul(np):
    A = np.fromfunction(lambda x, y: x + y, (20, 30))
    B = np.fromfunction(lambda x, y: x + y, (30, 40))
    return [_h((np.broadcast_to(A.T, (40, 30, 20)).T * np.broadcast_to(B, (20, 30, 40))).sum(axis=1)),
            _h((np.expand_dims(A, 2) * B).sum(axis=1))]


@case
def ref_setitem(np):
    a = np.ones((10, 20))
    a[2:5] = 5
    a[..., 8:12] += 12
    a[4, 17] = -3
    a[1, ..., 1] += 32
    a[7, 3] = onp.zeros((1, 1, 1))
    b = np.ones(120)
    v = b[10:50]
    b += 7
    return [_h(a), _h(v)]


@case
def ref_slices(np):
    a = np.arange(200)
    a[20:120] += 50
    b = a[40:140]
    b -= 20
    c = a[60:160] - 25
    d = b + a[80:180]
    r1 = b + c + d
    a = np.arange(200)
    a[20:120:3] += 50
    b = a[40:108:2]
    b -= 20
    c = a[60:196:4] - 25
    d = b + a[80:180:3]
    return [_h(r1), _h(b + c + d)]


@case
def ref_skipslice2(np):
    a = np.fromfunction(lambda i, j: i + j, (500, 50), dtype=int)
    b = a[40:340:3, 20::2]
    b[b > 50] -= 20
    c = np.broadcast_to(b.T, (70, 15, 100))
    d = c[15:25:2, 2:7, ::7] - c[20:30:2, 6:11, 1::7] + 4
    e = np.sum(d)
    return [_h(d + e)]


@case
def ref_negative_skipslice(np):
    a = np.fromfunction(lambda i, j: i + j, (500, 50), dtype=int)
    b = a[340:40:-3, :20:-2]
    b[b > 50] -= 20
    c = np.broadcast_to(b.T, (70, 15, 100))
    d = c[15:25:2, 7:2:-1, ::-3] - c[30:20:-2, 6:11, -1::-3] + 4
    e = np.sum(d)
    return [_h(d + e)]


# ---- test_clip1-4 (:1277-1306), TestReduction (:1308-1366)
@case
def ref_clip(np):
    a = np.arange(200)
    b = np.empty(200, dtype=int)
    a.clip(30, 50, out=b)
    c = np.arange(200)
    c.clip(30, 50, out=c)
    return [_h(a.clip(30, 50)), _h(np.clip(a, 30, 50)), _h(b), _h(c)]


@case
def ref_reduction_sum_prod(np):  # the sum / prod half of testFull, testAxis1, testAxis2 (min / max: see ref_reduction_min_max)
    f = np.fromfunction(lambda i, j, k: 0.01 * i + 0.7 * j + 0.3 * k + 1, (8, 6, 4))
    g = np.fromfunction(lambda i, j, k: 10 * i + 7 * j + k + 1, (8, 6, 4))
    out = []
    for op in ("sum", "prod"):
        out += [onp.asarray(getattr(f, op)()), _h(getattr(g, op)(axis=1)), _h(getattr(g, op)(axis=(1, 0)))]
    return out


@case
def ref_reduction_min_max(np):  # the reference cannot run these under NumPy 2 (np.NINF in getminmax)
    f = np.fromfunction(lambda i, j, k: 0.01 * i + 0.7 * j + 0.3 * k + 1, (8, 6, 4))
    g = np.fromfunction(lambda i, j, k: 10 * i + 7 * j + k + 1, (8, 6, 4))
    out = []
    for op in ("min", "max"):
        out += [onp.asarray(getattr(f, op)()), _h(getattr(g, op)(axis=1)), _h(getattr(g, op)(axis=(1, 0)))]
    return out


@case
def ref_transpose_reductions(np):
    out = []
    for sl in (slice(None), slice(50, 170), slice(150, 170)):
        def mk():
            return np.ones((200, 100), dtype=int)[sl]
        out += [_h(np.sum(mk(), axis=0)), _h(np.sum(mk(), axis=1)), _h(np.sum(mk().T, axis=0)), _h(np.sum(mk().T, axis=1)),
                onp.asarray(np.sum(mk())), onp.asarray(np.sum(mk().T))]
    return out


# ---- joining and padding (test_concatenate_1-2 :1075-1091, test_pad1 / pad2 / pad1_slice :1189-1276)
@case
def ref_concatenate(np):
    shape = (20, 4)
    a = np.fromfunction(lambda i, j: i + j, shape, dtype=int)
    b = np.fromfunction(lambda i, j: i + j, shape, dtype=int)
    c = np.fromfunction(lambda i, j: i * 7 - j, (130, 4), dtype=int)
    return [_h(np.concatenate([a, b], axis=0)), _h(np.concatenate([a, b], axis=1)), _h(np.concatenate([a, c, b], axis=0)), _h(np.concatenate([a, b]))]


@case
def ref_reshape_copy(np):
    """reshape_copy (ramba/ramba.py:9241-9277): merges, splits, an unrelated factorization, a sliced source; small sizes
    (the reference moves one element at a time in Python)."""
    def rc(a, shp):
        return a.reshape_copy(shp) if np is not onp else onp.reshape(a, shp).copy()

    a = np.fromfunction(lambda i, j: i * 15 + j, (24, 15), dtype=int)
    v = np.arange(360) * 2
    out = [_h(rc(a, (360,))), _h(rc(v, (24, 15))), _h(rc(a, (15, 24))), _h(rc(a, (4, 6, 15))), _h(rc(a, (20, 18)))]
    out.append(_h(rc(v[20:320], (30, 10)) + 1))
    return out


@case
def nan_reductions(np):  # the reference's nansum / nanmean give NaN here (its masked sum does not keep the NaNs out)
    v = onp.arange(240) * 0.5
    v[::7] = onp.nan
    a = v if np is onp else np.fromarray(v)
    # (also isclose(equal_nan=True): the reference compiles its isclose with fastmath, which drops the isnan tests)
    if np is onp:
        return [onp.asarray(onp.nansum(a)), onp.asarray(onp.nanmean(a)), onp.isclose(a, a + 1e-9, equal_nan=True)]
    return [onp.asarray(a.nansum()), onp.asarray(a.nanmean()), _h(a.isclose(a + 1e-9, equal_nan=True))]


@case
def isclose_and_rollaxis(np):
    v = onp.arange(240) * 0.5
    v[::7] = onp.nan
    a = v if np is onp else np.fromarray(v)
    b = np.fromfunction(lambda i, j, k: i * 100 + j * 10 + k, (6, 5, 4))
    close = onp.isclose(a, a + 1e-9) if np is onp else a.isclose(a + 1e-9)
    return [_h(close), _h(np.rollaxis(b, 2)), _h(np.rollaxis(b, 0, 3)), _h(np.rollaxis(b, 1, 0) * 2)]


@case
def scumulative_family(np):
    a = np.fromfunction(lambda i: (i * 37) % 101, (300,), dtype=int)
    g = np.fromfunction(lambda i, j: (i * 7 + j * 13) % 50, (40, 30), dtype=int)
    if np is onp:
        return [onp.cumsum(a), onp.cumsum(g, axis=0), onp.cumsum(g, axis=1)]
    return [_h(np.scumulative(lambda x, y: x + y, lambda x, y: x + y, a, 0)), _h(np.scumulative(lambda x, y: x + y, lambda x, y: x + y, g, 0)),
            _h(np.cumsum(g, axis=1))]


@case
def scumulative_forms(np):  # associative functions other than +, ramba functions inside, string lambdas
    a = np.fromfunction(lambda i: (i * 37) % 101, (300,), dtype=int)
    g = np.fromfunction(lambda i, j: (i * 7 + j * 13) % 50, (40, 30), dtype=int)
    if np is onp:
        return [onp.maximum.accumulate(a), onp.minimum.accumulate(g, axis=1)]
    return [_h(np.scumulative(lambda x, y: np.maximum(x, y), lambda x, y: np.maximum(x, y), a)),
            _h(np.scumulative("lambda x, y: numpy.minimum(x, y)", None, g, axis=1))]


@case
def split_family(np):
    a = np.fromfunction(lambda i, j: i * 10 + j, (120, 6), dtype=int)
    parts = np.split(a, 4)
    cols = np.split(a, 3, axis=1)
    return [_h(p) for p in parts] + [_h(c * 2) for c in cols] + [_h(parts[1] + parts[2])]


@case
def stack_family(np):  # the reference declares stack but its executor is a stub (ramba/ramba.py:9576-9577)
    a = np.fromfunction(lambda i, j: i + j, (20, 4), dtype=int)
    b = np.fromfunction(lambda i, j: i * j, (20, 4), dtype=int)
    return [_h(np.stack([a, b])), _h(np.stack([a, b, a], axis=2)), _h(np.stack([a, b], axis=-1))]


@case
def ref_pad_1d(np):
    out = []
    tests = [(2, {}), ((0, 1), {}), ((2, 0), {}), ((3, 4), {}), ((0, 3), {"constant_values": ((0, 7),)}), ((5, 0), {"constant_values": ((5, 0),)}),
             ((1, 2), {"constant_values": ((3, 4),)})]
    for mode in ("constant", "edge", "wrap"):
        for width, kw in tests:
            if kw and mode != "constant":
                continue
            out.append(_h(np.pad(np.arange(200), width, mode=mode, **kw)))
            out.append(_h(np.pad(np.arange(300)[25:225], width, mode=mode, **kw)))
    return out


@case
def ref_pad_2d(np):
    out = []
    tests = [2, (2, 3), ((0, 1), (0, 1)), ((2, 0), (3, 0)), ((2, 0), (0, 3)), ((0, 2), (3, 0)), ((2, 2), (0, 3)), ((0, 2), (3, 3)), ((4, 2), (3, 3))]
    for shape in [(20, 30), (400, 1), (1, 300)]:
        for mode in ("constant", "edge", "wrap"):
            for width in tests:
                if mode == "wrap" and 1 in shape and width in (2, (2, 3), ((2, 0), (3, 0)), ((2, 0), (0, 3)), ((0, 2), (3, 0)), ((2, 2), (0, 3)),
                                                                ((0, 2), (3, 3)), ((4, 2), (3, 3))):
                    continue  # wrapping wider than the axis
                a = np.fromfunction(lambda i, j: i + j, shape, dtype=int)
                out.append(_h(np.pad(a, width, mode=mode)))
    return out


# ---- the reference's TestBasic cases, as written there (masks :975-990, where :992-1021, linspace :1093-1133,
# identity/eye :773-792, transposes :1047-1073, transposed reductions :1332-1366)
@case
def ref_masked(np):
    a = np.arange(200)
    a[a % 5 == 1] -= 50
    g = np.fromfunction(lambda i, j: i + j, (50, 50), dtype=int)
    return [_h(a), onp.asarray(g[g < 20].sum())]


@case
def ref_where(np):
    a = np.arange(200)
    b = np.ones(200)
    g = np.fromfunction(lambda i, j: i + j, (50, 50), dtype=int)
    col = np.ones((50, 1))
    e = np.fromfunction(lambda i, j: 500 + i + j, (50, 50), dtype=int)
    bc = np.fromfunction(lambda i, j: i + 200, (50, 1), dtype=int)
    out = [np.where(a > 133, a, b), np.where(g > 33, g, col), np.where(bc > 233, g, e)]
    if np is onp:
        # the reference materialises `where` in the dtype of its first value operand (where_executor,
        # ramba/ramba.py:9785: empty(..., dtype=a.dtype)) even though it announces result_type(a, b)
        out = [o.astype(onp.int64) for o in out]
    return [_h(o) for o in out]


@case
def ref_linspace(np):
    l3 = np.linspace(1, 5, num=10, endpoint=False, retstep=True)
    l4 = np.linspace(1, 5, num=200, endpoint=False, retstep=True)
    return [_h(np.linspace(1, 5, num=10)), _h(np.linspace(1, 5, num=200)), _h(l3[0]), _h(l4[0]), onp.asarray(l3[1]), onp.asarray(l4[1]),
            _h(np.linspace(10, 30, dtype=int))]


@case
def ref_identity_eye(np):
    return [_h(np.identity(100)), _h(np.eye(100, 50)), _h(np.eye(50, 100, k=3))]


@case
def ref_transposes(np):
    a = np.fromfunction(lambda i, j: i * 100 + j, (30, 40))
    b = np.fromfunction(lambda i, j, k: i * 10000 + j * 100 + k, (12, 15, 10))
    return [_h(np.transpose(a)), _h(np.transpose(b)), _h(np.transpose(b, (1, 0, 2))), _h(b.transpose(2, 0, 1)),
            _h(a.T.sum(axis=0)), _h(a.T[5:30, 3:20].sum(axis=1)), onp.asarray(b.transpose(1, 2, 0)[2:9, :, 3:].sum())]


# ---- NumPy's own functions called on ramba arrays (__array_ufunc__ / __array_function__, ramba/ramba.py:6825-6894;
# the mechanism tests/test_xarray.py:36-47 relies on)
@case
def numpy_protocol(np):
    x = np.fromfunction(lambda i, j: i + j, (10, 20))
    chain = onp.sin((x + 10.0) * 7.1).transpose().sum()
    return [_h(onp.sin(x)), _h(onp.add(x, 10.0)), _h(onp.multiply(7.1, x)), onp.asarray(onp.sum(x)), _h(onp.sum(x, axis=0)),
            _h(onp.where(x > 5, x, 0 * x)), _h(onp.clip(x, 3, 9)), _h(onp.sqrt(x)), _h(onp.maximum(x, 7.0)),
            _h(onp.isnan(x)), _h(onp.square(x)), _h(onp.logical_and(x > 3, x < 9)), onp.asarray(chain)]


@case
def numpy_protocol_more(np):  # NumPy functions the reference does not register for its arrays
    x = np.fromfunction(lambda i, j: i + j, (10, 20))
    return [_h(onp.transpose(x)), _h(onp.concatenate([x, x])), _h(onp.expand_dims(x, 0)), onp.asarray(onp.mean(x)), onp.asarray(onp.min(x)),
            _h(onp.moveaxis(x, 0, 1)), _h(onp.squeeze(onp.expand_dims(x, 1))), _h(onp.abs(x - 10)), _h(onp.power(x, 2))]


# ---- deletion while ops are pending (TestDel :1398-1432)
@case
def delete_pending(np):
    a = np.ones(100)
    b = a + 3
    del a
    s = 0
    c = np.ones(100)
    d = c * 3
    del c
    for i in range(20):
        a = np.ones(200)
        v = a[37:137]
        c = v * 3
        s += c[42]
    d += s
    return [_h(b), _h(d)]


# ---- randomised elementwise / transpose / slice mixes (the commented-out TestGeneric :1435-1560, seeded)
@case
def random_generic(np):
    rng = onp.random.RandomState(20260922)
    fa = (lambda x: x) if np is onp else np.fromarray
    out = []
    for _ in range(6):
        x, y = int(rng.randint(1, 60)), int(rng.randint(1, 60))
        al, bl = rng.randint(200, size=(x, y)) * 0.5, rng.randint(200, size=(x, y)) * 0.5
        dl = rng.randint(200, size=(y, x)) * 0.5
        a, b, d = fa(al), fa(bl), fa(dl)
        out += [_h(2 * (a + b)), _h(d.T * b - a), _h((d.T + b).sum(axis=0)), onp.asarray((a * b).sum())]

    def rand_slices(k):
        lo = int(rng.randint(0, k - 1))
        hi = int(rng.randint(lo, k))
        n = hi - lo
        c = int(rng.randint(0, k - n + 1))
        return slice(lo, hi), slice(c, c + n)

    Al, Bl = rng.randint(20, size=(120, 90)), rng.randint(20, size=(150, 130))
    A, B = fa(Al), fa(Bl)
    for _ in range(6):
        s1, s2 = rand_slices(120)
        t1, t2 = rand_slices(90)
        out += [_h(A[s1, t1] * 3 + B[s2, t2]), _h(A[s1, t1] - A[s2, t2])]
        A[s1, t1] = B[s2, t2] + 1
    out.append(_h(A))
    return out


# ---- 0-d arrays and index terms (test_0d_* :705-759), newaxis
@case
def zero_d(np):
    a0 = np.array(7)
    a0[()] = 3
    a = np.arange(200)
    b = np.array(7, dtype=int)
    a[b] = 0
    o = np.ones((20, 20))
    o[:, b] = 0
    return [onp.asarray(np.array(7)[()]), onp.asarray(a0[()]), onp.asarray(a[b]), _h(a), _h(o[:, b]), _h(o), onp.asarray(float(np.array(7)))]


@case
def newaxis_views(np):
    c = np.ones((6, 7, 8)) * 3
    return [_h(np.arange(120)[:, None] * 1.0), _h(np.arange(120)[None, :] + np.arange(30)[:, None]), _h(c[2, None, ..., None, 1:5]),
            _h(c[None].sum(axis=0))]


# ---- unit-dim views (expand_dims / squeeze, ramba/ramba.py:9438-9476)
@case
def unit_dim_views(np):
    a = np.fromfunction(lambda i, j: i * 10 + j, (30, 8))
    v = np.arange(120) * 1.0
    e0, e1, e2 = np.expand_dims(a, 0), np.expand_dims(a, 1), np.expand_dims(a, (0, 3))
    col = np.expand_dims(v, 1)          # (120, 1)
    row = np.expand_dims(v[:8], 0)      # (1, 8)
    out = [_h(e0), _h(e1), _h(e2), _h(e1 * 2.0 + 1.0), _h(np.squeeze(e2)), _h(np.squeeze(e1, axis=1) - a),
           _h(col[:30] + row), _h((e0 + 1.0).sum(axis=0)), _h(np.reshape(v, (120, 1, 1))), _h(np.reshape(e2, (30, 8)))]
    return out


# ---- index-driven builders (test_triu1-3 :1023-1045, test_mgrid_1-4 :1135-1163, meshgrid, select)
@case
def triu_family(np):
    a = np.fromfunction(lambda i, j: i + j, (50, 50), dtype=int)
    r = np.fromfunction(lambda i, j: i * 100 + j, (30, 120), dtype=onp.float64)
    return [_h(np.triu(a)), _h(np.triu(a, k=-2)), _h(np.triu(a, k=2)), _h(np.triu(r, k=5))]


@case
def tril_family(np):  # NumPy's tril; the reference has only triu
    a = np.fromfunction(lambda i, j: i + j, (50, 50), dtype=int)
    r = np.fromfunction(lambda i, j: i * 100 + j, (30, 120), dtype=onp.float64)
    return [_h(np.tril(a)), _h(np.tril(r, k=-3))]


@case
def mgrid_meshgrid(np):
    out = [_h(np.mgrid[0:20, 0:20]), _h(np.mgrid[0:5, 0:5]), _h(np.mgrid[0:7, 0:30, 0:3])]
    m, n = np.mgrid[0:20, 0:20]
    out += [_h(m), _h(n)]
    x, y = np.arange(30) * 0.5, np.arange(12) * 2.0
    if np is onp:
        out.append(onp.stack(onp.meshgrid(x, y, indexing="ij")))
    else:
        out.append(_h(np.meshgrid(x, y, indexing="ij")))
    return out


@case
def mgrid_offsets(np):  # slice starts: NumPy semantics (the reference's mgrid counts from 0 whatever the start)
    return [_h(np.mgrid[2:9, 0:30, 1:4]), _h(np.mgrid[5:25])]


@case
def select_family(np):
    a = np.arange(200) - 50
    conds = [a < 0, a < 50, a < 100]
    choices = [a * 0, a * 2, a * 3]
    if np is onp:
        # the reference's select, as written (ramba/ramba.py:9079-9092): float64, assignment order 0, -1, -2, ...
        temp = onp.full(a.shape, -7.0)
        for i in range(len(choices)):
            temp[conds[-i]] = choices[-i][conds[-i]]
        return [temp]
    return [_h(np.select(conds, choices, default=-7))]


# ---- skeletons over user functions (test_smap1-3, test_smap_index1-3, :919-972) and cumsum (:1368-1386)
@case
def smap_family(np):
    a = np.arange(100)
    a2 = a * a
    g = np.fromfunction(lambda i, j: i + j, (100, 100))
    if np is onp:
        return [3 * a - 7, 3 * a - 7, 3 * g - 7, 3 * a2 - 7 * a, 3 * a2 - 7 * a, (a * 0.5).astype(onp.int64), a * 0.5,
                (onp.sin(a * 0.25) * 10).astype(onp.int64)]
    return [_h(np.smap("lambda x: 3*x-7", a)), _h(np.smap(lambda x: 3 * x - 7, a)), _h(np.smap("lambda x: 3*x-7", g)),
            _h(np.smap("lambda x,y: 3*x-7*y", a2, a)), _h(np.smap(lambda x, y: 3 * x - 7 * y, a2, a)),
            _h(np.smap(lambda x: x * 0.5, a)),  # result takes the dtype of the first array
            _h(np.smap(lambda x: x * 0.5, a, dtype=onp.float64)),
            _h(np.smap("lambda x: numpy.sin(x*0.25)*10", a, imports=["numpy"]))]  # string lambdas see the package as `numpy`


@case
def smap_index_family(np):
    a = np.arange(100) - 25
    a2 = np.arange(100) * a
    o = np.ones((100, 100))
    if np is onp:
        i = onp.arange(100)
        return [7 * i + a, 7 * i + a, onp.fromfunction(lambda i, j: 7 * i - j + 1, (100, 100)), 7 * i + a - 4 * a2]
    return [_h(np.smap_index("lambda i,x: 7*i+x", a)), _h(np.smap_index(lambda i, x: 7 * i + x, a)),
            _h(np.smap_index("lambda i,x: 7*i[0]-i[1]+x", o)), _h(np.smap_index(lambda i, x, y: 7 * i + x - 4 * y, a, a2))]


@case
def sreduce_family(np):
    a = np.arange(300) - 100
    if np is onp:
        return [onp.asarray((a * a).sum()), onp.asarray(onp.max(2 * a + 1)), onp.asarray(5 + (a * onp.arange(300)).sum())]
    return [onp.asarray(np.sreduce(lambda x: x * x, lambda p, q: p + q, 0, a)),
            onp.asarray(np.sreduce(lambda x: 2 * x + 1, lambda p, q: max(p, q), -10**9, a)),
            onp.asarray(np.sreduce_index(lambda i, x: i * x, lambda p, q: p + q, 5, a))]


@case
def sreduce_forms(np):  # string lambdas and SreduceReducer pairs (the reference takes functions only)
    a = np.arange(300) - 100
    if np is onp:
        return [onp.asarray(onp.max(2 * a + 1)), onp.asarray(5 + (a * onp.arange(300)).sum()), onp.asarray(onp.min(a * 0.5))]
    return [onp.asarray(np.sreduce("lambda x: 2*x+1", lambda p, q: max(p, q), -10**9, a)),
            onp.asarray(np.sreduce_index(lambda i, x: i * x, "lambda p,q: p+q", 5, a)),
            onp.asarray(np.sreduce(lambda x: x * 0.5, np.SreduceReducer(min, min), 1e300, a))]


@case
def cumsum_family(np):
    out = [_h(np.cumsum(np.arange(200))), _h(np.cumsum(np.arange(150) * 0.5)), _h(np.cumsum(np.arange(7)))]
    for shp in [(4, 50), (20, 20), (50, 4)]:  # the reference's own shapes (tests/…:1375-1386)
        a = onp.arange(shp[0] * shp[1]).reshape(shp)
        b = a if np is onp else np.fromarray(a)
        for axis in range(2):
            out.append(_h(np.cumsum(b, axis=axis)))
    return out


# ---- chains longer than the op-list tables of the C-ABI: the fuser cuts them, it never raises (the reference compiles
# whatever its fuser accumulated)
@case
def long_chain_many_arrays(np):  # 20 live arrays written by one flush (> 16 views)
    xs = [np.arange(200) * float(i) for i in range(20)]
    return [_h(x) for x in xs]


@case
def long_chain_many_operands(np):  # one expression over 21 arrays (> 16 views)
    xs = [np.arange(150) * (i + 1) for i in range(21)]
    if np is not onp:
        np.sync()
    t = xs[0]
    for x in xs[1:]:
        t = t + x
    return [_h(t)]


@case
def long_chain_live_temporaries(np):  # 14 temporaries alive at once (> 12 spill registers)
    a = np.arange(130) * 0.5
    ts = [a * float(i + 2) for i in range(14)]
    acc = ts[0]
    for i in range(1, 14):
        acc = acc * 0.5 + ts[i]
    for i in range(14):
        acc = acc - ts[i] * 0.25
    del ts  # dead handles: the 14 values are register temporaries, all alive until the second loop used them
    return [_h(acc)]


@case
def long_chain_many_instructions(np):  # > 96 instructions in one flush
    a = np.arange(180) * 0.25
    b = a
    for i in range(70):
        b = (b + float(i)) * 0.5 - a
    return [_h(b), onp.asarray((b * 2.0).sum())]


@case
def long_chain_many_reductions(np):  # > 4 global reductions pending in one fused op
    a = np.arange(210) * 0.5
    if np is onp:
        return [onp.asarray((a * float(i)).sum()) for i in range(7)]
    parts = [(a * float(i)).sum(asarray=True) for i in range(7)]
    return [onp.asarray(_h(p)[0]) for p in parts]


@case
def where_float_condition(np):  # the condition is tested in its own class before the branches' class applies
    c = (np.arange(120) - 60) * 0.25
    ia, ib = np.arange(120), np.arange(120) * -1
    f32 = (np.arange(120) * 1.0).astype(onp.float32)
    tiny = (np.arange(120) % 3) * 1e-300
    return [_h(np.where(c, ia, ib)), _h(np.where(tiny, f32, f32 * onp.float32(2.0)))]


# ---- operands every rank needs in full / partial rows that cross ranks (BASELINE config 5's shape, small)
@case
def broadcast_vector_axis_sum(np):
    M = np.fromfunction(lambda i, j: (i + 3 * j) % 8, (1024, 128)).astype(onp.float32)  # split by rows at 2..8 ranks
    v = (np.arange(128) % 8).astype(onp.float32)                                       # split into chunks: every rank needs all
    w = np.arange(128) * 0.5
    return [_h((M + v).sum(axis=0)), _h(M * v - w), onp.asarray((M * 2.0 + v).sum()), _h((M + v).sum(axis=1)),
            _h((M * v).prod(axis=0) * 0.0 + (M + 1.0).sum(axis=0))]


def _compare(got, exp, name):
    assert len(got) == len(exp), name
    for i, (g, e) in enumerate(zip(got, exp)):
        g, e = onp.asarray(g), onp.asarray(e)
        assert g.shape == e.shape, "%s[%d]: shape %s vs %s" % (name, i, g.shape, e.shape)
        if e.dtype.kind == "f":
            assert g.dtype.kind == "f"
            assert onp.allclose(g, e, rtol=1e-13 if e.dtype == onp.float64 else 1e-6, atol=1e-15), "%s[%d]" % (name, i)
        else:
            assert onp.array_equal(g, e), "%s[%d]: %r vs %r" % (name, i, g.reshape(-1)[:8], e.reshape(-1)[:8])


@pytest.mark.parametrize("f", CASES, ids=lambda f: f.__name__)
def test_run_both_oracle(oracle_engine, f):
    import ramba_b200 as rb

    _compare(f(rb), f(onp), f.__name__)


@pytest.mark.gpu
@pytest.mark.parametrize("f", CASES, ids=lambda f: f.__name__)
def test_run_both_cuda(gpu_engine, f):
    import ramba_b200 as rb
    from ramba_b200 import _cabi
    from ramba_b200.runtime import RT

    if f.__name__ in FIRST_GPU_RUN_IS_THE_DRIVERS:
        pytest.skip("runs in tests/test_zz_dag_fuzz_gpu.py (written after the round's GPU budget was spent: collected last)")
    before = _cabi.launch_count()
    got = f(rb)
    assert RT.is_cuda and _cabi.launch_count() > before
    _compare(got, f(onp), f.__name__)
