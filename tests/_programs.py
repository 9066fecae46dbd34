"""Programs used by the parity tests: each takes the array module (`ramba_b200` or `numpy`) and
returns NumPy results — the reference's own test pattern (`run_both`,
ramba/tests/test_distributed_array.py:255-259)."""
import numpy as onp


def _h(x):
    return x.asarray() if hasattr(x, "asarray") else onp.asarray(x)


def chain(np, N=100003):
    # the reference rewrites `/ c` to `* (1.0/c)` (ramba/ramba.py:6121-6126): the NumPy twin spells it out
    A = np.arange(N) * (1.0 / 1000.0) if np is onp else np.arange(N) / 1000.0
    B = np.sin(A)
    C = np.cos(A)
    D = B * B + C ** 2
    return [_h(A), _h(B), _h(C), _h(D)]


def arith_int(np, N=1000):
    a = np.arange(N)
    b = np.arange(N) * 3 - 7
    return [_h(a + b), _h(a * b), _h(a - b), _h(b // 5), _h(b % 5), _h(1000 - b), _h(a ** 2), _h(-b), _h(abs(b))]


def arith_float(np, N=1000):
    a = np.arange(N) * 0.25 - 30.0
    b = np.arange(N) * 1.5 + 1.0
    return [_h(a + b), _h(a * b), _h(a - b), _h(a * (1.0 / 7.0)) if np is onp else _h(a / 7.0), _h(7.0 * (1.0 / b)) if np is onp else _h(7.0 / b), _h(a // 3.0), _h(a % 3.0), _h(b ** 2), _h(b ** 0.5),
            _h(np.sqrt(b)), _h(abs(a)), _h(-a), _h(np.minimum(a, b)), _h(np.maximum(a, 2.0))]


def compare_ops(np, N=500):
    a = np.arange(N) % 7
    b = np.arange(N) % 5
    return [_h(a > b), _h(a < b), _h(a >= b), _h(a <= b), _h(a == b), _h(a != b), _h(np.logical_and(a > 2, b > 2)),
            _h(np.logical_or(a > 2, b > 2)), _h(np.where(a > b, a, b))]


def float32_mixed(np, N=4096):
    x = (np.arange(N) % 64).astype(onp.float32)
    y = x * 2.0 + 1.0  # float32 array, python float scalars: computed in float64, rounded on store
    z = x * x - y
    return [_h(y), _h(z), onp.asarray(_h(y).dtype == onp.float32)]


def inplace(np, N=1000):
    a = np.arange(N) * 1.0
    a += 1
    a *= 2.0
    a -= 3
    return [_h(a)]


def slices(np, N=1000):
    a = np.arange(N) * 1.0
    b = a[2:-2]
    c = b[::3] + 1.0
    d = a[::-1] * 2.0
    e = a[10:500:7]
    a[5:50] = 0.0
    return [_h(b), _h(c), _h(d), _h(e), _h(a)]


def stencil1d(np, N=1000):
    u = np.arange(N) * 0.5
    v = np.zeros(N)
    v[1:-1] = u[:-2] + u[2:] - 2.0 * u[1:-1]
    return [_h(v)]


def stencil2d(np, n=64, m=48):
    i = np.fromfunction(lambda a, b: (a * 3 + b * 5) % 16, (n, m))
    u = i * 1.0
    v = np.zeros((n, m))
    v[1:-1, 1:-1] = u[:-2, 1:-1] + u[2:, 1:-1] + u[1:-1, :-2] + u[1:-1, 2:] - 4.0 * u[1:-1, 1:-1]
    return [_h(u), _h(v)]


def reductions(np, n=120, m=50):
    x = np.fromfunction(lambda a, b: (a * 131 + b * 31) % 4, (n, m))
    xf = x.astype(onp.float32)
    return [onp.asarray(x.sum()), onp.asarray(xf.sum()), onp.asarray((xf * 2.0 + 1.0).sum()),
            _h(x.sum(axis=0)), _h(x.sum(axis=1)), _h(xf.sum(axis=0)), _h((x % 2 + 1).prod(axis=1)), onp.asarray((x > 1).any()),
            onp.asarray((x >= 0).all())]


def reductions_minmax(np, n=120, m=50):
    # the reference itself cannot run these under NumPy 2 (np.NINF in getminmax, ramba/ramba.py:5349-5355;
    # SURVEY.md §8c): pinned by the NumPy twin only
    x = np.fromfunction(lambda a, b: (a * 131 + b * 31) % 17 - 5, (n, m))
    xf = x.astype(onp.float32) * 0.5
    return [onp.asarray(x.min()), onp.asarray(x.max()), onp.asarray(xf.min()), onp.asarray(xf.max()), _h(x.min(axis=0)), _h(xf.max(axis=1))]


def broadcast_axis_sum(np, n=256, m=64):
    M = np.fromfunction(lambda i, j: (i + 3 * j) % 8, (n, m)).astype(onp.float32)
    v = (np.arange(m) % 8).astype(onp.float32)
    return [_h((M + v).sum(axis=0)), _h(M + v)]


def transpose(np, n=40, m=30):
    x = np.fromfunction(lambda a, b: a * 100 + b, (n, m))
    return [_h(x.T + 1), _h((x.T * 2).sum(axis=0)), _h(x.T[3:20:2, 5:])]


ALL = [chain, arith_int, arith_float, compare_ops, float32_mixed, inplace, slices, stencil1d, stencil2d, reductions,
       reductions_minmax, broadcast_axis_sum, transpose]
NOT_IN_REFERENCE = {"reductions_minmax"}
