"""Seeded random programs over the fused path: arrays of 1 to 4 dims, random slices (positive, skipping and negative
steps), elementwise mixes, global and axis sums, transposes, `where`, in-place updates of the sliced views.  Compared
with NumPy running the same program (the reference's run_both pattern); used on the oracle executor, across ranks
over gloo, and through the CUDA library."""
import numpy as onp


def _h(x):
    return x.asarray() if hasattr(x, "asarray") else onp.asarray(x)


def random_program(np, seed):
    rng = onp.random.RandomState(seed)
    fa = (lambda x: x) if np is onp else np.fromarray
    out = []
    nd = int(rng.randint(1, 4))
    shape = tuple(int(rng.randint(1, 40)) for _ in range(nd))
    if onp.prod(shape) < 100 and rng.rand() < 0.5:
        shape = shape + (int(rng.randint(20, 60)),)
    A = fa(rng.randint(-50, 50, size=shape))
    B = fa(rng.randint(1, 9, size=shape) * 0.5)
    def rs(n):
        lo = int(rng.randint(0, n)); hi = int(rng.randint(lo, n + 1)); st = int(rng.choice([1, 1, 2, 3, -1, -2]))
        return slice(lo, hi, st) if st > 0 else slice(hi - 1 if hi > 0 else None, lo - 1 if lo > 0 else None, st)
    for _ in range(4):
        sl = tuple(rs(n) for n in A.shape)
        va, vb = A[sl], B[sl]
        out.append(_h(va * 2 + vb))
        if va.size > 0:
            out.append(onp.asarray((va * vb).sum()))
            if va.ndim > 1:
                ax = int(rng.randint(0, va.ndim))
                out.append(_h((va + 1).sum(axis=ax)))
                out.append(_h(va.T * 3))
            m = va > 0
            w = np.where(m, va, vb)  # materialised in the dtype of its first value operand, like the reference
            out.append(_h(w.astype(va.dtype) if np is onp else w))
        B[sl] = vb * 2 + 1
        A[sl] += 3
    out += [_h(A), _h(B)]
    return out


def _case(seed):
    def f(np):
        return random_program(np, seed)

    f.__name__ = "random_program_%d" % seed
    return f


CASES = [_case(s) for s in range(40)]
