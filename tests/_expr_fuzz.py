"""Seeded programs over ONE 2-D shape that mix statement forms inside long pending stretches: integer, float and bool arrays,
arithmetic / comparisons / logical ops / where / minimum / maximum / abs / astype, floor division and modulo on integers,
in-place updates, strided and reversed slice views as operands, sliced assignments, transposed and broadcast (row / column)
operands, global and axis reductions.  Integer-valued data: every result is exact in NumPy and here.  Aimed at the lowering
(forwarding, dead stores, register allocation, SINCOS-free) and at the fuser's alias rules under the DAG's statement orders."""
import numpy as onp

R, C = 13, 22


def _h(x):
    return onp.array(x.asarray() if hasattr(x, "asarray") else x)


def expr_program(np, seed, n_actions=40):
    rng = onp.random.RandomState(9000 + seed)
    fa = (lambda x: x.copy()) if np is onp else np.fromarray
    F = [fa(rng.randint(-6, 7, size=(R, C)).astype(onp.float64)) for _ in range(3)]   # float pool
    I = [fa(rng.randint(-6, 7, size=(R, C)).astype(onp.int64)) for _ in range(2)]     # int pool
    row = fa(rng.randint(-3, 4, size=(C,)).astype(onp.float64))
    col = fa(rng.randint(-3, 4, size=(R, 1)).astype(onp.float64))
    sq = fa(rng.randint(-3, 4, size=(C, R)).astype(onp.float64))                     # read through .T
    out = []

    def f():
        return F[int(rng.randint(len(F)))]

    def i():
        return I[int(rng.randint(len(I)))]

    def put(pool, v):
        if len(pool) >= 6:
            del pool[int(rng.randint(len(pool)))]
        pool.append(v)

    def sl():
        a = int(rng.randint(0, R - 4)); b = int(rng.randint(0, C - 6))
        return (slice(a, a + 4), slice(b, b + 6))

    held = []   # temporaries built by one statement, consumed by a LATER one and dropped before anything flushes them
    for _ in range(n_actions):
        k = int(rng.randint(0, 27))
        if k == 0:
            put(F, f() + f() * 2.0)
        elif k == 1:
            put(F, np.where(f() > f(), f(), f() - 1.0))
        elif k == 2:
            put(F, np.minimum(f(), 3.0) - np.maximum(f(), -2.0))
        elif k == 3:
            put(F, abs(f()) + row)                      # row-broadcast operand
        elif k == 4:
            put(F, f() * col)                           # column-broadcast operand
        elif k == 5:
            put(F, f() + sq.T)                          # transposed operand
        elif k == 6:
            put(I, i() * 3 - i())
        elif k == 7:
            put(I, i() // 4 + i() % 5)
        elif k == 8:
            put(F, i().astype(onp.float64) * 0.5 + f())
        elif k == 9:
            put(I, (f() * 2.0).astype(onp.int64) + i())
        elif k == 10:
            m = np.logical_and(f() > 0.0, i() < 2)
            put(F, np.where(m, f(), -f()))
        elif k == 11:
            x = f(); x += f()
        elif k == 12:
            x = f(); x *= 2.0; x -= 1.0
        elif k == 13:
            x = i(); x += i()
        elif k == 14:
            x = f(); x[sl()] = 0.0
        elif k == 15:
            x, y = f(), f(); s1, s2 = sl(), sl(); x[s1] = y[s2] + 1.0          # maybe the same array, maybe overlapping
        elif k == 16:
            x = f(); put(F, x[::-1, :] - x[:, ::-1])                            # reversed views
        elif k == 17:
            x = f(); v = x[1:R:2, 0:C:3]; put(F, f() * 1.0); out.append(_h(v * 2.0))   # strided view read
        elif k == 18:
            out.append(onp.asarray(float(f().sum())))
        elif k == 19:
            out.append(_h((f() + 1.0).sum(axis=int(rng.randint(0, 2)))))
        elif k == 20:
            out.append(onp.asarray(float((f() * 2.0 - f()).max())))
        elif k in (21, 22):
            if len(held) < 2:
                held.append(f() * 2.0 - f() if k == 21 else abs(f()) + 1.0)      # built now ...
        elif k == 23:
            if held:
                x = f(); x[:, :] = held.pop(0)                                    # ... stored whole, statements later
        elif k == 24:
            if held:
                x = f(); w = sl(); x[w] = held.pop()[w]                           # ... or a window of it
        elif k == 25:
            if held:
                put(F, held.pop(0) + f())                                         # ... or consumed as an operand
        elif k == 26:
            if held:
                held.pop()                                                        # ... or dropped unused
        else:
            out.append(_h(F[int(rng.randint(len(F)))] if rng.rand() < 0.5 else I[int(rng.randint(len(I)))]))
    for x in F + I:
        out.append(_h(x))
    return out


def _case(seed):
    def f(np):
        return expr_program(np, seed)

    f.__name__ = "expr_program_%d" % seed
    return f


CASES = [_case(s) for s in range(200)]


def trig_mask_program(np, seed, n_actions=36):
    """sin / cos of shared operands (the SINCOS pairing and the store it moves), sqrt / exp, boolean-mask assignment and
    masked sums, float32 arrays next to float64 ones - around in-place updates and held temporaries.  Compared with a
    tolerance (transcendentals) by the caller."""
    rng = onp.random.RandomState(12000 + seed)
    fa = (lambda x: x.copy()) if np is onp else np.fromarray
    F = [fa(rng.randint(-6, 7, size=(R, C)).astype(onp.float64) * 0.25) for _ in range(3)]
    G = [fa(rng.randint(-6, 7, size=(R, C)).astype(onp.float32)) for _ in range(2)]      # float32 pool (array-with-array ops only)
    out = []
    held = []

    def f():
        return F[int(rng.randint(len(F)))]

    def g():
        return G[int(rng.randint(len(G)))]

    def put(pool, v):
        if len(pool) >= 5:
            del pool[int(rng.randint(len(pool)))]
        pool.append(v)

    for _ in range(n_actions):
        k = int(rng.randint(0, 18))
        if k == 0:
            x = f(); put(F, np.sin(x) * np.sin(x) + np.cos(x) * np.cos(x))
        elif k == 1:
            x = f(); s = np.sin(x); y = f(); y += 1.0; c = np.cos(x); put(F, s - c)       # an in-place update between the pair
        elif k == 2:
            x = f(); s = np.sin(x); y = f(); u = y * 2.0; y[:, :] = np.cos(x); put(F, u + s)   # the pair's second half overwrites y
        elif k == 3:
            put(F, np.sqrt(abs(f())) + np.exp(np.minimum(f(), 2.0)))
        elif k == 4:
            x = f(); m = f() > 0.0; x[m] = 0.5                                             # boolean-mask assignment
        elif k == 5:
            x = f(); m = x > f(); out.append(onp.asarray(float(x[m].sum())))              # masked sum
        elif k == 6:
            put(G, np.minimum(np.maximum(g() + g(), -8.0), 8.0))      # (kept small: float32 results stay exact, so NumPy's
        elif k == 7:                                                   # rounding of every temporary cannot differ)
            put(G, np.minimum(np.maximum(g() * g() - g(), -8.0), 8.0))
        elif k == 8:
            put(F, g().astype(onp.float64) * 0.5 + f())
        elif k == 9:
            put(G, np.minimum(np.maximum((g().astype(onp.float64) * 2.0).astype(onp.float32) - g(), -8.0), 8.0))
        elif k == 10:
            x = g(); x += g()
        elif k in (11, 12):
            if len(held) < 2:
                x = f(); held.append(np.cos(x) if k == 11 else np.sin(x) + 1.0)
        elif k == 13:
            if held:
                x = f(); x[:, :] = held.pop(0)
        elif k == 14:
            if held:
                put(F, held.pop() * 2.0)
        elif k == 15:
            out.append(_h(f()))
        elif k == 16:
            out.append(onp.asarray(float((np.sin(f()) + np.cos(f())).sum())))
        else:
            x = f(); x -= f()
    for x in F + G:
        out.append(_h(x))
    return out


def _tcase(seed):
    def f(np):
        return trig_mask_program(np, seed)

    f.__name__ = "trig_mask_program_%d" % seed
    return f


TRIG_CASES = [_tcase(s) for s in range(120)]
