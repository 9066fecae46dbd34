"""Seeded random programs that leave MANY statements pending before anything is read: a pool of variables over three
shapes, out-of-place / in-place / sliced updates, views that stay alive, rebinding and `del`, reductions, reads of single
variables at random points.  The same action list is applied to NumPy (eager, program order) and to ramba_b200 (lazy DAG):
every reordering, pruning and partial materialisation the DAG does must be invisible in the values."""
import numpy as onp


SHAPES = [(230,), (140,), (12, 17)]


def _h(x):
    # a snapshot: the program goes on updating the array in place after the read
    return onp.array(x.asarray() if hasattr(x, "asarray") else x)


def dag_program(np, seed, n_actions=45, trace=None):
    rng = onp.random.RandomState(1000 + seed)
    fa = (lambda x: x.copy()) if np is onp else np.fromarray
    out = []
    pool = {}   # name -> (array, shape index)
    count = [0]

    def fresh(arr, si):
        count[0] += 1
        pool["v%d" % count[0]] = (arr, si)

    for si, shp in enumerate(SHAPES):
        for _ in range(2):
            fresh(fa(rng.randint(-9, 10, size=shp).astype(onp.float64)), si)

    def pick(si=None):
        names = sorted(n for n, (_, s) in pool.items() if si is None or s == si)
        if not names:
            return None, None, None
        n = names[int(rng.randint(len(names)))]
        return n, pool[n][0], pool[n][1]

    def rslice(shp):
        sl = []
        for n in shp:
            lo = int(rng.randint(0, n - 1))
            hi = int(rng.randint(lo + 1, n + 1))
            sl.append(slice(lo, hi))
        return tuple(sl)

    def scalar():
        return float(rng.randint(-3, 4))

    for _ in range(n_actions):
        act = rng.choice(["binop", "binop", "unary", "inplace", "inplace", "setslice", "setslice_arr", "view", "rebind",
                          "del", "read", "sum", "axsum", "sync", "clamp", "dead"])
        n, x, si = pick()
        if x is None:
            break
        if trace is not None:
            trace.append((str(act), n, len(out)))
        if act == "binop":
            _, y, _ = pick(si)
            op = rng.choice(["add", "sub", "mul"])
            r = x + y if op == "add" else (x - y if op == "sub" else x * 0.5 + y)
            fresh(r, si)
        elif act == "unary":
            fresh(abs(x) if rng.rand() < 0.5 else -x, si)
        elif act == "inplace":
            if rng.rand() < 0.5:
                x += scalar()
            else:
                _, y, _ = pick(si)
                x -= y
        elif act == "setslice":
            if x.shape == SHAPES[si]:
                x[rslice(x.shape)] = scalar()
        elif act == "setslice_arr":
            _, y, _ = pick(si)
            if x.shape == SHAPES[si] and y.shape == SHAPES[si]:
                # same-extent windows of two arrays of one shape (possibly the same array, possibly overlapping)
                ext = [int(rng.randint(1, d + 1)) for d in x.shape]
                a0 = [int(rng.randint(0, d - e + 1)) for d, e in zip(x.shape, ext)]
                b0 = [int(rng.randint(0, d - e + 1)) for d, e in zip(x.shape, ext)]
                x[tuple(slice(a, a + e) for a, e in zip(a0, ext))] = y[tuple(slice(b, b + e) for b, e in zip(b0, ext))] * 2.0
        elif act == "view":
            if x.shape == SHAPES[si]:
                v = x[rslice(x.shape)]
                v += 1.0              # writes through to x
                out.append(_h(v * 3.0))
        elif act == "rebind":
            pool[n] = (x * 0.5 + 1.0, si)   # the old handle dies
        elif act == "del":
            if len(pool) > 3:
                del pool[n]
        elif act == "read":
            out.append(_h(x))
        elif act == "sum":
            out.append(onp.asarray(float(x.sum())))
        elif act == "axsum":
            if x.ndim == 2:
                out.append(_h((x + 1.0).sum(axis=int(rng.randint(0, 2)))))
        elif act == "sync":
            if np is not onp:
                np.sync()
        elif act == "clamp":
            pool[n] = (np.minimum(np.maximum(x, -500.0), 500.0), si)
        elif act == "dead":
            _ = x * 7.0 + 1.0     # never observed
            del _
        x = None
    names = sorted(pool)
    for i in rng.permutation(len(names)):   # (a seeded order: every rank of an SPMD run must read in the same one)
        out.append(_h(pool[names[int(i)]][0]))
    return out


def _case(seed):
    def f(np):
        return dag_program(np, seed)

    f.__name__ = "dag_program_%d" % seed
    return f


CASES = [_case(s) for s in range(300)]
