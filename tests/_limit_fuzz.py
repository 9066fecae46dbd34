"""Seeded programs that press on the fuser's table sizes (40 statements, 16 views, 12 spill registers, 4 reductions, 96
instructions): long runs of same-shaped statements with many distinct arrays alive, temporaries that die, reductions of
temporaries (the elided-operand path) placed at every offset from the statement limit, axis sums between plain statements."""
import numpy as onp


def _h(x):
    return onp.array(x.asarray() if hasattr(x, "asarray") else x)


def limit_program(np, seed):
    rng = onp.random.RandomState(5000 + seed)
    fa = (lambda x: x.copy()) if np is onp else np.fromarray
    two_d = bool(seed % 3 == 0)
    shape = (24, 11) if two_d else (300,)
    dt = onp.float32 if seed % 4 == 1 else onp.float64
    base = [fa((rng.randint(-4, 5, size=shape)).astype(dt)) for _ in range(int(rng.randint(2, 7)))]
    live = list(base)
    out = []
    n = int(rng.randint(25, 75))
    for step in range(n):
        r = rng.rand()
        a = live[int(rng.randint(len(live)))]
        b = live[int(rng.randint(len(live)))]
        if r < 0.45:
            t = a + b if rng.rand() < 0.5 else a - b
            live.append(t)
        elif r < 0.6:
            live.append(np.minimum(np.maximum(a * 2.0 - b, -64.0), 64.0))
        elif r < 0.7:
            out.append(onp.asarray(float((a * 2.0 + 1.0).sum())))      # reduction of a temporary nobody else sees
        elif r < 0.76:
            out.append(onp.asarray(float((a + b).sum()) + float(a.sum())))
        elif r < 0.82 and two_d:
            out.append(_h((a - b).sum(axis=int(rng.randint(0, 2)))))
        elif r < 0.9:
            if len(live) > 3:
                del live[int(rng.randint(len(live)))]                      # a handle dies: register temporary or pruned
        else:
            a += 1.0
        if len(live) > 22:                                                 # more arrays than one op list has views
            live = live[-22:]
        a = b = None
    for x in live[-6:]:
        out.append(_h(x))
    return out


def _case(seed):
    def f(np):
        return limit_program(np, seed)

    f.__name__ = "limit_program_%d" % seed
    return f


CASES = [_case(s) for s in range(160)]
