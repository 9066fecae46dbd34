#!/usr/bin/env python
"""Generate the golden fixtures by running the REAL reference (Python-for-HPC/ramba mounted at
/root/reference) in this container.  The reference cannot travel to the GPU box, so its outputs
are committed next to this script:

    partition_golden.json   outputs of ramba.common / ramba.shardview_array pure functions
                            (work division, slicing, broadcast, transpose, reduce distributions,
                            intersections, range splits) for several worker counts
    programs_golden.npz     outputs of tests/_programs.py run under `import ramba`
                            (RAMBA_NON_DIST=1, Numba CPU path) — pins the oracle and the CUDA path
    api_golden.npz          outputs of the API-level cases of tests/test_api_parity.py run under `import ramba`;
                            cases the reference cannot run here (functions it does not have, its NumPy-2
                            incompatibilities) are recorded with the reason in __status__

Usage (from the repo root; needs /root/reference, numba; Ray is replaced by a 10-line stub because
NON_DIST mode never calls it — SURVEY.md Appendix C):

    python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"

RAY_STUB = '''
def get(*a, **k): raise RuntimeError("ray stub")
def put(*a, **k): raise RuntimeError("ray stub")
def wait(*a, **k): raise RuntimeError("ray stub")
def is_initialized(): return False
def remote(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    def deco(f): return f
    return deco
'''

CHILD = r'''
import json, os, sys
import numpy as onp
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ramba
import ramba.common as rc
import ramba.shardview_array as sv
import _programs

mode = sys.argv[1]
out_path = sys.argv[2]

def svl(s):
    return [sv._size(s).tolist(), sv._index_start(s).tolist(), sv._axis_map(s).tolist(), sv._steps(s).tolist(),
            sv._base_offset(s).tolist()]

if mode == "partition":
    G = {"schedule": [], "slice": [], "broadcast": [], "remap": [], "reduce": [], "intersect": [], "splits": []}
    shapes = [(10,), (100,), (1000,), (1000003,), (6, 8), (120, 50), (100, 100), (32768, 32768), (7, 1000), (1000, 7),
              (1024, 1024, 1024), (30, 40, 50), (1048576, 4096), (4096,), (5, 6, 7, 8), (1000000000,)]
    for W in (1, 2, 3, 4, 5, 6, 8):
        rc.num_workers = W
        rc.num_nodes = 1
        sv.num_workers = W
        for shape in shapes:
            if onp.prod(shape) < 100:
                pass
            try:
                div = rc.compute_regular_schedule_internal(W, shape, ())
            except AssertionError:
                continue
            G["schedule"].append({"W": W, "shape": list(shape), "divisions": div.tolist()})
        # distributions + view algebra on a few shapes
        for shape in [(10,), (100,), (6, 8), (120, 50), (30, 40, 50)]:
            try:
                div = rc.compute_regular_schedule_internal(W, shape, ())
            except AssertionError:
                continue
            D = sv.divisions_to_distribution(div)
            k = len(shape)
            slices_list = []
            if k == 1:
                n = shape[0]
                slices_list = [(slice(2, n - 2, 1),), (slice(1, n, 3),), (slice(0, n, 1),), (slice(n - 1, -1, -1),), (slice(n - 2, 0, -2),),
                               (slice(3, 4, 1),), (slice(0, n // 2, 2),)]
            elif k == 2:
                n, m = shape
                slices_list = [(slice(1, n - 1, 1), slice(1, m - 1, 1)), (slice(0, n - 2, 1), slice(2, m, 1)), (slice(0, n, 2), slice(1, m, 3)),
                               (slice(n - 1, -1, -1), slice(0, m, 1)), (slice(2, 3, 1), slice(0, m, 1))]
            else:
                a, b, c = shape
                slices_list = [(slice(1, a - 1, 1), slice(1, b - 1, 1), slice(1, c - 1, 1)), (slice(0, a - 2, 1), slice(1, b - 1, 1), slice(2, c, 1)),
                               (slice(0, a, 2), slice(0, b, 1), slice(c - 1, -1, -3))]
            for sl in slices_list:
                S = sv.slice_distribution(sl, D)
                G["slice"].append({"W": W, "shape": list(shape), "slices": [[s.start, s.stop, s.step] for s in sl], "dist": [svl(x) for x in S]})
                # slice of a slice (positive steps only for the second)
                sshape = tuple(max(0, -(-(s.stop - s.start) // s.step)) if s.step > 0 else max(0, -(-(s.start - s.stop) // (-s.step))) for s in sl)
                if all(x > 3 for x in sshape):
                    sl2 = tuple(slice(1, x - 1, 1) for x in sshape)
                    S2 = sv.slice_distribution(sl2, S)
                    G["slice"].append({"W": W, "shape": list(shape), "slices": [[s.start, s.stop, s.step] for s in sl],
                                       "slices2": [[s.start, s.stop, s.step] for s in sl2], "dist": [svl(x) for x in S2]})
            if k == 1:
                big = (7, shape[0])
                B = sv.broadcast(D, [True, False], big)
                G["broadcast"].append({"W": W, "shape": list(shape), "bdims": [True, False], "size": list(big), "dist": [svl(x) for x in B]})
                D2 = sv.divisions_to_distribution(rc.compute_regular_schedule_internal(W, big, ())) if onp.prod(big) >= 1 and W <= onp.prod(big) else None
                if D2 is not None:
                    for i in range(W):
                        for j in range(W):
                            G["intersect"].append({"W": W, "shape": list(shape), "i": i, "j": j, "part": svl(sv.intersect(B[i], D2[j])),
                                                   "compat": bool(sv.is_compat(sv.clean_range(D2[j]), B[j]))})
            if k >= 2:
                perm = list(range(k))[::-1]
                ns, R = sv.remap_axis(shape, D, perm)
                G["remap"].append({"W": W, "shape": list(shape), "perm": perm, "new_shape": list(ns), "dist": [svl(x) for x in R]})
                for axes in ([0], [k - 1], list(range(k))):
                    rsz, rdist, bdist = sv.reduce_axes(shape, D, axes)
                    G["reduce"].append({"W": W, "shape": list(shape), "axes": axes, "rsz": [int(x) for x in rsz],
                                        "rdist": [svl(x) for x in rdist], "bdist": [svl(x) for x in bdist]})
            import numba
            lst = numba.typed.List()
            for x in D:
                lst.append(sv.clean_range(x))
            sp = sv.get_range_splits_list(lst)
            G["splits"].append({"W": W, "shape": list(shape), "splits": sorted([svl(x)[:2] for x in sp])})
    with open(out_path, "w") as f:
        json.dump(G, f)
elif mode == "api":
    import test_api_parity
    res = {}
    status = {}
    only = [n for n in os.environ.get("RB_GOLDEN_CASES", "").split(",") if n]
    for f in test_api_parity.CASES:
        if only and f.__name__ not in only:
            continue
        try:
            outs = f(ramba)
            ramba.sync()
            outs = [onp.asarray(o) for o in outs]
            for i, o in enumerate(outs):
                res["%s__%d" % (f.__name__, i)] = o
            status[f.__name__] = "ok"
        except BaseException as ex:  # the reference lacks the function, or trips over NumPy 2 (SURVEY §8c)
            status[f.__name__] = "reference failed: %s: %s" % (type(ex).__name__, str(ex)[:200])
            for k in [k for k in res if k.startswith(f.__name__ + "__")]:
                del res[k]
    res["__status__"] = onp.array(json.dumps(status))
    onp.savez_compressed(out_path, **res)
    print(json.dumps(status, indent=1))
else:
    res = {}
    status = {}
    for prog in _programs.ALL:
        try:
            outs = prog(ramba)
            ramba.sync()
            for i, o in enumerate(outs):
                res["%s__%d" % (prog.__name__, i)] = onp.asarray(o)
            status[prog.__name__] = "ok"
        except Exception as ex:  # reference limitation (e.g. NumPy-2 incompatibilities, SURVEY §8c)
            status[prog.__name__] = "reference failed: %s: %s" % (type(ex).__name__, str(ex)[:200])
    res["__status__"] = onp.array(json.dumps(status))
    onp.savez_compressed(out_path, **res)
    print(json.dumps(status, indent=1))
'''


def main():
    if not os.path.isdir(REF):
        raise SystemExit("needs the reference at /root/reference (run in the build container)")
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "ray"))
        with open(os.path.join(tmp, "ray", "__init__.py"), "w") as f:
            f.write(RAY_STUB)
        child = os.path.join(tmp, "child.py")
        with open(child, "w") as f:
            f.write("ROOT = %r\n" % ROOT + CHILD)
        env = dict(os.environ)
        env.update({"RAMBA_NON_DIST": "1", "RAMBA_NUM_THREADS": "2", "RAMBA_BIG_DATA": "1",
                    "PYTHONPATH": tmp + ":" + REF, "NUMBA_CACHE_DIR": os.path.join(tmp, "nbcache")})
        only = sys.argv[1:]
        if only and only[0] == "--add":
            # `make_golden.py --add case1,case2`: run only these API cases under the reference and merge their outputs
            # into the existing api_golden.npz (the full regeneration takes minutes)
            import numpy as onp

            out = os.path.join(HERE, "api_golden.npz")
            alone = os.path.join(tmp, "alone.npz")
            env2 = dict(env)
            env2["RB_GOLDEN_CASES"] = only[1]
            subprocess.check_call([sys.executable, child, "api", alone], env=env2, cwd=tmp)
            z, z2 = dict(onp.load(out)), dict(onp.load(alone))
            status = json.loads(str(z["__status__"]))
            status.update(json.loads(str(z2["__status__"])))
            for name in only[1].split(","):
                for k in [k for k in z if k.startswith(name + "__")]:
                    del z[k]
            for k, v in z2.items():
                if k != "__status__":
                    z[k] = v
            z["__status__"] = onp.array(json.dumps(status))
            onp.savez_compressed(out, **z)
            print("merged", only[1], "into", out)
            return
        for mode, name in (("partition", "partition_golden.json"), ("programs", "programs_golden.npz"), ("api", "api_golden.npz")):
            if only and mode not in only:
                continue
            out = os.path.join(HERE, name)
            subprocess.check_call([sys.executable, child, mode, out], env=env, cwd=tmp)
            if mode == "api":
                # a case that fails inside the reference can leave its pending-op state broken for the cases
                # after it: give every failed case a second run alone in a fresh process
                import numpy as onp

                z = dict(onp.load(out))
                status = json.loads(str(z["__status__"]))
                for name in [n for n, st in status.items() if st != "ok"]:
                    alone = os.path.join(tmp, "alone.npz")
                    env2 = dict(env)
                    env2["RB_GOLDEN_CASES"] = name
                    subprocess.check_call([sys.executable, child, mode, alone], env=env2, cwd=tmp)
                    z2 = dict(onp.load(alone))
                    st2 = json.loads(str(z2["__status__"]))
                    status[name] = st2[name]
                    for k, v in z2.items():
                        if k != "__status__":
                            z[k] = v
                z["__status__"] = onp.array(json.dumps(status))
                onp.savez_compressed(out, **z)
                print(json.dumps(status, indent=1))
            print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
