#!/usr/bin/env python
"""Run the seeded fuzz programs of tests/_dag_fuzz.py, _limit_fuzz.py and _expr_fuzz.py under the REAL reference
(Python-for-HPC/ramba installed in oracle/_ref, RAMBA_NON_DIST=1, Ray stubbed - like tests/golden/make_golden.py) and keep
its outputs as fixtures: tests/golden/fuzz_golden.npz.  One process per program (a failure inside the reference leaves its
fuser in an undefined state).  Programs the reference cannot run here (its NumPy-2 incompatibilities - np.NINF in min / max -
and its own defects, e.g. `missing argument 'ramba_tmp_var_…'` for the sum of a temporary) are recorded with the reason, and so
are programs the reference runs to a result that is NOT NumPy's (ordering defects of its own fuser on particular sequences of
pending statements: every statement of those programs is right there in isolation).

    python tests/golden/make_fuzz_golden.py            (from the repo root; needs oracle/_ref, i.e. /root/reference + build())"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
FAMILIES = {"dag_program": ("_dag_fuzz", 24), "limit_program": ("_limit_fuzz", 24), "view_program": ("_expr_fuzz", 30),
            "shape_program": ("_expr_fuzz", 24), "partition_program": ("_expr_fuzz", 16),
            # not comparable with NumPy by design (the reference's typing / rounding rules): kept whenever the reference runs it
            "typing_program": ("_expr_fuzz", 40), "typing2_program": ("_expr_fuzz", 40)}
NOT_NUMPY = {"typing_program", "typing2_program"}

CHILD = r'''
import sys, warnings, json
warnings.filterwarnings("ignore")
sys.path.insert(0, sys.argv[1])
import numpy as onp
import ramba   # the reference
mod = __import__(sys.argv[2])
fn = getattr(mod, sys.argv[3]); seed = int(sys.argv[4])
out = fn(ramba, seed)
twin = fn(onp, seed) if sys.argv[6] == "1" else out
same = len(out) == len(twin) and all(onp.asarray(a).shape == onp.asarray(b).shape and onp.array_equal(onp.asarray(a), onp.asarray(b))
                                     for a, b in zip(out, twin))
onp.savez(sys.argv[5], __same__=onp.array(same), **{"o%d" % i: onp.asarray(x) for i, x in enumerate(out)})
'''


def main():
    import numpy as onp

    env = dict(os.environ)
    env.update({"PYTHONPATH": os.path.join(ROOT, "oracle", "_ref") + os.pathsep + os.path.join(ROOT, "oracle", "ray_stub"),
                "RAMBA_NON_DIST": "1", "RAMBA_NUM_THREADS": "2"})
    res, status = {}, {}
    tmp = os.path.join(HERE, "_fuzz_tmp.npz")
    for fam, (mod, n) in FAMILIES.items():
        for seed in range(n):
            name = "%s_%d" % (fam, seed)
            p = subprocess.run([sys.executable, "-c", CHILD, os.path.join(ROOT, "tests"), mod, fam, str(seed), tmp, "0" if fam in NOT_NUMPY else "1"], env=env,
                               capture_output=True, text=True, timeout=600)
            if p.returncode == 0 and os.path.exists(tmp):
                z = onp.load(tmp)
                if bool(z["__same__"]):
                    for k in z.files:
                        if k != "__same__":
                            res["%s__%s" % (name, k)] = z[k]
                    status[name] = "ok"
                else:
                    # the reference ran but its result is not NumPy's: in every case looked at, each statement alone is right
                    # there and a particular SEQUENCE of pending statements is not (ordering defects of its fuser / DAG);
                    # NumPy is the specification, the outputs are not kept
                    status[name] = "reference differs from NumPy"
                os.remove(tmp)
            else:
                last = [ln for ln in p.stderr.strip().splitlines() if ln.strip()][-1:] or ["?"]
                status[name] = "reference failed: " + last[0][:160]
            print(name, status[name], flush=True)
    res["__status__"] = onp.array(json.dumps(status))
    onp.savez_compressed(os.path.join(HERE, "fuzz_golden.npz"), **res)
    print(sum(v == "ok" for v in status.values()), "of", len(status), "programs run under the reference")


if __name__ == "__main__":
    main()
