import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as onp
import ramba_b200 as rb
import test_stencil_tile as T
name = sys.argv[1] if len(sys.argv) > 1 else "lap3d_odd_f32"
prog = dict(T.CASES)[name]
got = prog(rb)
exp = prog(onp)
for g, e in zip(got, exp):
    print(name, g.shape, g.dtype, "max abs diff", float(onp.max(onp.abs(g.astype(onp.float64) - e.astype(onp.float64)))))
