#!/usr/bin/env python
"""BASELINE configs[0] / the reference's README example (sample/test-ramba.py) with the import switched:

    python examples/sincos_chain.py [N]            # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/sincos_chain.py 8e9

Needs a B200 and the built library (python -c "import __graft_entry__ as g; g.build()")."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import ramba_b200 as np  # noqa: E402  (the reference: `import ramba as np`)

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100 * 1000 * 1000

np.sync()
t0 = time.time()
A = np.arange(N) / 1000.0
np.sync()
print("Initialize array time:", time.time() - t0)

for i in range(5):
    t0 = time.time()
    B = np.sin(A)
    C = np.cos(A)
    D = B * B + C ** 2
    np.sync()
    dt = time.time() - t0
    print("Iteration", i + 1, "time:", dt, "-> %.1f GB/s of 32 B/element" % (N * 32 / dt / 1e9))
