#!/usr/bin/env python
"""Development probe: config 4 (7-point Laplacian through slice views) at a given size on one GPU; every iteration is
flushed (sync) so that each one is one launch; CUDA-event time per launch; slabs checked against NumPy."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as onp  # noqa: E402
import torch  # noqa: E402

import ramba_b200 as np  # noqa: E402
from ramba_b200.runtime import RT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--dtype", default="float32")
args = ap.parse_args()
m = args.n
dt = onp.dtype(args.dtype)
U = np.fromfunction(lambda i, j, k: (i + 2 * j + 3 * k) % 64, (m, m, m), dtype=dt)
V = np.zeros((m, m, m), dtype=dt)
np.sync()


def lap():
    V[1:-1, 1:-1, 1:-1] = (U[:-2, 1:-1, 1:-1] + U[2:, 1:-1, 1:-1] + U[1:-1, :-2, 1:-1] + U[1:-1, 2:, 1:-1]
                           + U[1:-1, 1:-1, :-2] + U[1:-1, 1:-1, 2:] - 6.0 * U[1:-1, 1:-1, 1:-1])
    np.sync()


lap()
lap()
RT.profile_events = []
for _ in range(args.iters):
    lap()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for (a, b, _) in RT.profile_events]
RT.profile_events = None
ok = True
for z0 in (0, m // 2 - 3, m - 8):
    sub = V[z0:z0 + 8].asarray()
    i, j, k = onp.meshgrid(onp.arange(z0 - 1, z0 + 9), onp.arange(m), onp.arange(m), indexing="ij")
    u = ((i + 2 * j + 3 * k) % 64).astype(dt)
    ref = onp.zeros((8, m, m), dtype=dt)
    full = (u[:-2, 1:-1, 1:-1] + u[2:, 1:-1, 1:-1] + u[1:-1, :-2, 1:-1] + u[1:-1, 2:, 1:-1] + u[1:-1, 1:-1, :-2] + u[1:-1, 1:-1, 2:]
            - 6.0 * u[1:-1, 1:-1, 1:-1]).astype(dt)
    ref[:, 1:-1, 1:-1] = full
    if z0 == 0:
        ref[0] = 0
    if z0 + 8 == m:
        ref[-1] = 0
    ok = ok and bool(onp.array_equal(sub, ref))
byt = (m - 2) ** 3 * 2 * dt.itemsize
print(json.dumps({"n": m, "dtype": str(dt), "launch_ms": ms, "best_ms": min(ms), "GBps": byt / min(ms) / 1e6, "frac_of_6486": byt / min(ms) / 1e6 / 6486.1, "exact": ok,
                  "launches": len(ms)}))
