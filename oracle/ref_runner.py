#!/usr/bin/env python
"""ORACLE / REFERENCE ARM (test and measurement infrastructure, not product code).

Runs the UNMODIFIED reference (Python-for-HPC/ramba, installed by oracle/build_ref.sh into the git-ignored
oracle/_ref/) on one of the BASELINE workloads, through the reference's own public API and stock code path:
RAMBA_NON_DIST=1 (one worker in-process, ramba/common.py:32-45), Numba `parallel=True` kernels on
RAMBA_NUM_THREADS host threads (ramba/ramba.py:266-267, 363), the README's timing loop
(sample/test-ramba.py:12-19).  bench.py launches it as a subprocess (`--impl reference`, and the cpu_baseline leg):

    PYTHONPATH=oracle/_ref:oracle/ray_stub RAMBA_NON_DIST=1 RAMBA_NUM_THREADS=<cores> python oracle/ref_runner.py \
        --config 2 --n 1e9 --steps 5 --warmup 1

and reads ONE JSON line: seconds per timed step, a checksum of the result, threads.  Ray is replaced by the stub in
oracle/ray_stub because NON_DIST mode never calls it."""
import argparse
import json
import os
import sys
import time
import warnings

warnings.filterwarnings("ignore")
ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--n", type=float, default=1e8, help="config 2: elements; 3: rows=cols; 4: edge; 5: rows")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=1)
args = ap.parse_args()

import numpy as onp  # noqa: E402
import ramba  # noqa: E402  (the reference)

n = int(args.n)
cfg = args.config
if cfg == 2:
    A = ramba.arange(n) / 1000.0
    ramba.sync()

    def step():
        B = ramba.sin(A)
        C = ramba.cos(A)
        D = B * B + C ** 2
        ramba.sync()
        return D

    def check(D):
        d = D[0:4096].asarray()
        return float(onp.abs(d - 1.0).max())
elif cfg == 3:
    X = ramba.fromfunction(lambda i, j: (i * 131 + j * 31) % 4, (n, n), dtype=onp.float32)
    ramba.sync()

    def step():
        return (X * 2.0 + 1.0).sum()

    def check(s):
        return float(s)
elif cfg == 4:
    U = ramba.fromfunction(lambda i, j, k: (i + 2 * j + 3 * k) % 64, (n, n, n), dtype=onp.float32)
    V = ramba.zeros((n, n, n), dtype=onp.float32)
    ramba.sync()

    def step():
        V[1:-1, 1:-1, 1:-1] = (U[:-2, 1:-1, 1:-1] + U[2:, 1:-1, 1:-1] + U[1:-1, :-2, 1:-1] + U[1:-1, 2:, 1:-1]
                               + U[1:-1, 1:-1, :-2] + U[1:-1, 1:-1, 2:] - 6.0 * U[1:-1, 1:-1, 1:-1])
        ramba.sync()
        return V

    def check(V):
        return float(V[1:3, 1:3, 1:3].asarray().sum())
elif cfg == 5:
    c = 4096
    M = ramba.fromfunction(lambda i, j: (i + 3 * j) % 8, (n, c), dtype=onp.float32)
    v = (ramba.arange(c) % 8).astype(onp.float32)
    ramba.sync()

    def step():
        r = (M + v).sum(axis=0)
        ramba.sync()
        return r

    def check(r):
        return float(r.asarray().sum())
else:
    raise SystemExit("unknown config")

for _ in range(max(1, args.warmup)):  # first call = Numba JIT (README.md:63)
    out = step()
times = []
for _ in range(args.steps):
    t0 = time.perf_counter()
    out = step()
    times.append(time.perf_counter() - t0)
print(json.dumps({"config": cfg, "n": n, "steps": args.steps, "seconds": times, "best": min(times), "median": sorted(times)[len(times) // 2],
                  "threads": int(os.environ.get("RAMBA_NUM_THREADS", "0")), "check": check(out), "ramba_file": ramba.__file__}))
