#!/usr/bin/env python
"""One secondary config in isolation (for ncu): python benchmarks/one_config.py {3|5} [scale]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as onp  # noqa: E402

import ramba_b200 as np  # noqa: E402

which = sys.argv[1]
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
if which == "3":
    n = int(32768 * scale)
    X = np.fromfunction(lambda i, j: (i * 131 + j * 31) % 4, (n, n), dtype=onp.float32)
    np.sync()
    for _ in range(3):
        s = float((X * 2.0 + 1.0).sum(asarray=True).asarray()[0])
    print(s)
else:
    r, c = int((1 << 20) * scale), 4096
    M = np.fromfunction(lambda i, j: (i + 3 * j) % 8, (r, c), dtype=onp.float32)
    v = (np.arange(c) % 8).astype(onp.float32)
    np.sync()
    for _ in range(3):
        res = (M + v).sum(axis=0)
        np.sync()
    print(res.asarray()[:4])
