#!/usr/bin/env python
"""A few steps of ONE BASELINE config through the public API, nothing else - the command ncu wraps
(profiles/capture.sh).  Sizes as in bench.py."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import ramba_b200 as rb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--n", type=float, default=1e9)
ap.add_argument("--scale", type=float, default=1.0)
args = ap.parse_args()
wl = bench.CLASSES[args.config](rb, 1, int(args.n), args)
for _ in range(args.steps):
    wl.step()
rb.sync()
print("ok", args.config, wl.check())
