#!/usr/bin/env python
"""Secondary BASELINE configs (3, 4, 5 of BASELINE.json) through the public API, with the synthetic
inputs of SURVEY.md §8(d): exact closed-form checks + algorithmic GB/s.  Not the driver's bench
(that is /bench.py = config 2); run by hand:

    python benchmarks/configs.py [--scale 1.0]                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N benchmarks/configs.py   # N GPUs
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as onp  # noqa: E402
import torch  # noqa: E402

import ramba_b200 as np  # noqa: E402
from ramba_b200 import common  # noqa: E402
from ramba_b200.runtime import RT  # noqa: E402


KERNEL_MS = {}


def timed(fn, iters=5, warm=2, name=None):
    """(wall seconds per iteration incl. Python, result); also records the summed CUDA-event time of
    the library's launches per iteration in KERNEL_MS[name]."""
    for _ in range(warm):
        r = fn()
    np.sync()
    torch.cuda.synchronize()
    RT.profile_events = []
    t0 = time.perf_counter()
    for _ in range(iters):
        r = fn()
    np.sync()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    ev = RT.profile_events
    RT.profile_events = None
    if name is not None:
        KERNEL_MS[name] = {"kernel_ms_per_iter": sum(a.elapsed_time(b) for (a, b, _) in ev) / iters, "launches_per_iter": len(ev) / iters}
    return dt, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="linear size factor (1.0 = BASELINE sizes)")
    args = ap.parse_args()
    W, rank = common.num_workers, common.worker_num
    RT.device
    if W > 1:
        RT.ensure_process_group()
    out = {"n_gpus": W}

    # ---- config 3: 32768^2 fp32 elementwise + global sum
    n = int(32768 * args.scale)
    X = np.fromfunction(lambda i, j: (i * 131 + j * 31) % 4, (n, n), dtype=onp.float32)
    np.sync()
    dt, s = timed(lambda: float((X * 2.0 + 1.0).sum(asarray=True).asarray()[0]), name="config3")
    # closed form: sum over i,j of 2*((131 i + 31 j) mod 4) + 1 — computed exactly with integers on the host for a stripe
    ii = onp.arange(n, dtype=onp.int64)
    cnt = onp.zeros(4, dtype=onp.int64)
    ri = (ii * 131) % 4
    rj = (ii * 31) % 4
    ci = onp.bincount(ri, minlength=4)
    cj = onp.bincount(rj, minlength=4)
    for a in range(4):
        for b in range(4):
            cnt[(a + b) % 4] += ci[a] * cj[b]
    expect = float(sum(cnt[v] * (2 * v + 1) for v in range(4)))
    out["config3"] = {"shape": [n, n], "seconds": dt, "GBps_algorithmic_4B": n * n * 4 / dt / 1e9, "exact": bool(s == float(onp.float32(expect))), "value": float(s), "expected": expect}

    # ---- config 5: (2^20, 4096) fp32 broadcast-add + axis-0 sum
    r, c = int((1 << 20) * args.scale), 4096
    M = np.fromfunction(lambda i, j: (i + 3 * j) % 8, (r, c), dtype=onp.float32)
    v = (np.arange(c) % 8).astype(onp.float32)
    np.sync()
    dt, res = timed(lambda: (M + v).sum(axis=0), iters=3, warm=1, name="config5")
    got = res.asarray()
    j = onp.arange(c, dtype=onp.int64)
    # column sum of (i + 3j) % 8 over i in [0, r): r/8 full cycles (r is a multiple of 8) -> 28 * r/8, plus v
    assert r % 8 == 0
    expect5 = 28 * (r // 8) + (j % 8) * r
    out["config5"] = {"shape": [r, c], "seconds": dt, "GBps_algorithmic_4B": r * c * 4 / dt / 1e9,
                      "exact": bool(onp.array_equal(got, onp.asarray(expect5, dtype=onp.float32)))}
    del M, v, res

    # ---- config 4: 1024^3 fp32 7-point Laplacian through slice views
    m = int(1024 * args.scale)
    U = np.fromfunction(lambda i, j, k: (i + 2 * j + 3 * k) % 64, (m, m, m), dtype=onp.float32)
    V = np.zeros((m, m, m), dtype=onp.float32)
    np.sync()

    def lap():
        V[1:-1, 1:-1, 1:-1] = (U[:-2, 1:-1, 1:-1] + U[2:, 1:-1, 1:-1] + U[1:-1, :-2, 1:-1] + U[1:-1, 2:, 1:-1]
                               + U[1:-1, 1:-1, :-2] + U[1:-1, 1:-1, 2:] - 6.0 * U[1:-1, 1:-1, 1:-1])
        return None

    dt, _ = timed(lap, iters=3, warm=1, name="config4")
    # check a sub-block against NumPy
    sub = V[1:9, 1:9, 1:m - 1].asarray()
    i, jj, k = onp.meshgrid(onp.arange(0, 10), onp.arange(0, 10), onp.arange(m), indexing="ij")
    u = ((i + 2 * jj + 3 * k) % 64).astype(onp.float32)
    ref = (u[:-2, 1:-1, 1:-1] + u[2:, 1:-1, 1:-1] + u[1:-1, :-2, 1:-1] + u[1:-1, 2:, 1:-1] + u[1:-1, 1:-1, :-2] + u[1:-1, 1:-1, 2:]
           - 6.0 * u[1:-1, 1:-1, 1:-1]).astype(onp.float32)
    out["config4"] = {"shape": [m, m, m], "seconds": dt, "GBps_algorithmic_8B": (m - 2) ** 3 * 8 / dt / 1e9,
                      "exact": bool(onp.array_equal(sub, ref)), "bytes_sent_per_rank": RT.bytes_sent}
    for k_, v_ in KERNEL_MS.items():
        out[k_].update(v_)
    if rank == 0:
        print(json.dumps(out))
    if W > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
