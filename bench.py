#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: fused-elementwise GB/s on the 1e9-element fp64
arange/sin/cos/mul/add chain (sample/test-ramba.py:12-19 of the reference), per GPU count.

    python bench.py --gpus N --steps K --warmup W             # our arm (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU path (oracle port)

A step = one pass of the hot path: `B = sin(A); C = cos(A); D = B*B + C**2; sync()` with A
resident in HBM (read A 8 B + write B, C, D 24 B = 32 algorithmic bytes / element).
Weak scaling: every GPU owns `--n` (default 1e9) elements of one global array.
Inputs (8 GB per GPU) are far larger than the 126 MB L2, so no explicit L2 flush is needed.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_ELEM = 32  # read A (8) + write B, C, D (24)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n", type=float, default=1e9, help="elements per GPU")
    p.add_argument("--cpu-n", type=float, default=1e8, help="elements of the bounded CPU sample")
    p.add_argument("--e2e-steps", type=int, default=2)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    return p.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """Start of the timed region (the sampler itself is started earlier: nvidia-smi needs ~100 ms
        to produce its first line)."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t1 = time.time()
        t0 = getattr(self, "t0", 0.0)
        inside = [ln for (ts, ln) in self.lines if t0 <= ts <= t1 + 0.05]
        if len(inside) < 2:  # region shorter than the sampling period: take the samples around it
            inside = [ln for (ts, ln) in self.lines if t0 - 0.3 <= ts <= t1 + 0.3]
        sm, mx, reasons, power = [], [], set(), []
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
def cpu_chain_baseline(n, iters, warm=1):
    """The reference's CPU path restated in C/OpenMP (oracle/fused_chain.c), all host threads,
    on a bounded sample of the same workload.  Returns (GB/s, threads, description)."""
    import numpy as np

    from oracle import chain  # bench.py's cpu_baseline / reference legs may execute the oracle

    n = int(n)
    # first touch in parallel (same static schedule as the timed loop): pages land next to their threads
    A = np.empty(n, dtype=np.float64)
    B = np.empty_like(A); C = np.empty_like(A); D = np.empty_like(A)
    chain.chain_f64(A, B, C, D, global_start=0, make_A=True)
    chain.calibrate_threads(A, B, C, D)  # all logical CPUs vs one thread per physical core: keep the faster
    for _ in range(warm):
        chain.chain_f64(A, B, C, D)
    best = []
    for _ in range(iters):
        t0 = time.perf_counter()
        chain.chain_f64(A, B, C, D)
        best.append(time.perf_counter() - t0)
    dt = sum(best) / len(best)
    return n * BYTES_PER_ELEM / dt / 1e9, chain.num_threads(), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = int(args.cpu_n)
    import numpy as np

    from oracle import chain

    A = np.empty(n, dtype=np.float64)
    B = np.empty_like(A); C = np.empty_like(A); D = np.empty_like(A)
    chain.chain_f64(A, B, C, D, global_start=0, make_A=True)  # parallel first touch
    chain.calibrate_threads(A, B, C, D)  # all logical CPUs vs one thread per physical core: keep the faster
    for _ in range(max(1, args.warmup)):
        chain.chain_f64(A, B, C, D)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        chain.chain_f64(A, B, C, D)
    dt = time.perf_counter() - t0
    val = n * BYTES_PER_ELEM * args.steps / dt / 1e9
    threads = chain.num_threads()
    sample = "oracle/fused_chain.c (C/OpenMP restatement of the reference's generated Numba loop), %d elements per step, %d threads" % (n, threads)
    out = {
        "impl": "reference", "metric": "fused-elementwise GB/s (fp64 sin/cos/mul/add chain)", "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "1e9-element fp64 arange/sin/cos/mul/add fused chain (sample/test-ramba.py loop); CPU arm runs a bounded %d-element sample per step" % n},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    import ramba_b200 as rb
    from ramba_b200 import _cabi, common
    from ramba_b200.runtime import RT

    W = common.num_workers
    rank = common.worker_num
    assert W == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, W)
    dev = RT.device
    dist = None
    if W > 1:
        RT.ensure_process_group()
        import torch.distributed as dist

    n_per_gpu = int(args.n)
    N = n_per_gpu * W

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    A = rb.arange(N) / 1000.0
    rb.sync()

    def step():
        B = rb.sin(A)
        C = rb.cos(A)
        D = B * B + C ** 2
        rb.sync()
        return B, C, D

    sampler = ClockSampler(common.local_rank)
    sampler.start()
    for _ in range(max(3, args.warmup)):
        out = step()
    del out
    barrier()
    sampler.mark()
    _cabi.reset_launch_count()
    RT.profile_events = []
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    gc.disable()  # like timeit: no collector pauses inside the timed region
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    gc.enable()
    launches = _cabi.launch_count()
    clocks = sampler.stop()
    dev_ms = e0.elapsed_time(e1)
    events = RT.profile_events
    RT.profile_events = None
    kern_ms = [a.elapsed_time(b) for (a, b, _) in events]
    t = torch.tensor([max(dev_ms / 1e3, wall)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    value = N * BYTES_PER_ELEM * args.steps / elapsed / 1e9
    # parity spot check of the timed result (cheap): D == 1 within 4 ulp on a slice
    B, C, D = out
    d_head = D[0:4096].asarray()
    assert np.max(np.abs(d_head - 1.0)) <= 4 * np.finfo(np.float64).eps
    del B, C, D, out

    peak, peak_src = measured_peak()
    k_ms = sum(kern_ms) / max(1, len(kern_ms))
    achieved = n_per_gpu * BYTES_PER_ELEM / (k_ms * 1e-3) / 1e9 if kern_ms else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if int(tj.get("elements_per_launch", 0)) == n_per_gpu:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "kernel": "vm_elementwise_kernel<8,1>", "kernel_ms": k_ms, "launches_timed": len(kern_ms),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": n_per_gpu * BYTES_PER_ELEM}

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        hA = torch.empty(n_per_gpu, dtype=torch.float64, pin_memory=True)
        hD = torch.empty(n_per_gpu, dtype=torch.float64, pin_memory=True)
        hA.copy_(torch.arange(n_per_gpu, dtype=torch.float64) * 0.001)
        hA_np, hD_np = hA.numpy(), hD.numpy()
        # The step is issued in chunks on two CUDA streams so that the upload of one chunk overlaps the
        # download of the previous one (PCIe is full duplex); every byte of A goes host->device and
        # every byte of D device->host inside the timed region.
        n_chunks = 8
        bounds = [n_per_gpu * c // n_chunks for c in range(n_chunks + 1)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

        def e2e_step():
            for c in range(n_chunks):
                lo, hi = bounds[c], bounds[c + 1]
                with torch.cuda.stream(streams[c % 2]):
                    if W == 1:
                        Ah = rb.fromarray(hA_np[lo:hi])
                    else:
                        Ah = rb.fromarray_local(hA_np[lo:hi], ((hi - lo) * W,))
                    Bh = rb.sin(Ah)
                    Ch = rb.cos(Ah)
                    Dh = Bh * Bh + Ch ** 2
                    if W == 1:
                        Dh.asarray(out=hD_np[lo:hi], non_blocking=True)
                    else:
                        rb.local_block_to_host(Dh, hD_np[lo:hi], non_blocking=True)
                    del Ah, Bh, Ch, Dh
            torch.cuda.synchronize(dev)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        assert abs(float(hD_np[12345]) - 1.0) < 1e-15
        e2e = {"value": N * BYTES_PER_ELEM * args.e2e_steps / dt / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": n_per_gpu * 8 * W, "d2h_bytes_per_step": n_per_gpu * 8 * W,
               "ms_per_step": dt / args.e2e_steps * 1e3, "steps": args.e2e_steps,
               "what": "A in pinned host memory -> fromarray (H2D) -> sin/cos/mul/add fused kernel -> D.asarray(out=pinned) (D2H); 8 chunks on 2 CUDA streams so that H2D and D2H overlap"}
        del hA, hD

    cpu = None
    if rank == 0 and W == 1 and not args.no_cpu:
        v, threads, dt = cpu_chain_baseline(args.cpu_n, iters=5)
        cpu = {"value": v, "unit": "GB/s", "cores": threads, "kind": "port",
               "sample": "oracle/fused_chain.c (C/OpenMP port of the reference's generated loop), %d elements x 5 iterations, %.2f s each" % (int(args.cpu_n), dt)}

    if rank == 0:
        out = {
            "metric": "fused-elementwise GB/s (fp64 sin/cos/mul/add chain)", "value": value, "unit": "GB/s",
            "n_gpus": W, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": elapsed / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "1e9-element fp64 arange/sin/cos/mul/add fused chain per B200 (BASELINE configs[1]) x %d GPU(s); timed loop of sample/test-ramba.py" % W,
                       "elements_per_gpu": n_per_gpu, "global_elements": N, "bytes_per_element": BYTES_PER_ELEM,
                       "l2": "inputs (8 GB/GPU) >> 126 MB L2, no flush needed", "parallelism": "block partition, %d rank(s), no collective" % W},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "device_ms_total": dev_ms,
        }
        print(json.dumps(out))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
