#!/usr/bin/env python
"""Reduce the source page of an ncu capture (`ncu -i X.ncu-rep --page source --csv`) to what matters for an HBM-bound
kernel: thread-instructions executed per output element by opcode class, and the instructions that hold the stall samples.
Usage: python profiles/hot_sass.py gpurun_out/r02_cfg4.ncu-rep <elements per launch> > profiles/r02_cfg4_hot_sass.txt"""
import csv
import io
import re
import subprocess
import sys

rep, elements = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
insns = []
for r in rows[2:]:
    if len(r) < len(hdr) - 2:
        continue
    try:
        insns.append((r[ci["Source"]].strip(), float(r[ci["Thread Instructions Executed"]]), float(r[ci["Instructions Executed"]]), float(r[ci["# Samples"]])))
    except ValueError:
        pass
tot_thread = sum(x[1] for x in insns)
tot_samples = sum(x[3] for x in insns) or 1.0
print("kernel: %s" % kname)
print("elements per launch: %.0f" % elements)
print("thread-instructions executed per element: %.1f   (warp-instructions per element: %.2f)" % (tot_thread / elements, sum(x[2] for x in insns) / elements))
cls = {}
for src, ti, wi, sm in insns:
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(1) if m else "?"
    c = cls.setdefault(op, [0.0, 0.0])
    c[0] += ti
    c[1] += sm
print("\nper opcode: thread-instructions per element, share of instructions, share of stall samples")
for op, (ti, sm) in sorted(cls.items(), key=lambda kv: -kv[1][0])[:28]:
    print("  %-12s %7.2f  %5.1f%%  %5.1f%%" % (op, ti / elements, 100 * ti / tot_thread, 100 * sm / tot_samples))
print("\ninstructions holding the most stall samples")
for src, ti, wi, sm in sorted(insns, key=lambda x: -x[3])[:25]:
    print("  %5.1f%%  %6.2f/elem  %s" % (100 * sm / tot_samples, ti / elements, src[:110]))
