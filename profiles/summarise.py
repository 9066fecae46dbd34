#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ (profiles/capture.sh) into the committed summaries:
  profiles/r02_traffic.json          dram bytes of the dominant kernel per config + sha256 of the library they were taken with
  profiles/r02_ncu_cfgN.csv          the metrics that matter (time, dram, issue, occupancy, stalls) of that launch
  profiles/r02_launches_cfgN.csv     copy of the launch list
  profiles/r02_sass_mnemonics.txt    TMA / bulk-copy / mbarrier mnemonics per kernel (cuobjdump -sass)
Run in the build container (needs ncu, cuobjdump; no GPU)."""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
GO = os.path.join(ROOT, "gpurun_out")
KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__dram_throughput.*|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct|smsp__inst_executed\.sum|launch__registers_per_thread|"
                  r"launch__grid_size|launch__block_size|launch__shared_mem_per_block_dynamic|launch__occupancy_limit.*|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|"
                  r"lts__t_bytes\.sum|lts__t_sector_hit_rate\.pct|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_.*|sm__inst_executed_pipe_.*\.sum)$")


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    return names, units, vals


def main():
    traffic = {}
    sha = None
    if "--sass-only" in sys.argv:  # (the library changed after the last capture: refresh the SASS evidence only)
        return sass_mnemonics()
    p = os.path.join(GO, "r02_lib.sha256")
    if os.path.exists(p):
        sha = open(p).read().split()[0]
    for c in (2, 3, 4, 5):
        rep = os.path.join(GO, "r02_cfg%d.ncu-rep" % c)
        if not os.path.exists(rep):
            continue
        names, units, vals = raw_rows(rep)
        if not vals:
            continue
        v = vals[0]
        d = {n: (x, u) for n, u, x in zip(names, units, v)}
        kname = d.get("Kernel Name", ("?", ""))[0]

        def num(key):
            x, u = d[key]
            f = float(x.replace(",", ""))
            mul = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}.get(u, 1)
            return f * mul

        ent = {"kernel": kname, "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
               "duration_s_under_ncu": num("gpu__time_duration.sum")}
        traffic["config%d" % c] = ent
        with open(os.path.join(ROOT, "profiles", "r02_ncu_cfg%d.csv" % c), "w") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit", "value"])
            w.writerow(["Kernel Name", "", kname])
            for n, u, x in zip(names, units, v):
                if KEEP.match(n):
                    w.writerow([n, u, x])
        ll = os.path.join(GO, "r02_launches_cfg%d.csv" % c)
        if os.path.exists(ll):
            shutil.copy(ll, os.path.join(ROOT, "profiles", "r02_launches_cfg%d.csv" % c))
    if traffic:
        traffic["lib_sha256"] = sha
        traffic["how"] = "ncu --set full --clock-control none -k regex:<kernel> -s 2 -c 1 python benchmarks/one_config.py --config N (profiles/capture.sh); one launch each"
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
        print(json.dumps(traffic, indent=1))
    sass_mnemonics()


def sass_mnemonics():
    so = os.path.join(ROOT, "ramba_b200", "lib", "libramba_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
    counts, fn = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"\b(UTMALDG\S*|UBLKCP\S*|LDGSTS\S*|SYNCS\S*|ARRIVES\S*|LDG\.E\.128\S*)", line)
        if m and fn:
            counts[(fn, m.group(1))] = counts.get((fn, m.group(1)), 0) + 1
    with open(os.path.join(ROOT, "profiles", "r02_sass_mnemonics.txt"), "w") as f:
        import hashlib

        f.write("# cuobjdump -sass ramba_b200/lib/libramba_b200.so (sha256 %s):\n# async-copy / TMA / mbarrier / 128-bit load mnemonics per kernel (count)\n"
                % hashlib.sha256(open(so, "rb").read()).hexdigest())
        for (fn, mn), n in sorted(counts.items()):
            dem = subprocess.run(["c++filt", fn], stdout=subprocess.PIPE, text=True).stdout.strip()
            f.write("%-28s %4d  %s\n" % (mn, n, dem[:150]))
    print("wrote profiles/r02_sass_mnemonics.txt")


if __name__ == "__main__":
    main()
