#!/bin/bash
# Run ON THE GPU BOX (gpurun --timeout 1500 -- 'bash profiles/capture.sh'): per BASELINE config
#   * the launch list (device time of every launch, cold-cache and serialised: compare shares),
#   * ONE `ncu --set full` capture of the dominant kernel (dram bytes, stalls, source page),
# into gpurun_out/; profiles/summarise.py (run in the build container) turns them into the committed summaries.
set -x
mkdir -p gpurun_out
declare -A KERN=( [2]="vm_elementwise_kernel" [3]="mapred_global" [4]="stencil_t" [5]="mapred_columns" )
for c in ${CONFIGS:-2 3 4 5}; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_cfg$c.csv \
      python benchmarks/one_config.py --config $c --steps 4 > gpurun_out/r02_launches_cfg$c.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:${KERN[$c]} -s 2 -c 1 -f -o gpurun_out/r02_cfg$c \
      python benchmarks/one_config.py --config $c --steps 4 > gpurun_out/r02_ncu_cfg$c.log 2>&1
done
sha256sum ramba_b200/lib/libramba_b200.so > gpurun_out/r02_lib.sha256
ls -la gpurun_out | tail -20
