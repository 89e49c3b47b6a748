#!/bin/bash
# --set full capture of the count-phase kernels (fused classify + sweep, segment scan).  usage: job_ncu_count.sh <tag>
tag=${1:-x}; out=gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"k_classify_sweep|k_seg_scan" --launch-skip 6 -c 2 -f -o $out/${tag}_count $B > $out/${tag}_ncu_count.log 2>&1
ncu -i $out/${tag}_count.ncu-rep --page raw --csv > $out/${tag}_count_raw.csv 2>/dev/null
grep PROF $out/${tag}_ncu_count.log | head
