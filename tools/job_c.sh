#!/bin/bash
tag=${1:-x}; out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -4 $out/${tag}_tests.log
python tools/fuzz_parity.py 60 33 > $out/${tag}_fuzz.log 2>&1; tail -2 $out/${tag}_fuzz.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -3 $out/${tag}_bench.err
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_project --launch-skip 0 -c 1 -f -o $out/${tag}_k_project python tools/proj_only.py > $out/${tag}_ncu_proj.log 2>&1
ls -la $out | grep ${tag}_
