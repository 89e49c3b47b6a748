#!/bin/bash
# single-GPU evidence job: full tests, full bench, ncu launch list + full summary of every kernel of a step.  usage: job_final.sh <tag>
tag=${1:-x}; out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -4 $out/${tag}_tests.log
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -3 $out/${tag}_bench.err; head -c 300 $out/${tag}_bench.json; echo
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launch.log 2>&1
# one --set full capture of each kernel of a (synchronous-API) step: launches 2.. of the bench are warm
ncu --set full --clock-control none --import-source on --launch-skip 40 -c 9 -f -o $out/${tag}_step $B > $out/${tag}_ncu_step.log 2>&1
ncu -i $out/${tag}_step.ncu-rep --page raw --csv > $out/${tag}_step_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_project -c 1 -f -o $out/${tag}_k_project python tools/proj_only.py > $out/${tag}_ncu_proj.log 2>&1
ncu -i $out/${tag}_k_project.ncu-rep --page raw --csv > $out/${tag}_proj_raw.csv 2>/dev/null
ls -la $out | grep ${tag}_
