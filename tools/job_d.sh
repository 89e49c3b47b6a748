#!/bin/bash
tag=${1:-x}; out=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q > $out/${tag}_tests.log 2>&1; tail -2 $out/${tag}_tests.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -3 $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launch.log 2>&1
for k in k_faces k_assign; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 4 -c 1 -f -o $out/${tag}_$k $B > $out/${tag}_ncu_$k.log 2>&1
done
for a in "512 0 0 0" "512 1 1 0"; do tests/cpp/CuberilleTest01 DropInBench $a; done > $out/${tag}_dropin.log 2>&1; cat $out/${tag}_dropin.log
