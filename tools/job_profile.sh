#!/bin/bash
# GPU job: parity subset, bench, ncu launch list, ncu --set full of the post-classify kernels.  usage: job_profile.sh <tag>
tag=${1:-x}
out=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q > $out/${tag}_tests.log 2>&1; tail -3 $out/${tag}_tests.log
make -C tests/cpp -s cuberille_mgpu && tests/cpp/cuberille_mgpu 1 256 32 1 1 > $out/${tag}_mgpu1.log 2>&1; tail -3 $out/${tag}_mgpu1.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -3 $out/${tag}_bench.err; head -c 600 $out/${tag}_bench.json; echo
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launch.log 2>&1
for k in k_vertices k_faces k_sweep k_seg_scan; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o $out/${tag}_$k $B > $out/${tag}_ncu_$k.log 2>&1
done
ls -la $out | grep ${tag}_
