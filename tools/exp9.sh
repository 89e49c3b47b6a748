timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -x -q 2>&1 | tail -2
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-extras --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["n_points"], d["n_quads"])'
echo "BENCH"; $B | python -c "$P"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r2w_ncu_launch.log 2>&1
python tools/fuzz_parity.py 30 95 | tail -1
