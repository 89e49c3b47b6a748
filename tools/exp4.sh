timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -x -q 2>&1 | tail -3
for r in 32 26 22 18 14 10 4; do echo "REFILL $r"; CUB_PROJ_REFILL=$r REPS=1 timeout 120 python tools/proj_only.py; done
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-extras --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["n_points"], d["n_quads"])'
echo "BENCH"; $B | python -c "$P"
