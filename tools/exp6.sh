timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -x -q 2>&1 | tail -2
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-extras --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["n_points"], d["n_quads"])'
echo "BENCH"; $B | python -c "$P"
echo "BENCH FUSE=0"; CUB_FUSE=0 $B | python -c "$P"
python tools/fuzz_parity.py 30 91 | tail -1
