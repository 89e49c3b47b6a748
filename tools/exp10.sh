B="timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-extras --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["n_points"], d["n_quads"])'
echo "BENCH"; $B | python -c "$P"
echo "BENCH TZ6"; CUB_FUSE_TZ=6 $B | python -c "$P"
echo "BENCH TZ12"; CUB_FUSE_TZ=12 $B | python -c "$P"
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -2
