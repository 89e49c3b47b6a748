timeout 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -5
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-extras --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["n_points"], d["n_quads"])'
echo "FUSE 1"; CUB_FUSE=1 $B | python -c "$P"
echo "FUSE 1 DBG 1 (producers only)"; CUB_FUSE_DBG=1 $B | python -c "$P"
echo "FUSE 1 DBG 2 (consumers only)"; CUB_FUSE_DBG=2 $B | python -c "$P"
for tz in 8 16; do echo "FUSE 1 TZ $tz"; CUB_FUSE_TZ=$tz $B | python -c "$P"; done
for c in 2 3; do echo "FUSE 1 CTAS $c"; CUB_FUSE_CTAS_PER_SM=$c $B | python -c "$P"; done
