for r in 18; do echo "REFILL $r"; CUB_PROJ_REFILL=$r REPS=1 timeout 120 python tools/proj_only.py; done
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-extras --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["n_points"], d["n_quads"])'
echo "BENCH"; $B | python -c "$P"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/fuzz_parity.py 40 81 | tail -1; python tools/fuzz_parity.py 30 82 big | tail -1
