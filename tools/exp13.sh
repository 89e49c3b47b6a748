B="timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-extras --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"])'
for p in 0 1; do echo "PDL $p"; CUB_PDL=$p $B | python -c "$P"; CUB_PDL=$p N=8 python tools/slab_time.py; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -x -q 2>&1 | tail -2
python tools/fuzz_parity.py 25 101 | tail -1
