B="python bench.py --steps 10 --warmup 3 --no-e2e --no-extras --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"])'
for c in 2 3 4 6 8 32; do echo "K1_CTAS $c"; CUB_K1_CTAS_PER_SM=$c $B | python -c "$P"; done
for c in 1 2 3 4 6 8; do echo "CORUN $c"; CUB_EXP_CORUN=$c $B | python -c "$P"; done
