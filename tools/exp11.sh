B="timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-extras --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"]["classify"])'
for c in 3 4; do for tz in 4 6 8; do echo "CTAS $c TZ $tz"; CUB_FUSE_CTAS_PER_SM=$c CUB_FUSE_TZ=$tz $B | python -c "$P"; done; done
