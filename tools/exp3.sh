B="timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-extras --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["n_points"], d["n_quads"])'
for tz in 4 6 8 12; do echo "TZ $tz"; CUB_FUSE_TZ=$tz $B | python -c "$P"; done
for tz in 4 8; do echo "CTAS 3 TZ $tz"; CUB_FUSE_CTAS_PER_SM=3 CUB_FUSE_TZ=$tz $B | python -c "$P"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_classify_sweep -c 1 -o gpurun_out/r2u_fused python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r2u_ncu.log 2>&1
ncu -i gpurun_out/r2u_fused.ncu-rep --page raw --csv > gpurun_out/r2u_fused_raw.csv 2>/dev/null
