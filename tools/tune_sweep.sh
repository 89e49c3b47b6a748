set -e
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 0 1; do
  echo "COUNT_CFG=$c"; CUB_COUNT_CFG=$c python bench.py --steps 5 --warmup 2 --no-e2e --no-extras --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_quads'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
for c in 0 5 6 7 8; do
  echo "EMIT_CFG=$c"; CUB_EMIT_CFG=$c python bench.py --steps 5 --warmup 2 --no-e2e --no-extras --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_quads'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
