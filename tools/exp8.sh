timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for c in 5 6; do for r in 14 18 24; do echo "CTAS $c REFILL $r"; CUB_PROJ_CTAS_PER_SM=$c CUB_PROJ_REFILL=$r REPS=1 timeout 120 python tools/proj_only.py | python -c "import sys; d=eval(sys.stdin.read()); print(d['project'])"; done; done
python tools/fuzz_parity.py 30 93 | tail -1
