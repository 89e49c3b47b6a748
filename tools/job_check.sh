#!/bin/bash
# last check of a round: the whole -m gpu suite, smoke(), a short bench.  usage: job_check.sh <tag>
tag=${1:-x}; out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -3 $out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 5 --no-e2e --no-extras --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -2 $out/${tag}_bench.err; head -c 250 $out/${tag}_bench.json; echo
