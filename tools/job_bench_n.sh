#!/bin/bash
# bench.py on N GPUs of one box (torchrun), nothing else.  usage: job_bench_n.sh <tag> <n_gpus>
tag=$1; N=$2; out=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -3 $out/${tag}_bench.err; head -c 400 $out/${tag}_bench.json; echo
