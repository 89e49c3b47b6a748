#!/bin/bash
# multi-GPU correctness with programmatic dependent launch forced on: usage job_mgpu_pdl.sh <tag> <n_gpus>
tag=$1; N=$2; out=gpurun_out
export CUB_PDL=1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "${N}-" > $out/${tag}_tests.log 2>&1; tail -3 $out/${tag}_tests.log
make -C tests/cpp -s cuberille_mgpu
timeout 120 tests/cpp/cuberille_mgpu $N 512 64 0 0 > $out/${tag}_mgpu_cpp.log 2>&1
timeout 120 tests/cpp/cuberille_mgpu $N 512 64 1 1 >> $out/${tag}_mgpu_cpp.log 2>&1; cat $out/${tag}_mgpu_cpp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-e2e > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -2 $out/${tag}_bench.err; head -c 300 $out/${tag}_bench.json; echo
