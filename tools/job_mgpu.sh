#!/bin/bash
# multi-GPU job: usage job_mgpu.sh <tag> <n_gpus> [tests]
tag=$1; N=$2; out=gpurun_out
if [ "$3" == "tests" ]; then
  python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -4 $out/${tag}_tests.log
else
  python -m pytest tests/test_gpu_multi.py -x -q -k "${N}-" > $out/${tag}_tests.log 2>&1; tail -4 $out/${tag}_tests.log
fi
make -C tests/cpp -s cuberille_mgpu
tests/cpp/cuberille_mgpu $N 512 64 0 0 > $out/${tag}_mgpu_cpp.log 2>&1
tests/cpp/cuberille_mgpu $N 512 64 1 1 >> $out/${tag}_mgpu_cpp.log 2>&1; cat $out/${tag}_mgpu_cpp.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -5 $out/${tag}_bench.err; head -c 600 $out/${tag}_bench.json; echo
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
