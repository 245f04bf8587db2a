#!/bin/bash
# N-GPU bench line under torchrun (strong scaling of config 4, K2 all-reduce inside the library, multi-device CLI leg)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -c 800 gpurun_out/bench_n$N.err; head -c 400 gpurun_out/bench_n$N.json
