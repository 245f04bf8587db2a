#!/bin/bash
# 2-GPU validation: the torchrun bench (strong scaling, K2 all-reduce inside the library, multi-device CLI leg) and the GPU tests that need two devices
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_n2.err
timeout 600 python -m pytest tests/test_host_cli.py -m gpu -q -k "check_distances or sharded" 2>&1 | tail -4
