#!/bin/bash
# 2-GPU validation: the torchrun bench (strong scaling, K2 all-reduce inside the library, multi-device CLI leg) and the GPU tests that need two devices
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"; tail -c 300 gpurun_out/bench_n2.err
python -c "
import json; l=json.load(open('gpurun_out/bench_n2.json')); print(round(l['ms_per_step'],4), 'e2e', round(l['e2e']['ms_per_step'],1), 'k2', round(l['mismatch']['k2_ms'],1), round(l['mismatch']['allreduce']['ms'],2), 'cli', l['cli_multi_device']['ok'], l['mismatch_allreduce_check'])"
timeout 600 python -m pytest tests/test_host_cli.py tests/test_sweep_gpu.py -m gpu -q -k "check_distances or sharded or stale_origin or two_engines" 2>&1 | tail -3
