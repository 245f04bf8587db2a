#!/bin/bash
# one GPU call: the whole GPU test suite, then the bench lines of configs 4 and 3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
nproc; free -g | head -2
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 600 gpurun_out/bench_c4.err
timeout 400 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.err
