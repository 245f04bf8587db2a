#!/bin/bash
# one GPU call: the whole GPU test suite, then the bench lines of configs 4 and 3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
nproc; free -g | head -2
python scripts/profile_k2.py > gpurun_out/k2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_mismatch_tiles -s 1 -c 1 -o gpurun_out/k2_full python scripts/profile_k2.py > gpurun_out/ncu_k2.log 2>&1; cat gpurun_out/k2_plain.log
timeout 400 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.err
