#!/bin/bash
# one GPU call: K2 variants, the GPU test suite, profiles, bench lines
mkdir -p gpurun_out
for v in 0 1 2; do echo "K2 variant $v: $(PGS_K2_VARIANT=$v python scripts/profile_k2.py)"; done
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
PGS_PLAN_TIMING=1 timeout 300 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>&1 | grep -E "^plan" | tail -8
bash scripts/profile_gpu.sh > gpurun_out/profile.log 2>&1; tail -5 gpurun_out/profile.log
unset PGS_NO_GRAPH
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 400 gpurun_out/bench_c4.err
timeout 400 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.err
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 300 gpurun_out/bench_c5.err
for s in 2 4 8; do timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-mismatch --shard-of $s > gpurun_out/shard$s.json 2>/dev/null; python -c "
import json
l=json.load(open('gpurun_out/shard$s.json')); print('shard-of', $s, round(l['ms_per_step'],4))"; done
