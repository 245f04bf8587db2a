mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
SHARDS="8 4 1" bash scripts/ab_shard.sh "registers PGS_NO_ASYNC_PROLOGUE=1" "staged X=1"
export PGS_NO_GRAPH=1
PGS_WALK_TRACE=gpurun_out/trace8_async.txt python bench.py --shard-of 8 --steps 2 --warmup 3 --no-e2e --no-cpu --no-mismatch > /dev/null 2>gpurun_out/tr.err
python scripts/walk_trace.py gpurun_out/trace8_async.txt | head -6
