mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
sed -i 's/for rep in 1 2 3/for rep in 1 2/' scripts/ab_shard.sh
SHARDS="8 4 2 1" bash scripts/ab_shard.sh "one PGS_TOP_TASKS_PER_CTA=1" "folded X=1" "two PGS_TOP_TASKS_PER_CTA=2" "four PGS_TOP_TASKS_PER_CTA=4"
export PGS_NO_GRAPH=1
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-mismatch --shard-of 8"
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 4 --csv --log-file gpurun_out/r02_launches_shard8.csv $CMD > gpurun_out/ncu_l8.log 2>&1
grep -v "^==" gpurun_out/r02_launches_shard8.csv | awk -F'","' '{print $5, $9, $NF}' | tail -4
