timeout 900 python -m pytest tests/test_sweep_gpu.py -m gpu -x -q -k pieces 2>&1 | tail -3
sed -i 's/for rep in 1 2 3/for rep in 1 2/' scripts/ab_shard.sh
SHARDS=8 bash scripts/ab_shard.sh "base X=1" \
  "u20c120 PGS_WALK_UNITS=20 PGS_WALK_CARVE=120" "u30c120 PGS_WALK_UNITS=30 PGS_WALK_CARVE=120" "u40c120 PGS_WALK_UNITS=40 PGS_WALK_CARVE=120" \
  "u60c120 PGS_WALK_UNITS=60 PGS_WALK_CARVE=120" "u80c120 PGS_WALK_UNITS=80 PGS_WALK_CARVE=120" \
  "u30c100 PGS_WALK_UNITS=30 PGS_WALK_CARVE=100" "u60c100 PGS_WALK_UNITS=60 PGS_WALK_CARVE=100" \
  "u30c140 PGS_WALK_UNITS=30 PGS_WALK_CARVE=140" "u60c140 PGS_WALK_UNITS=60 PGS_WALK_CARVE=140" "u80c160 PGS_WALK_UNITS=80 PGS_WALK_CARVE=160" "u120c160 PGS_WALK_CARVE=160"
export PGS_NO_GRAPH=1
PGS_WALK_UNITS=40 PGS_WALK_CARVE=120 PGS_WALK_TRACE=gpurun_out/trace8_carve.txt python bench.py --shard-of 8 --steps 2 --warmup 3 --no-e2e --no-cpu --no-mismatch > /dev/null 2>gpurun_out/tr.err
python scripts/walk_trace.py gpurun_out/trace8_carve.txt
