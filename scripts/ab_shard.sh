#!/bin/bash
# development: A/B of sweep variants on one rank's shard of an N-way split of config 4 (one GPU)
#   usage: ab_shard.sh "<label> ENV=.. ENV=.." ...   (each argument = one variant; X=1 for none)
mkdir -p gpurun_out
for s in ${SHARDS:-8 4 1}; do
  for v in "$@"; do
    set -- "$@"; label=${v%% *}; envs=${v#* }
    best=""
    for rep in 1 2 3; do
      env $envs timeout 200 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu --no-mismatch --shard-of $s > gpurun_out/t.json 2> gpurun_out/t.err
      ms=$(python -c "import json; print(round(json.load(open('gpurun_out/t.json'))['ms_per_step'],4))")
      best="$best $ms"
    done
    echo "shard-of $s  $label: $best"
  done
done
