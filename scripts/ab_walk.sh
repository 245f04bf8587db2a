#!/bin/bash
# development: old (round-1) library against the current one on the same box; per-launch times under ncu
run() { # label, shard, env...
  local label=$1; shift; local shard=$1; shift
  env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --shard-of $shard > gpurun_out/t.json 2> gpurun_out/t.err
  python -c "
import json
l=json.load(open('gpurun_out/t.json')); print('$label', 'shard', $shard, round(l['ms_per_step'],4), 'launches', l['roofline']['launches_per_sweep']+1)"
}
for s in 1 2 4 8; do
  run old $s PGS_LIBRARY=build/oldlib/libpgs.so
  run new $s X=1
done
for u in 60 80 100 120 140 200; do run new_units$u 8 PGS_WALK_UNITS=$u; done
for u in 30 50 60; do run new_units$u 1 PGS_WALK_UNITS=$u; done
for s in 1 8; do
PGS_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/l_new_$s.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --shard-of $s > /dev/null 2>&1
done
