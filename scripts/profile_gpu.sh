#!/bin/bash
# Run on the GPU box (under gpurun): plain run first, then the ncu launch list and one
# --set full capture per kernel (recipe: /opt/skills/guides/B200_PROFILING.md).
# Output (round 2): gpurun_out/r02_launches{,_c3}.csv, r02_{k1,k1sparse,k2}_full.{ncu-rep,csv,md}
set -u
mkdir -p gpurun_out
export PGS_NO_GRAPH=1
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-mismatch"
# config 4: 2 launches per sweep (k1_top_masks, k1_walk_dense)
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 12 --csv \
    --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k1_walk_dense|k1_top_masks" -s 6 -c 4 \
    -o gpurun_out/r02_k1_full $CMD > gpurun_out/ncu_full.log 2>&1
python scripts/summarize_ncu.py gpurun_out/r02_k1_full.ncu-rep gpurun_out/r02_k1_full
# config 3: 3 launches per sweep (k1_top_masks, k1_walk_dense, k1_walk_sparse)
CMD3="python bench.py --config 3 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD3 > gpurun_out/plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 12 --csv \
    --log-file gpurun_out/r02_launches_c3.csv $CMD3 > gpurun_out/ncu_launches3.log 2>&1
$CMD3 > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k1_walk_sparse" -s 2 -c 2 \
    -o gpurun_out/r02_k1sparse_full $CMD3 > gpurun_out/ncu_full3.log 2>&1
python scripts/summarize_ncu.py gpurun_out/r02_k1sparse_full.ncu-rep gpurun_out/r02_k1sparse_full
# K2
K2="python scripts/profile_k2.py"
$K2 > gpurun_out/k2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k2_mismatch_tiles|k2_split_planes" -s 2 -c 2 \
    -o gpurun_out/r02_k2_full $K2 > gpurun_out/ncu_k2.log 2>&1
python scripts/summarize_ncu.py gpurun_out/r02_k2_full.ncu-rep gpurun_out/r02_k2_full
cat gpurun_out/k2_plain.log
ls -la gpurun_out | tail -20
