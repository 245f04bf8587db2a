#!/bin/bash
# Run on the GPU box (under gpurun): plain run first, then the ncu launch list and one
# --set full capture per kernel (recipe: /opt/skills/guides/B200_PROFILING.md).
# Output: gpurun_out/{launches,launches_levelsync}.csv, {k1_walk_full,k1_levelsync_full,aux_full}.ncu-rep
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
export PGS_NO_GRAPH=1
# default sweep (fused column walk): 4 launches per sweep (K0, walk passes A, B, C)
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 20 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k1_walk_dense -s 9 -c 3 \
    -o gpurun_out/k1_walk_full $CMD > gpurun_out/ncu_full.log 2>&1
# level-synchronous sweep (PGS_FLAG_KEEP_INTERNAL): K0 + one k1_sweep_dense launch per level
$CMD --level-sync > gpurun_out/plain_ls.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 150 --csv \
    --log-file gpurun_out/launches_levelsync.csv $CMD --level-sync > gpurun_out/ncu_launches_ls.log 2>&1
$CMD --level-sync > gpurun_out/plain_ls2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k1_sweep_dense -s 87 -c 29 \
    -o gpurun_out/k1_levelsync_full $CMD --level-sync > gpurun_out/ncu_full_ls.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw metric table of the big reports, not the reports
python scripts/summarize_ncu.py gpurun_out/k1_levelsync_full.ncu-rep gpurun_out/k1_levelsync_full && rm -f gpurun_out/k1_levelsync_full.ncu-rep
AUX="python scripts/profile_aux.py"
$AUX > gpurun_out/aux_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k0_rootgen|k1_sweep_sparse|k2_mismatch|k3_genome_fasta|k3_maf" -c 40 \
    -o gpurun_out/aux_full $AUX > gpurun_out/ncu_aux.log 2>&1
python scripts/summarize_ncu.py gpurun_out/aux_full.ncu-rep gpurun_out/aux_full && rm -f gpurun_out/aux_full.ncu-rep
ls -la gpurun_out
