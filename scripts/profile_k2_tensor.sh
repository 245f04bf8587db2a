#!/bin/bash
# Run on the GPU box (under gpurun): K2 on the tensor cores, 10 000 genomes x 200 genes.
# plain run, ncu launch list, one --set full capture of k2_tc_pair and k2_expand_pm1.
set -u
mkdir -p gpurun_out
CMD="python scripts/bench_k2.py 10000 200 1000 tensor-fp4"
$CMD > gpurun_out/k2t_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k2_" -c 24 --csv \
    --log-file gpurun_out/r02_k2fp4_launches.csv $CMD > gpurun_out/k2t_ncu_launches.log 2>&1
$CMD > gpurun_out/k2t_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k2_tc_pair|k2_expand_pm1" -s 2 -c 2 \
    -o gpurun_out/r02_k2fp4_full $CMD > gpurun_out/k2t_ncu_full.log 2>&1
python scripts/summarize_ncu.py gpurun_out/r02_k2fp4_full.ncu-rep gpurun_out/r02_k2fp4_full
cat gpurun_out/k2t_plain.log
tail -3 gpurun_out/k2t_ncu_full.log
