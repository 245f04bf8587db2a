set -u
mkdir -p gpurun_out
export PGS_NO_GRAPH=1
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-mismatch"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 12 --csv \
    --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k1_walk_dense|k1_top_masks" -s 6 -c 4 \
    -o gpurun_out/r02_k1_full $CMD > gpurun_out/ncu_full.log 2>&1
python scripts/summarize_ncu.py gpurun_out/r02_k1_full.ncu-rep gpurun_out/r02_k1_full
tail -5 gpurun_out/r02_launches.csv | cut -c1-200
rm -f gpurun_out/r02_k1_full.ncu-rep
