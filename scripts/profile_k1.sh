set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
export PGS_NO_GRAPH=1
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 150 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k1_sweep_dense -s 87 -c 29 -o gpurun_out/k1_full $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | grep -E "k1_full|launches"
