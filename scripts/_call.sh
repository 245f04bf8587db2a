mkdir -p gpurun_out
run() { env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-mismatch --e2e-steps 1 2>gpurun_out/x.err >gpurun_out/x.json
python -c "
import json; l=json.load(open('gpurun_out/x.json')); m=l['e2e_files']['modes']; print('$*', {k:round(v['seconds'],3) for k,v in m.items()}, 'e2e host', round(l['e2e_host_expand']['ms_per_step'],1))"; }
run PGS_HX_RING=1
run PGS_X=1
run PGS_X=2
run PGS_HX_TASK_LINES=256
run PGS_X=3
