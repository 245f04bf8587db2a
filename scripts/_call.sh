mkdir -p gpurun_out
run() { env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-mismatch --no-files --e2e-steps 3 2>gpurun_out/x.err >gpurun_out/x.json
python -c "
import json; l=json.load(open('gpurun_out/x.json')); print('$*', 'e2e host', round(l['e2e_host_expand']['ms_per_step'],1), 'packed', round(l['e2e_packed']['ms_per_step'],1))"; }
run PGS_X=1
run PGS_SPIN_SYNC=1
run PGS_X=1
run PGS_SPIN_SYNC=1
run PGS_X=1 PGS_EXPAND_THREADS=15
run PGS_X=1 PGS_EXPAND_THREADS=17
