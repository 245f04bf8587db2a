mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_format_gpu.py -m gpu -x -q 2>&1 | tail -3
export PGS_OUT_TIMING=1
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-mismatch --no-files --e2e-steps 2 2>gpurun_out/x.err >gpurun_out/x.json; grep -E "^out" gpurun_out/x.err | grep -E "genome files|alignment.maf" | grep -v " 0.00 ms" | sed -n 3,6p
python -c "
import json; l=json.load(open('gpurun_out/x.json')); print('e2e host', round(l['e2e_host_expand']['ms_per_step'],1), 'packed', round(l['e2e_packed']['ms_per_step'],1))"; }
run PGS_X=1
run PGS_HX_ONE_COPY=1
run PGS_HX_ONE_COPY=1 PGS_HX_RING=2
run PGS_HX_ONE_COPY=1 PGS_HX_RING=4
run PGS_HX_ONE_COPY=1 PGS_STAGE_MB=512 PGS_HX_RING=2
run PGS_HX_ONE_COPY=1 PGS_STAGE_MB=1024 PGS_HX_RING=4
