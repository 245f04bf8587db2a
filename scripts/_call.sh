mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_format_gpu.py tests/test_host_cli.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -5
export PGS_OUT_TIMING=1
echo "== config 4"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-mismatch --no-files --e2e-steps 2 2>gpurun_out/c4.err >gpurun_out/c4.json; grep -E "^out" gpurun_out/c4.err | grep -v " 0.0[0-9] ms\| 0.1[0-9] ms" | tail -14
python -c "
import json; l=json.load(open('gpurun_out/c4.json')); print('e2e host', l['e2e_host_expand']['ms_per_step'], 'dev', l['e2e_device_formatter']['ms_per_step'], 'packed', l['e2e_packed']['ms_per_step'])"
for r in 1 2; do echo "== ring $r"; PGS_HX_RING=$r timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-mismatch --no-files --e2e-steps 2 2>&1 >/dev/null | grep -E "^out.*maf" | tail -2; done
echo "== task lines 256 / 4096"
for r in 256 4096; do PGS_HX_TASK_LINES=$r timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-mismatch --no-files --e2e-steps 2 2>&1 >/dev/null | grep -E "^out.*maf" | tail -2; done
