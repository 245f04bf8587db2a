mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_format_gpu.py -m gpu -q -x > gpurun_out/pytest_fmt.log 2>&1; echo rc=$?; tail -5 gpurun_out/pytest_fmt.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo rc=$?; tail -3 gpurun_out/pytest_full.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-mismatch --no-files > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 300 gpurun_out/bench_c4.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/bench_c4.json'))
for k in ('e2e_host_expand','e2e_device_formatter','e2e_packed'):
    d=l[k]; print(k, d['ms_per_step'], d.get('text_gb_per_s'), d.get('d2h_gb_per_s'), d.get('d2h_bytes_per_step'))
PY
PGS_HX_TWO_PASS=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-mismatch --no-files --no-cpu > gpurun_out/t.json 2>/dev/null; python -c "
import json
l=json.load(open('gpurun_out/t.json')); d=l['e2e_host_expand']; print('two-pass', d['ms_per_step'], d['text_gb_per_s'], d['d2h_bytes_per_step'])"
for mb in 128 512; do PGS_STAGE_MB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --no-mismatch --no-files --no-cpu > gpurun_out/t.json 2>/dev/null; python -c "
import json
l=json.load(open('gpurun_out/t.json')); d=l['e2e_host_expand']; print('stage', $mb, d['ms_per_step'], d['text_gb_per_s'])"; done
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.err
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --no-mismatch > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 300 gpurun_out/bench_c5.err
python - <<'PY'
import json
for c in (3,5):
    l=json.load(open(f'gpurun_out/bench_c{c}.json'))
    for k in ('e2e_host_expand','e2e_device_formatter'):
        d=l[k]; print(c, k, d['ms_per_step'], d.get('text_gb_per_s'), d.get('d2h_gb_per_s'), d.get('plan_ms_per_step'))
PY
