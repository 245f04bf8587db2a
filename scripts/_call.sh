mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "tests: $(( $(date +%s) - S )) s"; S=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
timeout 400 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
echo "bench: $(( $(date +%s) - S )) s"
python - <<'PY'
import json
for c in (4,3,5):
    l=json.load(open(f'gpurun_out/bench_c{c}.json'))
    h=l['e2e_host_expand']; d=l['e2e_device_formatter']
    print(c, 'sweep ms', round(l['ms_per_step'],4), 'frac', round(l['roofline']['frac'],3), 'e2e host', round(h['ms_per_step'],1), 'plan', round(h['plan_ms_per_step'],1), 'dev', round(d['ms_per_step'],1), 'plan', round(d['plan_ms_per_step'],1), 'packed', round(l['e2e_packed']['ms_per_step'],1), l['clocks']['reasons'])
PY
