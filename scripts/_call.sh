mkdir -p gpurun_out
nvidia-smi --query-gpu=power.limit,power.draw,clocks.sm,temperature.gpu --format=csv,noheader
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/bench_c4.json'))
print('value %.4g ms %.4f frac %.3f e2e %.4g (%.1f ms) k2 %.1f' % (l['value'], l['ms_per_step'], l['roofline']['frac'], l['e2e']['value'], l['e2e']['ms_per_step'], l['mismatch']['k2_ms']), l['clocks'])
PY
