for j in 1 2 3 4; do python scripts/flake.py pangenomesim_b200/bin/pangenomesim-b200 30 fixed$j > gpurun_out/flake_fixed$j.log 2>&1 & done; wait
grep -h "bad runs" gpurun_out/flake_fixed*.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for s in 1 8; do timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-mismatch --shard-of $s > gpurun_out/shard$s.json 2>/dev/null; python -c "
import json
l=json.load(open('gpurun_out/shard$s.json')); print('shard-of', $s, round(l['ms_per_step'],4))"; done
