mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo rc=$?; tail -3 gpurun_out/pytest_full.log
g++ -O2 -std=c++17 -pthread -Ipangenomesim_b200/csrc scripts/expand_bench.cpp pangenomesim_b200/csrc/expand.cpp -o /tmp/expand_bench && /tmp/expand_bench 16 | head -5 && PGS_EXPAND_ISA=avx2 /tmp/expand_bench 16 | head -5
for isa in avx512 avx2; do PGS_EXPAND_ISA=$isa timeout 300 python bench.py --steps 5 --warmup 3 --no-mismatch --no-files --no-cpu > gpurun_out/t.json 2>/dev/null; python -c "
import json
l=json.load(open('gpurun_out/t.json')); d=l['e2e_host_expand']; print('isa $isa', d['ms_per_step'], d['text_gb_per_s'], d['d2h_bytes_per_step'])"; done
