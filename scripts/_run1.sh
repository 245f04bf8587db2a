for mb in 64 256 512; do echo "== stage $mb"; PGS_STAGE_MB=$mb python bench.py --steps 5 --no-cpu --no-files 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print({k:e[k] for k in ('ms_per_step','k3_device_ms_per_step','d2h_gb_per_s','plan_ms_per_step')})"; done
