timeout 600 python -m pytest tests/test_sweep_gpu.py -m gpu -q -k "stale_origin" 2>&1 | tail -2
cp pangenomesim_b200/lib/libpgs.so /tmp/new.so; cp build/racy/libpgs.so pangenomesim_b200/lib/libpgs.so
echo "== the library with the hoistable loads (expected: the delayed variant fails)"
timeout 600 python -m pytest tests/test_sweep_gpu.py -m gpu -q -k "stale_origin" 2>&1 | grep -E "passed|failed|assert np.array_equal" | head -4
cp /tmp/new.so pangenomesim_b200/lib/libpgs.so
timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-mismatch --shard-of 8 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('shard-of 8', round(l['ms_per_step'],4))"
