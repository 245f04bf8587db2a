mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "tests: $(( $(date +%s) - S )) s"; S=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "smoke: $(( $(date +%s) - S )) s"; S=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_default.err
echo "bench: $(( $(date +%s) - S )) s"; S=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
echo "ref: $(( $(date +%s) - S )) s"
