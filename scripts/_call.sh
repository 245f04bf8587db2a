mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo rc=$?; tail -c 300 gpurun_out/bench_c4.err
timeout 400 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo rc=$?; tail -c 300 gpurun_out/bench_c3.err
timeout 900 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo rc=$?; tail -c 300 gpurun_out/bench_c5.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
python -c "
import __graft_entry__ as g
g.smoke(); print('smoke ok')"
