python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -3 gpurun_out/gputests.log
q() { python scripts/quick_sweep.py --reps 6 "$@" 2>&1 | tail -1 | cut -c1-40; }
for lib in pangenomesim_b200/lib/libpgs.so build/alt/libpgs_prev.so; do
 export PGS_LIBRARY=$lib; echo "#### $lib"
 for r in 0.01 0.1 0.5; do echo "== cfg4 mut $r: $(q --mut-rate $r)"; done
 echo "== cfg5 mut .5: $(q --n 50000 --genes 2000 --length 2000 --mut-rate 0.5)"
 echo "== cfg5 mut .01: $(q --n 50000 --genes 2000 --length 2000)"
done
