"""Development probe (not the bench): time pgs_sweep on a synthetic core-only workload."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import treegen  # noqa: E402
from pangenomesim_b200 import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--genes", type=int, default=5000)
ap.add_argument("--length", type=int, default=1000)
ap.add_argument("--mut-rate", type=float, default=0.01)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--mismatch", action="store_true")
a = ap.parse_args()

t0 = time.time()
parent, rate, leaf = treegen.coalescent(a.n, seed=1, mut_rate=a.mut_rate)
print(f"tree {time.time()-t0:.2f}s", flush=True)
e = Engine(1, a.length, flags=a.flags)
e.set_tree(parent, rate, leaf)
t0 = time.time()
e.set_core_genes(a.genes, len(parent) - 1)
print(f"plan+alloc {time.time()-t0:.2f}s", flush=True)
for r in range(a.reps):
    t0 = time.time()
    e.sweep()
    e.sync()
    wall = time.time() - t0
    s = e.stats()
    bytes_alg = (s["rows_written"] + s["rows_read"]) * a.length / 4
    bytes_real = (s["rows_written"] + s["rows_read"]) * s["blocks_per_gene"] * 16
    print(f"rep {r}: sweep {s['sweep_ms']:.3f} ms rootgen {s['rootgen_ms']:.3f} ms wall {wall*1e3:.2f} ms | "
          f"{bytes_alg/s['sweep_ms']/1e6:.0f} GB/s algorithmic ({bytes_real/s['sweep_ms']/1e6:.0f} moved) | "
          f"{s['leaf_nucleotides']/s['sweep_ms']/1e-3:.3e} leaf-nt/s | launches {s['kernel_launches']} levels {s['levels']} "
          f"| HBM {s['device_bytes']/2**30:.1f} GiB | mode {s['sweep_mode']} stored {s['rows_stored']/s['n_dense_genes']:.0f} loaded {s['rows_loaded']/s['n_dense_genes']:.0f} node-rows", flush=True)
if a.mismatch:
    t0 = time.time()
    mm, sites = e.mismatch()
    s = e.stats()
    pairs = a.n * (a.n - 1) / 2
    print(f"mismatch {s['mismatch_ms']:.2f} ms, {pairs*sites/s['mismatch_ms']/1e-3:.3e} site-pairs/s, wall {time.time()-t0:.2f}s")
