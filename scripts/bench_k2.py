"""K2 both ways on one GPU: popcount kernel and tensor-core kernels, same engine, same rows.
usage: bench_k2.py [n] [G] [L] [impl,impl,...]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

from pangenomesim_b200 import Engine  # noqa: E402
from pangenomesim_b200.capi import host_coalescent  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
G = int(sys.argv[2]) if len(sys.argv) > 2 else 96
L = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
impls = sys.argv[4].split(",") if len(sys.argv) > 4 else ["popc", "tensor", "tensor-i8", "tensor-fp4", "tensor-1cta"]
parent, rate, leaf = host_coalescent(n, 1, 0.05)
e = Engine(1, L)
e.set_tree(parent, rate, leaf)
e.set_core_genes(G, 2 * n - 2)
e.sweep()
ref = None
for impl in impls:
    os.environ["PGS_K2_IMPL"] = impl
    best = 1e30
    for _ in range(3):
        _, _, sites = e.mismatch_tiles_device()
        best = min(best, e.stats()["mismatch_ms"])
    line = f"K2 {impl:10s} n={n} sites={sites}: {best:.3f} ms = {n * (n - 1) / 2 * sites / best / 1e-3:.3e} site-pairs/s"
    if n <= 12000:
        mm, _ = e.mismatch()
        if ref is None:
            ref = mm
        line += "  equal to first: " + str(bool(np.array_equal(mm, ref)))
    print(line, flush=True)
