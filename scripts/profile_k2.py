"""Profiling target for K2 (k2_split_planes + k2_mismatch_tiles): 4096 genomes x 96 dense genes."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pangenomesim_b200 import Engine  # noqa: E402
from pangenomesim_b200.capi import host_coalescent  # noqa: E402

n, G, L = 4096, 96, 1000
parent, rate, leaf = host_coalescent(n, 1, 0.05)
e = Engine(1, L)
e.set_tree(parent, rate, leaf)
e.set_core_genes(G, 2 * n - 2)
e.sweep()
for _ in range(3):
    _, _, sites = e.mismatch_tiles_device()
ms = e.stats()["mismatch_ms"]
print(f"K2 {ms:.3f} ms = {n * (n - 1) / 2 * sites / ms / 1e-3:.3e} site-pairs/s")
