"""Profiling target for K0, the sparse K1 kernel, K2 (mismatch) and K3 (formatter): one mid-sized run, every stage once."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import treegen  # noqa: E402
from pangenomesim_b200 import Engine  # noqa: E402
from pangenomesim_b200.capi import OUT_DISCARD, OUT_GENOMES, OUT_MAF  # noqa: E402

n, G, L = 2000, 500, 1000
parent, rate, leaf = treegen.coalescent(n, seed=1, mut_rate=0.01)
e = Engine(1, L)
e.set_tree(parent, rate, leaf)
# 500 core genes plus 3000 accessory genes (random origin, random carriers below it) so that the
# sparse sweep kernel runs too
gene_id, origin, off, car = treegen.gene_table(parent, leaf, n_core=G, n_acc=3000, seed=3, shuffle=False)
e.set_genes(gene_id, origin, off, car)
for _ in range(2):
    e.sweep()
    mm, sites = e.mismatch()
    rep = e.write_outputs(None, OUT_GENOMES | OUT_MAF | OUT_DISCARD, list(range(len(gene_id))), [])
st = e.stats()
pairs = n * (n - 1) / 2
print(f"K2 {st['mismatch_ms']:.3f} ms = {pairs * sites / st['mismatch_ms'] / 1e-3:.3e} site-pairs/s; "
      f"K3 {rep['device_ms']:.3f} ms for {rep['bytes']} bytes = {rep['bytes'] * 1.25 / 1.03 / rep['device_ms'] / 1e6:.0f} GB/s; "
      f"wall {rep['wall_ms']:.1f} ms")
