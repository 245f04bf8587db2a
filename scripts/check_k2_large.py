import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from pangenomesim_b200 import Engine
from pangenomesim_b200.capi import host_coalescent
from oracle import oracle_py
n, G, L, seed = 50000, 20, 2000, 5
parent, rate, leaf = host_coalescent(n, seed, 0.5)
e = Engine(seed, L); e.set_tree(parent, rate, leaf); e.set_core_genes(G, 2*n-2); e.sweep()
t0 = time.time(); mm, sites = e.mismatch(); print("K2 n=50000:", e.stats()['mismatch_ms'], "ms; wall", time.time()-t0, "sites", sites, "matrix GB", mm.nbytes/1e9)
assert (mm == mm.T).all() and not mm.diagonal().any()
rng = np.random.default_rng(0)
for _ in range(6):
    i, j = int(rng.integers(n)), int(rng.integers(n))
    a, b = e.copy_dense(i, G), e.copy_dense(j, G)
    assert mm[i, j] == oracle_py.mismatch(a.ravel(), b.ravel()), (i, j)
print("spot checks ok; max p-distance", mm.max()/sites)
