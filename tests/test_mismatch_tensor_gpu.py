"""K2 on the tensor cores (mismatch_tc.cu: +-1 contraction, tcgen05 / TMEM / TMA) against the
popcount kernel and against numpy on the rows.  Integer counts: bit-exact."""
import numpy as np
import pytest

import treegen
from pangenomesim_b200 import Engine
from pangenomesim_b200.capi import PGS_FLAG_K2_POPC, PGS_FLAG_K2_TENSOR

pytestmark = pytest.mark.gpu


def core_table(n_core, root):
    ids = np.arange(n_core, dtype=np.int64)
    return ids, np.full(n_core, root, np.int32), np.arange(n_core + 1, dtype=np.int64), np.full(n_core, -1, np.int32)


def engine(n, L, G, seed, flags, mut_rate=0.5):
    parent, rate, leaf = treegen.coalescent(n, seed=n, mut_rate=mut_rate)
    e = Engine(seed, L, flags=flags)
    e.set_tree(parent, rate, leaf)
    e.set_genes(*core_table(G, len(parent) - 1))
    e.sweep()
    e.sync()
    return e


def numpy_counts(oracle, e, n, L, G):
    codes = np.stack([oracle.unpack_codes(e.copy_dense(v, G), L).reshape(-1) for v in range(n)])
    want = np.zeros((n, n), np.uint32)
    for i in range(n):
        want[i] = (codes != codes[i]).sum(-1)
    return want


# (n, L, G, operand budget in MB): one tile with the K range split over many CTAs; tile edges on
# both sides; an odd number of 64-site blocks (zero tail); a single block; several column chunks
# with a short last one
CASES = [(100, 1000, 7, None), (300, 130, 33, None), (129, 64, 1, None), (2, 100, 5, None), (333, 448, 21, "1"), (700, 1000, 3, "2"),
         (257, 1500, 2, None), (1100, 200, 5, None), (500, 64, 7, None), (1000, 192, 9, "1")]


@pytest.mark.parametrize("impl", ["tensor", "tensor-i8", "tensor-fp4", "tensor-1cta", "tensor-1cta-i8"])
@pytest.mark.parametrize("n,L,G,budget_mb", CASES)
def test_tensor_counts_equal_numpy_and_popc(oracle, monkeypatch, impl, n, L, G, budget_mb):
    if budget_mb:
        monkeypatch.setenv("PGS_K2_PLANES_MB", budget_mb)
    with engine(n, L, G, 90 + n, 0) as e:
        monkeypatch.setenv("PGS_K2_IMPL", "popc")
        popc, sites = e.mismatch()
        monkeypatch.setenv("PGS_K2_IMPL", impl)
        tens, sites_t = e.mismatch()
        assert sites == sites_t == G * L
        assert np.array_equal(tens, popc)
        assert np.array_equal(tens, numpy_counts(oracle, e, n, L, G))
        again, _ = e.mismatch()  # buffers are reused: a second call must not accumulate
        assert np.array_equal(again, popc)


def test_flag_selects_the_kernel(oracle, monkeypatch):
    monkeypatch.delenv("PGS_K2_IMPL", raising=False)
    n, L, G = 150, 300, 6
    with engine(n, L, G, 5, PGS_FLAG_K2_POPC) as a, engine(n, L, G, 5, PGS_FLAG_K2_TENSOR) as b:
        ma, _ = a.mismatch()
        mb, _ = b.mismatch()
        assert np.array_equal(ma, mb) and ma.any()


def test_tensor_saturated_and_identical_rows(oracle, monkeypatch):
    """mutation rate 0 (all genomes identical: every count 0) and saturation (p -> 3/4)"""
    monkeypatch.setenv("PGS_K2_IMPL", "tensor")
    n, L, G = 140, 1000, 4
    with engine(n, L, G, 7, 0, mut_rate=0.0) as e:
        mm, _ = e.mismatch()
        assert not mm.any()
    with engine(n, L, G, 7, 0, mut_rate=200.0) as e:
        mm, sites = e.mismatch()
        off = mm[~np.eye(n, dtype=bool)]
        assert abs(off.mean() / sites - 0.75) < 0.01
        assert np.array_equal(mm, numpy_counts(oracle, e, n, L, G))
