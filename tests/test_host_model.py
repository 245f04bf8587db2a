"""CPU tests of the retained host model (pangenomesim_b200/host/model.cpp) through its C API: the
tables it hands to pgs_set_tree / pgs_set_genes are consistent, the flat-array gain/loss process of
rng-fast runs is the same process as the reference-compatible one (same coalescent, same
distribution of the pangenome), and both are deterministic."""
import numpy as np
import pytest

from pangenomesim_b200.capi import HostModel, PgsError

PARAMS = dict(num_genomes=60, core_size=25, gene_length=30, theta=40, rho=2, mut_rate=0.05)


def ancestors(parent, v):
    out = {v}
    while parent[v] >= 0:
        v = int(parent[v])
        out.add(v)
    return out


@pytest.mark.parametrize("mode", ["fast", "compat"])
def test_tables_are_consistent(mode):
    m = HostModel(mode, seed=5, **PARAMS)
    n, N = m.n_genomes, m.n_nodes
    assert N == 2 * n - 1 and m.parent[m.root] == -1 and (m.parent[:-1] > np.arange(N - 1)).all()
    assert (m.leaf_genome[:n] == np.arange(n)).all() and (m.leaf_genome[n:] == -1).all()
    assert (m.rate[:-1] > 0).all() and m.rate[m.root] == 0
    assert sorted(m.all_order) == list(range(n))
    assert len(set(m.gene_id.tolist())) == m.n_genes and m.carrier_off[0] == 0 and m.carrier_off[-1] == len(m.carriers)
    count = np.zeros(n, np.int64)
    core = 0
    for k in range(m.n_genes):
        car = m.carriers[m.carrier_off[k]:m.carrier_off[k + 1]]
        assert len(car) >= 1
        if len(car) == 1 and car[0] == -1:  # carried by every genome, origin = root
            assert m.origin[k] == m.root
            count += 1
            core += 1
            continue
        assert len(set(car.tolist())) == len(car) and car.min() >= 0 and car.max() < n
        for g in car:  # the gene appeared on the branch into its origin: every carrier is below it
            assert int(m.origin[k]) in ancestors(m.parent, int(g))
        count[car] += 1
    assert core >= PARAMS["core_size"] and (count == m.genome_gene_count).all()
    # reference lists: every core gene, every accessory gene that some genome carries, nothing twice
    refs = list(m.ref_core) + list(m.ref_acc)
    assert sorted(refs) == list(range(m.n_genes))
    assert m.used_compat == (mode == "compat")
    d = m.distance_row(0)
    assert d[0] == 0 and (d[1:] > 0).all()
    m.close()


def test_fast_mode_keeps_the_coalescent_and_the_process():
    """same seed: the coalescent is drawn first and is identical in both modes; the pangenome differs
    (other draws, other orders) but follows the same law: E[accessory genes per genome] = theta / rho"""
    per_genome = {"fast": [], "compat": []}
    for seed in range(1, 13):
        a = HostModel("fast", seed=seed, **PARAMS)
        b = HostModel("compat", seed=seed, **PARAMS)
        assert (a.parent == b.parent).all() and np.array_equal(a.rate, b.rate)
        per_genome["fast"].append(a.genome_gene_count.mean() - PARAMS["core_size"])
        per_genome["compat"].append(b.genome_gene_count.mean() - PARAMS["core_size"])
        a.close()
        b.close()
    want = PARAMS["theta"] / PARAMS["rho"]  # 20 accessory genes per genome in equilibrium
    for mode, xs in per_genome.items():
        assert abs(np.mean(xs) - want) < 3.0, (mode, np.mean(xs))


def test_deterministic_and_parameter_errors():
    a = HostModel("fast", seed=9, **PARAMS)
    b = HostModel("fast", seed=9, **PARAMS)
    assert np.array_equal(a.gene_id, b.gene_id) and np.array_equal(a.carriers, b.carriers) and np.array_equal(a.origin, b.origin)
    # fast mode's own orders: genes by ascending id, carriers ascending
    assert (np.diff(a.gene_id) > 0).all()
    for k in range(a.n_genes):
        car = a.carriers[a.carrier_off[k]:a.carrier_off[k + 1]]
        assert (np.diff(car) > 0).all() or len(car) == 1
    a.close()
    b.close()
    with pytest.raises(PgsError):
        HostModel("fast", bogus_key=1)
    with pytest.raises(PgsError):
        HostModel("fast", num_genomes="x")
    with pytest.raises(PgsError):
        HostModel("fast", num_genomes=0)
