"""Parity at BASELINE.json's full sizes (config 4: 10000 genomes x 5000 genes x 1000 sites, 25.6 GB of
rows; config 5 shape at reduced gene count) through size-independent properties: bit-exact oracle
spot checks along whole root-to-leaf paths, determinism, shard invariance, per-branch substitution
counts within their binomial band, uniform base composition, mismatch symmetry."""
import numpy as np
import pytest
from scipy.stats import norm

from pangenomesim_b200 import Engine
from pangenomesim_b200.capi import host_coalescent

pytestmark = pytest.mark.gpu


def popcount_mismatch(oracle, a, b, L):
    return int((oracle.unpack_codes(a, L) != oracle.unpack_codes(b, L)).sum())


@pytest.fixture(scope="module")
def config4():
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs 26 GB of device memory")
    n, G, L, seed = 10000, 5000, 1000, 1
    parent, rate, leaf = host_coalescent(n, seed, 0.01)
    e = Engine(seed, L)
    e.set_tree(parent, rate, leaf)
    e.set_core_genes(G, 2 * n - 2)
    e.sweep()
    e.sync()
    yield e, parent, rate, leaf, n, G, L, seed
    e.close()


def test_config4_counts_and_stats(config4):
    e, parent, rate, leaf, n, G, L, seed = config4
    s = e.stats()
    assert s["leaf_nucleotides"] == n * G * L == 5 * 10**10
    assert s["site_edges"] == (2 * n - 2) * G * L
    assert s["rows_read"] == (n - 1) * G and s["rows_written"] == (2 * n - 2) * G   # algorithmic (level-synchronous) counts
    # the fused walk keeps rows for the genomes and the root plus a small pool of transient rows
    assert s["sweep_mode"] == 1 and (n + 1) * G <= s["rows_total"] < 1.3 * n * G
    assert s["rows_stored"] < 1.4 * n * G and s["rows_loaded"] < 0.4 * n * G
    assert 12.8e9 < s["device_bytes"] < 17e9


def test_config4_oracle_spot_checks(config4, oracle):
    e, parent, rate, leaf, n, G, L, seed = config4
    rng = np.random.default_rng(0)
    root = 2 * n - 2
    nodes = [0, n - 1, root] + [int(x) for x in rng.integers(0, n, 25)]  # leaves and root: what the walk keeps
    genes = [0, G - 1] + [int(x) for x in rng.integers(0, G, 6)]
    for v in nodes:
        for g in genes:
            got = e.copy_packed(v, g)
            assert np.array_equal(got, oracle.row(seed, parent, rate, v, root, g, L)), (v, g)


def test_config4_deterministic_and_shard_invariant(config4):
    e, parent, rate, leaf, n, G, L, seed = config4
    probe = [0, 1234, n - 1, 4321, 2 * n - 2]
    before = {v: e.copy_dense(v, G).copy() for v in probe}
    e.sweep()
    e.sync()
    for v in probe:
        assert np.array_equal(before[v], e.copy_dense(v, G))
    # a second engine holding only genes [2500, 2600) reproduces exactly those rows
    with Engine(seed, L) as part:
        part.set_tree(parent, rate, leaf)
        part.set_core_genes(100, 2 * n - 2, first_id=2500)
        part.sweep()
        for v in probe:
            assert np.array_equal(before[v][2500:2600], part.copy_dense(v, 100))


def test_config4_branch_substitution_counts(config4, oracle):
    """per-branch counts ~ Binomial(S, p_e): the longest branches, a few typical and tiny ones"""
    _, parent, rate, leaf, n, _, L, seed = config4
    from pangenomesim_b200.capi import PGS_FLAG_KEEP_INTERNAL
    G = 500  # internal rows are needed here: a second, smaller engine in level-synchronous mode
    e = Engine(seed, L, flags=PGS_FLAG_KEEP_INTERNAL)
    e.set_tree(parent, rate, leaf)
    e.set_core_genes(G, 2 * n - 2)
    e.sweep()
    S = G * L
    order = np.argsort(rate[:-1])
    picks = list(order[-6:]) + list(order[len(order) // 2 - 2: len(order) // 2 + 2]) + list(order[:3])
    z = norm.ppf(1 - 0.01 / (2 * len(picks)))
    total_obs = total_exp = 0.0
    for v in picks:
        v = int(v)
        child, par = e.copy_dense(v, G), e.copy_dense(int(parent[v]), G)
        obs = popcount_mismatch(oracle, child, par, L)
        p = oracle.jc_p(rate[v])
        assert abs(obs - S * p) <= z * np.sqrt(S * p * (1 - p)) + 1e-9, (v, rate[v], obs, S * p)
        total_obs += obs
        total_exp += S * p
    assert total_obs > 0
    # padding sites (1000..1023 of every gene) are zero everywhere
    codes = oracle.unpack_codes(e.copy_dense(int(order[-1]), G)[:50], 1024)
    assert not codes[:, 1000:].any()
    # and this engine's leaves equal the big walk-mode engine's (same gene ids 0..499)
    big = config4[0]
    for v in (0, 77, n - 1):
        assert np.array_equal(e.copy_dense(v, G), big.copy_dense(v, 5000)[:G])
    e.close()


def test_config4_leaf_composition_and_distance(config4, oracle):
    e, parent, rate, leaf, n, G, L, seed = config4
    S = G * L
    a, b = e.copy_dense(0, G), e.copy_dense(1, G)
    comp = np.bincount(oracle.unpack_codes(a, L).ravel(), minlength=4) / S
    assert np.all(np.abs(comp - 0.25) < 5 * np.sqrt(0.1875 / S))
    # observed distance of genomes 0 and 1 against the JC expectation of the tree path
    da, v, acc = {}, 0, 0.0
    while v != -1:
        da[v] = acc
        acc += rate[v]
        v = int(parent[v])
    v, accb = 1, 0.0
    while v not in da:
        accb += rate[v]
        v = int(parent[v])
    d = da[v] + accb
    p = 0.75 * (1 - np.exp(-4 * d / 3))
    obs = popcount_mismatch(oracle, a, b, L) / S
    assert abs(obs - p) < 4.5 * np.sqrt(p * (1 - p) / S) + 1e-9


def test_config5_shape_high_mutation_rate(oracle):
    """config 5's tree (50000 genomes, gene-length 2000, high mut-rate) at a reduced gene count:
    memory layout beyond 2^31 bytes, 38+ levels, dense rare path; oracle spot checks."""
    n, G, L, seed = 50000, 120, 2000, 3
    parent, rate, leaf = host_coalescent(n, seed, 0.5)
    root = 2 * n - 2
    with Engine(seed, L) as e:
        e.set_tree(parent, rate, leaf)
        e.set_core_genes(G, root)
        e.sweep()
        s = e.stats()
        assert s["rows_total"] * s["blocks_per_gene"] * 16 > 2**31
        rng = np.random.default_rng(1)
        for v in [0, n - 1, root] + [int(x) for x in rng.integers(0, n, 12)]:
            g = int(rng.integers(0, G))
            assert np.array_equal(e.copy_packed(v, g), oracle.row(seed, parent, rate, v, root, g, L)), (v, g)


# ---- BASELINE config 3: num-genomes=1000, core-size=3000, gene-length=1500, theta/rho raised for a
# large accessory genome (SURVEY.md §8d suggests theta=2000, rho=1): ~29 600 accessory genes, each
# present in its own small subtree -- the workload of gene::bulk_mutate (gene.h:26-40) and of the
# accessory traversal (img.cxx:278-333).  Gene table from the retained host model.

@pytest.fixture(scope="module")
def config3():
    from pangenomesim_b200.capi import HostModel
    m = HostModel("fast", num_genomes=1000, core_size=3000, gene_length=1500, theta=2000, rho=1, seed=1)
    e = Engine(1, 1500)
    e.set_tree(m.parent, m.rate, m.leaf_genome)
    e.set_genes(*m.gene_table())
    e.sweep()
    e.sync()
    yield e, m
    e.close()
    m.close()


def test_config3_shape_and_launches(config3):
    e, m = config3
    s = e.stats()
    assert m.n_genomes == 1000 and m.n_genes > 20000 and s["n_dense_genes"] >= 3000
    assert s["leaf_nucleotides"] == int(m.genome_gene_count.sum()) * 1500
    assert s["sweep_mode"] == 1
    # one launch for all sparse genes (origin rows included), two for the dense ones (masks, walk)
    assert s["kernel_launches"] == 3
    assert s["rows_stored"] < s["rows_written"] + s["rows_origin"]  # internal rows are not written


def test_config3_rows_match_oracle(config3, oracle):
    """dense and sparse rows of genomes and origins against the host replay, whole paths origin -> leaf"""
    e, m = config3
    rng = np.random.default_rng(3)
    L, seed = 1500, 1
    dense = np.where((m.carriers[m.carrier_off[:-1]] == -1) & (np.diff(m.carrier_off) == 1))[0]
    sparse = np.setdiff1d(np.arange(m.n_genes), dense)
    assert len(dense) >= 3000 and len(sparse) > 20000
    for gi in rng.choice(dense, 6, replace=False):
        for v in [m.root] + [int(x) for x in rng.integers(0, 1000, 4)]:
            want = oracle.row(seed, m.parent, m.rate, v, m.root, int(m.gene_id[gi]), L)
            assert np.array_equal(e.copy_packed(v, int(gi)), want), (gi, v)
    checked = 0
    n_car = np.diff(m.carrier_off)
    picks = list(rng.choice(sparse, 40, replace=False)) + [int(sparse[np.argmax(n_car[sparse])])]  # and the most widespread one
    for gi in picks:
        gi = int(gi)
        org, gid = int(m.origin[gi]), int(m.gene_id[gi])
        car = m.carriers[m.carrier_off[gi]:m.carrier_off[gi + 1]]
        assert np.array_equal(e.copy_packed(org, gi), oracle.row(seed, m.parent, m.rate, org, org, gid, L)), gi
        for g in list(car[:3]) + list(car[-2:]):
            v = int(g)  # genome g sits on leaf g in the reference's pool layout
            assert np.array_equal(e.copy_packed(v, gi), oracle.row(seed, m.parent, m.rate, v, org, gid, L)), (gi, v)
            checked += 1
        others = np.setdiff1d(rng.integers(0, 1000, 3), car)
        for g in others:
            assert e.copy_packed(int(g), gi) is None
    assert checked > 80


def test_config3_sparse_walk_equals_level_sweep(config3):
    """the fused sparse walk against the level-synchronous kernels (PGS_FLAG_KEEP_INTERNAL) on a
    second engine: same rows at the genomes, gene by gene"""
    from pangenomesim_b200.capi import PGS_FLAG_KEEP_INTERNAL
    e, m = config3
    rng = np.random.default_rng(5)
    with Engine(1, 1500, flags=PGS_FLAG_KEEP_INTERNAL) as k:
        k.set_tree(m.parent, m.rate, m.leaf_genome)
        k.set_genes(*m.gene_table())
        k.sweep()
        assert k.stats()["sweep_mode"] == 0 and k.stats()["kernel_launches"] > 20
        for g in rng.integers(0, 1000, 12):
            for gi in rng.integers(0, m.n_genes, 40):
                a, b = e.copy_packed(int(g), int(gi)), k.copy_packed(int(g), int(gi))
                assert (a is None) == (b is None) and (a is None or np.array_equal(a, b)), (g, gi)


def test_config3_fasta_and_maf_text(config3, oracle):
    """genome500.fasta and one accessory gene's MAF block out of the full-size text stream (15 GB),
    record by record against the host replay"""
    import ctypes
    from pangenomesim_b200.capi import OUT_GENOMES, OUT_MAF
    e, m = config3
    L, seed = 1500, 1
    n_car = np.diff(m.carrier_off)
    sparse = np.where(~((m.carriers[m.carrier_off[:-1]] == -1) & (n_car == 1)))[0]
    gm = int(sparse[np.argsort(n_car[sparse])[len(sparse) // 2]])  # an accessory gene of median spread
    marker = b"s genomeref.%d " % m.gene_id[gm]
    got = {"fasta": [], "maf": None}

    def sink(name, offset, data, length):
        if name == "genome500.fasta":
            got["fasta"].append(ctypes.string_at(data, length))
        elif name == "alignment.maf" and got["maf"] is None:
            chunk = ctypes.string_at(data, length)
            at = chunk.find(marker)
            if at >= 0:
                end = chunk.find(b"\n\n", at)
                got["maf"] = chunk[at - 2:end + 2]
    rep = e.write_outputs(None, OUT_GENOMES | OUT_MAF, m.ref_core, m.ref_acc, sink=sink)
    assert rep["files"] == 1001 and rep["bytes"] > 2 * int(m.genome_gene_count.sum()) * 1500
    text = b"".join(got["fasta"]).split(b"\n")
    g = 499
    carried = [k for k in range(m.n_genes)
               if (m.carriers[m.carrier_off[k]] == -1 and n_car[k] == 1) or g in m.carriers[m.carrier_off[k]:m.carrier_off[k + 1]]]
    assert len(text) == 2 * len(carried) + 1 and len(carried) == m.genome_gene_count[g]
    for r in [0, 1, len(carried) // 2, len(carried) - 1] + [i for i, k in enumerate(carried) if k in set(sparse[:2000])][:4]:
        k = carried[r]
        assert text[2 * r] == b">genome500.%d" % m.gene_id[k]
        want = oracle.row(seed, m.parent, m.rate, g, int(m.origin[k]), int(m.gene_id[k]), L)
        assert text[2 * r + 1] == oracle.unpack_ascii(want, L), (r, k)
    # the MAF block: a, reference line, one s line per carrier in the table's carrier order, blank line
    block = got["maf"].split(b"\n")
    car = m.carriers[m.carrier_off[gm]:m.carrier_off[gm + 1]]
    n_ref = len(m.ref_core) + len(m.ref_acc)
    org, gid = int(m.origin[gm]), int(m.gene_id[gm])
    assert block[0] == b"a" and len(block) == 2 + len(car) + 2
    ref_seq = oracle.unpack_ascii(oracle.row(seed, m.parent, m.rate, org, org, gid, L), L)
    assert block[1] == b"s genomeref.%d 0 %d + %d %s" % (gid, L, n_ref * L, ref_seq)
    for i, cg in enumerate(car):
        seq = oracle.unpack_ascii(oracle.row(seed, m.parent, m.rate, int(cg), org, gid, L), L)
        assert block[2 + i] == b"s genome%03d.%d 0 %d + %d %s" % (cg + 1, gid, L, m.genome_gene_count[cg] * L, seq), i


# ---- BASELINE config 5 at full size: num-genomes=50000, core-size=2000, gene-length=2000, high
# mut-rate (0.5): 48 GiB of rows on one GPU, most blocks of the long branches take the rare path

def test_config5_full_size(oracle):
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 80e9:
        pytest.skip("needs 55 GB of device memory")
    n, G, L, seed = 50000, 2000, 2000, 3
    parent, rate, leaf = host_coalescent(n, seed, 0.5)
    root = 2 * n - 2
    with Engine(seed, L) as e:
        e.set_tree(parent, rate, leaf)
        e.set_core_genes(G, root)
        e.sweep()
        s = e.stats()
        assert s["leaf_nucleotides"] == n * G * L == 2 * 10**11 and s["site_edges"] == 4 * 10**11 - 2 * G * L
        assert s["sweep_mode"] == 1 and s["kernel_launches"] == 2
        assert 50e9 < s["device_bytes"] < 60e9
        rng = np.random.default_rng(1)
        for v in [0, n - 1, root] + [int(x) for x in rng.integers(0, n, 20)]:
            for g in (0, G - 1, int(rng.integers(0, G))):
                assert np.array_equal(e.copy_packed(v, g), oracle.row(seed, parent, rate, v, root, g, L)), (v, g)
        before = e.copy_dense(12345, G).copy()
        e.sweep()  # repeatable
        e.sync()
        assert np.array_equal(before, e.copy_dense(12345, G))
        # saturating distances: two genomes far apart differ in close to 3/4 (1 - exp(-4d/3)) of the sites
        a, b = e.copy_dense(0, G), e.copy_dense(1, G)
        da, v, acc = {}, 0, 0.0
        while v != -1:
            da[v] = acc
            acc += rate[v]
            v = int(parent[v])
        v, accb = 1, 0.0
        while v not in da:
            accb += rate[v]
            v = int(parent[v])
        d = da[v] + accb
        p = 0.75 * (1 - np.exp(-4 * d / 3))
        S = G * L
        obs = popcount_mismatch(oracle, a, b, L) / S
        assert abs(obs - p) < 4.5 * np.sqrt(p * (1 - p) / S) + 1e-9, (obs, p)
