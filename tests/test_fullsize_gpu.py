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
