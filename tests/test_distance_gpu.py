"""north_star check (3): the observed distance matrix of kernel K2 against the reference's own
distance.mat (golden fixture of the unmodified reference at BASELINE config 2) and against the
reference's own sequences at matched parameters."""
import json
from pathlib import Path

import numpy as np
import pytest
from scipy.stats import norm

from pangenomesim_b200 import Engine

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def test_observed_distances_match_reference_distance_mat(oracle):
    fx = json.loads((GOLD / "config2.json").read_text())
    n, G, L, seed = 100, 1000, 1000, 1
    lines = fx["files"]["distance.mat"].splitlines()
    assert int(lines[0]) == n
    dmat = np.array([[float(x) for x in l.split()[1:]] for l in lines[1:]])
    # the tree the reference used (bit-exact restatement, pinned in test_ref_model.py)
    parent, left, right, time = oracle.RefEngine(seed).coalescent(n)
    rate = time * float(np.float32(0.01))
    leaf = np.where(np.arange(2 * n - 1) < n, np.arange(2 * n - 1), -1)
    with Engine(seed, L) as e:
        e.set_tree(parent, rate, leaf)
        e.set_core_genes(G, 2 * n - 2)
        e.sweep()
        mm, sites = e.mismatch()
    S = G * L
    assert sites == S and np.array_equal(mm, mm.T) and not mm.diagonal().any()
    pairs = n * (n - 1) // 2
    z = norm.ppf(1 - 0.01 / (2 * pairs))  # 99 % overall, Bonferroni
    iu = np.triu_indices(n, 1)
    d = dmat[iu]
    p = 0.75 * (1 - np.exp(-4.0 / 3.0 * d))          # Jukes-Cantor expectation of distance.mat
    obs = mm[iu] / S
    band = z * np.sqrt(p * (1 - p) / S) + 5e-4 * d + 1e-9   # + 4-significant-digit rounding of the file
    worst = np.max(np.abs(obs - p) / band)
    assert worst <= 1.0, f"worst pair at {worst:.2f} of its 99% band"
    # and the JC-corrected observed distance reproduces distance.mat
    dhat = -0.75 * np.log(1 - 4.0 / 3.0 * obs)
    assert np.allclose(dhat, d, rtol=0.05, atol=3e-4)


def test_head_to_head_with_reference_sequences_at_matched_parameters(oracle):
    """in the regime rate*L >> 1 the reference's 'force ~rate*L sites' model and
    Jukes-Cantor agree to first order; the allowance below is the model gap
    2/3 * sum(rate_e^2) along the path plus both sampling bands."""
    fx = json.loads((GOLD / "ref_distances.json").read_text())
    n, L = fx["params"]["num-genomes"], fx["params"]["gene-length"]
    parent = np.array(fx["parent"], dtype=np.int32)
    rate = np.array(fx["time"]) * float(np.float32(fx["params"]["mut-rate"]))
    leaf = np.where(np.arange(2 * n - 1) < n, np.arange(2 * n - 1), -1)
    G = fx["params"]["core-size"]
    with Engine(fx["params"]["seed"], L) as e:
        e.set_tree(parent, rate, leaf)
        e.set_core_genes(G, 2 * n - 2)
        e.sweep()
        mm, sites = e.mismatch()
    assert sites == fx["sites"]

    def path(a, b):
        anc, v = [], a
        while v != -1:
            anc.append(v)
            v = int(parent[v])
        out, v = [], b
        while v not in anc:
            out.append(v)
            v = int(parent[v])
        return anc[:anc.index(v)] + out
    for i in range(n):
        for j in range(i + 1, n):
            edges = path(i, j)
            p_new, p_ref = mm[i, j] / sites, fx["p_distance"][i][j]
            gap = sum(rate[k] ** 2 for k in edges) * 2.0 / 3.0 + len(edges) * 0.63 / L
            band = 2 * 2.576 * np.sqrt(max(p_ref, 1e-9) * (1 - p_ref) / sites)
            assert abs(p_new - p_ref) <= gap + band + 1e-9, (i, j, p_new, p_ref, gap, band)
