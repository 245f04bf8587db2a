"""GPU parity of K0/K1 (origin rows + level sweep) against the host Philox replay in oracle/,
through the C ABI.  Bit-exact: this is integer/byte work."""
import numpy as np
import pytest

import treegen
from pangenomesim_b200 import Engine, PgsError
from pangenomesim_b200.capi import PGS_FLAG_KEEP_INTERNAL, PGS_FLAG_NO_GRAPH

pytestmark = pytest.mark.gpu


def build(seed, L, parent, rate, leaf_genome, genes, flags=PGS_FLAG_KEEP_INTERNAL):
    """default here: the level-synchronous sweep that keeps every internal row, so that every node
    can be compared; the fused column walk (the product default) is covered by the *_walk tests"""
    e = Engine(seed, L, flags=flags)
    e.set_tree(parent, rate, leaf_genome)
    e.set_genes(*genes)
    e.sweep()
    e.sync()
    return e


def core_table(n_core, root, first=0):
    ids = np.arange(first, first + n_core, dtype=np.int64)
    return ids, np.full(n_core, root, np.int32), np.arange(n_core + 1, dtype=np.int64), np.full(n_core, -1, np.int32)


def check_all_nodes(e, oracle, seed, parent, rate, gene_id, origin, L, gene_indices):
    for gi in gene_indices:
        want = oracle.gene_all_nodes(seed, parent, rate, int(origin[gi]), int(gene_id[gi]), L)
        for v in range(len(parent)):
            got = e.copy_packed(v, int(gi))
            if got is None:
                continue
            assert np.array_equal(got, want[v]), f"gene index {gi} node {v}"


@pytest.mark.parametrize("L", [1, 63, 64, 65, 100, 130, 1000, 1500])
def test_core_genes_every_node(oracle, L):
    n, seed = 9, 7 + L
    parent, rate, leaf = treegen.coalescent(n, seed=3, mut_rate=0.3)
    genes = core_table(5, len(parent) - 1)
    with build(seed, L, parent, rate, leaf, genes) as e:
        s = e.stats()
        assert s["n_dense_genes"] == 5 and s["rows_total"] == 5 * len(parent)
        assert s["rows_written"] == 5 * (len(parent) - 1) and s["rows_read"] == 5 * (n - 1)
        check_all_nodes(e, oracle, seed, parent, rate, genes[0], genes[1], L, range(5))


@pytest.mark.parametrize("mut_rate", [1e-6, 0.01, 2.0, 50.0])
def test_rate_regimes(oracle, mut_rate):
    """ultra-sparse (copy), default, dense (K ~ 30 per block) and saturated (p = 3/4) branches"""
    n, L, seed = 12, 200, 99
    parent, rate, leaf = treegen.coalescent(n, seed=5, mut_rate=mut_rate)
    genes = core_table(7, len(parent) - 1, first=1000)
    with build(seed, L, parent, rate, leaf, genes) as e:
        check_all_nodes(e, oracle, seed, parent, rate, genes[0], genes[1], L, range(7))


def test_edge_tables_match_oracle(oracle):
    parent, rate, leaf = treegen.coalescent(40, seed=11, mut_rate=0.05)
    rate[3] = 0.0
    rate[4] = 1e-12
    rate[5] = 30.0
    with build(1, 64, parent, rate, leaf, core_table(1, len(parent) - 1)) as e:
        for v in range(len(parent) - 1):
            H, k = e.edge_table(v)
            Ho, ko = oracle.edge_table(rate[v])
            assert k == ko and np.array_equal(H, Ho), f"edge {v} rate {rate[v]}"


def test_accessory_presence_and_rows(oracle):
    n, L, seed = 30, 150, 2024
    parent, rate, leaf = treegen.coalescent(n, seed=8, mut_rate=0.2)
    gene_id, origin, off, car = treegen.gene_table(parent, leaf, n_core=6, n_acc=40, seed=4)
    with build(seed, L, parent, rate, leaf, (gene_id, origin, off, car)) as e:
        genome_leaf = {int(g): v for v, g in enumerate(leaf) if g >= 0}
        for gi in range(len(gene_id)):
            c = car[off[gi]:off[gi + 1]]
            if len(c) == 1 and c[0] == -1:
                needed = set(range(len(parent)))
            else:
                needed = set()
                for g in c:
                    v = genome_leaf[int(g)]
                    while True:
                        needed.add(v)
                        if v == origin[gi]:
                            break
                        v = int(parent[v])
            want = oracle.gene_all_nodes(seed, parent, rate, int(origin[gi]), int(gene_id[gi]), L)
            for v in range(len(parent)):
                got = e.copy_packed(v, gi)
                assert (got is not None) == (v in needed), f"presence gene {gene_id[gi]} node {v}"
                if got is not None:
                    assert np.array_equal(got, want[v]), f"gene {gene_id[gi]} node {v}"
        s = e.stats()
        assert s["n_dense_genes"] == 6


def test_graph_and_plain_launches_agree_and_repeat(oracle):
    n, L, seed = 50, 1000, 5
    parent, rate, leaf = treegen.coalescent(n, seed=2, mut_rate=0.05)
    genes = treegen.gene_table(parent, leaf, n_core=20, n_acc=15, seed=1)
    with build(seed, L, parent, rate, leaf, genes) as a, \
            build(seed, L, parent, rate, leaf, genes, flags=PGS_FLAG_NO_GRAPH | PGS_FLAG_KEEP_INTERNAL) as b:
        a.sweep()  # a second sweep overwrites every row with the same values
        a.sync()
        for v in (0, n - 1, n, len(parent) - 1):
            assert np.array_equal(a.copy_dense(v, 20), b.copy_dense(v, 20))
        for gi in range(len(genes[0])):
            for v in range(0, len(parent), 7):
                ra, rb = a.copy_packed(v, gi), b.copy_packed(v, gi)
                assert (ra is None) == (rb is None)
                if ra is not None:
                    assert np.array_equal(ra, rb)


def test_sharding_does_not_change_rows(oracle):
    """genes are independent: an engine holding half of the table produces the same rows"""
    n, L, seed = 20, 300, 31
    parent, rate, leaf = treegen.coalescent(n, seed=6, mut_rate=0.1)
    root = len(parent) - 1
    full = core_table(10, root)
    with build(seed, L, parent, rate, leaf, full) as e, \
            build(seed, L, parent, rate, leaf, core_table(5, root, first=5)) as half:
        for v in range(len(parent)):
            assert np.array_equal(e.copy_dense(v, 10)[5:], half.copy_dense(v, 5))


def test_config2_sample(oracle):
    """BASELINE config 2: num-genomes=100, core-size=1000, gene-length=1000."""
    n, L, seed = 100, 1000, 1
    parent, rate, leaf = treegen.coalescent(n, seed=1, mut_rate=0.01)
    root = len(parent) - 1
    with build(seed, L, parent, rate, leaf, core_table(1000, root)) as e:
        s = e.stats()
        assert s["leaf_nucleotides"] == 100 * 1000 * 1000
        assert s["site_edges"] == 198 * 1000 * 1000
        rng = np.random.default_rng(0)
        sample = sorted(rng.choice(1000, size=25, replace=False))
        dense = {v: e.copy_dense(v, 1000) for v in range(len(parent))}
        for g in sample:
            want = oracle.gene_all_nodes(seed, parent, rate, root, int(g), L)
            for v in range(len(parent)):
                assert np.array_equal(dense[v][g], want[v]), f"gene {g} node {v}"
        # every substitution is a real change and padding stays zero
        pad_sites = 16 * 64 - 1000
        for v in (0, 50, root):
            codes = oracle.unpack_codes(dense[v], 1024)
            assert not codes[:, 1000:].any() and pad_sites == 24


def test_base_composition_and_jc_distance(oracle):
    """statistical parity: uniform base composition at the root and observed
    pairwise p-distance within the 99% band of the JC expectation, Bonferroni over all pairs."""
    from scipy.stats import norm
    n, L, G, seed = 16, 1000, 400, 12345
    parent, rate, leaf = treegen.coalescent(n, seed=21, mut_rate=0.05)
    root = len(parent) - 1
    with build(seed, L, parent, rate, leaf, core_table(G, root)) as e:
        S = G * L
        rootc = oracle.unpack_codes(e.copy_dense(root, G), L)
        freq = np.bincount(rootc.ravel(), minlength=4) / S
        assert np.all(np.abs(freq - 0.25) < 4.5 * np.sqrt(0.25 * 0.75 / S))
        mm, sites = e.mismatch()
        assert sites == S
        # expected distance along the tree path
        def path_rate(a, b):
            da, db = {}, 0.0
            v, acc = a, 0.0
            while v != -1:
                da[v] = acc
                acc += rate[v]
                v = parent[v]
            v = b
            while v not in da:
                db += rate[v]
                v = parent[v]
            return da[v] + db
        pairs = n * (n - 1) // 2
        z = norm.ppf(1 - 0.01 / (2 * pairs))
        for i in range(n):
            for j in range(i + 1, n):
                d = path_rate(i, j)
                p = 0.75 * (1 - np.exp(-4.0 / 3.0 * d))
                assert abs(mm[i, j] / S - p) <= z * np.sqrt(p * (1 - p) / S) + 1e-12, (i, j)


def test_mismatch_matches_numpy(oracle):
    n, L, G, seed = 70, 130, 33, 8
    parent, rate, leaf = treegen.coalescent(n, seed=13, mut_rate=0.4)
    root = len(parent) - 1
    with build(seed, L, parent, rate, leaf, core_table(G, root)) as e:
        mm, sites = e.mismatch()
        assert sites == G * L
        codes = np.stack([oracle.unpack_codes(e.copy_dense(v, G), L).ravel() for v in range(n)])
        want = (codes[:, None, :] != codes[None, :, :]).sum(-1).astype(np.uint32)
        assert np.array_equal(mm, want)
        assert oracle.mismatch(e.copy_dense(0, G).ravel(), e.copy_dense(1, G).ravel()) == want[0, 1]


def test_error_behaviour():
    e = Engine(1, 100)
    with pytest.raises(PgsError) as ex:
        e.sweep()
    assert ex.value.code == -2
    with pytest.raises(PgsError):
        e.set_tree([-1, -1, 0], [0, 0, 0.1], [-1, -1, 0])  # two roots
    with pytest.raises(PgsError):
        e.set_tree([1, 0], [0.1, 0.1], [0, 1])  # cycle, no root
    e.set_tree([2, 2, -1], [0.1, 0.2, 0.0], [0, 1, -1])
    with pytest.raises(PgsError):  # carrier outside the origin's subtree
        e.set_genes([5], [0], [0, 1], [1])
    with pytest.raises(PgsError):  # duplicate gene id
        e.set_genes([5, 5], [2, 2], [0, 1, 2], [-1, -1])
    e.set_genes([5], [2], [0, 1], [-1])
    e.sweep()
    assert e.copy_packed(0, 0) is not None
    e.close()


def test_non_binary_trees_and_degenerate_inputs(oracle):
    """the ABI takes any rooted tree: children are swept in sibling pairs (c0,c1), (c2,c3), (c4)"""
    # root 0 with five children 1..5; node 2 has three children 6,7,8; node 8 has one child 9
    parent = np.array([-1, 0, 0, 0, 0, 0, 2, 2, 2, 8], dtype=np.int32)
    rate = np.array([0, 0.3, 0.2, 0.01, 1.5, 0.0, 0.4, 0.05, 0.6, 0.9])
    leaf = np.array([-1, 0, -1, 1, 2, 3, 4, 5, -1, 6], dtype=np.int32)
    seed, L = 77, 200
    genes = (np.array([3, 9, 11], np.int64), np.array([0, 0, 2], np.int32), np.array([0, 1, 2, 4], np.int64),
             np.array([-1, -1, 4, 6], np.int32))
    with build(seed, L, parent, rate, leaf, genes) as e:
        assert e.stats()["n_genomes"] == 7 and e.stats()["n_dense_genes"] == 2
        check_all_nodes(e, oracle, seed, parent, rate, genes[0], genes[1], L, range(3))
        assert e.copy_packed(1, 2) is None and e.copy_packed(9, 2) is not None and e.copy_packed(7, 2) is None
    # a single genome: the tree is one node, nothing to sweep
    with build(5, 100, [-1], [0.0], [0], core_table(3, 0)) as e:
        assert e.stats()["levels"] == 0 and e.stats()["rows_written"] == 0
        assert np.array_equal(e.copy_packed(0, 1), oracle.row(5, [-1], [0.0], 0, 0, 1, 100))
        mm, sites = e.mismatch()
        assert mm.shape == (1, 1) and mm[0, 0] == 0
    # an empty gene table is legal and sweeps nothing
    parent, rate, leaf = treegen.coalescent(4, seed=1)
    with Engine(1, 64) as e:
        e.set_tree(parent, rate, leaf)
        e.set_genes([], [], [0], [])
        e.sweep()
        assert e.stats()["rows_total"] == 0
        assert e.write_outputs(None, 31 | 32, [], [])["records"] == 0


@pytest.mark.parametrize("n,L,mut_rate", [(2, 128, 0.3), (9, 128, 0.3), (40, 1000, 0.01), (200, 1000, 2.0), (700, 256, 0.05),
                                          (300, 512, 0.02)])
def test_fused_walk_equals_level_sweep_and_oracle(oracle, n, L, mut_rate):
    """the product default (fused column walk: children carried in registers, parked siblings,
    internal rows not kept) produces exactly the leaf and root rows of the level-synchronous sweep"""
    seed = 1000 + n
    parent, rate, leaf = treegen.coalescent(n, seed=n, mut_rate=mut_rate)
    genes = treegen.gene_table(parent, leaf, n_core=9, n_acc=6, seed=n + 1)
    gene_id, origin = genes[0], genes[1]
    root = len(parent) - 1
    with build(seed, L, parent, rate, leaf, genes, flags=0) as walk, build(seed, L, parent, rate, leaf, genes) as keep:
        assert walk.stats()["sweep_mode"] == 1 and keep.stats()["sweep_mode"] == 0
        assert walk.stats()["rows_stored"] < keep.stats()["rows_stored"] or n == 2
        assert walk.stats()["rows_loaded"] <= keep.stats()["rows_loaded"]
        for v in list(range(n)) + [root]:
            assert np.array_equal(walk.copy_dense(v, 9), keep.copy_dense(v, 9)), v
        dense = [gi for gi in range(len(gene_id)) if gene_id[gi] < 9]
        for gi in dense[:3]:
            want = oracle.gene_all_nodes(seed, parent, rate, root, int(gene_id[gi]), L)
            for v in list(range(n)) + [root]:
                assert np.array_equal(walk.copy_packed(v, gi), want[v]), (gi, v)
        if n > 2:
            with pytest.raises(PgsError):
                walk.copy_dense(n, 9)  # an internal node other than the root is not kept
        # sparse genes have their own fused walk: genomes' rows and origin rows are kept, nothing else
        is_leaf = np.ones(len(parent), bool)
        is_leaf[parent[parent >= 0]] = False
        for gi in range(len(gene_id)):
            if gene_id[gi] >= 9:
                for v in range(len(parent)):
                    b = keep.copy_packed(v, gi)
                    if b is None or is_leaf[v] or v == origin[gi]:
                        a = walk.copy_packed(v, gi)
                        assert (a is None) == (b is None) and (a is None or np.array_equal(a, b)), (gi, v)
                    else:
                        with pytest.raises(PgsError) as ex:
                            walk.copy_packed(v, gi)
                        assert ex.value.code == -2  # PGS_ERR_STATE, as include/pgs.h documents
        if n > 2:
            with pytest.raises(PgsError) as ex:
                walk.copy_packed(n, dense[0])  # internal dense row: not kept either
            assert ex.value.code == -2
        mw, _ = walk.mismatch()
        mk, _ = keep.mismatch()
        assert np.array_equal(mw, mk)
        walk.sweep()  # repeatable (parked rows are rewritten, discarded lines do not matter)
        walk.sync()
        for v in (0, n - 1, root):
            assert np.array_equal(walk.copy_dense(v, 9), keep.copy_dense(v, 9))


def _caterpillar(n, rate):
    """leaves 0..n-1; internal node n+i joins leaf i+1 with the previous internal node (or leaf 0)"""
    N = 2 * n - 1
    parent = np.full(N, -1, dtype=np.int32)
    prev = 0
    for i in range(n - 1):
        p = n + i
        parent[prev] = p
        parent[i + 1] = p
        prev = p
    r = np.full(N, rate)
    r[-1] = 0.0
    leaf = np.full(N, -1, dtype=np.int32)
    leaf[:n] = np.arange(n)
    return parent, r, leaf


def _balanced(depth, rate):
    n = 2 ** depth
    N = 2 * n - 1
    parent = np.full(N, -1, dtype=np.int32)
    level = list(range(n))
    nxt = n
    while len(level) > 1:
        up = []
        for i in range(0, len(level), 2):
            parent[level[i]] = parent[level[i + 1]] = nxt
            up.append(nxt)
            nxt += 1
        level = up
    r = np.full(N, rate)
    r[-1] = 0.0
    leaf = np.full(N, -1, dtype=np.int32)
    leaf[:n] = np.arange(n)
    return parent, r, leaf


@pytest.mark.parametrize("shape", ["caterpillar", "balanced"])
def test_fused_walk_on_extreme_tree_shapes(oracle, shape):
    parent, rate, leaf = _caterpillar(700, 0.004) if shape == "caterpillar" else _balanced(9, 0.02)
    n, L, seed, G = int((leaf >= 0).sum()), 256, 31, 12
    root = len(parent) - 1
    genes = core_table(G, root)
    with build(seed, L, parent, rate, leaf, genes, flags=0) as walk, build(seed, L, parent, rate, leaf, genes) as keep:
        assert walk.stats()["sweep_mode"] == 1
        for v in list(range(0, n, 7)) + [n - 1, root]:
            assert np.array_equal(walk.copy_dense(v, G), keep.copy_dense(v, G)), v
        for v in (0, n // 2, n - 1):
            assert np.array_equal(walk.copy_packed(v, 3), oracle.row(seed, parent, rate, v, root, 3, L))


def test_fused_walk_is_race_free_under_repetition(oracle):
    """pool slots of parked rows are reused inside a sweep; repeat a mid-sized sweep and compare
    every genome with the level-synchronous engine each time"""
    n, L, seed, G = 3000, 1024, 77, 64
    parent, rate, leaf = treegen.coalescent(n, seed=12, mut_rate=0.02)
    root = len(parent) - 1
    genes = core_table(G, root)
    with build(seed, L, parent, rate, leaf, genes) as keep:
        want = np.stack([keep.copy_dense(v, G) for v in range(n)])
    with build(seed, L, parent, rate, leaf, genes, flags=0) as walk:
        for rep in range(6):
            if rep:
                walk.sweep()
                walk.sync()
            got = np.stack([walk.copy_dense(v, G) for v in range(n)])
            assert np.array_equal(got, want), f"repetition {rep}"


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("smem_levels", [None, 0, 2])
def test_fused_walk_kernel_variants_and_stack_placement(oracle, monkeypatch, variant, smem_levels):
    """every walk kernel (block pairs / quads per thread) and every placement of the parked rows
    (a thread's shared-memory stack, the global pool, or the first levels here and the rest there)
    gives the rows of the level-synchronous sweep"""
    monkeypatch.setenv("PGS_WALK_VARIANT", str(variant))
    if smem_levels is not None:
        monkeypatch.setenv("PGS_WALK_SMEM_LEVELS", str(smem_levels))
    n, seed, G = 600, 5, 20
    parent, rate, leaf = treegen.coalescent(n, seed=21, mut_rate=0.05)
    root = len(parent) - 1
    genes = core_table(G, root)
    for L in (1000, 400, 100):  # 16 blocks, 7 blocks (level-synchronous fallback), 2 blocks (pairs)
        with build(seed, L, parent, rate, leaf, genes) as keep:
            want = np.stack([keep.copy_dense(v, G) for v in list(range(n)) + [root]])
        with build(seed, L, parent, rate, leaf, genes, flags=0) as walk:
            assert walk.stats()["sweep_mode"] == (0 if L == 400 else 1)
            for rep in range(2):
                if rep:
                    walk.sweep()
                    walk.sync()
                got = np.stack([walk.copy_dense(v, G) for v in list(range(n)) + [root]])
                assert np.array_equal(got, want), (L, rep)


@pytest.mark.parametrize("units,piece,carve", [(1, 1, 0), (3, 2, 0), (8, 5, 0), (40, 16, 0), (2, 1 << 30, 0), (1, 1 << 30, 12), (4, 1 << 30, 40), (6, 20, 30)])
@pytest.mark.parametrize("shape", ["coalescent", "caterpillar", "balanced"])
def test_fused_walk_units_cut_into_pieces(oracle, monkeypatch, units, piece, carve, shape):
    """units of the dense walk cut into pieces (a piece reaches its root from the unit's root across entry
    steps in registers; the pieces cut out of it are skipped): every cut gives the level-synchronous rows"""
    monkeypatch.setenv("PGS_WALK_UNITS", str(units))
    monkeypatch.setenv("PGS_WALK_PIECE", str(piece))
    monkeypatch.setenv("PGS_WALK_CARVE", str(carve))
    if shape == "coalescent":
        parent, rate, leaf = treegen.coalescent(500, seed=8, mut_rate=0.05)
    else:
        parent, rate, leaf = _caterpillar(300, 0.004) if shape == "caterpillar" else _balanced(8, 0.02)
    n, L, seed, G = int((leaf >= 0).sum()), 1000, 91, 9
    root = len(parent) - 1
    genes = core_table(G, root)
    with build(seed, L, parent, rate, leaf, genes) as keep:
        want = np.stack([keep.copy_dense(v, G) for v in list(range(n)) + [root]])
    with build(seed, L, parent, rate, leaf, genes, flags=0) as walk:
        assert walk.stats()["sweep_mode"] == 1
        for rep in range(2):
            if rep:
                walk.sweep()
                walk.sync()
            got = np.stack([walk.copy_dense(v, G) for v in list(range(n)) + [root]])
            assert np.array_equal(got, want), rep
    for v in (0, n // 3, n - 1):
        assert np.array_equal(want[v][3], oracle.row(seed, parent, rate, v, root, 3, L))


def _check_sparse_walk(oracle, e, seed, parent, rate, leaf, genes, L):
    gene_id, origin, off, car = genes
    genome_leaf = {int(g): v for v, g in enumerate(leaf) if g >= 0}
    checked = 0
    for gi in range(len(gene_id)):
        c = car[off[gi]:off[gi + 1]]
        if len(c) == 1 and c[0] == -1:
            continue  # dense gene
        want = oracle.gene_all_nodes(seed, parent, rate, int(origin[gi]), int(gene_id[gi]), L)
        for v in [genome_leaf[int(g)] for g in c] + [int(origin[gi])]:
            got = e.copy_packed(v, gi)
            assert got is not None and np.array_equal(got, want[v]), f"gene {gene_id[gi]} node {v}"
            checked += 1
        absent = set(genome_leaf.values()) - {genome_leaf[int(g)] for g in c}
        for v in list(absent)[:5]:
            assert e.copy_packed(v, gi) is None
    return checked


@pytest.mark.parametrize("L", [1, 64, 65, 100, 130, 1000, 1500])
def test_sparse_walk_rows_match_oracle(oracle, L):
    """fused walk of the accessory genes (k1_walk_sparse): origin rows drawn in the kernel, every genome's
    row equal to the host replay; block counts 1, 2, 3 (odd), 16, 24"""
    n, seed = 60, 4000 + L
    parent, rate, leaf = treegen.coalescent(n, seed=17, mut_rate=0.3)
    genes = treegen.gene_table(parent, leaf, n_core=3, n_acc=80, seed=L)
    with build(seed, L, parent, rate, leaf, genes, flags=0) as e:
        st = e.stats()
        assert st["n_dense_genes"] == 3
        # one launch for the sparse genes, two (masks, walk) for the dense ones where the dense walk applies
        assert st["kernel_launches"] <= 3 or ((L + 63) // 64) % 2 == 1
        assert _check_sparse_walk(oracle, e, seed, parent, rate, leaf, genes, L) > 80
        e.sweep()
        e.sync()
        assert _check_sparse_walk(oracle, e, seed, parent, rate, leaf, genes, L) > 80


@pytest.mark.parametrize("sparse_walk", ["1", "0"])
def test_sparse_walk_deep_stacks_and_rates(oracle, monkeypatch, sparse_walk):
    """genes carried by hundreds of genomes: the stack of parked siblings outgrows the shared-memory
    levels (the deeper ones go to the siblings' own rows); high rate: most blocks take the rare path"""
    monkeypatch.setenv("PGS_SPARSE_WALK", sparse_walk)
    n, L, seed = 700, 256, 606
    parent, rate, leaf = treegen.coalescent(n, seed=23, mut_rate=1.0)
    gene_id, origin, off, car = treegen.gene_table(parent, leaf, n_core=2, n_acc=24, seed=9)
    # plus four genes that appeared at the root and are carried by 300 .. 699 of the 700 genomes
    rng = np.random.default_rng(1)
    extra = [sorted(rng.choice(n, size=k, replace=False)) for k in (300, 450, 600, 699)]
    gene_id = np.concatenate([gene_id, np.arange(100, 104)])
    origin = np.concatenate([origin, np.full(4, len(parent) - 1, np.int32)])
    off = np.concatenate([off, off[-1] + np.cumsum([len(x) for x in extra])])
    car = np.concatenate([car] + [np.array(x, np.int32) for x in extra])
    genes = (gene_id, origin, off, car)
    with build(seed, L, parent, rate, leaf, genes, flags=0) as e:
        assert _check_sparse_walk(oracle, e, seed, parent, rate, leaf, genes, L) > 2000


@pytest.mark.parametrize("n,L,G,planes_mb", [(70, 130, 33, None), (200, 1000, 9, "1"), (129, 64, 3, None), (2, 100, 5, None),
                                              (333, 448, 21, "1")])
def test_mismatch_tiles_chunks_and_edges(oracle, monkeypatch, n, L, G, planes_mb):
    """K2 (bit planes + 64 x 64 tiles of the upper triangle): genome counts off the tile size, column
    chunks (a 1 MB plane budget forces several), packed tiles against the square, all against numpy"""
    if planes_mb:
        monkeypatch.setenv("PGS_K2_PLANES_MB", planes_mb)
    seed = 50 + n
    parent, rate, leaf = treegen.coalescent(n, seed=n, mut_rate=0.5)
    root = len(parent) - 1
    with build(seed, L, parent, rate, leaf, core_table(G, root), flags=0) as e:
        mm, sites = e.mismatch()
        assert sites == G * L
        codes = np.stack([oracle.unpack_codes(e.copy_dense(v, G), L).reshape(-1) for v in range(n)])
        want = np.zeros((n, n), np.uint32)
        for i in range(n):
            want[i] = (codes != codes[i]).sum(-1)
        assert np.array_equal(mm, want)
        # the same through the device-side square and the packed tiles
        import torch
        sq = torch.zeros((n, n), dtype=torch.int32, device="cuda")
        assert e.mismatch_device(sq.data_ptr()) == sites
        assert np.array_equal(sq.cpu().numpy().astype(np.uint32), want)
        ptr, words, _ = e.mismatch_tiles_device()
        T = (n + 63) // 64
        assert words == T * (T + 1) // 2 * 4096
        tiles = np.zeros(words, np.uint32)
        torch.cuda.synchronize()
        import ctypes
        from pangenomesim_b200.capi import load_library
        cudart = ctypes.CDLL("libcudart.so")
        assert cudart.cudaMemcpy(tiles.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(ptr), ctypes.c_size_t(words * 4), 2) == 0
        tiles = tiles.reshape(-1, 64, 64)
        k = 0
        for ti in range(T):
            for tj in range(ti, T):
                blockw = np.zeros((64, 64), np.uint32)
                r, c = want[64 * ti:64 * ti + 64, 64 * tj:64 * tj + 64].shape
                blockw[:r, :c] = want[64 * ti:64 * ti + 64, 64 * tj:64 * tj + 64]
                if ti == tj:
                    blockw = np.triu(blockw, 1)
                assert np.array_equal(tiles[k], blockw), (ti, tj)
                k += 1
        del load_library


def test_two_engines_with_different_plans_share_a_device(oracle):
    """the walk kernels' dynamic shared-memory limit is a per-function, per-device attribute: engines
    with different stack depths (different trees and gene counts) on ONE device, planned and swept in
    turn and at the same time from two threads, must not lower it under each other's feet"""
    import threading
    cases = []
    for n, G, L in ((900, 40, 1000), (30, 7, 256), (2500, 16, 512)):
        parent, rate, leaf = treegen.coalescent(n, seed=n, mut_rate=0.03)
        root = len(parent) - 1
        genes = treegen.gene_table(parent, leaf, n_core=G, n_acc=10, seed=n)
        with build(77, L, parent, rate, leaf, genes) as keep:
            want = np.stack([keep.copy_dense(v, G) for v in (0, n // 2, n - 1, root)])
        cases.append((n, G, L, parent, rate, leaf, genes, root, want))
    engines = [build(77, c[2], c[3], c[4], c[5], c[6], flags=0) for c in cases]
    levels = {e.stats()["levels"] for e in engines}
    assert len(levels) == 3
    errors = []

    def hammer(k):
        try:
            e, (n, G, L, parent, rate, leaf, genes, root, want) = engines[k], cases[k]
            for rep in range(6):
                if rep % 2:  # re-plan: sets the attribute again while the other engines launch
                    e.set_tree(parent, rate, leaf)
                    e.set_genes(*genes)
                e.sweep()
                e.sync()
                got = np.stack([e.copy_dense(v, G) for v in (0, n // 2, n - 1, root)])
                assert np.array_equal(got, want), (k, rep)
        except Exception as ex:  # noqa: BLE001
            errors.append((k, ex))
    threads = [threading.Thread(target=hammer, args=(k,)) for k in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in engines:
        e.close()
    assert not errors, errors


@pytest.mark.parametrize("delay", [None, "400000"])
def test_walk_never_starts_from_a_stale_origin_row(oracle, monkeypatch, delay):
    """k1_walk_dense is a programmatic dependent of k1_top_masks, which writes the origin rows: an engine that
    alternates between two gene tables (same layout, other gene ids -> other origin rows at the same
    addresses) while other engines keep the GPU busy must never carry the OTHER table's origin row into its
    genomes (ptxas once scheduled the origin row's ld.global.nc above griddepcontrol.wait).  With
    PGS_TEST_ORIGIN_DELAY the origin rows are written 0.2 ms late: a load above the wait is stale every time."""
    import threading
    if delay:
        monkeypatch.setenv("PGS_TEST_ORIGIN_DELAY", delay)
    n, G, L = 300, 24, 128
    parent, rate, leaf = treegen.coalescent(n, seed=5, mut_rate=0.02)
    root = len(parent) - 1
    tables, wants = [], []
    for first_id in (0, 1000):
        genes = list(treegen.gene_table(parent, leaf, n_core=G, n_acc=0, seed=3))
        genes[0] = np.asarray(genes[0], dtype=np.int64) + first_id
        with build(91, L, parent, rate, leaf, genes) as e:  # level-synchronous engine: no dependent launch in it
            wants.append(np.stack([e.copy_dense(v, G) for v in (0, n // 3, n - 1, root)]))
        tables.append(genes)
    assert not np.array_equal(wants[0], wants[1])
    stop = threading.Event()

    def noise(k):  # other engines sweeping in a loop: the top-mask grid of the engine under test starts late
        p2, r2, l2 = treegen.coalescent(2000, seed=k, mut_rate=0.02)
        g2 = treegen.gene_table(p2, l2, n_core=64, n_acc=0, seed=k)
        with build(k, 1000, p2, r2, l2, g2, flags=0) as e2:
            while not stop.is_set():
                e2.sweep()
                e2.sync()
    threads = [threading.Thread(target=noise, args=(k,)) for k in (1, 2, 3)]
    for t in threads:
        t.start()
    try:
        with build(91, L, parent, rate, leaf, tables[0], flags=0) as e:
            for rep in range(40):
                which = rep % 2
                e.set_genes(*tables[which])
                e.sweep()
                e.sync()
                got = np.stack([e.copy_dense(v, G) for v in (0, n // 3, n - 1, root)])
                assert np.array_equal(got, wants[which]), rep
    finally:
        stop.set()
        for t in threads:
            t.join()
