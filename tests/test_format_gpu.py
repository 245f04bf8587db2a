"""GPU parity of K3 (FASTA / MAF text) against a host restatement of the reference's writers
(gene::to_fasta gene.h:42-52; pangenomesim.cxx:117-157, 236-293) fed with the oracle's rows."""
import numpy as np
import pytest

import treegen
from pangenomesim_b200 import Engine
from pangenomesim_b200.capi import genome_name, OUT_ALL, OUT_DISCARD, OUT_GENOMES, OUT_HOST_EXPAND, OUT_MAF

pytestmark = pytest.mark.gpu


def expected_files(oracle, seed, L, parent, rate, leaf, genes, ref_core, ref_acc, all_order=None):
    gene_id, origin, off, car = genes
    n = int((leaf >= 0).sum())
    genome_leaf = {int(g): v for v, g in enumerate(leaf) if g >= 0}
    below = treegen.leaves_below(parent, leaf)
    order = list(range(n)) if all_order is None else list(all_order)
    rows = {}
    carriers = []
    for gi in range(len(gene_id)):
        c = [int(x) for x in car[off[gi]:off[gi + 1]]]
        if c == [-1]:
            c = [g for g in order if g in set(below[int(origin[gi])])]
        carriers.append(c)
        allrows = oracle.gene_all_nodes(seed, parent, rate, int(origin[gi]), int(gene_id[gi]), L)
        rows[gi] = {v: oracle.unpack_ascii(allrows[v], L).decode() for v in [int(origin[gi])] + [genome_leaf[g] for g in c]}
    count = [sum(1 for c in carriers if g in c) for g in range(n)]
    files = {}

    def rec(name, gi, node):
        return ">%s.%d\n%s\n" % (name, gene_id[gi], rows[gi][node])
    files["ref.fasta"] = "".join(rec("genomeref", gi, int(origin[gi])) for gi in list(ref_core) + list(ref_acc))
    files["core.fasta"] = "".join(rec("genomeref", gi, int(origin[gi])) for gi in ref_core)
    files["accessory.fasta"] = "".join(rec("genomeref", gi, int(origin[gi])) for gi in ref_acc)
    for g in range(n):
        files[genome_name(g) + ".fasta"] = "".join(
            rec(genome_name(g), gi, genome_leaf[g]) for gi in range(len(gene_id)) if g in carriers[gi])
    ref_size = (len(ref_core) + len(ref_acc)) * L
    maf = ["##maf version=1 program=pangenomesim\n"]
    for gi in range(len(gene_id)):
        if not carriers[gi]:
            continue
        maf.append("a\n")
        maf.append("s genomeref.%d 0 %d + %d %s\n" % (gene_id[gi], L, ref_size, rows[gi][int(origin[gi])]))
        for g in carriers[gi]:
            maf.append("s %s.%d 0 %d + %d %s\n" % (genome_name(g), gene_id[gi], L, count[g] * L, rows[gi][genome_leaf[g]]))
        maf.append("\n")
    files["alignment.maf"] = "".join(maf)
    return files


def run_case(oracle, tmp_path, n, L, n_core, n_acc, seed, all_order=None, use_sink=False, what=OUT_ALL):
    parent, rate, leaf = treegen.coalescent(n, seed=seed, mut_rate=0.3)
    genes = treegen.gene_table(parent, leaf, n_core=n_core, n_acc=n_acc, seed=seed + 1)
    gene_id = genes[0]
    ref_core = [i for i in range(len(gene_id)) if gene_id[i] < n_core]
    ref_acc = [i for i in range(len(gene_id)) if gene_id[i] >= n_core][::-1]
    e = Engine(seed, L)
    e.set_tree(parent, rate, leaf)
    e.set_genes(*genes, all_order=all_order)
    e.sweep()
    want = expected_files(oracle, seed, L, parent, rate, leaf, genes, ref_core, ref_acc, all_order)
    if use_sink:
        import ctypes
        got = {}

        def sink(name, offset, data, length):
            buf = ctypes.string_at(data, length)
            assert len(got.get(name, b"")) == offset, "chunks must arrive in file order"
            got[name] = got.get(name, b"") + buf
        rep = e.write_outputs(None, what, ref_core, ref_acc, sink=sink)
        got = {k: v.decode() for k, v in got.items()}
    else:
        rep = e.write_outputs(tmp_path, what, ref_core, ref_acc)
        got = {p.name: p.read_text() for p in tmp_path.iterdir()}
    assert set(got) == set(want)
    for name in want:
        assert got[name] == want[name], name
    assert rep["files"] == len(want) and rep["bytes"] == sum(len(v) for v in want.values())
    if what & OUT_HOST_EXPAND:  # and the same totals when the text is produced and dropped
        rep2 = e.write_outputs(None, what | OUT_DISCARD, ref_core, ref_acc)
        assert rep2["files"] == rep["files"] and rep2["bytes"] == rep["bytes"] and rep2["records"] == rep["records"]
        assert 0 < rep2["d2h_bytes"] < rep2["bytes"]  # packed rows (and the text of the three small reference files)
    e.close()


@pytest.mark.parametrize("L", [1, 15, 16, 17, 63, 64, 100, 1000])
def test_fasta_and_maf_core_only(oracle, tmp_path, L):
    run_case(oracle, tmp_path, n=5, L=L, n_core=12, n_acc=0, seed=3 + L)


def test_fasta_and_maf_with_accessory_genes(oracle, tmp_path):
    run_case(oracle, tmp_path, n=14, L=77, n_core=4, n_acc=30, seed=9)


def test_accessory_only_and_sink(oracle, tmp_path):
    run_case(oracle, tmp_path, n=9, L=130, n_core=0, n_acc=25, seed=4, use_sink=True)


def test_all_order_drives_maf_line_order(oracle, tmp_path):
    order = np.random.default_rng(2).permutation(11).astype(np.int32)
    run_case(oracle, tmp_path, n=11, L=50, n_core=3, n_acc=8, seed=6, all_order=order)


def test_more_than_999_genomes_and_many_chunks(oracle):
    """names beyond genome999 (the reference aborts there) and a run that needs several chunks"""
    n, L, G, seed = 1200, 1000, 70, 5
    parent, rate, leaf = treegen.coalescent(n, seed=2, mut_rate=0.01)
    root = len(parent) - 1
    e = Engine(seed, L)
    e.set_tree(parent, rate, leaf)
    e.set_core_genes(G, root)
    e.sweep()
    import ctypes
    sizes, heads = {}, {}

    def sink(name, offset, data, length):
        assert sizes.get(name, 0) == offset
        if offset == 0:
            heads[name] = ctypes.string_at(data, min(length, 40))
        sizes[name] = offset + length
    rep = e.write_outputs(None, OUT_GENOMES | OUT_MAF, list(range(G)), [], sink=sink)
    assert heads["genome1000.fasta"].startswith(b">genome1000.0\n")
    assert heads["genome001.fasta"].startswith(b">genome001.0\n")
    assert len(sizes) == n + 1 and rep["files"] == n + 1
    digits = sum(len(str(g)) for g in range(G))
    for g in (0, 998, 999, 1199):
        name = genome_name(g)
        assert sizes[name + ".fasta"] == G * (len(name) + L + 4) + digits
    assert rep["bytes"] == sum(sizes.values()) and rep["bytes"] > (128 << 20)
    # spot-check one record deep inside a late chunk against the packed row
    row = e.copy_packed(int(np.where(leaf == 1100)[0][0]), 37)
    want = oracle.unpack_ascii(row, L)
    got = {}

    def sink2(name, offset, data, length):
        if name == "genome1101.fasta":
            got[name] = got.get(name, b"") + ctypes.string_at(data, length)
    e.write_outputs(None, OUT_GENOMES, list(range(G)), [], sink=sink2)
    recs = got["genome1101.fasta"].split(b"\n")
    assert recs[2 * 37] == b">genome1101.37" and recs[2 * 37 + 1] == want
    e.close()


@pytest.mark.parametrize("presize", [False, True])
@pytest.mark.parametrize("extra", [0, OUT_HOST_EXPAND])
def test_sharded_outputs_equal_single_engine(oracle, tmp_path, extra, presize):
    """two engines that each hold a slice of the gene table fill disjoint byte ranges of the same
    files (pgs_shard_layout); the result is byte-identical to one engine holding everything"""
    n, L, seed = 13, 90, 21
    parent, rate, leaf = treegen.coalescent(n, seed=5, mut_rate=0.2)
    gene_id, origin, off, car = treegen.gene_table(parent, leaf, n_core=7, n_acc=20, seed=8)
    G = len(gene_id)
    below = treegen.leaves_below(parent, leaf)
    total = np.zeros(n, dtype=np.int64)
    for g in range(G):
        c = car[off[g]:off[g + 1]]
        for x in (below[int(origin[g])] if list(c) == [-1] else c):
            total[int(x)] += 1
    order = np.random.default_rng(3).permutation(n).astype(np.int32)
    what = OUT_GENOMES | OUT_MAF | extra

    single = tmp_path / "single"
    single.mkdir()
    with Engine(seed, L) as e:
        e.set_tree(parent, rate, leaf)
        e.set_genes(gene_id, origin, off, car, all_order=order)
        e.sweep()
        e.write_outputs(single, OUT_GENOMES | OUT_MAF, list(range(G)), [])
        ref_seq = {int(gene_id[g]): e.copy_origin_ascii(g) for g in range(G)}

    cut = 11
    parts = []
    for lo, hi in ((0, cut), (cut, G)):
        e = Engine(seed, L)
        e.set_tree(parent, rate, leaf)
        e.set_genes(gene_id[lo:hi], origin[lo:hi], off[lo:hi + 1] - off[lo], car[off[lo]:off[hi]], all_order=order)
        e.sweep()
        parts.append(e)
    shard0 = dict(maf_offset=0, genome_gene_total=total, n_ref_total=G)
    gb0, mb0 = parts[0].output_sizes(shard=shard0)
    gb1, mb1 = parts[1].output_sizes(shard=dict(shard0, maf_offset=mb0))
    sharded = tmp_path / "sharded"
    sharded.mkdir()
    for i, name in enumerate([genome_name(i) + ".fasta" for i in range(n)] + ["alignment.maf"]):
        # created empty, or (as the host program does) at the final size: the engines never resize a shared file
        (sharded / name).write_bytes(b"\0" * int((gb0[i] + gb1[i]) if i < n else (mb0 + mb1)) if presize else b"")
    parts[1].write_outputs(sharded, what, [], [], shard=dict(shard0, genome_offset=gb0, maf_offset=mb0))
    parts[0].write_outputs(sharded, what, [], [], shard=dict(shard0, genome_offset=np.zeros(n, np.int64)))
    for p in sorted(single.iterdir()):
        assert (sharded / p.name).read_bytes() == p.read_bytes(), p.name
    assert (sharded / "alignment.maf").stat().st_size == mb0 + mb1
    # the host writes the reference files itself in sharded mode
    for g in range(G):
        eng, idx = (parts[0], g) if g < cut else (parts[1], g - cut)
        assert eng.copy_origin_ascii(idx) == ref_seq[int(gene_id[g])]
    for e in parts:
        e.close()


@pytest.mark.parametrize("n,L,n_core,n_acc,stage_mb", [(14, 77, 4, 30, None), (9, 130, 0, 25, None), (300, 1000, 40, 60, 1),
                                                      (5, 64, 3, 0, None)])
def test_packed_side_output_holds_the_fasta_sequences(tmp_path, monkeypatch, n, L, n_core, n_acc, stage_mb):
    """PGS_OUT_PACKED (genomes.pgs2: rows copied straight from the store behind an index, SURVEY.md
    §8f-3): unpacked by pangenomesim_b200/packed.py it gives every genomeNNN.fasta byte for byte;
    with a 1 MB staging buffer the rows arrive in many chunks"""
    from pangenomesim_b200.capi import OUT_PACKED
    from pangenomesim_b200.packed import PackedGenomes
    if stage_mb is not None:
        monkeypatch.setenv("PGS_STAGE_MB", str(stage_mb))
    parent, rate, leaf = treegen.coalescent(n, seed=n, mut_rate=0.3)
    genes = treegen.gene_table(parent, leaf, n_core=n_core, n_acc=n_acc, seed=n + 1)
    with Engine(99, L) as e:
        e.set_tree(parent, rate, leaf)
        e.set_genes(*genes)
        e.sweep()
        rep = e.write_outputs(tmp_path, OUT_GENOMES | OUT_PACKED, [], [])
        pk = PackedGenomes(tmp_path / "genomes.pgs2")
        assert pk.n_genomes == n and pk.gene_length == L
        total = 0
        for g in range(n):
            text = (tmp_path / (genome_name(g) + ".fasta")).read_bytes()
            assert pk.fasta(g) == text, g
            total += len(text)
        assert rep["files"] == n + 1 and rep["bytes"] == total + (tmp_path / "genomes.pgs2").stat().st_size
        # the same through a sink, chunks in file order
        import ctypes
        got = {}

        def sink(name, offset, data, length):
            assert len(got.get(name, b"")) == offset
            got[name] = got.get(name, b"") + ctypes.string_at(data, length)
        e.write_outputs(None, OUT_PACKED, [], [], sink=sink)
        assert got["genomes.pgs2"] == (tmp_path / "genomes.pgs2").read_bytes()


# ---- PGS_OUT_HOST_EXPAND: packed rows over PCIe, text written by the host's cores; same bytes

@pytest.mark.parametrize("L", [1, 16, 63, 64, 65, 100, 128, 1000, 1500])
def test_host_expand_equals_reference_layout(oracle, tmp_path, L):
    run_case(oracle, tmp_path, n=7, L=L, n_core=6, n_acc=9, seed=30 + L, what=OUT_ALL | OUT_HOST_EXPAND)


def test_host_expand_sink_and_order(oracle, tmp_path):
    order = np.random.default_rng(5).permutation(13).astype(np.int32)
    run_case(oracle, tmp_path, n=13, L=130, n_core=3, n_acc=25, seed=8, all_order=order, use_sink=True, what=OUT_ALL | OUT_HOST_EXPAND)


@pytest.mark.parametrize("ring,task_lines,mmap", [("1", "1024", "1"), ("2", "40", "1"), ("4", "40", "1"), ("4", "300", "0"), ("3", "1", "1")])
def test_host_expand_maf_ring(oracle, tmp_path, monkeypatch, ring, task_lines, mmap):
    """alignment.maf through the ring of staging slots (PGS_HX_RING slots per buffer, 1 MB staging: dozens of
    chunks in flight, retired in order), tasks of several 256-line runs / several tasks per block, mapped
    windows and pwrite: same bytes as the reference layout"""
    monkeypatch.setenv("PGS_STAGE_MB", "1")
    monkeypatch.setenv("PGS_HX_RING", ring)
    monkeypatch.setenv("PGS_HX_TASK_LINES", task_lines)
    monkeypatch.setenv("PGS_WRITE_MMAP", mmap)
    monkeypatch.setenv("PGS_EXPAND_THREADS", "7")
    run_case(oracle, tmp_path, n=600, L=1000, n_core=25, n_acc=40, seed=14, what=OUT_ALL | OUT_HOST_EXPAND)


@pytest.mark.parametrize("run_records", [None, "7", "1", "one-copy"])
@pytest.mark.parametrize("use_sink", [False, True])
def test_host_expand_many_chunks(oracle, tmp_path, monkeypatch, use_sink, run_records):
    """1 MB staging: the genomes and the MAF lines take several chunks each (double buffering,
    mapped windows of the MAF file at unaligned offsets)"""
    # ("one-copy": both texts from one copy of the rows, chunk of genes by chunk of genes; sink and PGS_OUT_DISCARD only)
    if run_records == "one-copy":
        monkeypatch.setenv("PGS_HX_ONE_COPY", "1")
    elif run_records:  # genomeNNN.fasta delivered in several pieces (runs of records through an L2-sized buffer)
        monkeypatch.setenv("PGS_HX_RUN_RECORDS", run_records)
    monkeypatch.setenv("PGS_STAGE_MB", "1")
    monkeypatch.setenv("PGS_EXPAND_THREADS", "5")
    run_case(oracle, tmp_path, n=150, L=1000, n_core=30, n_acc=40, seed=12, use_sink=use_sink, what=OUT_ALL | OUT_HOST_EXPAND)
