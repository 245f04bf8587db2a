"""pangenomesim_b200/packed.py (reader of genomes.pgs2, the optional 2-bit side output; layout in
include/pgs.h) against a file written here from known sequences -- no GPU needed."""
import numpy as np
import pytest

from pangenomesim_b200.packed import PackedGenomes, genome_name


def write_pgs2(path, L, dense, genomes):
    """dense: [(gene_id, key)] shared by all genomes; genomes: per genome (dense sequences,
    [(gene_id, key, sequence)] of its other records), rows in that order"""
    blocks = (L + 63) // 64
    n, D = len(genomes), len(dense)
    rec_off = np.zeros(n + 1, dtype="<i8")
    for i, (dseq, other) in enumerate(genomes):
        assert len(dseq) == D
        rec_off[i + 1] = rec_off[i] + D + len(other)
    total = int(rec_off[-1])
    n_other = total - n * D
    index = 56 + 8 * (n + 1) + 16 * (D + n_other)
    data_off = (index + 255) & ~255
    rows = np.zeros((total, blocks * 4), dtype="<u4")
    code = {c: k for k, c in enumerate("ACGT")}
    r = 0
    for dseq, other in genomes:
        for seq in list(dseq) + [o[2] for o in other]:
            for s, c in enumerate(seq):
                rows[r, s // 16] |= np.uint32(code[c] << (2 * (s % 16)))
            r += 1
    with open(path, "wb") as f:
        f.write(b"PGS2BIT\0" + np.array([1, data_off], dtype="<u4").tobytes())
        f.write(np.array([n, L, blocks, total, D], dtype="<i8").tobytes())
        f.write(rec_off.tobytes())
        f.write(np.array([d[0] for d in dense], dtype="<i8").tobytes() + np.array([d[1] for d in dense], dtype="<i8").tobytes())
        f.write(np.array([o[0] for g in genomes for o in g[1]], dtype="<i8").tobytes())
        f.write(np.array([o[1] for g in genomes for o in g[1]], dtype="<i8").tobytes())
        f.write(b"\0" * (data_off - index))
        f.write(rows.tobytes())


@pytest.mark.parametrize("L", [1, 16, 63, 64, 65, 200])
def test_reader_round_trip(tmp_path, L):
    rng = np.random.default_rng(L)
    seq = lambda: "".join(rng.choice(list("ACGT"), size=L))
    dense = [(3, 1), (8, 4)]  # (gene_id, key)
    genomes = [([seq(), seq()], [(7, 2, seq()), (12, 0, seq()), (5, 9, seq())]), ([seq(), seq()], []),
               ([seq(), seq()], [(5, 9, seq())])]
    write_pgs2(tmp_path / "genomes.pgs2", L, dense, genomes)
    pk = PackedGenomes(tmp_path / "genomes.pgs2")
    assert (pk.n_genomes, pk.gene_length, pk.n_records, pk.n_dense) == (3, L, 10, 2)
    for i, (dseq, other) in enumerate(genomes):
        recs = [(dense[k][1], dense[k][0], dseq[k]) for k in range(2)] + [(o[1], o[0], o[2]) for o in other]
        want = "".join(">%s.%d\n%s\n" % (genome_name(i), gid, s) for _, gid, s in sorted(recs))
        assert pk.fasta(i) == want.encode()
    assert pk.sequence(0) == genomes[0][0][0].encode()


def test_reader_rejects_foreign_and_truncated_files(tmp_path):
    (tmp_path / "x").write_bytes(b"not a packed file" * 10)
    with pytest.raises(ValueError):
        PackedGenomes(tmp_path / "x")
    write_pgs2(tmp_path / "y", 10, [(1, 0)], [(["ACGTACGTAC"], [])])
    data = (tmp_path / "y").read_bytes()
    (tmp_path / "z").write_bytes(data[:-4])
    with pytest.raises(ValueError):
        PackedGenomes(tmp_path / "z")
    assert genome_name(-1) == "genomeref" and genome_name(0) == "genome001" and genome_name(999) == "genome1000"
