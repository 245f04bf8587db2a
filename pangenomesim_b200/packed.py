"""Reader of ``genomes.pgs2`` (``PGS_OUT_PACKED``; layout in ``include/pgs.h``): the genomes'
sequences 2-bit packed behind an index -- the optional compact side output of the engine
(the reference has no such format).  Pure numpy, no GPU, no library needed."""
from __future__ import annotations

import numpy as np

_MAGIC = b"PGS2BIT\0"
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def genome_name(index: int) -> str:
    """genome_name of the reference, util.cxx:15-27 (without its abort beyond genome999)"""
    return "genomeref" if index < 0 else "genome%03d" % (index + 1)


class PackedGenomes:
    def __init__(self, path):
        self._m = np.memmap(path, dtype=np.uint8, mode="r")
        head = self._m[:56].tobytes()
        if head[:8] != _MAGIC:
            raise ValueError(f"{path}: not a genomes.pgs2 file")
        version, self.data_offset = np.frombuffer(head, dtype="<u4", count=2, offset=8)
        if version != 1:
            raise ValueError(f"{path}: version {version} is not supported")
        self.n_genomes, self.gene_length, self.blocks, self.n_records, self.n_dense = (
            int(x) for x in np.frombuffer(head, dtype="<i8", count=5, offset=16))
        n_other = self.n_records - self.n_genomes * self.n_dense
        at = 56
        self.record_off = np.frombuffer(self._m, dtype="<i8", count=self.n_genomes + 1, offset=at)
        at += 8 * (self.n_genomes + 1)
        self.dense_gene_id = np.frombuffer(self._m, dtype="<i8", count=self.n_dense, offset=at)
        self.dense_key = np.frombuffer(self._m, dtype="<i8", count=self.n_dense, offset=at + 8 * self.n_dense)
        at += 16 * self.n_dense
        self.other_gene_id = np.frombuffer(self._m, dtype="<i8", count=n_other, offset=at)
        self.other_key = np.frombuffer(self._m, dtype="<i8", count=n_other, offset=at + 8 * n_other)
        self.row_bytes = 16 * self.blocks
        derived = (at + 16 * n_other + 255) & ~255  # end of the index, rounded up to 256
        if self.data_offset not in (0, derived):
            raise ValueError(f"{path}: inconsistent header")
        self.data_offset = derived
        if int(self.data_offset) + self.n_records * self.row_bytes != self._m.size:
            raise ValueError(f"{path}: truncated ({self._m.size} bytes)")

    def row(self, record: int) -> np.ndarray:
        """packed row of a record: uint32 words, site s = bits 2(s%16).. of word s//16"""
        at = int(self.data_offset) + record * self.row_bytes
        return np.frombuffer(self._m, dtype="<u4", count=4 * self.blocks, offset=at)

    def sequence(self, record: int) -> bytes:
        """the record's nucleotides as ASCII"""
        w = self.row(record)
        codes = (w[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3
        return _ACGT[codes.reshape(-1)[: self.gene_length]].tobytes()

    def records(self, genome: int):
        """(gene_id, record index) of one genome in the order of its genomeNNN.fasta"""
        lo, hi = int(self.record_off[genome]), int(self.record_off[genome + 1])
        o0 = lo - genome * self.n_dense
        o1 = o0 + (hi - lo - self.n_dense)
        ids = np.concatenate([self.dense_gene_id, self.other_gene_id[o0:o1]])
        keys = np.concatenate([self.dense_key, self.other_key[o0:o1]])
        return [(int(ids[k]), lo + int(k)) for k in np.argsort(keys, kind="stable")]

    def fasta(self, genome: int) -> bytes:
        """exactly the bytes of genomeNNN.fasta (gene::to_fasta, reference gene.h:42-52)"""
        name = genome_name(genome).encode()
        out = []
        for gid, rec in self.records(genome):
            out.append(b">" + name + b"." + str(gid).encode() + b"\n" + self.sequence(rec) + b"\n")
        return b"".join(out)
