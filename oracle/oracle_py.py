"""ctypes binding of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "liboracle.so"
REF_BIN = HERE / "_ref" / "pangenomesim"
REF_SEQPATH = HERE / "_ref" / "ref_seqpath"
_LIB = None


def build(quiet: bool = True):
    """Compile the checkers (and, where /root/reference exists, the unmodified reference)."""
    subprocess.run(["make", "-C", str(HERE)], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        build()
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    L.oracle_philox4x32_10.argtypes = [vp, vp, vp]
    L.oracle_philox4x32_10.restype = None
    L.oracle_edge_table.argtypes = [dbl, vp]
    L.oracle_edge_table.restype = C.c_int
    L.oracle_jc_p.argtypes = [dbl]
    L.oracle_jc_p.restype = dbl
    L.oracle_root_block.argtypes = [u64, u32, u32, i64, vp]
    L.oracle_root_block.restype = None
    L.oracle_mutate_block.argtypes = [u64, u32, u32, u32, u32, u32, i64, vp, vp]
    L.oracle_sibling_pairs.argtypes = [i32, vp, vp, vp]
    L.oracle_sibling_pairs.restype = None
    L.oracle_mutate_block.restype = None
    L.oracle_row.argtypes = [u64, i32, vp, vp, i32, i32, u32, i64, vp]
    L.oracle_row.restype = C.c_int
    L.oracle_gene_all_nodes.argtypes = [u64, i32, vp, vp, i32, u32, i64, vp]
    L.oracle_gene_all_nodes.restype = C.c_int
    L.oracle_mismatch.argtypes = [vp, vp, i64]
    L.oracle_mismatch.restype = u64
    L.oracle_unpack_ascii.argtypes = [vp, i64, vp]
    L.oracle_unpack_ascii.restype = None
    L.ref_engine_seed.argtypes = [vp, u64]
    L.ref_engine_seed.restype = None
    L.ref_engine_next.argtypes = [vp]
    L.ref_engine_next.restype = u32
    L.ref_uniform_int.argtypes = [vp, u32]
    L.ref_uniform_int.restype = u32
    L.ref_uniform_real.argtypes = [vp]
    L.ref_uniform_real.restype = dbl
    L.ref_exponential.argtypes = [vp, dbl]
    L.ref_exponential.restype = dbl
    L.ref_coalescent.argtypes = [vp, i64, vp, vp, vp, vp]
    L.ref_coalescent.restype = None
    L.ref_gene_root.argtypes = [vp, i64, vp]
    L.ref_gene_root.restype = None
    L.ref_gene_mutate.argtypes = [vp, vp, i64, dbl, vp]
    L.ref_gene_mutate.restype = i64
    L.ref_gene_mutate_count.argtypes = [vp, i64, dbl]
    L.ref_gene_mutate_count.restype = i64
    L.ref_sim_core_gene.argtypes = [vp, i32, vp, vp, vp, vp, dbl, i64, vp, vp]
    L.ref_sim_core_gene.restype = None
    L.ref_expected_hits.argtypes = [dbl, i64]
    L.ref_expected_hits.restype = dbl
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def edge_table(rate: float):
    H = np.zeros(65, dtype=np.uint64)
    kmax = lib().oracle_edge_table(float(rate), _p(H))
    return H, kmax


def jc_p(rate: float) -> float:
    return lib().oracle_jc_p(float(rate))


def blocks_of(gene_length: int) -> int:
    return (gene_length + 63) // 64


def row(seed, parent, rate, node, origin, gene_id, gene_length):
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    rate = np.ascontiguousarray(rate, dtype=np.float64)
    dst = np.zeros(4 * blocks_of(gene_length), dtype=np.uint32)
    rc = lib().oracle_row(seed, len(parent), _p(parent), _p(rate), node, origin, gene_id, gene_length, _p(dst))
    if rc != 0:
        raise ValueError("node is not below origin")
    return dst


def gene_all_nodes(seed, parent, rate, origin, gene_id, gene_length):
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    rate = np.ascontiguousarray(rate, dtype=np.float64)
    dst = np.zeros((len(parent), 4 * blocks_of(gene_length)), dtype=np.uint32)
    rc = lib().oracle_gene_all_nodes(seed, len(parent), _p(parent), _p(rate), origin, gene_id, gene_length, _p(dst))
    if rc != 0:
        raise ValueError("bad origin")
    return dst


def mismatch(a, b) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    return int(lib().oracle_mismatch(_p(a), _p(b), len(a) // 4))


def unpack_ascii(row_words, gene_length: int) -> bytes:
    r = np.ascontiguousarray(row_words, dtype=np.uint32)
    buf = C.create_string_buffer(gene_length)
    lib().oracle_unpack_ascii(_p(r), gene_length, buf)
    return buf.raw


def unpack_codes(words, gene_length: int):
    """numpy: packed uint32 words [..., W] -> 2-bit codes [..., gene_length] (uint8)."""
    w = np.asarray(words, dtype=np.uint32)
    shifts = (2 * np.arange(16, dtype=np.uint32))
    codes = (w[..., :, None] >> shifts) & 3
    return codes.reshape(*w.shape[:-1], -1)[..., :gene_length].astype(np.uint8)


class RefEngine:
    """minstd_rand0 as the reference uses it (global.h:6)."""

    def __init__(self, seed: int):
        self.state = C.c_uint32(0)
        lib().ref_engine_seed(C.byref(self.state), seed)

    def ptr(self):
        return C.byref(self.state)

    def next(self) -> int:
        return lib().ref_engine_next(self.ptr())

    def uniform_int(self, hi: int) -> int:
        return lib().ref_uniform_int(self.ptr(), hi)

    def uniform_real(self) -> float:
        return lib().ref_uniform_real(self.ptr())

    def coalescent(self, n: int):
        """parent, left, right (int32) and branch time (float64) in the reference's pool layout"""
        m = 2 * n - 1
        parent = np.zeros(m, np.int32)
        left = np.zeros(m, np.int32)
        right = np.zeros(m, np.int32)
        time = np.zeros(m, np.float64)
        lib().ref_coalescent(self.ptr(), n, _p(parent), _p(left), _p(right), _p(time))
        return parent, left, right, time

    def gene_root(self, length: int) -> bytes:
        buf = C.create_string_buffer(length)
        lib().ref_gene_root(self.ptr(), length, buf)
        return buf.raw

    def gene_mutate(self, seq: bytes, rate: float):
        buf = C.create_string_buffer(len(seq))
        hits = lib().ref_gene_mutate(self.ptr(), seq, len(seq), float(rate), buf)
        return buf.raw, hits

    def gene_mutate_count(self, length: int, rate: float) -> int:
        return lib().ref_gene_mutate_count(self.ptr(), length, float(rate))

    def sim_core_gene(self, root, left, right, time, leaf_index, mut_rate, length, n_leaves):
        left = np.ascontiguousarray(left, dtype=np.int32)
        right = np.ascontiguousarray(right, dtype=np.int32)
        time = np.ascontiguousarray(time, dtype=np.float64)
        leaf_index = np.ascontiguousarray(leaf_index, dtype=np.int32)
        root_out = C.create_string_buffer(length)
        leaves = C.create_string_buffer(length * n_leaves)
        lib().ref_sim_core_gene(self.ptr(), root, _p(left), _p(right), _p(time), _p(leaf_index),
                                float(mut_rate), length, root_out, leaves)
        raw = leaves.raw
        return root_out.raw, [raw[i * length:(i + 1) * length] for i in range(n_leaves)]


def expected_hits(rate: float, length: int) -> float:
    return lib().ref_expected_hits(float(rate), length)


def run_reference(out_dir, params: dict, timeout=600):
    """Run the unmodified reference binary (oracle/_ref/pangenomesim)."""
    arg = ",".join(f"{k}={v}" for k, v in params.items())
    subprocess.run([str(REF_BIN), "-o", str(out_dir), "-p", arg], check=True, timeout=timeout,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
