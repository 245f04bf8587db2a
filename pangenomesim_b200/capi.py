"""ctypes binding of include/pgs.h (one class, no logic of its own).

Every method maps 1:1 onto a C entry point; arrays are numpy and are passed as plain pointers.
There is deliberately no fallback: if ``libpgs.so`` cannot be loaded, or no CUDA device is
usable, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

PGS_ALL_GENOMES = -1
PGS_FLAG_NO_GRAPH = 1
PGS_FLAG_KEEP_INTERNAL = 2
PGS_FLAG_K2_POPC = 4
PGS_FLAG_K2_TENSOR = 8
OUT_REF, OUT_CORE, OUT_ACCESSORY, OUT_GENOMES, OUT_MAF, OUT_ALL, OUT_DISCARD, OUT_PACKED, OUT_HOST_EXPAND = 1, 2, 4, 8, 16, 31, 32, 64, 128


class PgsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"pgs error {code}: {message}")
        self.code = code
        self.message = message


class _Config(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("gene_length", C.c_int64), ("device", C.c_int32), ("flags", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "n_nodes", "n_genomes", "n_genes", "n_dense_genes", "blocks_per_gene", "rows_total",
        "rows_written", "rows_read", "rows_origin", "leaf_nucleotides", "site_edges", "levels",
        "kernel_launches", "device_bytes", "sweep_mode", "rows_stored", "rows_loaded")] + [(k, C.c_double) for k in (
        "sweep_ms", "rootgen_ms", "mismatch_ms", "format_ms", "dense_ms", "sparse_ms")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64)


class ShardLayout(C.Structure):
    _fields_ = [("genome_offset", C.POINTER(C.c_int64)), ("maf_offset", C.c_int64),
                ("genome_gene_total", C.POINTER(C.c_int64)), ("n_ref_total", C.c_int64)]


class OutputSpec(C.Structure):
    _fields_ = [("out_dir", C.c_char_p), ("what", C.c_uint32), ("n_ref_core", C.c_int64),
                ("ref_core", C.POINTER(C.c_int64)), ("n_ref_acc", C.c_int64), ("ref_acc", C.POINTER(C.c_int64)),
                ("sink", SINK_FN), ("sink_user", C.c_void_p), ("shard", C.POINTER(ShardLayout))]


class OutputReport(C.Structure):
    _fields_ = [("files", C.c_int64), ("bytes", C.c_int64), ("records", C.c_int64),
                ("device_ms", C.c_double), ("wall_ms", C.c_double), ("d2h_bytes", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def lib_path() -> Path:
    return Path(os.environ.get("PGS_LIBRARY", _HERE / "lib" / "libpgs.so"))


def load_library():
    """Load libpgs.so (built in-tree by ``make lib`` / ``__graft_entry__.build()``)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not path.exists():
        raise PgsError(-6, f"{path} is missing: build it with `make lib`; there is no CPU fallback")
    lib = C.CDLL(str(path))
    P = C.POINTER
    lib.pgs_abi_version.restype = C.c_int
    lib.pgs_device_count.restype = C.c_int
    lib.pgs_last_error.restype = C.c_char_p
    lib.pgs_last_error.argtypes = [C.c_void_p]
    lib.pgs_create.argtypes = [P(_Config), P(C.c_void_p)]
    lib.pgs_destroy.argtypes = [C.c_void_p]
    lib.pgs_destroy.restype = None
    lib.pgs_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.pgs_sync.argtypes = [C.c_void_p]
    lib.pgs_set_tree.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pgs_set_genes.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pgs_sweep.argtypes = [C.c_void_p]
    lib.pgs_get_stats.argtypes = [C.c_void_p, P(Stats)]
    lib.pgs_mismatch.argtypes = [C.c_void_p, C.c_void_p, P(C.c_int64)]
    lib.pgs_mismatch_device.argtypes = [C.c_void_p, C.c_void_p, P(C.c_int64)]
    lib.pgs_mismatch_tile_words.argtypes = [C.c_int32]
    lib.pgs_mismatch_tile_words.restype = C.c_int64
    lib.pgs_mismatch_tiles_device.argtypes = [C.c_void_p, P(C.c_void_p), P(C.c_int64)]
    lib.pgs_mismatch_multi.argtypes = [P(C.c_void_p), C.c_int, C.c_void_p, P(C.c_int64)]
    lib.pgs_comm_unique_id.argtypes = [C.c_void_p]
    lib.pgs_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.pgs_mismatch_allreduce.argtypes = [C.c_void_p, C.c_void_p, P(C.c_int64), P(C.c_double)]
    lib.pgs_write_outputs.argtypes = [C.c_void_p, P(OutputSpec), P(OutputReport)]
    lib.pgs_output_sizes.argtypes = [C.c_void_p, P(OutputSpec), C.c_void_p, P(C.c_int64)]
    lib.pgs_copy_origin_ascii.argtypes = [C.c_void_p, C.c_int64, C.c_char_p]
    lib.pgs_genome_name.argtypes = [C.c_int64, C.c_char_p, C.c_size_t]
    lib.pgs_copy_packed.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
    lib.pgs_copy_dense.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.pgs_row_index.argtypes = [C.c_void_p, C.c_int32, C.c_int64]
    lib.pgs_row_index.restype = C.c_int64
    lib.pgs_get_edge_table.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, P(C.c_int32)]
    lib.pgs_expand_bases.argtypes = [C.c_void_p, C.c_int64, C.c_char_p]
    lib.pgs_expand_bases.restype = C.c_char_p
    _LIB = lib
    return lib


def genome_name(index: int) -> str:
    buf = C.create_string_buffer(64)
    load_library().pgs_genome_name(index, buf, 64)
    return buf.value.decode()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Engine:
    """One engine = one CUDA device.  Mirrors the call order of include/pgs.h."""

    def __init__(self, seed: int, gene_length: int, device: int = -1, flags: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        cfg = _Config(seed, gene_length, device, flags)
        rc = self._lib.pgs_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise PgsError(rc, self._lib.pgs_last_error(None).decode())
        self.gene_length = gene_length
        self.n_nodes = 0
        self.n_genomes = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pgs_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc < 0:
            raise PgsError(rc, self._lib.pgs_last_error(self._h).decode())
        return rc

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._lib.pgs_set_stream(self._h, cuda_stream_ptr))

    def sync(self):
        self._check(self._lib.pgs_sync(self._h))

    def set_tree(self, parent, rate, leaf_genome):
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        rate = np.ascontiguousarray(rate, dtype=np.float64)
        leaf_genome = np.ascontiguousarray(leaf_genome, dtype=np.int32)
        assert parent.shape == rate.shape == leaf_genome.shape
        self._check(self._lib.pgs_set_tree(self._h, len(parent), _ptr(parent), _ptr(rate), _ptr(leaf_genome)))
        self.n_nodes = len(parent)
        self.n_genomes = int((leaf_genome >= 0).sum())

    def set_genes(self, gene_id, origin, carrier_off, carrier_genome, all_order=None):
        gene_id = np.ascontiguousarray(gene_id, dtype=np.int64)
        origin = np.ascontiguousarray(origin, dtype=np.int32)
        carrier_off = np.ascontiguousarray(carrier_off, dtype=np.int64)
        carrier_genome = np.ascontiguousarray(carrier_genome, dtype=np.int32)
        if all_order is not None:
            all_order = np.ascontiguousarray(all_order, dtype=np.int32)
        assert len(origin) == len(gene_id) and len(carrier_off) == len(gene_id) + 1
        self._check(self._lib.pgs_set_genes(self._h, len(gene_id), _ptr(gene_id), _ptr(origin),
                                            _ptr(carrier_off), _ptr(carrier_genome), _ptr(all_order)))
        self.n_genes = len(gene_id)

    def set_core_genes(self, n_genes: int, root: int, first_id: int = 0):
        """Convenience: n_genes genes with ids first_id.., origin = root, carried by every genome."""
        ids = np.arange(first_id, first_id + n_genes, dtype=np.int64)
        self.set_genes(ids, np.full(n_genes, root, np.int32), np.arange(n_genes + 1, dtype=np.int64),
                       np.full(n_genes, PGS_ALL_GENOMES, np.int32))

    def sweep(self):
        self._check(self._lib.pgs_sweep(self._h))

    def stats(self) -> dict:
        s = Stats()
        self._check(self._lib.pgs_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def mismatch(self):
        out = np.zeros((self.n_genomes, self.n_genomes), dtype=np.uint32)
        sites = C.c_int64()
        self._check(self._lib.pgs_mismatch(self._h, _ptr(out), C.byref(sites)))
        return out, sites.value

    def mismatch_device(self, device_ptr: int) -> int:
        sites = C.c_int64()
        self._check(self._lib.pgs_mismatch_device(self._h, device_ptr, C.byref(sites)))
        return sites.value

    def mismatch_tiles_device(self):
        """(device pointer of the packed upper-triangular tiles, number of uint32 words, sites)"""
        ptr, sites = C.c_void_p(), C.c_int64()
        self._check(self._lib.pgs_mismatch_tiles_device(self._h, C.byref(ptr), C.byref(sites)))
        return ptr.value, self._lib.pgs_mismatch_tile_words(self.n_genomes), sites.value

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """join the NCCL communicator of a multi-process run (one engine per process)"""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.pgs_comm_init(self._h, buf, rank, world))

    def mismatch_allreduce(self, want_matrix: bool = True):
        """K2 on this rank's genes, NCCL all-reduce of the partial counts; returns (matrix or None,
        this rank's sites, all-reduce milliseconds)"""
        out = np.zeros((self.n_genomes, self.n_genomes), dtype=np.uint32) if want_matrix else None
        sites, ms = C.c_int64(), C.c_double()
        self._check(self._lib.pgs_mismatch_allreduce(self._h, _ptr(out), C.byref(sites), C.byref(ms)))
        return out, sites.value, ms.value

    def blocks_per_gene(self) -> int:
        return (self.gene_length + 63) // 64

    def copy_packed(self, node: int, gene_index: int):
        """Packed row as uint32 words, or None when the gene is absent at the node."""
        dst = np.zeros(self.blocks_per_gene() * 4, dtype=np.uint32)
        rc = self._check(self._lib.pgs_copy_packed(self._h, node, gene_index, _ptr(dst)))
        return None if rc == 1 else dst

    def copy_dense(self, node: int, n_dense: int):
        dst = np.zeros((n_dense, self.blocks_per_gene() * 4), dtype=np.uint32)
        self._check(self._lib.pgs_copy_dense(self._h, node, _ptr(dst)))
        return dst

    def row_index(self, node: int, gene_index: int) -> int:
        return self._lib.pgs_row_index(self._h, node, gene_index)

    def edge_table(self, node: int):
        H = np.zeros(65, dtype=np.uint64)
        k = C.c_int32()
        self._check(self._lib.pgs_get_edge_table(self._h, node, _ptr(H), C.byref(k)))
        return H, k.value

    @staticmethod
    def _shard(genome_offset, maf_offset, genome_gene_total, n_ref_total):
        keep = []
        sh = ShardLayout()
        if genome_offset is not None:
            a = np.ascontiguousarray(genome_offset, dtype=np.int64)
            keep.append(a)
            sh.genome_offset = a.ctypes.data_as(C.POINTER(C.c_int64))
        b = np.ascontiguousarray(genome_gene_total, dtype=np.int64)
        keep.append(b)
        sh.genome_gene_total = b.ctypes.data_as(C.POINTER(C.c_int64))
        sh.maf_offset = maf_offset
        sh.n_ref_total = n_ref_total
        return sh, keep

    def output_sizes(self, n_ref_core=0, n_ref_acc=0, shard=None):
        """(bytes per genomeNNN.fasta [n_genomes], bytes of alignment.maf) this engine will write;
        shard = dict(maf_offset=, genome_gene_total=, n_ref_total=) for a sliced gene table."""
        spec = OutputSpec()
        spec.what = OUT_GENOMES | OUT_MAF
        dummy = np.zeros(max(n_ref_core, n_ref_acc, 1), dtype=np.int64)
        spec.n_ref_core, spec.n_ref_acc = 0, 0
        keep = None
        if shard is not None:
            sh, keep = self._shard(None, shard.get("maf_offset", 0), shard["genome_gene_total"], shard["n_ref_total"])
            spec.shard = C.pointer(sh)
        else:
            # ref_size only depends on the counts: pass index lists of the right length
            idx = np.zeros(n_ref_core + n_ref_acc, dtype=np.int64)
            spec.n_ref_core = len(idx)
            spec.ref_core = idx.ctypes.data_as(C.POINTER(C.c_int64))
        gb = np.zeros(self.n_genomes, dtype=np.int64)
        mb = C.c_int64()
        self._check(self._lib.pgs_output_sizes(self._h, C.byref(spec), _ptr(gb), C.byref(mb)))
        del dummy, keep
        return gb, mb.value

    def copy_origin_ascii(self, gene_index: int) -> bytes:
        buf = C.create_string_buffer(self.gene_length)
        self._check(self._lib.pgs_copy_origin_ascii(self._h, gene_index, buf))
        return buf.raw

    def write_outputs(self, out_dir=None, what=OUT_ALL, ref_core=(), ref_acc=(), sink=None, shard=None):
        rc_arr = np.ascontiguousarray(ref_core, dtype=np.int64)
        ra_arr = np.ascontiguousarray(ref_acc, dtype=np.int64)
        spec = OutputSpec()
        keep = None
        if shard is not None:
            sh, keep = self._shard(shard["genome_offset"], shard["maf_offset"], shard["genome_gene_total"],
                                   shard["n_ref_total"])
            spec.shard = C.pointer(sh)
        spec.out_dir = str(out_dir).encode() if out_dir is not None else None
        spec.what = what
        spec.n_ref_core = len(rc_arr)
        spec.ref_core = rc_arr.ctypes.data_as(C.POINTER(C.c_int64))
        spec.n_ref_acc = len(ra_arr)
        spec.ref_acc = ra_arr.ctypes.data_as(C.POINTER(C.c_int64))
        cb = None
        if sink is not None:
            def _cb(user, name, offset, data, length):
                try:
                    return int(sink(name.decode(), offset, data, length) or 0)
                except Exception:  # never unwind through C
                    return 1
            cb = SINK_FN(_cb)
            spec.sink = cb
        rep = OutputReport()
        self._check(self._lib.pgs_write_outputs(self._h, C.byref(spec), C.byref(rep)))
        return rep.as_dict()


def expand_bases(packed_words, sites: int):
    """host-side expander of PGS_OUT_HOST_EXPAND: (ASCII bytes, code path name)"""
    w = np.ascontiguousarray(packed_words, dtype=np.uint32)
    guard = b"\xaa" * 64  # nothing may be written behind the last site
    buf = C.create_string_buffer(b"\x00" * int(sites) + guard)
    kind = load_library().pgs_expand_bases(_ptr(w), sites, buf)
    if buf.raw[sites:sites + 64] != guard:
        raise PgsError(-1, "pgs_expand_bases wrote behind the last site")
    return buf.raw[:sites], kind.decode()


def comm_unique_id() -> bytes:
    """NCCL unique id (128 bytes) for Engine.comm_init: draw it on one rank, hand it to all"""
    buf = C.create_string_buffer(128)
    rc = load_library().pgs_comm_unique_id(buf)
    if rc != 0:
        raise PgsError(rc, "pgs_comm_unique_id: NCCL is not available")
    return buf.raw


def mismatch_multi(engines):
    """engines of this process, one per device: (summed n x n matrix, sites)"""
    lib = load_library()
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    n = engines[0].n_genomes
    out = np.zeros((n, n), dtype=np.uint32)
    sites = C.c_int64()
    rc = lib.pgs_mismatch_multi(arr, len(engines), _ptr(out), C.byref(sites))
    if rc < 0:
        raise PgsError(rc, lib.pgs_last_error(engines[0]._h).decode())
    return out, sites.value


def host_coalescent(n: int, seed: int, mut_rate: float = 0.01, return_time: bool = False):
    """The reference's coalescent (img.cxx:71-116) as drawn by the host program for `seed`:
    returns parent[int32], rate[float64] = branch time * mut_rate, leaf_genome[int32]."""
    path = _HERE / "lib" / "libpgshost.so"
    if not path.exists():
        raise PgsError(-6, f"{path} is missing: build it with `make cli`")
    lib = C.CDLL(str(path))
    lib.pgsh_coalescent.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    N = 2 * n - 1
    parent = np.zeros(N, dtype=np.int32)
    time = np.zeros(N, dtype=np.float64)
    if lib.pgsh_coalescent(n, seed, _ptr(parent), _ptr(time)) != 0:
        raise PgsError(-1, "pgsh_coalescent: bad arguments (n >= 1, seed >= 1)")
    leaf = np.full(N, -1, dtype=np.int32)
    leaf[:n] = np.arange(n)
    if return_time:
        return parent, time, leaf
    return parent, time * float(np.float32(mut_rate)), leaf


class HostModel:
    """The host program's model (pangenomesim_b200/host/model.cpp: coalescent + IMG gain/loss, the
    reference's img_model without nucleotides) for Python callers: parameters as on the command line,
    results as the arrays pgs_set_tree / pgs_set_genes / pgs_write_outputs take."""

    def __init__(self, rng_mode: str = "auto", **params):
        path = _HERE / "lib" / "libpgshost.so"
        if not path.exists():
            raise PgsError(-6, f"{path} is missing: build it with `make cli`")
        lib = C.CDLL(str(path))
        lib.pgsh_model_create.restype = C.c_void_p
        lib.pgsh_model_destroy.argtypes = [C.c_void_p]
        lib.pgsh_model_error.restype = C.c_char_p
        lib.pgsh_model_error.argtypes = [C.c_void_p]
        lib.pgsh_model_param.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        lib.pgsh_model_simulate.argtypes = [C.c_void_p, C.c_int]
        lib.pgsh_model_sizes.argtypes = [C.c_void_p, C.c_void_p]
        lib.pgsh_model_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 11
        lib.pgsh_model_distance_row.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        self._lib = lib
        self._h = lib.pgsh_model_create()
        try:
            for k, v in params.items():
                if lib.pgsh_model_param(self._h, k.replace("_", "-").encode(), str(v).encode()) != 0:
                    raise PgsError(-1, lib.pgsh_model_error(self._h).decode())
            if lib.pgsh_model_simulate(self._h, {"auto": 0, "compat": 1, "fast": 2}[rng_mode]) != 0:
                raise PgsError(-1, lib.pgsh_model_error(self._h).decode())
            sz = np.zeros(7, dtype=np.int64)
            lib.pgsh_model_sizes(self._h, _ptr(sz))
            N, n, G, nc, nrc, nra, compat = (int(x) for x in sz)
            self.n_nodes, self.n_genomes, self.n_genes, self.used_compat = N, n, G, bool(compat)
            self.parent = np.zeros(N, np.int32)
            self.rate = np.zeros(N, np.float64)
            self.leaf_genome = np.zeros(N, np.int32)
            self.gene_id = np.zeros(G, np.int64)
            self.origin = np.zeros(G, np.int32)
            self.carrier_off = np.zeros(G + 1, np.int64)
            self.carriers = np.zeros(nc, np.int32)
            self.all_order = np.zeros(n, np.int32)
            self.ref_core = np.zeros(nrc, np.int64)
            self.ref_acc = np.zeros(nra, np.int64)
            self.genome_gene_count = np.zeros(n, np.int64)
            lib.pgsh_model_tables(self._h, _ptr(self.parent), _ptr(self.rate), _ptr(self.leaf_genome), _ptr(self.gene_id),
                                  _ptr(self.origin), _ptr(self.carrier_off), _ptr(self.carriers), _ptr(self.all_order),
                                  _ptr(self.ref_core), _ptr(self.ref_acc), _ptr(self.genome_gene_count))
        except Exception:
            self.close()
            raise

    @property
    def root(self) -> int:
        return self.n_nodes - 1

    def gene_table(self):
        """arguments of Engine.set_genes"""
        return self.gene_id, self.origin, self.carrier_off, self.carriers, self.all_order

    def distance_row(self, i: int):
        row = np.zeros(self.n_genomes, np.float64)
        self._lib.pgsh_model_distance_row(self._h, i, _ptr(row))
        return row

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pgsh_model_destroy(self._h)
            self._h = None

    __del__ = close
