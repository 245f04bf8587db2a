/* include/pgs.h -- C ABI of the B200 sequence-evolution engine for pangenomesim.
 *
 * The reference (kloetzl/pangenomesim) has no plugin or FFI layer: its de-facto boundary is the
 * `gene` class used by img.cxx / pangenomesim.cxx.  Each entry point below names the reference
 * interface it replaces (file:line into the reference tree).  The host program keeps the
 * coalescent, the IMG gain/loss process, the CLI and the small text outputs, and calls this
 * library for everything that touches nucleotides.
 *
 * Conventions: opaque handle; every call returns PGS_OK (0) or a negative pgs_status and never
 * exits or throws across the boundary (the reference exits via err/errx, util.cxx:37,44,69);
 * pgs_last_error() gives the message.  Inputs are copied during the call.  All pointers are plain
 * host pointers unless the name says `_device`.  One engine drives one CUDA device; a host thread
 * may own several engines (one per device) -- genes shard across engines with no communication
 * during the sweep.  There is no CPU fallback: without a usable CUDA device pgs_create fails.
 *
 * Data model
 *   tree     N nodes, parent[i] (root: -1), rate[i] = mu * t of the branch above node i
 *            (reference: tree_node::get_time() * mut_rate, img.cxx:218,286), leaf_genome[i] =
 *            genome index of a leaf or -1.  Node index i is also the RNG edge id of the branch
 *            above i.
 *   genes    G entries: gene_id (reference gene::gene_id, gene.h:11), origin node (root for
 *            core / start genes, img.cxx:210,266; node v for a gene gained on the branch into v,
 *            img.cxx:302 -- such a gene is NOT mutated on that branch), and the genomes that
 *            carry it at the leaves (the IMG presence bitmask, as a list).
 *   rows     one 2-bit packed sequence per present (node, gene): A=0 C=1 G=2 T=3, site s of a
 *            gene in 32-bit word s/16 at bits 2(s%16); genes padded to 64-site blocks (16 B).
 */
#ifndef PGS_H
#define PGS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGS_ABI_VERSION 3

typedef struct pgs_engine pgs_engine;

typedef enum pgs_status {
	PGS_OK = 0,
	PGS_ERR_ARG = -1,     /* invalid argument / inconsistent tables */
	PGS_ERR_STATE = -2,   /* call order (e.g. sweep before set_genes) */
	PGS_ERR_CUDA = -3,    /* CUDA runtime error, message has the detail */
	PGS_ERR_NOMEM = -4,   /* host or device allocation failed */
	PGS_ERR_IO = -5,      /* file output failed (errno in the message) */
	PGS_ERR_NO_DEVICE = -6 /* no usable CUDA device: there is no CPU path */
} pgs_status;

typedef struct pgs_config {
	uint64_t seed;       /* resolved seed (reference: img_model::seed, img.cxx:239-243) */
	int64_t gene_length; /* reference: img_model::gene_length, img.h:15 */
	int32_t device;      /* CUDA ordinal, -1 = current device */
	uint32_t flags;      /* PGS_FLAG_* */
} pgs_config;

#define PGS_FLAG_NO_GRAPH 1u /* launch the level sweep kernel by kernel instead of one CUDA graph */
/* Keep the row of every internal node in memory (level-synchronous sweep).  Without this flag
 * the genes are swept by the fused walks (binary trees), which only materialise what the outputs
 * need: the rows of the genomes (leaves) and of the origins; pgs_copy_packed / pgs_copy_dense on
 * any other node then fail with PGS_ERR_STATE. */
#define PGS_FLAG_KEEP_INTERNAL 2u
/* Which kernel counts the pairwise mismatches (pgs_mismatch*): the popcount-XOR kernel, or the
 * same counts as a +-1 contraction on the tensor cores (tcgen05, FP4 operands with unit block
 * scales, FP32 accumulation -- exact: the counts are identical integers, checked against each
 * other by tests/test_mismatch_tensor_gpu.py).  Neither flag: the tensor-core kernel. */
#define PGS_FLAG_K2_POPC 4u
#define PGS_FLAG_K2_TENSOR 8u

/* sentinel carrier entry: "every genome below the origin node carries the gene" */
#define PGS_ALL_GENOMES (-1)

/* ---- lifetime ------------------------------------------------------------------------- */

/* Replaces: construction of the sequence store, img.h:32-35. */
int pgs_create(const pgs_config *cfg, pgs_engine **out);
void pgs_destroy(pgs_engine *h);
/* Message of the last failing call on h (h may be NULL: last pgs_create failure of this thread). */
const char *pgs_last_error(const pgs_engine *h);
int pgs_abi_version(void);
/* Number of usable CUDA devices (0 if none: there is no CPU path). */
int pgs_device_count(void);

/* Use a caller-owned stream (a cudaStream_t cast to void*; NULL = the engine's own stream). */
int pgs_set_stream(pgs_engine *h, void *cuda_stream);
int pgs_sync(pgs_engine *h);

/* ---- inputs --------------------------------------------------------------------------- */

/* Replaces: the tree_node pool walked by tree_node::traverse (img.h:316-329) in
 * sim_core_gene (img.cxx:215-231) and the accessory traversal (img.cxx:278-333). */
int pgs_set_tree(pgs_engine *h, int32_t n_nodes, const int32_t *parent, const double *rate,
				 const int32_t *leaf_genome);

/* Replaces: gene::gene(len, -1, id) at img.cxx:210,266,302 (registration of a gene and its
 * origin) and img_model::add at the leaves (img.cxx:226,328; img.h:44-49) (presence).
 * carrier_off has n_genes+1 entries into carrier_genome (genome indices; a gene whose list is
 * the single entry PGS_ALL_GENOMES is carried by every genome below its origin).  The order of a
 * gene's carriers is the order of its `s` lines in alignment.maf; all_order (n_genomes entries,
 * may be NULL = ascending) gives that order for PGS_ALL_GENOMES genes.  Table order is the
 * order of records in genomeNNN.fasta and of blocks in alignment.maf (reference:
 * img_model::gene_ids(), img.cxx:429-432).  Plans and allocates device memory. */
int pgs_set_genes(pgs_engine *h, int64_t n_genes, const int64_t *gene_id, const int32_t *origin_node,
				  const int64_t *carrier_off, const int32_t *carrier_genome, const int32_t *all_order);

/* ---- the hot path --------------------------------------------------------------------- */

/* Replaces: every gene::gene(len,..) root draw (gene.cxx:13-25) and every gene::mutate /
 * gene::bulk_mutate call (gene.cxx:27-61, gene.h:26-40) of img_model::simulate.  Root rows are
 * generated (kernel K0), then the tree is swept level by level (kernel K1).  Asynchronous on the
 * engine's stream; pgs_sync waits. */
int pgs_sweep(pgs_engine *h);

typedef struct pgs_stats {
	int64_t n_nodes, n_genomes, n_genes, n_dense_genes;
	int64_t blocks_per_gene;     /* 64-site blocks (16 B) per gene row */
	int64_t rows_total;          /* present (node, gene) rows held in HBM */
	int64_t rows_written;        /* rows produced by K1 (W of SURVEY.md §8d) */
	int64_t rows_read;           /* rows read by K1 (R) */
	int64_t rows_origin;         /* rows produced by K0 */
	int64_t leaf_nucleotides;    /* sum over genomes of carried genes * gene_length */
	int64_t site_edges;          /* rows_written * gene_length */
	int64_t levels;              /* tree levels below the root */
	int64_t kernel_launches;     /* kernels launched by the last pgs_sweep (fused walks: masks + walk for the dense genes, one for the sparse genes) */
	int64_t device_bytes;        /* bytes of HBM held by the engine */
	int64_t sweep_mode;          /* 0 = level-synchronous, 1 = fused column walk (dense genes) */
	int64_t rows_stored;         /* rows K1 actually writes to memory */
	int64_t rows_loaded;         /* rows K1 actually reads from memory */
	double sweep_ms, rootgen_ms; /* device time of the last completed pgs_sweep (CUDA events) */
	double mismatch_ms, format_ms;
	/* K1 split into the dense (core) and the sparse (accessory) genes; measured by sweeps that are
	 * launched kernel by kernel (the first sweep of a plan, or PGS_FLAG_NO_GRAPH) */
	double dense_ms, sparse_ms;
} pgs_stats;
int pgs_get_stats(pgs_engine *h, pgs_stats *out); /* synchronises */

/* Replaces nothing in the reference (its distance.mat is the EXPECTED distance, img.cxx:121-165,
 * printed at pangenomesim.cxx:214-234): OBSERVED pairwise mismatch counts over the dense
 * (all-genome, root-origin) genes, kernel K2 -- what that matrix is checked against.
 * out is n_genomes x n_genomes uint32, symmetric, zero diagonal.  *sites_out = sites compared.
 * On the device the counts are kept as packed upper-triangular 64 x 64 tiles (never the square);
 * the result reaches the host through pinned buffers in pieces. */
int pgs_mismatch(pgs_engine *h, uint32_t *out_host, int64_t *sites_out);
/* Same, full square left in caller-provided DEVICE memory (n_genomes^2 uint32). */
int pgs_mismatch_device(pgs_engine *h, void *out_device_u32, int64_t *sites_out);
/* The packed form: tile (ti <= tj) number ti*T - ti(ti-1)/2 + (tj - ti), T = ceil(n/64), holds
 * counts[r][c] of genomes (64 ti + r, 64 tj + c) for 64 ti + r < 64 tj + c, zero elsewhere.
 * *tiles_device_u32 = the engine's own buffer (pgs_mismatch_tile_words(n) uint32, valid until the
 * next mismatch call on h): one contiguous range to all-reduce across engines. */
int64_t pgs_mismatch_tile_words(int32_t n_genomes);
int pgs_mismatch_tiles_device(pgs_engine *h, void **tiles_device_u32, int64_t *sites_out);

/* The path's only collective: genes shard across engines with no communication during the sweep;
 * every engine's partial counts are summed with an NCCL all-reduce (NVLink / NVSwitch).  NCCL is
 * loaded at run time (libnccl.so.2); without it these calls fail with PGS_ERR_STATE.
 *  - engines of ONE process, one per device (the CLI's --devices N): */
int pgs_mismatch_multi(pgs_engine *const *engines, int n_engines, uint32_t *out_host, int64_t *sites_out);
/*  - one engine per PROCESS (one rank per GPU, MPI-style launchers): rank 0 draws an id (128 bytes), the caller hands it
 *    to every rank, every rank joins; then every rank calls pgs_mismatch_allreduce together.
 *    out_host may be NULL on ranks that do not need the matrix.  *sites_out = THIS rank's sites. */
int pgs_comm_unique_id(void *id128);
int pgs_comm_init(pgs_engine *h, const void *id128, int rank, int world);
int pgs_mismatch_allreduce(pgs_engine *h, uint32_t *out_host, int64_t *sites_out, double *allreduce_ms);

/* ---- outputs -------------------------------------------------------------------------- */

/* Replaces: gene::to_fasta (gene.h:42-52) and the writers pangenomesim.cxx:117-157 (ref.fasta,
 * core.fasta, accessory.fasta, genomeNNN.fasta) and :236-293 (alignment.maf).  Text is produced
 * on the device (kernel K3) and streamed through pinned buffers. */
#define PGS_OUT_REF 1u
#define PGS_OUT_CORE 2u
#define PGS_OUT_ACCESSORY 4u
#define PGS_OUT_GENOMES 8u
#define PGS_OUT_MAF 16u
#define PGS_OUT_ALL 31u
#define PGS_OUT_DISCARD 32u /* produce, copy to pinned host memory, then drop (benchmarking) */
/* Not a format of the reference (optional): genomes.pgs2, the genomes' sequences
 * 2-bit packed -- an eighth of the FASTA text's bytes -- behind an index.  Little-endian:
 *   0  char[8]  "PGS2BIT\0"          8  uint32 version (1)      12  uint32 offset of the row data (0 if
 *      beyond 4 GiB: it is the end of the index below rounded up to a multiple of 256)
 *  16  int64 n_genomes   24  int64 gene_length   32  int64 blocks (a row = blocks * 16 bytes)
 *  40  int64 n_records   48  int64 n_dense (genes carried by every genome: their records lead
 *      every genome's rows, in the same order)
 *  56  int64 record_off[n_genomes + 1]   rows of genome i: [record_off[i], record_off[i+1]),
 *                                        n_dense dense records, then the genome's other records
 *      int64 dense_gene_id[n_dense], dense_key[n_dense]
 *      int64 other_gene_id[n_other], other_key[n_other]   n_other = n_records - n_genomes * n_dense;
 *                                        genome i's entries start at record_off[i] - i * n_dense
 *      (key = the gene's position in gene_ids() order: genomeNNN.fasta lists the genome's records
 *      by ascending key)
 *      zero padding up to the row data; then record r's row at data offset + r * blocks * 16:
 *      site s is bits 2(s%16)..2(s%16)+1 of 32-bit word s/16, A=0 C=1 G=2 T=3, padding sites 0.
 * Not part of PGS_OUT_ALL; not available with a shard layout.  pangenomesim_b200/packed.py reads it. */
#define PGS_OUT_PACKED 64u
/* genomeNNN.fasta and alignment.maf from PACKED rows: the rows cross PCIe 2-bit packed (a quarter
 * of the text's bytes) and the host's cores write the text -- headers and bases -- where it has to
 * end up (threaded, SIMD; PGS_EXPAND_THREADS overrides the thread count).  Same bytes as without
 * the flag.  With a sink: genomeNNN.fasta arrives whole in one call per file, from worker threads
 * (one call at a time, in no particular order across files); alignment.maf in consecutive pieces. */
#define PGS_OUT_HOST_EXPAND 128u

/* Sharded output: engines that each hold a contiguous slice of the gene table fill disjoint byte
 * ranges of the same genomeNNN.fasta / alignment.maf.  The caller creates (truncates) the files
 * -- best at their final size (the sum of pgs_output_sizes over the engines): the engines never
 * resize a shared file, and write a file that is already large enough through a mapping --,
 * asks every engine for its sizes (pgs_output_sizes), prefix-sums them over the engines and
 * passes the resulting offsets.  ref/core/accessory.fasta are then written by the caller from
 * pgs_copy_origin_ascii. */
typedef struct pgs_shard_layout {
	const int64_t *genome_offset;     /* [n_genomes] offset of this engine's first record in genomeNNN.fasta */
	int64_t maf_offset;               /* offset of this engine's first block in alignment.maf; 0 = writes the header */
	const int64_t *genome_gene_total; /* [n_genomes] genes carried by each genome over ALL engines (MAF size field) */
	int64_t n_ref_total;              /* reference genes over ALL engines (MAF size field of genomeref) */
} pgs_shard_layout;

typedef struct pgs_output_spec {
	const char *out_dir;      /* files are created in out_dir (must exist); NULL with a sink / DISCARD */
	uint32_t what;            /* PGS_OUT_* */
	int64_t n_ref_core;       /* reference: img_model::get_core(), img.cxx:381-384 */
	const int64_t *ref_core;  /* gene table indices */
	int64_t n_ref_acc;        /* reference: img_model::get_accessory(), img.cxx:386-389 */
	const int64_t *ref_acc;
	/* Optional sink: called with consecutive chunks of each file instead of writing out_dir/name.
	 * data is pinned host memory valid during the call.  Return non-zero to abort. */
	int (*sink)(void *user, const char *file_name, int64_t offset, const char *data, int64_t len);
	void *sink_user;
	const pgs_shard_layout *shard; /* NULL: this engine holds the whole gene table */
} pgs_output_spec;

typedef struct pgs_output_report {
	int64_t files, bytes, records;
	double device_ms; /* K3 kernel time */
	double wall_ms;   /* whole call */
	int64_t d2h_bytes; /* bytes copied device -> host by the call: the text (K3), or the packed rows behind it
	                    * (PGS_OUT_HOST_EXPAND: once per record and once per MAF line).  ABI 3 */
} pgs_output_report;

int pgs_write_outputs(pgs_engine *h, const pgs_output_spec *spec, pgs_output_report *report);

/* Bytes this engine will contribute to every genomeNNN.fasta (genome_bytes[n_genomes]) and to
 * alignment.maf (*maf_bytes, including the file header iff spec->shard is NULL or has maf_offset 0).
 * spec->shard may carry only genome_gene_total / n_ref_total at this point. */
int pgs_output_sizes(pgs_engine *h, const pgs_output_spec *spec, int64_t *genome_bytes, int64_t *maf_bytes);
/* The gene's sequence at its origin node (its genomeref record) as gene_length ASCII bases. */
int pgs_copy_origin_ascii(pgs_engine *h, int64_t gene_index, char *dst);

/* genome_name of the reference (util.cxx:15-27) with the >= 1000 genomes fix of SURVEY.md H5:
 * "genome" + zero-padded(3) (index+1), "genomeref" for -1.  Returns the length written. */
int pgs_genome_name(int64_t genome_index, char *dst, size_t cap);

/* ---- test hooks ----------------------------------------------------------------------- */

/* Copy the packed row of (node, gene table index) to dst (blocks_per_gene*16 bytes).
 * Returns 1 if the row does not exist (gene absent at node), PGS_ERR_STATE if the gene is present
 * at the node but the sweep does not keep that row (an internal node other than the gene's origin,
 * without PGS_FLAG_KEEP_INTERNAL). */
int pgs_copy_packed(pgs_engine *h, int32_t node, int64_t gene_index, void *dst);
/* Copy the dense-gene region of a node: n_dense_genes rows in ascending gene_id order. */
int pgs_copy_dense(pgs_engine *h, int32_t node, void *dst);
/* Row index of (node, gene) or -1. */
int64_t pgs_row_index(pgs_engine *h, int32_t node, int64_t gene_index);
/* The host-side expander of PGS_OUT_HOST_EXPAND on its own (no engine, no device): `sites` bases of a
 * packed row -> ASCII.  Returns the name of the code path in use ("avx2", "ssse3", "scalar"). */
const char *pgs_expand_bases(const void *packed_row, int64_t sites, char *dst);
/* Decision table of the branch above node: H[k] = floor(2^64 P(K >= k)), k = 0..64. */
int pgs_get_edge_table(pgs_engine *h, int32_t node, uint64_t H[65], int32_t *kmax);

#ifdef __cplusplus
}
#endif
#endif
