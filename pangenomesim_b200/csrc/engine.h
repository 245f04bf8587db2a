// engine.h -- host-side state of one engine (one CUDA device): tree, gene table, row plan,
// device tables.  Internal; the public surface is include/pgs.h.
#pragma once
#include "../../include/pgs.h"
#include "kernels.cuh"

#include <cstdint>
#include <string>
#include <vector>

namespace pgs {

struct DevBuf {
	void *p = nullptr;
	size_t bytes = 0;
	DevBuf() = default;
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	~DevBuf() { release(); }
	void release();
	template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct Engine {
	pgs_config cfg{};
	int device = 0;
	bool plan_only = false; // PGS_PLAN_ONLY: no device, planning only (development knob)
	cudaStream_t own_stream = nullptr, stream = nullptr;
	cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // 0-2 sweep, 3-4 mismatch, 5 format, 6 between the dense and the sparse genes
	std::string err;

	// ---- tree (pgs_set_tree)
	int32_t n_nodes = 0, n_genomes = 0, root = -1, max_level = 0;
	std::vector<int32_t> parent, leaf_genome, genome_leaf, level;
	std::vector<double> rate;
	std::vector<int32_t> child_off, child; // CSR children, ascending node index
	bool have_tree = false;

	// ---- genes (pgs_set_genes)
	int64_t n_genes = 0;
	std::vector<int64_t> gene_id;
	std::vector<int32_t> origin;
	std::vector<int64_t> carrier_off;  // CSR over carrier_genome, sparse genes only expanded
	std::vector<int32_t> carrier_genome;
	std::vector<int32_t> all_order;    // MAF order of genomes for dense genes
	std::vector<int32_t> dense_slot;   // per gene: slot in the dense region or -1
	std::vector<uint32_t> dense_gid;   // per dense slot: gene_id (ascending)
	std::vector<int64_t> dense_gene;   // per dense slot: gene table index
	std::vector<int64_t> sp_off;       // [N+1] CSR: sparse genes present at node (ascending index)
	std::vector<int64_t> sp_gene;
	std::vector<int64_t> row_base;     // [N]
	std::vector<int64_t> genome_gene_count; // [n_genomes] genes carried
	bool have_genes = false, swept = false;

	// ---- plan
	int64_t blocks = 0;       // 64-site blocks per gene
	int32_t last_valid = 64;  // real sites in the last block
	int64_t rows_total = 0, rows_written = 0, rows_read = 0, rows_origin = 0;
	std::vector<int64_t> dense_level_off;  // [max_level+2] into dense tasks (index = child level)
	std::vector<int64_t> sparse_level_off; // [max_level+2]
	int64_t n_dense_tasks = 0, n_sparse_tasks = 0, n_root_tasks = 0;
	// fused column walk of the dense region (default for binary trees)
	bool walk = false;
	std::vector<int32_t> walk_pass_units; // units by kind (crown, rest of the top, subtrees), in launch order
	int32_t n_walk_units = 0, n_top_tasks = 0;
	bool sparse_walk = false;             // sparse genes: fused walk (k1_walk_sparse); internal sparse rows are not kept
	int64_t n_sp_units = 0, sparse_rows_written = 0, sparse_rows_read = 0; // (W and R of the sparse region)
	int walk_variant = 2;                 // kernel of the pass over the subtrees (choose_walk_variant)
	int32_t walk_smem_levels = 0;         // deepest shared-memory stack level used by any step, plus one
	int64_t rows_parked_smem = 0;         // rows per sweep parked in shared memory instead of the global pool
	int64_t rows_stored = 0, rows_loaded = 0;
	std::vector<char> materialised; // [N] dense region of the node holds valid rows after a sweep
	std::vector<uint64_t> tab_host; // concatenated H[1..kmax]
	std::vector<int32_t> tab_off_host, kmax_host;
	std::vector<uint64_t> h1_host;

	// ---- device
	DevBuf d_rows, d_row_base, d_h1, d_tab, d_tab_off, d_kmax, d_dense_gid;
	DevBuf d_dense_tasks, d_sparse_tasks, d_root_tasks, d_leaf_base_words, d_mismatch, d_tab8, d_walk_steps, d_walk_off, d_unit_path_off, d_unit_path, d_unit_store, d_top_tasks, d_sp_units, d_sp_steps, d_sp_paths, d_planes, d_tiles;
	char *h_mm_stage[2] = {nullptr, nullptr}; // pinned staging of pgs_mismatch's result
	size_t mm_stage_bytes = 0;
	cudaStream_t aux_stream = nullptr; // K2 on the tensor cores: the operand expansion runs beside the tile kernel
	cudaEvent_t mm_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
	void *comm = nullptr;       // ncclComm_t of pgs_comm_init (one engine per process), else nullptr
	int comm_world = 1, comm_rank = 0;
	// K3's tables and per-chunk index buffers: kept between pgs_write_outputs calls (they change with
	// pgs_set_genes and with the caller's totals only), so that a call allocates nothing after the first
	DevBuf d_fmt[16], d_fmt_aux_a[2], d_fmt_aux_b[2];
	uint64_t fmt_key = 0;       // what d_fmt holds: plan serial + the spec's totals (0: nothing)
	uint64_t plan_serial = 0;   // bumped by every pgs_set_genes
	DevBuf d_stage[2];          // K3 staging, kept between pgs_write_outputs calls
	char *h_stage[2] = {nullptr, nullptr}; // pinned
	int64_t stage_bytes = 0;
	DeviceTables tables{};
	cudaGraph_t graph = nullptr;
	cudaGraphExec_t graph_exec = nullptr;
	int64_t launches_last_sweep = 0, sweeps_done = 0;
	double mismatch_ms = 0.0, format_ms = 0.0, dense_ms = 0.0, sparse_ms = 0.0;
	bool phase_split = false; // the last sweep recorded ev[6]

	~Engine();
	int fail(int code, const std::string &msg);
	int cuda_fail(cudaError_t e, const char *what);
	int alloc(DevBuf &b, size_t bytes, const char *what);
	int upload(DevBuf &b, const void *src, size_t bytes, const char *what);
	int upload_keep(DevBuf &b, const void *src, size_t bytes, const char *what); // grow-only, no synchronisation
	int64_t device_bytes() const;
	int64_t row_index(int32_t node, int64_t gene_index) const;
	bool row_kept(int32_t node, int64_t gene_index) const;
	void drop_graph();
	void drop_comm();
	void enqueue_levels(cudaStream_t s, int64_t *launches, cudaEvent_t after_dense);
};

// plan.cu: everything behind pgs_set_genes (validation, layout, tasks, tables, uploads)
int plan_genes(Engine &e, int64_t n_genes, const int64_t *gene_id, const int32_t *origin_node,
			   const int64_t *carrier_off, const int32_t *carrier_genome, const int32_t *all_order);

// mismatch.cu: K2 and the all-reduce of its partial results
int64_t mismatch_tile_words(int32_t n);
int mismatch_tiles(Engine &e, int64_t *sites_out);
int mismatch_impl(const Engine &e);
int mismatch_tiles_tensor(Engine &e, int impl); // mismatch_tc.cu: tcgen05 path into e.d_tiles (allocated and cleared by mismatch_tiles)
int mismatch_square_device(Engine &e, uint32_t *out_device);
int mismatch_tiles_to_host(Engine &e, uint32_t *out_host);
int comm_unique_id(void *id128);
int comm_init(Engine &e, const void *id128, int rank, int world);
int allreduce_tiles(Engine &e);
int mismatch_multi(Engine *const *engines, int n_engines, uint32_t *out_host, int64_t *sites_out);

// format.cu
int write_outputs(Engine &e, const pgs_output_spec *spec, pgs_output_report *report);
int output_sizes(Engine &e, const pgs_output_spec *spec, int64_t *genome_bytes, int64_t *maf_bytes);
int copy_origin_ascii(Engine &e, int64_t gene_index, char *dst);
int format_genome_name(int64_t genome_index, char *dst, size_t cap);

} // namespace pgs

struct pgs_engine {
	pgs::Engine e;
};
