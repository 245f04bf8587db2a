// engine.cu -- C ABI (include/pgs.h): planning of the row layout and of the level sweep, device
// tables, launches.  There is no CPU path: every compute entry point ends in a CUDA kernel of
// kernels.cu / format.cu and fails with PGS_ERR_NO_DEVICE / PGS_ERR_CUDA otherwise.
#include "engine.h"
#include "expand.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>

namespace pgs {

static thread_local std::string g_create_error;

void DevBuf::release()
{
	if (p) cudaFree(p);
	p = nullptr;
	bytes = 0;
}

Engine::~Engine()
{
	drop_graph();
	drop_comm();
	for (auto &p : h_stage)
		if (p) cudaFreeHost(p);
	for (auto &p : h_mm_stage)
		if (p) cudaFreeHost(p);
	for (auto &x : ev)
		if (x) cudaEventDestroy(x);
	for (auto &x : mm_ev)
		if (x) cudaEventDestroy(x);
	if (aux_stream) cudaStreamDestroy(aux_stream);
	if (own_stream) cudaStreamDestroy(own_stream);
}

int Engine::fail(int code, const std::string &msg)
{
	err = msg;
	return code;
}

int Engine::cuda_fail(cudaError_t e, const char *what)
{
	err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
	return e == cudaErrorMemoryAllocation ? PGS_ERR_NOMEM : PGS_ERR_CUDA;
}

int Engine::alloc(DevBuf &b, size_t bytes, const char *what)
{
	if (plan_only) return fail(PGS_ERR_NO_DEVICE, std::string("plan-only engine (PGS_PLAN_ONLY): no device memory for ") + what);
	b.release();
	if (bytes == 0) bytes = 16;
	cudaError_t e = cudaMalloc(&b.p, bytes);
	if (e != cudaSuccess) {
		b.p = nullptr;
		cudaGetLastError();
		return cuda_fail(e, what);
	}
	b.bytes = bytes;
	return PGS_OK;
}

int Engine::upload(DevBuf &b, const void *src, size_t bytes, const char *what)
{
	// a buffer that fits without being more than twice too large is used again: cudaFree / cudaMalloc synchronise
	// the device, and a re-plan (every pgs_set_genes) replaces some twenty tables of unchanged size
	if (!b.p || b.bytes < bytes || b.bytes > 2 * bytes + ((size_t)1 << 20) || plan_only) {
		const int rc = alloc(b, bytes, what);
		if (rc) return rc;
	}
	if (bytes) {
		cudaError_t e = cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
		if (e != cudaSuccess) return cuda_fail(e, what);
	}
	return PGS_OK;
}

// The buffer is only replaced when it is too small; the copy is queued on the engine's stream (the
// source is pageable: the runtime has taken its copy when the call returns).
int Engine::upload_keep(DevBuf &b, const void *src, size_t bytes, const char *what)
{
	if (b.bytes < bytes || !b.p) {
		const int rc = alloc(b, bytes, what);
		if (rc) return rc;
	}
	if (bytes) {
		const cudaError_t e = cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, stream);
		if (e != cudaSuccess) return cuda_fail(e, what);
	}
	return PGS_OK;
}

int64_t Engine::device_bytes() const
{
	const DevBuf *all[] = {&d_rows, &d_row_base, &d_h1, &d_tab, &d_tab_off, &d_kmax, &d_dense_gid,
						   &d_dense_tasks, &d_sparse_tasks, &d_root_tasks, &d_leaf_base_words,
						   &d_tab8, &d_walk_steps, &d_walk_off, &d_unit_path_off, &d_unit_path, &d_unit_store, &d_top_tasks, &d_sp_units, &d_sp_steps, &d_sp_paths, &d_planes, &d_tiles};
	int64_t s = 0;
	for (auto b : all) s += (int64_t)b->bytes;
	return s;
}

void Engine::drop_graph()
{
	if (graph_exec) cudaGraphExecDestroy(graph_exec);
	if (graph) cudaGraphDestroy(graph);
	graph_exec = nullptr;
	graph = nullptr;
}

int64_t Engine::row_index(int32_t node, int64_t g) const
{
	if (!have_genes || node < 0 || node >= n_nodes || g < 0 || g >= n_genes) return -1;
	if (dense_slot[g] >= 0) return materialised[node] ? row_base[node] + dense_slot[g] : -1;
	const int64_t *b = sp_gene.data() + sp_off[node], *e = sp_gene.data() + sp_off[node + 1];
	const int64_t *it = std::lower_bound(b, e, g);
	if (it == e || *it != g) return -1;
	return row_base[node] + (materialised[node] ? (int64_t)dense_gid.size() : 0) + (it - b);
}

// Does the sweep leave a valid row for (node, gene) in memory?  The fused walks keep only what the
// outputs read: genomes' rows and origin rows.  (True also where the gene is absent: row_index says so.)
bool Engine::row_kept(int32_t node, int64_t g) const
{
	if (dense_slot[g] >= 0) return materialised[node] != 0;
	if (!sparse_walk) return true;
	return child_off[node + 1] == child_off[node] || node == origin[g] || row_index(node, g) < 0;
}

// K1 for every level below the root: dense region, then sparse region.  Levels are separated by
// stream order (each launch reads rows written by the previous level's launches).
void Engine::enqueue_levels(cudaStream_t s, int64_t *launches, cudaEvent_t after_dense)
{
	const DenseTask *dt = d_dense_tasks.as<DenseTask>();
	const SparseTask *st = d_sparse_tasks.as<SparseTask>();
	if (walk) { // dense genes: origin rows and the masks of the top branches, then ONE launch for every unit of the tree
		launch_top_masks(tables, d_top_tasks.as<TopMaskTask>(), n_top_tasks, s);
		(*launches)++;
		WalkUnits u{};
		u.steps = d_walk_steps.as<WalkStep>();
		u.unit_off = d_walk_off.as<int64_t>();
		u.unit_path_off = d_unit_path_off.as<int32_t>();
		u.unit_path = d_unit_path.as<int64_t>();
		u.unit_store = d_unit_store.as<int64_t>();
		u.root_off = row_base[root] * blocks;
		u.n_units = n_walk_units;
		launch_walk_dense(tables, u, walk_smem_levels, walk_variant, s);
		(*launches)++;
	}
	if (after_dense && (walk || tables.n_dense == 0)) cudaEventRecord(after_dense, s); // (plain launches only: splits K1's time)
	if (sparse_walk) { // sparse genes: origin rows and every branch in one launch
		launch_walk_sparse(tables, d_sp_units.as<SparseUnit>(), d_sp_steps.as<SparseStep>(), d_sp_paths.as<SparsePathEdge>(), n_sp_units, s);
		if (n_sp_units > 0) (*launches)++;
	}
	for (int32_t lvl = 1; lvl <= max_level; lvl++) {
		const int64_t nd = dense_level_off[lvl + 1] - dense_level_off[lvl];
		if (nd > 0 && tables.n_dense > 0 && !walk) {
			launch_sweep_dense(tables, dt + dense_level_off[lvl], (int32_t)nd, s);
			(*launches)++;
		}
		const int64_t ns = sparse_level_off[lvl + 1] - sparse_level_off[lvl];
		if (ns > 0 && !sparse_walk) {
			launch_sweep_sparse(tables, st + sparse_level_off[lvl], ns, s);
			(*launches)++;
		}
	}
}

} // namespace pgs

using pgs::Engine;

#define ENGINE_OR_RETURN(h)                \
	if (!(h)) return PGS_ERR_ARG;          \
	Engine &e = (h)->e;                    \
	if (!e.plan_only) {                    \
		cudaError_t ce_ = cudaSetDevice(e.device); \
		if (ce_ != cudaSuccess) return e.cuda_fail(ce_, "cudaSetDevice"); \
	}

extern "C" {

int pgs_abi_version(void)
{
	return PGS_ABI_VERSION;
}

int pgs_device_count(void)
{
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return count;
}

const char *pgs_last_error(const pgs_engine *h)
{
	return h ? h->e.err.c_str() : pgs::g_create_error.c_str();
}

int pgs_create(const pgs_config *cfg, pgs_engine **out)
{
	if (!cfg || !out) {
		pgs::g_create_error = "pgs_create: null argument";
		return PGS_ERR_ARG;
	}
	*out = nullptr;
	if (cfg->gene_length < 1 || cfg->gene_length > (int64_t)1 << 36) {
		pgs::g_create_error = "pgs_create: gene_length out of range";
		return PGS_ERR_ARG;
	}
	// Development knob: an engine without a device that can only PLAN (pgs_set_tree / pgs_set_genes up to
	// the first device allocation, where it fails with PGS_ERR_NO_DEVICE) -- for timing and testing the
	// host-side planner on machines without a GPU.  It computes nothing: there is no CPU path.
	if (const char *v = std::getenv("PGS_PLAN_ONLY"))
		if (std::atoi(v)) {
			pgs_engine *h = new (std::nothrow) pgs_engine();
			if (!h) return PGS_ERR_NOMEM;
			h->e.cfg = *cfg;
			h->e.plan_only = true;
			h->e.blocks = (cfg->gene_length + 63) / 64;
			h->e.last_valid = (int32_t)(cfg->gene_length - (h->e.blocks - 1) * 64);
			*out = h;
			return PGS_OK;
		}
	int count = 0;
	cudaError_t ce = cudaGetDeviceCount(&count);
	if (ce != cudaSuccess || count == 0) {
		cudaGetLastError();
		pgs::g_create_error = std::string("pgs_create: no usable CUDA device (") +
							  (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0") +
							  "); this engine has no CPU path";
		return PGS_ERR_NO_DEVICE;
	}
	int dev = cfg->device;
	if (dev < 0) {
		ce = cudaGetDevice(&dev);
		if (ce != cudaSuccess) dev = 0;
	}
	if (dev >= count) {
		pgs::g_create_error = "pgs_create: device ordinal out of range";
		return PGS_ERR_ARG;
	}
	pgs_engine *h = new (std::nothrow) pgs_engine();
	if (!h) {
		pgs::g_create_error = "pgs_create: out of host memory";
		return PGS_ERR_NOMEM;
	}
	Engine &e = h->e;
	e.cfg = *cfg;
	e.device = dev;
	ce = cudaSetDevice(dev);
	if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e.own_stream, cudaStreamNonBlocking);
	for (int i = 0; i < 7 && ce == cudaSuccess; i++) ce = cudaEventCreate(&e.ev[i]);
	if (ce != cudaSuccess) {
		pgs::g_create_error = std::string("pgs_create: ") + cudaGetErrorString(ce);
		delete h;
		return PGS_ERR_CUDA;
	}
	e.stream = e.own_stream;
	if (const char *v = std::getenv("PGS_K1_VARIANT")) pgs::g_dense_variant = std::atoi(v);
	if (const char *v = std::getenv("PGS_WALK_VARIANT")) pgs::g_walk_variant = std::atoi(v);
	if (const char *v = std::getenv("PGS_NO_GRAPH")) // development knob: plain launches (profiling, first-sweep timing)
		if (std::atoi(v)) e.cfg.flags |= PGS_FLAG_NO_GRAPH;
	if (const char *v = std::getenv("PGS_KEEP_INTERNAL")) // development knob: force the level-synchronous sweep
		if (std::atoi(v)) e.cfg.flags |= PGS_FLAG_KEEP_INTERNAL;
	e.blocks = (cfg->gene_length + 63) / 64;
	e.last_valid = (int32_t)(cfg->gene_length - (e.blocks - 1) * 64);
	*out = h;
	return PGS_OK;
}

void pgs_destroy(pgs_engine *h)
{
	if (!h) return;
	if (!h->e.plan_only) {
		cudaSetDevice(h->e.device);
		cudaStreamSynchronize(h->e.stream);
	}
	delete h;
}

int pgs_set_stream(pgs_engine *h, void *cuda_stream)
{
	ENGINE_OR_RETURN(h);
	cudaStreamSynchronize(e.stream);
	e.stream = cuda_stream ? (cudaStream_t)cuda_stream : e.own_stream;
	return PGS_OK;
}

int pgs_sync(pgs_engine *h)
{
	ENGINE_OR_RETURN(h);
	cudaError_t ce = cudaStreamSynchronize(e.stream);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_sync");
	return PGS_OK;
}

int pgs_set_tree(pgs_engine *h, int32_t n_nodes, const int32_t *parent, const double *rate,
				 const int32_t *leaf_genome)
{
	ENGINE_OR_RETURN(h);
	if (n_nodes < 1 || !parent || !rate || !leaf_genome) return e.fail(PGS_ERR_ARG, "pgs_set_tree: null or empty");
	e.have_tree = e.have_genes = e.swept = false;
	e.drop_graph();

	int32_t root = -1;
	std::vector<int32_t> deg(n_nodes + 1, 0);
	for (int32_t v = 0; v < n_nodes; v++) {
		const int32_t p = parent[v];
		if (p == -1) {
			if (root != -1) return e.fail(PGS_ERR_ARG, "pgs_set_tree: more than one root");
			root = v;
		} else if (p < 0 || p >= n_nodes || p == v) {
			return e.fail(PGS_ERR_ARG, "pgs_set_tree: parent index out of range");
		} else {
			deg[p]++;
		}
		if (p != -1 && !(rate[v] >= 0.0)) return e.fail(PGS_ERR_ARG, "pgs_set_tree: negative or NaN rate");
	}
	if (root == -1) return e.fail(PGS_ERR_ARG, "pgs_set_tree: no root");

	e.child_off.assign(n_nodes + 1, 0);
	for (int32_t v = 0; v < n_nodes; v++) e.child_off[v + 1] = e.child_off[v] + deg[v];
	e.child.assign(e.child_off[n_nodes], 0);
	{
		std::vector<int32_t> fill(e.child_off.begin(), e.child_off.end() - 1);
		for (int32_t v = 0; v < n_nodes; v++)
			if (parent[v] >= 0) e.child[fill[parent[v]]++] = v;
	}
	// levels by breadth-first walk from the root; unreachable nodes mean a cycle
	e.level.assign(n_nodes, -1);
	std::vector<int32_t> queue;
	queue.reserve(n_nodes);
	queue.push_back(root);
	e.level[root] = 0;
	int32_t max_level = 0;
	for (size_t qi = 0; qi < queue.size(); qi++) {
		const int32_t v = queue[qi];
		for (int32_t c = e.child_off[v]; c < e.child_off[v + 1]; c++) {
			const int32_t u = e.child[c];
			e.level[u] = e.level[v] + 1;
			max_level = std::max(max_level, e.level[u]);
			queue.push_back(u);
		}
	}
	if ((int32_t)queue.size() != n_nodes) return e.fail(PGS_ERR_ARG, "pgs_set_tree: cycle or disconnected node");

	int32_t n_genomes = 0;
	for (int32_t v = 0; v < n_nodes; v++) {
		const int32_t g = leaf_genome[v];
		if (g < -1) return e.fail(PGS_ERR_ARG, "pgs_set_tree: leaf_genome < -1");
		if (g >= 0) {
			if (deg[v] != 0) return e.fail(PGS_ERR_ARG, "pgs_set_tree: genome assigned to an internal node");
			n_genomes = std::max(n_genomes, g + 1);
		}
	}
	e.genome_leaf.assign(n_genomes, -1);
	for (int32_t v = 0; v < n_nodes; v++) {
		const int32_t g = leaf_genome[v];
		if (g >= 0) {
			if (e.genome_leaf[g] != -1) return e.fail(PGS_ERR_ARG, "pgs_set_tree: duplicate genome index");
			e.genome_leaf[g] = v;
		}
	}
	for (int32_t g = 0; g < n_genomes; g++)
		if (e.genome_leaf[g] < 0) return e.fail(PGS_ERR_ARG, "pgs_set_tree: genome indices are not 0..n-1");

	e.n_nodes = n_nodes;
	e.n_genomes = n_genomes;
	e.root = root;
	e.max_level = max_level;
	e.parent.assign(parent, parent + n_nodes);
	e.rate.assign(rate, rate + n_nodes);
	e.leaf_genome.assign(leaf_genome, leaf_genome + n_nodes);
	e.have_tree = true;
	return PGS_OK;
}

int pgs_set_genes(pgs_engine *h, int64_t n_genes, const int64_t *gene_id, const int32_t *origin_node,
				  const int64_t *carrier_off, const int32_t *carrier_genome, const int32_t *all_order)
{
	ENGINE_OR_RETURN(h);
	if (!e.have_tree) return e.fail(PGS_ERR_STATE, "pgs_set_genes: call pgs_set_tree first");
	if (n_genes < 0 || (n_genes > 0 && (!gene_id || !origin_node || !carrier_off || !carrier_genome)))
		return e.fail(PGS_ERR_ARG, "pgs_set_genes: null table");
	e.have_genes = e.swept = false;
	e.sweeps_done = 0;
	e.plan_serial++;
	e.fmt_key = 0;
	e.drop_graph();
	const int rc = pgs::plan_genes(e, n_genes, gene_id, origin_node, carrier_off, carrier_genome, all_order); // plan.cu
	if (rc) return rc;
	e.have_genes = true;
	return PGS_OK;
}

int pgs_sweep(pgs_engine *h)
{
	ENGINE_OR_RETURN(h);
	if (!e.have_genes) return e.fail(PGS_ERR_STATE, "pgs_sweep: call pgs_set_tree and pgs_set_genes first");
	cudaError_t ce;
	int64_t launches = 0;
	cudaEventRecord(e.ev[0], e.stream);
	pgs::launch_rootgen(e.tables, e.d_root_tasks.as<pgs::RootTask>(), e.n_root_tasks, e.stream);
	if (e.n_root_tasks) launches++;
	cudaEventRecord(e.ev[1], e.stream);
	// The first sweep of a plan is launched kernel by kernel (capturing and instantiating a graph
	// costs more than it saves for a single sweep, e.g. the CLI); repeated sweeps replay a graph.
	const bool plain = (e.cfg.flags & PGS_FLAG_NO_GRAPH) || e.sweeps_done == 0;
	e.sweeps_done++;
	e.phase_split = false;
	if (plain) {
		e.enqueue_levels(e.stream, &launches, e.ev[6]);
		e.phase_split = e.walk || e.tables.n_dense == 0;
	} else {
		if (!e.graph_exec) {
			// capture the level launches once; replays cost one graph launch per sweep
			int64_t captured = 0;
			ce = cudaStreamBeginCapture(e.stream, cudaStreamCaptureModeThreadLocal);
			if (ce != cudaSuccess) return e.cuda_fail(ce, "cudaStreamBeginCapture");
			e.enqueue_levels(e.stream, &captured, nullptr);
			ce = cudaStreamEndCapture(e.stream, &e.graph);
			if (ce != cudaSuccess) return e.cuda_fail(ce, "cudaStreamEndCapture");
			ce = cudaGraphInstantiate(&e.graph_exec, e.graph, 0);
			if (ce != cudaSuccess) return e.cuda_fail(ce, "cudaGraphInstantiate");
			e.launches_last_sweep = captured;
		}
		launches += e.launches_last_sweep;
		ce = cudaGraphLaunch(e.graph_exec, e.stream);
		if (ce != cudaSuccess) return e.cuda_fail(ce, "cudaGraphLaunch");
	}
	cudaEventRecord(e.ev[2], e.stream);
	ce = cudaGetLastError();
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_sweep launch");
	if (plain) e.launches_last_sweep = launches - (e.n_root_tasks ? 1 : 0);
	(void)launches;
	e.swept = true;
	return PGS_OK;
}

int pgs_get_stats(pgs_engine *h, pgs_stats *out)
{
	ENGINE_OR_RETURN(h);
	if (!out) return e.fail(PGS_ERR_ARG, "pgs_get_stats: null");
	std::memset(out, 0, sizeof(*out));
	out->n_nodes = e.n_nodes;
	out->n_genomes = e.n_genomes;
	out->blocks_per_gene = e.blocks;
	if (e.have_genes) {
		out->n_genes = e.n_genes;
		out->n_dense_genes = (int64_t)e.dense_gid.size();
		out->rows_total = e.rows_total;
		out->rows_written = e.rows_written;
		out->rows_read = e.rows_read;
		out->rows_origin = e.rows_origin;
		int64_t leaf_genes = 0;
		for (auto c : e.genome_gene_count) leaf_genes += c;
		out->leaf_nucleotides = leaf_genes * e.cfg.gene_length;
		out->site_edges = e.rows_written * e.cfg.gene_length;
		out->levels = e.max_level;
		out->kernel_launches = e.launches_last_sweep + (e.n_root_tasks ? 1 : 0);
		out->device_bytes = e.device_bytes();
		out->sweep_mode = e.walk ? 1 : 0;
		out->rows_stored = e.rows_stored;
		out->rows_loaded = e.rows_loaded;
	}
	if (e.swept) {
		cudaError_t ce = cudaStreamSynchronize(e.stream);
		if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_get_stats sync");
		float a = 0, b = 0;
		if (cudaEventElapsedTime(&a, e.ev[0], e.ev[1]) == cudaSuccess) out->rootgen_ms = a;
		if (cudaEventElapsedTime(&b, e.ev[1], e.ev[2]) == cudaSuccess) out->sweep_ms = b;
		if (e.phase_split) {
			float c = 0, d = 0;
			if (cudaEventElapsedTime(&c, e.ev[1], e.ev[6]) == cudaSuccess) e.dense_ms = c;
			if (cudaEventElapsedTime(&d, e.ev[6], e.ev[2]) == cudaSuccess) e.sparse_ms = d;
		}
		cudaGetLastError();
	}
	out->mismatch_ms = e.mismatch_ms;
	out->format_ms = e.format_ms;
	out->dense_ms = e.dense_ms;
	out->sparse_ms = e.sparse_ms;
	return PGS_OK;
}

// K2 on the engine's stream, timed; the packed tiles stay in e.d_tiles
static int run_mismatch(Engine &e, int64_t *sites_out)
{
	if (!e.swept) return e.fail(PGS_ERR_STATE, "pgs_mismatch: call pgs_sweep first");
	cudaEventRecord(e.ev[3], e.stream);
	const int rc = pgs::mismatch_tiles(e, sites_out);
	if (rc) return rc;
	cudaEventRecord(e.ev[4], e.stream);
	cudaError_t ce = cudaStreamSynchronize(e.stream);
	if (ce == cudaSuccess) ce = cudaGetLastError();
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch");
	float ms = 0;
	cudaEventElapsedTime(&ms, e.ev[3], e.ev[4]);
	e.mismatch_ms = ms;
	return PGS_OK;
}

int pgs_mismatch_device(pgs_engine *h, void *out_device_u32, int64_t *sites_out)
{
	ENGINE_OR_RETURN(h);
	if (!out_device_u32) return e.fail(PGS_ERR_ARG, "pgs_mismatch: null output");
	int rc = run_mismatch(e, sites_out);
	if (rc) return rc;
	if ((rc = pgs::mismatch_square_device(e, static_cast<uint32_t *>(out_device_u32)))) return rc;
	const cudaError_t ce = cudaStreamSynchronize(e.stream);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch");
	return PGS_OK;
}

int pgs_mismatch(pgs_engine *h, uint32_t *out_host, int64_t *sites_out)
{
	ENGINE_OR_RETURN(h);
	if (!out_host) return e.fail(PGS_ERR_ARG, "pgs_mismatch: null output");
	const int rc = run_mismatch(e, sites_out);
	if (rc) return rc;
	return pgs::mismatch_tiles_to_host(e, out_host);
}

int64_t pgs_mismatch_tile_words(int32_t n_genomes)
{
	return n_genomes < 0 ? 0 : pgs::mismatch_tile_words(n_genomes);
}

int pgs_mismatch_tiles_device(pgs_engine *h, void **tiles_device_u32, int64_t *sites_out)
{
	ENGINE_OR_RETURN(h);
	const int rc = run_mismatch(e, sites_out);
	if (rc) return rc;
	if (tiles_device_u32) *tiles_device_u32 = e.d_tiles.p;
	return PGS_OK;
}

int pgs_comm_unique_id(void *id128)
{
	return pgs::comm_unique_id(id128);
}

int pgs_comm_init(pgs_engine *h, const void *id128, int rank, int world)
{
	ENGINE_OR_RETURN(h);
	return pgs::comm_init(e, id128, rank, world);
}

int pgs_mismatch_allreduce(pgs_engine *h, uint32_t *out_host, int64_t *sites_out, double *allreduce_ms)
{
	ENGINE_OR_RETURN(h);
	int64_t sites = 0;
	int rc = run_mismatch(e, &sites);
	if (rc) return rc;
	cudaEventRecord(e.ev[3], e.stream);
	if ((rc = pgs::allreduce_tiles(e))) return rc;
	cudaEventRecord(e.ev[4], e.stream);
	const cudaError_t ce = cudaStreamSynchronize(e.stream);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch_allreduce");
	float ms = 0;
	cudaEventElapsedTime(&ms, e.ev[3], e.ev[4]);
	if (allreduce_ms) *allreduce_ms = ms;
	if (sites_out) *sites_out = sites; // this rank's sites: the caller sums them (it knows how the genes were split)
	return out_host ? pgs::mismatch_tiles_to_host(e, out_host) : PGS_OK;
}

int pgs_mismatch_multi(pgs_engine *const *engines, int n_engines, uint32_t *out_host, int64_t *sites_out)
{
	if (!engines || n_engines < 1 || !engines[0]) return PGS_ERR_ARG;
	std::vector<Engine *> es(n_engines);
	for (int d = 0; d < n_engines; d++) {
		if (!engines[d]) return PGS_ERR_ARG;
		es[d] = &engines[d]->e;
		if (!es[d]->swept) return es[0]->fail(PGS_ERR_STATE, "pgs_mismatch_multi: call pgs_sweep on every engine first");
	}
	return pgs::mismatch_multi(es.data(), n_engines, out_host, sites_out);
}

int pgs_write_outputs(pgs_engine *h, const pgs_output_spec *spec, pgs_output_report *report)
{
	ENGINE_OR_RETURN(h);
	if (!e.swept) return e.fail(PGS_ERR_STATE, "pgs_write_outputs: call pgs_sweep first");
	if (!spec) return e.fail(PGS_ERR_ARG, "pgs_write_outputs: null spec");
	return pgs::write_outputs(e, spec, report);
}

int pgs_output_sizes(pgs_engine *h, const pgs_output_spec *spec, int64_t *genome_bytes, int64_t *maf_bytes)
{
	ENGINE_OR_RETURN(h);
	if (!e.have_genes) return e.fail(PGS_ERR_STATE, "pgs_output_sizes: call pgs_set_genes first");
	if (!spec) return e.fail(PGS_ERR_ARG, "pgs_output_sizes: null spec");
	if (spec->shard && !spec->shard->genome_gene_total) return e.fail(PGS_ERR_ARG, "pgs_output_sizes: shard layout without totals");
	pgs_shard_layout tmp;
	pgs_output_spec local = *spec;
	std::vector<int64_t> zeros;
	if (spec->shard && !spec->shard->genome_offset) { // offsets are not known yet: that is what the caller is computing
		tmp = *spec->shard;
		zeros.assign(e.n_genomes, 0);
		tmp.genome_offset = zeros.data();
		local.shard = &tmp;
	}
	return pgs::output_sizes(e, &local, genome_bytes, maf_bytes);
}

int pgs_copy_origin_ascii(pgs_engine *h, int64_t gene_index, char *dst)
{
	ENGINE_OR_RETURN(h);
	if (!e.swept) return e.fail(PGS_ERR_STATE, "pgs_copy_origin_ascii: call pgs_sweep first");
	return pgs::copy_origin_ascii(e, gene_index, dst);
}

int pgs_genome_name(int64_t genome_index, char *dst, size_t cap)
{
	return pgs::format_genome_name(genome_index, dst, cap);
}

int64_t pgs_row_index(pgs_engine *h, int32_t node, int64_t gene_index)
{
	if (!h) return -1;
	return h->e.row_index(node, gene_index);
}

int pgs_copy_packed(pgs_engine *h, int32_t node, int64_t gene_index, void *dst)
{
	ENGINE_OR_RETURN(h);
	if (!e.swept) return e.fail(PGS_ERR_STATE, "pgs_copy_packed: call pgs_sweep first");
	if (!dst) return e.fail(PGS_ERR_ARG, "pgs_copy_packed: null dst");
	if (node < 0 || node >= e.n_nodes || gene_index < 0 || gene_index >= e.n_genes)
		return e.fail(PGS_ERR_ARG, "pgs_copy_packed: index out of range");
	if (!e.row_kept(node, gene_index))
		return e.fail(PGS_ERR_STATE, "pgs_copy_packed: this internal row is not kept (create the engine with PGS_FLAG_KEEP_INTERNAL)");
	const int64_t row = e.row_index(node, gene_index);
	if (row < 0) return 1;
	cudaError_t ce = cudaStreamSynchronize(e.stream);
	if (ce == cudaSuccess)
		ce = cudaMemcpy(dst, e.tables.rows + row * e.blocks, (size_t)e.blocks * 16, cudaMemcpyDeviceToHost);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_copy_packed");
	return PGS_OK;
}

int pgs_copy_dense(pgs_engine *h, int32_t node, void *dst)
{
	ENGINE_OR_RETURN(h);
	if (!e.swept) return e.fail(PGS_ERR_STATE, "pgs_copy_dense: call pgs_sweep first");
	if (!dst || node < 0 || node >= e.n_nodes) return e.fail(PGS_ERR_ARG, "pgs_copy_dense: bad argument");
	if (!e.materialised[node])
		return e.fail(PGS_ERR_STATE, "pgs_copy_dense: this internal node is not kept (create the engine with PGS_FLAG_KEEP_INTERNAL)");
	cudaError_t ce = cudaStreamSynchronize(e.stream);
	if (ce == cudaSuccess)
		ce = cudaMemcpy(dst, e.tables.rows + e.row_base[node] * e.blocks,
						(size_t)e.dense_gid.size() * (size_t)e.blocks * 16, cudaMemcpyDeviceToHost);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_copy_dense");
	return PGS_OK;
}

const char *pgs_expand_bases(const void *packed_row, int64_t sites, char *dst)
{
	if (packed_row && dst && sites > 0) pgs::expand_bases(packed_row, sites, dst);
	return pgs::expand_kind();
}

int pgs_get_edge_table(pgs_engine *h, int32_t node, uint64_t H[65], int32_t *kmax)
{
	if (!h) return PGS_ERR_ARG;
	Engine &e = h->e;
	if (!e.have_genes) return e.fail(PGS_ERR_STATE, "pgs_get_edge_table: call pgs_set_genes first");
	if (node < 0 || node >= e.n_nodes || !H) return e.fail(PGS_ERR_ARG, "pgs_get_edge_table: bad argument");
	H[0] = std::numeric_limits<uint64_t>::max();
	for (int k = 1; k <= 64; k++) H[k] = k <= e.kmax_host[node] ? e.tab_host[e.tab_off_host[node] + k - 1] : 0;
	if (kmax) *kmax = e.kmax_host[node];
	return PGS_OK;
}

} // extern "C"
