// engine.cu -- C ABI (include/pgs.h): planning of the row layout and of the level sweep, device
// tables, launches.  There is no CPU path: every compute entry point ends in a CUDA kernel of
// kernels.cu / format.cu and fails with PGS_ERR_NO_DEVICE / PGS_ERR_CUDA otherwise.
#include "engine.h"
#include "edge_table.h"

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
	for (auto &p : h_stage)
		if (p) cudaFreeHost(p);
	for (auto &x : ev)
		if (x) cudaEventDestroy(x);
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
	int rc = alloc(b, bytes, what);
	if (rc) return rc;
	if (bytes) {
		cudaError_t e = cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
		if (e != cudaSuccess) return cuda_fail(e, what);
	}
	return PGS_OK;
}

int64_t Engine::device_bytes() const
{
	const DevBuf *all[] = {&d_rows, &d_row_base, &d_h1, &d_tab, &d_tab_off, &d_kmax, &d_dense_gid,
						   &d_dense_tasks, &d_sparse_tasks, &d_root_tasks, &d_leaf_base_words, &d_mismatch,
						   &d_tab8, &d_walk_steps, &d_walk_off};
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

// K1 for every level below the root: dense region, then sparse region.  Levels are separated by
// stream order (each launch reads rows written by the previous level's launches).
void Engine::enqueue_levels(cudaStream_t s, int64_t *launches)
{
	const DenseTask *dt = d_dense_tasks.as<DenseTask>();
	const SparseTask *st = d_sparse_tasks.as<SparseTask>();
	if (walk) { // dense genes: the top of the tree (one unit), then all subtrees side by side
		if (n_walk_top_units) {
			launch_walk_dense(tables, d_walk_steps.as<WalkStep>(), d_walk_off.as<int64_t>(), n_walk_top_units, s);
			(*launches)++;
		}
		launch_walk_dense(tables, d_walk_steps.as<WalkStep>(), d_walk_off.as<int64_t>() + n_walk_top_units, n_walk_units, s);
		(*launches)++;
	}
	for (int32_t lvl = 1; lvl <= max_level; lvl++) {
		const int64_t nd = dense_level_off[lvl + 1] - dense_level_off[lvl];
		if (nd > 0 && tables.n_dense > 0 && !walk) {
			launch_sweep_dense(tables, dt + dense_level_off[lvl], (int32_t)nd, s);
			(*launches)++;
		}
		const int64_t ns = sparse_level_off[lvl + 1] - sparse_level_off[lvl];
		if (ns > 0) {
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
	{                                      \
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
	for (int i = 0; i < 6 && ce == cudaSuccess; i++) ce = cudaEventCreate(&e.ev[i]);
	if (ce != cudaSuccess) {
		pgs::g_create_error = std::string("pgs_create: ") + cudaGetErrorString(ce);
		delete h;
		return PGS_ERR_CUDA;
	}
	e.stream = e.own_stream;
	if (const char *v = std::getenv("PGS_K1_VARIANT")) pgs::g_dense_variant = std::atoi(v);
	if (const char *v = std::getenv("PGS_WALK_VARIANT")) pgs::g_walk_variant = std::atoi(v);
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
	cudaSetDevice(h->e.device);
	cudaStreamSynchronize(h->e.stream);
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
	e.drop_graph();
	const int32_t N = e.n_nodes, n = e.n_genomes;

	// ---- copy and validate the gene table
	e.n_genes = n_genes;
	e.gene_id.assign(gene_id, gene_id + n_genes);
	e.origin.assign(origin_node, origin_node + n_genes);
	e.all_order.resize(n);
	if (all_order) {
		std::vector<char> seen(n, 0);
		for (int32_t i = 0; i < n; i++) {
			const int32_t g = all_order[i];
			if (g < 0 || g >= n || seen[g]) return e.fail(PGS_ERR_ARG, "pgs_set_genes: all_order is not a permutation");
			seen[g] = 1;
			e.all_order[i] = g;
		}
	} else {
		for (int32_t i = 0; i < n; i++) e.all_order[i] = i;
	}
	e.dense_slot.assign(n_genes, -1);
	std::vector<std::pair<uint32_t, int64_t>> dense; // (gene_id, table index)
	for (int64_t g = 0; g < n_genes; g++) {
		if (gene_id[g] < 0 || gene_id[g] >= (int64_t)0xFFFFFFFFll)
			return e.fail(PGS_ERR_ARG, "pgs_set_genes: gene_id outside [0, 2^32-1)");
		if (origin_node[g] < 0 || origin_node[g] >= N) return e.fail(PGS_ERR_ARG, "pgs_set_genes: origin out of range");
		const int64_t b = carrier_off[g], c = carrier_off[g + 1];
		if (b < 0 || c < b) return e.fail(PGS_ERR_ARG, "pgs_set_genes: carrier_off not monotone");
		if (c - b == 1 && carrier_genome[b] == PGS_ALL_GENOMES && origin_node[g] == e.root)
			dense.emplace_back((uint32_t)gene_id[g], g);
	}
	std::sort(dense.begin(), dense.end());
	const int64_t D = (int64_t)dense.size();
	e.dense_gid.resize(D);
	e.dense_gene.resize(D);
	for (int64_t s = 0; s < D; s++) {
		e.dense_gid[s] = dense[s].first;
		e.dense_gene[s] = dense[s].second;
		e.dense_slot[dense[s].second] = (int32_t)s;
	}

	// ---- sparse genes: expand carriers, mark the nodes on the paths carrier -> origin
	e.carrier_off.assign(n_genes + 1, 0);
	e.carrier_genome.clear();
	std::vector<std::vector<int64_t>> at_node(N);
	std::vector<int64_t> stamp(N, -1);
	e.genome_gene_count.assign(n, D);
	for (int64_t g = 0; g < n_genes; g++) {
		e.carrier_off[g] = (int64_t)e.carrier_genome.size();
		if (e.dense_slot[g] >= 0) continue;
		const int64_t b = carrier_off[g], c = carrier_off[g + 1];
		const int32_t org = origin_node[g];
		if (c - b == 1 && carrier_genome[b] == PGS_ALL_GENOMES) {
			// every genome below a non-root origin, in all_order
			for (int32_t i = 0; i < n; i++) {
				int32_t v = e.genome_leaf[e.all_order[i]];
				while (v != -1 && v != org) v = e.parent[v];
				if (v == org) e.carrier_genome.push_back(e.all_order[i]);
			}
		} else {
			for (int64_t k = b; k < c; k++) {
				const int32_t cg = carrier_genome[k];
				if (cg < 0 || cg >= n) return e.fail(PGS_ERR_ARG, "pgs_set_genes: carrier genome out of range");
				e.carrier_genome.push_back(cg);
			}
		}
		for (int64_t k = e.carrier_off[g]; k < (int64_t)e.carrier_genome.size(); k++) {
			int32_t v = e.genome_leaf[e.carrier_genome[k]];
			if (stamp[v] == g) return e.fail(PGS_ERR_ARG, "pgs_set_genes: genome listed twice for a gene");
			e.genome_gene_count[e.carrier_genome[k]]++;
			while (stamp[v] != g) {
				stamp[v] = g;
				at_node[v].push_back(g);
				if (v == org) break;
				v = e.parent[v];
				if (v < 0) return e.fail(PGS_ERR_ARG, "pgs_set_genes: carrier is not below the gene's origin");
			}
		}
	}
	e.carrier_off[n_genes] = (int64_t)e.carrier_genome.size();
	{ // gene ids must be unique across the whole table
		std::vector<int64_t> ids(e.gene_id);
		std::sort(ids.begin(), ids.end());
		if (std::adjacent_find(ids.begin(), ids.end()) != ids.end())
			return e.fail(PGS_ERR_ARG, "pgs_set_genes: duplicate gene_id");
	}

	e.sp_off.assign(N + 1, 0);
	for (int32_t v = 0; v < N; v++) e.sp_off[v + 1] = e.sp_off[v] + (int64_t)at_node[v].size();
	e.sp_gene.resize(e.sp_off[N]);
	for (int32_t v = 0; v < N; v++) std::copy(at_node[v].begin(), at_node[v].end(), e.sp_gene.begin() + e.sp_off[v]);
	at_node.clear();
	at_node.shrink_to_fit();

	// ---- sweep mode.  The fused column walk (k1_walk_dense) needs a binary tree and an even number
	// of blocks per gene; it keeps dense rows only where an output reads them (genomes and the root).
	bool walk_mode = N > 1 && D > 0 && (e.blocks % 2) == 0 && !(e.cfg.flags & PGS_FLAG_KEEP_INTERNAL);
	for (int32_t v = 0; v < N && walk_mode; v++) {
		const int32_t nc = e.child_off[v + 1] - e.child_off[v];
		if (nc != 0 && nc != 2) walk_mode = false;
	}
	e.materialised.assign(N, 1);
	if (walk_mode)
		for (int32_t v = 0; v < N; v++) e.materialised[v] = (e.child_off[v + 1] == e.child_off[v] || v == e.root) ? 1 : 0;
	auto dense_rows = [&](int32_t v) -> int64_t { return e.materialised[v] ? D : 0; };

	// ---- row layout: node-major, dense region (where kept) first, then the node's sparse rows;
	// the walk's transient rows (parked siblings, unit roots) live in a small pool behind the nodes
	e.row_base.assign(N, 0);
	int64_t rows = 0;
	for (int32_t v = 0; v < N; v++) {
		e.row_base[v] = rows;
		rows += dense_rows(v) + (e.sp_off[v + 1] - e.sp_off[v]);
	}
	if (D * e.blocks >= (int64_t)1 << 31) return e.fail(PGS_ERR_ARG, "pgs_set_genes: dense region of a node exceeds 2^31 blocks; shard the genes");

	// ---- tasks, grouped by the level of the children
	const int32_t ML = e.max_level;
	std::vector<std::vector<pgs::DenseTask>> dl(ML + 2);
	std::vector<std::vector<pgs::SparseTask>> sl(ML + 2);
	e.rows_written = e.rows_read = 0;
	for (int32_t p = 0; p < N; p++) {
		const int32_t nc = e.child_off[p + 1] - e.child_off[p];
		if (nc == 0) continue;
		const int32_t lvl = e.level[p] + 1;
		const int64_t *pg = e.sp_gene.data() + e.sp_off[p];
		const int64_t np = e.sp_off[p + 1] - e.sp_off[p];
		for (int32_t ci = 0; ci < nc; ci += 2) {
			const int32_t l = e.child[e.child_off[p] + ci];
			const int32_t r = (ci + 1 < nc) ? e.child[e.child_off[p] + ci + 1] : -1;
			if (D > 0) {
				dl[lvl].push_back(pgs::DenseTask{p, l, r});
				e.rows_read += D;
				e.rows_written += D * (r >= 0 ? 2 : 1);
			}
			// merge the parent's sparse list with the children's (all ascending)
			const int64_t *lg = e.sp_gene.data() + e.sp_off[l];
			const int64_t nl = e.sp_off[l + 1] - e.sp_off[l];
			const int64_t *rg = r >= 0 ? e.sp_gene.data() + e.sp_off[r] : nullptr;
			const int64_t nr = r >= 0 ? e.sp_off[r + 1] - e.sp_off[r] : 0;
			int64_t il = 0, ir = 0;
			for (int64_t ip = 0; ip < np; ip++) {
				const int64_t g = pg[ip];
				while (il < nl && lg[il] < g) il++;
				while (ir < nr && rg[ir] < g) ir++;
				const bool hl = il < nl && lg[il] == g, hr = ir < nr && rg[ir] == g;
				if (!hl && !hr) continue;
				pgs::SparseTask t;
				t.src = (uint32_t)(e.row_base[p] + dense_rows(p) + ip);
				t.dst_left = hl ? (uint32_t)(e.row_base[l] + dense_rows(l) + il) : pgs::kNoRow;
				t.dst_right = hr ? (uint32_t)(e.row_base[r] + dense_rows(r) + ir) : pgs::kNoRow;
				t.gid = (uint32_t)e.gene_id[g];
				t.edge_left = l;
				t.edge_right = r;
				sl[lvl].push_back(t);
				e.rows_read += 1;
				e.rows_written += (hl ? 1 : 0) + (hr ? 1 : 0);
			}
		}
	}
	std::vector<pgs::DenseTask> dense_tasks;
	std::vector<pgs::SparseTask> sparse_tasks;
	e.dense_level_off.assign(ML + 2, 0);
	e.sparse_level_off.assign(ML + 2, 0);
	for (int32_t lvl = 0; lvl <= ML; lvl++) {
		e.dense_level_off[lvl] = (int64_t)dense_tasks.size();
		e.sparse_level_off[lvl] = (int64_t)sparse_tasks.size();
		dense_tasks.insert(dense_tasks.end(), dl[lvl].begin(), dl[lvl].end());
		sparse_tasks.insert(sparse_tasks.end(), sl[lvl].begin(), sl[lvl].end());
	}
	e.dense_level_off[ML + 1] = (int64_t)dense_tasks.size();
	e.sparse_level_off[ML + 1] = (int64_t)sparse_tasks.size();
	e.n_dense_tasks = (int64_t)dense_tasks.size();
	e.n_sparse_tasks = (int64_t)sparse_tasks.size();

	std::vector<pgs::RootTask> root_tasks;
	root_tasks.reserve(n_genes);
	for (int64_t s = 0; s < D; s++) root_tasks.push_back(pgs::RootTask{(uint32_t)(e.row_base[e.root] + s), e.dense_gid[s]});
	for (int64_t g = 0; g < n_genes; g++) {
		if (e.dense_slot[g] >= 0) continue;
		const int32_t v = e.origin[g];
		const int64_t *b = e.sp_gene.data() + e.sp_off[v], *en = e.sp_gene.data() + e.sp_off[v + 1];
		const int64_t *it = std::lower_bound(b, en, g);
		if (it == en || *it != g) continue; // gene carried by no genome: no rows at all
		root_tasks.push_back(pgs::RootTask{(uint32_t)(e.row_base[v] + dense_rows(v) + (it - b)), (uint32_t)e.gene_id[g]});
	}
	e.n_root_tasks = (int64_t)root_tasks.size();
	e.rows_origin = e.n_root_tasks;

	// ---- per-branch decision tables
	e.h1_host.assign(N, 0);
	e.tab_off_host.assign(N, 0);
	e.kmax_host.assign(N, 0);
	e.tab_host.clear();
	for (int32_t v = 0; v < N; v++) {
		uint64_t H[65];
		const int km = (v == e.root) ? 0 : pgs::build_edge_table(e.rate[v], H);
		e.tab_off_host[v] = (int32_t)e.tab_host.size();
		e.kmax_host[v] = km;
		e.h1_host[v] = km > 0 ? H[1] : 0;
		for (int k = 1; k <= km; k++) e.tab_host.push_back(H[k]);
	}

	// ---- fused column walk of the dense region (see k1_walk_dense): cut the tree into units,
	// order every unit depth-first with the smaller subtree first
	int rc = 0;
	e.walk = false;
	e.n_walk_units = e.n_walk_top_units = 0;
	e.rows_stored = e.rows_written;
	e.rows_loaded = e.rows_read;
	std::vector<pgs::WalkStep> walk_steps;
	std::vector<int64_t> walk_off;
	int64_t pool_slots = 0; // transient dense regions (D rows each) behind the node regions
	if (walk_mode) {
		auto internal = [&](int32_t v) { return e.child_off[v + 1] > e.child_off[v]; };
		// subtree sizes (nodes), children before parents: process by decreasing level
		std::vector<int32_t> order(N);
		for (int32_t v = 0; v < N; v++) order[v] = v;
		std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return e.level[a] > e.level[b]; });
		std::vector<int64_t> size(N, 1);
		for (int32_t v : order)
			if (e.parent[v] >= 0) size[e.parent[v]] += size[v];
		// units: split the largest subtree until there are enough of them
		size_t target_units = 80;
		if (const char *u = std::getenv("PGS_WALK_UNITS")) target_units = (size_t)std::max(1, std::atoi(u)); // tuning knob
		std::vector<int32_t> roots{e.root};
		std::vector<char> top(N, 0);
		size_t n_top = 0;
		while (roots.size() < target_units && n_top < 4 * target_units) { // (a caterpillar never widens)
			size_t best = 0;
			for (size_t i = 1; i < roots.size(); i++)
				if (size[roots[i]] > size[roots[best]]) best = i;
			const int32_t v = roots[best];
			if (size[v] < 64) break;
			top[v] = 1;
			n_top++;
			roots.erase(roots.begin() + best);
			for (int32_t c = e.child_off[v]; c < e.child_off[v + 1]; c++)
				if (internal(e.child[c])) roots.push_back(e.child[c]);
		}
		std::sort(roots.begin(), roots.end(), [&](int32_t a, int32_t b) { return size[a] > size[b]; });
		const int64_t B = e.blocks;
		// where a node's dense region lives (uint4 units): its own region for genomes and the root,
		// a pool slot for unit roots (written by pass 1, read by pass 2) and parked siblings
		std::vector<int64_t> loc(N, -1);
		for (int32_t v = 0; v < N; v++)
			if (e.materialised[v]) loc[v] = e.row_base[v] * B;
		auto slot_off = [&](int64_t slot) { return (rows + slot * D) * B; };
		for (size_t u = 0; u < roots.size(); u++)
			if (roots[u] != e.root) loc[roots[u]] = slot_off(pool_slots++);
		auto make_step = [&](int32_t v, uint32_t flags) {
			const int32_t a = e.child[e.child_off[v]], b = e.child[e.child_off[v] + 1]; // a < b: a leads the pair
			pgs::WalkStep s{};
			s.in_off = loc[v] >= 0 ? loc[v] : 0;
			s.a_off = loc[a] >= 0 ? loc[a] : 0;
			s.b_off = loc[b] >= 0 ? loc[b] : 0;
			s.edge_a = (uint32_t)a;
			s.edge_b = (uint32_t)b;
			s.h32a = (uint32_t)(e.h1_host[a] >> 32);
			s.h32b = (uint32_t)(e.h1_host[b] >> 32);
			s.flags = flags;
			return s;
		};
		e.rows_stored = e.rows_loaded = 0;
		// Depth-first step list of the subtree below `start`, smaller child first.  `descend(c)`
		// says whether an internal child belongs to this walk; an internal child that does not
		// (the root of another unit) is written to its pool slot and picked up by its own unit.
		auto emit_walk = [&](int32_t start, auto descend) {
			const int64_t stack_base = pool_slots; // this walk's parked rows: slots stack_base + depth
			int64_t deepest = 0;
			std::vector<int32_t> parked;
			int32_t v = start;
			bool in_reg = false, have = true;
			while (have) {
				const int32_t a = e.child[e.child_off[v]], b = e.child[e.child_off[v] + 1];
				const bool ia = internal(a), ib = internal(b);
				const bool da = ia && descend(a), db = ib && descend(b);
				uint32_t f = 0;
				if (!in_reg) {
					f |= pgs::kWalkLoad;
					e.rows_loaded += D;
				}
				if (!ia) f |= pgs::kWalkStoreA;                       // a genome's row
				if (!ib) f |= pgs::kWalkStoreB;
				if (ia && !da) f |= pgs::kWalkStoreA | pgs::kWalkParkA; // root of another unit
				if (ib && !db) f |= pgs::kWalkStoreB | pgs::kWalkParkB;
				int32_t carry = -1;
				if (da && db) {
					const bool a_first = size[a] <= size[b];
					carry = a_first ? a : b;
					const int32_t later = a_first ? b : a;
					f |= a_first ? (pgs::kWalkCarryA | pgs::kWalkStoreB | pgs::kWalkParkB)
								 : (pgs::kWalkCarryB | pgs::kWalkStoreA | pgs::kWalkParkA);
					loc[later] = slot_off(stack_base + (int64_t)parked.size());
					parked.push_back(later);
					deepest = std::max(deepest, (int64_t)parked.size());
				} else if (da) {
					carry = a;
					f |= pgs::kWalkCarryA;
				} else if (db) {
					carry = b;
					f |= pgs::kWalkCarryB;
				}
				if (f & pgs::kWalkStoreA) e.rows_stored += D;
				if (f & pgs::kWalkStoreB) e.rows_stored += D;
				walk_steps.push_back(make_step(v, f));
				if (carry >= 0) {
					v = carry;
					in_reg = true;
				} else if (!parked.empty()) {
					v = parked.back();
					parked.pop_back();
					in_reg = false;
				} else {
					have = false;
				}
			}
			pool_slots += deepest;
		};
		// pass 1 (one unit): the top of the tree, down to the roots of the units of pass 2;
		// pass 2: every unit loads its root row and walks its own subtree
		walk_off.push_back(0);
		if (top[e.root]) {
			emit_walk(e.root, [&](int32_t c) { return top[c] != 0; });
			walk_off.push_back((int64_t)walk_steps.size());
			e.n_walk_top_units = 1;
		}
		for (int32_t r : roots) {
			emit_walk(r, [](int32_t) { return true; });
			walk_off.push_back((int64_t)walk_steps.size());
		}
		e.walk = true;
		e.n_walk_units = (int32_t)roots.size();
	}
	rows += pool_slots * D;
	if (rows >= (int64_t)pgs::kNoRow) return e.fail(PGS_ERR_ARG, "pgs_set_genes: more than 2^32-2 rows on one engine; shard the genes");
	e.rows_total = rows;
	if ((rc = e.upload(e.d_walk_steps, walk_steps.data(), sizeof(pgs::WalkStep) * walk_steps.size(), "walk steps"))) return rc;
	if ((rc = e.upload(e.d_walk_off, walk_off.data(), sizeof(int64_t) * walk_off.size(), "walk units"))) return rc;

	// ---- device
	const size_t row_bytes = (size_t)e.blocks * 16;
	// the sequence store is by far the largest allocation: keep it when a re-plan fits
	if (e.d_rows.bytes < (size_t)rows * row_bytes || e.d_rows.bytes > 2 * (size_t)rows * row_bytes + (64u << 20))
		if ((rc = e.alloc(e.d_rows, (size_t)rows * row_bytes, "rows (HBM sequence store)"))) return rc;
	if ((rc = e.upload(e.d_row_base, e.row_base.data(), sizeof(int64_t) * N, "row_base"))) return rc;
	if ((rc = e.upload(e.d_h1, e.h1_host.data(), sizeof(uint64_t) * N, "h1"))) return rc;
	if ((rc = e.upload(e.d_tab, e.tab_host.data(), sizeof(uint64_t) * e.tab_host.size(), "tab"))) return rc;
	if ((rc = e.upload(e.d_tab_off, e.tab_off_host.data(), sizeof(int32_t) * N, "tab_off"))) return rc;
	if ((rc = e.upload(e.d_kmax, e.kmax_host.data(), sizeof(int32_t) * N, "kmax"))) return rc;
	if ((rc = e.upload(e.d_dense_gid, e.dense_gid.data(), sizeof(uint32_t) * D, "dense_gid"))) return rc;
	if ((rc = e.upload(e.d_dense_tasks, dense_tasks.data(), sizeof(pgs::DenseTask) * dense_tasks.size(), "dense tasks"))) return rc;
	if ((rc = e.upload(e.d_sparse_tasks, sparse_tasks.data(), sizeof(pgs::SparseTask) * sparse_tasks.size(), "sparse tasks"))) return rc;
	if ((rc = e.upload(e.d_root_tasks, root_tasks.data(), sizeof(pgs::RootTask) * root_tasks.size(), "root tasks"))) return rc;
	std::vector<int64_t> leaf_words(n);
	for (int32_t g = 0; g < n; g++) leaf_words[g] = e.row_base[e.genome_leaf[g]] * e.blocks * 4;
	if ((rc = e.upload(e.d_leaf_base_words, leaf_words.data(), sizeof(int64_t) * n, "leaf bases"))) return rc;
	e.d_mismatch.release();
	{ // first eight thresholds of every branch at a fixed stride, for the kernels' shared-memory copy
		std::vector<uint64_t> tab8((size_t)N * 8, 0);
		for (int32_t v = 0; v < N; v++)
			for (int k = 0; k < 8 && k < e.kmax_host[v]; k++) tab8[(size_t)v * 8 + k] = e.tab_host[e.tab_off_host[v] + k];
		if ((rc = e.upload(e.d_tab8, tab8.data(), sizeof(uint64_t) * tab8.size(), "tab8"))) return rc;
	}

	pgs::DeviceTables &t = e.tables;
	t.rows = e.d_rows.as<uint4>();
	t.row_base = e.d_row_base.as<int64_t>();
	t.h1 = e.d_h1.as<uint64_t>();
	t.tab = e.d_tab.as<uint64_t>();
	t.tab_off = e.d_tab_off.as<int32_t>();
	t.kmax = e.d_kmax.as<int32_t>();
	t.dense_gid = e.d_dense_gid.as<uint32_t>();
	t.n_dense = (uint32_t)D;
	t.blocks = (uint32_t)e.blocks;
	t.last_valid = e.last_valid;
	t.div_blocks = pgs::make_fastdiv((uint32_t)e.blocks);
	t.tab8 = e.d_tab8.as<uint64_t>();
	t.keys = pgs::make_philox_keys(e.cfg.seed);
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
	if (e.cfg.flags & PGS_FLAG_NO_GRAPH) {
		e.enqueue_levels(e.stream, &launches);
	} else {
		if (!e.graph_exec) {
			// capture the level launches once; replays cost one graph launch per sweep
			int64_t captured = 0;
			ce = cudaStreamBeginCapture(e.stream, cudaStreamCaptureModeThreadLocal);
			if (ce != cudaSuccess) return e.cuda_fail(ce, "cudaStreamBeginCapture");
			e.enqueue_levels(e.stream, &captured);
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
	if (e.cfg.flags & PGS_FLAG_NO_GRAPH) e.launches_last_sweep = launches - (e.n_root_tasks ? 1 : 0);
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
		cudaGetLastError();
	}
	out->mismatch_ms = e.mismatch_ms;
	out->format_ms = e.format_ms;
	return PGS_OK;
}

int pgs_mismatch_device(pgs_engine *h, void *out_device_u32, int64_t *sites_out)
{
	ENGINE_OR_RETURN(h);
	if (!e.swept) return e.fail(PGS_ERR_STATE, "pgs_mismatch: call pgs_sweep first");
	if (!out_device_u32) return e.fail(PGS_ERR_ARG, "pgs_mismatch: null output");
	const int64_t D = (int64_t)e.dense_gid.size();
	cudaEvent_t t0 = e.ev[3], t1 = e.ev[4];
	cudaEventRecord(t0, e.stream);
	pgs::launch_mismatch(e.tables.rows, e.d_leaf_base_words.as<int64_t>(), e.n_genomes, D * e.blocks * 4,
						 static_cast<uint32_t *>(out_device_u32), e.stream);
	cudaEventRecord(t1, e.stream);
	cudaError_t ce = cudaStreamSynchronize(e.stream);
	if (ce == cudaSuccess) ce = cudaGetLastError();
	float ms = 0;
	if (ce == cudaSuccess) cudaEventElapsedTime(&ms, t0, t1);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch");
	e.mismatch_ms = ms;
	if (sites_out) *sites_out = D * e.cfg.gene_length;
	return PGS_OK;
}

int pgs_mismatch(pgs_engine *h, uint32_t *out_host, int64_t *sites_out)
{
	ENGINE_OR_RETURN(h);
	if (!out_host) return e.fail(PGS_ERR_ARG, "pgs_mismatch: null output");
	const size_t bytes = sizeof(uint32_t) * (size_t)e.n_genomes * (size_t)e.n_genomes;
	if (e.d_mismatch.bytes < bytes) {
		int rc = e.alloc(e.d_mismatch, bytes, "mismatch matrix");
		if (rc) return rc;
	}
	int rc = pgs_mismatch_device(h, e.d_mismatch.p, sites_out);
	if (rc) return rc;
	cudaError_t ce = cudaMemcpy(out_host, e.d_mismatch.p, bytes, cudaMemcpyDeviceToHost);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch copy");
	return PGS_OK;
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
	const int64_t row = e.row_index(node, gene_index);
	if (row < 0) return 1;
	if (e.dense_slot[gene_index] >= 0 && !e.materialised[node])
		return e.fail(PGS_ERR_STATE, "pgs_copy_packed: this internal row is not kept (create the engine with PGS_FLAG_KEEP_INTERNAL)");
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
