// plan.cu -- host-side planning behind pgs_set_genes: gene table -> row layout -> level tasks ->
// per-branch tables -> fused-walk step lists -> device tables.  Pure host code (compiled by nvcc
// only because it fills the launch structures of kernels.cuh).
#include "edge_table.h"
#include "engine.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

namespace pgs {

namespace {

// Everything a plan produces that is not kept in the Engine itself.
struct SweepPlan {
	int64_t D = 0;          // dense genes
	bool walk_mode = false; // fused column walk for the dense region
	int64_t rows = 0;       // rows of the store (node regions, then the walk's pool)
	std::vector<DenseTask> dense_tasks;
	std::vector<SparseTask> sparse_tasks;
	std::vector<RootTask> root_tasks;
	std::vector<WalkStep> walk_steps;
	std::vector<int64_t> walk_off;
	std::vector<TopMaskTask> top_tasks; // k1_top_masks: origin rows, then the branches below the top nodes
	std::vector<int32_t> unit_path_off; // per unit: its entries of unit_path
	std::vector<int64_t> unit_path;     // offsets of the mask regions on the path from the tree's root to the unit's root
	std::vector<int64_t> unit_store;    // per unit: where the root value is stored (a genome below a top node), -1 = nowhere
	bool binary = false;       // every internal node has exactly two children
	bool sparse_walk = false;  // fused walk for the sparse region
	// per sparse gene: the nodes that hold a row of it and the row's rank inside the node's sparse list
	std::vector<int64_t> gn_off;
	std::vector<int32_t> gn_node, gn_rank;
	std::vector<SparseUnit> sp_units;
	// steps and paths stay in the pieces the planner's threads produced (joined on the device: no host copy)
	std::vector<std::vector<SparseStep>> sp_steps;
	std::vector<std::vector<SparsePathEdge>> sp_paths;
};

int64_t dense_rows(const Engine &e, const SweepPlan &p, int32_t v)
{
	return e.materialised[v] ? p.D : 0;
}

// Copy and validate the gene table; classify dense genes; expand the carriers of sparse genes and
// mark the nodes on the paths carrier -> origin (the rows a sparse gene needs).
int plan_gene_table(Engine &e, SweepPlan &p, int64_t n_genes, const int64_t *gene_id, const int32_t *origin_node,
					const int64_t *carrier_off, const int32_t *carrier_genome, const int32_t *all_order)
{
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
	const int64_t D = p.D = (int64_t)dense.size();
	e.dense_gid.resize(D);
	e.dense_gene.resize(D);
	for (int64_t s = 0; s < D; s++) {
		e.dense_gid[s] = dense[s].first;
		e.dense_gene[s] = dense[s].second;
		e.dense_slot[dense[s].second] = (int32_t)s;
	}

	// ---- sparse genes: expand carriers, mark the nodes on the paths carrier -> origin.  Genes are
	// independent: contiguous ranges of the table are marked side by side on a few host threads (each
	// with its own stamp array), then joined in gene order.
	struct Part {
		std::vector<int32_t> carriers, nodes; // concatenated over the part's sparse genes
		std::vector<int64_t> n_carriers, n_nodes; // per gene of the range (0 for dense genes)
		std::string error;
	};
	auto mark_range = [&](int64_t g_lo, int64_t g_hi, Part &out) {
		std::vector<int64_t> stamp(N, -1);
		out.n_carriers.assign(g_hi - g_lo, 0);
		out.n_nodes.assign(g_hi - g_lo, 0);
		for (int64_t g = g_lo; g < g_hi; g++) {
			if (e.dense_slot[g] >= 0) continue;
			const int64_t b = carrier_off[g], c = carrier_off[g + 1];
			const int32_t org = origin_node[g];
			const size_t first_carrier = out.carriers.size(), first_node = out.nodes.size();
			if (c - b == 1 && carrier_genome[b] == PGS_ALL_GENOMES) {
				// every genome below a non-root origin, in all_order
				for (int32_t i = 0; i < n; i++) {
					int32_t v = e.genome_leaf[e.all_order[i]];
					while (v != -1 && v != org) v = e.parent[v];
					if (v == org) out.carriers.push_back(e.all_order[i]);
				}
			} else {
				for (int64_t k = b; k < c; k++) {
					const int32_t cg = carrier_genome[k];
					if (cg < 0 || cg >= n) {
						out.error = "pgs_set_genes: carrier genome out of range";
						return;
					}
					out.carriers.push_back(cg);
				}
			}
			for (size_t k = first_carrier; k < out.carriers.size(); k++) {
				int32_t v = e.genome_leaf[out.carriers[k]];
				if (stamp[v] == g) {
					out.error = "pgs_set_genes: genome listed twice for a gene";
					return;
				}
				while (stamp[v] != g) {
					stamp[v] = g;
					out.nodes.push_back(v);
					if (v == org) break;
					v = e.parent[v];
					if (v < 0) {
						out.error = "pgs_set_genes: carrier is not below the gene's origin";
						return;
					}
				}
			}
			out.n_carriers[g - g_lo] = (int64_t)(out.carriers.size() - first_carrier);
			out.n_nodes[g - g_lo] = (int64_t)(out.nodes.size() - first_node);
		}
	};
	const int64_t entries = n_genes > 0 ? carrier_off[n_genes] : 0;
	unsigned workers = entries >= 100000 ? std::min(16u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
	if (const char *v = std::getenv("PGS_PLAN_THREADS")) workers = (unsigned)std::max(1, std::atoi(v)); // tuning knob
	std::vector<Part> parts(workers);
	std::vector<int64_t> cuts(workers + 1, n_genes);
	cuts[0] = 0;
	for (unsigned w = 1; w < workers; w++) // balanced by carrier entries
		cuts[w] = std::lower_bound(carrier_off, carrier_off + n_genes, entries * w / workers) - carrier_off;
	{
		std::vector<std::thread> pool;
		for (unsigned w = 1; w < workers; w++) pool.emplace_back(mark_range, cuts[w], cuts[w + 1], std::ref(parts[w]));
		mark_range(cuts[0], cuts[1], parts[0]);
		for (auto &th : pool) th.join();
	}
	for (const Part &part : parts)
		if (!part.error.empty()) return e.fail(PGS_ERR_ARG, part.error);
	// join in gene order: carrier lists, per-gene node lists, and the per-node lists of sparse genes
	// (ascending gene index inside a node).  The parts hold disjoint, ascending gene ranges, so every part
	// owns a known slice of each list -- of each node's list too, from the per-part node counts -- and the
	// parts are copied side by side on the threads that marked them.
	e.carrier_off.assign(n_genes + 1, 0);
	e.genome_gene_count.assign(n, D);
	p.gn_off.assign(n_genes + 1, 0);
	e.sp_off.assign(N + 1, 0);
	std::vector<size_t> carrier_base(workers + 1, 0), node_base(workers + 1, 0);
	std::vector<std::vector<int64_t>> fill(workers); // per part: where its next entry of a node's list goes
	for (unsigned w = 0; w < workers; w++) {
		carrier_base[w + 1] = carrier_base[w] + parts[w].carriers.size();
		node_base[w + 1] = node_base[w] + parts[w].nodes.size();
	}
	{
		std::vector<std::thread> pool;
		auto count = [&](unsigned w) {
			fill[w].assign(N, 0);
			for (int32_t v : parts[w].nodes) fill[w][v]++;
		};
		for (unsigned w = 1; w < workers; w++) pool.emplace_back(count, w);
		count(0);
		for (auto &th : pool) th.join();
	}
	for (int32_t v = 0; v < N; v++) {
		int64_t at = e.sp_off[v];
		for (unsigned w = 0; w < workers; w++) {
			const int64_t c = fill[w][v];
			fill[w][v] = at;
			at += c;
		}
		e.sp_off[v + 1] = at;
	}
	e.carrier_genome.resize(carrier_base[workers]);
	p.gn_node.resize(node_base[workers]);
	p.gn_rank.resize(node_base[workers]);
	e.sp_gene.resize(e.sp_off[N]);
	{
		std::vector<std::vector<int64_t>> per_genome(workers);
		auto place = [&](unsigned w) {
			const Part &part = parts[w];
			std::vector<int64_t> &count = per_genome[w], &next = fill[w];
			count.assign(n, 0);
			size_t ci = 0, ni = 0;
			for (int64_t g = cuts[w]; g < cuts[w + 1]; g++) {
				e.carrier_off[g] = (int64_t)(carrier_base[w] + ci);
				p.gn_off[g] = (int64_t)(node_base[w] + ni);
				const int64_t nc = part.n_carriers[g - cuts[w]], nn = part.n_nodes[g - cuts[w]];
				for (int64_t k = 0; k < nc; k++) {
					const int32_t cg = part.carriers[ci + k];
					e.carrier_genome[carrier_base[w] + ci + k] = cg;
					count[cg]++;
				}
				for (int64_t k = 0; k < nn; k++) {
					const int32_t v = part.nodes[ni + k];
					const int64_t at = next[v]++;
					p.gn_rank[node_base[w] + ni + k] = (int32_t)(at - e.sp_off[v]);
					p.gn_node[node_base[w] + ni + k] = v;
					e.sp_gene[at] = g;
				}
				ci += (size_t)nc;
				ni += (size_t)nn;
			}
		};
		std::vector<std::thread> pool;
		for (unsigned w = 1; w < workers; w++) pool.emplace_back(place, w);
		place(0);
		for (auto &th : pool) th.join();
		for (unsigned w = 0; w < workers; w++)
			for (int32_t i = 0; i < n; i++) e.genome_gene_count[i] += per_genome[w][i];
	}
	e.carrier_off[n_genes] = (int64_t)e.carrier_genome.size();
	p.gn_off[n_genes] = (int64_t)p.gn_node.size();
	parts.clear();
	{ // gene ids must be unique across the whole table
		std::vector<int64_t> ids(e.gene_id);
		std::sort(ids.begin(), ids.end());
		if (std::adjacent_find(ids.begin(), ids.end()) != ids.end())
			return e.fail(PGS_ERR_ARG, "pgs_set_genes: duplicate gene_id");
	}

	return PGS_OK;
}

// Sweep mode and row layout.
int plan_layout(Engine &e, SweepPlan &p)
{
	const int32_t N = e.n_nodes;
	const int64_t D = p.D;

	// ---- sweep mode.  The fused column walk (k1_walk_dense) needs a binary tree and an even number
	// of blocks per gene; it keeps dense rows only where an output reads them (genomes and the root).
	bool &walk_mode = p.walk_mode;
	p.binary = true;
	for (int32_t v = 0; v < N && p.binary; v++) {
		const int32_t nc = e.child_off[v + 1] - e.child_off[v];
		if (nc != 0 && nc != 2) p.binary = false;
	}
	const bool keep = (e.cfg.flags & PGS_FLAG_KEEP_INTERNAL) != 0;
	walk_mode = N > 1 && D > 0 && (e.blocks % 2) == 0 && !keep && p.binary;
	// the sparse region has its own fused walk (any number of blocks); internal sparse rows are then
	// not kept either (PGS_SPARSE_WALK=0, development knob: level by level as with KEEP_INTERNAL)
	p.sparse_walk = N > 1 && !keep && p.binary && !e.sp_gene.empty();
	if (const char *v = std::getenv("PGS_SPARSE_WALK")) p.sparse_walk = p.sparse_walk && std::atoi(v) != 0;
	e.materialised.assign(N, 1);
	if (walk_mode)
		for (int32_t v = 0; v < N; v++) e.materialised[v] = (e.child_off[v + 1] == e.child_off[v] || v == e.root) ? 1 : 0;

	// ---- row layout: node-major, dense region (where kept) first, then the node's sparse rows;
	// the walk's transient rows (parked siblings, unit roots) live in a small pool behind the nodes
	e.row_base.assign(N, 0);
	int64_t &rows = p.rows;
	rows = 0;
	for (int32_t v = 0; v < N; v++) {
		e.row_base[v] = rows;
		rows += dense_rows(e, p, v) + (e.sp_off[v + 1] - e.sp_off[v]);
	}
	if (D * e.blocks >= (int64_t)1 << 31) return e.fail(PGS_ERR_ARG, "pgs_set_genes: dense region of a node exceeds 2^31 blocks; shard the genes");
	return PGS_OK;
}

// Level-synchronous tasks (dense node pairs, sparse parent rows) grouped by the level of the
// children, and the origin-row tasks of K0.
void plan_level_tasks(Engine &e, SweepPlan &plan)
{
	const int32_t N = e.n_nodes;
	const int64_t D = plan.D, n_genes = e.n_genes;
	std::vector<DenseTask> &dense_tasks = plan.dense_tasks;
	std::vector<SparseTask> &sparse_tasks = plan.sparse_tasks;
	std::vector<RootTask> &root_tasks = plan.root_tasks;
	auto dr = [&](int32_t v) -> int64_t { return dense_rows(e, plan, v); };

	// ---- tasks, grouped by the level of the children
	const int32_t ML = e.max_level;
	std::vector<std::vector<pgs::DenseTask>> dl(ML + 2);
	std::vector<std::vector<pgs::SparseTask>> sl(ML + 2);
	e.rows_written = e.rows_read = 0;
	e.sparse_rows_written = e.sparse_rows_read = 0;
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
				if (!plan.sparse_walk) {
					pgs::SparseTask t;
					t.src = (uint32_t)(e.row_base[p] + dr(p) + ip);
					t.dst_left = hl ? (uint32_t)(e.row_base[l] + dr(l) + il) : pgs::kNoRow;
					t.dst_right = hr ? (uint32_t)(e.row_base[r] + dr(r) + ir) : pgs::kNoRow;
					t.gid = (uint32_t)e.gene_id[g];
					t.edge_left = l;
					t.edge_right = r;
					sl[lvl].push_back(t);
				}
				e.rows_read += 1;
				e.rows_written += (hl ? 1 : 0) + (hr ? 1 : 0);
				e.sparse_rows_read += 1;
				e.sparse_rows_written += (hl ? 1 : 0) + (hr ? 1 : 0);
			}
		}
	}
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

	root_tasks.reserve(n_genes);
	// (the fused column walk draws the dense genes' origin rows itself: no K0 work for them)
	for (int64_t s = 0; s < D && !plan.walk_mode; s++)
		root_tasks.push_back(pgs::RootTask{(uint32_t)(e.row_base[e.root] + s), e.dense_gid[s]});
	int64_t sparse_origins = 0;
	for (int64_t g = 0; g < n_genes; g++) {
		if (e.dense_slot[g] >= 0) continue;
		const int32_t v = e.origin[g];
		const int64_t *b = e.sp_gene.data() + e.sp_off[v], *en = e.sp_gene.data() + e.sp_off[v + 1];
		const int64_t *it = std::lower_bound(b, en, g);
		if (it == en || *it != g) continue; // gene carried by no genome: no rows at all
		sparse_origins++;
		if (!plan.sparse_walk) // (the fused sparse walk draws the origin rows itself)
			root_tasks.push_back(pgs::RootTask{(uint32_t)(e.row_base[v] + dr(v) + (it - b)), (uint32_t)e.gene_id[g]});
	}
	e.n_root_tasks = (int64_t)root_tasks.size();
	e.rows_origin = D + sparse_origins;
}

// Per-branch decision tables (edge_table.cpp) in the layouts the kernels read.
void plan_edge_tables(Engine &e)
{
	const int32_t N = e.n_nodes;

	// ---- per-branch decision tables
	e.h1_host.assign(N, 0);
	e.tab_off_host.assign(N, 0);
	e.kmax_host.assign(N, 0);
	e.tab_host.clear();
	// one table per branch, each a few hundred floating-point operations: computed side by side on a
	// few host threads for big trees (every table is independent; the results do not depend on it)
	std::vector<uint64_t> all((size_t)N * 65);
	auto build = [&](int32_t lo, int32_t hi) {
		for (int32_t v = lo; v < hi; v++)
			e.kmax_host[v] = (v == e.root) ? 0 : pgs::build_edge_table(e.rate[v], all.data() + (size_t)v * 65);
	};
	unsigned workers = N >= 4096 ? std::min(16u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
	std::vector<std::thread> pool;
	for (unsigned w = 1; w < workers; w++)
		pool.emplace_back(build, (int32_t)((int64_t)N * w / workers), (int32_t)((int64_t)N * (w + 1) / workers));
	build(0, (int32_t)((int64_t)N / workers));
	for (auto &th : pool) th.join();
	for (int32_t v = 0; v < N; v++) {
		const uint64_t *H = all.data() + (size_t)v * 65;
		const int km = e.kmax_host[v];
		e.tab_off_host[v] = (int32_t)e.tab_host.size();
		e.h1_host[v] = km > 0 ? H[1] : 0;
		for (int k = 1; k <= km; k++) e.tab_host.push_back(H[k]);
	}
}

// Fused column walk of the dense region: top of the tree, units, depth-first step lists, pool of
// transient rows.
//
// A substitution XORs a value-independent mask into the sequence (philox.cuh), so the value at any
// node is  origin row ^ XOR of the masks of the branches on the path from the root.  The tree is
// cut into a small TOP (the nodes whose subtrees were split) and the UNITS hanging below it.
// k1_top_masks draws the origin rows and the masks of all branches below top nodes side by side (no
// node waits for its parent); k1_walk_dense then walks every unit on its own: a unit's root value is
// the origin row XORed with the few masks on its path (at most ~16 for 160 units of a 10 000-leaf
// coalescent).  No pass waits for another one.
int plan_walk(Engine &e, SweepPlan &p)
{
	const int32_t N = e.n_nodes;
	const int64_t D = p.D;
	const bool walk_mode = p.walk_mode;
	int64_t &rows = p.rows;
	std::vector<WalkStep> &walk_steps = p.walk_steps;
	std::vector<int64_t> &walk_off = p.walk_off;

	e.walk = false;
	e.walk_pass_units.clear();
	e.rows_stored = e.rows_written;
	e.rows_loaded = e.rows_read;
	int64_t pool_slots = 0; // transient dense regions (D rows each) behind the node regions
	e.walk_smem_levels = 0;
	e.rows_parked_smem = 0;
	e.n_walk_units = 0;
	if (walk_mode) {
		// a thread's parked siblings go to its shared-memory stack; levels beyond what fits go to the pool
		double mean_hit = 0.0; // mean over the branches of P(a block changes) = H[1] / 2^64
		for (int32_t v = 0; v < N; v++) mean_hit += (double)e.h1_host[v] * 0x1p-64;
		e.walk_variant = pgs::choose_walk_variant((uint32_t)e.blocks, N > 1 ? mean_hit / (N - 1) : 0.0);
		const int smem_cap = pgs::walk_smem_capacity(e.walk_variant);
		std::vector<int32_t> smem_level(N, -1); // stack level of a node parked in shared memory
		auto internal = [&](int32_t v) { return e.child_off[v + 1] > e.child_off[v]; };
		// subtree sizes (nodes), children before parents: process by decreasing level
		std::vector<int32_t> order(N);
		for (int32_t v = 0; v < N; v++) order[v] = v;
		std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return e.level[a] > e.level[b]; });
		std::vector<int64_t> size(N, 1);
		for (int32_t v : order)
			if (e.parent[v] >= 0) size[e.parent[v]] += size[v];
		// units: split the largest subtree until there are enough of them.  Enough = the walk gets
		// about four waves of CTAs (units x column slabs; 444 CTAs are resident), but no fewer than 40
		// units and no more than 120 (every split adds a top node, whose branches are the long ones:
		// the work of k1_top_masks grows with the unit count).  Config 4 on one GPU: 79 slabs -> 40
		// units; an eighth of it: 10 slabs -> 120 units (0.288 ms against 0.296 with 160, 0.313 with 80).
		size_t target_units = 0;
		{
			const size_t slabs = std::max<size_t>(1, pgs::walk_slabs(e.walk_variant, (uint32_t)D, (uint32_t)e.blocks));
			// (a small tree has few steps to hand out: units of at least ~64 steps, so that a CTA's
			// fixed cost -- prologue, step staging -- stays small beside its walk)
			const size_t lo = std::min<size_t>(40, std::max<size_t>(8, (size_t)N / 128));
			target_units = std::min<size_t>(120, std::max<size_t>(lo, (1800 + slabs - 1) / slabs));
		}
		// Many units are wanted when there are few column slabs.  Beyond ~60 the last third of them is cut
		// without top nodes (carve_units below: pieces reached across entry steps in registers): no mask
		// rows, no longer mask paths.  An eighth of config 4: 0.2747 ms against 0.2768 with 120 top-cut
		// units, 0.2755 / 0.2793 with 60 / 40 of 120 (profiles/r02_shard_of.txt).
		size_t carve_units = 0;
		if (target_units > 60) {
			carve_units = target_units;
			target_units = std::max<size_t>(60, target_units * 2 / 3);
		}
		if (const char *u = std::getenv("PGS_WALK_UNITS")) { // tuning knob
			target_units = (size_t)std::max(1, std::atoi(u));
			carve_units = 0;
		}
		if (const char *v = std::getenv("PGS_WALK_CARVE")) carve_units = (size_t)std::max(0, std::atoi(v)); // tuning knob
		std::vector<int32_t> roots{e.root};
		std::vector<char> top(N, 0), piece_root(N, 0);
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
		// longest unit first (CTAs are numbered unit-major: the shortest units fill the tail of the launch)
		std::sort(roots.begin(), roots.end(), [&](int32_t a, int32_t b) { return size[a] > size[b]; });
		// genomes hanging directly below a top node: units without steps (root value = the genome's row)
		for (int32_t v = 0; v < N; v++)
			if (top[v])
				for (int32_t c = e.child_off[v]; c < e.child_off[v + 1]; c++)
					if (!internal(e.child[c])) roots.push_back(e.child[c]);
		const int64_t B = e.blocks;
		// where a node's dense region lives (uint4 units): its own region for genomes and the root,
		// a pool slot for the deeper levels of a unit's stack of parked siblings
		std::vector<int64_t> loc(N, -1);
		for (int32_t v = 0; v < N; v++)
			if (e.materialised[v]) loc[v] = e.row_base[v] * B;
		auto slot_off = [&](int64_t slot) { return (rows + slot * D) * B; };
		// k1_top_masks: task 0 draws the origin rows, then one task per top node (both branches below it)
		std::vector<int64_t> mask_off(N, -1); // per child of a top node: where its branch's mask region lives
		{
			pgs::TopMaskTask tk{};
			tk.ma_off = loc[e.root];
			tk.edge_a = pgs::kRootEdge;
			// test knob: the origin rows are written this many clock cycles late, so that a walk that read one above
			// its griddepcontrol.wait gets a stale row every time (tests/test_sweep_gpu.py)
			if (const char *v = std::getenv("PGS_TEST_ORIGIN_DELAY")) tk.h1a = (uint64_t)std::max(0L, std::atol(v));
			p.top_tasks.push_back(tk);
		}
		for (int32_t v = 0; v < N; v++) {
			if (!top[v]) continue;
			const int32_t a = e.child[e.child_off[v]], b = e.child[e.child_off[v] + 1];
			pgs::TopMaskTask tk{};
			tk.ma_off = mask_off[a] = slot_off(pool_slots++);
			tk.mb_off = mask_off[b] = slot_off(pool_slots++);
			tk.edge_a = (uint32_t)a;
			tk.edge_b = (uint32_t)b;
			tk.h1a = e.h1_host[a];
			tk.h1b = e.h1_host[b];
			tk.h2a = e.kmax_host[a] >= 2 ? e.tab_host[e.tab_off_host[a] + 1] : 0;
			tk.h2b = e.kmax_host[b] >= 2 ? e.tab_host[e.tab_off_host[b] + 1] : 0;
			p.top_tasks.push_back(tk);
		}
		// CTAs start in task order and a task costs what its branches' substitutions cost: the long
		// branches next to the root (the highest node indices) go first, not into the tail of the launch
		if (!std::getenv("PGS_TOP_UNSORTED")) // (A/B knob)
			std::stable_sort(p.top_tasks.begin() + 1, p.top_tasks.end(), [&](const pgs::TopMaskTask &x, const pgs::TopMaskTask &y) {
				return e.rate[x.edge_a] + e.rate[x.edge_b] > e.rate[y.edge_a] + e.rate[y.edge_b];
			});
		auto make_step = [&](int32_t v, uint32_t flags) {
			const int32_t a = e.child[e.child_off[v]], b = e.child[e.child_off[v] + 1]; // a < b: a leads the pair
			pgs::WalkStep s{};
			s.in_off = loc[v] >= 0 ? loc[v] : 0;
			s.a_off = loc[a] >= 0 ? loc[a] : 0;
			s.b_off = loc[b] >= 0 ? loc[b] : 0;
			s.edge_a = (uint32_t)a;
			s.edge_b = (uint32_t)b;
			s.hca = decision_cmp(e.h1_host[a]);
			s.hcb = decision_cmp(e.h1_host[b]);
			s.h1a = e.h1_host[a];
			s.h1b = e.h1_host[b];
			s.h2a = e.kmax_host[a] >= 2 ? e.tab_host[e.tab_off_host[a] + 1] : 0;
			s.h2b = e.kmax_host[b] >= 2 ? e.tab_host[e.tab_off_host[b] + 1] : 0;
			s.flags = flags;
			if (flags & pgs::kWalkSmemA) s.slots |= (uint32_t)smem_level[a];
			if (flags & pgs::kWalkSmemB) s.slots |= (uint32_t)smem_level[b];
			if (flags & pgs::kWalkLoadSmem) s.slots |= (uint32_t)smem_level[v] << 8;
			return s;
		};
		// rows K1 really moves: the sparse region's as before (plan_sparse_walk corrects them if it
		// walks too), the dense region's counted below
		e.rows_stored = e.sparse_rows_written + D * (int64_t)(1 + 2 * n_top); // k1_top_masks: origin rows and masks
		e.rows_loaded = e.sparse_rows_read;
		// Depth-first step list of the subtree below `start`, smaller child first.  The unit's root
		// value is in registers when the first step begins (unit prologue of k1_walk_dense).
		// A PIECE of a unit: the subtree below `start` without the subtrees of the pieces cut out of it
		// (piece_root[] marks their roots).  Its walk begins at the unit's root -- that is the value the
		// prologue delivers -- and follows `entry`, the nodes from the unit's root down to start's parent,
		// in registers: one step each that keeps only the child on the path (nothing stored or parked; the
		// piece that owns the node stores its other child).  Cheap, because the branches inside a unit are
		// short; no mask rows and no longer mask paths as another top node would cost.
		auto emit_walk = [&](int32_t start, const std::vector<int32_t> &entry) {
			const int64_t stack_base = pool_slots; // this walk's deep parked rows: slots stack_base + depth - smem_cap
			int64_t deepest = 0;
			std::vector<int32_t> parked;
			for (size_t i = 0; i < entry.size(); i++) {
				const int32_t v = entry[i], next = i + 1 < entry.size() ? entry[i + 1] : start;
				walk_steps.push_back(make_step(v, next == e.child[e.child_off[v]] ? pgs::kWalkCarryA : pgs::kWalkCarryB));
			}
			int32_t v = start;
			bool in_reg = true, have = internal(start);
			while (have) {
				const int32_t a = e.child[e.child_off[v]], b = e.child[e.child_off[v] + 1];
				const bool ca = piece_root[a] != 0, cb = piece_root[b] != 0; // cut out: another piece's business
				const bool ia = internal(a) && !ca, ib = internal(b) && !cb;
				uint32_t f = 0;
				if (!in_reg && smem_level[v] >= 0) {
					f |= pgs::kWalkLoadSmem;
				} else if (!in_reg) {
					f |= pgs::kWalkLoad;
					e.rows_loaded += D;
				}
				if (!ia && !ca) f |= pgs::kWalkStoreA; // a genome's row
				if (!ib && !cb) f |= pgs::kWalkStoreB;
				int32_t carry = -1;
				if (ia && ib) {
					const bool a_first = size[a] <= size[b];
					carry = a_first ? a : b;
					const int32_t later = a_first ? b : a;
					const int64_t depth = (int64_t)parked.size();
					if (depth < smem_cap) {
						f |= a_first ? (pgs::kWalkCarryA | pgs::kWalkSmemB) : (pgs::kWalkCarryB | pgs::kWalkSmemA);
						smem_level[later] = (int32_t)depth;
						e.walk_smem_levels = std::max(e.walk_smem_levels, (int32_t)depth + 1);
						e.rows_parked_smem += D;
					} else {
						f |= a_first ? (pgs::kWalkCarryA | pgs::kWalkStoreB | pgs::kWalkParkB)
									 : (pgs::kWalkCarryB | pgs::kWalkStoreA | pgs::kWalkParkA);
						loc[later] = slot_off(stack_base + depth - smem_cap);
						deepest = std::max(deepest, depth - smem_cap + 1);
					}
					parked.push_back(later);
				} else if (ia) {
					carry = a;
					f |= pgs::kWalkCarryA;
				} else if (ib) {
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
		// Pieces.  A unit may be cut further without another top node: the nodes marked piece_root[] start
		// walks of their own (emit_walk's entry steps).  Two rules:
		//  * carve_units = T2 (above; PGS_WALK_CARVE): the top-down split continues until there are T2
		//    subtrees; the nodes split in this stage stay with the piece above them;
		//  * PGS_WALK_PIECE=P (off by default): bottom up, a node whose part of the unit has grown to P steps
		//    becomes a piece root (pieces of P .. 2P steps).  Measured on an eighth of config 4 this evens
		//    out the end of the launch (the SMs finish within 7 us instead of 22) but costs more than that in
		//    prologues (origin row + path masks, ~4.5 us per CTA) and entry steps: 0.289-0.313 ms against 0.276.
		// Neither cuts deeper below the unit's root than kMaxEntry steps.
		struct Piece {
			int32_t unit_root, start;
			int64_t steps; // its own steps, entry included
			std::vector<int32_t> entry;
		};
		std::vector<Piece> pieces;
		{
			int64_t piece_steps = (int64_t)1 << 40;
			if (const char *v = std::getenv("PGS_WALK_PIECE")) piece_steps = std::max(1, std::atoi(v)); // tuning knob
			constexpr int kMaxEntry = 24;
			std::vector<int64_t> acc(N, 0);
			std::vector<int32_t> depth(N, 0), nodes;
			{ // depth below the unit's root, for every internal node of every unit
				for (int32_t r : roots) {
					if (!internal(r)) continue;
					nodes.assign(1, r);
					depth[r] = 0;
					for (size_t i = 0; i < nodes.size(); i++)
						for (int32_t c = e.child_off[nodes[i]]; c < e.child_off[nodes[i] + 1]; c++)
							if (internal(e.child[c])) {
								depth[e.child[c]] = depth[nodes[i]] + 1;
								nodes.push_back(e.child[c]);
							}
				}
			}
			if (carve_units > 0) {
				std::vector<int32_t> cand;
				for (int32_t r : roots)
					if (internal(r)) cand.push_back(r);
				size_t splits = 0;
				while (cand.size() < carve_units && splits < 4 * carve_units) {
					size_t best = cand.size();
					for (size_t i = 0; i < cand.size(); i++)
						if (depth[cand[i]] < kMaxEntry && (best == cand.size() || size[cand[i]] > size[cand[best]])) best = i;
					if (best == cand.size() || size[cand[best]] < 64) break; // (as the top cut: a small tree is cut no further than before)
					const int32_t v = cand[best];
					piece_root[v] = 0; // stays with the piece above it (a unit's root: with the unit's remainder)
					splits++;
					cand.erase(cand.begin() + best);
					for (int32_t c = e.child_off[v]; c < e.child_off[v + 1]; c++)
						if (internal(e.child[c])) {
							piece_root[e.child[c]] = 1;
							cand.push_back(e.child[c]);
						}
				}
			}
			for (int32_t r : roots) {
				if (!internal(r)) {
					pieces.push_back(Piece{r, r, 0, {}});
					continue;
				}
				// the unit's internal nodes, parents before children
				nodes.assign(1, r);
				for (size_t i = 0; i < nodes.size(); i++)
					for (int32_t c = e.child_off[nodes[i]]; c < e.child_off[nodes[i] + 1]; c++)
						if (internal(e.child[c])) nodes.push_back(e.child[c]);
				for (size_t i = nodes.size(); i-- > 0;) {
					const int32_t v = nodes[i];
					acc[v] = 1;
					for (int32_t c = e.child_off[v]; c < e.child_off[v + 1]; c++)
						if (internal(e.child[c]) && !piece_root[e.child[c]]) acc[v] += acc[e.child[c]];
					if (v != r && acc[v] >= piece_steps && depth[v] <= kMaxEntry) piece_root[v] = 1;
					if (v != r && piece_root[v]) {
						Piece pc{r, v, acc[v] + depth[v], {}};
						for (int32_t w = e.parent[v];; w = e.parent[w]) {
							pc.entry.push_back(w);
							if (w == r) break;
						}
						std::reverse(pc.entry.begin(), pc.entry.end());
						pieces.push_back(std::move(pc));
					}
				}
				pieces.push_back(Piece{r, r, acc[r], {}});
			}
			// longest piece first (CTAs are numbered piece-major: the shortest pieces fill the tail of the launch)
			std::stable_sort(pieces.begin(), pieces.end(), [](const Piece &x, const Piece &y) { return x.steps > y.steps; });
		}
		walk_off.push_back(0);
		p.unit_path_off.push_back(0);
		for (const Piece &pc : pieces) {
			const int32_t r = pc.unit_root;
			// the masks on the path root -> r, and where the value goes if r is a genome
			for (int32_t v = r; v != e.root; v = e.parent[v]) p.unit_path.push_back(mask_off[v]);
			p.unit_path_off.push_back((int32_t)p.unit_path.size());
			e.rows_loaded += D * (1 + (p.unit_path_off.back() - p.unit_path_off[p.unit_path_off.size() - 2]));
			const bool genome = !internal(r);
			p.unit_store.push_back(genome ? loc[r] : -1);
			if (genome) e.rows_stored += D;
			emit_walk(pc.start, pc.entry);
			walk_off.push_back((int64_t)walk_steps.size());
		}
		e.walk_pass_units.push_back((int32_t)pieces.size());
		e.n_walk_units = (int32_t)pieces.size();
		e.walk = true;
	}
	rows += pool_slots * D;
	if (rows >= (int64_t)pgs::kNoRow) return e.fail(PGS_ERR_ARG, "pgs_set_genes: more than 2^32-2 rows on one engine; shard the genes");
	e.rows_total = rows;

	return PGS_OK;
}

// Fused walk of the sparse region (k1_walk_sparse).  A gene's presence subtree is cut into units of
// at most ~kSpUnitSteps steps (a widespread gene would otherwise be one serial chain of up to 2n
// steps); every unit gets a depth-first step list, smaller subtree first (a thread's stack of
// parked siblings stays shallow: at most log2(carriers) levels), and the list of branches from the
// gene's origin down to its root.
int plan_sparse_walk(Engine &e, SweepPlan &p)
{
	e.sparse_walk = p.sparse_walk;
	e.n_sp_units = 0;
	if (!p.sparse_walk) return PGS_OK;
	const int32_t N = e.n_nodes;
	const int smem_cap = pgs::sparse_walk_smem_levels();
	// steps per unit (a unit below the origin pays 4 + depth Philox calls to start).  Config 3's
	// accessory genes: 0.37 ms with 48, 0.41 with 96, 0.62 with 192, 1.26 uncut.
	int32_t unit_steps = 48;
	if (const char *v = std::getenv("PGS_SPARSE_UNIT_STEPS")) unit_steps = std::max(1, std::atoi(v)); // tuning knob
	auto internal = [&](int32_t v) { return e.child_off[v + 1] > e.child_off[v]; };

	// Genes are independent: contiguous ranges of the gene table (balanced by rows) are planned side by
	// side on a few host threads, each into its own lists; the lists are then joined in gene order, so
	// the result does not depend on the number of threads.
	struct Part {
		std::vector<SparseUnit> units;
		std::vector<SparseStep> steps;
		std::vector<SparsePathEdge> paths;
		int64_t stored = 0, loaded = 0; // rows the sparse walk writes to / reads from memory
	};
	auto plan_range = [&](int64_t g_lo, int64_t g_hi, Part &out) {
		std::vector<int64_t> stamp(N, -1);
		std::vector<int32_t> rank_at(N, 0), sub(N, 0), acc(N, 0);
		std::vector<char> cut(N, 0);
		std::vector<int32_t> order, parked, unit_roots, level_count;
		auto row_of = [&](int32_t v) { return (uint32_t)(e.row_base[v] + dense_rows(e, p, v) + rank_at[v]); };
		// one allocation up front (a step per internal row at most: about half of the range's rows)
		out.steps.reserve((size_t)(p.gn_off[g_hi] - p.gn_off[g_lo]) / 2 + 64);
		for (int64_t g = g_lo; g < g_hi; g++) {
			const int64_t nb = p.gn_off[g], ne = p.gn_off[g + 1];
			if (nb == ne) continue;
			for (int64_t k = nb; k < ne; k++) {
				const int32_t v = p.gn_node[k];
				stamp[v] = g;
				rank_at[v] = p.gn_rank[k];
				sub[v] = 1;
				acc[v] = internal(v) ? 1 : 0; // steps (internal nodes) of the piece of the subtree that is not cut off yet
				cut[v] = 0;
			}
			// children before parents: subtree sizes, and a cut wherever a piece has grown to a unit.
			// Deepest level first by a counting sort (a few dozen levels; std::sort was half of this loop's time).
			{
				level_count.assign((size_t)e.max_level + 2, 0);
				for (int64_t k = nb; k < ne; k++) level_count[e.max_level - e.level[p.gn_node[k]] + 1]++;
				for (size_t l = 1; l < level_count.size(); l++) level_count[l] += level_count[l - 1];
				order.resize((size_t)(ne - nb));
				for (int64_t k = nb; k < ne; k++) order[level_count[e.max_level - e.level[p.gn_node[k]]]++] = p.gn_node[k];
			}
			const int32_t org = e.origin[g];
			unit_roots.assign(1, org);
			for (int32_t v : order) {
				if (v == org) continue;
				sub[e.parent[v]] += sub[v];
				if (acc[v] >= unit_steps) {
					cut[v] = 1;
					unit_roots.push_back(v);
				} else {
					acc[e.parent[v]] += acc[v];
				}
			}
			for (int32_t ur : unit_roots) {
				pgs::SparseUnit su{};
				su.gid = (uint32_t)e.gene_id[g];
				su.origin_row = ur == org ? row_of(org) : pgs::kNoRow;
				su.step_begin = (uint32_t)out.steps.size();
				su.path_begin = (uint32_t)out.paths.size();
				if (ur == org) out.stored++;
				{ // branches origin -> unit root, top down
					const size_t first = out.paths.size();
					for (int32_t v = ur; v != org; v = e.parent[v]) {
						const int32_t par = e.parent[v], leader = e.child[e.child_off[par]];
						pgs::SparsePathEdge pe{};
						pe.leader = (uint32_t)leader;
						pe.edge = (uint32_t)v;
						pe.side_h16 = (uint32_t)(e.h1_host[v] >> 48) | (v == leader ? 0u : 0x80000000u);
						out.paths.push_back(pe);
					}
					std::reverse(out.paths.begin() + first, out.paths.end());
					su.path_len = (uint32_t)(out.paths.size() - first);
				}
				parked.clear();
				int32_t v = ur;
				uint32_t load = 0, load_level = 0, load_row = 0; // how the next step gets its input (0: carried in registers)
				bool have = internal(ur);
				while (have) {
					const int32_t a = e.child[e.child_off[v]], b = e.child[e.child_off[v] + 1];
					// a child that roots another unit is none of this unit's business
					const bool ha = stamp[a] == g && !cut[a], hb = stamp[b] == g && !cut[b];
					const bool da = ha && internal(a), db = hb && internal(b);
					pgs::SparseStep st{};
					st.dst_a = (ha && !da) ? row_of(a) : pgs::kNoRow; // a genome's row
					st.dst_b = (hb && !db) ? row_of(b) : pgs::kNoRow;
					st.edge_a = (uint32_t)a;
					st.edge_b = (uint32_t)b;
					st.h16 = (uint32_t)(e.h1_host[a] >> 48) | ((uint32_t)(e.h1_host[b] >> 48) << 16);
					st.flags = (ha ? pgs::kSpHasA : 0u) | (hb ? pgs::kSpHasB : 0u) | load;
					st.in_row = load_row;
					st.levels = load_level << 8;
					if (load == pgs::kSpLoadRow) out.loaded++;
					int32_t carry = -1;
					if (da && db) {
						const bool a_first = sub[a] <= sub[b];
						carry = a_first ? a : b;
						const int32_t later = a_first ? b : a;
						const int depth = (int)parked.size();
						st.flags |= a_first ? pgs::kSpCarryA : pgs::kSpCarryB;
						if (depth < smem_cap) {
							st.flags |= a_first ? pgs::kSpParkB : pgs::kSpParkA;
							st.levels |= (uint32_t)depth;
						} else { // deeper than the shared-memory stack: parked in its own row
							(a_first ? st.dst_b : st.dst_a) = row_of(later);
						}
						parked.push_back(later);
					} else if (da) {
						carry = a;
						st.flags |= pgs::kSpCarryA;
					} else if (db) {
						carry = b;
						st.flags |= pgs::kSpCarryB;
					}
					out.stored += (st.dst_a != pgs::kNoRow) + (st.dst_b != pgs::kNoRow);
					if (ha || hb) out.steps.push_back(st); // (both children cut off: nothing to do here)
					if (carry >= 0) {
						v = carry;
						load = 0;
					} else if (!parked.empty()) {
						v = parked.back();
						parked.pop_back();
						const int depth = (int)parked.size();
						if (depth < smem_cap) {
							load = pgs::kSpLoadSmem;
							load_level = (uint32_t)depth;
						} else {
							load = pgs::kSpLoadRow;
							load_row = row_of(v);
						}
					} else {
						have = false;
					}
				}
				su.n_steps = (uint32_t)(out.steps.size() - su.step_begin);
				out.units.push_back(su);
			}
		}
	};
	const int64_t total_rows = p.gn_off[e.n_genes];
	if (total_rows + 8 >= (int64_t)0xF0000000ll) return e.fail(PGS_ERR_ARG, "pgs_set_genes: more than 2^32 sparse rows on one engine; shard the genes");
	unsigned workers = total_rows >= 200000 ? std::min(16u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
	if (const char *v = std::getenv("PGS_PLAN_THREADS")) workers = (unsigned)std::max(1, std::atoi(v)); // tuning knob
	std::vector<Part> parts(workers);
	std::vector<int64_t> cuts(workers + 1, e.n_genes);
	cuts[0] = 0;
	for (unsigned w = 1; w < workers; w++) // first gene whose rows start at or beyond the w-th share
		cuts[w] = std::lower_bound(p.gn_off.begin(), p.gn_off.begin() + e.n_genes, total_rows * w / workers) - p.gn_off.begin();
	const bool timing = std::getenv("PGS_PLAN_TIMING") != nullptr;
	auto tp0 = std::chrono::steady_clock::now();
	auto lap = [&](const char *what) {
		if (!timing) return;
		const auto t1 = std::chrono::steady_clock::now();
		std::fprintf(stderr, "plan:   sparse %-10s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - tp0).count());
		tp0 = t1;
	};
	{
		std::vector<std::thread> pool;
		for (unsigned w = 1; w < workers; w++) pool.emplace_back(plan_range, cuts[w], cuts[w + 1], std::ref(parts[w]));
		plan_range(cuts[0], cuts[1], parts[0]);
		for (auto &th : pool) th.join();
	}
	lap("ranges");
	int64_t stored = 0, loaded = 0;
	uint32_t step0 = 0, path0 = 0;
	for (Part &part : parts) { // join in gene order; a part's indices move behind what is already there
		for (SparseUnit &u : part.units) {
			u.step_begin += step0;
			u.path_begin += path0;
		}
		step0 += (uint32_t)part.steps.size();
		path0 += (uint32_t)part.paths.size();
		p.sp_units.insert(p.sp_units.end(), part.units.begin(), part.units.end());
		p.sp_steps.push_back(std::move(part.steps));
		p.sp_paths.push_back(std::move(part.paths));
		stored += part.stored;
		loaded += part.loaded;
	}
	lap("join");
	if (timing) { // order-independent checksum of the units and their step lists (compare two planner builds without a device)
		auto fnv = [](const void *q, size_t bytes) {
			uint64_t h = 1469598103934665603ull;
			const unsigned char *b = static_cast<const unsigned char *>(q);
			for (size_t i = 0; i < bytes; i++) h = (h ^ b[i]) * 1099511628211ull;
			return h;
		};
		uint64_t sum = 0;
		size_t piece = 0, base = 0;
		for (const SparseUnit &u : p.sp_units) { // units are still in gene order: a unit's steps lie in one piece
			while (u.step_begin >= base + p.sp_steps[piece].size() && piece + 1 < p.sp_steps.size()) base += p.sp_steps[piece++].size();
			const uint32_t key[4] = {u.gid, u.n_steps, u.path_len, u.origin_row};
			uint64_t h = fnv(key, sizeof key);
			if (u.n_steps) h ^= fnv(p.sp_steps[piece].data() + (u.step_begin - base), sizeof(SparseStep) * u.n_steps);
			sum += h * 0x9E3779B97F4A7C15ull;
		}
		std::fprintf(stderr, "plan:   sparse units %zu checksum %016llx\n", p.sp_units.size(), (unsigned long long)sum);
		tp0 = std::chrono::steady_clock::now();
	}
	p.sp_steps.emplace_back(4, pgs::SparseStep{}); // the kernel prefetches a few steps past a unit's end
	// lanes of a warp walk different units: sort by length so that they finish together, longest first
	std::stable_sort(p.sp_units.begin(), p.sp_units.end(),
					 [](const pgs::SparseUnit &a, const pgs::SparseUnit &b) { return a.n_steps + a.path_len > b.n_steps + b.path_len; });
	lap("sort");
	e.n_sp_units = (int64_t)p.sp_units.size();
	// rows K1 really moves: these instead of every sparse row written and every sparse parent row read
	e.rows_stored += stored - e.sparse_rows_written;
	e.rows_loaded += loaded - e.sparse_rows_read;
	return PGS_OK;
}

// Allocate the row store and upload every table.
int upload_plan(Engine &e, SweepPlan &p)
{
	const int32_t N = e.n_nodes, n = e.n_genomes;
	const int64_t D = p.D, rows = p.rows;
	const int32_t ML = e.max_level;
	(void)ML;
	std::vector<DenseTask> &dense_tasks = p.dense_tasks;
	std::vector<SparseTask> &sparse_tasks = p.sparse_tasks;
	std::vector<RootTask> &root_tasks = p.root_tasks;
	int rc = 0;
	if ((rc = e.upload(e.d_walk_steps, p.walk_steps.data(), sizeof(WalkStep) * p.walk_steps.size(), "walk steps"))) return rc;
	if ((rc = e.upload(e.d_walk_off, p.walk_off.data(), sizeof(int64_t) * p.walk_off.size(), "walk units"))) return rc;
	if ((rc = e.upload(e.d_sp_units, p.sp_units.data(), sizeof(SparseUnit) * p.sp_units.size(), "sparse walk units"))) return rc;
	{ // pieces -> one device array each
		auto upload_pieces = [&](DevBuf &b, const auto &pieces, const char *what) -> int {
			size_t total = 0;
			for (const auto &v : pieces) total += v.size() * sizeof(v[0]);
			if (!b.p || b.bytes < total || b.bytes > 2 * total + ((size_t)1 << 20)) { // (kept across re-plans, as Engine::upload does)
				const int r = e.alloc(b, total, what);
				if (r) return r;
			}
			size_t at = 0;
			for (const auto &v : pieces) {
				if (v.empty()) continue;
				const cudaError_t ce = cudaMemcpyAsync(static_cast<char *>(b.p) + at, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice, e.stream);
				if (ce != cudaSuccess) return e.cuda_fail(ce, what);
				at += v.size() * sizeof(v[0]);
			}
			const cudaError_t ce = cudaStreamSynchronize(e.stream);
			return ce == cudaSuccess ? PGS_OK : e.cuda_fail(ce, what);
		};
		if ((rc = upload_pieces(e.d_sp_steps, p.sp_steps, "sparse walk steps"))) return rc;
		if ((rc = upload_pieces(e.d_sp_paths, p.sp_paths, "sparse walk paths"))) return rc;
	}
	if (e.sparse_walk && pgs::prepare_walk_sparse() != cudaSuccess)
		return e.fail(PGS_ERR_CUDA, "pgs_set_genes: the sparse walk kernel's shared-memory stack does not fit this device");
	if ((rc = e.upload(e.d_top_tasks, p.top_tasks.data(), sizeof(TopMaskTask) * p.top_tasks.size(), "top mask tasks"))) return rc;
	e.n_top_tasks = (int32_t)p.top_tasks.size();
	if ((rc = e.upload(e.d_unit_path_off, p.unit_path_off.data(), sizeof(int32_t) * p.unit_path_off.size(), "walk unit paths"))) return rc;
	if ((rc = e.upload(e.d_unit_path, p.unit_path.data(), sizeof(int64_t) * p.unit_path.size(), "walk unit path masks"))) return rc;
	if ((rc = e.upload(e.d_unit_store, p.unit_store.data(), sizeof(int64_t) * p.unit_store.size(), "walk unit stores"))) return rc;
	if (e.walk && pgs::prepare_walk_dense() != cudaSuccess)
		return e.fail(PGS_ERR_CUDA, "pgs_set_genes: the walk kernel's shared-memory stack does not fit this device");

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

} // namespace

int plan_genes(Engine &e, int64_t n_genes, const int64_t *gene_id, const int32_t *origin_node,
			   const int64_t *carrier_off, const int32_t *carrier_genome, const int32_t *all_order)
{
	SweepPlan p;
	int rc;
	const bool timing = std::getenv("PGS_PLAN_TIMING") != nullptr; // development knob: phase times on stderr
	auto t0 = std::chrono::steady_clock::now();
	auto lap = [&](const char *what) {
		if (!timing) return;
		const auto t1 = std::chrono::steady_clock::now();
		std::fprintf(stderr, "plan: %-14s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
		t0 = t1;
	};
	if ((rc = plan_gene_table(e, p, n_genes, gene_id, origin_node, carrier_off, carrier_genome, all_order))) return rc;
	lap("gene table");
	if (timing) { // FNV-1a over the joined tables: two builds of the planner can be compared without a device
		uint64_t h = 1469598103934665603ull;
		auto mix = [&](const void *q, size_t bytes) {
			const unsigned char *b = static_cast<const unsigned char *>(q);
			for (size_t i = 0; i < bytes; i++) h = (h ^ b[i]) * 1099511628211ull;
		};
		mix(e.carrier_off.data(), e.carrier_off.size() * sizeof(int64_t));
		mix(e.carrier_genome.data(), e.carrier_genome.size() * sizeof(int32_t));
		mix(e.genome_gene_count.data(), e.genome_gene_count.size() * sizeof(e.genome_gene_count[0]));
		mix(e.sp_off.data(), e.sp_off.size() * sizeof(int64_t));
		mix(e.sp_gene.data(), e.sp_gene.size() * sizeof(int64_t));
		mix(p.gn_off.data(), p.gn_off.size() * sizeof(int64_t));
		mix(p.gn_node.data(), p.gn_node.size() * sizeof(int32_t));
		mix(p.gn_rank.data(), p.gn_rank.size() * sizeof(int32_t));
		std::fprintf(stderr, "plan: gene table checksum %016llx\n", (unsigned long long)h);
		t0 = std::chrono::steady_clock::now();
	}
	if ((rc = plan_layout(e, p))) return rc;
	lap("layout");
	plan_level_tasks(e, p);
	lap("level tasks");
	plan_edge_tables(e);
	lap("edge tables");
	if ((rc = plan_walk(e, p))) return rc;
	lap("dense walk");
	if ((rc = plan_sparse_walk(e, p))) return rc;
	lap("sparse walk");
	rc = upload_plan(e, p);
	lap("upload");
	return rc;
}

} // namespace pgs
