#include "model.h"

#include <algorithm>
#include <cmath>
#include <functional>
#include <sstream>
#include <stdexcept>

namespace pgshost {

std::string genome_name(ssize_t index)
{
	// util.cxx:15-27 pads to ceil(log10(101)) = 3 digits and throws std::length_error when the
	// number has more (genome 1000 onwards); here longer numbers are simply not padded.
	if (index == -1) return "genomeref";
	std::string num = std::to_string(index + 1);
	if (num.size() < 3) num.insert(0, 3 - num.size(), '0');
	return "genome" + num;
}

bool Model::parse_param(const std::string &key, const std::string &value)
{
	// img.cxx:16-48: floats go through std::stof (float precision), `seed` re-seeds at parse time
	if (key == "theta") {
		theta = std::stof(value);
	} else if (key == "rho") {
		rho = std::stof(value);
	} else if (key == "core-size") {
		core_size = std::stoul(value);
	} else if (key == "gene-length") {
		gene_length = std::stoul(value);
	} else if (key == "num-genomes") {
		num_genomes = std::stoul(value);
	} else if (key == "seed") {
		seed = std::stoul(value);
		rng = std::default_random_engine(seed);
	} else if (key == "mut-rate") {
		mut_rate = std::stof(value);
	} else {
		return false;
	}
	return true;
}

std::string Model::parameters() const
{
	std::stringstream s;
	s << "--param core-size=" << core_size << "\n";
	s << "--param gene-length=" << gene_length << "\n";
	s << "--param mut-rate=" << mut_rate << "\n";
	s << "--param num-genomes=" << num_genomes << "\n";
	s << "--param rho=" << rho << "\n";
	s << "--param seed=" << seed << "\n";
	s << "--param theta=" << theta << "\n";
	return s.str();
}

// ---- draws, in the reference's wrappers' own words (util.cxx:48-64)

static size_t draw_int(std::default_random_engine &rng, size_t lower, size_t upper)
{
	return std::uniform_int_distribution<size_t>(lower, upper - 1)(rng);
}
static double draw_exp(std::default_random_engine &rng, double lambda)
{
	return std::exponential_distribution<double>(lambda)(rng);
}
static size_t draw_poisson(std::default_random_engine &rng, double mean)
{
	return std::poisson_distribution<size_t>(mean)(rng);
}

// Kingman coalescent, img.cxx:71-116: i-th merger takes two random members of the shrinking
// `active` table; waiting times Exp(k(k-1)/2); branch = parent abs time - own abs time.
void Model::create_coalescent()
{
	const size_t n = num_genomes, m = 2 * n - 1;
	tree.n = n;
	tree.parent.assign(m, -1);
	tree.left.assign(m, -1);
	tree.right.assign(m, -1);
	tree.time.assign(m, 0.0);
	tree.abs_time.assign(m, 0.0);
	std::vector<int32_t> active(m);
	for (size_t i = 0; i < m; i++) active[i] = (int32_t)i;
	for (size_t i = 0; i + 1 < n; i++) {
		const int32_t p = (int32_t)(n + i);
		const size_t a = draw_int(rng, 0, n - i);
		tree.left[p] = active[a];
		tree.parent[active[a]] = p;
		active[a] = active[n - i - 1];
		const size_t b = draw_int(rng, 0, n - i - 1);
		tree.right[p] = active[b];
		tree.parent[active[b]] = p;
		active[b] = p;
	}
	double t = 0.0;
	size_t k = n;
	for (size_t i = 0; i + 1 < n; i++, k--) {
		const size_t pairs = k * (k - 1) / 2;
		t += draw_exp(rng, pairs);
		tree.abs_time[n + i] = t;
	}
	for (size_t v = 0; v + 1 < m; v++) tree.time[v] = tree.abs_time[tree.parent[v]] - tree.abs_time[v];
}

// The L draws of gene::gene(len, ..) (gene.cxx:19-24), without the sequence.
void Model::replay_root_draws()
{
	std::uniform_int_distribution<int> base(0, 3);
	for (size_t i = 0; i < gene_length; i++) (void)base(rng);
}

// The engine traffic of one gene::mutate(rate) (gene.cxx:36-58), without the sequence: site
// selection runs on a COPY of the engine (std::bind copies, gene.cxx:37), each hit costs one
// uniform_int(0,2) draw on the real engine.
void Model::replay_mutate(double rate)
{
	auto site = std::bind(std::uniform_real_distribution<double>(0, 1), rng);
	std::uniform_int_distribution<int> other(0, 2);
	double nucleotides = (double)gene_length;
	double mutations = rate * nucleotides;
	for (size_t i = 0; i < gene_length; i++) {
		if (site() < mutations / nucleotides) {
			(void)other(rng);
			mutations--;
		}
		nucleotides--;
	}
}

namespace {

// tree_node::traverse(pre, process, post), img.h:316-329, over the array tree
template <typename Pre, typename Process, typename Post>
void traverse(const Tree &t, Pre pre, Process process, Post post)
{
	struct Frame {
		int32_t node;
		int state;
	};
	std::vector<Frame> stack;
	stack.push_back({t.root(), 0});
	pre(t.root());
	while (!stack.empty()) {
		Frame &f = stack.back();
		const int32_t v = f.node;
		if (f.state == 0) {
			f.state = 1;
			if (t.left[v] >= 0) {
				const int32_t c = t.left[v];
				stack.push_back({c, 0});
				pre(c);
			}
		} else if (f.state == 1) {
			f.state = 2;
			process(v);
			if (t.right[v] >= 0) {
				const int32_t c = t.right[v];
				stack.push_back({c, 0});
				pre(c);
			}
		} else {
			post(v);
			stack.pop_back();
		}
	}
}

} // namespace

void Model::simulate()
{
	if (num_genomes < 1) throw std::invalid_argument("num-genomes must be at least 1");
	if (seed == 0) { // img.cxx:239-243
		std::random_device rd;
		seed = rd();
		rng = std::default_random_engine(seed);
	}
	const size_t n = num_genomes;
	genes.clear(); // a second simulate() on the same Model starts from scratch
	ref_core.clear();
	ref_acc.clear();
	ref_core_ids.clear();
	ref_acc_ids.clear();
	all_order.clear();
	genome_gene_count.clear();
	create_coalescent(); // first consumer of the engine: coalescent.newick is bit-exact in every mode
	const Tree &t = tree;

	bool compat = rng_mode == RngMode::Compat;
	if (rng_mode == RngMode::Auto) {
		const double work = (double)core_size * (double)(2 * n - 2) * (double)gene_length;
		compat = work <= 1.0e9;
	}
	used_compat = compat;

	if (!compat) { // nothing to keep in lock step with: flat arrays instead of the reference's hash maps
		simulate_flat();
		build_distance_graph();
		return;
	}

	using id_map = std::unordered_map<ssize_t, char>; // stands for unordered_map<ssize_t, gene>
	std::unordered_set<ssize_t> gene_id_list;          // img.h:32
	std::unordered_map<ssize_t, id_map> store;         // img.h:34-35, accessory genes only
	std::unordered_map<ssize_t, int32_t> origin_of;

	// ---- core genes, img.cxx:253-255 -> sim_core_gene :203-232.  Every core gene reaches every
	// leaf in the same order, so one inner map stands for all of them.
	id_map core_carriers;
	traverse(t, [](int32_t) {}, [&](int32_t v) { if (t.left[v] < 0) core_carriers[(ssize_t)v] = 1; }, [](int32_t) {});
	all_order.clear();
	for (const auto &kv : core_carriers) all_order.push_back((int32_t)kv.first);
	for (size_t g = 0; g < core_size; g++) {
		if (compat) {
			replay_root_draws();
			traverse(t, [&](int32_t v) { if (t.parent[v] >= 0) replay_mutate(t.time[v] * mut_rate); },
					 [](int32_t) {}, [](int32_t) {});
		}
		gene_id_list.insert((ssize_t)g);
		origin_of[(ssize_t)g] = t.root();
		ref_core_ids.push_back((ssize_t)g);
	}

	// ---- accessory genes, img.cxx:257-333
	id_map start;
	const size_t start_size = draw_poisson(rng, theta / rho);
	start.reserve(start_size);
	for (size_t k = 0; k < start_size; k++) {
		const ssize_t id = (ssize_t)(k + core_size);
		if (compat) replay_root_draws();
		start.insert(std::make_pair(id, (char)1));
		origin_of[id] = t.root();
	}
	ssize_t next_id = (ssize_t)(start_size + core_size);
	std::vector<ssize_t> acc_ref; // ref_acc before pruning
	for (const auto &kv : start) acc_ref.push_back(kv.first);

	std::vector<id_map> stack;
	stack.reserve(n);
	stack.push_back(start);
	traverse(
		t,
		[&](int32_t v) {
			if (t.parent[v] < 0) return;
			const double branch = t.time[v];
			// gene::bulk_mutate, gene.h:26-40: fresh map, reserve, insert in iteration order
			id_map neu;
			{
				const id_map &top = stack.back();
				neu.reserve(top.size());
				for (const auto &kv : top) {
					if (compat) replay_mutate(mut_rate * branch);
					neu.insert(neu.end(), kv);
				}
			}
			double time = 0.0;
			while (time < branch) {
				const double to_go = branch - time;
				const double gain = draw_exp(rng, theta);
				const double loss = draw_exp(rng, rho * neu.size());
				if (gain > to_go && loss > to_go) break;
				if (gain < loss || neu.empty()) {
					const ssize_t id = next_id++;
					if (compat) replay_root_draws();
					neu[id] = 1;
					acc_ref.push_back(id);
					origin_of[id] = v;
					time += gain;
				} else {
					const size_t loser = draw_int(rng, 0, neu.size());
					auto it = neu.begin();
					std::advance(it, loser);
					neu.erase(it);
					time += loss;
				}
			}
			stack.push_back(neu);
		},
		[&](int32_t v) {
			if (t.left[v] >= 0) return;
			for (const auto &kv : stack.back()) { // img_model::add, img.h:44-49
				gene_id_list.insert(kv.first);
				store[kv.first][(ssize_t)v] = 1;
			}
		},
		[&](int32_t) { stack.pop_back(); });

	// ---- reference lists, img.cxx:355-368
	auto carriers_of = [&](ssize_t id) -> size_t {
		if ((size_t)id < core_size) return n;
		auto it = store.find(id);
		return it == store.end() ? 0 : it->second.size();
	};
	acc_ref.erase(std::remove_if(acc_ref.begin(), acc_ref.end(), [&](ssize_t id) { return carriers_of(id) == 0; }),
				  acc_ref.end());
	for (ssize_t id : acc_ref)
		if (carriers_of(id) == n) ref_core_ids.push_back(id);
	acc_ref.erase(std::remove_if(acc_ref.begin(), acc_ref.end(), [&](ssize_t id) { return carriers_of(id) == n; }),
				  acc_ref.end());
	ref_acc_ids = acc_ref;

	// ---- gene table in gene_ids() order (img.cxx:429-432)
	genes.clear();
	std::unordered_map<ssize_t, int64_t> index_of;
	genome_gene_count.assign(n, 0);
	for (ssize_t id : gene_id_list) {
		GeneRecord r;
		r.id = id;
		r.origin = origin_of[id];
		if ((size_t)id < core_size) {
			r.all = true;
			for (auto &c : genome_gene_count) c++;
		} else {
			for (const auto &kv : store[id]) r.carriers.push_back((int32_t)kv.first);
			for (int32_t g : r.carriers) genome_gene_count[g]++;
		}
		index_of[id] = (int64_t)genes.size();
		genes.push_back(std::move(r));
	}
	ref_core.clear();
	ref_acc.clear();
	for (ssize_t id : ref_core_ids) ref_core.push_back(index_of.at(id));
	for (ssize_t id : ref_acc_ids) ref_acc.push_back(index_of.at(id));

	build_distance_graph();
}

// The IMG gain/loss process (img.cxx:257-333) on flat arrays, for runs where the reference's draws are
// not replayed (rng-fast): the same process -- Poisson(theta/rho) start set, competing exponential
// clocks for gain (theta) and loss (rho per gene) along every branch, a uniformly chosen victim --
// drawn from the same engine, but with vectors instead of unordered_map copies per node (config 3:
// 2.5 s -> 0.1 s of host time).  Orders are this program's own: genes by ascending id, carriers by
// ascending genome (the reference's orders are hash-table iteration orders, which only rng-compat
// runs reproduce).
void Model::simulate_flat()
{
	const Tree &t = tree;
	const size_t n = num_genomes;
	std::vector<int32_t> origin_of;               // per accessory gene (id - core_size)
	std::vector<std::vector<int32_t>> carriers;   // per accessory gene: genomes, ascending after the sort below
	auto new_gene = [&](int32_t node) {
		origin_of.push_back(node);
		carriers.emplace_back();
		return (ssize_t)(core_size + origin_of.size() - 1);
	};
	std::vector<ssize_t> start;
	const size_t start_size = draw_poisson(rng, theta / rho);
	for (size_t k = 0; k < start_size; k++) start.push_back(new_gene(t.root()));
	std::vector<std::vector<ssize_t>> stack;
	stack.push_back(std::move(start));
	traverse(
		t,
		[&](int32_t v) {
			if (t.parent[v] < 0) return;
			const double branch = t.time[v];
			std::vector<ssize_t> neu = stack.back();
			double time = 0.0;
			while (time < branch) {
				const double to_go = branch - time;
				const double gain = draw_exp(rng, theta);
				const double loss = draw_exp(rng, rho * neu.size());
				if (gain > to_go && loss > to_go) break;
				if (gain < loss || neu.empty()) {
					neu.push_back(new_gene(v));
					time += gain;
				} else {
					const size_t loser = draw_int(rng, 0, neu.size());
					neu[loser] = neu.back();
					neu.pop_back();
					time += loss;
				}
			}
			stack.push_back(std::move(neu));
		},
		[&](int32_t v) {
			if (t.left[v] >= 0) return;
			for (ssize_t id : stack.back()) carriers[(size_t)id - core_size].push_back(v);
		},
		[&](int32_t) { stack.pop_back(); });

	all_order.resize(n);
	for (size_t i = 0; i < n; i++) all_order[i] = (int32_t)i;
	genome_gene_count.assign(n, (int64_t)core_size);
	for (size_t g = 0; g < core_size; g++) {
		GeneRecord r;
		r.id = (ssize_t)g;
		r.origin = t.root();
		r.all = true;
		ref_core.push_back((int64_t)genes.size());
		ref_core_ids.push_back(r.id);
		genes.push_back(std::move(r));
	}
	std::vector<int64_t> acc_index;
	for (size_t a = 0; a < carriers.size(); a++) {
		if (carriers[a].empty()) continue; // lost everywhere (img.cxx:355-358)
		GeneRecord r;
		r.id = (ssize_t)(core_size + a);
		r.origin = origin_of[a];
		std::sort(carriers[a].begin(), carriers[a].end());
		for (int32_t g : carriers[a]) genome_gene_count[g]++;
		const bool everywhere = carriers[a].size() == n;
		r.carriers = std::move(carriers[a]);
		// an accessory gene that no genome has lost joins the core reference (img.cxx:361-368)
		(everywhere ? ref_core : ref_acc).push_back((int64_t)genes.size());
		(everywhere ? ref_core_ids : ref_acc_ids).push_back(r.id);
		genes.push_back(std::move(r));
	}
}

void Model::build_distance_graph()
{
	// undirected adjacency (CSR) with edge weight time * mut_rate (img.cxx:142) for distance rows
	const Tree &t = tree;
	const size_t N = t.nodes();
	adj_off.assign(N + 1, 0);
	for (size_t v = 0; v + 1 < N; v++) {
		adj_off[v + 1]++;
		adj_off[(size_t)t.parent[v] + 1]++;
	}
	for (size_t v = 0; v < N; v++) adj_off[v + 1] += adj_off[v];
	adj_to.assign(adj_off[N], 0);
	adj_w.assign(adj_off[N], 0.0);
	std::vector<int64_t> fill(adj_off.begin(), adj_off.end() - 1);
	for (size_t v = 0; v + 1 < N; v++) {
		const double w = t.time[v] * mut_rate;
		const size_t p = (size_t)t.parent[v];
		adj_to[fill[v]] = (int32_t)p;
		adj_w[fill[v]++] = w;
		adj_to[fill[p]] = (int32_t)v;
		adj_w[fill[p]++] = w;
	}
}

// Expected distances from genome i to every genome: the sum of the branch weights along the
// unique path, accumulated outwards from i (the value the reference's Floyd-Warshall converges to,
// img.cxx:147-154; identical as printed, tests/test_host_cli.py).
void Model::distance_row(size_t i, std::vector<double> &row) const
{
	const size_t n = num_genomes;
	row.assign(n, 0.0);
	struct Item {
		int32_t node, from;
		double dist;
	};
	thread_local std::vector<Item> todo;
	todo.clear();
	todo.push_back({(int32_t)i, -1, 0.0});
	while (!todo.empty()) {
		const Item it = todo.back();
		todo.pop_back();
		if ((size_t)it.node < n) row[it.node] = it.dist;
		for (int64_t k = adj_off[it.node]; k < adj_off[it.node + 1]; k++)
			if (adj_to[k] != it.from) todo.push_back({adj_to[k], it.node, it.dist + adj_w[k]});
	}
}

} // namespace pgshost
