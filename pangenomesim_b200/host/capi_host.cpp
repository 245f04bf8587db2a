// capi_host.cpp -- C entry points of the host model for Python callers (bench.py, tests): the
// coalescent and the IMG gain/loss process exactly as the host program runs them (reference
// create_coalescent, img.cxx:71-116, and img_model::simulate, img.cxx:237-369, same libstdc++
// engine and distributions), so benchmarks and tests run on the reference's own trees and gene
// tables.  No CUDA in here.
#include "model.h"

#include <cstring>
#include <exception>
#include <string>

extern "C" {

// parent[2n-1] (root = -1), time[2n-1] (branch length above the node, 0 for the root).
// Returns 0, or -1 on bad arguments.
int pgsh_coalescent(uint64_t n, uint64_t seed, int32_t *parent, double *time)
{
	if (n < 1 || !parent || !time) return -1;
	pgshost::Model m;
	m.num_genomes = n;
	m.core_size = 0;
	m.theta = 0.0; // no gain events: with theta = 0 the waiting time is infinite
	m.rho = 1.0;
	m.rng_mode = pgshost::RngMode::Fast;
	m.seed = seed;
	m.rng = std::default_random_engine(seed);
	if (seed == 0) return -1;
	m.simulate();
	std::memcpy(parent, m.tree.parent.data(), sizeof(int32_t) * m.tree.nodes());
	std::memcpy(time, m.tree.time.data(), sizeof(double) * m.tree.nodes());
	return 0;
}

// ---- the whole host model: parameters in (the -p KEY=VALUE strings of the CLI), tables out

struct pgsh_model {
	pgshost::Model m;
	std::string error;
	// flattened gene table (the arguments of pgs_set_genes)
	std::vector<int64_t> gene_id, carrier_off;
	std::vector<int32_t> origin, carriers;
	std::vector<double> rate;
	std::vector<int32_t> leaf_genome;
};

pgsh_model *pgsh_model_create(void)
{
	return new (std::nothrow) pgsh_model();
}

void pgsh_model_destroy(pgsh_model *h)
{
	delete h;
}

const char *pgsh_model_error(const pgsh_model *h)
{
	return h ? h->error.c_str() : "null model";
}

// key/value as on the command line (img.cxx:16-48).  0, or -1 (unknown key / bad number).
int pgsh_model_param(pgsh_model *h, const char *key, const char *value)
{
	if (!h || !key || !value) return -1;
	try {
		if (!h->m.parse_param(key, value)) {
			h->error = std::string("unkown key ") + key + ".";
			return -1;
		}
	} catch (const std::exception &ex) {
		h->error = ex.what();
		return -1;
	}
	return 0;
}

// rng_mode: 0 = auto, 1 = compat (replay the reference's nucleotide draws), 2 = fast.
int pgsh_model_simulate(pgsh_model *h, int rng_mode)
{
	if (!h) return -1;
	try {
		pgshost::Model &m = h->m;
		m.rng_mode = rng_mode == 1 ? pgshost::RngMode::Compat : rng_mode == 2 ? pgshost::RngMode::Fast : pgshost::RngMode::Auto;
		m.simulate();
		const size_t N = m.tree.nodes(), n = m.num_genomes, G = m.genes.size();
		h->rate.assign(N, 0.0);
		h->leaf_genome.assign(N, -1);
		for (size_t v = 0; v < N; v++) {
			h->rate[v] = m.tree.time[v] * m.mut_rate; // img.cxx:218,286
			if (v < n) h->leaf_genome[v] = (int32_t)v;
		}
		h->gene_id.resize(G);
		h->origin.resize(G);
		h->carrier_off.assign(G + 1, 0);
		h->carriers.clear();
		for (size_t k = 0; k < G; k++) { // as host/main.cpp hands the table to pgs_set_genes
			const pgshost::GeneRecord &g = m.genes[k];
			h->gene_id[k] = g.id;
			h->origin[k] = g.origin;
			if (g.all && g.origin == m.tree.root())
				h->carriers.push_back(-1); // PGS_ALL_GENOMES
			else if (g.all)
				h->carriers.insert(h->carriers.end(), m.all_order.begin(), m.all_order.end());
			else
				h->carriers.insert(h->carriers.end(), g.carriers.begin(), g.carriers.end());
			h->carrier_off[k + 1] = (int64_t)h->carriers.size();
		}
	} catch (const std::exception &ex) {
		h->error = ex.what();
		return -1;
	}
	return 0;
}

// sizes: [0] nodes, [1] genomes, [2] genes, [3] carrier entries, [4] reference core genes,
// [5] reference accessory genes, [6] 1 if the nucleotide draws were replayed (rng-compat)
void pgsh_model_sizes(const pgsh_model *h, int64_t out[7])
{
	const pgshost::Model &m = h->m;
	out[0] = (int64_t)m.tree.nodes();
	out[1] = (int64_t)m.num_genomes;
	out[2] = (int64_t)m.genes.size();
	out[3] = (int64_t)h->carriers.size();
	out[4] = (int64_t)m.ref_core.size();
	out[5] = (int64_t)m.ref_acc.size();
	out[6] = m.used_compat ? 1 : 0;
}

// Every pointer may be NULL (skipped).  Array sizes as reported by pgsh_model_sizes.
void pgsh_model_tables(const pgsh_model *h, int32_t *parent, double *rate, int32_t *leaf_genome, int64_t *gene_id,
					   int32_t *origin, int64_t *carrier_off, int32_t *carriers, int32_t *all_order, int64_t *ref_core,
					   int64_t *ref_acc, int64_t *genome_gene_count)
{
	const pgshost::Model &m = h->m;
	auto put = [](auto *dst, const auto &v) {
		if (dst && !v.empty()) std::memcpy(dst, v.data(), sizeof(v[0]) * v.size());
	};
	put(parent, m.tree.parent);
	put(rate, h->rate);
	put(leaf_genome, h->leaf_genome);
	put(gene_id, h->gene_id);
	put(origin, h->origin);
	put(carrier_off, h->carrier_off);
	put(carriers, h->carriers);
	put(all_order, m.all_order);
	put(ref_core, m.ref_core);
	put(ref_acc, m.ref_acc);
	put(genome_gene_count, m.genome_gene_count);
}

// expected distance from genome i to every genome (distance.mat row, img.cxx:121-165)
void pgsh_model_distance_row(const pgsh_model *h, int64_t i, double *row)
{
	std::vector<double> r;
	h->m.distance_row((size_t)i, r);
	std::memcpy(row, r.data(), sizeof(double) * r.size());
}

} // extern "C"
