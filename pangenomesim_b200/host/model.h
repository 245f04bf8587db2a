// model.h -- host side of pangenomesim-b200: everything the north_star keeps on the CPU.
//
// Mirrors the reference's img_model (img.h:12-88, img.cxx) with nucleotides removed: the
// coalescent, the "infinitely many genes" gain/loss process, the expected distance matrix and
// the bookkeeping that fixes the ORDER of genes and genomes in every output file.  Sequences are
// represented by ids only; the rows themselves live on the GPU behind include/pgs.h.
//
// Bit-exactness against the reference (same seed => same coalescent.newick, genefrequency.gfs,
// panmatrix.tsv, distance.mat, reproducible.seed) rests on two things:
//   * the same libstdc++ engine and distributions, called in the same order (util.cxx:48-64);
//   * the same container types with the same insert / erase / copy history, because the
//     reference picks loss victims and orders its output by unordered_map / unordered_set
//     iteration (img.cxx:308-312, :429-432).
// In rng-compat mode the draws the reference spends on nucleotides (gene.cxx:19-24, :36-58) are
// replayed without sequences so that the engine state stays in lock step.
#pragma once
#include <cstddef>
#include <cstdint>
#include <random>
#include <string>
#include <sys/types.h>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace pgshost {

enum class RngMode { Auto, Compat, Fast };

struct Tree {
	// pool layout of the reference (img.cxx:65-70): [0,n) leaves, [n,2n-1) mergers, root last
	size_t n = 0;
	std::vector<int32_t> parent, left, right;
	std::vector<double> time;     // branch length above the node (tree_node::up_time)
	std::vector<double> abs_time; // tree_node::abs_time
	size_t nodes() const { return parent.size(); }
	int32_t root() const { return (int32_t)nodes() - 1; }
};

struct GeneRecord {
	ssize_t id;
	int32_t origin;                // node where the gene appeared (root for core / start genes)
	bool all = false;              // carried by every genome, in Model::all_order (carriers empty)
	std::vector<int32_t> carriers; // genomes, in the reference's store[gene] iteration order
};

class Model {
  public:
	// parameters, defaults of img.h:15-21
	size_t gene_length = 100, num_genomes = 3, seed = 0, core_size = 4;
	double theta = 0.1, rho = 0.1, mut_rate = 0.01;
	RngMode rng_mode = RngMode::Auto;
	bool used_compat = false;

	// throws std::invalid_argument / std::out_of_range like std::stof / std::stoul do;
	// returns false for an unknown key (the reference: errx(1, "unkown key %s."))
	bool parse_param(const std::string &key, const std::string &value);
	std::string parameters() const; // reproducible.seed body, img.cxx:50-63

	void simulate();

	// results
	Tree tree;
	std::vector<GeneRecord> genes;       // in gene_ids() order (img.cxx:429-432): the output order
	std::vector<int64_t> ref_core, ref_acc; // indices into `genes`, reference order (img.cxx:371-389)
	std::vector<ssize_t> ref_core_ids, ref_acc_ids; // all reference ids incl. genes carried by no genome
	std::vector<int32_t> all_order;      // store[] iteration order of a gene carried by every genome
	std::vector<int64_t> genome_gene_count;

	// expected distance between genomes i and every j (reference create_distmatrix, img.cxx:121-165,
	// O(m^3) Floyd-Warshall there; here one O(n) tree walk per row)
	void distance_row(size_t i, std::vector<double> &row) const;

	std::default_random_engine rng; // the reference's global RNG (global.h:6)

  private:
	void create_coalescent();
	void simulate_flat();        // rng-fast: the gain/loss process on flat arrays
	void build_distance_graph();
	void replay_root_draws();
	void replay_mutate(double rate);
	void build_children_lists();
	std::vector<int64_t> adj_off; // undirected tree in CSR form, for distance rows
	std::vector<int32_t> adj_to;
	std::vector<double> adj_w;
};

std::string genome_name(ssize_t index); // util.cxx:15-27 with the >= 1000 fix

} // namespace pgshost
