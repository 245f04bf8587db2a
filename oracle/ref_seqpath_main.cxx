// oracle/ref_seqpath_main.cxx -- TEST/BENCH INFRASTRUCTURE, not product code.
//
// Timing harness "T2" of SURVEY.md §8(d): links the UNMODIFIED reference objects
// (gene.cxx, img.cxx, util.cxx compiled from /root/reference/src by oracle/Makefile) and drives
// the reference's own sequence path for configurations its CLI cannot run (n >= 1000 aborts in
// util.cxx:21-22; the O(m^3) Floyd-Warshall of img.cxx:147-154).
//
// What it calls, all of it reference code with external linkage:
//   create_coalescent(n)            img.cxx:71-116
//   gene::gene(len, -1, id)         gene.cxx:13-25   (random root)
//   gene::mutate(time * mut_rate)   gene.cxx:27-61   (one call per edge, DFS order)
//   tree_node::traverse(pre, process, post)   img.h:316-329
// The DFS below has the same shape as img_model::sim_core_gene (img.cxx:203-232) but keeps no
// leaf store: leaves are only counted (and optionally folded into a checksum so the work cannot
// be optimised away).
//
// usage: ref_seqpath <num-genomes> <gene-length> <genes> <mut-rate> <seed> [tree-dump-file]
// prints one JSON line.
#include "gene.h"
#include "global.h"
#include "img.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

// the three globals the reference objects expect (defined in pangenomesim.cxx:26-28 there)
std::string OUT_DIR = "./";
std::default_random_engine RNG;
bool VERBOSE = false;

std::vector<tree_node> create_coalescent(size_t n);

int main(int argc, char **argv)
{
	if (argc < 6) {
		std::fprintf(stderr, "usage: %s n L genes mut_rate seed [tree-dump]\n", argv[0]);
		return 2;
	}
	const size_t n = std::strtoull(argv[1], nullptr, 10);
	const size_t len = std::strtoull(argv[2], nullptr, 10);
	const size_t genes = std::strtoull(argv[3], nullptr, 10);
	const double mut_rate = std::strtof(argv[4], nullptr); // the reference parses with stof
	const size_t seed = std::strtoull(argv[5], nullptr, 10);

	RNG = std::default_random_engine(seed);
	auto pool = create_coalescent(n);
	const tree_node &root = pool.back();

	if (argc > 6) { // dump parent index + branch time per pool slot, for cross-checking the host port
		FILE *f = std::fopen(argv[6], "w");
		if (!f) return 3;
		for (size_t i = 0; i < pool.size(); i++) {
			long par = pool[i].has_parent() ? (long)(&pool[i].get_parent() - &pool[0]) : -1;
			std::fprintf(f, "%zu %ld %.17g %zd\n", i, par, pool[i].get_time(), pool[i].get_index());
		}
		std::fclose(f);
	}

	unsigned long long checksum = 0, leaves_seen = 0, edges_done = 0;
	const auto t0 = std::chrono::steady_clock::now();
	for (size_t g = 0; g < genes; g++) {
		std::vector<gene> stack;
		stack.reserve(64);
		stack.push_back(gene((ssize_t)len, -1, (ssize_t)g));
		root.traverse(
			[&](const tree_node &self) {
				if (self.has_parent()) {
					stack.push_back(stack.back().mutate(self.get_time() * mut_rate));
					edges_done++;
				}
			},
			[&](const tree_node &self) {
				if (self.is_leaf()) {
					const std::string &s = stack.back().get_nucl();
					checksum += (unsigned char)s[s.size() / 2] + s.size();
					leaves_seen++;
				}
			},
			[&](const tree_node &) { stack.pop_back(); });
	}
	const auto t1 = std::chrono::steady_clock::now();
	const double sec = std::chrono::duration<double>(t1 - t0).count();
	const double site_edges = (double)edges_done * (double)len;
	const double leaf_nt = (double)leaves_seen * (double)len;
	std::printf("{\"n\": %zu, \"gene_length\": %zu, \"genes\": %zu, \"mut_rate\": %g, \"seed\": %zu, "
				"\"seconds\": %.6f, \"site_edges\": %.0f, \"leaf_nt\": %.0f, "
				"\"site_edges_per_s\": %.6g, \"leaf_nt_per_s\": %.6g, \"checksum\": %llu}\n",
				n, len, genes, mut_rate, seed, sec, site_edges, leaf_nt, site_edges / sec,
				leaf_nt / sec, checksum);
	return 0;
}
