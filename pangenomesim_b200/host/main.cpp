// pangenomesim-b200 -- the host program: same command line and the same small text outputs as
// the reference's main (pangenomesim.cxx:30-327); every nucleotide is produced by the CUDA engine
// behind include/pgs.h.
//
//   reference step (pangenomesim.cxx)              here
//   :43-98   getopt_long + -p KEY=VALUE regex       same parser, same keys, same error texts
//   :103     model.simulate()                       pgshost::Model::simulate (ids only) + pgs_sweep
//   :105-115 reproducible.seed                      host, verbatim
//   :117-157 ref/core/accessory/genomeNNN.fasta     pgs_write_outputs (GPU formatter)
//   :159-234 gfs, newick, distance.mat              host, verbatim formats
//   :236-293 alignment.maf                          pgs_write_outputs (GPU formatter)
//   :295-324 panmatrix.tsv                          host, verbatim
#include "model.h"
#include "pgs.h"

#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <regex>
#include <sstream>
#include <thread>

extern "C" {
#include <err.h>
#include <getopt.h>
#include <sys/stat.h>
#include <unistd.h>
}

#define PGS_HOST_VERSION "2.0" // printed in reproducible.seed exactly like the reference's VERSION

using pgshost::Model;

static bool verbose = false;

[[noreturn]] static void usage(int code)
{
	static const char text[] =
		"usage: pangenomesim-b200 [OPTIONS...]\n"
		"Simulate a pangenome according to the 'infinitely many genes' model;\n"
		"sequences evolve on the GPU (B200, sm_100a).\n\n"
		"Options:\n"
		"  -o, --out-dir DIR        The directory to write files to\n"
		"  -p, --param KEY=VALUE    Set simulation parameter\n"
		"  -v, --verbose            Print additional information\n"
		"      --help               Display help and exit\n"
		"      --version            Output version information and exit\n"
		"      --rng-compat         Replay the reference's nucleotide draws on the host engine so that\n"
		"                           genefrequency.gfs / panmatrix.tsv match the reference bit for bit\n"
		"      --rng-fast           Never replay (default for large runs)\n"
		"      --no-sequences       Host-side files only (no GPU needed)\n"
		"      --device N           CUDA device ordinal\n\n"
		"Simulation Parameters:\n"
		"  core-size=INT            Minimum size of the core genome\n"
		"  gene-length=INT          Length of a gene\n"
		"  mut-rate=FLOAT           Substitution rate\n"
		"  num-genomes=INT          Number of genomes\n"
		"  rho=FLOAT                Rate of gene loss\n"
		"  seed=INT                 Seed for random number generator (use 0 for random)\n"
		"  theta=FLOAT              Rate of gene gain\n";
	std::fputs(text, code == EXIT_SUCCESS ? stdout : stderr);
	std::exit(code);
}

static void make_path(const std::string &path)
{
	// util.cxx:29-46: mkdir -p
	if (path.empty()) return;
	if (::mkdir(path.c_str(), 0755) == 0 || errno == EEXIST) return;
	if (errno != ENOENT) err(errno, "%s", path.c_str());
	const auto cut = path.find_last_of('/');
	if (cut == std::string::npos || cut == 0) err(errno, "%s", path.c_str());
	make_path(path.substr(0, cut));
	if (::mkdir(path.c_str(), 0755) != 0 && errno != EEXIST) err(errno, "%s", path.c_str());
}

static std::ofstream open_out(const std::string &name)
{
	std::ofstream f(name);
	if (!f.good()) err(errno, "%s: couldn't write to file", name.c_str());
	return f;
}

static void write_newick(const Model &m, std::ostream &out)
{
	// pangenomesim.cxx:186-209
	const pgshost::Tree &t = m.tree;
	struct Frame {
		int32_t node;
		int state;
	};
	std::vector<Frame> stack{{t.root(), 0}};
	while (!stack.empty()) {
		Frame &f = stack.back();
		const int32_t v = f.node;
		const bool branch = t.left[v] >= 0;
		if (f.state == 0) {
			f.state = 1;
			if (branch) {
				out << "(";
				stack.push_back({t.left[v], 0});
			}
		} else if (f.state == 1) {
			f.state = 2;
			if (branch) {
				out << ",";
				stack.push_back({t.right[v], 0});
			} else {
				out << pgshost::genome_name(v);
			}
		} else {
			if (branch) out << ")";
			if (t.time[v] != 0.0) out << ":" << t.time[v];
			stack.pop_back();
		}
	}
	out << ";\n";
}

static void write_distance_matrix(const Model &m, const std::string &name)
{
	// pangenomesim.cxx:214-234: "  " + setw(8) + setprecision(4) per entry == "  %8.4g"
	const size_t n = m.num_genomes;
	std::ofstream out = open_out(name);
	out << n << "\n";
	const size_t batch = 256;
	unsigned workers = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
	if (n < 512) workers = 1;
	for (size_t r0 = 0; r0 < n; r0 += batch) {
		const size_t r1 = std::min(n, r0 + batch);
		std::vector<std::string> lines(r1 - r0);
		std::atomic<size_t> next{r0};
		auto work = [&]() {
			std::vector<double> row;
			std::string line;
			char buf[64];
			for (size_t i = next++; i < r1; i = next++) {
				m.distance_row(i, row);
				line = pgshost::genome_name((ssize_t)i);
				for (double d : row) {
					const int len = std::snprintf(buf, sizeof buf, "  %8.4g", d);
					line.append(buf, (size_t)len);
				}
				line += "\n";
				lines[i - r0].swap(line);
			}
		};
		std::vector<std::thread> pool;
		for (unsigned w = 1; w < workers; w++) pool.emplace_back(work);
		work();
		for (auto &th : pool) th.join();
		for (const auto &l : lines) out << l;
	}
	if (!out.good()) err(errno, "%s: couldn't write to file", name.c_str());
}

int main(int argc, char *argv[])
{
	static const struct option long_options[] = {
		{"version", no_argument, nullptr, 0},      {"help", no_argument, nullptr, 0},
		{"param", required_argument, nullptr, 'p'}, {"out-dir", required_argument, nullptr, 'o'},
		{"verbose", no_argument, nullptr, 'v'},     {"rng-compat", no_argument, nullptr, 0},
		{"rng-fast", no_argument, nullptr, 0},      {"no-sequences", no_argument, nullptr, 0},
		{"device", required_argument, nullptr, 0},  {nullptr, 0, nullptr, 0}};

	Model model;
	std::string out_dir = "./";
	bool sequences = true;
	int device = -1;

	for (;;) {
		int index = 0;
		const int c = getopt_long(argc, argv, "o:p:v", long_options, &index);
		if (c == -1) break;
		switch (c) {
			case 0: {
				const std::string name = long_options[index].name;
				if (name == "help") usage(EXIT_SUCCESS);
				if (name == "version") {
					std::printf("pangenomesim-b200 (pangenomesim " PGS_HOST_VERSION " compatible), pgs ABI %d\n",
								pgs_abi_version());
					return 0;
				}
				if (name == "rng-compat") model.rng_mode = pgshost::RngMode::Compat;
				if (name == "rng-fast") model.rng_mode = pgshost::RngMode::Fast;
				if (name == "no-sequences") sequences = false;
				if (name == "device") device = std::atoi(optarg);
				break;
			}
			case 'o': out_dir = std::string(optarg) + "/"; break;
			case 'p': {
				// getsubopt-like: KEY[=VALUE] separated by commas (pangenomesim.cxx:68-90)
				std::string rest = optarg;
				const std::regex item("^([\\w\\-]+)(?:=([a-z_0-9.]+))?,?");
				std::smatch m;
				while (std::regex_search(rest, m, item)) {
					const std::string key = m[1], value = m.size() == 3 ? std::string(m[2]) : std::string();
					if (!model.parse_param(key, value)) errx(1, "unkown key %s.", key.c_str());
					rest = m.suffix();
				}
				if (!rest.empty()) errx(1, "invalid parameter %s", optarg);
				break;
			}
			case 'v': verbose = true; break;
			default: usage(EXIT_FAILURE);
		}
	}

	make_path(out_dir);
	const auto t0 = std::chrono::steady_clock::now();
	model.simulate();
	const auto t1 = std::chrono::steady_clock::now();
	const size_t n = model.num_genomes;

	{ // reproducible.seed, pangenomesim.cxx:105-115
		std::ofstream f = open_out(out_dir + "reproducible.seed");
		f << "pangenomesim " PGS_HOST_VERSION << "\n\n";
		f << "full options:\n";
		f << model.parameters() << std::endl;
	}
	{ // genefrequency.gfs, pangenomesim.cxx:159-179
		std::vector<size_t> gfs(n + 1);
		for (const auto &g : model.genes) gfs.at(g.all ? n : g.carriers.size())++;
		std::ofstream f = open_out(out_dir + "genefrequency.gfs");
		for (size_t k = 1; k <= n; k++) f << gfs[k] << " ";
		f << std::endl;
	}
	{ // coalescent.newick
		std::ofstream f = open_out(out_dir + "coalescent.newick");
		write_newick(model, f);
	}
	write_distance_matrix(model, out_dir + "distance.mat");
	{ // panmatrix.tsv, pangenomesim.cxx:295-324
		std::ofstream f = open_out(out_dir + "panmatrix.tsv");
		f << "GeneNumber";
		for (const auto &g : model.genes) f << "\t" << g.id;
		f << std::endl;
		std::vector<std::vector<char>> has(n, std::vector<char>(model.genes.size(), 0));
		for (size_t k = 0; k < model.genes.size(); k++) {
			if (model.genes[k].all) {
				for (size_t i = 0; i < n; i++) has[i][k] = 1;
			} else {
				for (int32_t i : model.genes[k].carriers) has[i][k] = 1;
			}
		}
		std::string line;
		for (size_t i = 0; i < n; i++) {
			line = pgshost::genome_name((ssize_t)i);
			for (size_t k = 0; k < model.genes.size(); k++) {
				line += '\t';
				line += has[i][k] ? '1' : '0';
			}
			f << line << std::endl;
		}
	}
	const auto t2 = std::chrono::steady_clock::now();

	if (sequences) {
		pgs_config cfg{};
		cfg.seed = model.seed;
		cfg.gene_length = (int64_t)model.gene_length;
		cfg.device = device;
		pgs_engine *h = nullptr;
		if (pgs_create(&cfg, &h) != PGS_OK) errx(2, "%s", pgs_last_error(nullptr));
		auto check = [&](int rc) {
			if (rc < 0) errx(2, "%s", pgs_last_error(h));
		};
		const pgshost::Tree &t = model.tree;
		const size_t N = t.nodes();
		std::vector<double> rate(N, 0.0);
		std::vector<int32_t> leaf_genome(N, -1);
		for (size_t v = 0; v < N; v++) {
			rate[v] = t.time[v] * model.mut_rate; // img.cxx:218,286
			if (v < n) leaf_genome[v] = (int32_t)v;
		}
		check(pgs_set_tree(h, (int32_t)N, t.parent.data(), rate.data(), leaf_genome.data()));

		const size_t G = model.genes.size();
		std::vector<int64_t> gene_id(G), carrier_off(G + 1, 0);
		std::vector<int32_t> origin(G), carriers;
		for (size_t k = 0; k < G; k++) {
			const auto &g = model.genes[k];
			gene_id[k] = g.id;
			origin[k] = g.origin;
			if (g.all && g.origin == t.root()) {
				carriers.push_back(PGS_ALL_GENOMES);
			} else if (g.all) {
				carriers.insert(carriers.end(), model.all_order.begin(), model.all_order.end());
			} else {
				carriers.insert(carriers.end(), g.carriers.begin(), g.carriers.end());
			}
			carrier_off[k + 1] = (int64_t)carriers.size();
		}
		check(pgs_set_genes(h, (int64_t)G, gene_id.data(), origin.data(), carrier_off.data(), carriers.data(),
							model.all_order.data()));
		check(pgs_sweep(h));
		pgs_output_spec spec{};
		const std::string dir = out_dir.substr(0, out_dir.size() - 1);
		spec.out_dir = dir.empty() ? "/" : dir.c_str();
		spec.what = PGS_OUT_ALL;
		spec.n_ref_core = (int64_t)model.ref_core.size();
		spec.ref_core = model.ref_core.data();
		spec.n_ref_acc = (int64_t)model.ref_acc.size();
		spec.ref_acc = model.ref_acc.data();
		pgs_output_report rep{};
		check(pgs_write_outputs(h, &spec, &rep));
		if (verbose) {
			pgs_stats st{};
			pgs_get_stats(h, &st);
			std::fprintf(stderr,
						 "sequences: %lld rows, K0 %.3f ms, K1 %.3f ms (%lld launches), K3 %.3f ms, %lld files, "
						 "%lld bytes, output wall %.1f ms\n",
						 (long long)st.rows_total, st.rootgen_ms, st.sweep_ms, (long long)st.kernel_launches,
						 rep.device_ms, (long long)rep.files, (long long)rep.bytes, rep.wall_ms);
		}
		pgs_destroy(h);
	}
	if (verbose) {
		const auto t3 = std::chrono::steady_clock::now();
		auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
		std::fprintf(stderr, "host: simulate %.1f ms (%s), small files %.1f ms, sequences %.1f ms\n", ms(t0, t1),
					 model.used_compat ? "rng-compat" : "rng-fast", ms(t1, t2), ms(t2, t3));
	}
	return 0;
}
