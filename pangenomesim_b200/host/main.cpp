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
#include <charconv>
#include <cmath>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <future>
#include <iomanip>
#include <iostream>
#include <regex>
#include <sstream>
#include <thread>
#include <unordered_map>

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
		"      --device N           CUDA device ordinal\n"
		"      --devices N          Shard the genes over N engines (devices 0..N-1, wrapping around)\n"
		"      --packed             Also write genomes.pgs2: all genomes 2-bit packed behind an index\n"
		"                           (one device only; pangenomesim_b200/packed.py reads it)\n"
		"      --gpu-format         Format the text on the GPU (kernel K3) instead of shipping packed rows\n"
		"                           and expanding them on the host's cores; same bytes either way\n"
		"      --check-distances    Count the observed pairwise differences over the core genes on the GPU(s)\n"
		"                           (partial counts summed with NCCL), write distance.observed.mat (Jukes-Cantor\n"
		"                           corrected) and check every pair against distance.mat's expectation (99 %% band)\n\n"
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

// A fatal error after threads have been started: thrown to main, which joins the threads before it
// exits (exit() while another thread is still writing is undefined behaviour and leaves partial files).
struct Fatal {
	int code;
	std::string msg;
};
[[noreturn]] static void fatal(int code, const std::string &msg)
{
	throw Fatal{code, msg};
}

static std::ofstream open_out(const std::string &name)
{
	std::ofstream f(name);
	if (!f.good()) fatal(errno ? errno : 1, name + ": couldn't write to file: " + std::strerror(errno)); // reference: check_io, util.cxx:66-70
	return f;
}

static void check_out(std::ostream &f, const std::string &name)
{
	f.flush();
	if (!f.good()) fatal(errno ? errno : 1, name + ": couldn't write to file: " + std::strerror(errno));
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
	std::future<void> writing;
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
				for (double d : row) { // "  %8.4g": std::to_chars(general, 4) is printf's %.4g, 4x faster
					const std::to_chars_result r = std::to_chars(buf, buf + sizeof buf, d, std::chars_format::general, 4);
					const size_t len = (size_t)(r.ptr - buf);
					line.append(len < 8 ? 10 - len : 2, ' ');
					line.append(buf, len);
				}
				line += "\n";
				lines[i - r0].swap(line);
			}
		};
		std::vector<std::thread> pool;
		for (unsigned w = 1; w < workers; w++) pool.emplace_back(work);
		work();
		for (auto &th : pool) th.join();
		// the batch is written while the next one is computed
		if (writing.valid()) writing.get();
		writing = std::async(std::launch::async, [&out, done = std::move(lines)] {
			for (const auto &l : done) out << l;
		});
	}
	if (writing.valid()) writing.get();
	check_out(out, name);
}

static int run(int argc, char *argv[], std::thread &matrices, std::string &matrices_error)
{
	static const struct option long_options[] = {
		{"version", no_argument, nullptr, 0},      {"help", no_argument, nullptr, 0},
		{"param", required_argument, nullptr, 'p'}, {"out-dir", required_argument, nullptr, 'o'},
		{"verbose", no_argument, nullptr, 'v'},     {"rng-compat", no_argument, nullptr, 0},
		{"rng-fast", no_argument, nullptr, 0},      {"no-sequences", no_argument, nullptr, 0},
		{"device", required_argument, nullptr, 0},  {"devices", required_argument, nullptr, 0},
		{"packed", no_argument, nullptr, 0},        {"gpu-format", no_argument, nullptr, 0},
		{"check-distances", no_argument, nullptr, 0},
		{nullptr, 0, nullptr, 0}};

	Model model;
	std::string out_dir = "./";
	bool sequences = true, packed = false, gpu_format = false, check_distances = false;
	int device = -1, n_devices = 1;

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
				if (name == "devices") n_devices = std::atoi(optarg);
				if (name == "packed") packed = true;
				if (name == "gpu-format") gpu_format = true;
				if (name == "check-distances") check_distances = true;
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
					try {
						if (!model.parse_param(key, value)) errx(1, "unkown key %s.", key.c_str());
					} catch (const std::exception &) { // the reference dies of an uncaught std::invalid_argument here
						errx(1, "invalid value '%s' for parameter %s", value.c_str(), key.c_str());
					}
					rest = m.suffix();
				}
				if (!rest.empty()) errx(1, "invalid parameter %s", optarg);
				break;
			}
			case 'v': verbose = true; break;
			default: usage(EXIT_FAILURE);
		}
	}

	if (model.num_genomes < 1) errx(1, "num-genomes must be at least 1");
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
	// distance.mat (n^2 entries: 1 GB at 10 000 genomes) and panmatrix.tsv take seconds of host time
	// and depend on nothing the GPU produces: they are written while the sequences are made
	auto write_matrices = [&]() {
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
	};
	if (sequences)
		matrices = std::thread([&] {
			try {
				write_matrices();
			} catch (const Fatal &f) {
				matrices_error = f.msg;
			}
		});
	else
		write_matrices();
	const auto t2 = std::chrono::steady_clock::now();

	if (sequences) {
		const pgshost::Tree &t = model.tree;
		const size_t N = t.nodes(), G = model.genes.size();
		std::vector<double> rate(N, 0.0);
		std::vector<int32_t> leaf_genome(N, -1);
		for (size_t v = 0; v < N; v++) {
			rate[v] = t.time[v] * model.mut_rate; // img.cxx:218,286
			if (v < n) leaf_genome[v] = (int32_t)v;
		}
		const int available = pgs_device_count();
		if (available < 1) fatal(2, "no usable CUDA device; this program has no CPU path for sequences");
		if (n_devices < 1) n_devices = 1;
		if (packed && n_devices > 1) fatal(1, "--packed needs a single engine (no --devices)");

		// contiguous slices of the gene table, balanced by the rows they occupy
		std::vector<size_t> cut(n_devices + 1, G);
		{
			std::vector<double> weight(G + 1, 0.0);
			for (size_t k = 0; k < G; k++)
				weight[k + 1] = weight[k] + (model.genes[k].all ? (double)N : (double)model.genes[k].carriers.size() * 2.0 + 1.0);
			cut[0] = 0;
			size_t k = 0;
			for (int d = 1; d < n_devices; d++) {
				while (k < G && weight[k] < weight[G] * d / n_devices) k++;
				cut[d] = k;
			}
		}
		struct Shard {
			pgs_engine *h = nullptr;
			std::string error;
			pgs_output_report rep{};
			pgs_stats st{};
			std::vector<int64_t> genome_bytes, genome_offset, ref_core, ref_acc;
			int64_t maf_bytes = 0, maf_offset = 0;
		};
		std::vector<Shard> shards(n_devices);
		auto each = [&](auto fn) { // run fn(d) for every shard on its own thread, stop on the first error
			std::vector<std::thread> pool;
			for (int d = 0; d < n_devices; d++) pool.emplace_back([&, d] { fn(d); });
			for (auto &th : pool) th.join();
			for (int d = 0; d < n_devices; d++)
				if (!shards[d].error.empty()) fatal(2, "device " + std::to_string(d) + ": " + shards[d].error);
		};
		std::unordered_map<int64_t, std::pair<int, int64_t>> where; // gene table index -> (shard, local index)
		each([&](int d) {
			Shard &s = shards[d];
			auto check = [&](int rc) {
				if (rc < 0 && s.error.empty()) s.error = pgs_last_error(s.h);
				return rc >= 0;
			};
			pgs_config cfg{};
			cfg.seed = model.seed;
			cfg.gene_length = (int64_t)model.gene_length;
			cfg.device = device >= 0 && n_devices == 1 ? device : d % available;
			if (pgs_create(&cfg, &s.h) != PGS_OK) {
				s.error = pgs_last_error(nullptr);
				return;
			}
			if (!check(pgs_set_tree(s.h, (int32_t)N, t.parent.data(), rate.data(), leaf_genome.data()))) return;
			const size_t lo = cut[d], hi = cut[d + 1];
			std::vector<int64_t> gene_id(hi - lo), carrier_off(hi - lo + 1, 0);
			std::vector<int32_t> origin(hi - lo), carriers;
			for (size_t k = lo; k < hi; k++) {
				const auto &g = model.genes[k];
				gene_id[k - lo] = g.id;
				origin[k - lo] = g.origin;
				if (g.all && g.origin == t.root())
					carriers.push_back(PGS_ALL_GENOMES);
				else if (g.all)
					carriers.insert(carriers.end(), model.all_order.begin(), model.all_order.end());
				else
					carriers.insert(carriers.end(), g.carriers.begin(), g.carriers.end());
				carrier_off[k - lo + 1] = (int64_t)carriers.size();
			}
			if (!check(pgs_set_genes(s.h, (int64_t)(hi - lo), gene_id.data(), origin.data(), carrier_off.data(),
									 carriers.data(), model.all_order.data())))
				return;
			check(pgs_sweep(s.h));
		});
		for (int d = 0; d < n_devices; d++)
			for (size_t k = cut[d]; k < cut[d + 1]; k++) where[(int64_t)k] = {d, (int64_t)(k - cut[d])};

		const std::string dir = out_dir.substr(0, out_dir.size() - 1);
		const std::string dir_c = dir.empty() ? "/" : dir;
		if (n_devices == 1) {
			Shard &s = shards[0];
			pgs_output_spec spec{};
			spec.out_dir = dir_c.c_str();
			spec.what = PGS_OUT_ALL | (packed ? PGS_OUT_PACKED : 0u) | (gpu_format ? 0u : PGS_OUT_HOST_EXPAND);
			spec.n_ref_core = (int64_t)model.ref_core.size();
			spec.ref_core = model.ref_core.data();
			spec.n_ref_acc = (int64_t)model.ref_acc.size();
			spec.ref_acc = model.ref_acc.data();
			if (pgs_write_outputs(s.h, &spec, &s.rep) < 0) fatal(2, pgs_last_error(s.h));
		} else {
			// several engines fill disjoint byte ranges of the same files
			const int64_t n_ref_total = (int64_t)(model.ref_core.size() + model.ref_acc.size());
			each([&](int d) {
				Shard &s = shards[d];
				pgs_shard_layout lay{};
				lay.genome_gene_total = model.genome_gene_count.data();
				lay.n_ref_total = n_ref_total;
				lay.maf_offset = d; // only "is this the first shard" matters for the size
				pgs_output_spec spec{};
				spec.what = PGS_OUT_GENOMES | PGS_OUT_MAF;
				spec.shard = &lay;
				s.genome_bytes.assign(n, 0);
				if (pgs_output_sizes(s.h, &spec, s.genome_bytes.data(), &s.maf_bytes) < 0) s.error = pgs_last_error(s.h);
			});
			std::vector<int64_t> at(n, 0);
			int64_t maf_at = 0;
			for (int d = 0; d < n_devices; d++) {
				shards[d].genome_offset = at;
				shards[d].maf_offset = maf_at;
				for (size_t i = 0; i < n; i++) at[i] += shards[d].genome_bytes[i];
				maf_at += shards[d].maf_bytes;
			}
			// the files are created at their final size: the engines fill disjoint ranges and never resize them
			auto create_sized = [&](const std::string &path, int64_t size) {
				open_out(path);
				if (::truncate(path.c_str(), (off_t)size) != 0) fatal(errno, path + ": couldn't set the file size: " + std::strerror(errno));
			};
			for (size_t i = 0; i < n; i++) create_sized(out_dir + pgshost::genome_name((ssize_t)i) + ".fasta", at[i]);
			create_sized(out_dir + "alignment.maf", maf_at);
			each([&](int d) {
				Shard &s = shards[d];
				pgs_shard_layout lay{};
				lay.genome_offset = s.genome_offset.data();
				lay.maf_offset = s.maf_offset;
				lay.genome_gene_total = model.genome_gene_count.data();
				lay.n_ref_total = n_ref_total;
				pgs_output_spec spec{};
				spec.out_dir = dir_c.c_str();
				spec.what = PGS_OUT_GENOMES | PGS_OUT_MAF | (gpu_format ? 0u : PGS_OUT_HOST_EXPAND);
				spec.shard = &lay;
				if (pgs_write_outputs(s.h, &spec, &s.rep) < 0) s.error = pgs_last_error(s.h);
			});
			// ref.fasta = core.fasta + accessory.fasta, gene::to_fasta layout (gene.h:42-52)
			std::string seq(model.gene_length, 'N');
			auto record = [&](int64_t k) {
				const auto w = where.at(k);
				if (pgs_copy_origin_ascii(shards[w.first].h, w.second, &seq[0]) < 0) fatal(2, pgs_last_error(shards[w.first].h));
				return ">genomeref." + std::to_string(model.genes[k].id) + "\n" + seq + "\n";
			};
			std::ofstream ref = open_out(out_dir + "ref.fasta"), cor = open_out(out_dir + "core.fasta"),
						  acc = open_out(out_dir + "accessory.fasta");
			for (int64_t k : model.ref_core) {
				const std::string r = record(k);
				ref << r;
				cor << r;
			}
			for (int64_t k : model.ref_acc) {
				const std::string r = record(k);
				ref << r;
				acc << r;
			}
			check_out(ref, out_dir + "ref.fasta"); // the reference checks every file it has written (check_io)
			check_out(cor, out_dir + "core.fasta");
			check_out(acc, out_dir + "accessory.fasta");
		}
		if (check_distances) {
			// Observed pairwise differences over the core genes (kernel K2 on every engine's genes, the
			// partial counts summed with NCCL over NVLink) against what distance.mat promises: a pair at
			// expected distance d differs in p = 3/4 (1 - exp(-4d/3)) of its sites (Jukes-Cantor), and the
			// observed fraction must lie within the 99 % band of a Binomial(S, p), Bonferroni over all pairs.
			std::vector<uint32_t> mm(n * n);
			int64_t sites = 0;
			std::vector<pgs_engine *> hs;
			for (auto &sh : shards) hs.push_back(sh.h);
			if (n_devices == 1) {
				if (pgs_mismatch(hs[0], mm.data(), &sites) < 0) fatal(2, pgs_last_error(hs[0]));
			} else if (n_devices <= available) {
				if (pgs_mismatch_multi(hs.data(), n_devices, mm.data(), &sites) < 0) fatal(2, pgs_last_error(hs[0]));
			} else { // engines share a device: no NCCL rank per device, the partial matrices are summed here
				std::vector<uint32_t> part(n * n);
				for (auto h : hs) {
					int64_t ps = 0;
					if (pgs_mismatch(h, part.data(), &ps) < 0) fatal(2, pgs_last_error(h));
					for (size_t i = 0; i < n * n; i++) mm[i] += part[i];
					sites += ps;
				}
			}
			const double pairs = (double)n * (double)(n - 1) / 2.0;
			double zc = 0.0; // two-sided critical value at 0.01 / pairs
			if (pairs >= 1) {
				double lo = 0.0, hi = 40.0;
				for (int it = 0; it < 200; it++) {
					zc = 0.5 * (lo + hi);
					(std::erfc(zc / std::sqrt(2.0)) > 0.01 / pairs ? lo : hi) = zc;
				}
			}
			std::ofstream out = open_out(out_dir + "distance.observed.mat");
			out << n << "\n";
			std::vector<double> row;
			size_t outside = 0;
			double worst = 0.0;
			std::string line;
			char buf[64];
			for (size_t i = 0; i < n; i++) {
				model.distance_row(i, row);
				line = pgshost::genome_name((ssize_t)i);
				for (size_t j = 0; j < n; j++) {
					const double ph = sites > 0 ? (double)mm[i * n + j] / (double)sites : 0.0;
					const double dh = ph < 0.75 ? -0.75 * std::log1p(-4.0 * ph / 3.0) : INFINITY; // Jukes-Cantor correction
					const std::to_chars_result r = std::to_chars(buf, buf + sizeof buf, dh, std::chars_format::general, 4);
					const size_t len = (size_t)(r.ptr - buf);
					line.append(len < 8 ? 10 - len : 2, ' ');
					line.append(buf, len);
					if (j > i && sites > 0) {
						const double pe = 0.75 * (1.0 - std::exp(-4.0 * row[j] / 3.0));
						const double sd = std::sqrt(std::max(pe * (1.0 - pe), 0.0) / (double)sites);
						const double z = sd > 0 ? std::fabs(ph - pe) / sd : (ph == pe ? 0.0 : INFINITY);
						worst = std::max(worst, z);
						if (std::fabs(ph - pe) > zc * sd + 1e-12) outside++;
					}
				}
				line += "\n";
				out << line;
			}
			check_out(out, out_dir + "distance.observed.mat");
			std::fprintf(stderr, "check-distances: %lld sites, %.0f pairs, largest |z| %.2f (99 %% band: %.2f), %zu pairs outside\n",
						 (long long)sites, pairs, worst, zc, outside);
			if (outside > 0) fatal(3, "observed distances disagree with distance.mat");
		}
		for (int d = 0; d < n_devices; d++) {
			Shard &s = shards[d];
			if (verbose) {
				pgs_get_stats(s.h, &s.st);
				std::fprintf(stderr,
							 "device %d: genes [%zu,%zu), %lld rows, K0 %.3f ms, K1 %.3f ms (%lld launches), K3 %.3f ms, "
							 "%lld files, %lld bytes, output wall %.1f ms\n",
							 d, cut[d], cut[d + 1], (long long)s.st.rows_total, s.st.rootgen_ms, s.st.sweep_ms,
							 (long long)s.st.kernel_launches, s.rep.device_ms, (long long)s.rep.files, (long long)s.rep.bytes,
							 s.rep.wall_ms);
			}
			pgs_destroy(s.h);
		}
	}
	if (matrices.joinable()) matrices.join();
	if (!matrices_error.empty()) fatal(2, matrices_error);
	if (verbose) {
		const auto t3 = std::chrono::steady_clock::now();
		auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
		std::fprintf(stderr, "host: simulate %.1f ms (%s), small files %.1f ms, sequences (and the two matrices beside them) %.1f ms\n", ms(t0, t1),
					 model.used_compat ? "rng-compat" : "rng-fast", ms(t1, t2), ms(t2, t3));
	}
	return 0;
}

int main(int argc, char *argv[])
{
	std::thread matrices; // distance.mat and panmatrix.tsv are written beside the GPU work
	std::string matrices_error;
	int rc = 0;
	Fatal failure{0, ""};
	try {
		rc = run(argc, argv, matrices, matrices_error);
	} catch (const Fatal &f) {
		failure = f;
		if (failure.code == 0) failure.code = 1;
	}
	if (matrices.joinable()) matrices.join(); // never exit under a running thread
	if (failure.code) errx(failure.code, "%s", failure.msg.c_str());
	return rc;
}
