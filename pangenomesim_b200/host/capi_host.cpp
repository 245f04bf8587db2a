// capi_host.cpp -- C entry points of the host model for Python callers (bench.py, tests): the
// coalescent exactly as the host program draws it (reference create_coalescent, img.cxx:71-116,
// same libstdc++ engine and distributions), so benchmarks run on the reference's own trees.
#include "model.h"

#include <cstring>

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

} // extern "C"
