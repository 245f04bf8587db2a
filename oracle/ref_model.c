/* oracle/ref_model.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference's own algorithm for the sequence path, so that the
 * reference's behaviour (including its quirks Q1-Q3 of SURVEY.md §8a) can be replayed, timed
 * ("port" CPU baseline) and modelled without its C++ object graph.  Pinned bit-exactly against
 * core.fasta / genomeNNN.fasta written by the unmodified reference binary (tests/golden/,
 * tests/test_ref_model.py).
 *
 * Third-party arithmetic restated here (absent from /root/reference): GNU libstdc++ 13
 * <random> -- std::default_random_engine = minstd_rand0 (bits/random.h), the "two divisions +
 * rejection" branch of uniform_int_distribution (bits/uniform_int_dist.h) and
 * generate_canonical<double,53> (bits/random.tcc).  Reference call sites: gene.cxx:19-20,36-40.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define REF_M 2147483647u /* 2^31 - 1 */
#define REF_A 16807u

void ref_engine_seed(ref_engine *e, uint64_t seed)
{
	/* linear_congruential_engine::seed: c == 0 and s mod m == 0  ->  state 1 */
	const uint32_t s = (uint32_t)(seed % REF_M);
	e->x = s == 0 ? 1u : s;
}

uint32_t ref_engine_next(ref_engine *e)
{
	e->x = (uint32_t)(((uint64_t)REF_A * e->x) % REF_M);
	return e->x; /* in [1, 2^31 - 2] */
}

uint32_t ref_uniform_int(ref_engine *e, uint32_t hi)
{
	/* urng range 2^31-3 (max - min), target range hi (>= 0, smaller): downscale with rejection */
	const uint64_t urngrange = 2147483645u;
	const uint64_t uerange = (uint64_t)hi + 1u;
	const uint64_t scaling = urngrange / uerange;
	const uint64_t past = uerange * scaling;
	uint64_t ret;
	do {
		ret = (uint64_t)ref_engine_next(e) - 1u;
	} while (ret >= past);
	return (uint32_t)(ret / scaling);
}

double ref_uniform_real(ref_engine *e)
{
	/* generate_canonical<double,53>: range r = 2^31 - 2, log2 r -> 30, m = (53+30-1)/30 = 2 calls */
	const long double r = 2147483646.0L;
	double sum = 0.0, tmp = 1.0;
	for (int k = 0; k < 2; k++) {
		sum += (double)(ref_engine_next(e) - 1u) * tmp;
		tmp = (double)((long double)tmp * r);
	}
	double ret = sum / tmp;
	if (ret >= 1.0) ret = nextafter(1.0, 0.0);
	return ret;
}

/* std::exponential_distribution<double>(lambda)(engine): -log(1 - canonical) / lambda (random.h) */
double ref_exponential(ref_engine *e, double lambda)
{
	return -log(1.0 - ref_uniform_real(e)) / lambda;
}

/* create_coalescent(n) -- reference img.cxx:71-116.  Pool layout: slots [0,n) leaves, slot n+i
 * the i-th merger, root last.  Topology: two rand_int picks per merger over a shrinking
 * indirection table (img.cxx:87-96); times: cumulative Exp(k(k-1)/2) waits (img.cxx:103-107);
 * branch time = parent abs time - own abs time (img.cxx:111-113, img.h:165-171).
 * All output arrays have 2n-1 entries. */
void ref_coalescent(ref_engine *global, int64_t n, int32_t *parent, int32_t *left, int32_t *right,
					double *time)
{
	const int64_t m = 2 * n - 1;
	int32_t *indirect = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
	double *abs_time = (double *)calloc((size_t)m, sizeof(double));
	for (int64_t i = 0; i < m; i++) {
		indirect[i] = (int32_t)i;
		parent[i] = left[i] = right[i] = -1;
		time[i] = 0.0;
	}
	for (int64_t i = 0; i + 1 < n; i++) {
		const int32_t p = (int32_t)(n + i);
		const uint32_t c = ref_uniform_int(global, (uint32_t)(n - i - 1)); /* rand_int(0, n-i) */
		left[indirect[p]] = indirect[c];
		parent[indirect[c]] = indirect[p];
		indirect[c] = indirect[n - i - 1];
		const uint32_t d = ref_uniform_int(global, (uint32_t)(n - i - 2)); /* rand_int(0, n-i-1) */
		right[indirect[p]] = indirect[d];
		parent[indirect[d]] = indirect[p];
		indirect[d] = indirect[p];
	}
	double t = 0.0;
	uint64_t sample = (uint64_t)n;
	for (int64_t i = 0; i + 1 < n; i++, sample--) {
		const uint64_t pairs = sample * (sample - 1) / 2;
		t += ref_exponential(global, (double)pairs);
		abs_time[n + i] = t;
	}
	for (int64_t i = 0; i + 1 < m; i++) time[i] = abs_time[parent[i]] - abs_time[i];
	free(indirect);
	free(abs_time);
}

/* reference gene.cxx:13-25 -- one uniform_int(0,3) draw per site on the global engine */
void ref_gene_root(ref_engine *global, int64_t len, char *dst)
{
	static const char acgt[4] = {'A', 'C', 'G', 'T'};
	for (int64_t i = 0; i < len; i++) dst[i] = acgt[ref_uniform_int(global, 3)];
}

static char other_base(char c, uint32_t idx)
{
	/* reference gene.cxx:31-34,41-47: the three bases that are not c, alphabetical */
	switch (c) {
		case 'A': return "CGT"[idx];
		case 'C': return "AGT"[idx];
		case 'G': return "ACT"[idx];
		case 'T': return "ACG"[idx];
		default: return 'X';
	}
}

/* reference gene.cxx:27-61.  Site selection draws from a COPY of the global engine taken on
 * entry (gene.cxx:37, std::bind copies); only the replacement-base draw (gene.cxx:40) advances
 * the global engine. */
int64_t ref_gene_mutate(ref_engine *global, const char *src, int64_t len, double rate, char *dst)
{
	ref_engine site = *global;
	double nucleotides = (double)len;
	double mutations = rate * nucleotides;
	int64_t hits = 0;
	if (dst != src) memcpy(dst, src, (size_t)len);
	for (int64_t i = 0; i < len; i++) {
		if (ref_uniform_real(&site) < mutations / nucleotides) {
			dst[i] = other_base(dst[i], ref_uniform_int(global, 2));
			mutations--;
			hits++;
		}
		nucleotides--;
	}
	return hits;
}

int64_t ref_gene_mutate_count(ref_engine *global, int64_t len, double rate)
{
	ref_engine site = *global;
	double nucleotides = (double)len;
	double mutations = rate * nucleotides;
	int64_t hits = 0;
	for (int64_t i = 0; i < len; i++) {
		if (ref_uniform_real(&site) < mutations / nucleotides) {
			(void)ref_uniform_int(global, 2);
			mutations--;
			hits++;
		}
		nucleotides--;
	}
	return hits;
}

/* reference img.cxx:203-232 over tree_node::traverse(pre, process, post) (img.h:316-329):
 * pre = mutate from the sequence on top of the stack (non-root), process = store leaf,
 * post = pop.  Written iteratively: frame state 0 = entering, 1 = left done, 2 = right done. */
void ref_sim_core_gene(ref_engine *global, int32_t root, const int32_t *left, const int32_t *right,
					   const double *time, const int32_t *leaf_index, double mut_rate, int64_t len,
					   char *root_out, char *leaves_out)
{
	int32_t cap = 64, top = 0;
	int32_t *node = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
	int32_t *state = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
	char *seqs = (char *)malloc((size_t)cap * (size_t)len);

	ref_gene_root(global, len, seqs);
	memcpy(root_out, seqs, (size_t)len);
	node[0] = root;
	state[0] = 0;
	while (top >= 0) {
		const int32_t v = node[top];
		int32_t child = -1;
		if (state[top] == 0) {
			state[top] = 1;
			child = left[v];
			if (child < 0) continue;
		} else if (state[top] == 1) {
			if (left[v] < 0 && leaf_index[v] >= 0) {
				memcpy(leaves_out + (size_t)leaf_index[v] * (size_t)len, seqs + (size_t)top * (size_t)len,
					   (size_t)len);
			}
			state[top] = 2;
			child = right[v];
			if (child < 0) continue;
		} else {
			top--;
			continue;
		}
		if (top + 1 == cap) {
			cap *= 2;
			node = (int32_t *)realloc(node, sizeof(int32_t) * (size_t)cap);
			state = (int32_t *)realloc(state, sizeof(int32_t) * (size_t)cap);
			seqs = (char *)realloc(seqs, (size_t)cap * (size_t)len);
		}
		ref_gene_mutate(global, seqs + (size_t)top * (size_t)len, len, time[child] * mut_rate,
						seqs + (size_t)(top + 1) * (size_t)len);
		top++;
		node[top] = child;
		state[top] = 0;
	}
	free(node);
	free(state);
	free(seqs);
}

/* Exact expectation of the number of substituted sites of one gene::mutate(rate) call.
 * With quota q = rate*len and h hits so far, site i (0-based) is hit with probability
 * clamp((q - h) / (len - i), 0, 1); dist[h] carries P(h hits after i sites). */
double ref_expected_hits(double rate, int64_t len)
{
	const double quota = rate * (double)len;
	if (!(quota > 0.0)) return 0.0;
	if (quota >= (double)len) return (double)len;
	if (quota > 64.0) {
		/* the programme is O(len * quota); for large quotas the count is floor or ceil of the quota
		 * (+0.3..+0.6 hits), so return the midpoint: relative error < 8e-3 of one branch, absolute < 0.5 hit */
		return quota == floor(quota) ? quota : floor(quota) + 0.5;
	}
	int64_t hmax = (int64_t)ceil(quota) + 1;
	if (hmax > len) hmax = len;
	double *dist = (double *)calloc((size_t)hmax + 2, sizeof(double));
	dist[0] = 1.0;
	for (int64_t i = 0; i < len; i++) {
		const double left = (double)(len - i);
		for (int64_t h = (hmax < i ? hmax : i); h >= 0; h--) {
			if (dist[h] == 0.0) continue;
			double ph = (quota - (double)h) / left;
			if (ph <= 0.0) continue;
			if (ph > 1.0) ph = 1.0;
			const double moved = dist[h] * ph;
			dist[h] -= moved;
			dist[h + 1] += moved;
		}
	}
	double e = 0.0;
	for (int64_t h = 0; h <= hmax + 1; h++) e += (double)h * dist[h];
	free(dist);
	return e;
}
