/* oracle/pgs_replay.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Host replay of the Philox-keyed Jukes-Cantor substitution scheme of DESIGN.md §3, written
 * from the specification; shares no source with the CUDA engine.  It replaces, for checking
 * purposes, the reference's gene::gene (gene.cxx:13-25) and gene::mutate (gene.cxx:27-61) by the
 * north_star's scheme: a counter-based generator keyed by (seed, edge, gene, site-block) -- the
 * decision call of a block is shared by a sibling pair of branches and four adjacent blocks, see
 * oracle_mutate_block -- and a per-branch substitution probability p = 3/4 (1 - exp(-4/3 mu t)),
 * uniform among the 3 other bases.  Pinned by the Random123 Philox4x32-10 known-answer vectors (tests/test_oracle.py).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ Philox4x32-10 */

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for (int round = 0; round < 10; round++) {
		const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
		const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
		const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		const uint32_t n1 = (uint32_t)p1;
		const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		const uint32_t n3 = (uint32_t)p0;
		c0 = n0;
		c1 = n1;
		c2 = n2;
		c3 = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0;
	out[1] = c1;
	out[2] = c2;
	out[3] = c3;
}

/* ------------------------------------------------------------------ per-edge table */

double oracle_jc_p(double rate)
{
	if (!(rate > 0.0)) return 0.0;
	double p = 0.75 * (-expm1(-(4.0 / 3.0) * rate));
	if (p > 0.75) p = 0.75;
	if (p < 0.0) p = 0.0;
	return p;
}

int oracle_edge_table(double rate, uint64_t H[65])
{
	double pmf[65], tail[66];
	const double p = oracle_jc_p(rate);
	H[0] = UINT64_MAX;
	if (p <= 0.0) {
		for (int k = 1; k <= 64; k++) H[k] = 0;
		return 0;
	}
	const double r = p / (1.0 - p);
	pmf[0] = exp(64.0 * log1p(-p));
	for (int j = 0; j < 64; j++) {
		pmf[j + 1] = ((pmf[j] * (double)(64 - j)) / (double)(j + 1)) * r;
	}
	tail[65] = 0.0;
	for (int k = 64; k >= 1; k--) tail[k] = tail[k + 1] + pmf[k];
	int kmax = 0;
	for (int k = 1; k <= 64; k++) {
		const double scaled = tail[k] * 18446744073709551616.0; /* 2^64 */
		if (scaled >= 18446744073709551615.0) { /* rounds to 2^64 in double */
			H[k] = UINT64_MAX;
		} else {
			H[k] = (uint64_t)floor(scaled);
		}
		if (H[k] > 0) kmax = k;
	}
	return kmax;
}

/* ------------------------------------------------------------------ blocks */

static void valid_mask(uint32_t blk, int64_t gene_length, uint32_t v[4])
{
	const int64_t first = (int64_t)blk * ORACLE_BLOCK_SITES;
	for (int w = 0; w < 4; w++) {
		uint32_t m = 0;
		for (int s = 0; s < 16; s++) {
			if (first + 16 * w + s < gene_length) m |= 3u << (2 * s);
		}
		v[w] = m;
	}
}

void oracle_root_block(uint64_t seed, uint32_t gene_id, uint32_t blk, int64_t gene_length,
					   uint32_t out[4])
{
	const uint32_t ctr[4] = {ORACLE_ROOT_EDGE, gene_id, blk, 0u};
	const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
	uint32_t v[4];
	oracle_philox4x32_10(ctr, key, out);
	valid_mask(blk, gene_length, v);
	for (int w = 0; w < 4; w++) out[w] &= v[w];
}

/* Sibling pairs: the children of a node in ascending index are paired (c0,c1), (c2,c3), ...;
 * leader[v] = first node of v's pair, side[v] = 0 for the leader, 1 for the other.  Root: itself. */
void oracle_sibling_pairs(int32_t n_nodes, const int32_t *parent, int32_t *leader, int32_t *side)
{
	/* rank of v among its siblings = number of smaller-index nodes with the same parent */
	int32_t *seen = (int32_t *)calloc((size_t)n_nodes, sizeof(int32_t));   /* children met so far */
	int32_t *last = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes); /* previous child met */
	for (int32_t v = 0; v < n_nodes; v++) {
		const int32_t p = parent[v];
		if (p < 0) {
			leader[v] = v;
			side[v] = 0;
			continue;
		}
		if (seen[p] % 2 == 0) {
			leader[v] = v;
			side[v] = 0;
		} else {
			leader[v] = last[p];
			side[v] = 1;
		}
		last[p] = v;
		seen[p]++;
	}
	free(seen);
	free(last);
}

/* word stream of one (edge, gene, block): words 0..3 of stream 1, then of stream 2, ... */
typedef struct {
	uint32_t ctr[4];
	uint32_t key[2];
	uint32_t buf[4];
	int next; /* index into buf of the next unread word */
} word_stream;

static uint32_t stream_next(word_stream *s)
{
	if (s->next == 4) {
		s->ctr[3] += 1u;
		oracle_philox4x32_10(s->ctr, s->key, s->buf);
		s->next = 0;
	}
	return s->buf[s->next++];
}

void oracle_mutate_block(uint64_t seed, uint32_t leader, uint32_t side, uint32_t edge, uint32_t gene_id,
						 uint32_t blk, int64_t gene_length, const uint64_t H[65], uint32_t x[4])
{
	/* decision field: one call is shared by the sibling pair and by the quad of blocks 4q..4q+3;
	 * word 2*side + (pair of the quad), 16-bit field (even block: low half, odd block: high half) */
	const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
	const uint32_t pctr[4] = {leader, gene_id, blk / 4u, 0u};
	uint32_t pw[4];
	oracle_philox4x32_10(pctr, key, pw);
	const uint32_t pair_word = pw[2u * side + ((blk / 2u) % 2u)];
	const uint32_t D = (blk % 2u) ? (pair_word >> 16) : (pair_word & 0xFFFFu);
	if (D > (uint32_t)(H[1] >> 48)) return; /* U >= H[1] whatever the low 48 bits: no substitution */

	word_stream s;
	s.ctr[0] = edge;
	s.ctr[1] = gene_id;
	s.ctr[2] = blk;
	s.ctr[3] = 0u; /* stream_next pre-increments: the first call is stream 1 */
	s.key[0] = key[0];
	s.key[1] = key[1];
	s.next = 4;
	/* U = D : first word : top half of the second word (the rest of that word is not used) */
	const uint64_t u1 = stream_next(&s), u2 = stream_next(&s);
	const uint64_t U = ((uint64_t)D << 48) | (u1 << 16) | (u2 >> 16);

	int K = 0;
	while (K < 64 && U < H[K + 1]) K++;
	if (K == 0) return;

	uint64_t taken = 0;
	uint32_t mask[4] = {0, 0, 0, 0};
	for (int j = 0; j < K; j++) {
		const uint32_t w = stream_next(&s);
		const uint64_t prod = (uint64_t)w * (uint32_t)(64 - j);
		uint32_t r = (uint32_t)(prod >> 32);           /* uniform in [0, 64-j) */
		const uint32_t lo = (uint32_t)prod;
		const uint32_t delta = 1u + (uint32_t)(((uint64_t)lo * 3u) >> 32); /* 1..3 */
		int pos = 0;
		for (;; pos++) { /* r-th site (0-based, ascending) not yet taken */
			if (!((taken >> pos) & 1u)) {
				if (r == 0) break;
				r--;
			}
		}
		taken |= (uint64_t)1 << pos;
		mask[pos >> 4] |= delta << (2 * (pos & 15));
	}
	uint32_t v[4];
	valid_mask(blk, gene_length, v);
	for (int w = 0; w < 4; w++) x[w] ^= mask[w] & v[w];
}

/* ------------------------------------------------------------------ rows */

static int64_t blocks_of(int64_t gene_length)
{
	return (gene_length + ORACLE_BLOCK_SITES - 1) / ORACLE_BLOCK_SITES;
}

int oracle_row(uint64_t seed, int32_t n_nodes, const int32_t *parent, const double *rate,
			   int32_t node, int32_t origin, uint32_t gene_id, int64_t gene_length, uint32_t *dst)
{
	if (node < 0 || node >= n_nodes || origin < 0 || origin >= n_nodes) return -1;
	/* path node -> origin */
	int32_t *path = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
	int len = 0;
	int32_t v = node;
	while (v != origin) {
		if (v < 0 || len >= n_nodes) {
			free(path);
			return -1;
		}
		path[len++] = v;
		v = parent[v];
	}
	const int64_t nb = blocks_of(gene_length);
	for (int64_t b = 0; b < nb; b++) {
		oracle_root_block(seed, gene_id, (uint32_t)b, gene_length, dst + 4 * b);
	}
	int32_t *leader = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
	int32_t *side = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
	oracle_sibling_pairs(n_nodes, parent, leader, side);
	uint64_t H[65];
	for (int i = len - 1; i >= 0; i--) {
		const int32_t e = path[i];
		oracle_edge_table(rate[e], H);
		for (int64_t b = 0; b < nb; b++) {
			oracle_mutate_block(seed, (uint32_t)leader[e], (uint32_t)side[e], (uint32_t)e, gene_id, (uint32_t)b,
								gene_length, H, dst + 4 * b);
		}
	}
	free(leader);
	free(side);
	free(path);
	return 0;
}

int oracle_gene_all_nodes(uint64_t seed, int32_t n_nodes, const int32_t *parent, const double *rate,
						  int32_t origin, uint32_t gene_id, int64_t gene_length, uint32_t *dst)
{
	if (origin < 0 || origin >= n_nodes) return -1;
	const int64_t nb = blocks_of(gene_length);
	const size_t words = (size_t)(4 * nb);
	memset(dst, 0, sizeof(uint32_t) * words * (size_t)n_nodes);
	/* depth below origin, or -1 when not in the subtree; done[] marks computed rows */
	int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
	int32_t *depth = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
	int32_t maxd = 0;
	for (int32_t v = 0; v < n_nodes; v++) {
		int32_t d = 0, u = v;
		while (u != origin && u >= 0 && d <= n_nodes) {
			u = parent[u];
			d++;
		}
		depth[v] = (u == origin) ? d : -1;
		if (depth[v] > maxd) maxd = depth[v];
	}
	for (int64_t b = 0; b < nb; b++) {
		oracle_root_block(seed, gene_id, (uint32_t)b, gene_length, dst + words * (size_t)origin + 4 * b);
	}
	int32_t *leader = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
	int32_t *side = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
	oracle_sibling_pairs(n_nodes, parent, leader, side);
	uint64_t H[65];
	for (int32_t d = 1; d <= maxd; d++) {
		for (int32_t v = 0; v < n_nodes; v++) {
			if (depth[v] != d) continue;
			uint32_t *row = dst + words * (size_t)v;
			memcpy(row, dst + words * (size_t)parent[v], sizeof(uint32_t) * words);
			oracle_edge_table(rate[v], H);
			for (int64_t b = 0; b < nb; b++) {
				oracle_mutate_block(seed, (uint32_t)leader[v], (uint32_t)side[v], (uint32_t)v, gene_id,
									(uint32_t)b, gene_length, H, row + 4 * b);
			}
		}
	}
	free(leader);
	free(side);
	free(order);
	free(depth);
	return 0;
}

uint64_t oracle_mismatch(const uint32_t *a, const uint32_t *b, int64_t blocks)
{
	uint64_t total = 0;
	for (int64_t w = 0; w < 4 * blocks; w++) {
		const uint32_t d = a[w] ^ b[w];
		for (int s = 0; s < 16; s++) total += ((d >> (2 * s)) & 3u) != 0;
	}
	return total;
}

void oracle_unpack_ascii(const uint32_t *row, int64_t gene_length, char *dst)
{
	static const char acgt[4] = {'A', 'C', 'G', 'T'};
	for (int64_t s = 0; s < gene_length; s++) {
		dst[s] = acgt[(row[s >> 4] >> (2 * (s & 15))) & 3u];
	}
}
