/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU checkers for the pangenomesim sequence path.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (pangenomesim_b200/lib/libpgs.so) never links or calls it and
 * fails loudly without its CUDA kernels.
 *
 * Two independent pieces:
 *
 *  (1) pgs_replay.c -- host replay of the Philox-keyed substitution scheme that the GPU engine
 *      implements (specified in DESIGN.md §3).  Written from the specification, sharing no code
 *      with pangenomesim_b200/csrc.  Pinned by the three Random123 Philox4x32-10 known-answer
 *      vectors and by closed-form binomial tail values.
 *
 *  (2) ref_model.c -- plain-C restatement of the reference's own algorithm for the path
 *      (minstd_rand0 engine + libstdc++ distributions + gene::gene / gene::mutate /
 *      sim_core_gene), pinned bit-exactly against FASTA produced by the unmodified reference
 *      binary (oracle/_ref/pangenomesim, tests/golden/).
 */
#ifndef PGS_ORACLE_H
#define PGS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- (1) Philox replay */

#define ORACLE_ROOT_EDGE 0xFFFFFFFFu
#define ORACLE_BLOCK_SITES 64

/* Philox4x32-10 (Salmon et al., SC'11; Random123).  out = philox(ctr, key). */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Per-edge decision table for one 64-site block.
 *   p      = 3/4 (1 - exp(-4/3 rate))                                  (Jukes-Cantor)
 *   K      ~ Binomial(64, p)
 *   H[k]   = floor(2^64 * P(K >= k)), k = 1..64, clamped to 2^64-1; H[0] = 2^64-1.
 * Returns kmax = largest k with H[k] > 0 (0 if none).  H must have room for 65 entries. */
int oracle_edge_table(double rate, uint64_t H[65]);

/* JC substitution probability for a branch, as used by oracle_edge_table. */
double oracle_jc_p(double rate);

/* Root (origin) block b of gene gene_id: 64 uniform 2-bit bases, sites >= gene_length zeroed. */
void oracle_root_block(uint64_t seed, uint32_t gene_id, uint32_t blk, int64_t gene_length,
					   uint32_t out[4]);

/* Sibling pairs (children of a node in ascending index: (c0,c1), (c2,c3), ...): leader[v] = first
 * node of v's pair, side[v] = 0 for the leader and 1 for the other. */
void oracle_sibling_pairs(int32_t n_nodes, const int32_t *parent, int32_t *leader, int32_t *side);

/* Carry block x across edge `edge` (= child node index) in place.  The top 16 bits of the 64-bit
 * uniform are the (blk%2)-th 16-bit field of word [2*side + (blk/2)%2] of
 * Philox(leader, gene, blk/4, stream 0); the low 48 bits and the site/base draws come from
 * Philox(edge, gene, blk, stream 1, 2, ...). */
void oracle_mutate_block(uint64_t seed, uint32_t leader, uint32_t side, uint32_t edge, uint32_t gene_id,
						 uint32_t blk, int64_t gene_length, const uint64_t H[65], uint32_t x[4]);

/* Row of (node, gene): root block at `origin`, then every edge on the path origin -> node.
 * dst receives 4*ceil(gene_length/64) words.  Returns 0, or -1 if node is not in origin's subtree. */
int oracle_row(uint64_t seed, int32_t n_nodes, const int32_t *parent, const double *rate,
			   int32_t node, int32_t origin, uint32_t gene_id, int64_t gene_length, uint32_t *dst);

/* All nodes of origin's subtree for one gene, top-down.  dst is [n_nodes][4*blocks] words;
 * rows of nodes outside the subtree are zero-filled.  Returns 0 on success. */
int oracle_gene_all_nodes(uint64_t seed, int32_t n_nodes, const int32_t *parent, const double *rate,
						  int32_t origin, uint32_t gene_id, int64_t gene_length, uint32_t *dst);

/* number of differing sites between two packed rows of `blocks` 64-site blocks */
uint64_t oracle_mismatch(const uint32_t *a, const uint32_t *b, int64_t blocks);

/* 2-bit -> ASCII ("ACGT"), gene_length chars, no terminator */
void oracle_unpack_ascii(const uint32_t *row, int64_t gene_length, char *dst);

/* ---------------------------------------------------------------- (2) reference restatement */

/* minstd_rand0 as libstdc++'s std::default_random_engine (bits/random.h): x <- 16807 x mod 2^31-1;
 * seeding maps 0 -> 1. */
typedef struct {
	uint32_t x;
} ref_engine;
void ref_engine_seed(ref_engine *e, uint64_t seed);
uint32_t ref_engine_next(ref_engine *e);

/* std::uniform_int_distribution<int>(0, hi)(engine), hi < engine range (uniform_int_dist.h). */
uint32_t ref_uniform_int(ref_engine *e, uint32_t hi);
/* std::uniform_real_distribution<double>(0,1)(engine) = generate_canonical<double,53> (random.tcc) */
double ref_uniform_real(ref_engine *e);

/* std::exponential_distribution<double>(lambda)(engine) */
double ref_exponential(ref_engine *e, double lambda);
/* create_coalescent(n) -- reference img.cxx:71-116; arrays of 2n-1 entries (root = 2n-2). */
void ref_coalescent(ref_engine *global, int64_t n, int32_t *parent, int32_t *left, int32_t *right,
					double *time);

/* gene::gene(len, ...) -- reference gene.cxx:13-25.  Writes len chars. */
void ref_gene_root(ref_engine *global, int64_t len, char *dst);
/* gene::mutate(rate) -- reference gene.cxx:27-61.  src/dst len chars; returns number of hits. */
int64_t ref_gene_mutate(ref_engine *global, const char *src, int64_t len, double rate, char *dst);
/* The same call without sequences: advances `global` exactly as gene::mutate does and returns
 * the number of substituted sites (SURVEY.md H3 "rng-compat"). */
int64_t ref_gene_mutate_count(ref_engine *global, int64_t len, double rate);

/* img_model::sim_core_gene (reference img.cxx:203-232) for one gene over a tree given as arrays
 * (left/right child or -1, branch time above each node).  leaves_out is [n_leaves][len] chars,
 * indexed by leaf_index[node]; root_out receives the root sequence.  DFS order and engine use are
 * those of tree_node::traverse (img.h:316-329). */
void ref_sim_core_gene(ref_engine *global, int32_t root, const int32_t *left, const int32_t *right,
					   const double *time, const int32_t *leaf_index, double mut_rate, int64_t len,
					   char *root_out, char *leaves_out);

/* E[#substituted sites] of one gene::mutate(rate) call on a gene of length len (selection
 * sampling with fractional quota): exact dynamic programme for quotas up to 64
 * hits, floor(quota)+1/2 beyond (the count is floor or ceil of the quota there). */
double ref_expected_hits(double rate, int64_t len);

#ifdef __cplusplus
}
#endif
#endif
