// philox.cuh -- counter-based RNG and the per-block substitution scheme.
//
// Replaces the reference's single global std::default_random_engine (global.h:6) that makes
// gene::gene (gene.cxx:13-25) and gene::mutate (gene.cxx:27-61) order-dependent.  Every draw is
// a pure function of (seed, edge, gene_id, site-block, stream), so results do not depend on
// thread order, grid shape or gene sharding.
//
//   key = (seed_lo, seed_hi); the ten round keys are launch constants (PhiloxKeys)
//   ctr = (0xFFFFFFFF, gene_id, block, 0)            origin rows
//   ctr = (leader edge, gene_id, block >> 2, 0)      decision fields of a sibling pair (see below)
//   ctr = (edge, gene_id, block, stream >= 1)        everything else a branch needs for a block
// edge = child node index of the branch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pgs {

constexpr uint32_t kRootEdge = 0xFFFFFFFFu;
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

struct PhiloxKeys {
	uint32_t k0[10];
	uint32_t k1[10];
};

inline PhiloxKeys make_philox_keys(uint64_t seed)
{
	PhiloxKeys k;
	uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
	for (int i = 0; i < 10; i++) {
		k.k0[i] = a;
		k.k1[i] = b;
		a += kPhiloxW0;
		b += kPhiloxW1;
	}
	return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
											   const PhiloxKeys &keys)
{
#pragma unroll
	for (int r = 0; r < 10; r++) {
		const uint32_t hi0 = __umulhi(kPhiloxM0, c0);
		const uint32_t lo0 = kPhiloxM0 * c0;
		const uint32_t hi1 = __umulhi(kPhiloxM1, c2);
		const uint32_t lo1 = kPhiloxM1 * c2;
		c0 = hi1 ^ c1 ^ keys.k0[r];
		c1 = lo1;
		c2 = hi0 ^ c3 ^ keys.k1[r];
		c3 = lo0;
	}
	return make_uint4(c0, c1, c2, c3);
}

// Per-edge decision data.  h1 = floor(2^64 P(K >= 1)) for K ~ Binomial(64, p_edge): a block is
// copied unchanged iff U >= h1.  tab points at H[1..kmax] (descending tail thresholds).
struct EdgeDecision {
	uint64_t h1;
	const uint64_t *tab; // tab[k-1] = H[k]
	int kmax;
};

// ---- decision fields ----------------------------------------------------------------------
// The branches below a node are taken in sibling pairs (children in ascending node index:
// (c0,c1), (c2,c3), ...); the first of a pair is its `leader`, side = 0 for the leader and 1 for
// the other.  One Philox call serves the EIGHT (edge, block) decisions of a sibling pair and a
// quad of four adjacent blocks (256 sites of both branches), 16 bits each:
//     W = Philox(leader, gid, blk >> 2, stream 0)
//     D(edge, gid, blk) = 16-bit field (blk & 1) of word W[2*side + ((blk >> 1) & 1)]
// (field 0 = low half, field 1 = high half).  D is the TOP 16 bits of the 64-bit uniform
// U = D * 2^48 + lo48 that is compared with the branch's thresholds H[k] = floor(2^64 P(K >= k)).
// With h16 = H[1] >> 48:  D > h16  =>  U >= H[1]  =>  no substitution in the block (the common
// case).  Otherwise the low 48 bits and everything else come from the branch's own stream
// Philox(edge, gid, blk, stream >= 1).
__device__ __forceinline__ uint4 quad_words(uint32_t leader, uint32_t gid, uint32_t blk, const PhiloxKeys &keys)
{
	return philox4x32_10(leader, gid, blk >> 2, 0u, keys);
}
// the word that holds the two fields of blocks (blk & ~1, blk | 1) on one side
__device__ __forceinline__ uint32_t pick_pair_word(const uint4 &w, uint32_t side, uint32_t blk)
{
	const uint32_t i = 2u * side + ((blk >> 1) & 1u);
	return i == 0 ? w.x : (i == 1 ? w.y : (i == 2 ? w.z : w.w));
}
__device__ __forceinline__ uint32_t pair_field(uint32_t pair_word, uint32_t odd)
{
	return odd ? (pair_word >> 16) : (pair_word & 0xFFFFu);
}
__device__ __forceinline__ uint32_t pick_field(const uint4 &w, uint32_t side, uint32_t blk)
{
	return pair_field(pick_pair_word(w, side, blk), blk & 1u);
}
// Compare constant of a branch: hc = (h16 << 16) | 0xFFFF.  For a pair word v:
//   block field 1 (high half) decides "may change"  iff  v <= hc
//   block field 0 (low half)                        iff  (v << 16) <= hc
__host__ __device__ __forceinline__ uint32_t decision_cmp(uint64_t h1)
{
	return ((uint32_t)(h1 >> 48) << 16) | 0xFFFFu;
}
// bit 0: even block of the pair may change, bit 1: odd block
__device__ __forceinline__ uint32_t pair_hits(uint32_t v, uint32_t hc)
{
	return ((v << 16) <= hc ? 1u : 0u) | (v <= hc ? 2u : 0u);
}

// Rare path (D <= h16): the block may change.  Stream 1 of the branch gives lo48 = x:y[31:16],
// then the words for K distinct sites (r-th free site, r = mulhi(word, 64-j)) each with a uniform
// non-zero 2-bit delta; XOR with a non-zero delta is "uniform pick among the 3 other bases".
// Words are consumed in order: z, w of stream 1, then x, y, z, w of stream 2, ...
// valid = number of real sites in this block (64 except for the last block of a gene); padding
// stays zero.
// tab8 = H[1..8] of the branch (any address space, e.g. a shared-memory copy), tab = H[1..kmax].
// K substitutions in block x: the picks read the branch's stream from word z of stream 1 (= w) on.
static __device__ __forceinline__ uint4 substitute_sites(uint4 x, uint4 w, int K, uint32_t edge, uint32_t gid,
														  uint32_t blk, int valid, const PhiloxKeys &keys)
{
	if (K == 1) { // nearly every block that changes at all changes in one site: no loop, no 64-bit shifts
		const uint32_t pos = __umulhi(w.z, 64u);
		const uint32_t delta = 1u + __umulhi(w.z * 64u, 3u);
		if ((int)pos < valid) { // (padding sites never change)
			const uint32_t m = delta << (2u * (pos & 15u)), wi = pos >> 4;
			x.x ^= wi == 0u ? m : 0u;
			x.y ^= wi == 1u ? m : 0u;
			x.z ^= wi == 2u ? m : 0u;
			x.w ^= wi == 3u ? m : 0u;
		}
		return x;
	}
	uint32_t stream = 1;
	uint64_t taken = 0, mlo = 0, mhi = 0;
	int have = 2; // unread words of the current call
	for (int j = 0; j < K; j++) {
		if (have == 0) {
			stream++;
			w = philox4x32_10(edge, gid, blk, stream, keys);
			have = 4;
		}
		uint32_t word;
		switch (have) {
			case 4: word = w.x; break;
			case 3: word = w.y; break;
			case 2: word = w.z; break;
			default: word = w.w; break;
		}
		have--;
		const uint32_t m = 64u - (uint32_t)j;
		const uint32_t r = __umulhi(word, m);
		const uint32_t lo = word * m;
		const uint64_t delta = 1u + __umulhi(lo, 3u);
		// r-th (0-based, ascending) site not yet taken; the first pick (most blocks on the rare path
		// change in exactly one site) needs no search
		int pos = (int)r;
		if (j > 0) {
			const uint64_t freebits = ~taken;
			const uint32_t flo = (uint32_t)freebits, fhi = (uint32_t)(freebits >> 32);
			const uint32_t nlo = __popc(flo);
			pos = r < nlo ? (int)__fns(flo, 0u, (int)r + 1) : 32 + (int)__fns(fhi, 0u, (int)(r - nlo) + 1);
		}
		taken |= 1ull << pos;
		if (pos < 32)
			mlo |= delta << (2 * pos);
		else
			mhi |= delta << (2 * (pos - 32));
	}
	if (valid < 64) { // last block of the gene: padding sites never change
		if (valid <= 32) {
			mhi = 0;
			mlo &= (valid == 32) ? ~0ull : ((1ull << (2 * valid)) - 1);
		} else {
			mhi &= (1ull << (2 * (valid - 32))) - 1;
		}
	}
	x.x ^= (uint32_t)mlo;
	x.y ^= (uint32_t)(mlo >> 32);
	x.z ^= (uint32_t)mhi;
	x.w ^= (uint32_t)(mhi >> 32);
	return x;
}

static __device__ __forceinline__ uint4 mutate_block_slow(uint4 x, uint32_t edge, uint32_t gid, uint32_t blk,
														   uint32_t D, const uint64_t *tab8, const uint64_t *__restrict__ tab,
														   int kmax, int valid, const PhiloxKeys &keys)
{
	const uint4 w = philox4x32_10(edge, gid, blk, 1u, keys);
	const uint64_t U = ((uint64_t)D << 48) | ((uint64_t)w.x << 16) | (uint64_t)(w.y >> 16);
	int K = 0;
	while (K < kmax && K < 8 && U < tab8[K]) K++;
	if (K == 8)
		while (K < kmax && U < tab[K]) K++;
	return substitute_sites(x, w, K, edge, gid, blk, valid, keys);
}

// The same with the branch's first two thresholds at hand (h1 = H[1], h2 = H[2], 0 beyond kmax): a
// block that changes in one site -- nearly all that change -- needs no table access at all; the
// table (tab_all + tab_off[edge], kmax[edge] entries) is read only when U < H[2].
static __device__ __forceinline__ uint4 mutate_block_h12(uint4 x, uint32_t edge, uint32_t gid, uint32_t blk, uint32_t D,
														  uint64_t h1, uint64_t h2, const uint64_t *__restrict__ tab_all,
														  const int32_t *__restrict__ tab_off, const int32_t *__restrict__ kmax_all,
														  int valid, const PhiloxKeys &keys)
{
	const uint4 w = philox4x32_10(edge, gid, blk, 1u, keys);
	const uint64_t U = ((uint64_t)D << 48) | ((uint64_t)w.x << 16) | (uint64_t)(w.y >> 16);
	if (U >= h1) return x;
	int K = 1;
	if (U < h2) {
		const int kmax = __ldg(kmax_all + edge);
		const uint64_t *tab = tab_all + __ldg(tab_off + edge);
		K = 2;
		while (K < kmax && U < __ldg(tab + K)) K++;
	}
	return substitute_sites(x, w, K, edge, gid, blk, valid, keys);
}

// Carry one 64-site block across one edge (generic form used where blocks are handled singly).
__device__ __forceinline__ uint4 mutate_block(uint4 x, uint32_t leader, uint32_t side, uint32_t edge, uint32_t gid,
											  uint32_t blk, const EdgeDecision &d, int valid, const PhiloxKeys &keys)
{
	const uint4 w = quad_words(leader, gid, blk, keys);
	const uint32_t D = pick_field(w, side, blk);
	if (D <= (uint32_t)(d.h1 >> 48)) return mutate_block_slow(x, edge, gid, blk, D, d.tab, d.tab, d.kmax, valid, keys);
	return x;
}

// Origin row block: 64 uniform bases, padding zeroed.
__device__ __forceinline__ uint4 root_block(uint32_t gid, uint32_t blk, int valid, const PhiloxKeys &keys)
{
	uint4 w = philox4x32_10(kRootEdge, gid, blk, 0u, keys);
	if (valid < 64) {
		uint64_t lo = ((uint64_t)w.y << 32) | w.x, hi = ((uint64_t)w.w << 32) | w.z;
		if (valid <= 32) {
			hi = 0;
			lo &= (valid == 32) ? ~0ull : ((1ull << (2 * valid)) - 1);
		} else {
			hi &= (1ull << (2 * (valid - 32))) - 1;
		}
		w = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
	}
	return w;
}

} // namespace pgs
