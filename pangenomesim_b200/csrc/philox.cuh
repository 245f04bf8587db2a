// philox.cuh -- counter-based RNG and the per-block substitution scheme.
//
// Replaces the reference's single global std::default_random_engine (global.h:6) that makes
// gene::gene (gene.cxx:13-25) and gene::mutate (gene.cxx:27-61) order-dependent.  Every draw is
// a pure function of (seed, edge, gene_id, site-block, stream), so results do not depend on
// thread order, grid shape or gene sharding.
//
//   key = (seed_lo, seed_hi); the ten round keys are launch constants (PhiloxKeys)
//   ctr = (0xFFFFFFFF, gene_id, block, 0)            origin rows
//   ctr = (leader edge, gene_id, block >> 1, 0)      decision words of a sibling pair (see below)
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

// ---- decision words -----------------------------------------------------------------------
// The branches below a node are taken in sibling pairs (children in ascending node index:
// (c0,c1), (c2,c3), ...); the first of a pair is its `leader`, side = 0 for the leader and 1 for
// the other.  One Philox call serves the four (edge, block) decisions of a sibling pair and two
// adjacent blocks:
//     D(edge, gid, blk) = word[2*side + (blk & 1)] of Philox(leader, gid, blk >> 1, stream 0)
// D is the HIGH half of the 64-bit uniform U = D * 2^32 + lo that is compared with the branch's
// thresholds H[k] = floor(2^64 P(K >= k)).  With h32 = H[1] >> 32:  D > h32  =>  U >= H[1]  =>
// no substitution in the block (the common case: one compare).  Otherwise the low half and
// everything else comes from the branch's own stream Philox(edge, gid, blk, stream >= 1).
__device__ __forceinline__ uint4 pair_words(uint32_t leader, uint32_t gid, uint32_t blk, const PhiloxKeys &keys)
{
	return philox4x32_10(leader, gid, blk >> 1, 0u, keys);
}
__device__ __forceinline__ uint32_t pick_word(const uint4 &w, uint32_t side, uint32_t blk)
{
	const uint32_t i = 2u * side + (blk & 1u);
	return i == 0 ? w.x : (i == 1 ? w.y : (i == 2 ? w.z : w.w));
}

// Rare path (D <= h32): the block may change.  Stream 1 of the branch gives lo, then the words
// for K distinct sites (r-th free site, r = mulhi(word, 64-j)) each with a uniform non-zero
// 2-bit delta; XOR with a non-zero delta is "uniform pick among the 3 other bases".  Words are
// consumed in order: x (= lo), y, z, w of stream 1, then x, y, z, w of stream 2, ...
// valid = number of real sites in this block (64 except for the last block of a gene); padding
// stays zero.
// tab8 = H[1..8] of the branch (any address space, e.g. a shared-memory copy), tab = H[1..kmax].
static __device__ __forceinline__ uint4 mutate_block_slow(uint4 x, uint32_t edge, uint32_t gid, uint32_t blk,
													   uint32_t D, const uint64_t *tab8, const uint64_t *__restrict__ tab,
													   int kmax, int valid, const PhiloxKeys &keys)
{
	uint32_t stream = 1;
	uint4 w = philox4x32_10(edge, gid, blk, stream, keys);
	const uint64_t U = ((uint64_t)D << 32) | w.x;
	int K = 0;
	while (K < kmax && K < 8 && U < tab8[K]) K++;
	if (K == 8)
		while (K < kmax && U < tab[K]) K++;

	uint64_t taken = 0, mlo = 0, mhi = 0;
	int have = 3; // unread words of the current call
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
		// r-th (0-based, ascending) site not yet taken
		const uint64_t freebits = ~taken;
		const uint32_t flo = (uint32_t)freebits, fhi = (uint32_t)(freebits >> 32);
		const uint32_t nlo = __popc(flo);
		const int pos = r < nlo ? (int)__fns(flo, 0u, (int)r + 1) : 32 + (int)__fns(fhi, 0u, (int)(r - nlo) + 1);
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

// Carry one 64-site block across one edge (generic form used where blocks are handled singly).
__device__ __forceinline__ uint4 mutate_block(uint4 x, uint32_t leader, uint32_t side, uint32_t edge, uint32_t gid,
											  uint32_t blk, const EdgeDecision &d, int valid, const PhiloxKeys &keys)
{
	const uint4 w = pair_words(leader, gid, blk, keys);
	const uint32_t D = pick_word(w, side, blk);
	if (D <= (uint32_t)(d.h1 >> 32)) return mutate_block_slow(x, edge, gid, blk, D, d.tab, d.tab, d.kmax, valid, keys);
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
