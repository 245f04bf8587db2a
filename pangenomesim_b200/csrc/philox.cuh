// philox.cuh -- counter-based RNG and the per-block substitution scheme.
//
// Replaces the reference's single global std::default_random_engine (global.h:6) that makes
// gene::gene (gene.cxx:13-25) and gene::mutate (gene.cxx:27-61) order-dependent.  Every draw is
// a pure function of (seed, edge, gene_id, site-block, stream), so results do not depend on
// thread order, grid shape or gene sharding.
//
//   ctr = (edge, gene_id, block, stream)   edge = child node index, 0xFFFFFFFF for origin rows
//   key = (seed_lo, seed_hi); the ten round keys are launch constants (PhiloxKeys)
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

// Rare path: the block has K >= 1 substitutions.  w is Philox(edge, gid, blk, stream 0), of which
// words x,y formed U.  K-1 further thresholds are scanned, then K distinct sites are drawn
// (r-th free site, r = mulhi(word, 64-j)) each with a uniform non-zero 2-bit delta; XOR with a
// non-zero delta is "uniform pick among the 3 other bases".  valid = number of real sites in
// this block (64 except for the last block of a gene); padding stays zero.
static __device__ __noinline__ uint4 mutate_block_slow(uint4 x, uint32_t edge, uint32_t gid, uint32_t blk,
												uint4 w, uint64_t U, const uint64_t *__restrict__ tab,
												int kmax, int valid, const PhiloxKeys &keys)
{
	int K = 1;
	while (K < kmax && U < tab[K]) K++;

	uint64_t taken = 0, mlo = 0, mhi = 0;
	uint32_t stream = 0;
	int have = 2; // unread words of the current call; stream 0 has w.z, w.w left
	for (int j = 0; j < K; j++) {
		if (have == 0) {
			stream++;
			w = philox4x32_10(edge, gid, blk, stream, keys);
			have = 4;
		}
		uint32_t word;
		switch (have) { // words are consumed in index order x,y,z,w
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
		uint64_t freebits = ~taken;
		for (uint32_t i = 0; i < r; i++) freebits &= freebits - 1; // drop the r lowest free sites
		const int pos = __ffsll((long long)freebits) - 1;
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

// Carry one 64-site block across one edge.
__device__ __forceinline__ uint4 mutate_block(uint4 x, uint32_t edge, uint32_t gid, uint32_t blk,
											  const EdgeDecision &d, int valid, const PhiloxKeys &keys)
{
	const uint4 w = philox4x32_10(edge, gid, blk, 0u, keys);
	const uint64_t U = ((uint64_t)w.y << 32) | w.x;
	if (U < d.h1) return mutate_block_slow(x, edge, gid, blk, w, U, d.tab, d.kmax, valid, keys);
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
