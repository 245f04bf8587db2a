// kernels.cu -- K0 (origin rows) and K1 (tree sweep: fused walks and the level-synchronous kernels).
// Hand-written for sm_100a; integer / byte work bounded by HBM,
// so no tensor cores: coalesced 16-byte accesses, several loads in flight per thread, grids that
// are whole multiples of what the 148 SMs hold.
#include "kernels.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace pgs {

// ---------------------------------------------------------------------------------------------
// memory helpers: rows are streamed (read once, written once per level), so keep them out of L1.

__device__ __forceinline__ uint4 ld_stream(const uint4 *p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
				 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
				 : "l"(p));
	return r;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v)
{
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
				 "r"(v.z), "r"(v.w)
				 : "memory");
}

__device__ __forceinline__ EdgeDecision load_edge(const DeviceTables &t, int32_t node)
{
	EdgeDecision d;
	d.h1 = __ldg(t.h1 + node);
	d.tab = t.tab + __ldg(t.tab_off + node);
	d.kmax = __ldg(t.kmax + node);
	return d;
}

// ---------------------------------------------------------------------------------------------
// K0: origin rows.  One thread per 64-site block; write-only.
// Replaces gene::gene(len, -1, id), reference gene.cxx:13-25.

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
k0_rootgen(const __grid_constant__ DeviceTables t, const RootTask *__restrict__ tasks, int64_t n_items)
{
	const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (idx >= n_items) return;
	const uint32_t task = (uint32_t)(idx / t.blocks);
	const uint32_t blk = (uint32_t)(idx - (int64_t)task * t.blocks);
	const RootTask rt = tasks[task];
	const int valid = (blk + 1 == t.blocks) ? t.last_valid : 64;
	const uint4 v = root_block(rt.gid, blk, valid, t.keys);
	st_stream(t.rows + (int64_t)rt.row * t.blocks + blk, v);
}

void launch_rootgen(const DeviceTables &t, const RootTask *tasks, int64_t n_tasks, cudaStream_t s)
{
	const int64_t items = n_tasks * t.blocks;
	if (items == 0) return;
	const int64_t grid = (items + kThreads - 1) / kThreads;
	k0_rootgen<<<(unsigned)grid, kThreads, 0, s>>>(t, tasks, items);
}

// ---------------------------------------------------------------------------------------------
// K1, dense region.  The dense genes (carried by every genome, origin = root) occupy the first
// n_dense rows of every node, in the same order, so a (parent, left, right) triple is a plain
// 1-D stream: item i of the parent region produces item i of both child regions.  Each parent
// block is read once and both children are produced from it -- the read-once / write-once
// traffic of SURVEY.md §8(d): 48 bytes per (block, parent).
// Replaces the inner loop of img_model::sim_core_gene, reference img.cxx:215-231, whose
// per-edge body is gene::mutate, gene.cxx:27-61.
//
// Main variant (even number of blocks per gene): a thread owns PAIRS pairs of adjacent blocks,
// moved with 256-bit loads/stores (LDG.256 / STG.256).  One Philox call yields the eight 16-bit
// decision fields of {left, right} x {four blocks of a quad}.  QUAD (blocks per gene a multiple
// of four): pairs 2i and 2i+1 of a thread are the two halves of one quad and share the call;
// otherwise every pair makes the call of its quad itself and uses half of it.  The fast path is
// branch-free (all loads first, independent Philox chains, compares folded into a bit mask) and
// substitutions -- rare events -- are applied afterwards.

struct __align__(32) U32x8 {
	uint4 lo, hi; // blocks 2p and 2p+1
};

__device__ __forceinline__ U32x8 ld_stream256(const U32x8 *p)
{
	U32x8 r;
	asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				 : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y),
				   "=r"(r.hi.z), "=r"(r.hi.w)
				 : "l"(p));
	return r;
}
// The same through the coherent path, for rows that the grid BEFORE this one wrote while this grid may already
// have been resident (programmatic dependent launch): ptxas treats ld.global.nc as a load of read-only data and
// schedules it ABOVE griddepcontrol.wait (SASS: LDG...CONSTANT before ACQBULK) -- the walk then read its origin
// row before k1_top_masks had written it whenever that grid was slow to start (several processes on one GPU).
__device__ __forceinline__ U32x8 ld_written256(const U32x8 *p)
{
	U32x8 r;
	asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				 : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y),
				   "=r"(r.hi.z), "=r"(r.hi.w)
				 : "l"(p)
				 : "memory");
	return r;
}
__device__ __forceinline__ void st_stream256(U32x8 *p, const U32x8 &v)
{
	asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.lo.x),
				 "r"(v.lo.y), "r"(v.lo.z), "r"(v.lo.w), "r"(v.hi.x), "r"(v.hi.y), "r"(v.hi.z), "r"(v.hi.w)
				 : "memory");
}

// pair index of a thread's j-th pair inside its tile
template <int PAIRS, bool QUAD>
__device__ __forceinline__ uint32_t tile_pair(uint32_t tile, int j)
{
	return QUAD ? 2u * ((tile * (PAIRS / 2) + (uint32_t)(j >> 1)) * kThreads + threadIdx.x) + (uint32_t)(j & 1)
				: (tile * PAIRS + (uint32_t)j) * kThreads + threadIdx.x;
}

template <int PAIRS, int MINB, bool QUAD, bool COPY_ONLY = false>
__global__ void __launch_bounds__(kThreads, MINB)
k1_sweep_dense(const __grid_constant__ DeviceTables t, const DenseTask *__restrict__ tasks,
			   uint32_t tiles_per_task, uint32_t pairs_per_task)
{
	static_assert(!QUAD || PAIRS % 2 == 0, "a quad is two pairs");
	constexpr uint32_t kTile = kThreads * PAIRS;
	const uint32_t task = blockIdx.x / tiles_per_task;
	const uint32_t tile = blockIdx.x - task * tiles_per_task;
	const DenseTask dt = tasks[task];
	const bool has_right = dt.right >= 0;
	const int32_t right = has_right ? dt.right : dt.left;

	const U32x8 *__restrict__ src = reinterpret_cast<const U32x8 *>(t.rows + __ldg(t.row_base + dt.parent) * t.blocks);
	const bool full = (tile + 1) * kTile <= pairs_per_task; // CTA-uniform

	U32x8 x[PAIRS];
#pragma unroll
	for (int j = 0; j < PAIRS; j++) {
		const uint32_t p = tile_pair<PAIRS, QUAD>(tile, j);
		if (full || p < pairs_per_task) x[j] = ld_stream256(src + p);
	}

	// Copy first: in the common case a child block equals its parent block, so the stores do not
	// wait for the random numbers at all.  Blocks that do change are rewritten below (the second
	// store of the same thread to the same address lands in L2 long before the line is evicted).
	U32x8 *__restrict__ dl = reinterpret_cast<U32x8 *>(t.rows + __ldg(t.row_base + dt.left) * t.blocks);
	U32x8 *__restrict__ dr = reinterpret_cast<U32x8 *>(t.rows + __ldg(t.row_base + right) * t.blocks);
#pragma unroll
	for (int j = 0; j < PAIRS; j++) {
		const uint32_t p = tile_pair<PAIRS, QUAD>(tile, j);
		if (full || p < pairs_per_task) {
			st_stream256(dl + p, x[j]);
			if (has_right) st_stream256(dr + p, x[j]);
		}
	}
	if (COPY_ONLY) return;

	const uint32_t hcl = decision_cmp(__ldg(t.h1 + dt.left));
	const uint32_t hcr = decision_cmp(__ldg(t.h1 + right));
	uint32_t hits = 0; // 4 bits per pair: left even, left odd, right even, right odd
	uint4 w = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
	for (int j = 0; j < PAIRS; j++) {
		const uint32_t idx = 2u * tile_pair<PAIRS, QUAD>(tile, j); // even block index inside the region
		const bool live = full || idx < 2u * pairs_per_task;
		uint32_t wl, wr;
		if (!QUAD || (j & 1) == 0) {
			const uint32_t slot = fastdiv(idx, t.div_blocks);
			const uint32_t blk = idx - slot * t.blocks;
			const uint32_t gid = live ? __ldg(t.dense_gid + slot) : 0u;
			w = quad_words((uint32_t)dt.left, gid, blk, t.keys);
			wl = QUAD ? w.x : pick_pair_word(w, 0u, blk);
			wr = QUAD ? w.z : pick_pair_word(w, 1u, blk);
		} else {
			wl = w.y;
			wr = w.w;
		}
		const uint32_t m = pair_hits(wl, hcl) | (pair_hits(wr, hcr) << 2);
		hits |= live ? (m << (4 * j)) : 0u;
	}
	if (!has_right) hits &= 0x33333333u;

	if (!__any_sync(0xFFFFFFFFu, hits != 0)) return;

	// This warp has work on the rare path.  Its lanes fetch the two branches' first thresholds
	// (H[1..8] each) and table lengths in ONE round trip into a per-warp shared-memory copy, so
	// the per-block work below has no dependent global load.
	__shared__ uint64_t s_tab[kThreads / 32][2][8];
	__shared__ int s_kmax[kThreads / 32][2];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane < 16)
		s_tab[warp][lane >> 3][lane & 7] = __ldg(t.tab8 + (size_t)(lane < 8 ? dt.left : right) * 8 + (lane & 7));
	else if (lane < 18)
		s_kmax[warp][lane - 16] = __ldg(t.kmax + (lane == 16 ? dt.left : right));
	__syncwarp();

	// rare: some of this thread's child blocks (may) change.  Every lane works through its own
	// flagged blocks one at a time, so a warp needs as many passes as its busiest lane has hits
	// (usually one).  The thresholds come from this warp's shared-memory copy: no dependent global
	// load sits between the decision and the rewrite.
	while (hits) {
		const uint32_t b = (uint32_t)__ffs((int)hits) - 1u;
		hits &= hits - 1u;
		const uint32_t j = b >> 2, side = (b >> 1) & 1u, odd = b & 1u;
		const uint32_t p = tile_pair<PAIRS, QUAD>(tile, (int)j);
		const uint32_t idx = 2u * p;
		const uint32_t slot = fastdiv(idx, t.div_blocks);
		const uint32_t blk0 = idx - slot * t.blocks;
		const uint32_t gid = __ldg(t.dense_gid + slot);
		const uint32_t blk = blk0 + odd;
		const uint32_t D = pick_field(quad_words((uint32_t)dt.left, gid, blk, t.keys), side, blk);
		const int valid = (blk + 1 == t.blocks) ? t.last_valid : 64;
		uint4 xin = x[0].lo;
#pragma unroll
		for (int jj = 0; jj < PAIRS; jj++)
			if (j == (uint32_t)jj) xin = odd ? x[jj].hi : x[jj].lo;
		const int kmax = s_kmax[warp][side];
		const uint32_t edge = side ? (uint32_t)right : (uint32_t)dt.left;
		// the full table is only touched when more than eight sites of one block change
		const uint64_t *tab = kmax > 8 ? t.tab + __ldg(t.tab_off + edge) : nullptr;
		uint4 *dst = reinterpret_cast<uint4 *>((side ? dr : dl) + p) + odd;
		st_stream(dst, mutate_block_slow(xin, edge, gid, blk, D, s_tab[warp][side], tab, kmax, valid, t.keys));
	}
}

// Fallback for an odd number of blocks per gene (adjacent blocks may straddle genes and rows are
// only 16-byte aligned): one block per thread, each computing its pair's Philox call itself.
__global__ void __launch_bounds__(kThreads)
k1_sweep_dense_single(const __grid_constant__ DeviceTables t, const DenseTask *__restrict__ tasks,
					  uint32_t tiles_per_task, uint32_t items_per_task)
{
	const uint32_t task = blockIdx.x / tiles_per_task;
	const uint32_t tile = blockIdx.x - task * tiles_per_task;
	const DenseTask dt = tasks[task];
	const uint32_t idx = tile * kThreads + threadIdx.x;
	if (idx >= items_per_task) return;
	const uint4 x = ld_stream(t.rows + __ldg(t.row_base + dt.parent) * t.blocks + idx);
	const uint32_t slot = fastdiv(idx, t.div_blocks);
	const uint32_t blk = idx - slot * t.blocks;
	const uint32_t gid = __ldg(t.dense_gid + slot);
	const int valid = (blk + 1 == t.blocks) ? t.last_valid : 64;
	{
		const EdgeDecision d = load_edge(t, dt.left);
		st_stream(t.rows + __ldg(t.row_base + dt.left) * t.blocks + idx,
				  mutate_block(x, (uint32_t)dt.left, 0u, (uint32_t)dt.left, gid, blk, d, valid, t.keys));
	}
	if (dt.right >= 0) {
		const EdgeDecision d = load_edge(t, dt.right);
		st_stream(t.rows + __ldg(t.row_base + dt.right) * t.blocks + idx,
				  mutate_block(x, (uint32_t)dt.left, 1u, (uint32_t)dt.right, gid, blk, d, valid, t.keys));
	}
}

template <int PAIRS, int MINB, bool QUAD, bool COPY_ONLY = false>
static void launch_dense_variant(const DeviceTables &t, const DenseTask *tasks, int32_t n_tasks, cudaStream_t s)
{
	const uint32_t pairs = t.n_dense * t.blocks / 2;
	const uint32_t tile = kThreads * PAIRS;
	const uint32_t tiles = (pairs + tile - 1) / tile;
	// one launch covers at most 2^31-1 CTAs; chunk the task list beyond that
	const int64_t max_tasks = (int64_t)0x7FFFFFFF / tiles;
	for (int64_t done = 0; done < n_tasks; done += max_tasks) {
		const int64_t now = (n_tasks - done < max_tasks) ? (n_tasks - done) : max_tasks;
		k1_sweep_dense<PAIRS, MINB, QUAD, COPY_ONLY><<<(unsigned)(now * tiles), kThreads, 0, s>>>(t, tasks + done, tiles, pairs);
	}
}

int g_dense_variant = -1; // development knob (PGS_K1_VARIANT); -1 = default

void launch_sweep_dense(const DeviceTables &t, const DenseTask *tasks, int32_t n_tasks, cudaStream_t s)
{
	if (n_tasks == 0 || t.n_dense == 0) return;
	if (t.blocks & 1u) {
		const uint32_t items = t.n_dense * t.blocks;
		const uint32_t tiles = (items + kThreads - 1) / kThreads;
		const int64_t max_tasks = (int64_t)0x7FFFFFFF / tiles;
		for (int64_t done = 0; done < n_tasks; done += max_tasks) {
			const int64_t now = (n_tasks - done < max_tasks) ? (n_tasks - done) : max_tasks;
			k1_sweep_dense_single<<<(unsigned)(now * tiles), kThreads, 0, s>>>(t, tasks + done, tiles, items);
		}
		return;
	}
	if (t.blocks & 3u) { // blocks = 2 mod 4: every pair makes its quad's call itself
		launch_dense_variant<4, 3, false>(t, tasks, n_tasks, s);
		return;
	}
	switch (g_dense_variant) { // development knob
		case 1: launch_dense_variant<4, 3, false>(t, tasks, n_tasks, s); break;
		case 2: launch_dense_variant<4, 3, true, true>(t, tasks, n_tasks, s); break; // memory ceiling probe: no RNG
		default: launch_dense_variant<4, 3, true>(t, tasks, n_tasks, s); break;
	}
}

// ---------------------------------------------------------------------------------------------
// K1, fused column walk (default for binary trees).  Columns are independent given the tree, so a
// thread can take one pair of blocks of one gene and carry it through a whole subtree by itself:
// the value of the child it descends into stays in REGISTERS, only the sibling that is visited
// later is parked in memory (and re-read from L2 a few steps later), and internal nodes that no
// output needs are never written.  HBM traffic falls from "every row written + every parent
// read" (0.375 B per site-edge) towards "leaves written once" (0.25 B per leaf nucleotide).
// The host cuts the tree into a few dozen subtrees (units) and orders each unit's steps
// depth-first, smaller subtree first, so at most log2(n) parked rows per unit are alive;
// a CTA = (unit, slab of 256 block pairs); the RNG keying makes the result independent of all this.

// leaf rows are not touched again during the sweep: first in line for eviction from L2 ...
__device__ __forceinline__ void st_leaf256(U32x8 *p, const U32x8 &v)
{
	asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
				 "r"(v.lo.x), "r"(v.lo.y), "r"(v.lo.z), "r"(v.lo.w), "r"(v.hi.x), "r"(v.hi.y), "r"(v.hi.z), "r"(v.hi.w)
				 : "memory");
}
// ... parked siblings come back within a few steps: last in line
__device__ __forceinline__ void st_park256(U32x8 *p, const U32x8 &v)
{
	asm volatile("st.global.L1::no_allocate.L2::evict_last.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
				 "r"(v.lo.x), "r"(v.lo.y), "r"(v.lo.z), "r"(v.lo.w), "r"(v.hi.x), "r"(v.hi.y), "r"(v.hi.z), "r"(v.hi.w)
				 : "memory");
}

// Predicated re-load of a parked row straight into the registers that carry the running value
// (read-write operands, two 128-bit loads: a 256-bit load wants an 8-aligned register octet, which
// made ptxas keep a second copy of the value and shuffle 16 registers per step).
__device__ __forceinline__ void ld_own256_if(U32x8 &v, const U32x8 *p, uint32_t pred)
{
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %9, 0;\n\t"
				 "@q ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%8];\n\t"
				 "@q ld.global.L1::no_allocate.v4.u32 {%4,%5,%6,%7}, [%8+16];\n\t}"
				 : "+r"(v.lo.x), "+r"(v.lo.y), "+r"(v.lo.z), "+r"(v.lo.w), "+r"(v.hi.x), "+r"(v.hi.y), "+r"(v.hi.z),
				   "+r"(v.hi.w)
				 : "l"(p), "r"(pred)
				 : "memory");
}

__device__ __forceinline__ U32x8 operator^(const U32x8 &a, const U32x8 &b)
{
	U32x8 r;
	r.lo = make_uint4(a.lo.x ^ b.lo.x, a.lo.y ^ b.lo.y, a.lo.z ^ b.lo.z, a.lo.w ^ b.lo.w);
	r.hi = make_uint4(a.hi.x ^ b.hi.x, a.hi.y ^ b.hi.y, a.hi.z ^ b.hi.z, a.hi.w ^ b.hi.w);
	return r;
}

#ifndef PGS_WALK_CHUNK
#define PGS_WALK_CHUNK 12
#endif
constexpr int kWalkChunk = PGS_WALK_CHUNK; // steps a warp stages in shared memory at a time

// QUAD (blocks per gene a multiple of four): pairs 2i, 2i+1 of a thread are the halves of one
// quad of blocks -- 64 contiguous bytes, one Philox call per step for the eight decisions.
template <int MINB, int PAIRS, bool QUAD, bool HINTS = true>
__global__ void __launch_bounds__(kThreads, MINB)
k1_walk_dense(const __grid_constant__ DeviceTables t, const __grid_constant__ WalkUnits u, uint32_t pairs_per_node)
{
	static_assert(!QUAD || PAIRS % 2 == 0, "a quad is two pairs");
	// Units are sorted by decreasing length and CTAs start in index order: unit-major numbering
	// (all slabs of the longest unit first, the shortest units last) keeps the tail of the launch short.
	// (config 4: 2.28 -> 2.11 ms against slab-major numbering; 625 genes: 0.49 -> 0.36 ms)
	const WalkStep *__restrict__ const steps = u.steps;
	const uint32_t n_slabs = gridDim.x / (uint32_t)u.n_units;
	const uint32_t unit = blockIdx.x / n_slabs, slab = blockIdx.x - unit * n_slabs;
	// this thread's PAIRS columns (block pairs) inside any node's dense region
	uint32_t p[PAIRS], gid[PAIRS], blk[PAIRS];
	bool active[PAIRS];
	int valid_hi[PAIRS];
	U32x8 *column[PAIRS];
#pragma unroll
	for (int j = 0; j < PAIRS; j++) {
		p[j] = tile_pair<PAIRS, QUAD>(slab, j);
		active[j] = p[j] < pairs_per_node;
		const uint32_t idx = 2u * p[j];
		const uint32_t slot = fastdiv(idx, t.div_blocks);
		blk[j] = idx - slot * t.blocks;
		gid[j] = active[j] ? __ldg(t.dense_gid + slot) : 0u;
		valid_hi[j] = (blk[j] + 2 == t.blocks) ? t.last_valid : 64;
		column[j] = reinterpret_cast<U32x8 *>(t.rows) + p[j]; // step offsets are in uint4 units
	}

	// every warp stages its own copy of the next steps: warps of a CTA drift apart (rare path) and
	// must not wait for each other at a block barrier
	__shared__ WalkStep s_steps_all[kThreads / 32][kWalkChunk];
	// Parked siblings: a thread's own stack, [level][pair][half][thread] (thread-private, so no
	// synchronisation; consecutive threads hit consecutive 16-byte words: conflict-free).
	extern __shared__ uint4 s_park[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	U32x8 cur[PAIRS]; // the values carried from step to step
#pragma unroll
	for (int j = 0; j < PAIRS; j++) {
		cur[j].lo = make_uint4(0u, 0u, 0u, 0u);
		cur[j].hi = cur[j].lo;
	}
	// Launched as a programmatic dependent of k1_top_masks (launch_walk_variant): this CTA may have become
	// resident while the last CTAs of that grid were still drawing masks.  Everything above reads only
	// tables uploaded by pgs_set_genes; the rows below are valid once the whole preceding grid has
	// completed and flushed, which is what this waits for (a no-op without the launch attribute).
	if (u.trace && threadIdx.x == 0) {
		unsigned long long now;
		uint32_t sm;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
		asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
		u.trace[4 * blockIdx.x] = now;
		u.trace[4 * blockIdx.x + 3] = sm;
	}
	const int32_t path_end = __ldg(u.unit_path_off + unit + 1), path_begin = __ldg(u.unit_path_off + unit);
	const int64_t store = __ldg(u.unit_store + unit);
	const int64_t begin = __ldg(u.unit_off + unit), end = __ldg(u.unit_off + unit + 1);
	// (the operands tie the table loads above to this side of the wait: ptxas otherwise moves the wait to the kernel's top)
	asm volatile("griddepcontrol.wait; // %0 %1 %2 %3 %4 %5 %6" ::"r"(gid[0]), "r"(gid[PAIRS - 1]), "r"(path_end), "r"(path_begin),
				 "l"(store), "l"(begin), "l"(end)
				 : "memory");
	{
		// The unit's root value: the origin row XORed with the masks of the branches on the path from
		// the tree's root (k1_top_masks wrote both; every load is independent of the others).
		const uint4 *const base = reinterpret_cast<const uint4 *>(t.rows);
#pragma unroll
		for (int j = 0; j < PAIRS; j++)
			if (active[j]) cur[j] = ld_written256(reinterpret_cast<const U32x8 *>(base + u.root_off) + p[j]);
		const int32_t pe = path_end;
		int32_t k = path_begin;
		for (; k + 2 <= pe; k += 2) { // two masks in flight per pair
			const int64_t off0 = __ldg(u.unit_path + k), off1 = __ldg(u.unit_path + k + 1);
#pragma unroll
			for (int j = 0; j < PAIRS; j++) {
				if (!active[j]) continue;
				const U32x8 m0 = ld_written256(reinterpret_cast<const U32x8 *>(base + off0) + p[j]);
				const U32x8 m1 = ld_written256(reinterpret_cast<const U32x8 *>(base + off1) + p[j]);
				cur[j] = cur[j] ^ m0 ^ m1;
			}
		}
		if (k < pe) {
			const int64_t off = __ldg(u.unit_path + k);
#pragma unroll
			for (int j = 0; j < PAIRS; j++)
				if (active[j]) cur[j] = cur[j] ^ ld_written256(reinterpret_cast<const U32x8 *>(base + off) + p[j]);
		}
		if (store >= 0) { // the unit's root is a genome directly below a top node
#pragma unroll
			for (int j = 0; j < PAIRS; j++)
				if (active[j]) st_leaf256(reinterpret_cast<U32x8 *>(reinterpret_cast<uint4 *>(t.rows) + store) + p[j], cur[j]);
		}
	}
	if (u.trace && threadIdx.x == 0) { // (the operand makes the clock wait for the root value)
		unsigned long long now;
		asm volatile("mov.u64 %0, %%globaltimer; // %1" : "=l"(now) : "r"(cur[0].lo.x ^ cur[PAIRS - 1].hi.w));
		u.trace[4 * blockIdx.x + 1] = now;
	}
	for (int64_t i0 = begin; i0 < end; i0 += kWalkChunk) {
		const int cnt = (int)min((int64_t)kWalkChunk, end - i0);
		WalkStep *const s_steps = s_steps_all[warp];
		__syncwarp();
		for (int q = lane; q < kWalkStepQuads * cnt; q += 32)
			reinterpret_cast<uint4 *>(s_steps)[q] = __ldg(reinterpret_cast<const uint4 *>(steps + i0) + q);
		__syncwarp();
		for (int k = 0; k < cnt; k++) {
			const WalkStep &st = s_steps[k];
			const uint32_t f = st.flags;
			// threads beyond the last pair of the region (tail slab only) keep in step with no effect
			if (f & kWalkLoadSmem) {
				const uint4 *src = s_park + ((st.slots >> 8) & 255u) * (PAIRS * 2 * kThreads) + threadIdx.x;
#pragma unroll
				for (int j = 0; j < PAIRS; j++) {
					cur[j].lo = src[(2 * j) * kThreads];
					cur[j].hi = src[(2 * j + 1) * kThreads];
				}
			} else {
#pragma unroll
				for (int j = 0; j < PAIRS; j++)
					ld_own256_if(cur[j], reinterpret_cast<const U32x8 *>(reinterpret_cast<const uint4 *>(column[j]) + st.in_off),
								 (f & kWalkLoad) && active[j] ? 1u : 0u);
			}
			const uint32_t edge_a = st.edge_a, hca = st.hca, hcb = st.hcb;
			uint32_t wa[PAIRS], wb[PAIRS]; // the pair words of branch a and branch b
			uint32_t m = 0;
#pragma unroll
			for (int j = 0; j < PAIRS; j++) {
				if (!QUAD || (j & 1) == 0) {
					const uint4 w = quad_words(edge_a, gid[j], blk[j], t.keys);
					wa[j] = QUAD ? w.x : pick_pair_word(w, 0u, blk[j]);
					wb[j] = QUAD ? w.z : pick_pair_word(w, 1u, blk[j]);
					if constexpr (QUAD) {
						wa[j | 1] = w.y;
						wb[j | 1] = w.w;
					}
				}
				const uint32_t mj = pair_hits(wa[j], hca) | (pair_hits(wb[j], hcb) << 2);
				if (active[j]) m |= mj << (4 * j);
			}
			uint4 *const park = s_park + (st.slots & 255u) * (PAIRS * 2 * kThreads) + threadIdx.x;
#pragma unroll
			for (int j = 0; j < PAIRS; j++) {
				if (!active[j]) continue;
				uint32_t mj = (m >> (4 * j)) & 15u;
				U32x8 *const dst_a = reinterpret_cast<U32x8 *>(reinterpret_cast<uint4 *>(column[j]) + st.a_off);
				U32x8 *const dst_b = reinterpret_cast<U32x8 *>(reinterpret_cast<uint4 *>(column[j]) + st.b_off);
				// child `side` (0 = a, 1 = b) takes value v: to its row, to the global pool or to the stack
				auto put = [&](uint32_t side, const U32x8 &v) {
					if (f & (side ? kWalkStoreB : kWalkStoreA)) {
						U32x8 *const dst = side ? dst_b : dst_a;
						if (!HINTS) st_stream256(dst, v); else if (f & (side ? kWalkParkB : kWalkParkA)) st_park256(dst, v); else st_leaf256(dst, v);
					} else if (f & (side ? kWalkSmemB : kWalkSmemA)) {
						park[(2 * j) * kThreads] = v.lo;
						park[(2 * j + 1) * kThreads] = v.hi;
					}
				};
				if (mj == 0) { // common case: both children are copies of the carried value
					put(0u, cur[j]);
					put(1u, cur[j]);
				} else { // rare: a block of a child may change
					// The child that is NOT carried on goes first, into a temporary; the carried child (or
					// b when nothing is carried: the value is dead after this step) is then changed in
					// place, so the running value never lives in two register sets.
					const uint32_t edge_b = st.edge_b;
					const uint32_t second = (f & kWalkCarryA) ? 0u : 1u, first = second ^ 1u;
					{
						U32x8 y = cur[j];
						uint32_t mf = (mj >> (2 * first)) & 3u;
						while (mf) {
							const uint32_t odd = (mf & 1u) ? 0u : 1u;
							mf &= mf - 1u;
							const uint32_t edge = first ? edge_b : edge_a;
							const uint4 out = mutate_block_h12(odd ? cur[j].hi : cur[j].lo, edge, gid[j], blk[j] + odd,
															   pair_field(first ? wb[j] : wa[j], odd), first ? st.h1b : st.h1a,
															   first ? st.h2b : st.h2a, t.tab, t.tab_off, t.kmax,
															   odd ? valid_hi[j] : 64, t.keys);
							if (odd) y.hi = out; else y.lo = out;
						}
						put(first, y);
					}
					{
						uint32_t ms = (mj >> (2 * second)) & 3u;
						const uint4 lo0 = cur[j].lo, hi0 = cur[j].hi;
						while (ms) {
							const uint32_t odd = (ms & 1u) ? 0u : 1u;
							ms &= ms - 1u;
							const uint32_t edge = second ? edge_b : edge_a;
							const uint4 out = mutate_block_h12(odd ? hi0 : lo0, edge, gid[j], blk[j] + odd,
															   pair_field(second ? wb[j] : wa[j], odd), second ? st.h1b : st.h1a,
															   second ? st.h2b : st.h2a, t.tab, t.tab_off, t.kmax,
															   odd ? valid_hi[j] : 64, t.keys);
							if (odd) cur[j].hi = out; else cur[j].lo = out;
						}
						put(second, cur[j]);
					}
				}
			}
		}
	}
	if (u.trace) { // (warp 0's clock after the CTA's last warp would need a barrier; the per-warp maximum does it)
		unsigned long long now;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
		if (lane == 0) atomicMax(u.trace + 4 * blockIdx.x + 2, now);
	}
}

// ---------------------------------------------------------------------------------------------
// K1, masks of the branches below the top nodes.  A substitution XORs a non-zero delta into a
// base, and which sites change (and by which delta) depends only on (seed, edge, gene, block) --
// never on the sequence.  So the effect of a branch on a block is an XOR mask that can be drawn
// before the parent's value exists.  The top of the tree -- long branches, where most blocks take the
// rare path, walked by few threads one node after the other -- is therefore split in two: all
// masks are drawn here, side by side over (top node, block pair), and the walk's steps over the
// top reduce to "load two masks, XOR, store" (kWalkMask).  Thread = one pair of blocks of one
// top node's two branches: 1 Philox call for the four decisions, then the flagged blocks' rare paths.
__global__ void __launch_bounds__(kThreads)
k1_top_masks(const __grid_constant__ DeviceTables t, const TopMaskTask *__restrict__ tasks, uint32_t slabs_per_task,
			 uint32_t pairs_per_node)
{
	const uint32_t task = blockIdx.x / slabs_per_task, slab = blockIdx.x - task * slabs_per_task;
	const uint32_t p = slab * kThreads + threadIdx.x;
	// the walk behind this launch may be placed once every CTA of this grid has got this far (it still
	// waits for the grid's completion before it reads a mask: griddepcontrol.wait in k1_walk_dense)
	asm volatile("griddepcontrol.launch_dependents;");
	if (p >= pairs_per_node) return;
	const TopMaskTask &tk = tasks[task];
	const uint32_t idx = 2u * p;
	const uint32_t slot = fastdiv(idx, t.div_blocks);
	const uint32_t blk = idx - slot * t.blocks; // the pair's even block
	const uint32_t gid = __ldg(t.dense_gid + slot);
	const int valid_hi = (blk + 2 == t.blocks) ? t.last_valid : 64;
	const uint32_t edge_a = tk.edge_a, edge_b = tk.edge_b;
	if (edge_a == kRootEdge) { // the dense genes' origin rows: what gene::gene(len, -1, id) draws (gene.cxx:13-25)
		if (tk.h1a) { // (test knob PGS_TEST_ORIGIN_DELAY: write them late)
			const long long t0 = clock64();
			while (clock64() - t0 < (long long)tk.h1a) {}
		}
		U32x8 r;
		r.lo = root_block(gid, blk, 64, t.keys);
		r.hi = root_block(gid, blk + 1u, valid_hi, t.keys);
		st_park256(reinterpret_cast<U32x8 *>(t.rows + tk.ma_off) + p, r);
		return;
	}
	const uint64_t h1a = tk.h1a, h1b = tk.h1b;
	const uint4 w = quad_words(edge_a, gid, blk, t.keys);
	const uint32_t wa = pick_pair_word(w, 0u, blk), wb = pick_pair_word(w, 1u, blk);
	uint32_t hits = pair_hits(wa, decision_cmp(h1a)) | (pair_hits(wb, decision_cmp(h1b)) << 2);
	const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
	U32x8 ma, mb;
	ma.lo = ma.hi = mb.lo = mb.hi = zero;
	while (hits) { // every lane works through its own flagged blocks
		const uint32_t b = (uint32_t)__ffs((int)hits) - 1u;
		hits &= hits - 1u;
		const uint32_t side = b >> 1, odd = b & 1u;
		const uint4 out = mutate_block_h12(zero, side ? edge_b : edge_a, gid, blk + odd, pair_field(side ? wb : wa, odd),
										   side ? h1b : h1a, side ? tk.h2b : tk.h2a, t.tab, t.tab_off, t.kmax,
										   odd ? valid_hi : 64, t.keys);
		if (side) {
			if (odd) mb.hi = out; else mb.lo = out;
		} else {
			if (odd) ma.hi = out; else ma.lo = out;
		}
	}
	// read again by the walk within microseconds: keep them in L2
	st_park256(reinterpret_cast<U32x8 *>(t.rows + tk.ma_off) + p, ma);
	st_park256(reinterpret_cast<U32x8 *>(t.rows + tk.mb_off) + p, mb);
}

void launch_top_masks(const DeviceTables &t, const TopMaskTask *tasks, int32_t n_tasks, cudaStream_t s)
{
	if (n_tasks == 0 || t.n_dense == 0) return;
	const uint32_t pairs = t.n_dense * t.blocks / 2;
	const uint32_t slabs = (pairs + kThreads - 1) / kThreads;
	k1_top_masks<<<(unsigned)((uint64_t)slabs * (uint32_t)n_tasks), kThreads, 0, s>>>(t, tasks, slabs, pairs);
}

int g_walk_variant = -1; // development knob (PGS_WALK_VARIANT): force a kernel, see choose_walk_variant

// Which walk kernel sweeps the subtrees.  0 = quads, 2 CTAs/SM (six stack levels); 1 = pairs;
// 2 = quads, 3 CTAs/SM (four stack levels); 3 = quads without cache hints.
// mean_hit = mean over the branches of P(a 64-site block changes).  Measured on config 4's tree:
// variant 2 wins while the rare path is rare (2.2 ms against 2.3 / 2.6 for variants 0 / 1 at
// mut-rate 0.01), variant 0 once it is not (3.6 against 3.85 ms at mut-rate 0.1).
int choose_walk_variant(uint32_t blocks, double mean_hit)
{
	if (blocks & 3u) return 1; // blocks = 2 mod 4: every pair makes its quad's call itself
	if (g_walk_variant >= 0 && g_walk_variant <= 3) return g_walk_variant;
	return mean_hit > 0.003 ? 0 : 2;
}

// column slabs (CTAs per unit)
uint32_t walk_slabs(int variant, uint32_t n_dense, uint32_t blocks)
{
	const uint32_t pairs = n_dense * blocks / 2, per_cta = kThreads * (variant == 1 ? 1 : 2);
	return (pairs + per_cta - 1) / per_cta;
}

// Shared memory of one SM that the resident CTAs' stacks may take: 227 KB less the static part
// (steps, control words) and the 1 KB the driver reserves per CTA.
static int walk_smem_hw_levels(int variant)
{
	const int ctas = (variant == 1 || variant == 2) ? 3 : 2, pairs = variant == 1 ? 1 : 2;
	const int per_cta = (227 * 1024) / ctas - 1024 - (int)(sizeof(WalkStep) * kWalkChunk * (kThreads / 32)) - 256;
	return per_cta / (pairs * 2 * 16 * kThreads);
}

int walk_smem_capacity(int variant)
{
	int levels = walk_smem_hw_levels(variant);
	if (const char *v = std::getenv("PGS_WALK_SMEM_LEVELS")) levels = std::min(levels, std::max(0, std::atoi(v))); // tuning knob
	return levels;
}

template <int MINB, int PAIRS, bool QUAD, bool HINTS>
static void launch_walk_variant(const DeviceTables &t, const WalkUnits &u, int32_t smem_levels, cudaStream_t s)
{
	const uint32_t pairs = t.n_dense * t.blocks / 2;
	const uint32_t per_cta = kThreads * PAIRS;
	const unsigned grid = (unsigned)((uint64_t)((pairs + per_cta - 1) / per_cta) * (uint32_t)u.n_units);
	const size_t dyn = (size_t)smem_levels * PAIRS * 2 * sizeof(uint4) * kThreads; // allowed by prepare_walk_dense
	// Programmatic dependent launch behind k1_top_masks (the launch before this one in the stream, also
	// when captured into the sweep's graph): CTAs are placed while that grid drains and run their
	// table-only set-up; the kernel's griddepcontrol.wait orders every read of a row after the masks.
	static const bool pdl = [] { const char *v = std::getenv("PGS_NO_PDL"); return !(v && std::atoi(v)); }(); // (A/B knob)
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(kThreads);
	cfg.dynamicSmemBytes = dyn;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = pdl ? 1 : 0;
	if (const char *path = std::getenv("PGS_WALK_TRACE")) { // development: per-CTA clocks of this launch into a file (plain launches only)
		WalkUnits ut = u;
		const size_t bytes = (size_t)grid * 4 * sizeof(unsigned long long);
		if (cudaMalloc(&ut.trace, bytes) != cudaSuccess) return;
		cudaMemsetAsync(ut.trace, 0, bytes, s);
		cudaLaunchKernelEx(&cfg, k1_walk_dense<MINB, PAIRS, QUAD, HINTS>, t, ut, pairs);
		std::vector<unsigned long long> host((size_t)grid * 4);
		cudaStreamSynchronize(s);
		cudaMemcpy(host.data(), ut.trace, bytes, cudaMemcpyDeviceToHost);
		cudaFree(ut.trace);
		if (FILE *f = std::fopen(path, "w")) {
			const unsigned n_slabs = grid / (unsigned)u.n_units;
			for (unsigned c = 0; c < grid; c++)
				std::fprintf(f, "%u %u %llu %llu %llu %llu\n", c / n_slabs, c % n_slabs, host[4 * c], host[4 * c + 1], host[4 * c + 2], host[4 * c + 3]);
			std::fclose(f);
		}
		return;
	}
	cudaLaunchKernelEx(&cfg, k1_walk_dense<MINB, PAIRS, QUAD, HINTS>, t, u, pairs);
}

template <int MINB, int PAIRS, bool QUAD, bool HINTS>
static cudaError_t allow_walk_smem(int variant)
{
	return cudaFuncSetAttribute(k1_walk_dense<MINB, PAIRS, QUAD, HINTS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
								(int)((size_t)walk_smem_hw_levels(variant) * PAIRS * 2 * sizeof(uint4) * kThreads));
}

// The attribute belongs to the function on the current device, not to an engine: every engine sets
// the same value (the most any plan can ask for), so engines that share a device cannot lower it
// under each other's feet.
cudaError_t prepare_walk_dense()
{
	cudaError_t ce = allow_walk_smem<2, 2, true, true>(0);
	if (ce == cudaSuccess) ce = allow_walk_smem<3, 1, false, true>(1);
	if (ce == cudaSuccess) ce = allow_walk_smem<3, 2, true, true>(2);
	if (ce == cudaSuccess) ce = allow_walk_smem<2, 2, true, false>(3);
	return ce;
}

void launch_walk_dense(const DeviceTables &t, const WalkUnits &u, int32_t smem_levels, int variant, cudaStream_t s)
{
	if (u.n_units == 0 || t.n_dense == 0) return;
	switch (variant) {
		case 0: launch_walk_variant<2, 2, true, true>(t, u, smem_levels, s); break;
		case 2: launch_walk_variant<3, 2, true, true>(t, u, smem_levels, s); break;
		case 3: launch_walk_variant<2, 2, true, false>(t, u, smem_levels, s); break;
		default: launch_walk_variant<3, 1, false, true>(t, u, smem_levels, s); break;
	}
}

// ---------------------------------------------------------------------------------------------
// K1, sparse region.  Accessory genes are present at few nodes; their rows are
// compacted per node and a 24-byte task per parent row names source and destination rows.
// Replaces gene::bulk_mutate over the per-node gene map, reference gene.h:26-40 called at
// img.cxx:285-286.

// PAIR: a thread takes two adjacent blocks (256-bit accesses, one Philox call for the four
// decisions); used whenever the number of blocks per gene is even.
template <bool PAIR>
__global__ void __launch_bounds__(kThreads)
k1_sweep_sparse(const __grid_constant__ DeviceTables t, const SparseTask *__restrict__ tasks, int64_t n_items)
{
	const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (idx >= n_items) return;
	const uint32_t per_row = PAIR ? t.blocks / 2 : t.blocks;
	const uint32_t task = (uint32_t)(idx / per_row);
	const uint32_t blk = (uint32_t)(idx - (int64_t)task * per_row) * (PAIR ? 2u : 1u);
	const SparseTask st = tasks[task];
	U32x8 x;
	if (PAIR) {
		x = ld_stream256(reinterpret_cast<const U32x8 *>(t.rows + (int64_t)st.src * t.blocks + blk));
	} else {
		x.lo = ld_stream(t.rows + (int64_t)st.src * t.blocks + blk);
		x.hi = x.lo;
	}
	const uint32_t last = PAIR ? blk + 1 : blk;
	const int valid_last = (last + 1 == t.blocks) ? t.last_valid : 64;
	const uint4 w = quad_words((uint32_t)st.edge_left, st.gid, blk, t.keys);
#pragma unroll
	for (uint32_t side = 0; side < 2; side++) {
		const uint32_t dst = side ? st.dst_right : st.dst_left;
		if (dst == kNoRow) continue;
		const uint32_t edge = (uint32_t)(side ? st.edge_right : st.edge_left);
		const uint32_t h16 = (uint32_t)(__ldg(t.h1 + edge) >> 48); // the fast path needs nothing else of the branch
		U32x8 y = x;
		const uint32_t D0 = pick_field(w, side, blk);
		if (D0 <= h16) {
			const EdgeDecision d = load_edge(t, (int32_t)edge);
			y.lo = mutate_block_slow(x.lo, edge, st.gid, blk, D0, d.tab, d.tab, d.kmax, PAIR ? 64 : valid_last, t.keys);
		}
		if (PAIR) {
			const uint32_t D1 = pick_field(w, side, blk + 1);
			if (D1 <= h16) {
				const EdgeDecision d = load_edge(t, (int32_t)edge);
				y.hi = mutate_block_slow(x.hi, edge, st.gid, blk + 1, D1, d.tab, d.tab, d.kmax, valid_last, t.keys);
			}
			st_stream256(reinterpret_cast<U32x8 *>(t.rows + (int64_t)dst * t.blocks + blk), y);
		} else {
			st_stream(t.rows + (int64_t)dst * t.blocks + blk, y.lo);
		}
	}
}

void launch_sweep_sparse(const DeviceTables &t, const SparseTask *tasks, int64_t n_tasks, cudaStream_t s)
{
	const bool pair = (t.blocks & 1u) == 0;
	const int64_t items = n_tasks * (pair ? t.blocks / 2 : t.blocks);
	if (items == 0) return;
	const int64_t grid = (items + kThreads - 1) / kThreads;
	if (pair)
		k1_sweep_sparse<true><<<(unsigned)grid, kThreads, 0, s>>>(t, tasks, items);
	else
		k1_sweep_sparse<false><<<(unsigned)grid, kThreads, 0, s>>>(t, tasks, items);
}

// ---------------------------------------------------------------------------------------------
// K1, fused walk of the sparse region (default for binary trees).  The level-synchronous kernel
// above needs one launch per tree level with a handful of CTAs each: launch-bound.  Here ONE launch
// does everything for the accessory genes: a thread owns one quad of blocks (64 B, 256 sites) of one
// gene, draws the gene's origin row itself (K0's work), and carries the quad depth-first down the
// gene's presence subtree -- the child it descends into stays in registers, a sibling that is
// visited later is parked in the thread's shared-memory stack (deeper levels: in the sibling's own
// row), and only genomes' rows and origin rows are written.  The lanes of a warp walk different
// genes with their own step lists (32-byte steps, read through L1); genes are sorted by the number
// of steps so that the lanes of a warp finish together.
// Replaces gene::bulk_mutate over the per-node gene map, reference gene.h:26-40 (img.cxx:285-286),
// and the gained gene's fresh sequence, img.cxx:302.

constexpr int kSpStackLevels = 4; // shared-memory stack levels per thread (64 B each)

__device__ __forceinline__ uint4 ld_cg(const uint4 *p)
{
	uint4 r;
	asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
	return r;
}

__global__ void __launch_bounds__(kThreads, 3)
k1_walk_sparse(const __grid_constant__ DeviceTables t, const SparseUnit *__restrict__ units,
			   const SparseStep *__restrict__ steps, const SparsePathEdge *__restrict__ paths, uint32_t n_items,
			   uint32_t quads_per_gene)
{
	extern __shared__ uint4 s_stack[]; // [level][block of the quad][thread]
	const uint32_t item = blockIdx.x * kThreads + threadIdx.x;
	if (item >= n_items) return;
	const uint32_t rank = item / quads_per_gene;
	const uint32_t blk0 = (item - rank * quads_per_gene) * 4u;
	const uint32_t nb = min(4u, t.blocks - blk0); // blocks of this quad that exist
	const uint4 u0 = __ldg(reinterpret_cast<const uint4 *>(units + rank)), u1 = __ldg(reinterpret_cast<const uint4 *>(units + rank) + 1);
	const uint32_t gid = u0.x, n_steps = u0.w;
	const uint32_t live = (1u << nb) - 1u;
	auto valid_of = [&](uint32_t i) { return (blk0 + i + 1u == t.blocks) ? t.last_valid : 64; };
	auto row_ptr = [&](uint32_t row) { return t.rows + (int64_t)row * t.blocks + blk0; };
	// rare path of block i on the branch `edge`; `from` is the value of the block above the branch
	auto rare = [&](uint32_t edge, uint32_t word, uint32_t i, const uint4 &from) {
		const uint64_t h1 = __ldg(t.tab8 + (size_t)edge * 8), h2 = __ldg(t.tab8 + (size_t)edge * 8 + 1);
		return mutate_block_h12(from, edge, gid, blk0 + i, pair_field(word, i & 1u), h1, h2, t.tab, t.tab_off, t.kmax, valid_of(i),
								t.keys);
	};

	uint4 cur[4];
	// origin row: what gene::gene(len, -1, id) draws (gene.cxx:13-25); the gene's origin unit keeps it
#pragma unroll
	for (uint32_t i = 0; i < 4; i++) {
		cur[i] = make_uint4(0u, 0u, 0u, 0u);
		if (i < nb) cur[i] = root_block(gid, blk0 + i, valid_of(i), t.keys);
	}
	if (u0.y != kNoRow) {
		uint4 *const dst = row_ptr(u0.y);
#pragma unroll
		for (uint32_t i = 0; i < 4; i++)
			if (i < nb) st_stream(dst + i, cur[i]);
	}
	// a unit below the origin: carry the value across the branches of its path (no memory involved)
	for (uint32_t k = 0; k < u1.y; k++) {
		const uint4 pe = __ldg(reinterpret_cast<const uint4 *>(paths + u1.x + k));
		const uint4 w = quad_words(pe.x, gid, blk0, t.keys);
		const bool second = (pe.z >> 31) != 0;
		const uint32_t hc = (pe.z << 16) | 0xFFFFu;
		const uint32_t w01 = second ? w.z : w.x, w23 = second ? w.w : w.y;
		const uint32_t h = (pair_hits(w01, hc) | (pair_hits(w23, hc) << 2)) & live;
		if (h) {
#pragma unroll
			for (uint32_t i = 0; i < 4; i++)
				if (h & (1u << i)) cur[i] = rare(pe.y, (i & 2u) ? w23 : w01, i, cur[i]);
		}
	}

	const uint4 *sp = reinterpret_cast<const uint4 *>(steps + u0.z);
	for (uint32_t k = 0; k < n_steps; k++, sp += 2) {
		// the descriptors two and three steps ahead are pulled into L1 while this one is worked on (a
		// prefetch holds no register; reading a little past the unit's last step is harmless)
		asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + 4));
		const uint4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
		const uint32_t f = s1.y;
		if (f & kSpLoadSmem) {
			const uint4 *const src = s_stack + ((s1.w >> 8) & 255u) * (4 * kThreads) + threadIdx.x;
#pragma unroll
			for (int i = 0; i < 4; i++) cur[i] = src[i * kThreads];
		} else if (f & kSpLoadRow) {
			const uint4 *const src = row_ptr(s1.z);
#pragma unroll
			for (uint32_t i = 0; i < 4; i++)
				if (i < nb) cur[i] = ld_cg(src + i); // written by this thread a few steps ago
		}
		// one Philox call: the eight decisions {a, b} x {four blocks}
		const uint4 w = quad_words(s0.z, gid, blk0, t.keys);
		const uint32_t hca = (s1.x << 16) | 0xFFFFu, hcb = (s1.x & 0xFFFF0000u) | 0xFFFFu;
		const uint32_t hits_a = (f & kSpHasA) ? ((pair_hits(w.x, hca) | (pair_hits(w.y, hca) << 2)) & live) : 0u;
		const uint32_t hits_b = (f & kSpHasB) ? ((pair_hits(w.z, hcb) | (pair_hits(w.w, hcb) << 2)) & live) : 0u;
		auto emit = [&](uint32_t side, const uint4 (&v)[4]) {
			const uint32_t dst = side ? s0.y : s0.x;
			if (dst != kNoRow) {
				uint4 *const d = row_ptr(dst);
#pragma unroll
				for (uint32_t i = 0; i < 4; i++)
					if (i < nb) st_stream(d + i, v[i]);
			}
			if (f & (side ? kSpParkB : kSpParkA)) {
				uint4 *const d = s_stack + (s1.w & 255u) * (4 * kThreads) + threadIdx.x;
#pragma unroll
				for (int i = 0; i < 4; i++) d[i * kThreads] = v[i];
			}
		};
		// The child that is NOT carried on goes first; the carried child's value replaces the running one.
		const uint32_t second = (f & kSpCarryA) ? 0u : 1u;
#pragma unroll 1
		for (uint32_t pass = 0; pass < 2; pass++) {
			const uint32_t side = pass ? second : (second ^ 1u);
			if (!(f & (side ? kSpHasB : kSpHasA))) continue;
			uint4 y[4] = {cur[0], cur[1], cur[2], cur[3]};
			const uint32_t h = side ? hits_b : hits_a;
			if (h) { // rare: some block of this child may change (static indices: the values stay in registers)
				const uint32_t edge = side ? s0.w : s0.z;
#pragma unroll
				for (uint32_t i = 0; i < 4; i++)
					if (h & (1u << i)) y[i] = rare(edge, side ? ((i & 2u) ? w.w : w.z) : ((i & 2u) ? w.y : w.x), i, cur[i]);
			}
			emit(side, y);
			if (pass) {
#pragma unroll
				for (int i = 0; i < 4; i++) cur[i] = y[i];
			}
		}
	}
}

int sparse_walk_smem_levels()
{
	return kSpStackLevels;
}

cudaError_t prepare_walk_sparse()
{
	return cudaFuncSetAttribute(k1_walk_sparse, cudaFuncAttributeMaxDynamicSharedMemorySize,
								kSpStackLevels * 4 * (int)sizeof(uint4) * kThreads);
}

void launch_walk_sparse(const DeviceTables &t, const SparseUnit *units, const SparseStep *steps, const SparsePathEdge *paths,
						int64_t n_units, cudaStream_t s)
{
	if (n_units == 0) return;
	const uint32_t quads = (t.blocks + 3u) / 4u;
	const int64_t items = n_units * quads;
	const size_t dyn = (size_t)kSpStackLevels * 4 * sizeof(uint4) * kThreads;
	k1_walk_sparse<<<(unsigned)((items + kThreads - 1) / kThreads), kThreads, dyn, s>>>(t, units, steps, paths, (uint32_t)items, quads);
}

} // namespace pgs
