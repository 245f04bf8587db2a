// kernels.cu -- K0 (origin rows), K1 (level-synchronous tree sweep), K2 (pairwise mismatch).
// Hand-written for sm_100a; integer / byte work bounded by HBM (K0, K1) or the integer pipes (K2),
// so no tensor cores: coalesced 16-byte accesses, several loads in flight per thread, grids that
// are whole multiples of what the 148 SMs hold.
#include "kernels.cuh"

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
// 1-D stream: item i of the parent region produces item i of both child regions.  Each thread
// takes kItems blocks: all loads are issued first (kItems x 16 B in flight per thread), each
// parent block is read once and both children are produced from it -- the read-once /
// write-once traffic of SURVEY.md §8(d): 48 bytes per (block, parent).
// Replaces the inner loop of img_model::sim_core_gene, reference img.cxx:215-231, whose
// per-edge body is gene::mutate, gene.cxx:27-61.

constexpr int kItems = 4;
constexpr int kTile = kThreads * kItems;

__global__ void __launch_bounds__(kThreads)
k1_sweep_dense(const __grid_constant__ DeviceTables t, const DenseTask *__restrict__ tasks,
			   uint32_t tiles_per_task, uint32_t items_per_task)
{
	const uint32_t task = blockIdx.x / tiles_per_task;
	const uint32_t tile = blockIdx.x - task * tiles_per_task;
	const DenseTask dt = tasks[task];

	const uint4 *__restrict__ src = t.rows + __ldg(t.row_base + dt.parent) * t.blocks;
	const uint32_t first = tile * kTile + threadIdx.x;

	uint4 x[kItems];
#pragma unroll
	for (int j = 0; j < kItems; j++) {
		const uint32_t idx = first + j * kThreads;
		if (idx < items_per_task) x[j] = ld_stream(src + idx);
	}

	uint32_t gid[kItems], blk[kItems];
#pragma unroll
	for (int j = 0; j < kItems; j++) {
		const uint32_t idx = first + j * kThreads;
		const uint32_t slot = fastdiv(idx, t.div_blocks);
		blk[j] = idx - slot * t.blocks;
		gid[j] = (idx < items_per_task) ? __ldg(t.dense_gid + slot) : 0u;
	}

	{
		const EdgeDecision d = load_edge(t, dt.left);
		uint4 *__restrict__ dst = t.rows + __ldg(t.row_base + dt.left) * t.blocks;
#pragma unroll
		for (int j = 0; j < kItems; j++) {
			const uint32_t idx = first + j * kThreads;
			if (idx < items_per_task) {
				const int valid = (blk[j] + 1 == t.blocks) ? t.last_valid : 64;
				st_stream(dst + idx, mutate_block(x[j], (uint32_t)dt.left, gid[j], blk[j], d, valid, t.keys));
			}
		}
	}
	if (dt.right >= 0) {
		const EdgeDecision d = load_edge(t, dt.right);
		uint4 *__restrict__ dst = t.rows + __ldg(t.row_base + dt.right) * t.blocks;
#pragma unroll
		for (int j = 0; j < kItems; j++) {
			const uint32_t idx = first + j * kThreads;
			if (idx < items_per_task) {
				const int valid = (blk[j] + 1 == t.blocks) ? t.last_valid : 64;
				st_stream(dst + idx, mutate_block(x[j], (uint32_t)dt.right, gid[j], blk[j], d, valid, t.keys));
			}
		}
	}
}

void launch_sweep_dense(const DeviceTables &t, const DenseTask *tasks, int32_t n_tasks, cudaStream_t s)
{
	if (n_tasks == 0 || t.n_dense == 0) return;
	const uint32_t items = t.n_dense * t.blocks;
	const uint32_t tiles = (items + kTile - 1) / kTile;
	// one launch covers at most 2^31-1 CTAs; chunk the task list beyond that
	const int64_t max_tasks = (int64_t)0x7FFFFFFF / tiles;
	for (int64_t done = 0; done < n_tasks; done += max_tasks) {
		const int64_t now = (n_tasks - done < max_tasks) ? (n_tasks - done) : max_tasks;
		k1_sweep_dense<<<(unsigned)(now * tiles), kThreads, 0, s>>>(t, tasks + done, tiles, items);
	}
}

// ---------------------------------------------------------------------------------------------
// K1, sparse region.  Accessory genes are present at few nodes; their rows are
// compacted per node and a 24-byte task per parent row names source and destination rows.
// Replaces gene::bulk_mutate over the per-node gene map, reference gene.h:26-40 called at
// img.cxx:285-286.

__global__ void __launch_bounds__(kThreads)
k1_sweep_sparse(const __grid_constant__ DeviceTables t, const SparseTask *__restrict__ tasks, int64_t n_items)
{
	const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (idx >= n_items) return;
	const uint32_t task = (uint32_t)(idx / t.blocks);
	const uint32_t blk = (uint32_t)(idx - (int64_t)task * t.blocks);
	const SparseTask st = tasks[task];
	const uint4 x = ld_stream(t.rows + (int64_t)st.src * t.blocks + blk);
	const int valid = (blk + 1 == t.blocks) ? t.last_valid : 64;
	if (st.dst_left != kNoRow) {
		const EdgeDecision d = load_edge(t, st.edge_left);
		st_stream(t.rows + (int64_t)st.dst_left * t.blocks + blk,
				  mutate_block(x, (uint32_t)st.edge_left, st.gid, blk, d, valid, t.keys));
	}
	if (st.dst_right != kNoRow) {
		const EdgeDecision d = load_edge(t, st.edge_right);
		st_stream(t.rows + (int64_t)st.dst_right * t.blocks + blk,
				  mutate_block(x, (uint32_t)st.edge_right, st.gid, blk, d, valid, t.keys));
	}
}

void launch_sweep_sparse(const DeviceTables &t, const SparseTask *tasks, int64_t n_tasks, cudaStream_t s)
{
	const int64_t items = n_tasks * t.blocks;
	if (items == 0) return;
	const int64_t grid = (items + kThreads - 1) / kThreads;
	k1_sweep_sparse<<<(unsigned)grid, kThreads, 0, s>>>(t, tasks, items);
}

// ---------------------------------------------------------------------------------------------
// K2: observed pairwise mismatch counts (popcount of folded XOR) over the dense regions of the
// leaves.  A CTA owns a 64 x 64 tile of genome pairs and a slice of the columns; 32-bit words
// (16 sites) are staged through shared memory in chunks, each thread accumulates a 4 x 4
// sub-tile in registers.  Bound by the integer pipes (n^2 S / 2 site pairs), not by HBM.

constexpr int kMmTile = 64;   // genomes per tile side
constexpr int kMmChunk = 32;  // 32-bit words per staged chunk

__device__ __forceinline__ uint32_t mismatch16(uint32_t a, uint32_t b)
{
	const uint32_t d = a ^ b;
	return __popc((d | (d >> 1)) & 0x55555555u);
}

__global__ void __launch_bounds__(256)
k2_mismatch(const uint32_t *__restrict__ words, const int64_t *__restrict__ leaf_base_words, int32_t n,
			int64_t words_per_leaf, int64_t words_per_split, uint32_t *__restrict__ out)
{
	// tile coordinates in the upper triangle: blockIdx.x enumerates (ti <= tj)
	const int tiles = (n + kMmTile - 1) / kMmTile;
	int ti = 0, rem = blockIdx.x;
	while (rem >= tiles - ti) {
		rem -= tiles - ti;
		ti++;
	}
	const int tj = ti + rem;

	__shared__ uint32_t sa[kMmTile][kMmChunk + 1];
	__shared__ uint32_t sb[kMmTile][kMmChunk + 1];

	const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4; // 16 x 16 threads, 4 x 4 pairs each
	uint32_t acc[4][4];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) acc[i][j] = 0;

	const int64_t w_begin = (int64_t)blockIdx.y * words_per_split;
	int64_t w_end = w_begin + words_per_split;
	if (w_end > words_per_leaf) w_end = words_per_leaf;

	for (int64_t w0 = w_begin; w0 < w_end; w0 += kMmChunk) {
		// stage 64 rows x 32 words of each side: 2048 words per side, 8 per thread, coalesced in w
		for (int e = threadIdx.x; e < kMmTile * kMmChunk; e += 256) {
			const int row = e / kMmChunk, col = e % kMmChunk;
			const int64_t w = w0 + col;
			const int ga = ti * kMmTile + row, gb = tj * kMmTile + row;
			sa[row][col] = (ga < n && w < w_end) ? __ldg(words + leaf_base_words[ga] + w) : 0u;
			sb[row][col] = (gb < n && w < w_end) ? __ldg(words + leaf_base_words[gb] + w) : 0u;
		}
		__syncthreads();
#pragma unroll 4
		for (int k = 0; k < kMmChunk; k++) {
			uint32_t a[4], b[4];
#pragma unroll
			for (int i = 0; i < 4; i++) a[i] = sa[ty * 4 + i][k];
#pragma unroll
			for (int j = 0; j < 4; j++) b[j] = sb[tx * 4 + j][k];
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) acc[i][j] += mismatch16(a[i], b[j]);
		}
		__syncthreads();
	}

#pragma unroll
	for (int i = 0; i < 4; i++) {
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int gi = ti * kMmTile + ty * 4 + i, gj = tj * kMmTile + tx * 4 + j;
			if (gi < n && gj < n && gi < gj && acc[i][j]) {
				atomicAdd(out + (int64_t)gi * n + gj, acc[i][j]);
			}
		}
	}
}

__global__ void k2_mirror(uint32_t *__restrict__ out, int32_t n)
{
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (int64_t)n * n) return;
	const int32_t i = (int32_t)(idx / n), j = (int32_t)(idx % n);
	if (i > j) out[idx] = out[(int64_t)j * n + i];
}

void launch_mismatch(const uint4 *rows, const int64_t *leaf_base_words, int32_t n, int64_t words_per_leaf,
					 uint32_t *out, cudaStream_t s)
{
	cudaMemsetAsync(out, 0, sizeof(uint32_t) * (size_t)n * (size_t)n, s);
	if (n < 2 || words_per_leaf == 0) return;
	const int tiles = (n + kMmTile - 1) / kMmTile;
	const int64_t tri = (int64_t)tiles * (tiles + 1) / 2;
	// split the columns so that small n still fills the 148 SMs (4 CTAs each)
	int64_t splits = (148 * 4 + tri - 1) / tri;
	const int64_t chunks = (words_per_leaf + kMmChunk - 1) / kMmChunk;
	if (splits > chunks) splits = chunks;
	if (splits > 65535) splits = 65535;
	if (splits < 1) splits = 1;
	const int64_t words_per_split = ((chunks + splits - 1) / splits) * kMmChunk;
	const int64_t used = (words_per_leaf + words_per_split - 1) / words_per_split;
	dim3 grid((unsigned)tri, (unsigned)used);
	k2_mismatch<<<grid, 256, 0, s>>>(reinterpret_cast<const uint32_t *>(rows), leaf_base_words, n,
									  words_per_leaf, words_per_split, out);
	const int64_t cells = (int64_t)n * n;
	k2_mirror<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(out, n);
}

} // namespace pgs
