// mismatch.cu -- K2: observed pairwise mismatch counts between the genomes over the dense
// (all-genome, root-origin) genes, and the path's only collective: the sum of the engines'
// partial counts (NCCL all-reduce over NVLink).
//
// The reference has no counterpart: its distance.mat is the EXPECTED distance (img.cxx:121-165,
// printed at pangenomesim.cxx:214-234); north_star part 3 asks for the observed matrix it is
// checked against.
//
// n^2 S / 2 site pairs: bound by the integer pipe, not by HBM.  Three pieces:
//   k2_split_planes   genomes' 2-bit rows -> two bit planes (low / high bit of every base), 32
//                     sites per word, interleaved (lo, hi) per 32-site group.  Done once per column
//                     chunk, so the pair kernel's staging is a plain asynchronous copy.
//   k2_mismatch_tiles CTA = 64 x 64 tile of genome pairs (upper triangle only), 4 x 4 pairs per
//                     thread; chunks of 576 sites staged with cp.async, double buffered; per pair
//                     and 32 sites  mask = (lo_i ^ lo_j) | (hi_i ^ hi_j), three masks through a
//                     carry-save adder, count += popc(sum) + 2 popc(carry).
//   output            upper-triangular 64 x 64 tiles, packed: n = 50 000 needs 5 GB instead of the
//                     10 GB square, and a partial result is ONE contiguous buffer for the all-reduce.
#include "engine.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <thread>
#include <vector>

namespace pgs {

constexpr int kMmTile = 64;     // genomes per tile side
constexpr int kMmGroups = 18;   // 32-site groups per staged chunk (a multiple of three: see the adder below)
constexpr int kMmRowWords = 44; // shared-memory row stride in words: 18 groups x (lo, hi), one group of skew, padding.
                                // Rows 8..15 of every 16 are skewed by one group (8 bytes): the 16 rows a warp reads 64 bits
                                // of then start in banks 0, 12, 24, 4, ... (rows 0-7) and 2, 14, 26, ... (rows 8-15) -- all 32
                                // banks once, no conflict (unskewed, rows r and r + 8 collide: 53 % of the LSU's wavefronts).
__device__ __forceinline__ int mm_skew(int row)
{
	return (row >> 3) & 1; // in groups
}

// even bits of x -> low 16 bits
__device__ __forceinline__ uint32_t even_bits(uint32_t x)
{
	x &= 0x55555555u;
	x = (x | (x >> 1)) & 0x33333333u;
	x = (x | (x >> 2)) & 0x0F0F0F0Fu;
	x = (x | (x >> 4)) & 0x00FF00FFu;
	x = (x | (x >> 8)) & 0x0000FFFFu;
	return x;
}

// One thread = one 64-site block of one genome: 16 bytes in, two (lo, hi) groups out.
// planes[g * chunk_groups + (group - first_group)]; groups beyond the genome's last one stay zero
// (the buffer is cleared first), and so do padding sites (they are zero in every row).
__global__ void __launch_bounds__(256)
k2_split_planes(const uint4 *__restrict__ rows, const int64_t *__restrict__ leaf_base_words, int32_t n, int64_t first_block,
				int64_t n_blocks, int64_t chunk_groups, uint2 *__restrict__ planes)
{
	const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
	if (idx >= (int64_t)n * n_blocks) return;
	const int32_t g = (int32_t)(idx / n_blocks);
	const int64_t b = idx - (int64_t)g * n_blocks;
	const uint4 v = __ldg(rows + (leaf_base_words[g] >> 2) + first_block + b);
	uint4 out;
	out.x = even_bits(v.x) | (even_bits(v.y) << 16);
	out.y = even_bits(v.x >> 1) | (even_bits(v.y >> 1) << 16);
	out.z = even_bits(v.z) | (even_bits(v.w) << 16);
	out.w = even_bits(v.z >> 1) | (even_bits(v.w >> 1) << 16);
	*reinterpret_cast<uint4 *>(planes + (int64_t)g * chunk_groups + 2 * b) = out;
}

__device__ __forceinline__ void cp_async_commit()
{
	asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void cp_async_wait()
{
	asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// number of tiles in the upper triangle before tile row ti
__host__ __device__ inline int64_t tri_row_start(int64_t tiles, int64_t ti)
{
	return ti * tiles - ti * (ti - 1) / 2;
}

// Where the time goes (ncu, B200): POPC issues on the XU pipe at 16 lanes per clock per SM, a
// quarter of the ALU pipe's 64 -- one POPC per mask made XU the bound (88 % busy, ALU 47 %).  So
// three masks go through a carry-save adder first (sum = a ^ b ^ c, carry = majority: two LOP3) and
// only sum and carry are counted: per three masks 8 ALU instructions (16 pipe cycles per warp) and
// 2 POPC (16 XU cycles) -- the two pipes are level; a deeper adder tree would only load ALU more.
// `one` is 1: the adds are written as multiply-adds so that they issue on the FMA pipe (IMAD).
template <int MINB, int UNROLL>
__global__ void __launch_bounds__(256, MINB)
k2_mismatch_tiles(const uint2 *__restrict__ planes, int32_t n, int64_t chunk_groups, int64_t groups_per_split,
				  uint32_t *__restrict__ tiles_out, uint32_t one)
{
	// tile coordinates in the upper triangle: blockIdx.x enumerates (ti <= tj) row by row
	const int tiles = (n + kMmTile - 1) / kMmTile;
	int ti = 0, rem = blockIdx.x;
	while (rem >= tiles - ti) {
		rem -= tiles - ti;
		ti++;
	}
	const int tj = ti + rem;

	__shared__ __align__(16) uint32_t sm[2][2][kMmTile][kMmRowWords]; // [stage][side a/b][row][group: lo, hi]

	const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4; // pairs (ty + 16 i, tx + 16 j)
	uint32_t acc[4][4];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) acc[i][j] = 0;

	const int64_t g_begin = (int64_t)blockIdx.y * groups_per_split;
	const int64_t g_end = min(g_begin + groups_per_split, chunk_groups);
	const int64_t n_stages = (g_end - g_begin + kMmGroups - 1) / kMmGroups;

	// stage s: 2 sides x 64 rows x 18 groups of 8 bytes (lo, hi).  Two threads per row, nine
	// consecutive groups each: one global pointer and one shared-memory address per thread, the
	// pieces at immediate offsets (the 32-byte sectors a piece leaves behind in L1 serve the next ones).
	const int c_side = threadIdx.x >> 7, c_row = (threadIdx.x >> 1) & 63, c_half = threadIdx.x & 1;
	const int c_genome = (c_side ? tj : ti) * kMmTile + c_row;
	const uint2 *const c_src = planes + (int64_t)min(c_genome, n - 1) * chunk_groups + g_begin + 9 * c_half;
	const uint32_t c_dst = (uint32_t)__cvta_generic_to_shared(&sm[0][c_side][c_row][2 * (9 * c_half + mm_skew(c_row))]);
	constexpr uint32_t kStageBytes = 2 * kMmTile * kMmRowWords * 4;
	auto issue = [&](int64_t s, int buf) {
		const int64_t left = (g_end - g_begin) - s * kMmGroups - 9 * c_half; // groups of this thread's run that exist
		const int n_valid = c_genome < n ? (int)min((int64_t)9, max((int64_t)0, left)) : 0;
		const uint2 *src = c_src + s * kMmGroups;
		const uint32_t dst = c_dst + (uint32_t)buf * kStageBytes;
#pragma unroll
		for (int q = 0; q < 9; q++) {
			const int bytes = q < n_valid ? 8 : 0; // 0: zero-filled, nothing is read
			asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + 8u * q), "l"(n_valid ? src + q : c_src), "r"(bytes) : "memory");
		}
		cp_async_commit();
	};

	if (n_stages > 0) issue(0, 0);
	for (int64_t s = 0; s < n_stages; s++) {
		const int buf = (int)(s & 1);
		if (s + 1 < n_stages) {
			issue(s + 1, buf ^ 1);
			cp_async_wait<1>();
		} else {
			cp_async_wait<0>();
		}
		__syncthreads();
		const uint32_t two = one + one;
		// this thread's rows a = ty + 16 i and b = tx + 16 j: the skew of a row depends on bit 3 only, the same for every i / j
		const uint32_t *const arow = &sm[buf][0][ty][2 * mm_skew(ty)], *const brow = &sm[buf][1][tx][2 * mm_skew(tx)];
#pragma unroll UNROLL
		for (int k = 0; k < kMmGroups; k += 3) {
			uint2 a[4][3], b[4][3]; // three groups each: (lo, hi)
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int g = 0; g < 3; g++) {
					a[i][g] = *reinterpret_cast<const uint2 *>(arow + 16 * i * kMmRowWords + 2 * (k + g));
					b[i][g] = *reinterpret_cast<const uint2 *>(brow + 16 * i * kMmRowWords + 2 * (k + g));
				}
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const uint32_t m0 = (a[i][0].x ^ b[j][0].x) | (a[i][0].y ^ b[j][0].y);
					const uint32_t m1 = (a[i][1].x ^ b[j][1].x) | (a[i][1].y ^ b[j][1].y);
					const uint32_t m2 = (a[i][2].x ^ b[j][2].x) | (a[i][2].y ^ b[j][2].y);
					const uint32_t sum = m0 ^ m1 ^ m2, carry = (m0 & m1) | (m2 & (m0 | m1));
					acc[i][j] = __popc(sum) * one + acc[i][j];
					acc[i][j] = __popc(carry) * two + acc[i][j];
				}
		}
		__syncthreads(); // everyone is done with `buf` before the copy of stage s + 2 lands in it
	}

	uint32_t *const tile = tiles_out + (tri_row_start(tiles, ti) + (tj - ti)) * (int64_t)(kMmTile * kMmTile);
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int r = ty + 16 * i, c = tx + 16 * j;
			const int gi = ti * kMmTile + r, gj = tj * kMmTile + c;
			if (gi < n && gj < n && gi < gj && acc[i][j]) atomicAdd(tile + r * kMmTile + c, acc[i][j]);
		}
}

// packed upper-triangular tiles -> full symmetric n x n (zero diagonal)
__global__ void k2_tiles_to_square(const uint32_t *__restrict__ tiles_in, int32_t n, uint32_t *__restrict__ out)
{
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (int64_t)n * n) return;
	int32_t i = (int32_t)(idx / n), j = (int32_t)(idx % n);
	if (i == j) {
		out[idx] = 0;
		return;
	}
	if (i > j) {
		const int32_t t = i;
		i = j;
		j = t;
	}
	const int64_t tiles = (n + kMmTile - 1) / kMmTile;
	const int64_t ti = i / kMmTile, tj = j / kMmTile;
	out[idx] = tiles_in[(tri_row_start(tiles, ti) + (tj - ti)) * (kMmTile * kMmTile) + (i % kMmTile) * kMmTile + (j % kMmTile)];
}

int64_t mismatch_tile_words(int32_t n)
{
	const int64_t tiles = (n + kMmTile - 1) / kMmTile;
	return tiles * (tiles + 1) / 2 * (kMmTile * kMmTile);
}

// Which K2: 0 = popcount kernel (this file); tensor cores (mismatch_tc.cu): 1 = CTA pairs with E4M3
// operands, 2 = CTA pairs with INT8 operands, 3 / 4 = the single-CTA kernel (E4M3 / INT8), 5 = CTA pairs with
// E2M1 operands and unit block scales.  The
// counts are the same integers either way.
// PGS_K2_IMPL=popc|tensor|tensor-i8|tensor-fp4|tensor-1cta|tensor-1cta-i8 overrides the engine's flags.
int mismatch_impl(const Engine &e)
{
	if (const char *v = std::getenv("PGS_K2_IMPL")) {
		if (!std::strcmp(v, "popc")) return 0;
		if (!std::strcmp(v, "tensor")) return 1;
		if (!std::strcmp(v, "tensor-i8")) return 2;
		if (!std::strcmp(v, "tensor-1cta")) return 3;
		if (!std::strcmp(v, "tensor-1cta-i8")) return 4;
		if (!std::strcmp(v, "tensor-fp4")) return 5;
	}
	if (e.cfg.flags & PGS_FLAG_K2_POPC) return 0;
	return 5; // PGS_FLAG_K2_TENSOR or neither flag: the fastest of the exact kernels
}

// Partial mismatch counts of this engine's dense genes into e.d_tiles (packed upper-triangular
// tiles), asynchronous on the engine's stream.
int mismatch_tiles(Engine &e, int64_t *sites_out)
{
	const int32_t n = e.n_genomes;
	const int64_t D = (int64_t)e.dense_gid.size();
	const int64_t words = mismatch_tile_words(n);
	int rc;
	if (e.d_tiles.bytes < (size_t)words * 4 && (rc = e.alloc(e.d_tiles, (size_t)words * 4, "mismatch tiles"))) return rc;
	cudaError_t ce = cudaMemsetAsync(e.d_tiles.p, 0, (size_t)words * 4, e.stream);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "mismatch tiles");
	if (sites_out) *sites_out = D * e.cfg.gene_length;
	const int64_t blocks_total = D * e.blocks; // 64-site blocks per genome
	if (n < 2 || blocks_total == 0) return PGS_OK;
	if (const int impl = mismatch_impl(e)) return mismatch_tiles_tensor(e, impl);

	// column chunks: the bit planes of a chunk take n * chunk_groups * 8 bytes (default: at most 2 GB)
	int64_t budget = (int64_t)2048 << 20;
	if (const char *v = std::getenv("PGS_K2_PLANES_MB")) budget = std::max<int64_t>(1, std::atol(v)) << 20; // tuning knob
	static_assert(kMmGroups % 6 == 0 && 2 * kMmTile * kMmGroups == 9 * 256, "a stage is whole 64-site blocks and whole adder steps; nine pieces per thread");
	int64_t chunk_blocks = std::max<int64_t>(kMmGroups / 2, budget / ((int64_t)n * 16));
	chunk_blocks = std::min(chunk_blocks, blocks_total);
	chunk_blocks = (chunk_blocks + kMmGroups / 2 - 1) / (kMmGroups / 2) * (kMmGroups / 2); // whole stages
	const int64_t chunk_groups = 2 * chunk_blocks;
	const size_t plane_bytes = (size_t)n * (size_t)chunk_groups * 8;
	if (e.d_planes.bytes < plane_bytes && (rc = e.alloc(e.d_planes, plane_bytes, "mismatch bit planes"))) return rc;

	const int tiles = (n + kMmTile - 1) / kMmTile;
	const int64_t tri = (int64_t)tiles * (tiles + 1) / 2;
	for (int64_t b0 = 0; b0 < blocks_total; b0 += chunk_blocks) {
		const int64_t nb = std::min(chunk_blocks, blocks_total - b0);
		if (nb < chunk_blocks) { // last chunk: the tail of the buffer must read as zero
			ce = cudaMemsetAsync(e.d_planes.p, 0, plane_bytes, e.stream);
			if (ce != cudaSuccess) return e.cuda_fail(ce, "mismatch bit planes");
		}
		const int64_t items = (int64_t)n * nb;
		k2_split_planes<<<(unsigned)((items + 255) / 256), 256, 0, e.stream>>>(e.tables.rows, e.d_leaf_base_words.as<int64_t>(), n, b0, nb,
																			   chunk_groups, e.d_planes.as<uint2>());
		// split the columns so that small n still fills the 148 SMs (2 CTAs each)
		const int64_t stages = (2 * nb + kMmGroups - 1) / kMmGroups;
		int64_t splits = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(stages, 65535), (148 * 4 + tri - 1) / tri));
		const int64_t groups_per_split = (stages + splits - 1) / splits * kMmGroups;
		splits = (2 * nb + groups_per_split - 1) / groups_per_split;
		dim3 grid((unsigned)tri, (unsigned)splits);
		static const int variant = std::getenv("PGS_K2_VARIANT") ? std::atoi(std::getenv("PGS_K2_VARIANT")) : 0; // development knob
		if (variant == 1)
			k2_mismatch_tiles<3, 1><<<grid, 256, 0, e.stream>>>(e.d_planes.as<uint2>(), n, chunk_groups, groups_per_split, e.d_tiles.as<uint32_t>(), 1u);
		else if (variant == 2)
			k2_mismatch_tiles<2, 1><<<grid, 256, 0, e.stream>>>(e.d_planes.as<uint2>(), n, chunk_groups, groups_per_split, e.d_tiles.as<uint32_t>(), 1u);
		else
			k2_mismatch_tiles<2, 2><<<grid, 256, 0, e.stream>>>(e.d_planes.as<uint2>(), n, chunk_groups, groups_per_split, e.d_tiles.as<uint32_t>(), 1u);
	}
	ce = cudaGetLastError();
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch launch");
	return PGS_OK;
}

// e.d_tiles -> full symmetric matrix in caller-provided device memory
int mismatch_square_device(Engine &e, uint32_t *out_device)
{
	const int32_t n = e.n_genomes;
	const int64_t cells = (int64_t)n * n;
	if (cells == 0) return PGS_OK;
	k2_tiles_to_square<<<(unsigned)((cells + 255) / 256), 256, 0, e.stream>>>(e.d_tiles.as<uint32_t>(), n, out_device);
	const cudaError_t ce = cudaGetLastError();
	return ce == cudaSuccess ? PGS_OK : e.cuda_fail(ce, "pgs_mismatch expand");
}

// e.d_tiles -> host n x n (symmetric, zero diagonal), through the engine's pinned staging buffers in
// pieces: the device never holds the square.
int mismatch_tiles_to_host(Engine &e, uint32_t *out_host)
{
	const int32_t n = e.n_genomes;
	const int64_t tiles = (n + kMmTile - 1) / kMmTile;
	const int64_t tile_words = kMmTile * kMmTile;
	std::memset(out_host, 0, sizeof(uint32_t) * (size_t)n * (size_t)n);
	if (n < 2) return PGS_OK;
	// pinned staging: two buffers of up to 64 MB, each holds whole tiles
	const int64_t per_buf = std::max<int64_t>(1, ((int64_t)64 << 20) / (tile_words * 4));
	const size_t buf_bytes = (size_t)per_buf * tile_words * 4;
	if (e.mm_stage_bytes < buf_bytes) {
		for (auto &p : e.h_mm_stage) {
			if (p) cudaFreeHost(p);
			p = nullptr;
		}
		e.mm_stage_bytes = 0;
		for (auto &p : e.h_mm_stage) {
			const cudaError_t ce = cudaHostAlloc((void **)&p, buf_bytes, cudaHostAllocDefault);
			if (ce != cudaSuccess) {
				p = nullptr;
				return e.cuda_fail(ce, "pinned mismatch staging");
			}
		}
		e.mm_stage_bytes = buf_bytes;
	}
	const int64_t total = tiles * (tiles + 1) / 2;
	cudaEvent_t done[2] = {e.ev[3], e.ev[4]};
	auto scatter = [&](const uint32_t *src, int64_t first, int64_t count) {
		// tile index -> (ti, tj): walk the rows of the triangle
		int64_t ti = 0, start = 0;
		while (start + (tiles - ti) <= first) {
			start += tiles - ti;
			ti++;
		}
		int64_t tj = ti + (first - start);
		for (int64_t k = 0; k < count; k++) {
			const uint32_t *t = src + k * tile_words;
			const int64_t r0 = ti * kMmTile, c0 = tj * kMmTile;
			const int64_t rows = std::min<int64_t>(kMmTile, n - r0), cols = std::min<int64_t>(kMmTile, n - c0);
			for (int64_t r = 0; r < rows; r++)
				for (int64_t c = (ti == tj ? r + 1 : 0); c < cols; c++) {
					const uint32_t v = t[r * kMmTile + c];
					out_host[(r0 + r) * n + (c0 + c)] = v;
					out_host[(c0 + c) * n + (r0 + r)] = v;
				}
			if (++tj == tiles) {
				ti++;
				tj = ti;
			}
		}
	};
	int64_t pending_first[2] = {0, 0}, pending_count[2] = {0, 0};
	int slot = 0;
	for (int64_t first = 0; first < total; first += per_buf, slot ^= 1) {
		const int64_t count = std::min(per_buf, total - first);
		cudaError_t ce = cudaMemcpyAsync(e.h_mm_stage[slot], e.d_tiles.as<uint32_t>() + first * tile_words, (size_t)count * tile_words * 4,
										 cudaMemcpyDeviceToHost, e.stream);
		if (ce == cudaSuccess) ce = cudaEventRecord(done[slot], e.stream);
		if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch copy");
		pending_first[slot] = first;
		pending_count[slot] = count;
		if (pending_count[slot ^ 1]) { // unpack the previous piece while this one is copied
			ce = cudaEventSynchronize(done[slot ^ 1]);
			if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch copy");
			scatter(reinterpret_cast<const uint32_t *>(e.h_mm_stage[slot ^ 1]), pending_first[slot ^ 1], pending_count[slot ^ 1]);
			pending_count[slot ^ 1] = 0;
		}
	}
	for (int s = 0; s < 2; s++)
		if (pending_count[s]) {
			const cudaError_t ce = cudaEventSynchronize(done[s]);
			if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch copy");
			scatter(reinterpret_cast<const uint32_t *>(e.h_mm_stage[s]), pending_first[s], pending_count[s]);
		}
	return PGS_OK;
}

// ---------------------------------------------------------------------------------------------
// NCCL, loaded at run time (libnccl.so.2): the library has no hard dependency on it, and a host
// process that already carries an NCCL (PyTorch) shares that copy.

namespace {

struct Nccl {
	typedef struct ncclComm *comm_t;
	struct unique_id {
		char internal[128];
	};
	int (*GetUniqueId)(unique_id *) = nullptr;
	int (*CommInitRank)(comm_t *, int, unique_id, int) = nullptr;
	int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
	int (*CommDestroy)(comm_t) = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
	bool ok = false;
	std::string why;
};
constexpr int kNcclUint32 = 3, kNcclSum = 0; // ncclUint32, ncclSum (nccl.h)

Nccl &nccl()
{
	static Nccl lib;
	static std::once_flag once;
	std::call_once(once, [] {
		void *h = nullptr;
		for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
			h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (h) break;
		}
		if (!h) {
			lib.why = std::string("NCCL is not available (") + dlerror() + ")";
			return;
		}
		auto sym = [&](const char *name) { return dlsym(h, name); };
		lib.GetUniqueId = reinterpret_cast<decltype(lib.GetUniqueId)>(sym("ncclGetUniqueId"));
		lib.CommInitRank = reinterpret_cast<decltype(lib.CommInitRank)>(sym("ncclCommInitRank"));
		lib.CommInitAll = reinterpret_cast<decltype(lib.CommInitAll)>(sym("ncclCommInitAll"));
		lib.CommDestroy = reinterpret_cast<decltype(lib.CommDestroy)>(sym("ncclCommDestroy"));
		lib.AllReduce = reinterpret_cast<decltype(lib.AllReduce)>(sym("ncclAllReduce"));
		lib.GroupStart = reinterpret_cast<decltype(lib.GroupStart)>(sym("ncclGroupStart"));
		lib.GroupEnd = reinterpret_cast<decltype(lib.GroupEnd)>(sym("ncclGroupEnd"));
		lib.GetErrorString = reinterpret_cast<decltype(lib.GetErrorString)>(sym("ncclGetErrorString"));
		lib.ok = lib.GetUniqueId && lib.CommInitRank && lib.CommInitAll && lib.CommDestroy && lib.AllReduce && lib.GroupStart &&
				 lib.GroupEnd && lib.GetErrorString;
		if (!lib.ok) lib.why = "libnccl.so.2 lacks a required symbol";
	});
	return lib;
}

} // namespace

void Engine::drop_comm()
{
	if (comm && nccl().ok) nccl().CommDestroy(static_cast<Nccl::comm_t>(comm));
	comm = nullptr;
	comm_world = 1;
}

int comm_unique_id(void *id128)
{
	Nccl &l = nccl();
	if (!l.ok || !id128) return PGS_ERR_STATE;
	Nccl::unique_id id;
	if (l.GetUniqueId(&id) != 0) return PGS_ERR_CUDA;
	std::memcpy(id128, &id, sizeof id);
	return PGS_OK;
}

int comm_init(Engine &e, const void *id128, int rank, int world)
{
	Nccl &l = nccl();
	if (!l.ok) return e.fail(PGS_ERR_STATE, "pgs_comm_init: " + l.why);
	if (!id128 || world < 1 || rank < 0 || rank >= world) return e.fail(PGS_ERR_ARG, "pgs_comm_init: bad argument");
	e.drop_comm();
	Nccl::unique_id id;
	std::memcpy(&id, id128, sizeof id);
	Nccl::comm_t c = nullptr;
	const int r = l.CommInitRank(&c, world, id, rank);
	if (r != 0) return e.fail(PGS_ERR_CUDA, std::string("ncclCommInitRank: ") + l.GetErrorString(r));
	e.comm = c;
	e.comm_world = world;
	e.comm_rank = rank;
	return PGS_OK;
}

// Sum of the ranks' partial tiles, in place, on the engine's stream.  Pieces of at most 256 MB:
// each is one NCCL call (NVSwitch reduces in the fabric where NVLS is up; the piece size only
// bounds NCCL's staging).
int allreduce_tiles(Engine &e)
{
	Nccl &l = nccl();
	if (!e.comm) return e.fail(PGS_ERR_STATE, "pgs_mismatch_allreduce: call pgs_comm_init first");
	const int64_t words = mismatch_tile_words(e.n_genomes);
	const int64_t piece = (int64_t)64 << 20; // words
	for (int64_t at = 0; at < words; at += piece) {
		const int64_t now = std::min(piece, words - at);
		uint32_t *p = e.d_tiles.as<uint32_t>() + at;
		const int r = l.AllReduce(p, p, (size_t)now, kNcclUint32, kNcclSum, static_cast<Nccl::comm_t>(e.comm), e.stream);
		if (r != 0) return e.fail(PGS_ERR_CUDA, std::string("ncclAllReduce: ") + l.GetErrorString(r));
	}
	return PGS_OK;
}

// Several engines of ONE process, one per device (the CLI's --devices N): a communicator over their
// devices, every engine's partial tiles, one grouped all-reduce per piece, result from engine 0.
int mismatch_multi(Engine *const *engines, int n_engines, uint32_t *out_host, int64_t *sites_out)
{
	Engine &e0 = *engines[0];
	int64_t sites = 0;
	std::vector<int> rcs(n_engines, 0);
	std::vector<int64_t> part_sites(n_engines, 0);
	{
		std::vector<std::thread> pool;
		for (int d = 0; d < n_engines; d++)
			pool.emplace_back([&, d] {
				Engine &e = *engines[d];
				cudaSetDevice(e.device);
				rcs[d] = mismatch_tiles(e, &part_sites[d]);
			});
		for (auto &t : pool) t.join();
	}
	for (int d = 0; d < n_engines; d++) {
		if (rcs[d]) return d == 0 ? rcs[d] : e0.fail(rcs[d], "device " + std::to_string(d) + ": " + engines[d]->err);
		if (engines[d]->n_genomes != e0.n_genomes) return e0.fail(PGS_ERR_ARG, "pgs_mismatch_multi: the engines hold different trees");
		sites += part_sites[d];
	}
	if (sites_out) *sites_out = sites;
	if (n_engines > 1) {
		Nccl &l = nccl();
		if (!l.ok) return e0.fail(PGS_ERR_STATE, "pgs_mismatch_multi: " + l.why);
		std::vector<int> devs(n_engines);
		for (int d = 0; d < n_engines; d++) devs[d] = engines[d]->device;
		for (int a = 0; a < n_engines; a++)
			for (int b = a + 1; b < n_engines; b++)
				if (devs[a] == devs[b]) return e0.fail(PGS_ERR_ARG, "pgs_mismatch_multi: two engines share a device (NCCL needs one rank per device)");
		std::vector<Nccl::comm_t> comms(n_engines, nullptr);
		int r = l.CommInitAll(comms.data(), n_engines, devs.data());
		if (r != 0) return e0.fail(PGS_ERR_CUDA, std::string("ncclCommInitAll: ") + l.GetErrorString(r));
		const int64_t words = mismatch_tile_words(e0.n_genomes);
		const int64_t piece = (int64_t)64 << 20;
		for (int64_t at = 0; at < words && r == 0; at += piece) {
			const int64_t now = std::min(piece, words - at);
			l.GroupStart();
			for (int d = 0; d < n_engines && r == 0; d++) {
				uint32_t *p = engines[d]->d_tiles.as<uint32_t>() + at;
				r = l.AllReduce(p, p, (size_t)now, kNcclUint32, kNcclSum, comms[d], engines[d]->stream);
			}
			const int g = l.GroupEnd();
			if (r == 0) r = g;
		}
		for (int d = 0; d < n_engines; d++) {
			cudaSetDevice(engines[d]->device);
			cudaStreamSynchronize(engines[d]->stream);
		}
		for (auto c : comms)
			if (c) l.CommDestroy(c);
		if (r != 0) return e0.fail(PGS_ERR_CUDA, std::string("ncclAllReduce: ") + l.GetErrorString(r));
	}
	cudaSetDevice(e0.device);
	return out_host ? mismatch_tiles_to_host(e0, out_host) : PGS_OK;
}

} // namespace pgs
