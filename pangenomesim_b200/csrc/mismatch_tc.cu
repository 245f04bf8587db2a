// mismatch_tc.cu -- K2 on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), the same counts
// as the popcount kernel of mismatch.cu, bit for bit.
//
// The reference has no counterpart (its distance.mat is the EXPECTED distance, img.cxx:121-165,
// printed at pangenomesim.cxx:214-234); north_star part 3 asks for the observed matrix.
//
// Why this is a contraction.  Give base c (2 bits: lo, hi) the vector
//     v(c) = ((-1)^lo, (-1)^hi, (-1)^(lo ^ hi))                (the corners of a tetrahedron)
// then v(a) . v(b) = 3 if a = b and -1 otherwise: the XOR of two different codes is non-zero and
// exactly two of (d_lo, d_hi, d_lo ^ d_hi) are set.  Over S sites with m mismatches the dot
// product of two genomes' expanded rows is 3 (S - m) - m = 3 S - 4 m, so
//     m = (3 S - dot) / 4                                     (exact: all terms are integers)
// Three +-1 entries per site is the minimum (four equidistant points need three dimensions).
// +-1 are exact in E4M3 (0x38 / 0xB8) and in INT8; the accumulator is FP32 (exact below 2^24: a
// column chunk holds far fewer than 5.6 M sites) or INT32.
//
//   k2_expand_pm1   genomes' 2-bit rows -> X[genome][3 bytes per site] for one column chunk
//                   (per 64-site block: 192 bytes; K order is the same for every row, which is all
//                   a dot product needs).
//   k2_tc_tiles     C = X X^T over the upper triangle.  CTA = 128 x 256 tile of genome pairs; one
//                   warp streams 128-byte K slices of both operands with TMA (128-byte swizzle)
//                   through a 4-stage mbarrier ring, one thread issues tcgen05.mma
//                   (kind::f8f6f4 or kind::i8, M = 128, N = 256, K = 32), the accumulator lives in
//                   256 TMEM columns, four warps read it back (tcgen05.ld), turn dot products into
//                   mismatch counts and add them to the packed 64 x 64 tiles of mismatch.cu.
#include "engine.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <mutex>
#include <vector>

namespace pgs {

namespace {

constexpr int kTcM = 128;          // genomes per tile, M side
constexpr int kTcN = 256;          // genomes per tile, N side
constexpr int kTcKBytes = 128;     // bytes of K per pipeline stage (one 128-byte swizzle row)
constexpr int kTcStages = 4;
constexpr int kTcBlockBytes = 192; // expanded bytes per 64-site block
constexpr int kTcThreads = 192;    // warp 0: TMA, warp 1: MMA + TMEM allocation, warps 2-5: epilogue
constexpr uint32_t kTcStageA = kTcM * kTcKBytes, kTcStageB = kTcN * kTcKBytes;
constexpr uint32_t kTcSmemBytes = kTcStages * (kTcStageA + kTcStageB) + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int kMmTileTc = 64; // side of the packed output tiles (mismatch.cu: kMmTile)

__host__ __device__ inline int64_t tri_row_start_tc(int64_t tiles, int64_t ti)
{
	return ti * tiles - ti * (ti - 1) / 2;
}

// ---- PTX wrappers ----------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Waits for the phase with the given parity.  A wait that lasts two seconds is a protocol error
// (a lost arrive): trap instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	uint32_t done = 0;
	long long t0 = 0;
	for (uint32_t spins = 0; !done; spins++) {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
					 : "=r"(done)
					 : "r"(bar), "r"(parity)
					 : "memory");
		if (!done && (spins & 1023u) == 1023u) {
			long long now;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			if (!t0) t0 = now;
			else if (now - t0 > 2000000000ll) {
				printf("k2_tc_tiles: barrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
				__trap();
			}
		}
	}
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int32_t c0, int32_t c1)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
				 "l"(map), "r"(bar), "r"(c0), "r"(c1)
				 : "memory");
}
__device__ __forceinline__ void tc_fence_before()
{
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after()
{
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool INT8> __device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	if (INT8)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
					 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
					 : "memory");
	else
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
					 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
					 : "memory");
}
// 32 lanes x 32 columns of 32 bits: thread t of the warp gets columns [col, col + 32) of TMEM lane (lane base + t)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32])
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
				 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
				 "%26, %27, %28, %29, %30, %31}, [%32];"
				 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
				   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
				   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
				   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
				 : "r"(taddr)
				 : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor of a K-major operand tile whose rows are 128 bytes (one 128-byte
// swizzle span), 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor: start address >> 4 in
// bits 0-13, leading byte offset >> 4 in 16-29, stride byte offset >> 4 in 32-45, version 1 in
// 46-47, layout type SWIZZLE_128B = 2 in 61-63).
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr)
{
	return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): accumulator format in bits 4-5 (1 = F32,
// 2 = S32), A / B formats in 7-9 / 10-12 (f8f6f4: 0 = E4M3; i8: 1 = signed), both operands K-major
// (bits 15, 16 = 0), N >> 3 in 17-22, M >> 4 in 24-28.
template <bool INT8> __host__ __device__ constexpr uint32_t tc_instr_desc()
{
	return (INT8 ? (2u << 4) | (1u << 7) | (1u << 10) : (1u << 4)) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
}

// ---- expansion ---------------------------------------------------------------------------------

// even bits of x -> low 16 bits
__device__ __forceinline__ uint32_t even_bits16(uint32_t x)
{
	x &= 0x55555555u;
	x = (x | (x >> 1)) & 0x33333333u;
	x = (x | (x >> 2)) & 0x0F0F0F0Fu;
	x = (x | (x >> 4)) & 0x00FF00FFu;
	x = (x | (x >> 8)) & 0x0000FFFFu;
	return x;
}

// bits 0, 2, 4, 6 of b -> the low bit of bytes 0..3 (the partial products that collide land on bits 7, 13, 19: masked)
__device__ __forceinline__ uint32_t spread4(uint32_t b)
{
	return ((b & 0x55u) * 0x00041041u) & 0x01010101u;
}
template <bool INT8> __device__ __forceinline__ uint32_t pm1_bytes(uint32_t bits01)
{
	// byte = +1 where the bit is 0, -1 where it is 1
	return INT8 ? (bits01 * 0xFEu) | 0x01010101u : (bits01 << 7) | 0x38383838u;
}

// One thread = one 16-site word of one genome: 4 bytes in, 48 bytes out (16 lo, 16 hi, 16 lo^hi).
// X[g * pitch + (block - first_block) * 192 + word * 48 ...].  Padding sites (code 0 in every row)
// are expanded like any other site: they always match and the epilogue counts them in S.
// Blocks [n_blocks, out_blocks) -- the odd block that completes the last K triple -- are written as
// zeros (they add nothing to a dot product).  Rows of padding genomes are never written and may
// hold anything: an output element depends on its own two rows only, and those outputs are dropped.
template <bool INT8>
__global__ void __launch_bounds__(256)
k2_expand_pm1(const uint32_t *__restrict__ rows, const int64_t *__restrict__ leaf_base_words, int32_t n, int64_t first_block, int64_t n_blocks,
			  int64_t out_blocks, int64_t pitch, uint8_t *__restrict__ X)
{
	const int64_t words = out_blocks * 4;
	const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
	if (idx >= (int64_t)n * words) return;
	const int32_t g = (int32_t)(idx / words);
	const int64_t w = idx - (int64_t)g * words;
	if (w >= n_blocks * 4) {
		uint4 *const z = reinterpret_cast<uint4 *>(X + (int64_t)g * pitch + w * 48);
		z[0] = z[1] = z[2] = make_uint4(0, 0, 0, 0);
		return;
	}
	const uint32_t v = __ldg(rows + leaf_base_words[g] + first_block * 4 + w);
	uint4 lo, hi, xo;
	uint32_t *const lp = &lo.x, *const hp = &hi.x, *const xp = &xo.x;
#pragma unroll
	for (int q = 0; q < 4; q++) {
		const uint32_t b = (v >> (8 * q)) & 0xFFu;
		const uint32_t l = spread4(b), h = spread4(b >> 1);
		lp[q] = pm1_bytes<INT8>(l);
		hp[q] = pm1_bytes<INT8>(h);
		xp[q] = pm1_bytes<INT8>(l ^ h);
	}
	uint4 *const out = reinterpret_cast<uint4 *>(X + (int64_t)g * pitch + w * 48);
	out[0] = lo;
	out[1] = hi;
	out[2] = xo;
}

// ---- the tile kernel ---------------------------------------------------------------------------

struct TcTile {
	int32_t tm, tn; // rows [128 tm, 128 tm + 128) x columns [256 tn, 256 tn + 256)
};

// A K "triple" = three 128-byte slices = 384 bytes = two 64-site blocks: the unit at which the K
// range is split between CTAs (a site's three entries never straddle a split).
template <bool INT8>
__global__ void __launch_bounds__(kTcThreads, 1)
k2_tc_tiles(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcTile *__restrict__ tile_list,
			int32_t n, int32_t triples_total, int32_t triples_per_split, int32_t real_blocks, uint32_t *__restrict__ tiles_out)
{
	extern __shared__ uint8_t smem_raw[];
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u; // the swizzle pattern is a function of address bits 4-9
	const uint32_t smem_a = base, smem_b = base + kTcStages * kTcStageA;
	const uint32_t bars = smem_b + kTcStages * kTcStageB;
	const uint32_t full0 = bars, empty0 = bars + 8 * kTcStages, tmem_full = bars + 16 * kTcStages, tmem_slot = tmem_full + 8;

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const TcTile tile = tile_list[blockIdx.x];
	const int32_t t0 = blockIdx.y * triples_per_split, t1 = min(t0 + triples_per_split, triples_total);
	const int32_t n_slices = 3 * (t1 - t0);

	if (threadIdx.x == 0) {
		for (int s = 0; s < kTcStages; s++) {
			mbar_init(full0 + 8 * s, 1);
			mbar_init(empty0 + 8 * s, 1);
		}
		mbar_init(tmem_full, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) { // this warp owns the TMEM allocation: 256 columns (the FP32 / INT32 accumulator of a 128 x 256 tile)
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	uint32_t tmem_acc;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot) : "memory");

	if (warp == 0) {
		if (lane == 0) {
			for (int32_t k = 0; k < n_slices; k++) {
				const int s = k % kTcStages;
				const uint32_t ph = (uint32_t)(k / kTcStages) & 1u;
				mbar_wait(empty0 + 8 * s, ph ^ 1u); // a fresh barrier passes the wait for the phase "before the first"
				mbar_expect_tx(full0 + 8 * s, kTcStageA + kTcStageB);
				const int32_t kx = (3 * t0 + k) * kTcKBytes;
				tma_load_2d(smem_a + s * kTcStageA, &map_a, full0 + 8 * s, kx, tile.tm * kTcM);
				tma_load_2d(smem_b + s * kTcStageB, &map_b, full0 + 8 * s, kx, tile.tn * kTcN);
			}
		}
	} else if (warp == 1) {
		if (lane == 0) {
			constexpr uint32_t idesc = tc_instr_desc<INT8>();
			for (int32_t k = 0; k < n_slices; k++) {
				const int s = k % kTcStages;
				const uint32_t ph = (uint32_t)(k / kTcStages) & 1u;
				mbar_wait(full0 + 8 * s, ph);
				tc_fence_after();
				const uint64_t adesc = tc_smem_desc(smem_a + s * kTcStageA), bdesc = tc_smem_desc(smem_b + s * kTcStageB);
#pragma unroll
				for (int j = 0; j < kTcKBytes / 32; j++) // K = 32 bytes per instruction: the start address advances inside the swizzle span
					tc_mma<INT8>(tmem_acc, adesc + 2 * j, bdesc + 2 * j, idesc, (uint32_t)(k | j));
				tc_commit(empty0 + 8 * s); // the stage is free once these MMAs have read it
			}
			tc_commit(tmem_full);
		}
	} else {
		// epilogue: warp w may touch TMEM lanes 32 (w % 4) .. + 31
		const int quarter = warp & 3;
		const int row = 32 * quarter + lane; // accumulator row = TMEM lane
		const int32_t gi = tile.tm * kTcM + row;
		const int32_t blocks_here = max(0, min(2 * t1, real_blocks) - 2 * t0); // real (expanded) blocks in this CTA's K range
		const int32_t k_real = blocks_here * kTcBlockBytes;                     // = 3 * sites
		const int64_t tiles64 = (n + kMmTileTc - 1) / kMmTileTc;
		const int64_t ti = gi / kMmTileTc;
		if (n_slices > 0) {
			mbar_wait(tmem_full, 0);
			tc_fence_after();
#pragma unroll 1
			for (int c = 0; c < kTcN / 32; c++) {
				uint32_t r[32];
				__syncwarp();
				tc_ld32(tmem_acc + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(32 * c), r);
				const int32_t gj0 = tile.tn * kTcN + 32 * c;
				const int64_t tj = gj0 / kMmTileTc;
				if (gi < n && ti <= tj && gj0 < n) {
					uint32_t *const dst = tiles_out + (tri_row_start_tc(tiles64, ti) + (tj - ti)) * (int64_t)(kMmTileTc * kMmTileTc) +
										  (gi % kMmTileTc) * kMmTileTc + (gj0 % kMmTileTc);
#pragma unroll
					for (int j = 0; j < 32; j++) {
						const int32_t dot = INT8 ? (int32_t)r[j] : __float2int_rn(__uint_as_float(r[j]));
						const uint32_t mm = (uint32_t)(k_real - dot) >> 2;
						if (gi < gj0 + j && gj0 + j < n && mm) atomicAdd(dst + j, mm);
					}
				}
			}
		}
		tc_fence_before();
	}
	__syncthreads();
	if (warp == 1) {
		tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(256) : "memory");
	}
}

// ---- the CTA-pair kernel -----------------------------------------------------------------------
//
// Two CTAs of one TPC (a cluster of two) work on one 256 x 512 tile with tcgen05.mma.cta_group::2:
// CTA r holds rows [128 r, 128 r + 128) of the A side and, for each column half h (N_h <= 256
// columns, a multiple of 16), rows [N_h / 2 * r, + N_h / 2) of the B side -- three 128 x 128-byte
// boxes per K slice, 48 KB per stage, 48 bytes of operands per clock per SM at full tensor rate
// instead of the 96 of k2_tc_tiles.  Each CTA's accumulator is its 128 rows x 512 columns: all of
// its TMEM.  The leader (rank 0) issues the MMAs (M = 256, N = N_h, twice per K step); both CTAs
// stream their boxes with TMA, all of them signalling the LEADER's full barrier; tcgen05.commit
// multicasts the "stage is free" / "accumulator is complete" arrivals to both CTAs.
//
// MODE: operands E4M3 (kind::f8f6f4), INT8 (kind::i8), or E2M1 -- +-1 are exact in FP4 as well --
// with kind::mxf4 block scaling, every scale 2^0: twice the K per instruction (64), half the
// operand bytes per site.  The scale factors live in TMEM; all of them are the UE8M0 byte 127, so
// the epilogue warps fill 32 columns with 0x7F7F7F7F once and every scale-factor read of any MMA
// finds 1.0 there, whatever its layout.  Those 32 columns come out of the second column half
// (N_1 <= 224: a 256 x 480 tile).

enum : int { kModeE4M3 = 0, kModeInt8 = 1, kModeFp4 = 2 };

constexpr int kPairM = 256;
constexpr uint32_t kPairBox = 128 * kTcKBytes;     // one 128-row box
constexpr uint32_t kPairStage = 3 * kPairBox;      // A, B half 0, B half 1
constexpr int kPairStages = 4;
constexpr uint32_t kPairSmemBytes = kPairStages * kPairStage + 1024 + 256;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the same offset in CTA 0 of the pair (cute: Sm100MmaPeerBitMask)
constexpr uint32_t kSfColumn = 480;                // FP4: TMEM columns [480, 512) hold the scale factors
constexpr int kPairPrefetchLead = 48;              // K slices before the end at which the epilogue warps start prefetching

__host__ __device__ constexpr int pair_tile_n(int mode)
{
	return mode == kModeFp4 ? 480 : 512;
}
__host__ __device__ constexpr int pair_block_bytes(int mode) // expanded bytes per 64-site block
{
	return mode == kModeFp4 ? 96 : 192;
}
__host__ __device__ constexpr int pair_blocks_per_triple(int mode) // 64-site blocks in three 128-byte K slices
{
	return 384 / pair_block_bytes(mode);
}

struct PairUnit {
	int32_t tm, tn;         // rows [256 tm, + 256) x columns [TILE_N tn, + TILE_N)
	int32_t t_begin, t_end; // K range in triples (three 128-byte slices)
	int32_t n0, n1;         // columns of the two halves that are computed (multiples of 16; 0: the half is skipped --
	                        // entirely below the diagonal or beyond the last genome)
	int32_t atomic;         // 1: other units add to the same output tiles (K split) -- atomicAdd
	int32_t pad;
};

__device__ __forceinline__ void cluster_sync_all()
{
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int32_t c0, int32_t c1)
{
	asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
				 "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
				 : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar)
{
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
				 : "memory");
}
template <int MODE>
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t tmem_sf)
{
	if (MODE == kModeInt8)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
					 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
					 : "memory");
	else if (MODE == kModeE4M3)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
					 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
					 : "memory");
	else
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%5], "
					 "p;\n\t}" ::"r"(tmem_d),
					 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sf)
					 : "memory");
}
// M = 256; N is added at run time ((N >> 3) << 17).  Block-scaled form (cute::UMMA::InstrDescriptorBlockScaled):
// A / B format 1 = E2M1 (bits 7-9, 10-12), scale format UE8M0 (bit 23), scale-factor ids 0, K = 64 (bit 31 = 0).
template <int MODE> __host__ __device__ constexpr uint32_t tc_instr_desc_pair()
{
	return (MODE == kModeInt8 ? (2u << 4) | (1u << 7) | (1u << 10) : MODE == kModeE4M3 ? (1u << 4) : (1u << 7) | (1u << 10) | (1u << 23)) |
		   ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void tc_st32_const(uint32_t taddr, uint32_t v)
{
	asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
				 "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
				 "r"(v)
				 : "memory");
	asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
k2_tc_pair(const __grid_constant__ CUtensorMap map, const PairUnit *__restrict__ units, int32_t n, int32_t real_blocks, uint32_t *__restrict__ tiles_out)
{
	extern __shared__ uint8_t smem_raw[];
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bars = base + kPairStages * kPairStage;
	const uint32_t full0 = bars, empty0 = bars + 8 * kPairStages, tmem_full = bars + 16 * kPairStages, tmem_slot = tmem_full + 8;
	const uint32_t near_end = tmem_full + 16; // "the last K slices are under way" (multicast commit, like tmem_full)

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t rank;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
	const PairUnit u = units[blockIdx.x >> 1];
	const int32_t n_slices = 3 * (u.t_end - u.t_begin);
	const int32_t col0[2] = {u.tn * pair_tile_n(MODE), u.tn * pair_tile_n(MODE) + 256}; // first genome of each column half
	const int32_t nh[2] = {u.n0, u.n1};

	if (threadIdx.x == 0) {
		for (int s = 0; s < kPairStages; s++) {
			mbar_init(full0 + 8 * s, 1);  // the leader's arrive.expect_tx; the bytes of both CTAs complete it
			mbar_init(empty0 + 8 * s, 1); // one multicast commit per use
		}
		mbar_init(tmem_full, 1);
		mbar_init(near_end, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) { // the same warp of both CTAs: all 512 columns of both SMs
		asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	uint32_t tmem_acc;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot) : "memory");
	if (MODE == kModeFp4 && warp >= 2) // every scale factor = 2^0
		tc_st32_const(tmem_acc + ((uint32_t)(32 * (warp & 3)) << 16) + kSfColumn, 0x7F7F7F7Fu);
	tc_fence_before();
	cluster_sync_all(); // the peer's barriers (and scale factors) exist before anything uses them
	tc_fence_after();

	if (warp == 0) {
		if (lane == 0) {
			const int32_t row_a = u.tm * kPairM + 128 * (int32_t)rank;
			const uint32_t bytes_both = 2u * kPairBox * (uint32_t)(1 + (nh[0] > 0) + (nh[1] > 0));
			for (int32_t k = 0; k < n_slices; k++) {
				const int s = k % kPairStages;
				const uint32_t ph = (uint32_t)(k / kPairStages) & 1u;
				mbar_wait(empty0 + 8 * s, ph ^ 1u);
				if (rank == 0) mbar_expect_tx(full0 + 8 * s, bytes_both);
				const uint32_t leader_full = (full0 + 8 * s) & kPeerBitMask;
				const uint32_t st = base + s * kPairStage;
				const int32_t kx = (3 * u.t_begin + k) * kTcKBytes;
				tma_load_2d_pair(st, &map, leader_full, kx, row_a);
#pragma unroll
				for (int h = 0; h < 2; h++) // rows past the last genome are zero-filled by the TMA unit
					if (nh[h]) tma_load_2d_pair(st + (1 + h) * kPairBox, &map, leader_full, kx, col0[h] + (nh[h] >> 1) * (int32_t)rank);
			}
		}
	} else if (warp == 1) {
		if (lane == 0 && rank == 0) {
			constexpr uint32_t idesc = tc_instr_desc_pair<MODE>();
			for (int32_t k = 0; k < n_slices; k++) {
				const int s = k % kPairStages;
				const uint32_t ph = (uint32_t)(k / kPairStages) & 1u;
				mbar_wait(full0 + 8 * s, ph);
				tc_fence_after();
				const uint32_t st = base + s * kPairStage;
				const uint64_t adesc = tc_smem_desc(st);
#pragma unroll
				for (int h = 0; h < 2; h++) {
					if (!nh[h]) continue;
					const uint64_t bdesc = tc_smem_desc(st + (1 + h) * kPairBox);
					const uint32_t idesc_h = idesc | ((uint32_t)(nh[h] >> 3) << 17);
#pragma unroll
					for (int j = 0; j < kTcKBytes / 32; j++) // 32 bytes of K per instruction (32 E4M3 / INT8 or 64 E2M1 entries)
						tc_mma_pair<MODE>(tmem_acc + 256u * h, adesc + 2 * j, bdesc + 2 * j, idesc_h, (uint32_t)(k | j), tmem_acc + kSfColumn);
				}
				tc_commit_pair(empty0 + 8 * s);
				if (k == n_slices - kPairPrefetchLead) tc_commit_pair(near_end); // the epilogue warps may start fetching their output lines
			}
			if (n_slices < kPairPrefetchLead) tc_commit_pair(near_end);
			tc_commit_pair(tmem_full);
		}
	} else {
		const int quarter = warp & 3;
		const int row = 32 * quarter + lane;
		const int32_t gi = u.tm * kPairM + 128 * (int32_t)rank + row;
		constexpr int kBpt = pair_blocks_per_triple(MODE);
		const int32_t blocks_here = max(0, min(kBpt * u.t_end, real_blocks) - kBpt * u.t_begin);
		const int32_t k_real = blocks_here * 192; // K entries that are real sites: 3 per site
		const int64_t tiles64 = (n + kMmTileTc - 1) / kMmTileTc;
		const int64_t ti = gi / kMmTileTc;
		if (n_slices > 0) {
			// the counts are added to what earlier column chunks left in memory: pull those lines into L2 while the
			// last K slices are still being multiplied
			mbar_wait(near_end, 0);
			if (!u.atomic && gi < n)
				for (int h = 0; h < 2; h++)
					for (int c = 0; 32 * c < nh[h]; c++) {
						const int32_t gj0 = col0[h] + 32 * c;
						const int64_t tj = gj0 / kMmTileTc;
						if (ti <= tj && gj0 < n)
							asm volatile("prefetch.global.L2 [%0];" ::"l"(tiles_out + (tri_row_start_tc(tiles64, ti) + (tj - ti)) * (int64_t)(kMmTileTc * kMmTileTc) +
																		  (gi % kMmTileTc) * kMmTileTc + (gj0 % kMmTileTc)));
					}
			mbar_wait(tmem_full, 0);
			tc_fence_after();
#pragma unroll 1
			for (int h = 0; h < 2; h++) {
#pragma unroll 1
				for (int c = 0; 32 * c < nh[h]; c++) {
					uint32_t r[32];
					__syncwarp();
					tc_ld32(tmem_acc + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(256 * h + 32 * c), r);
					const int32_t gj0 = col0[h] + 32 * c;
					const int32_t gj_end = min(n, col0[h] + nh[h]); // columns past N_h hold whatever an earlier CTA left in TMEM
					const int64_t tj = gj0 / kMmTileTc;
					if (gi < n && ti <= tj && gj0 < gj_end) {
						uint32_t *const dst = tiles_out + (tri_row_start_tc(tiles64, ti) + (tj - ti)) * (int64_t)(kMmTileTc * kMmTileTc) +
											  (gi % kMmTileTc) * kMmTileTc + (gj0 % kMmTileTc);
#pragma unroll
						for (int j = 0; j < 32; j++) {
							const int32_t dot = MODE == kModeInt8 ? (int32_t)r[j] : __float2int_rn(__uint_as_float(r[j]));
							const uint32_t mm = (uint32_t)(k_real - dot) >> 2;
							r[j] = (gi < gj0 + j && gj0 + j < gj_end) ? mm : 0u;
						}
						if (u.atomic) {
#pragma unroll
							for (int j = 0; j < 32; j++)
								if (r[j]) atomicAdd(dst + j, r[j]);
						} else { // this unit owns its output tiles during this launch
							uint4 *const d4 = reinterpret_cast<uint4 *>(dst);
#pragma unroll
							for (int q = 0; q < 8; q++) {
								uint4 v = d4[q];
								v.x += r[4 * q], v.y += r[4 * q + 1], v.z += r[4 * q + 2], v.w += r[4 * q + 3];
								d4[q] = v;
							}
						}
					}
				}
			}
		}
		tc_fence_before();
	}
	cluster_sync_all(); // neither CTA may leave (or free TMEM) while the pair's MMAs read its shared memory
	if (warp == 1) {
		tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512) : "memory");
	}
}

// FP4 operands: one thread = two 16-site words (half a block) of one genome: 8 bytes in, 48 bytes
// out -- 16 bytes of lo entries (32 nibbles), 16 of hi, 16 of lo ^ hi; nibble = E2M1 +1 (0x2) or -1
// (0xA).  The order of the entries inside a row is the same for every genome, which is all a dot
// product needs.  Blocks [n_blocks, out_blocks) are written as zeros.
__device__ __forceinline__ uint32_t spread8_to_nibble_msb(uint32_t b) // bit k of the low byte -> bit 4 k + 3
{
	uint32_t x = b & 0xFFu;
	x = (x | (x << 12)) & 0x000F000Fu;
	x = (x | (x << 6)) & 0x03030303u;
	x = (x | (x << 3)) & 0x11111111u;
	return x << 3;
}
__global__ void __launch_bounds__(256)
k2_expand_fp4(const uint2 *__restrict__ rows, const int64_t *__restrict__ leaf_base_words, int32_t n, int64_t first_block, int64_t n_blocks,
			  int64_t out_blocks, int64_t pitch, uint8_t *__restrict__ X)
{
	const int64_t halves = out_blocks * 2;
	const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
	if (idx >= (int64_t)n * halves) return;
	const int32_t g = (int32_t)(idx / halves);
	const int64_t w = idx - (int64_t)g * halves;
	uint4 *const out = reinterpret_cast<uint4 *>(X + (int64_t)g * pitch + w * 48);
	if (w >= n_blocks * 2) {
		out[0] = out[1] = out[2] = make_uint4(0, 0, 0, 0);
		return;
	}
	const uint2 v = __ldg(rows + (leaf_base_words[g] >> 1) + first_block * 2 + w);
	const uint32_t l0 = even_bits16(v.x), h0 = even_bits16(v.x >> 1), l1 = even_bits16(v.y), h1 = even_bits16(v.y >> 1);
	uint4 lo, hi;
	lo.x = spread8_to_nibble_msb(l0), lo.y = spread8_to_nibble_msb(l0 >> 8), lo.z = spread8_to_nibble_msb(l1), lo.w = spread8_to_nibble_msb(l1 >> 8);
	hi.x = spread8_to_nibble_msb(h0), hi.y = spread8_to_nibble_msb(h0 >> 8), hi.z = spread8_to_nibble_msb(h1), hi.w = spread8_to_nibble_msb(h1 >> 8);
	constexpr uint32_t one = 0x22222222u;
	out[0] = make_uint4(lo.x | one, lo.y | one, lo.z | one, lo.w | one);
	out[1] = make_uint4(hi.x | one, hi.y | one, hi.z | one, hi.w | one);
	out[2] = make_uint4((lo.x ^ hi.x) | one, (lo.y ^ hi.y) | one, (lo.z ^ hi.z) | one, (lo.w ^ hi.w) | one);
}

// ---- host side ---------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
								  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
	static EncodeTiledFn fn = nullptr;
	static std::once_flag once;
	std::call_once(once, [] {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
			fn = reinterpret_cast<EncodeTiledFn>(p);
	});
	return fn;
}

// X as a 2-D byte tensor [rows][pitch]; a box = `box_rows` rows x 128 bytes, 128-byte swizzle
int make_map(Engine &e, CUtensorMap *map, void *X, int64_t rows, int64_t pitch, int box_rows)
{
	EncodeTiledFn fn = encode_tiled();
	if (!fn) return e.fail(PGS_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
	const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
	const cuuint64_t strides[1] = {(cuuint64_t)pitch};
	const cuuint32_t box[2] = {(cuuint32_t)kTcKBytes, (cuuint32_t)box_rows};
	const cuuint32_t elem[2] = {1, 1};
	const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, X, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
						  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) return e.fail(PGS_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
	return PGS_OK;
}

// Tiles of the upper triangle in launch order: bands of 8 tile rows, column by column inside a band,
// so that the CTAs resident at any time share a few row blocks of X (L2 reuse).
std::vector<TcTile> tile_order(int32_t n)
{
	const int tm_count = (n + kTcM - 1) / kTcM, tn_count = (n + kTcN - 1) / kTcN;
	std::vector<TcTile> v;
	constexpr int kBand = 8;
	for (int band = 0; band < tm_count; band += kBand)
		for (int tn = 0; tn < tn_count; tn++)
			for (int tm = band; tm < std::min(band + kBand, tm_count); tm++)
				if ((int64_t)tm * kTcM < (int64_t)tn * kTcN + kTcN - 1) // some pair i < j inside
					v.push_back(TcTile{tm, tn});
	return v;
}

template <bool INT8> int run_tensor(Engine &e, int64_t budget)
{
	const int32_t n = e.n_genomes;
	const int64_t D = (int64_t)e.dense_gid.size();
	const int64_t blocks_total = D * e.blocks;
	const int64_t n_pad = ((int64_t)n + kTcN - 1) / kTcN * kTcN;
	// column chunks: whole K triples (two blocks); FP32 accumulation stays exact below 2^24 = 3 * 64 * 87381 blocks
	int64_t chunk_blocks = std::max<int64_t>(2, budget / (n_pad * kTcBlockBytes));
	chunk_blocks = std::min<int64_t>(chunk_blocks, 80000);
	chunk_blocks = std::min(chunk_blocks, blocks_total + (blocks_total & 1));
	chunk_blocks &= ~(int64_t)1;
	const int64_t pitch = chunk_blocks * kTcBlockBytes;
	const size_t x_bytes = (size_t)n_pad * (size_t)pitch;
	int rc;
	if (e.d_planes.bytes < x_bytes && (rc = e.alloc(e.d_planes, x_bytes, "mismatch operand matrix"))) return rc;
	cudaError_t ce;

	CUtensorMap map_a, map_b;
	if ((rc = make_map(e, &map_a, e.d_planes.p, n_pad, pitch, kTcM))) return rc;
	if ((rc = make_map(e, &map_b, e.d_planes.p, n_pad, pitch, kTcN))) return rc;

	const std::vector<TcTile> order = tile_order(n);
	if ((rc = e.upload_keep(e.d_mismatch, order.data(), order.size() * sizeof(TcTile), "mismatch tile list"))) return rc;

	// the attribute is per function and per device: set it on every call (cheap; once per process would miss other devices)
	ce = cudaFuncSetAttribute(k2_tc_tiles<INT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "k2_tc_tiles shared memory");

	for (int64_t b0 = 0; b0 < blocks_total; b0 += chunk_blocks) {
		const int64_t nb = std::min(chunk_blocks, blocks_total - b0);
		const int64_t items = (int64_t)n * (nb + (nb & 1)) * 4;
		k2_expand_pm1<INT8><<<(unsigned)((items + 255) / 256), 256, 0, e.stream>>>(reinterpret_cast<const uint32_t *>(e.tables.rows),
																					 e.d_leaf_base_words.as<int64_t>(), n, b0, nb, nb + (nb & 1), pitch,
																					 e.d_planes.as<uint8_t>());
		const int64_t triples = (nb + 1) / 2;
		// split K so that few tiles still fill the 148 SMs
		int64_t splits = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(triples, 65535), (148 + (int64_t)order.size() - 1) / (int64_t)order.size()));
		const int64_t per_split = (triples + splits - 1) / splits;
		splits = (triples + per_split - 1) / per_split;
		dim3 grid((unsigned)order.size(), (unsigned)splits);
		k2_tc_tiles<INT8><<<grid, kTcThreads, kTcSmemBytes, e.stream>>>(map_a, map_b, e.d_mismatch.as<TcTile>(), n, (int32_t)triples, (int32_t)per_split,
																		 (int32_t)nb, e.d_tiles.as<uint32_t>());
	}
	ce = cudaGetLastError();
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch (tensor) launch");
	return PGS_OK;
}

// Pair tiles of the upper triangle in launch order (bands of 4 tile rows, column by column), cut
// into units: whole tiles while they fill complete waves of `slots` CTA pairs, the remaining tiles
// split along K into equal pieces so that the last wave ends together (those pieces add atomically).
std::vector<PairUnit> pair_units(int32_t n, int32_t triples, int slots, int tile_n, int n1_max)
{
	const int tm_count = (n + kPairM - 1) / kPairM, tn_count = (n + tile_n - 1) / tile_n;
	std::vector<PairUnit> tiles;
	constexpr int kBand = 4;
	auto cols = [&](int64_t first, int cap) { // computed columns of a half starting at genome `first`: up to the last genome, in 16s
		return (int32_t)std::max<int64_t>(0, std::min<int64_t>(cap, (n - first + 15) / 16 * 16));
	};
	for (int band = 0; band < tm_count; band += kBand)
		for (int tn = 0; tn < tn_count; tn++)
			for (int tm = band; tm < std::min(band + kBand, tm_count); tm++) {
				const int64_t row_first = (int64_t)tm * kPairM, c0 = (int64_t)tn * tile_n, c1 = c0 + 256;
				PairUnit u{tm, tn, 0, triples, cols(c0, 256), cols(c1, n1_max), 0, 0};
				if (c0 + u.n0 - 1 <= row_first) u.n0 = 0; // no pair i < j in this half
				if (c1 + u.n1 - 1 <= row_first) u.n1 = 0;
				if (u.n0 || u.n1) tiles.push_back(u);
			}
	const size_t whole = tiles.size() / (size_t)slots * (size_t)slots;
	const size_t rest = tiles.size() - whole;
	// pieces per remaining tile: the last rounds take ceil(rest * p / slots) / p of a tile's time; every piece
	// pays its own epilogue (about 2 % of a whole tile)
	int pieces = 1;
	double best = 1e30;
	for (int p = 1; rest && p <= std::max<int64_t>(12, std::min<int64_t>(slots / (int64_t)rest, 32)) && p <= triples; p++) {
		const double cost = (double)((rest * p + slots - 1) / slots) / p + 0.02 * p;
		if (cost < best - 1e-9) {
			best = cost;
			pieces = p;
		}
	}
	if (pieces <= 1) return tiles;
	std::vector<PairUnit> units(tiles.begin(), tiles.begin() + (std::ptrdiff_t)whole);
	const int32_t per = (triples + pieces - 1) / pieces;
	for (size_t i = whole; i < tiles.size(); i++)
		for (int32_t t = 0; t < triples; t += per) {
			PairUnit u = tiles[i];
			u.t_begin = t;
			u.t_end = std::min(t + per, triples);
			u.atomic = 1;
			units.push_back(u);
		}
	return units;
}

// The expansion of chunk c + 1 runs on a second stream beside the tile kernel of chunk c (two operand
// buffers): it is memory-bound work that fits into the SMs next to the one resident tile CTA.
int aux_stream(Engine &e)
{
	if (e.aux_stream) return PGS_OK;
	int lo = 0, hi = 0;
	cudaDeviceGetStreamPriorityRange(&lo, &hi);
	cudaError_t ce = cudaStreamCreateWithPriority(&e.aux_stream, cudaStreamNonBlocking, hi);
	for (int i = 0; i < 5 && ce == cudaSuccess; i++) ce = cudaEventCreateWithFlags(&e.mm_ev[i], cudaEventDisableTiming);
	return ce == cudaSuccess ? PGS_OK : e.cuda_fail(ce, "mismatch helper stream");
}

template <int MODE> int run_tensor_pair(Engine &e, int64_t budget)
{
	constexpr int kBpt = pair_blocks_per_triple(MODE), kBlockBytes = pair_block_bytes(MODE), kTileN = pair_tile_n(MODE);
	const int32_t n = e.n_genomes;
	const int64_t D = (int64_t)e.dense_gid.size();
	const int64_t blocks_total = D * e.blocks;
	// column chunks of whole K triples, two operand buffers inside the budget
	int64_t chunk_blocks = std::max<int64_t>(kBpt, budget / 2 / ((int64_t)n * kBlockBytes));
	chunk_blocks = std::min<int64_t>(chunk_blocks, 80000); // FP32 accumulation stays exact: 3 * 64 * 80000 < 2^24
	chunk_blocks = std::min(chunk_blocks, (blocks_total + kBpt - 1) / kBpt * kBpt);
	chunk_blocks = chunk_blocks / kBpt * kBpt;
	const int64_t pitch = chunk_blocks * kBlockBytes;
	const size_t x_bytes = ((size_t)n * (size_t)pitch + 1023) / 1024 * 1024;
	const int64_t n_chunks = (blocks_total + chunk_blocks - 1) / chunk_blocks;
	const int n_bufs = n_chunks > 1 ? 2 : 1;
	int rc;
	if (e.d_planes.bytes < n_bufs * x_bytes && (rc = e.alloc(e.d_planes, n_bufs * x_bytes, "mismatch operand matrix"))) {
		if (rc == PGS_ERR_NOMEM && budget > ((int64_t)256 << 20)) return run_tensor_pair<MODE>(e, budget / 2); // shorter chunks
		return rc;
	}
	if ((rc = aux_stream(e))) return rc;
	CUtensorMap map[2];
	for (int b = 0; b < n_bufs; b++)
		if ((rc = make_map(e, &map[b], e.d_planes.as<uint8_t>() + b * x_bytes, n, pitch, 128))) return rc;
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e.device);
	const int slots = std::max(1, sms / 2);

	// unit lists of a full chunk and of the (shorter) last chunk, uploaded once
	const int64_t last_nb = blocks_total % chunk_blocks;
	const std::vector<PairUnit> units_full = pair_units(n, (int32_t)(chunk_blocks / kBpt), slots, kTileN, kTileN - 256);
	const std::vector<PairUnit> units_last =
		last_nb ? pair_units(n, (int32_t)((last_nb + kBpt - 1) / kBpt), slots, kTileN, kTileN - 256) : std::vector<PairUnit>();
	std::vector<PairUnit> all(units_full);
	all.insert(all.end(), units_last.begin(), units_last.end());
	if ((rc = e.upload_keep(e.d_mismatch, all.data(), all.size() * sizeof(PairUnit), "mismatch unit list"))) return rc;

	cudaError_t ce = cudaFuncSetAttribute(k2_tc_pair<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "k2_tc_pair shared memory");

	// mm_ev[0]: the rows exist (main stream); [1], [2]: buffer b is expanded; [3], [4]: buffer b has been consumed
	ce = cudaEventRecord(e.mm_ev[0], e.stream);
	if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e.aux_stream, e.mm_ev[0], 0);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch (tensor) streams");
	for (int64_t c = 0; c < n_chunks; c++) {
		const int b = (int)(c & 1) % n_bufs;
		const int64_t b0 = c * chunk_blocks;
		const int64_t nb = std::min(chunk_blocks, blocks_total - b0);
		const int64_t out_blocks = (nb + kBpt - 1) / kBpt * kBpt;
		const bool last = nb < chunk_blocks;
		uint8_t *const X = e.d_planes.as<uint8_t>() + b * x_bytes;
		if (c >= 2 && (ce = cudaStreamWaitEvent(e.aux_stream, e.mm_ev[3 + b], 0)) != cudaSuccess) break;
		if (MODE == kModeFp4) {
			const int64_t items = (int64_t)n * out_blocks * 2;
			k2_expand_fp4<<<(unsigned)((items + 255) / 256), 256, 0, e.aux_stream>>>(reinterpret_cast<const uint2 *>(e.tables.rows),
																					  e.d_leaf_base_words.as<int64_t>(), n, b0, nb, out_blocks, pitch, X);
		} else {
			const int64_t items = (int64_t)n * out_blocks * 4;
			k2_expand_pm1<MODE == kModeInt8><<<(unsigned)((items + 255) / 256), 256, 0, e.aux_stream>>>(
				reinterpret_cast<const uint32_t *>(e.tables.rows), e.d_leaf_base_words.as<int64_t>(), n, b0, nb, out_blocks, pitch, X);
		}
		if ((ce = cudaEventRecord(e.mm_ev[1 + b], e.aux_stream)) != cudaSuccess) break;
		if ((ce = cudaStreamWaitEvent(e.stream, e.mm_ev[1 + b], 0)) != cudaSuccess) break;
		const PairUnit *const list = e.d_mismatch.as<PairUnit>() + (last ? units_full.size() : 0);
		const size_t count = last ? units_last.size() : units_full.size();
		k2_tc_pair<MODE><<<(unsigned)(2 * count), kTcThreads, kPairSmemBytes, e.stream>>>(map[b], list, n, (int32_t)nb, e.d_tiles.as<uint32_t>());
		if ((ce = cudaEventRecord(e.mm_ev[3 + b], e.stream)) != cudaSuccess) break;
	}
	if (ce == cudaSuccess) ce = cudaGetLastError();
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_mismatch (tensor) launch");
	return PGS_OK;
}

} // namespace

// Same contract as mismatch_tiles (mismatch.cu): partial counts of this engine's dense genes into
// e.d_tiles, asynchronous on the engine's stream.  e.d_tiles is allocated and cleared by the caller.
// impl: 1 = CTA pairs, E4M3; 2 = CTA pairs, INT8; 3 / 4 = the single-CTA kernel, E4M3 / INT8;
// 5 = CTA pairs, E2M1 with unit block scales.
int mismatch_tiles_tensor(Engine &e, int impl)
{
	// Two operand buffers.  A tile's epilogue (read-modify-write of its counts) is paid once per column chunk, so a
	// chunk should hold some 2000 blocks: 4 GB up to n = 10 000, more for more genomes (config 5: 16 GB of the 180).
	int64_t budget = std::min<int64_t>((int64_t)16384 << 20, std::max<int64_t>((int64_t)4096 << 20, (int64_t)e.n_genomes * 96 * 2 * 2048));
	if (const char *v = std::getenv("PGS_K2_PLANES_MB")) budget = std::max<int64_t>(1, std::atol(v)) << 20; // tuning knob
	switch (impl) {
	case 1: return run_tensor_pair<kModeE4M3>(e, budget);
	case 2: return run_tensor_pair<kModeInt8>(e, budget);
	case 3: return run_tensor<false>(e, budget / 2);
	case 4: return run_tensor<true>(e, budget / 2);
	default: return run_tensor_pair<kModeFp4>(e, budget);
	}
}

} // namespace pgs
