// kernels.cuh -- device tables and launchers of the sm_100a kernels (K0 root rows, K1 level
// sweep, K2 pairwise mismatch, K3 text formatter).  See DESIGN.md §4 for the roofline of each.
#pragma once
#include "philox.cuh"
#include <cstdint>
#include <cuda_runtime.h>

namespace pgs {

constexpr uint32_t kNoRow = 0xFFFFFFFFu;

// One K1 work unit of the dense region: parent node and up to two children (right = -1: none).
struct DenseTask {
	int32_t parent, left, right;
};
// One K1 work unit of the sparse region: one parent row and the rows it produces.
struct SparseTask {
	uint32_t src, dst_left, dst_right; // row indices, kNoRow = child does not carry the gene
	uint32_t gid;                      // gene_id (Philox counter word 1)
	int32_t edge_left, edge_right;     // child node indices
};
struct RootTask {
	uint32_t row, gid;
};

// One step of the fused column walk (k1_walk_dense): carry the dense region of one node across the
// two branches below it.  Offsets are in uint4 units from the start of the row store.
struct __align__(16) WalkStep {
	int64_t in_off, a_off, b_off; // dense regions of the node and of its children a (pair leader) and b
	uint32_t edge_a, edge_b;      // child node indices = RNG edge ids (edge_a < edge_b)
	uint32_t sig_a, sig_b;        // kWalkSignalA/B: index of the ready flag of the unit rooted at child a / b
	uint32_t flags;               // kWalk* below
	uint32_t slots;               // bits 0-7: shared-memory level the parked child goes to, bits 8-15: level the input comes from
	uint64_t h1a, h2a, h1b, h2b;  // H[1], H[2] of the two branches (0 beyond kmax): the rare path needs no table for K <= 1
};
constexpr int kWalkStepQuads = 5;
static_assert(sizeof(WalkStep) == 16 * kWalkStepQuads, "WalkStep is staged as five uint4");
constexpr uint32_t kWalkLoad = 1u;     // input comes from memory (else: the value carried in registers)
constexpr uint32_t kWalkStoreA = 2u;   // write child a's row
constexpr uint32_t kWalkStoreB = 4u;
constexpr uint32_t kWalkCarryA = 8u;   // child a's value is the next step's input
constexpr uint32_t kWalkCarryB = 16u;
constexpr uint32_t kWalkParkA = 64u;   // the stored child a is an internal node that is re-read soon (keep in L2)
constexpr uint32_t kWalkParkB = 128u;
constexpr uint32_t kWalkSmemA = 256u;  // child a is parked in the thread's shared-memory stack (no global store)
constexpr uint32_t kWalkSmemB = 512u;
constexpr uint32_t kWalkLoadSmem = 1024u; // input comes from the shared-memory stack
constexpr uint32_t kWalkSignalA = 2048u;  // the stored child a is the root of another unit: raise its ready flag
constexpr uint32_t kWalkSignalB = 4096u;
constexpr int32_t kWalkWaitGenRoot = -2;  // unit_wait value of the unit that starts at the tree's root: it draws the origin rows

// Control block of the fused walk launch (device memory, uint32 words): a ticket counter that
// numbers the CTAs in the order they start, a count of finished CTAs, the epoch of the last
// completed sweep, then one ready flag per (unit root, column slab, warp) holding the epoch in
// which it was last raised.  The last CTA of a launch resets the counters and advances the epoch.
constexpr int kWalkCtlTicket = 0, kWalkCtlDone = 1, kWalkCtlEpoch = 2, kWalkCtlFlags = 4;

// Unsigned division by a launch constant: q = umulhi(x, mul) >> shift (x < 2^31).
struct FastDiv {
	uint32_t mul, shift, div;
};
inline FastDiv make_fastdiv(uint32_t d)
{
	FastDiv f;
	f.div = d;
	if (d == 1) {
		f.mul = 0;
		f.shift = 0;
		return f;
	}
	uint32_t s = 0;
	while ((1ull << s) < d) s++;
	// ceil(2^(32+s) / d) - 2^32 would need 33 bits; use the 2^(31+s) variant valid for x < 2^31
	f.shift = s - 1;
	f.mul = (uint32_t)(((1ull << (31 + s)) + d - 1) / d);
	return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t x, const FastDiv &f)
{
	return f.div == 1 ? x : (__umulhi(x, f.mul) >> f.shift);
}

struct DeviceTables {
	uint4 *rows;              // all (node, gene) rows, blocks_per_gene uint4 each
	const int64_t *row_base;  // [N] first row of node
	const uint64_t *h1;       // [N] floor(2^64 P(K>=1)) of the branch above node
	const uint64_t *tab;      // concatenated H[1..kmax] per node
	const int32_t *tab_off;   // [N] offset into tab
	const int32_t *kmax;      // [N]
	const uint32_t *dense_gid; // [D] gene_id of dense slot (ascending)
	uint32_t n_dense;         // D
	uint32_t blocks;          // blocks per gene (Bg)
	int32_t last_valid;       // real sites in the last block of a gene (1..64)
	FastDiv div_blocks;
	const uint64_t *tab8;     // [N][8] first eight thresholds H[1..8] of every branch (0 beyond kmax)
	PhiloxKeys keys;
};

extern int g_dense_variant;
extern int g_walk_variant;
void launch_rootgen(const DeviceTables &t, const RootTask *tasks, int64_t n_tasks, cudaStream_t s);
void launch_sweep_dense(const DeviceTables &t, const DenseTask *tasks, int32_t n_tasks, cudaStream_t s);
void launch_sweep_sparse(const DeviceTables &t, const SparseTask *tasks, int64_t n_tasks, cudaStream_t s);
// Fused dense sweep.  variant = choose_walk_variant(...); smem_levels = deepest shared-memory stack
// level any step uses, plus one.
int choose_walk_variant(uint32_t blocks_per_gene, double mean_hit);
// how many parked rows per thread that kernel can keep in shared memory; deeper levels of a unit's
// stack go to the global pool
int walk_smem_capacity(int variant);
// column slabs (CTAs per unit) that kernel makes of a dense region
uint32_t walk_slabs(int variant, uint32_t n_dense, uint32_t blocks_per_gene);
// before the first launch on a device: opt in to the largest dynamic shared memory any plan may ask
// for (the attribute is per function and device, not per engine: it is only ever set to this value)
cudaError_t prepare_walk_dense();
// One launch sweeps every unit: unit u executes steps[unit_off[u] .. unit_off[u+1]) for every column
// slab; unit_wait[u] = index of the ready flag its first load waits for (-1: none); ctl = control block.
void launch_walk_dense(const DeviceTables &t, const WalkStep *steps, const int64_t *unit_off, const int32_t *unit_wait,
					   uint32_t *ctl, int32_t n_units, int32_t smem_levels, int variant, cudaStream_t s);
// words of the control block for n_signals ready flags
size_t walk_ctl_words(int variant, uint32_t n_dense, uint32_t blocks_per_gene, int32_t n_signals);

// K2: mismatch counts between the dense regions of n leaves.  leaf_base[i] = first row of genome
// i's leaf.  out is n x n uint32 (zeroed by the launcher, upper triangle filled then mirrored).
void launch_mismatch(const uint4 *rows, const int64_t *leaf_base, int32_t n, int64_t words_per_leaf,
					 uint32_t *out, cudaStream_t s);

} // namespace pgs
