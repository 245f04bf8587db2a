// kernels.cuh -- device tables and launchers of the sm_100a kernels (K0 root rows, K1 level
// sweep; K2 pairwise mismatch lives in mismatch.cu, the K3 text formatter in format.cu).  See DESIGN.md §4 for the roofline of each.
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
	uint32_t hca, hcb;            // decision_cmp(H[1]) of the two branches
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

// One task of k1_top_masks: the two branches below a node of the top of the tree (or, with
// edge_a == kRootEdge, the origin rows of the dense genes, written to ma_off).
struct TopMaskTask {
	int64_t ma_off, mb_off;       // where the masks of branch a / b go (uint4 units from the start of the row store)
	uint32_t edge_a, edge_b;      // child node indices (edge_a leads the sibling pair)
	uint32_t pad0, pad1;
	uint64_t h1a, h2a, h1b, h2b;  // H[1], H[2] of the two branches
};

// Fused walk of the sparse (accessory) genes, k1_walk_sparse: a thread carries one quad of blocks
// of ONE gene down the gene's own presence subtree (the nodes on the paths from the gene's origin
// to the genomes that carry it).  A widespread gene's subtree is cut into UNITS of bounded length;
// a unit below the origin starts from the origin row (drawn again, 4 Philox calls) carried across
// the branches on its path -- one Philox call per branch, the substitution masks being independent
// of the sequence -- so no unit waits for another one.
struct __align__(16) SparseUnit {
	uint32_t gid;        // gene_id (Philox counter word 1)
	uint32_t origin_row; // the gene's origin unit stores the origin row here; kNoRow for the other units
	uint32_t step_begin; // first of the unit's steps
	uint32_t n_steps;
	uint32_t path_begin; // branches from the origin down to the unit's root (SparsePathEdge)
	uint32_t path_len;
	uint32_t pad0, pad1;
};
static_assert(sizeof(SparseUnit) == 32, "SparseUnit is read as two uint4");
struct __align__(16) SparsePathEdge {
	uint32_t leader;   // first child of the branch's parent node (RNG key of the decision fields)
	uint32_t edge;     // child node index of the branch
	uint32_t side_h16; // bit 31: the branch is the second of its sibling pair; low 16 bits: H[1] >> 48
	uint32_t pad;
};
struct __align__(16) SparseStep {
	uint32_t dst_a, dst_b;   // rows the two children's values are stored to (kNoRow: not stored)
	uint32_t edge_a, edge_b; // child node indices; edge_a leads the sibling pair (RNG key) even if child a lacks the gene
	uint32_t h16;            // decision thresholds H[1] >> 48 of branch a (low half) and branch b (high half)
	uint32_t flags;          // kSp* below
	uint32_t in_row;         // kSpLoadRow: the row the input is read from (a sibling parked in its own row)
	uint32_t levels;         // bits 0-7: stack level the parked child goes to, bits 8-15: level the input comes from
};
static_assert(sizeof(SparseStep) == 32, "SparseStep is read as two uint4");
constexpr uint32_t kSpHasA = 1u, kSpHasB = 2u;       // the child carries the gene
constexpr uint32_t kSpCarryA = 4u, kSpCarryB = 8u;   // the child's value is the next step's input
constexpr uint32_t kSpParkA = 16u, kSpParkB = 32u;   // the child is parked in the thread's shared-memory stack
constexpr uint32_t kSpLoadSmem = 64u, kSpLoadRow = 128u; // where the input comes from (neither: carried in registers)

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
// Origin rows of the dense genes and the substitution masks of the branches below the top nodes
// (fully parallel: no node waits for its parent).
void launch_top_masks(const DeviceTables &t, const TopMaskTask *tasks, int32_t n_tasks, cudaStream_t s);
// One launch walks every unit: unit u starts from  origin row ^ masks on its path
// (unit_path[unit_path_off[u] .. unit_path_off[u+1])), stores that value at unit_store[u] if >= 0
// (a genome directly below a top node), then executes steps[unit_off[u] .. unit_off[u+1]) -- for
// every column slab.  root_off = the origin rows (all offsets in uint4 units from the row store).
struct WalkUnits {
	const WalkStep *steps;
	const int64_t *unit_off;
	const int32_t *unit_path_off;
	const int64_t *unit_path;
	const int64_t *unit_store;
	int64_t root_off;
	int32_t n_units;
	unsigned long long *trace; // development (PGS_WALK_TRACE): per CTA {start, walk begins, end [ns], SM}; nullptr = off
};
void launch_walk_dense(const DeviceTables &t, const WalkUnits &u, int32_t smem_levels, int variant, cudaStream_t s);

// Fused sparse sweep: origin rows and every branch of every sparse gene in one launch.
int sparse_walk_smem_levels();
cudaError_t prepare_walk_sparse();
void launch_walk_sparse(const DeviceTables &t, const SparseUnit *units, const SparseStep *steps, const SparsePathEdge *paths,
						int64_t n_units, cudaStream_t s);

} // namespace pgs
