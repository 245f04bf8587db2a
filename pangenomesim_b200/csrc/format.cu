// format.cu -- K3: 2-bit rows -> ASCII FASTA / MAF text on the device, streamed to the host
// through double-buffered pinned memory.
//
// Replaces gene::to_fasta (reference gene.h:42-52), the FASTA writers pangenomesim.cxx:117-157
// and the MAF writer pangenomesim.cxx:236-293.  The reference builds every record as a
// std::string on the host; here the whole text (headers included) is produced by the GPU in
// file order, so the host only hands contiguous chunks to write(2) or to a sink callback.
//
// Record offsets are closed-form (no per-record descriptor tables): inside genomeNNN.fasta the
// k-th record of a genome starts at  k * (name_len + L + 4) + sum of decimal digits of the
// gene ids before it, and both terms come from two small prefix tables; a MAF block is laid out
// the same way from a per-gene block offset and a per-carrier prefix.
#include "engine.h"
#include "expand.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

namespace pgs {

int format_genome_name(int64_t genome_index, char *dst, size_t cap)
{
	// reference util.cxx:15-27: "genome" + zero padding to 3 digits + (index + 1); "genomeref"
	// for -1.  The reference aborts for >= 1000 genomes (negative padding length); here the
	// padding is clamped at zero, which leaves every name below 1000 unchanged.
	if (genome_index < 0) return std::snprintf(dst, cap, "genomeref");
	return std::snprintf(dst, cap, "genome%03lld", (long long)(genome_index + 1));
}

// ---------------------------------------------------------------------------------------------
// device text helpers

__host__ __device__ inline int dec_digits(uint64_t x)
{
	// 64-bit division is a long software routine on the GPU: stay in 32 bits whenever possible
	int d = 1;
	while (x > 0xFFFFFFFFull) {
		x /= 10;
		d++;
	}
	uint32_t y = (uint32_t)x;
	while (y >= 10) {
		y /= 10;
		d++;
	}
	return d;
}

__host__ __device__ inline int genome_name_len(int64_t genome)
{
	if (genome < 0) return 9; // "genomeref"
	const int d = dec_digits((uint64_t)genome + 1);
	return 6 + (d < 3 ? 3 : d);
}

__host__ __device__ inline char *put_uint(char *p, uint64_t x, int min_digits)
{
	int d = dec_digits(x);
	if (d < min_digits) d = min_digits;
	int i = d - 1;
	while (x > 0xFFFFFFFFull) {
		p[i--] = (char)('0' + (int)(x % 10));
		x /= 10;
	}
	uint32_t y = (uint32_t)x; // division by the constant 10 is a multiply and a shift here
	for (; i >= 0; i--) {
		const uint32_t q = y / 10u;
		p[i] = (char)('0' + (int)(y - q * 10u));
		y = q;
	}
	return p + d;
}

__host__ __device__ inline char *put_str(char *p, const char *s)
{
	while (*s) *p++ = *s++;
	return p;
}

__host__ __device__ inline char *put_genome_name(char *p, int64_t genome)
{
	p = put_str(p, "genome");
	if (genome < 0) return put_str(p, "ref");
	return put_uint(p, (uint64_t)genome + 1, 3);
}

// Host-side decimal conversion for the PGS_OUT_HOST_EXPAND headers (one per kilobyte of text on sixteen cores:
// the division loops above were as expensive as expanding the record's bases): digit count by comparison,
// two digits per division through a table.  Same characters as put_uint.
static const char kDigitPairs[201] = "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
									 "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
									 "8081828384858687888990919293949596979899";
static inline int host_dec_digits(uint64_t x)
{
	int d = 1;
	while (x >= 10000) {
		x /= 10000;
		d += 4;
	}
	return d + (x >= 1000 ? 3 : x >= 100 ? 2 : x >= 10 ? 1 : 0);
}
static inline char *host_put_uint(char *p, uint64_t x, int min_digits)
{
	int d = host_dec_digits(x);
	if (d < min_digits) d = min_digits;
	char *q = p + d;
	while (x >= 100) {
		const uint64_t r = x % 100;
		x /= 100;
		q -= 2;
		std::memcpy(q, kDigitPairs + 2 * r, 2);
	}
	if (x >= 10) {
		q -= 2;
		std::memcpy(q, kDigitPairs + 2 * x, 2);
	} else {
		*--q = (char)('0' + (int)x);
	}
	while (q > p) *--q = '0';
	return p + d;
}

// four sites (8 bits) -> four ASCII bytes: spread the 2-bit codes to nibbles, then let PRMT pick
// the bytes of "ACGT"
__device__ __forceinline__ uint32_t ascii4(uint32_t bits8)
{
	uint32_t t = (bits8 | (bits8 << 4)) & 0x0F0Fu;
	t = (t | (t << 2)) & 0x3333u;
	return __byte_perm(0x54474341u, 0u, t);
}

// Warp-cooperative: prefix text (prefix_len bytes, staged in shared memory by lane 0), then L
// bases of `row`, then `tail_len` bytes of tail ("\n" or "\n\n").  dst has any alignment; the
// body is written with aligned 16-byte stores.
__device__ void emit_record(char *__restrict__ dst, const char *prefix, int prefix_len,
							const uint32_t *__restrict__ row, int64_t L, int tail_len, int lane)
{
	for (int i = lane; i < prefix_len; i += 32) dst[i] = prefix[i];
	char *seq = dst + prefix_len;
	const int64_t head = min((int64_t)((16 - ((uintptr_t)seq & 15)) & 15), L);
	for (int64_t s = lane; s < head; s += 32) seq[s] = "ACGT"[(row[s >> 4] >> (2 * (s & 15))) & 3u];
	const int64_t chunks = (L - head) / 16;
	const int sh = 2 * (int)head; // bit offset of the first body site inside word 0 (head < 16)
	uint4 *body = reinterpret_cast<uint4 *>(seq + head);
	// two 16-byte chunks per lane and round: all four loads are in flight before the first is used
	// (a 1000-site record is 62 chunks: one round)
	auto expand = [&](int64_t c, uint32_t lo, uint32_t hi) {
		const uint32_t bits = __funnelshift_r(lo, hi, sh);
		uint4 v;
		v.x = ascii4(bits & 0xFFu);
		v.y = ascii4((bits >> 8) & 0xFFu);
		v.z = ascii4((bits >> 16) & 0xFFu);
		v.w = ascii4(bits >> 24);
		body[c] = v;
	};
	for (int64_t c = lane; c < chunks; c += 64) {
		const int64_t c2 = c + 32;
		const bool two = c2 < chunks;
		const uint32_t lo0 = __ldg(row + c);
		const uint32_t hi0 = sh ? __ldg(row + c + 1) : 0u;
		const uint32_t lo1 = two ? __ldg(row + c2) : 0u;
		const uint32_t hi1 = (two && sh) ? __ldg(row + c2 + 1) : 0u;
		expand(c, lo0, hi0);
		if (two) expand(c2, lo1, hi1);
	}
	const int64_t done = head + chunks * 16;
	for (int64_t s = done + lane; s < L; s += 32) seq[s] = "ACGT"[(row[s >> 4] >> (2 * (s & 15))) & 3u];
	if (lane < tail_len) seq[L + lane] = '\n';
}

struct FormatTables {
	const uint4 *rows;
	const int64_t *row_base;     // [N]
	const int32_t *genome_leaf;  // [n]
	const int64_t *genome_count; // [n] genes carried by genome
	const int64_t *sp_off;       // [N+1]
	const int64_t *sp_gene;      // sparse gene table index per sparse row
	const int64_t *sp_digits;    // [sp rows + N] per node ms+1 entries: digits of earlier sparse rows
	const int32_t *dense_slot;   // [G]
	const int64_t *dense_gene;   // [D]
	const int32_t *dense_before; // [G+1]
	const int64_t *dense_digits_before; // [G+1]
	const int64_t *gene_id;      // [G]
	const uint32_t *origin_row;  // [G] row of the gene at its origin (kNoRow: no rows)
	int64_t n_dense, blocks, L;
};

__host__ __device__ inline int64_t sparse_rank(const FormatTables &t, int32_t node, int64_t g)
{
	const int64_t *b = t.sp_gene + t.sp_off[node];
	int64_t lo = 0, hi = t.sp_off[node + 1] - t.sp_off[node];
	while (lo < hi) {
		const int64_t mid = (lo + hi) >> 1;
		if (b[mid] < g)
			lo = mid + 1;
		else
			hi = mid;
	}
	return lo;
}

constexpr int kFmtWarps = 8;
constexpr int kRecPerCta = 32; // records per CTA (64 / 128 / 256 measured: no faster): headers are built by 32 lanes side by side (phase 1),
							   // then every warp expands four of the records (phase 2)

struct Record {
	char *dst;           // start of the record in the staging buffer
	const uint32_t *row; // packed sequence
	int plen, tail;      // header bytes, trailing newlines
};

// Where a record goes and what it holds, the same on the device (K3) and on the host
// (PGS_OUT_HOST_EXPAND): offset inside its file / MAF block, the row of the store, header text.
struct RecordInfo {
	int64_t off; // start of the record inside genomeNNN.fasta / inside the gene's MAF block
	int64_t row; // row of the store that holds the sequence
	int plen, tail; // header bytes, trailing newlines
};

// record `slot` of a genome's leaf region (dense rows first, then the leaf's sparse rows); the
// header ">genomeNNN.<gene id>\n" is written to text
__host__ __device__ inline RecordInfo fasta_record(const FormatTables &t, int32_t genome, int64_t slot, char *text)
{
	const int32_t leaf = t.genome_leaf[genome];
	int64_t g, srank;
	if (slot < t.n_dense) {
		g = t.dense_gene[slot];
		srank = sparse_rank(t, leaf, g);
	} else {
		srank = slot - t.n_dense;
		g = t.sp_gene[t.sp_off[leaf] + srank];
	}
	const int64_t k = t.dense_before[g] + srank;
	const int name_len = genome_name_len(genome);
	char *p = text;
	*p++ = '>';
	p = put_genome_name(p, genome);
	*p++ = '.';
	p = put_uint(p, (uint64_t)t.gene_id[g], 1);
	*p++ = '\n';
	RecordInfo r;
	r.off = k * (name_len + t.L + 4) + t.dense_digits_before[g] + t.sp_digits[t.sp_off[leaf] + leaf + srank];
	r.row = t.row_base[leaf] + slot;
	r.plen = (int)(p - text);
	r.tail = 1;
	return r;
}

// ---- genomeNNN.fasta.  rows_prefix[i] = number of leaf rows of the chunk's genomes before genome
// first_genome + i; file_off[i] = offset of that genome's file in the staging buffer.
__device__ Record genome_fasta_record(const FormatTables &t, char *out, int32_t first_genome, int32_t n_chunk_genomes,
									  const int64_t *__restrict__ rows_prefix, const int64_t *__restrict__ file_off,
									  int64_t item, char *text)
{
	int lo = 0, hi = n_chunk_genomes - 1; // last i with rows_prefix[i] <= item
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (rows_prefix[mid] <= item)
			lo = mid;
		else
			hi = mid - 1;
	}
	const RecordInfo ri = fasta_record(t, first_genome + lo, item - rows_prefix[lo], text);
	Record r;
	r.dst = out + file_off[lo] + ri.off;
	r.row = reinterpret_cast<const uint32_t *>(t.rows + ri.row * t.blocks);
	r.plen = ri.plen;
	r.tail = ri.tail;
	return r;
}

__global__ void __launch_bounds__(kFmtWarps * 32)
k3_genome_fasta(const __grid_constant__ FormatTables t, char *__restrict__ out, int32_t first_genome,
				int32_t n_chunk_genomes, const int64_t *__restrict__ rows_prefix,
				const int64_t *__restrict__ file_off)
{
	__shared__ char s_text[kRecPerCta][64];
	__shared__ Record s_rec[kRecPerCta];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t item0 = (int64_t)blockIdx.x * kRecPerCta, total = rows_prefix[n_chunk_genomes];
	if (threadIdx.x < kRecPerCta && item0 + threadIdx.x < total)
		s_rec[threadIdx.x] = genome_fasta_record(t, out, first_genome, n_chunk_genomes, rows_prefix, file_off,
												 item0 + threadIdx.x, s_text[threadIdx.x]);
	__syncthreads();
	for (int r = warp; r < kRecPerCta && item0 + r < total; r += kFmtWarps)
		emit_record(s_rec[r].dst, s_text[r], s_rec[r].plen, s_rec[r].row, t.L, s_rec[r].tail, lane);
}

// ---- ref.fasta / core.fasta / accessory.fasta: one warp per listed gene, offsets from the host.
__global__ void __launch_bounds__(kFmtWarps * 32)
k3_ref_fasta(const __grid_constant__ FormatTables t, char *__restrict__ out, const int64_t *__restrict__ genes,
			 const int64_t *__restrict__ offs, int64_t n_records)
{
	__shared__ char text[kFmtWarps][64];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t item = (int64_t)blockIdx.x * kFmtWarps + warp;
	if (item >= n_records) return;
	const int64_t g = genes[item];
	int plen = 0;
	if (lane == 0) {
		char *p = text[warp];
		*p++ = '>';
		p = put_genome_name(p, -1);
		*p++ = '.';
		p = put_uint(p, (uint64_t)t.gene_id[g], 1);
		*p++ = '\n';
		plen = (int)(p - text[warp]);
	}
	plen = __shfl_sync(0xFFFFFFFFu, plen, 0);
	__syncwarp();
	const uint32_t *row = reinterpret_cast<const uint32_t *>(t.rows + (int64_t)t.origin_row[g] * t.blocks);
	emit_record(out + offs[item], text[warp], plen, row, t.L, 1, lane);
}

// ---- alignment.maf: one warp per `s` line.  The chunk covers table genes [first, first + n);
// lines_prefix[i] = lines of the chunk's genes before gene i (1 + carriers each, 0 for genes
// without carriers); block_off[i] = offset of the gene's block in the staging buffer.
struct MafTables {
	const int64_t *carrier_off;    // [G+1] sparse genes: into carrier_genome
	const int32_t *carrier_genome;
	const int64_t *carrier_prefix; // [carriers + G] per gene nc+1 entries: sum(name_len + digits(size))
	const int32_t *all_order;      // [n]
	const int64_t *all_prefix;     // [n+1]
	int64_t ref_size;              // (#reference genes) * L
	int32_t n_genomes;
};

// line j of gene g's block (0 = the reference line, which also carries the "a\n"; 1.. = carriers)
__host__ __device__ inline RecordInfo maf_line(const FormatTables &t, const MafTables &m, int64_t g, int64_t j, int64_t lines, char *text)
{
	const uint64_t gid = (uint64_t)t.gene_id[g];
	const bool dense = t.dense_slot[g] >= 0;
	const int idd = dec_digits(gid), ld = dec_digits((uint64_t)t.L);
	const int64_t ref_len = 2 + 9 + 1 + idd + 3 + ld + 3 + dec_digits((uint64_t)m.ref_size) + 1 + t.L + 1;

	RecordInfo r;
	int64_t genome = -1, size;
	if (j == 0) {
		r.off = 0; // "a\n" is part of this line's prefix
		size = m.ref_size;
		r.row = (int64_t)t.origin_row[g];
	} else {
		const int64_t c = j - 1;
		int64_t pre;
		if (dense) {
			genome = m.all_order[c];
			pre = m.all_prefix[c];
		} else {
			genome = m.carrier_genome[m.carrier_off[g] + c];
			pre = m.carrier_prefix[m.carrier_off[g] + g + c];
		}
		r.off = 2 + ref_len + c * (idd + ld + t.L + 11) + pre;
		size = t.genome_count[genome] * t.L;
		const int32_t leaf = t.genome_leaf[genome];
		const int64_t slot = dense ? t.dense_slot[g] : t.n_dense + sparse_rank(t, leaf, g);
		r.row = t.row_base[leaf] + slot;
	}
	char *p = text;
	if (j == 0) p = put_str(p, "a\n");
	p = put_str(p, "s ");
	p = put_genome_name(p, genome);
	*p++ = '.';
	p = put_uint(p, gid, 1);
	p = put_str(p, " 0 ");
	p = put_uint(p, (uint64_t)t.L, 1);
	p = put_str(p, " + ");
	p = put_uint(p, (uint64_t)size, 1);
	*p++ = ' ';
	r.plen = (int)(p - text);
	// the last line of a block is followed by the blank line (std::endl at pangenomesim.cxx:289)
	r.tail = (j == lines - 1) ? 2 : 1;
	return r;
}

// item -> (gene of the chunk, line of its block)
__device__ inline int maf_item_gene(const int64_t *__restrict__ lines_prefix, int32_t n_chunk_genes, int64_t item)
{
	int lo = 0, hi = n_chunk_genes - 1;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (lines_prefix[mid] <= item)
			lo = mid;
		else
			hi = mid - 1;
	}
	return lo;
}

__device__ Record maf_record(const FormatTables &t, const MafTables &m, char *out, int64_t first_gene, int32_t n_chunk_genes,
							 const int64_t *__restrict__ lines_prefix, const int64_t *__restrict__ block_off, int64_t item,
							 char *text)
{
	const int lo = maf_item_gene(lines_prefix, n_chunk_genes, item);
	const RecordInfo ri = maf_line(t, m, first_gene + lo, item - lines_prefix[lo], lines_prefix[lo + 1] - lines_prefix[lo], text);
	Record r;
	r.dst = out + block_off[lo] + ri.off;
	r.row = reinterpret_cast<const uint32_t *>(t.rows + ri.row * t.blocks);
	r.plen = ri.plen;
	r.tail = ri.tail;
	return r;
}

// PGS_OUT_HOST_EXPAND: the rows of a MAF chunk's lines, gathered in line order (one warp per row);
// the host writes the headers and expands the bases.
__global__ void __launch_bounds__(kFmtWarps * 32)
k3_gather_maf_rows(const __grid_constant__ FormatTables t, const __grid_constant__ MafTables m, uint4 *__restrict__ out,
				   int64_t first_gene, int32_t n_chunk_genes, const int64_t *__restrict__ lines_prefix)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t item = (int64_t)blockIdx.x * kFmtWarps + warp, total = lines_prefix[n_chunk_genes];
	if (item >= total) return;
	const int lo = maf_item_gene(lines_prefix, n_chunk_genes, item);
	const int64_t g = first_gene + lo, j = item - lines_prefix[lo];
	int64_t row;
	if (j == 0) {
		row = (int64_t)t.origin_row[g];
	} else {
		const bool dense = t.dense_slot[g] >= 0;
		const int64_t genome = dense ? m.all_order[j - 1] : m.carrier_genome[m.carrier_off[g] + j - 1];
		const int32_t leaf = t.genome_leaf[genome];
		row = t.row_base[leaf] + (dense ? t.dense_slot[g] : t.n_dense + sparse_rank(t, leaf, g));
	}
	const uint4 *src = t.rows + row * t.blocks;
	uint4 *dst = out + item * t.blocks;
	for (int64_t b = lane; b < t.blocks; b += 32) dst[b] = __ldg(src + b);
}

__global__ void __launch_bounds__(kFmtWarps * 32, 8)
k3_maf(const __grid_constant__ FormatTables t, const __grid_constant__ MafTables m, char *__restrict__ out,
	   int64_t first_gene, int32_t n_chunk_genes, const int64_t *__restrict__ lines_prefix,
	   const int64_t *__restrict__ block_off)
{
	__shared__ char s_text[kRecPerCta][96];
	__shared__ Record s_rec[kRecPerCta];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t item0 = (int64_t)blockIdx.x * kRecPerCta, total = lines_prefix[n_chunk_genes];
	if (threadIdx.x < kRecPerCta && item0 + threadIdx.x < total)
		s_rec[threadIdx.x] = maf_record(t, m, out, first_gene, n_chunk_genes, lines_prefix, block_off, item0 + threadIdx.x,
										s_text[threadIdx.x]);
	__syncthreads();
	for (int r = warp; r < kRecPerCta && item0 + r < total; r += kFmtWarps)
		emit_record(s_rec[r].dst, s_text[r], s_rec[r].plen, s_rec[r].row, t.L, s_rec[r].tail, lane);
}

// ---------------------------------------------------------------------------------------------
// host side: tables, chunking, pinned double buffering, sinks

namespace {

struct Pinned {
	char *p = nullptr;
	~Pinned()
	{
		if (p) cudaFreeHost(p);
	}
};

// Small pool of writer threads: a chunk that has arrived in pinned memory is cut into ranges that
// are pwrite()n in parallel (one thread cannot keep up with PCIe: ~3 GB/s into the page cache).
class WritePool {
  public:
	explicit WritePool(unsigned n)
	{
		for (unsigned i = 0; i < n; i++) threads.emplace_back([this] { run(); });
	}
	~WritePool()
	{
		{
			std::lock_guard<std::mutex> l(m);
			stop = true;
		}
		cv.notify_all();
		for (auto &t : threads) t.join();
	}
	void submit(std::function<void()> f)
	{
		{
			std::lock_guard<std::mutex> l(m);
			q.push_back(std::move(f));
			pending++;
		}
		cv.notify_one();
	}
	void wait_all()
	{
		std::unique_lock<std::mutex> l(m);
		done_cv.wait(l, [this] { return pending == 0; });
	}

  private:
	void run()
	{
		for (;;) {
			std::function<void()> f;
			{
				std::unique_lock<std::mutex> l(m);
				cv.wait(l, [this] { return stop || !q.empty(); });
				if (q.empty()) return;
				f = std::move(q.front());
				q.pop_front();
			}
			f();
			{
				std::lock_guard<std::mutex> l(m);
				if (--pending == 0) done_cv.notify_all();
			}
		}
	}
	std::vector<std::thread> threads;
	std::mutex m;
	std::condition_variable cv, done_cv;
	std::deque<std::function<void()>> q;
	int pending = 0;
	bool stop = false;
};

struct Writer {
	Engine &e;
	const pgs_output_spec *spec;
	pgs_output_report rep{};
	std::unique_ptr<WritePool> pool;
	std::mutex err_mutex;
	std::string error;
	int maf_fd = -1; // a file that spans chunks stays open
	std::string maf_name;
	int64_t maf_size = 0;                          // current size of that file
	bool use_mmap = true;                          // PGS_WRITE_MMAP=0: pwrite only
	std::vector<std::pair<void *, size_t>> mapped; // this chunk's mappings of that file

	explicit Writer(Engine &eng, const pgs_output_spec *s) : e(eng), spec(s)
	{
		if (!(s->what & PGS_OUT_DISCARD) && !s->sink) {
			unsigned n = std::thread::hardware_concurrency();
			n = n < 2 ? 1 : (n > 8 ? 8 : n);
			if (const char *v = std::getenv("PGS_WRITE_THREADS")) n = (unsigned)std::max(1, std::atoi(v)); // tuning knob
			if (const char *v = std::getenv("PGS_WRITE_MMAP")) use_mmap = std::atoi(v) != 0;               // tuning knob
			pool.reset(new WritePool(n));
		}
	}
	~Writer() { close_file(); }

	void unmap_all()
	{
		for (auto &m : mapped) ::munmap(m.first, m.second);
		mapped.clear();
	}
	void close_file()
	{
		if (pool) pool->wait_all();
		unmap_all();
		if (maf_fd >= 0) ::close(maf_fd);
		maf_fd = -1;
	}
	void fail(const std::string &msg)
	{
		std::lock_guard<std::mutex> l(err_mutex);
		if (error.empty()) error = msg;
	}
	static bool write_range(int fd, const char *data, int64_t len, int64_t offset, std::string &msg)
	{
		while (len > 0) {
			const ssize_t w = ::pwrite(fd, data, (size_t)len, (off_t)offset);
			if (w < 0) {
				if (errno == EINTR) continue;
				msg = std::strerror(errno);
				return false;
			}
			data += w;
			offset += w;
			len -= w;
		}
		return true;
	}
	// deliver len bytes of file `name` starting at file offset `offset`; `whole` = the piece is the
	// complete file.  File output is asynchronous: finish() waits for the chunk's writes.
	bool put(const std::string &name, int64_t offset, const char *data, int64_t len, bool whole)
	{
		if (offset == 0) rep.files++;
		rep.bytes += len;
		if (spec->what & PGS_OUT_DISCARD) return true;
		if (spec->sink) {
			if (spec->sink(spec->sink_user, name.c_str(), offset, data, len) != 0) {
				error = "sink aborted at " + name;
				return false;
			}
			return true;
		}
		const std::string path = std::string(spec->out_dir) + "/" + name;
		constexpr int64_t kSlice = (int64_t)8 << 20;
		// with a shard layout several engines fill one file: never truncate (the caller created it)
		const int trunc = (offset == 0 && !spec->shard) ? O_TRUNC : 0;
		if ((whole || spec->shard) && len <= 2 * kSlice) { // small piece: one task opens, writes, closes
			pool->submit([this, path, data, len, offset, trunc] {
				const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | trunc, 0644);
				std::string msg;
				if (fd < 0) return fail(path + ": " + std::strerror(errno));
				if (!write_range(fd, data, len, offset, msg)) fail(path + ": " + msg);
				::close(fd);
			});
			return true;
		}
		if (maf_name != name) { // a file that arrives in several pieces / slices
			pool->wait_all();
			unmap_all();
			if (maf_fd >= 0) ::close(maf_fd);
			maf_fd = ::open(path.c_str(), O_RDWR | O_CREAT | trunc, 0644);
			if (maf_fd < 0) {
				error = path + ": " + std::strerror(errno);
				return false;
			}
			maf_name = name;
			struct stat sb;
			maf_size = ::fstat(maf_fd, &sb) == 0 ? (int64_t)sb.st_size : 0;
		}
		// Buffered pwrite()s to ONE file serialise on the inode lock (one writer at a time, ~3 GB/s);
		// page faults on a shared mapping do not.  So the piece's range of the file is mapped and the
		// pool copies slices into the mapping, each thread populating its own pages first
		// (MADV_POPULATE_WRITE reports a full disk as an error where a plain store would SIGBUS).
		char *window = nullptr;
		if (use_mmap && len > 0) {
			// (with a shard layout other engines write to the same file: it is never resized here --
			// an engine with lower offsets would cut off what another one has written; the caller
			// sizes the file beforehand, else the piece goes through pwrite)
			bool sized = maf_size >= offset + len;
			if (!sized && !spec->shard && ::ftruncate(maf_fd, (off_t)(offset + len)) == 0) {
				maf_size = offset + len;
				sized = true;
			}
			if (sized) {
				const int64_t page = (int64_t)::sysconf(_SC_PAGESIZE);
				const int64_t map_start = offset & ~(page - 1);
				const size_t map_len = (size_t)(offset + len - map_start);
				void *m = ::mmap(nullptr, map_len, PROT_READ | PROT_WRITE, MAP_SHARED, maf_fd, (off_t)map_start);
				if (m != MAP_FAILED) {
					mapped.emplace_back(m, map_len);
					window = static_cast<char *>(m) + (offset - map_start);
				}
			}
		}
		for (int64_t at = 0; at < len; at += kSlice) {
			const int64_t now = std::min(kSlice, len - at);
			const int fd = maf_fd;
			if (window) {
				pool->submit([this, fd, path, window, data, at, now, offset] {
					// Storing into a page the file system cannot back raises SIGBUS: the pages are
					// populated first (an error instead of a signal); where that is not possible --
					// no MADV_POPULATE_WRITE, or it fails for any reason -- the slice goes through pwrite.
					bool populated = false;
#ifdef MADV_POPULATE_WRITE
					const uintptr_t page = (uintptr_t)::sysconf(_SC_PAGESIZE);
					const uintptr_t b = (uintptr_t)(window + at) & ~(page - 1);
					const uintptr_t en = ((uintptr_t)(window + at + now) + page - 1) & ~(page - 1);
					populated = ::madvise(reinterpret_cast<void *>(b), (size_t)(en - b), MADV_POPULATE_WRITE) == 0;
#endif
					if (populated) {
						std::memcpy(window + at, data + at, (size_t)now);
						return;
					}
					std::string msg;
					if (!write_range(fd, data + at, now, offset + at, msg)) fail(path + ": " + msg);
				});
				continue;
			}
			pool->submit([this, fd, path, data, at, now, offset] {
				std::string msg;
				if (!write_range(fd, data + at, now, offset + at, msg)) fail(path + ": " + msg);
			});
		}
		return true;
	}
	// all writes of the chunk have reached the kernel; its pinned buffer may be reused
	bool finish()
	{
		if (pool) pool->wait_all();
		unmap_all();
		std::lock_guard<std::mutex> l(err_mutex);
		return error.empty();
	}
};

struct FilePiece {
	std::string name;
	int64_t file_offset; // offset inside the file
	int64_t buf_offset;  // offset inside the staging buffer
	int64_t len;
	bool whole;          // the piece is the complete file
};

} // namespace

// Host-side layout of the text outputs of one engine: digit tables, closed-form offsets, sizes.
struct OutputPlan {
	std::vector<int32_t> dense_before;
	std::vector<int64_t> dense_digits_before, sp_digits, all_prefix, carrier_prefix, genome_count;
	std::vector<uint32_t> origin_row;
	std::vector<int> id_digits;
	std::vector<int64_t> genome_bytes, genome_rows, maf_lines, maf_bytes;
	int64_t ref_size = 0;
};

static int build_output_plan(Engine &e, const pgs_output_spec *spec, OutputPlan &p)
{
	const int64_t G = e.n_genes, D = (int64_t)e.dense_gid.size(), L = e.cfg.gene_length;
	const int32_t N = e.n_nodes, n = e.n_genomes;
	const pgs_shard_layout *shard = spec->shard;
	for (int pass = 0; pass < 2; pass++) {
		const int64_t cnt = pass ? spec->n_ref_acc : spec->n_ref_core;
		const int64_t *lst = pass ? spec->ref_acc : spec->ref_core;
		if (cnt < 0 || (cnt > 0 && !lst)) return e.fail(PGS_ERR_ARG, "pgs_write_outputs: bad reference list");
		for (int64_t i = 0; i < cnt; i++)
			if (lst[i] < 0 || lst[i] >= G) return e.fail(PGS_ERR_ARG, "pgs_write_outputs: reference gene index out of range");
	}
	if (shard && (!shard->genome_offset || !shard->genome_gene_total))
		return e.fail(PGS_ERR_ARG, "pgs_write_outputs: incomplete shard layout");

	p.dense_before.assign(G + 1, 0);
	p.dense_digits_before.assign(G + 1, 0);
	p.origin_row.assign(G, kNoRow);
	p.id_digits.resize(G);
	for (int64_t g = 0; g < G; g++) {
		p.id_digits[g] = dec_digits((uint64_t)e.gene_id[g]);
		const bool dn = e.dense_slot[g] >= 0;
		p.dense_before[g + 1] = p.dense_before[g] + (dn ? 1 : 0);
		p.dense_digits_before[g + 1] = p.dense_digits_before[g] + (dn ? p.id_digits[g] : 0);
		const int64_t r = e.row_index(e.origin[g], g);
		if (r >= 0) p.origin_row[g] = (uint32_t)r;
	}
	p.sp_digits.assign(e.sp_gene.size() + N, 0);
	for (int32_t v = 0; v < N; v++) {
		int64_t acc = 0;
		const int64_t base = e.sp_off[v] + v;
		for (int64_t j = e.sp_off[v]; j < e.sp_off[v + 1]; j++) {
			p.sp_digits[base + (j - e.sp_off[v])] = acc;
			acc += p.id_digits[e.sp_gene[j]];
		}
		p.sp_digits[base + (e.sp_off[v + 1] - e.sp_off[v])] = acc;
	}
	// genes per genome over the whole alignment (all shards): the MAF "source size" field
	p.genome_count = shard ? std::vector<int64_t>(shard->genome_gene_total, shard->genome_gene_total + n) : e.genome_gene_count;
	// (one term per genome, looked up per carrier entry: config 3 has two million of those)
	std::vector<int32_t> term(n);
	for (int32_t i = 0; i < n; i++) term[i] = genome_name_len(i) + dec_digits((uint64_t)(p.genome_count[i] * L));
	auto carrier_term = [&](int32_t genome) { return (int64_t)term[genome]; };
	p.all_prefix.assign(n + 1, 0);
	for (int32_t i = 0; i < n; i++) p.all_prefix[i + 1] = p.all_prefix[i] + carrier_term(e.all_order[i]);
	p.carrier_prefix.assign(e.carrier_genome.size() + G + 1, 0);
	for (int64_t g = 0; g < G; g++) {
		int64_t acc = 0;
		const int64_t base = e.carrier_off[g] + g;
		for (int64_t k = e.carrier_off[g]; k < e.carrier_off[g + 1]; k++) {
			p.carrier_prefix[base + (k - e.carrier_off[g])] = acc;
			acc += carrier_term(e.carrier_genome[k]);
		}
		p.carrier_prefix[base + (e.carrier_off[g + 1] - e.carrier_off[g])] = acc;
	}
	p.ref_size = (shard ? shard->n_ref_total : spec->n_ref_core + spec->n_ref_acc) * L;

	// genome i: every carried gene contributes name_len + L + 4 + digits(gene_id)
	p.genome_bytes.resize(n);
	p.genome_rows.resize(n);
	for (int32_t i = 0; i < n; i++) {
		const int32_t v = e.genome_leaf[i];
		const int64_t ms = e.sp_off[v + 1] - e.sp_off[v];
		p.genome_rows[i] = D + ms;
		p.genome_bytes[i] = (D + ms) * (genome_name_len(i) + L + 4) + p.dense_digits_before[G] + p.sp_digits[e.sp_off[v] + v + ms];
	}
	// MAF block of gene g (only genes carried by at least one genome and having rows)
	const int ld = dec_digits((uint64_t)L);
	const int rsd = dec_digits((uint64_t)p.ref_size);
	p.maf_lines.assign(G, 0);
	p.maf_bytes.assign(G, 0);
	for (int64_t g = 0; g < G; g++) {
		const bool dn = e.dense_slot[g] >= 0;
		const int64_t nc = dn ? n : e.carrier_off[g + 1] - e.carrier_off[g];
		if (nc == 0 || p.origin_row[g] == kNoRow) continue;
		const int64_t pre = dn ? p.all_prefix[n] : p.carrier_prefix[e.carrier_off[g] + g + nc];
		const int64_t ref_len = 2 + 9 + 1 + p.id_digits[g] + 3 + ld + 3 + rsd + 1 + L + 1;
		p.maf_lines[g] = nc + 1;
		p.maf_bytes[g] = 2 + ref_len + nc * (p.id_digits[g] + ld + L + 11) + pre + 1;
	}
	return PGS_OK;
}

static const char kMafHeader[] = "##maf version=1 program=pangenomesim\n"; // pangenomesim.cxx:240

int output_sizes(Engine &e, const pgs_output_spec *spec, int64_t *genome_bytes, int64_t *maf_bytes)
{
	OutputPlan p;
	int rc = build_output_plan(e, spec, p);
	if (rc) return rc;
	if (genome_bytes) std::copy(p.genome_bytes.begin(), p.genome_bytes.end(), genome_bytes);
	if (maf_bytes) {
		int64_t total = (!spec->shard || spec->shard->maf_offset == 0) ? (int64_t)sizeof(kMafHeader) - 1 : 0;
		for (auto b : p.maf_bytes) total += b;
		*maf_bytes = total;
	}
	return PGS_OK;
}

int copy_origin_ascii(Engine &e, int64_t gene_index, char *dst)
{
	if (gene_index < 0 || gene_index >= e.n_genes || !dst) return e.fail(PGS_ERR_ARG, "pgs_copy_origin_ascii: bad argument");
	const int64_t row = e.row_index(e.origin[gene_index], gene_index);
	if (row < 0) return e.fail(PGS_ERR_ARG, "pgs_copy_origin_ascii: the gene has no rows");
	std::vector<uint32_t> words((size_t)e.blocks * 4);
	cudaError_t ce = cudaStreamSynchronize(e.stream);
	if (ce == cudaSuccess)
		ce = cudaMemcpy(words.data(), e.tables.rows + row * e.blocks, (size_t)e.blocks * 16, cudaMemcpyDeviceToHost);
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_copy_origin_ascii");
	for (int64_t s = 0; s < e.cfg.gene_length; s++) dst[s] = "ACGT"[(words[s >> 4] >> (2 * (s & 15))) & 3u];
	return PGS_OK;
}

int write_outputs(Engine &e, const pgs_output_spec *spec, pgs_output_report *report)
{
	const auto wall0 = std::chrono::steady_clock::now();
	const bool timing = std::getenv("PGS_OUT_TIMING") != nullptr; // development knob: phase times on stderr
	auto lap_t0 = wall0;
	auto lap = [&](const char *what) {
		if (!timing) return;
		const auto t1 = std::chrono::steady_clock::now();
		std::fprintf(stderr, "out: %-18s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - lap_t0).count());
		lap_t0 = t1;
	};
	if (!spec->sink && !spec->out_dir && !(spec->what & PGS_OUT_DISCARD))
		return e.fail(PGS_ERR_ARG, "pgs_write_outputs: neither out_dir nor sink");
	const int64_t G = e.n_genes, D = (int64_t)e.dense_gid.size(), L = e.cfg.gene_length;
	const int32_t n = e.n_genomes;
	const pgs_shard_layout *shard = spec->shard;
	if (shard && (spec->what & (PGS_OUT_REF | PGS_OUT_CORE | PGS_OUT_ACCESSORY)))
		return e.fail(PGS_ERR_ARG, "pgs_write_outputs: reference files are written by the host in sharded mode (pgs_copy_origin_ascii)");
	OutputPlan plan;
	int rc = build_output_plan(e, spec, plan);
	if (rc) return rc;
	lap("output plan");
	auto &dense_before = plan.dense_before;
	auto &dense_digits_before = plan.dense_digits_before;
	auto &sp_digits = plan.sp_digits;
	auto &all_prefix = plan.all_prefix;
	auto &carrier_prefix = plan.carrier_prefix;
	auto &origin_row = plan.origin_row;
	auto &id_digits = plan.id_digits;
	auto &genome_bytes = plan.genome_bytes;
	auto &genome_rows = plan.genome_rows;
	auto &maf_lines = plan.maf_lines;
	auto &maf_bytes = plan.maf_bytes;
	const int64_t ref_size = plan.ref_size;

	// the tables live in the engine between calls: uploaded again only when the plan or the caller's
	// totals (shard layout, number of reference genes) have changed
	DevBuf &d_genome_leaf = e.d_fmt[0], &d_genome_count = e.d_fmt[1], &d_sp_off = e.d_fmt[2], &d_sp_gene = e.d_fmt[3],
		   &d_sp_digits = e.d_fmt[4], &d_dense_slot = e.d_fmt[5], &d_dense_gene = e.d_fmt[6], &d_dense_before = e.d_fmt[7],
		   &d_dense_digits_before = e.d_fmt[8], &d_gene_id = e.d_fmt[9], &d_origin_row = e.d_fmt[10], &d_carrier_off = e.d_fmt[11],
		   &d_carrier_genome = e.d_fmt[12], &d_carrier_prefix = e.d_fmt[13], &d_all_order = e.d_fmt[14], &d_all_prefix = e.d_fmt[15];
	uint64_t key = 1469598103934665603ull; // FNV-1a over what the tables depend on besides the plan
	auto mix = [&](uint64_t v) {
		for (int i = 0; i < 8; i++) {
			key ^= (v >> (8 * i)) & 0xFF;
			key *= 1099511628211ull;
		}
	};
	mix(e.plan_serial);
	mix((uint64_t)plan.ref_size);
	for (int64_t c : plan.genome_count) mix((uint64_t)c);
	if (key == 0) key = 1;
	if (e.fmt_key != key) {
		e.fmt_key = 0;
#define UP(buf, vec) \
	if ((rc = e.upload_keep(buf, (vec).data(), sizeof((vec)[0]) * (vec).size(), #vec))) return rc
		UP(d_genome_leaf, e.genome_leaf);
		UP(d_genome_count, plan.genome_count);
		UP(d_sp_off, e.sp_off);
		UP(d_sp_gene, e.sp_gene);
		UP(d_sp_digits, sp_digits);
		UP(d_dense_slot, e.dense_slot);
		UP(d_dense_gene, e.dense_gene);
		UP(d_dense_before, dense_before);
		UP(d_dense_digits_before, dense_digits_before);
		UP(d_gene_id, e.gene_id);
		UP(d_origin_row, origin_row);
		UP(d_carrier_off, e.carrier_off);
		UP(d_carrier_genome, e.carrier_genome);
		UP(d_carrier_prefix, carrier_prefix);
		UP(d_all_order, e.all_order);
		UP(d_all_prefix, all_prefix);
#undef UP
		e.fmt_key = key;
	}
	FormatTables t{};
	t.rows = e.tables.rows;
	t.row_base = e.d_row_base.as<int64_t>();
	t.genome_leaf = d_genome_leaf.as<int32_t>();
	t.genome_count = d_genome_count.as<int64_t>();
	t.sp_off = d_sp_off.as<int64_t>();
	t.sp_gene = d_sp_gene.as<int64_t>();
	t.sp_digits = d_sp_digits.as<int64_t>();
	t.dense_slot = d_dense_slot.as<int32_t>();
	t.dense_gene = d_dense_gene.as<int64_t>();
	t.dense_before = d_dense_before.as<int32_t>();
	t.dense_digits_before = d_dense_digits_before.as<int64_t>();
	t.gene_id = d_gene_id.as<int64_t>();
	t.origin_row = d_origin_row.as<uint32_t>();
	t.n_dense = D;
	t.blocks = e.blocks;
	t.L = L;
	MafTables mt{};
	mt.carrier_off = d_carrier_off.as<int64_t>();
	mt.carrier_genome = d_carrier_genome.as<int32_t>();
	mt.carrier_prefix = d_carrier_prefix.as<int64_t>();
	mt.all_order = d_all_order.as<int32_t>();
	mt.all_prefix = d_all_prefix.as<int64_t>();
	mt.ref_size = ref_size;
	mt.n_genomes = n;

	lap("table upload");
	// ---- staging: two device buffers + two pinned buffers
	int64_t stage = (int64_t)256 << 20; // 64 MB chunks: K3 123 ms per config-4 output, 256 MB: 60 ms (launch-bound below)
	if (const char *v = std::getenv("PGS_STAGE_MB")) stage = std::max<int64_t>(1, std::atol(v)) << 20; // tuning knob
	{
		// A small job does not pin half a gigabyte: cudaHostAlloc costs ~0.3 ms per MB and serialises
		// between the engines of one process (--devices 8 on a 650 MB output: 1 s of 1.1 s per engine).
		// No chunk is larger than everything this call writes.
		int64_t everything = (int64_t)4 << 20;
		for (auto b : genome_bytes) everything += b;
		for (auto b : maf_bytes) everything += b + 64;
		stage = std::min(stage, everything);
	}
	if (spec->what & PGS_OUT_GENOMES)
		for (auto b : genome_bytes) stage = std::max(stage, b);
	if (spec->what & PGS_OUT_MAF)
		for (auto b : maf_bytes) stage = std::max(stage, b + 64);
	if (spec->what & PGS_OUT_PACKED)
		for (int32_t i = 0; i < n; i++) {
			const int32_t v = e.genome_leaf[i];
			stage = std::max(stage, (D + e.sp_off[v + 1] - e.sp_off[v]) * e.blocks * 16);
		}
	{
		int64_t refs = 0;
		for (int pass = 0; pass < 2; pass++) {
			const int64_t cnt = pass ? spec->n_ref_acc : spec->n_ref_core;
			const int64_t *lst = pass ? spec->ref_acc : spec->ref_core;
			for (int64_t i = 0; i < cnt; i++) refs += 9 + id_digits[lst[i]] + L + 4;
		}
		stage = std::max(stage, refs);
	}
	stage = (stage + 255) & ~(int64_t)255;
	// staging buffers (2 device + 2 pinned) are kept by the engine between calls: cudaHostAlloc of
	// tens of megabytes costs milliseconds each time
	DevBuf *const d_aux_a = e.d_fmt_aux_a, *const d_aux_b = e.d_fmt_aux_b;
	DevBuf *d_stage = e.d_stage;
	struct HostView {
		char *p;
	} h_stage[2];
	if (e.stage_bytes < stage) {
		for (int i = 0; i < 2; i++) {
			if (e.h_stage[i]) cudaFreeHost(e.h_stage[i]);
			e.h_stage[i] = nullptr;
			e.d_stage[i].release();
		}
		e.stage_bytes = 0;
	}
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t formatted[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
	cudaEvent_t k0 = nullptr, k1 = nullptr;
	auto cleanup = [&]() {
		for (int i = 0; i < 2; i++) {
			if (formatted[i]) cudaEventDestroy(formatted[i]);
			if (copied[i]) cudaEventDestroy(copied[i]);
			if (consumed[i]) cudaEventDestroy(consumed[i]);
		}
		if (k0) cudaEventDestroy(k0);
		if (k1) cudaEventDestroy(k1);
		if (copy_stream) cudaStreamDestroy(copy_stream);
	};
	const unsigned copy_wait_flag = std::getenv("PGS_SPIN_SYNC") ? 0u : (unsigned)cudaEventBlockingSync; // tuning knob: spin as before
	cudaError_t ce = cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking);
	for (int i = 0; i < 2 && ce == cudaSuccess; i++) {
		if (e.stage_bytes < stage) {
			if ((rc = e.alloc(d_stage[i], (size_t)stage, "format staging"))) {
				cleanup();
				return rc;
			}
			ce = cudaHostAlloc((void **)&e.h_stage[i], (size_t)stage, cudaHostAllocDefault);
			if (ce != cudaSuccess) e.h_stage[i] = nullptr;
		}
		h_stage[i].p = e.h_stage[i];
		if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&formatted[i], cudaEventDisableTiming);
		// (blocking: the thread that waits for a copy sleeps instead of spinning on a core the expander's pool wants)
		if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming | copy_wait_flag);
		if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming);
	}
	if (ce == cudaSuccess) ce = cudaEventCreate(&k0);
	if (ce == cudaSuccess) ce = cudaEventCreate(&k1);
	if (ce != cudaSuccess) {
		cleanup();
		return e.cuda_fail(ce, "pgs_write_outputs setup");
	}
	if (e.stage_bytes < stage) e.stage_bytes = stage;
	// small per-chunk index tables live in preallocated buffers (no cudaMalloc/cudaFree, which
	// would synchronise the device, inside the pipeline)
	const size_t aux_entries = (size_t)std::max<int64_t>(65537, spec->n_ref_core + spec->n_ref_acc + 1);
	for (int i = 0; i < 2; i++) {
		if ((d_aux_a[i].bytes < aux_entries * sizeof(int64_t) && (rc = e.alloc(d_aux_a[i], aux_entries * sizeof(int64_t), "format aux"))) ||
			(d_aux_b[i].bytes < aux_entries * sizeof(int64_t) && (rc = e.alloc(d_aux_b[i], aux_entries * sizeof(int64_t), "format aux")))) {
			cleanup();
			return rc;
		}
	}
	auto h2d = [&](DevBuf &dst, const std::vector<int64_t> &src) -> int {
		cudaError_t err2 = cudaMemcpyAsync(dst.p, src.data(), sizeof(int64_t) * src.size(), cudaMemcpyHostToDevice, e.stream);
		return err2 == cudaSuccess ? 0 : e.cuda_fail(err2, "format aux upload");
	};

	Writer w(e, spec);
	double device_ms = 0.0;
	int64_t d2h_bytes = 0; // what the call copies from the device (text, or packed rows)
	int64_t records = 0;

	// A chunk = kernels writing into d_stage[slot] + the list of file pieces it holds.  The chunk
	// is copied to pinned memory on copy_stream; the previous chunk is handed to the writer while
	// this one is formatted and copied.
	struct Pending {
		bool live = false;
		std::vector<FilePiece> pieces;
		int64_t bytes = 0;
	} pending[2];
	int slot = 0;
	auto drain = [&](int s) -> bool {
		if (!pending[s].live) return true;
		cudaError_t err2 = cudaEventSynchronize(copied[s]);
		if (err2 != cudaSuccess) {
			w.error = std::string("D2H copy: ") + cudaGetErrorString(err2);
			return false;
		}
		for (const auto &pc : pending[s].pieces)
			if (!w.put(pc.name, pc.file_offset, h_stage[s].p + pc.buf_offset, pc.len, pc.whole)) return false;
		pending[s].live = false;
		return w.finish();
	};
	// submit: record kernel-done, copy used bytes, mark pending; then drain the OTHER slot
	// `direct`: instead of the device staging buffer, these (device pointer, offset in the pinned
	// buffer, bytes) ranges of the row store are copied (no kernel ran for the chunk)
	struct Segment {
		const char *src;
		int64_t dst_off, bytes;
	};
	// every CUDA call of the pipeline is checked: the first failure ends the call with its message
	cudaError_t pipe_err = cudaSuccess;
	auto ck = [&](cudaError_t c) {
		if (c != cudaSuccess && pipe_err == cudaSuccess) pipe_err = c;
	};
	auto pipe_ok = [&]() -> bool {
		if (pipe_err == cudaSuccess) return true;
		if (w.error.empty()) w.error = std::string("output pipeline: ") + cudaGetErrorName(pipe_err) + " (" + cudaGetErrorString(pipe_err) + ")";
		return false;
	};
	auto submit = [&](int s, int64_t bytes, std::vector<FilePiece> &&pieces, const std::vector<Segment> *direct = nullptr) -> bool {
		if (!direct) ck(cudaEventRecord(k1, e.stream));
		ck(cudaGetLastError()); // the chunk's kernel launches
		ck(cudaEventRecord(formatted[s], e.stream));
		ck(cudaStreamWaitEvent(copy_stream, formatted[s], 0));
		if (direct) {
			for (const Segment &sg : *direct)
				ck(cudaMemcpyAsync(h_stage[s].p + sg.dst_off, sg.src, (size_t)sg.bytes, cudaMemcpyDeviceToHost, copy_stream)), d2h_bytes += sg.bytes;
		} else {
			ck(cudaMemcpyAsync(h_stage[s].p, d_stage[s].p, (size_t)bytes, cudaMemcpyDeviceToHost, copy_stream)), d2h_bytes += bytes;
		}
		ck(cudaEventRecord(copied[s], copy_stream));
		if (!pipe_ok()) return false;
		pending[s].live = true;
		pending[s].pieces = std::move(pieces);
		pending[s].bytes = bytes;
		if (!drain(s ^ 1)) return false;
		// kernel time of this chunk (the stream is idle by the time the next chunk needs the slot)
		if (!direct && cudaEventSynchronize(k1) == cudaSuccess) {
			float ms = 0;
			if (cudaEventElapsedTime(&ms, k0, k1) == cudaSuccess) device_ms += ms;
		}
		return true;
	};
	auto begin_chunk = [&](int s) {
		// the staging buffer of slot s is free once its previous copy finished
		ck(cudaStreamWaitEvent(e.stream, copied[s], 0));
	};
	auto fail_out = [&](int code, const std::string &msg) {
		cudaStreamSynchronize(e.stream);
		cudaStreamSynchronize(copy_stream);
		cleanup();
		return e.fail(code, msg);
	};
	// make cudaStreamWaitEvent on never-recorded events harmless
	for (int i = 0; i < 2; i++) ck(cudaEventRecord(copied[i], copy_stream));

	// ---- PGS_OUT_HOST_EXPAND: genomeNNN.fasta / alignment.maf from PACKED rows shipped over PCIe
	// (a quarter of the text's bytes) and expanded by the host's cores -- same bytes as K3's.
	const bool host_expand = (spec->what & PGS_OUT_HOST_EXPAND) != 0;
	FormatTables ht = t; // the same tables with host pointers: the record layout code is shared
	MafTables hm = mt;
	ht.rows = nullptr;
	ht.row_base = e.row_base.data();
	ht.genome_leaf = e.genome_leaf.data();
	ht.genome_count = plan.genome_count.data();
	ht.sp_off = e.sp_off.data();
	ht.sp_gene = e.sp_gene.data();
	ht.sp_digits = sp_digits.data();
	ht.dense_slot = e.dense_slot.data();
	ht.dense_gene = e.dense_gene.data();
	ht.dense_before = dense_before.data();
	ht.dense_digits_before = dense_digits_before.data();
	ht.gene_id = e.gene_id.data();
	ht.origin_row = origin_row.data();
	hm.carrier_off = e.carrier_off.data();
	hm.carrier_genome = e.carrier_genome.data();
	hm.carrier_prefix = carrier_prefix.data();
	hm.all_order = e.all_order.data();
	hm.all_prefix = all_prefix.data();
	std::unique_ptr<WritePool> xpool;
	std::atomic<int64_t> hx_files{0}, hx_bytes{0};
	std::mutex hx_sink_mutex;
	const int64_t row_bytes = e.blocks * 16;
	if (host_expand) {
		unsigned nthreads = std::max(2u, std::thread::hardware_concurrency());
		if (const char *v = std::getenv("PGS_EXPAND_THREADS")) nthreads = (unsigned)std::max(1, std::atoi(v)); // tuning knob
		xpool.reset(new WritePool(std::min(nthreads, 128u)));
	}
	// a finished piece of a file, from a worker thread
	auto hx_deliver = [&](const std::string &name, int64_t offset, const char *data, int64_t len, bool first_piece) {
		if (first_piece) hx_files++;
		hx_bytes += len;
		if (spec->what & PGS_OUT_DISCARD) return;
		if (spec->sink) {
			std::lock_guard<std::mutex> l(hx_sink_mutex);
			if (spec->sink(spec->sink_user, name.c_str(), offset, data, len) != 0) w.fail("sink aborted at " + name);
			return;
		}
		const std::string path = std::string(spec->out_dir) + "/" + name;
		const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | ((offset == 0 && !shard) ? O_TRUNC : 0), 0644);
		if (fd < 0) return w.fail(path + ": " + std::strerror(errno));
		std::string msg;
		if (!Writer::write_range(fd, data, len, offset, msg)) w.fail(path + ": " + msg);
		::close(fd);
	};
	// records expanded into a worker's buffer at a time: ~256 KB of text (PGS_HX_RUN_RECORDS: tuning / test knob)
	int64_t kHxRunRecords = std::max<int64_t>(1, ((int64_t)256 << 10) / (e.cfg.gene_length + 24));
	if (const char *v = std::getenv("PGS_HX_RUN_RECORDS")) kHxRunRecords = std::max(1, std::atoi(v));
	// a worker thread's own buffer, kept between tasks (no zero-filling, no reallocation per file)
	auto thread_buffer = [](size_t bytes) -> char * {
		thread_local std::unique_ptr<char[]> buf;
		thread_local size_t cap = 0;
		if (cap < bytes) {
			buf.reset(new char[bytes + bytes / 8]);
			cap = bytes + bytes / 8;
		}
		return buf.get();
	};
	if (host_expand) { // the copies wait for the sweep (and whatever else is queued on the engine's stream)
		ck(cudaEventRecord(formatted[0], e.stream));
		ck(cudaStreamWaitEvent(copy_stream, formatted[0], 0));
	}

	lap("staging, pool");
	// (timing knob) where the host-expanded phases wait: for the copy engine (PCIe-bound) or for the pool (host-bound)
	double wait_copy_ms = 0, wait_pool_ms = 0;
	auto timed = [&](double &acc, auto &&fn) {
		if (!timing) return fn();
		const auto t0 = std::chrono::steady_clock::now();
		auto r = fn();
		acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		return r;
	};
	auto lap_waits = [&](const char *what) {
		if (timing) std::fprintf(stderr, "out:   %s waited %.2f ms for copies, %.2f ms for the pool\n", what, wait_copy_ms, wait_pool_ms);
		wait_copy_ms = wait_pool_ms = 0;
	};
	// ---- ref.fasta, core.fasta, accessory.fasta
	struct RefFile {
		uint32_t flag;
		const char *name;
		std::vector<int64_t> genes;
	} ref_files[3] = {{PGS_OUT_REF, "ref.fasta", {}}, {PGS_OUT_CORE, "core.fasta", {}}, {PGS_OUT_ACCESSORY, "accessory.fasta", {}}};
	ref_files[0].genes.assign(spec->ref_core, spec->ref_core + spec->n_ref_core);
	ref_files[0].genes.insert(ref_files[0].genes.end(), spec->ref_acc, spec->ref_acc + spec->n_ref_acc);
	ref_files[1].genes.assign(spec->ref_core, spec->ref_core + spec->n_ref_core);
	ref_files[2].genes.assign(spec->ref_acc, spec->ref_acc + spec->n_ref_acc);
	for (auto &rf : ref_files) {
		if (!(spec->what & rf.flag)) continue;
		std::vector<int64_t> offs(rf.genes.size());
		int64_t total = 0;
		for (size_t i = 0; i < rf.genes.size(); i++) {
			if (origin_row[rf.genes[i]] == kNoRow) return fail_out(PGS_ERR_ARG, "pgs_write_outputs: reference gene has no rows");
			offs[i] = total;
			total += 9 + id_digits[rf.genes[i]] + L + 4;
		}
		begin_chunk(slot);
		ck(cudaEventRecord(k0, e.stream));
		if (!rf.genes.empty()) {
			if ((rc = h2d(d_aux_a[slot], rf.genes)) || (rc = h2d(d_aux_b[slot], offs))) {
				cleanup();
				return rc;
			}
			const int64_t nrec = (int64_t)rf.genes.size();
			ck(cudaEventRecord(k0, e.stream));
			k3_ref_fasta<<<(unsigned)((nrec + kFmtWarps - 1) / kFmtWarps), kFmtWarps * 32, 0, e.stream>>>(
				t, d_stage[slot].as<char>(), d_aux_a[slot].as<int64_t>(), d_aux_b[slot].as<int64_t>(), nrec);
			records += nrec;
		}
		std::vector<FilePiece> pieces{{rf.name, 0, 0, total, true}};
		if (!submit(slot, total, std::move(pieces))) return fail_out(PGS_ERR_IO, w.error);
		slot ^= 1;
	}

	// the host-expanded outputs use the pinned buffers themselves: what the reference files left in
	// them is delivered first
	if (host_expand && (!drain(0) || !drain(1))) return fail_out(PGS_ERR_IO, w.error);

	// ---- both texts from ONE copy of the rows (PGS_HX_ONE_COPY=1; off by default).  genomeNNN.fasta wants the rows
	// genome by genome, alignment.maf gene by gene; they are shipped twice (25.6 GB for config 4).  With the knob, when
	// both are asked for and the text goes to memory (sink / PGS_OUT_DISCARD), the rows cross PCIe once, in MAF line
	// order (k3_gather_maf_rows), and every chunk of genes yields its MAF blocks AND, for every genome, the piece of
	// genomeNNN.fasta that holds the chunk's genes (records follow the gene table in the file, so the piece is one
	// contiguous range).  Measured on the 16-core GPU box (config 4): 0.676 s per step against 0.538 with two copies --
	// the call is bound by the host's cores once PCIe carries a quarter of the text, and a genome's rows of a chunk lie
	// 2.5 MB apart (one 4 KB page per gene and group of genomes; 25 % slower per core in scripts/expand_bench.cpp's
	// strided variant).  It pays where PCIe is slower or the host has more cores per GPU.
	const bool fused = host_expand && (spec->what & PGS_OUT_GENOMES) && (spec->what & PGS_OUT_MAF) && !shard &&
					   ((spec->what & PGS_OUT_DISCARD) || spec->sink) && std::getenv("PGS_HX_ONE_COPY") && std::atoi(std::getenv("PGS_HX_ONE_COPY"));
	std::vector<int32_t> order_pos;  // genome -> its position in all_order (= its line in a dense gene's MAF block, less one)
	std::vector<int32_t> sp_line;    // per sparse row of a LEAF: the genome's index in the gene's carrier list
	std::vector<int64_t> dense_by_gene((size_t)ht.n_dense); // dense slots in the order of the gene table (= of the files)
	if (host_expand && (spec->what & PGS_OUT_GENOMES)) {
		for (int64_t sl = 0; sl < ht.n_dense; sl++) dense_by_gene[(size_t)sl] = sl;
		std::sort(dense_by_gene.begin(), dense_by_gene.end(), [&](int64_t a, int64_t b) { return ht.dense_gene[a] < ht.dense_gene[b]; });
	}
	if (fused) {
		order_pos.assign(n, 0);
		for (int32_t i = 0; i < n; i++) order_pos[e.all_order[i]] = i;
		sp_line.assign(e.sp_gene.size(), 0);
		const int64_t per_task = 4096;
		for (int64_t g0 = 0; g0 < G; g0 += per_task)
			xpool->submit([&, g0] {
				for (int64_t g = g0; g < std::min(G, g0 + per_task); g++) {
					if (e.dense_slot[g] >= 0) continue;
					for (int64_t k = e.carrier_off[g]; k < e.carrier_off[g + 1]; k++) {
						const int32_t leaf = e.genome_leaf[e.carrier_genome[k]];
						sp_line[(size_t)(e.sp_off[leaf] + sparse_rank(ht, leaf, g))] = (int32_t)(k - e.carrier_off[g]);
					}
				}
			});
		xpool->wait_all();
	}

	lap("reference files");
	// ---- genomeNNN.fasta, host-expanded: a genome's rows lie together in the store (its leaf's
	// region), so a chunk is a handful of plain copies; one task per genome writes the whole file
	if ((spec->what & PGS_OUT_GENOMES) && host_expand && !fused) {
		struct Chunk {
			int32_t i = 0, j = 0;
			int slot = 0;
			std::vector<int64_t> region_off;
		};
		const char *store = reinterpret_cast<const char *>(e.tables.rows);
		auto issue = [&](int32_t i, int s) {
			Chunk c;
			c.i = i;
			c.slot = s;
			int64_t bytes = 0;
			int32_t j = i;
			const char *run_src = nullptr;
			int64_t run_dst = 0, run_len = 0;
			auto flush = [&]() {
				if (run_len) ck(cudaMemcpyAsync(h_stage[s].p + run_dst, run_src, (size_t)run_len, cudaMemcpyDeviceToHost, copy_stream)), d2h_bytes += run_len;
				run_len = 0;
			};
			while (j < n) {
				const int64_t gb = genome_rows[j] * row_bytes;
				if (bytes + gb > stage && j > i) break;
				const char *src = store + e.row_base[e.genome_leaf[j]] * row_bytes;
				c.region_off.push_back(bytes);
				if (run_len && run_src + run_len == src) {
					run_len += gb; // neighbours in the store: one copy
				} else {
					flush();
					run_src = src;
					run_dst = bytes;
					run_len = gb;
				}
				bytes += gb;
				j++;
			}
			flush();
			ck(cudaEventRecord(copied[s], copy_stream));
			c.j = j;
			return c;
		};
		Chunk cur = issue(0, slot);
		while (cur.i < n) {
			Chunk next;
			next.i = n;
			if (cur.j < n) {
				timed(wait_pool_ms, [&] { xpool->wait_all(); return 0; }); // the tasks that read the other pinned buffer are done
				next = issue(cur.j, cur.slot ^ 1);
			}
			ck(timed(wait_copy_ms, [&] { return cudaEventSynchronize(copied[cur.slot]); }));
			if (!pipe_ok()) return fail_out(PGS_ERR_CUDA, w.error);
			for (int32_t g = cur.i; g < cur.j; g++) {
				const char *region = h_stage[cur.slot].p + cur.region_off[g - cur.i];
				xpool->submit([&, g, region] {
					const int64_t bytes = genome_bytes[g], nrows = genome_rows[g], L = ht.L;
					char name[64];
					const int nl = format_genome_name(g, name, sizeof name);
					const std::string file = std::string(name, (size_t)nl) + ".fasta";
					const int64_t at = shard ? shard->genome_offset[g] : 0;
					// The file is written in runs of kHxRunRecords records through a buffer that stays in the
					// core's L2: storing a whole 5 MB file first streams it through the L3 / DRAM at about half
					// the rate (scripts/expand_bench.cpp: 11.5 against 22.7 GB/s of text per core on the GPU box).
					// Records follow the gene table in the file, rows lie dense genes first (by gene id), then
					// the leaf's sparse genes: the two lists, each ascending in the table index, are merged so
					// that a run of records is one contiguous range of the file.
					const int32_t leaf = ht.genome_leaf[g];
					const int64_t *sp = ht.sp_gene + ht.sp_off[leaf];
					const int64_t n_sp = ht.sp_off[leaf + 1] - ht.sp_off[leaf];
					int64_t di = 0, si = 0; // next dense slot in table order, next sparse row
					// The record's place in the file as fasta_record computes it, without its search: while the two
					// lists are merged, si IS the number of the leaf's sparse genes before the record's gene.
					char text[64];
					text[0] = '>';
					char *const after_name = put_genome_name(text + 1, g);
					*after_name = '.';
					const int64_t per_record = genome_name_len(g) + L + 4;
					const int64_t *const leaf_digits = ht.sp_digits + ht.sp_off[leaf] + leaf;
					int64_t done = 0, run_begin = 0;
					do {
						const int64_t now = std::min(nrows - done, kHxRunRecords);
						int64_t run_len = 0;
						char *buf = nullptr;
						for (int64_t k = 0; k < now; k++) {
							int64_t sl, gene;
							const int64_t srank = si;
							if (si >= n_sp || (di < ht.n_dense && ht.dense_gene[dense_by_gene[di]] < sp[si])) {
								sl = dense_by_gene[di++];
								gene = ht.dense_gene[sl];
							} else {
								sl = ht.n_dense + si;
								gene = sp[si++];
							}
							char *const text_end = host_put_uint(after_name + 1, (uint64_t)ht.gene_id[gene], 1);
							*text_end = '\n';
							const int plen = (int)(text_end + 1 - text);
							const int64_t off = (ht.dense_before[gene] + srank) * per_record + ht.dense_digits_before[gene] + leaf_digits[srank];
							if (k == 0) { // (records of a run have the same length up to the digits of the gene id: a bound)
								run_begin = off;
								buf = thread_buffer((size_t)(now * (plen + 16 + L + 1)));
							}
							char *d = buf + (off - run_begin);
							std::memcpy(d, text, (size_t)plen);
							expand_bases(region + sl * row_bytes, L, d + plen);
							d[plen + L] = '\n';
							run_len = off - run_begin + plen + L + 1;
						}
						hx_deliver(file, at + run_begin, buf, run_len, at == 0 && done == 0);
						done += now;
					} while (done < nrows);
					(void)bytes;
				});
			}
			for (int32_t g = cur.i; g < cur.j; g++) records += genome_rows[g];
			cur = std::move(next);
		}
		timed(wait_pool_ms, [&] { xpool->wait_all(); return 0; });
		lap_waits("genome files");
		slot = 0;
		{
			std::lock_guard<std::mutex> l(w.err_mutex);
			if (!w.error.empty()) return fail_out(PGS_ERR_IO, w.error);
		}
	}

	// ---- genomeNNN.fasta
	if ((spec->what & PGS_OUT_GENOMES) && !host_expand) {
		int32_t i = 0;
		while (i < n) {
			int32_t j = i;
			int64_t bytes = 0;
			std::vector<int64_t> rows_prefix{0}, file_off;
			std::vector<FilePiece> pieces;
			while (j < n && bytes + genome_bytes[j] <= stage && (j - i) < 65536) {
				char name[64];
				format_genome_name(j, name, sizeof name);
				file_off.push_back(bytes);
				pieces.push_back({std::string(name) + ".fasta", shard ? shard->genome_offset[j] : 0, bytes, genome_bytes[j], !shard});
				bytes += genome_bytes[j];
				rows_prefix.push_back(rows_prefix.back() + genome_rows[j]);
				j++;
			}
			begin_chunk(slot);
			if ((rc = h2d(d_aux_a[slot], rows_prefix)) || (rc = h2d(d_aux_b[slot], file_off))) {
				cleanup();
				return rc;
			}
			const int64_t items = rows_prefix.back();
			ck(cudaEventRecord(k0, e.stream));
			if (items > 0)
				k3_genome_fasta<<<(unsigned)((items + kRecPerCta - 1) / kRecPerCta), kFmtWarps * 32, 0, e.stream>>>(
					t, d_stage[slot].as<char>(), i, j - i, d_aux_a[slot].as<int64_t>(), d_aux_b[slot].as<int64_t>());
			records += items;
			if (!submit(slot, bytes, std::move(pieces))) return fail_out(PGS_ERR_IO, w.error);
			slot ^= 1;
			i = j;
		}
	}

	// (fused: genomes without a gene still get their empty file; the FASTA records are counted here)
	if (fused) {
		for (int32_t g = 0; g < n; g++) {
			records += genome_rows[g];
			if (genome_rows[g] == 0) {
				char name[64];
				const int nl = format_genome_name(g, name, sizeof name);
				hx_deliver(std::string(name, (size_t)nl) + ".fasta", 0, nullptr, 0, true);
			}
		}
	}

	lap("genome files");
	// ---- alignment.maf, host-expanded: a kernel gathers the rows of a chunk's lines in line order,
	// the host writes the headers and expands the bases -- into the mapped file, into a buffer handed
	// to the sink, or (PGS_OUT_DISCARD) block by block into the workers' own buffers
	if ((spec->what & PGS_OUT_MAF) && host_expand) {
		const bool with_header = !shard || shard->maf_offset == 0;
		const int64_t hlen = with_header ? (int64_t)sizeof(kMafHeader) - 1 : 0;
		const int64_t file_begin = shard ? shard->maf_offset : 0;
		int64_t maf_total = hlen;
		for (auto b : maf_bytes) maf_total += b;
		const bool to_file = !(spec->what & PGS_OUT_DISCARD) && !spec->sink;
		const bool to_sink = !(spec->what & PGS_OUT_DISCARD) && spec->sink;
		int fd = -1;
		const std::string path = to_file ? std::string(spec->out_dir) + "/alignment.maf" : std::string();
		if (to_file) {
			fd = ::open(path.c_str(), O_RDWR | O_CREAT | (shard ? 0 : O_TRUNC), 0644);
			if (fd < 0) return fail_out(PGS_ERR_IO, path + ": " + std::strerror(errno));
			// one engine: the file gets its final size here; several: the caller has sized it
			if (!shard && ::ftruncate(fd, (off_t)maf_total) != 0) {
				const std::string msg = path + ": " + std::strerror(errno);
				::close(fd);
				return fail_out(PGS_ERR_IO, msg);
			}
		}
		// a file that is not large enough yet (several engines, not sized by the caller) is never
		// resized here -- another engine may be writing to it -- and never mapped: pwrite extends it
		bool can_map = to_file;
		if (to_file && shard) {
			struct stat sb;
			can_map = ::fstat(fd, &sb) == 0 && (int64_t)sb.st_size >= file_begin + maf_total;
		}
		if (const char *v = std::getenv("PGS_WRITE_MMAP")) can_map = can_map && std::atoi(v) != 0; // tuning knob
		int64_t kMafTaskLines = 1024;
		if (const char *v = std::getenv("PGS_HX_TASK_LINES")) kMafTaskLines = std::max(1, std::atoi(v)); // tuning knob
		int64_t packed_cap = std::min<int64_t>(stage, (int64_t)64 << 20); // 64 MB of rows = ~260 MB of text per chunk
		if (fused) packed_cap = stage; // (the pieces of the FASTA files grow with the chunk: ~100 records each at 256 MB)
		if (const char *v = std::getenv("PGS_HX_CHUNK_MB")) packed_cap = std::min<int64_t>(stage, std::max<int64_t>(1, std::atol(v)) << 20); // tuning knob
		int64_t largest = 0;
		for (int64_t g = 0; g < G; g++) largest = std::max(largest, maf_lines[g] * row_bytes);
		packed_cap = std::max(packed_cap, largest);
		if (packed_cap > stage) {
			if (fd >= 0) ::close(fd);
			return fail_out(PGS_ERR_NOMEM, "pgs_write_outputs: a gene's rows do not fit the staging buffer");
		}
		// A ring of slots.  With two slots and a barrier between chunks the copy engine and the pool took turns: the
		// next copy was issued only once the tasks of the chunk before had finished (config 4: 1.5 ms per 64 MB chunk
		// against 1.2 ms of copy).  Each staging buffer is cut into up to four slots (rows, then the chunk's
		// lines_prefix table on the device side); a slot is busy from the issue of its copy until the chunk's last task
		// is done, chunks retire in order.  Text handed to a sink is built in one buffer per slot: two slots there.
		const int64_t aux_room = (int64_t)64 << 10; // lines_prefix of at most 8191 genes
		auto stride_of = [&](int per_buffer) { return (stage / per_buffer) & ~(int64_t)255; }; // (16-byte stores, 8-byte table)
		int spb = 1;
		if (!fused && !to_sink)
			for (spb = 4; spb > 1; spb--) {
				const int64_t cap = stride_of(spb) - aux_room;
				if (cap >= largest && cap >= ((int64_t)8 << 20)) break;
			}
		if (const char *v = std::getenv("PGS_HX_RING")) // tuning / test knob: slots per staging buffer
			if (!to_sink) spb = std::max(1, std::min(4, std::atoi(v)));
		if (spb > 1) {
			const int64_t cap = stride_of(spb) - aux_room;
			if (cap < largest || cap < row_bytes) spb = 1; else packed_cap = std::min(packed_cap, cap);
		}
		const int R = 2 * spb;
		const int64_t slot_stride = spb > 1 ? stride_of(spb) : 0;
		const int64_t max_chunk_genes = spb > 1 ? aux_room / 8 - 1 : 65536;
		struct Chunk {
			int64_t g = 0, g2 = 0, bytes = 0, file_pos = 0;
			int slot = 0;
			bool first = false;
			std::vector<int64_t> lines_prefix, block_off;
			char *h_rows = nullptr; // the slot's rows in pinned memory
			char *window = nullptr; // where the chunk's text goes (mapped file / sink buffer), nullptr: workers' own buffers
			void *map = nullptr;
			size_t map_len = 0;
			std::atomic<int64_t> pending{0}; // tasks of the chunk that have not finished
		};
		struct RingEvents { // one pair per slot, destroyed on every way out
			std::vector<cudaEvent_t> formatted, copied;
			~RingEvents()
			{
				for (auto x : formatted) if (x) cudaEventDestroy(x);
				for (auto x : copied) if (x) cudaEventDestroy(x);
			}
		} ring;
		ring.formatted.assign(R, nullptr);
		ring.copied.assign(R, nullptr);
		for (int i = 0; i < R; i++) {
			ck(cudaEventCreateWithFlags(&ring.formatted[i], cudaEventDisableTiming));
			ck(cudaEventCreateWithFlags(&ring.copied[i], cudaEventDisableTiming | copy_wait_flag));
		}
		if (!pipe_ok()) {
			if (fd >= 0) ::close(fd);
			return fail_out(PGS_ERR_CUDA, w.error);
		}
		std::mutex ring_m;
		std::condition_variable ring_cv;
		auto task_done = [&](Chunk *c) {
			if (c->pending.fetch_sub(1) == 1) {
				std::lock_guard<std::mutex> l(ring_m);
				ring_cv.notify_all();
			}
		};
		auto submit_for = [&](Chunk *c, std::function<void()> fn) {
			c->pending.fetch_add(1);
			xpool->submit([&task_done, c, fn = std::move(fn)] {
				fn();
				task_done(c);
			});
		};
		auto wait_chunk = [&](Chunk &c) {
			std::unique_lock<std::mutex> l(ring_m);
			ring_cv.wait(l, [&] { return c.pending.load() == 0; });
		};
		std::vector<char> sink_buf[2];
		auto issue = [&](int64_t g, int s, bool first, int64_t file_pos) {
			std::unique_ptr<Chunk> cp(new Chunk()); // on the heap: the chunk's tasks keep a pointer to it
			Chunk &c = *cp;
			c.g = g;
			c.slot = s;
			c.first = first;
			c.file_pos = file_pos;
			c.bytes = first ? hlen : 0;
			c.lines_prefix.push_back(0);
			int64_t g2 = g, packed = 0;
			while (g2 < G && packed + maf_lines[g2] * row_bytes <= packed_cap && (g2 - g) < max_chunk_genes) {
				c.block_off.push_back(c.bytes);
				c.bytes += maf_bytes[g2];
				packed += maf_lines[g2] * row_bytes;
				c.lines_prefix.push_back(c.lines_prefix.back() + maf_lines[g2]);
				g2++;
			}
			c.g2 = g2;
			const int64_t items = c.lines_prefix.back();
			// the slot: a whole staging buffer (its table in d_aux_a), or a quarter of one with the table behind the rows
			const int buf = s / spb;
			const int64_t at = (int64_t)(s % spb) * slot_stride;
			char *const d_rows = d_stage[buf].as<char>() + at;
			c.h_rows = h_stage[buf].p + at;
			int64_t *const d_prefix = spb > 1 ? reinterpret_cast<int64_t *>(d_rows + (slot_stride - aux_room)) : d_aux_a[s].as<int64_t>();
			if (items > 0) {
				ck(cudaMemcpyAsync(d_prefix, c.lines_prefix.data(), sizeof(int64_t) * c.lines_prefix.size(), cudaMemcpyHostToDevice, e.stream));
				ck(cudaEventRecord(k0, e.stream));
				k3_gather_maf_rows<<<(unsigned)((items + kFmtWarps - 1) / kFmtWarps), kFmtWarps * 32, 0, e.stream>>>(
					t, mt, reinterpret_cast<uint4 *>(d_rows), g, (int32_t)(g2 - g), d_prefix);
				ck(cudaGetLastError());
				ck(cudaEventRecord(k1, e.stream));
			}
			ck(cudaEventRecord(ring.formatted[s], e.stream));
			ck(cudaStreamWaitEvent(copy_stream, ring.formatted[s], 0));
			if (items > 0) ck(cudaMemcpyAsync(c.h_rows, d_rows, (size_t)(items * row_bytes), cudaMemcpyDeviceToHost, copy_stream)), d2h_bytes += items * row_bytes;
			ck(cudaEventRecord(ring.copied[s], copy_stream));
			return cp;
		};
		auto finish_chunk = [&](Chunk &c) { // all of the chunk's tasks are done
			if (c.map) ::munmap(c.map, c.map_len);
			c.map = nullptr;
			if (c.first && with_header) hx_files++;
			if (to_sink && c.bytes > 0) {
				if (spec->sink(spec->sink_user, "alignment.maf", c.file_pos, c.window, c.bytes) != 0) w.fail("sink aborted at alignment.maf");
				hx_bytes += c.bytes;
			}
		};
		// the staging buffers are free: the reference files were drained, the genome files' tasks have finished
		ck(cudaStreamWaitEvent(e.stream, copied[0], 0));
		ck(cudaStreamWaitEvent(e.stream, copied[1], 0));
		std::deque<std::unique_ptr<Chunk>> copying, expanding; // copy issued / tasks submitted, both in chunk order
		auto retire_front = [&]() {
			timed(wait_pool_ms, [&] { wait_chunk(*expanding.front()); return 0; });
			finish_chunk(*expanding.front());
			expanding.pop_front();
		};
		int64_t g_next = 0, pos_next = file_begin;
		int slot_next = 0;
		bool first_chunk = true, gathered = false;
		for (;;) {
			while ((first_chunk || g_next < G) && (int)(copying.size() + expanding.size()) < R) {
				std::unique_ptr<Chunk> c = issue(g_next, slot_next, first_chunk, pos_next);
				slot_next = (slot_next + 1) % R;
				g_next = c->g2;
				pos_next += c->bytes;
				first_chunk = false;
				gathered = gathered || c->lines_prefix.back() > 0;
				copying.push_back(std::move(c));
			}
			if (copying.empty()) break;
			std::unique_ptr<Chunk> cur = std::move(copying.front());
			copying.pop_front();
			ck(timed(wait_copy_ms, [&] { return cudaEventSynchronize(ring.copied[cur->slot]); }));
			if (!pipe_ok()) {
				xpool->wait_all(); // (tasks of earlier chunks point into them)
				for (auto &c : expanding)
					if (c->map) ::munmap(c->map, c->map_len);
				if (fd >= 0) ::close(fd);
				return fail_out(PGS_ERR_CUDA, w.error);
			}
			// where the text goes
			if (can_map && cur->bytes > 0) {
				const int64_t page = (int64_t)::sysconf(_SC_PAGESIZE);
				const int64_t map_start = cur->file_pos & ~(page - 1);
				cur->map_len = (size_t)(cur->file_pos + cur->bytes - map_start);
				cur->map = ::mmap(nullptr, cur->map_len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)map_start);
				if (cur->map == MAP_FAILED) {
					const std::string msg = path + ": mmap: " + std::strerror(errno);
					cur->map = nullptr;
					xpool->wait_all();
					for (auto &c : expanding)
						if (c->map) ::munmap(c->map, c->map_len);
					::close(fd);
					return fail_out(PGS_ERR_IO, msg);
				}
				cur->window = static_cast<char *>(cur->map) + (cur->file_pos - map_start);
			} else if (to_sink) {
				sink_buf[cur->slot].resize((size_t)cur->bytes);
				cur->window = sink_buf[cur->slot].data();
			}
			if (cur->first && hlen > 0) {
				if (cur->window) {
					std::memcpy(cur->window, kMafHeader, (size_t)hlen);
				} else if (to_file) {
					std::string msg;
					if (!Writer::write_range(fd, kMafHeader, hlen, cur->file_pos, msg)) w.fail(path + ": " + msg);
				}
				if (!to_sink) hx_bytes += hlen;
			}
			// tasks: up to kMafTaskLines lines of one gene's block, written in runs of at most 256 lines (a run is what
			// passes through a worker's L2-sized buffer).  One task per run made the submitting thread the bottleneck
			// of the phase: 1000 submissions per 64 MB chunk, 0.8 ms of the chunk's 1.5 (the copy takes 1.2).
			Chunk *const cp = cur.get();
			cp->pending.fetch_add(1); // held until every task has been submitted
			for (int64_t k = 0; k < cur->g2 - cur->g; k++) {
				const int64_t lines = cur->lines_prefix[k + 1] - cur->lines_prefix[k];
				for (int64_t t0 = 0; t0 < lines; t0 += kMafTaskLines) {
					const int64_t t1 = std::min(lines, t0 + kMafTaskLines);
					submit_for(cp, [&, cp, k, lines, t0, t1] {
						for (int64_t j0 = t0; j0 < t1; j0 += 256) {
						const int64_t j1 = std::min(t1, j0 + 256);
						const int64_t g = cp->g + k, L = ht.L;
						const char *rows_in = cp->h_rows + (cp->lines_prefix[k] + j0) * row_bytes;
						char text[96];
						// line offsets grow with j: the run [j0, j1) is one contiguous range of the block
						const int64_t run_begin = maf_line(ht, hm, g, j0, lines, text).off;
						int64_t run_end;
						{
							const RecordInfo last = maf_line(ht, hm, g, j1 - 1, lines, text);
							run_end = last.off + last.plen + L + last.tail;
						}
						// dst of block offset x = base + (x - base_off)
						char *base = cp->window ? cp->window + cp->block_off[k] : nullptr;
						int64_t base_off = 0;
						bool spill = to_file && !cp->map; // the run is built in the worker's own buffer and written with pwrite
						if (cp->map) { // pages of a mapped file are populated before they are stored to (full disk: an error, not SIGBUS)
							bool populated = false;
#ifdef MADV_POPULATE_WRITE
							const uintptr_t page = (uintptr_t)::sysconf(_SC_PAGESIZE);
							const uintptr_t b = (uintptr_t)(base + run_begin) & ~(page - 1);
							const uintptr_t en = ((uintptr_t)(base + run_end) + page - 1) & ~(page - 1);
							populated = ::madvise(reinterpret_cast<void *>(b), (size_t)(en - b), MADV_POPULATE_WRITE) == 0;
#endif
							spill = !populated;
						}
						if (!base || spill) {
							base = thread_buffer((size_t)(run_end - run_begin));
							base_off = run_begin;
						}
						// what the block's carrier lines share (maf_line's layout): "s genomeNNN" + ".<id> 0 <L> + " + size + ' '
						char mid[48];
						int mid_len;
						{
							char *q = mid;
							*q++ = '.';
							q = host_put_uint(q, (uint64_t)ht.gene_id[g], 1);
							q = put_str(q, " 0 ");
							q = host_put_uint(q, (uint64_t)L, 1);
							q = put_str(q, " + ");
							mid_len = (int)(q - mid);
						}
						const bool dense = ht.dense_slot[g] >= 0;
						const int idd = host_dec_digits((uint64_t)ht.gene_id[g]), ld = host_dec_digits((uint64_t)L);
						const int64_t ref_len = 2 + 9 + 1 + idd + 3 + ld + 3 + host_dec_digits((uint64_t)hm.ref_size) + 1 + L + 1;
						const int64_t per_line = idd + ld + L + 11;
						const int32_t *const genomes = dense ? hm.all_order : hm.carrier_genome + hm.carrier_off[g];
						const int64_t *const prefix = dense ? hm.all_prefix : hm.carrier_prefix + hm.carrier_off[g] + g;
						std::memcpy(text, "s genome", 8);
						for (int64_t j = j0; j < j1; j++) {
							int64_t off;
							int plen;
							if (j == 0) { // the reference line (with the block's "a\n")
								const RecordInfo ri = maf_line(ht, hm, g, 0, lines, text);
								off = ri.off;
								plen = ri.plen;
							} else {
								const int64_t c = j - 1;
								const int32_t genome = genomes[c];
								char *q = host_put_uint(text + 8, (uint64_t)genome + 1, 3);
								std::memcpy(q, mid, (size_t)mid_len);
								q = host_put_uint(q + mid_len, (uint64_t)(ht.genome_count[genome] * L), 1);
								*q++ = ' ';
								plen = (int)(q - text);
								off = 2 + ref_len + c * per_line + prefix[c];
							}
							char *d = base + (off - base_off);
							std::memcpy(d, text, (size_t)plen);
							if (j == 0) std::memcpy(text, "s genome", 8); // (maf_line wrote "a\ns genomeref..." there)
							expand_bases(rows_in + (j - j0) * row_bytes, L, d + plen);
							d[plen + L] = '\n';
							if (j == lines - 1) d[plen + L + 1] = '\n'; // the blank line behind a block (std::endl at pangenomesim.cxx:289)
						}
						if (!to_sink) hx_bytes += run_end - run_begin;
						if (spill) {
							std::string msg;
							if (!Writer::write_range(fd, base, run_end - run_begin, cp->file_pos + cp->block_off[k] + run_begin, msg))
								w.fail(path + ": " + msg);
						}
						}
					});
				}
			}
			records += cur->lines_prefix.back();
			if (fused) {
				// genomes in groups of eight neighbours of all_order: their rows of a dense gene are eight
				// neighbouring lines of the chunk (2 KB), and eight pieces of ~100 records fit the L2
				// (fewer where a genome's piece is large -- many genes per chunk: about 1 MB of text per task)
				constexpr int kGroup = 8;
				const int64_t piece_guess = (cur->lines_prefix.back() / std::max<int32_t>(n, 1) + 1) * (ht.L + 32);
				const int group = (int)std::max<int64_t>(1, std::min<int64_t>(kGroup, ((int64_t)1 << 20) / piece_guess));
				for (int32_t p0 = 0; p0 < n; p0 += group)
					submit_for(cp, [&, cp, p0, group] {
						const int64_t L = ht.L;
						const int cnt = (int)std::min<int64_t>(group, n - p0);
						struct Cursor {
							int32_t genome, leaf;
							const int64_t *sp; // the leaf's sparse genes (ascending table index)
							int64_t n_sp, si;
							char *buf;
							int64_t begin, len; // the piece: offset of its first record in the file, bytes so far
							char name[24];      // ">genomeNNN"
							int name_len;
							int64_t per_record;          // bytes of a record less the digits of its gene id
							const int64_t *leaf_digits;  // digits of the leaf's sparse gene ids before its r-th
						} cs[kGroup];
						// a record: '>' name(<= 16) '.' id(<= 10) '\n' bases '\n'; a genome's piece: its genes of the chunk
						const int64_t dense_here = ht.dense_before[cp->g2] - ht.dense_before[cp->g];
						int64_t need = 0, caps[kGroup];
						for (int i = 0; i < cnt; i++) {
							Cursor &c = cs[i];
							c.genome = e.all_order[p0 + i];
							c.leaf = ht.genome_leaf[c.genome];
							c.sp = ht.sp_gene + ht.sp_off[c.leaf];
							c.n_sp = ht.sp_off[c.leaf + 1] - ht.sp_off[c.leaf];
							c.si = std::lower_bound(c.sp, c.sp + c.n_sp, cp->g) - c.sp;
							const int64_t sparse_here = (std::lower_bound(c.sp, c.sp + c.n_sp, cp->g2) - c.sp) - c.si;
							caps[i] = (dense_here + sparse_here) * (L + 32);
							need += caps[i];
							c.begin = -1;
							c.len = 0;
							c.name[0] = '>';
							c.name_len = (int)(put_genome_name(c.name + 1, c.genome) - c.name);
							c.per_record = genome_name_len(c.genome) + L + 4;
							c.leaf_digits = ht.sp_digits + ht.sp_off[c.leaf] + c.leaf;
						}
						char *all = thread_buffer((size_t)need);
						for (int i = 0; i < cnt; i++) {
							cs[i].buf = all;
							all += caps[i];
						}
						const char *rows_in = cp->h_rows;
						char text[64];
						text[0] = '.';
						for (int64_t k = cp->g; k < cp->g2; k++) {
							const int64_t first_line = cp->lines_prefix[k - cp->g];
							const bool dense = ht.dense_slot[k] >= 0;
							if (k + 2 < cp->g2 && ht.dense_slot[k + 2] >= 0) { // the group's rows two genes ahead
								const char *nx = rows_in + (cp->lines_prefix[k + 2 - cp->g] + 1 + p0) * row_bytes;
								for (int64_t b = 0; b < cnt * row_bytes; b += 64) __builtin_prefetch(nx + b, 0, 3);
							}
							// ".<gene id>\n" is the same for the whole group
							char *const id_end = host_put_uint(text + 1, (uint64_t)ht.gene_id[k], 1);
							*id_end = '\n';
							const int id_len = (int)(id_end + 1 - text);
							for (int i = 0; i < cnt; i++) {
								Cursor &c = cs[i];
								int64_t line;
								const int64_t srank = c.si; // the leaf's sparse genes before gene k (fasta_record's search, for free)
								if (dense) {
									line = 1 + p0 + i;
								} else {
									if (c.si >= c.n_sp || c.sp[c.si] != k) continue;
									line = 1 + sp_line[(size_t)(ht.sp_off[c.leaf] + c.si)];
									c.si++;
								}
								const int64_t off = (ht.dense_before[k] + srank) * c.per_record + ht.dense_digits_before[k] + c.leaf_digits[srank];
								if (c.begin < 0) c.begin = off;
								char *d = c.buf + (off - c.begin);
								std::memcpy(d, c.name, (size_t)c.name_len); // ">genomeNNN"
								std::memcpy(d + c.name_len, text, (size_t)id_len);
								const int plen = c.name_len + id_len;
								expand_bases(rows_in + (first_line + line) * row_bytes, L, d + plen);
								d[plen + L] = '\n';
								c.len = off - c.begin + plen + L + 1;
							}
						}
						for (int i = 0; i < cnt; i++) {
							const Cursor &c = cs[i];
							if (c.begin < 0) continue;
							char name[64];
							const int nl = format_genome_name(c.genome, name, sizeof name);
							hx_deliver(std::string(name, (size_t)nl) + ".fasta", c.begin, c.buf, c.len, c.begin == 0);
						}
					});
			}
			task_done(cp);
			expanding.push_back(std::move(cur));
			// chunks that are done leave; with every slot taken the oldest one has to
			while (!expanding.empty() &&
				   ((int)(copying.size() + expanding.size()) >= R || expanding.front()->pending.load() == 0))
				retire_front();
		}
		while (!expanding.empty()) retire_front();
		lap_waits("alignment.maf");
		if (gathered && cudaEventSynchronize(k1) == cudaSuccess) {
			float ms = 0; // (the gather kernels are a few microseconds each: the last one stands for the call's K3 time)
			if (cudaEventElapsedTime(&ms, k0, k1) == cudaSuccess) device_ms += ms;
		}
		if (fd >= 0) ::close(fd);
		slot = 0;
		{
			std::lock_guard<std::mutex> l(w.err_mutex);
			if (!w.error.empty()) return fail_out(PGS_ERR_IO, w.error);
		}
	}

	// ---- alignment.maf
	if ((spec->what & PGS_OUT_MAF) && !host_expand) {
		const char *header = kMafHeader;
		const bool with_header = !shard || shard->maf_offset == 0;
		const int64_t hlen = with_header ? (int64_t)sizeof(kMafHeader) - 1 : 0;
		int64_t file_pos = shard ? shard->maf_offset : 0;
		int64_t g = 0;
		bool first = true;
		while (g < G || first) {
			int64_t bytes = first ? hlen : 0;
			int64_t g2 = g;
			std::vector<int64_t> lines_prefix{0}, block_off;
			while (g2 < G && bytes + maf_bytes[g2] <= stage && (g2 - g) < 65536) {
				block_off.push_back(bytes);
				bytes += maf_bytes[g2];
				lines_prefix.push_back(lines_prefix.back() + maf_lines[g2]);
				g2++;
			}
			begin_chunk(slot);
			if (first && hlen > 0) {
				ce = cudaMemcpyAsync(d_stage[slot].p, header, (size_t)hlen, cudaMemcpyHostToDevice, e.stream);
				if (ce != cudaSuccess) {
					cleanup();
					return e.cuda_fail(ce, "maf header");
				}
			}
			const int64_t items = lines_prefix.back();
			ck(cudaEventRecord(k0, e.stream));
			if (items > 0) {
				if ((rc = h2d(d_aux_a[slot], lines_prefix)) || (rc = h2d(d_aux_b[slot], block_off))) {
					cleanup();
					return rc;
				}
				ck(cudaEventRecord(k0, e.stream));
				k3_maf<<<(unsigned)((items + kRecPerCta - 1) / kRecPerCta), kFmtWarps * 32, 0, e.stream>>>(
					t, mt, d_stage[slot].as<char>(), g, (int32_t)(g2 - g), d_aux_a[slot].as<int64_t>(),
					d_aux_b[slot].as<int64_t>());
				records += items;
			}
			std::vector<FilePiece> pieces{{"alignment.maf", file_pos, 0, bytes, false}};
			file_pos += bytes;
			if (!submit(slot, bytes, std::move(pieces))) return fail_out(PGS_ERR_IO, w.error);
			slot ^= 1;
			g = g2;
			first = false;
		}
	}

	lap("alignment.maf");
	// ---- genomes.pgs2 (optional): the genomes' packed 2-bit rows exactly as they
	// lie in the store, behind an index -- an eighth of the text's bytes over PCIe and to disk.
	// No kernel: the rows are copied straight from the store.  Layout: include/pgs.h.
	std::vector<char> packed_head; // must outlive the asynchronous writes
	if (spec->what & PGS_OUT_PACKED) {
		if (shard) return fail_out(PGS_ERR_ARG, "pgs_write_outputs: PGS_OUT_PACKED is not available with a shard layout");
		const int64_t row_bytes = e.blocks * 16;
		std::vector<int64_t> rec_off((size_t)n + 1, 0);
		for (int32_t i = 0; i < n; i++) {
			const int32_t v = e.genome_leaf[i];
			rec_off[i + 1] = rec_off[i] + D + (e.sp_off[v + 1] - e.sp_off[v]);
		}
		const int64_t total = rec_off[n], n_sparse = total - (int64_t)n * D;
		const int64_t index_bytes = 56 + 8 * ((int64_t)n + 1) + 16 * (D + n_sparse);
		const int64_t data_off = (index_bytes + 255) & ~(int64_t)255;
		packed_head.assign((size_t)data_off, 0);
		{
			char *h = packed_head.data();
			std::memcpy(h, "PGS2BIT", 8);
			const uint32_t version = 1, head32 = data_off <= 0xFFFFFFFFll ? (uint32_t)data_off : 0u; // 0: derive it from the counts
			std::memcpy(h + 8, &version, 4);
			std::memcpy(h + 12, &head32, 4);
			const int64_t fields[5] = {n, L, e.blocks, total, D};
			std::memcpy(h + 16, fields, 40);
			std::memcpy(h + 56, rec_off.data(), 8 * ((size_t)n + 1));
			// the dense records lead every genome's rows and are the same genes everywhere: listed once
			int64_t *dense_id = reinterpret_cast<int64_t *>(h + 56 + 8 * ((size_t)n + 1));
			int64_t *dense_key = dense_id + D, *sparse_id = dense_key + D, *sparse_key = sparse_id + n_sparse;
			for (int64_t sl = 0; sl < D; sl++) {
				dense_id[sl] = e.gene_id[e.dense_gene[sl]];
				dense_key[sl] = e.dense_gene[sl]; // gene table index = position key of genomeNNN.fasta
			}
			int64_t r = 0;
			for (int32_t i = 0; i < n; i++) {
				const int32_t v = e.genome_leaf[i];
				for (int64_t j = e.sp_off[v]; j < e.sp_off[v + 1]; j++, r++) {
					sparse_id[r] = e.gene_id[e.sp_gene[j]];
					sparse_key[r] = e.sp_gene[j];
				}
			}
		}
		if (!w.put("genomes.pgs2", 0, packed_head.data(), data_off, false) || !w.finish())
			return fail_out(PGS_ERR_IO, w.error);
		const char *store = reinterpret_cast<const char *>(e.tables.rows);
		int32_t i = 0;
		while (i < n) {
			std::vector<Segment> segs;
			int64_t bytes = 0;
			int32_t j = i;
			while (j < n) {
				const int64_t gb = (rec_off[j + 1] - rec_off[j]) * row_bytes;
				if (bytes + gb > stage && j > i) break;
				const char *src = store + e.row_base[e.genome_leaf[j]] * row_bytes;
				if (!segs.empty() && segs.back().src + segs.back().bytes == src)
					segs.back().bytes += gb; // neighbours in the store: one copy
				else if (gb > 0)
					segs.push_back({src, bytes, gb});
				bytes += gb;
				j++;
			}
			std::vector<FilePiece> pieces{{"genomes.pgs2", data_off + rec_off[i] * row_bytes, 0, bytes, false}};
			if (!submit(slot, bytes, std::move(pieces), &segs)) return fail_out(PGS_ERR_IO, w.error);
			records += rec_off[j] - rec_off[i];
			slot ^= 1;
			i = j;
		}
	}

	if (!drain(slot) || !drain(slot ^ 1)) return fail_out(PGS_ERR_IO, w.error);
	ce = cudaStreamSynchronize(e.stream);
	if (ce == cudaSuccess) ce = cudaStreamSynchronize(copy_stream);
	if (ce == cudaSuccess) ce = cudaGetLastError();
	cleanup();
	w.close_file();
	if (ce != cudaSuccess) return e.cuda_fail(ce, "pgs_write_outputs");
	e.format_ms = device_ms;
	w.rep.files += hx_files.load();
	w.rep.bytes += hx_bytes.load();
	if (report) {
		*report = w.rep;
		lap("packed, drain");
		report->records = records;
		report->d2h_bytes = d2h_bytes;
		report->device_ms = device_ms;
		report->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
	}
	return PGS_OK;
}

} // namespace pgs
