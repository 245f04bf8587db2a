// expand.cpp -- host side of PGS_OUT_HOST_EXPAND: 2-bit packed rows -> ASCII bases on the CPU.
//
// The reference formats every record on the host anyway (gene::to_fasta, gene.h:42-52; the MAF `s`
// lines, pangenomesim.cxx:268-287).  Text is 4.1x the size of the packed rows, and PCIe (~53 GB/s)
// is what bounds the device formatter end to end; this path ships the packed rows instead and lets
// the host's cores write the text where it has to end up.  Output bytes are identical.
#include "expand.h"

#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace pgs {

namespace {

// byte (four sites, two bits each, site 0 in the low bits) -> four ASCII bases
struct Lut4 {
	uint32_t v[256];
	Lut4()
	{
		static const char base[5] = "ACGT";
		for (int b = 0; b < 256; b++) {
			uint32_t w = 0;
			for (int s = 0; s < 4; s++) w |= (uint32_t)(unsigned char)base[(b >> (2 * s)) & 3] << (8 * s);
			v[b] = w;
		}
	}
};
const Lut4 g_lut;

void expand_scalar(const unsigned char *src, int64_t sites, char *dst)
{
	const int64_t full = sites / 4;
	for (int64_t i = 0; i < full; i++) std::memcpy(dst + 4 * i, &g_lut.v[src[i]], 4);
	for (int64_t s = 4 * full; s < sites; s++) dst[s] = "ACGT"[(src[s >> 2] >> (2 * (s & 3))) & 3];
}

#if defined(__x86_64__)
// 16 packed bytes (64 sites) -> 64 ASCII bytes: both nibbles of every byte through two 16-entry
// PSHUFB tables (first base of a nibble, second base of a nibble), then two rounds of interleaving.
__attribute__((target("ssse3"))) inline void block64_ssse3(const unsigned char *src, char *dst)
{
	const __m128i first = _mm_setr_epi8('A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T');
	const __m128i second = _mm_setr_epi8('A', 'A', 'A', 'A', 'C', 'C', 'C', 'C', 'G', 'G', 'G', 'G', 'T', 'T', 'T', 'T');
	const __m128i low = _mm_set1_epi8(0x0F);
	const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src));
	const __m128i lo = _mm_and_si128(x, low), hi = _mm_and_si128(_mm_srli_epi16(x, 4), low);
	const __m128i c0 = _mm_shuffle_epi8(first, lo), c1 = _mm_shuffle_epi8(second, lo);
	const __m128i c2 = _mm_shuffle_epi8(first, hi), c3 = _mm_shuffle_epi8(second, hi);
	const __m128i a01 = _mm_unpacklo_epi8(c0, c1), b01 = _mm_unpackhi_epi8(c0, c1); // c0[i] c1[i] pairs
	const __m128i a23 = _mm_unpacklo_epi8(c2, c3), b23 = _mm_unpackhi_epi8(c2, c3);
	_mm_storeu_si128(reinterpret_cast<__m128i *>(dst), _mm_unpacklo_epi16(a01, a23));
	_mm_storeu_si128(reinterpret_cast<__m128i *>(dst + 16), _mm_unpackhi_epi16(a01, a23));
	_mm_storeu_si128(reinterpret_cast<__m128i *>(dst + 32), _mm_unpacklo_epi16(b01, b23));
	_mm_storeu_si128(reinterpret_cast<__m128i *>(dst + 48), _mm_unpackhi_epi16(b01, b23));
}

// The rest of a row after the full vectors: when the row is at least one vector long and the rest starts on a byte
// of the packed row (sites a multiple of four -- gene lengths 1000, 1500, 2000), the LAST vector's worth of sites is
// expanded once more, overlapping what is already there (same bytes): no scalar tail, nothing read or written beyond
// the row.  (L = 1000: 7 vectors of 128 + 104 sites = one 64-site vector and 40 sites through the byte table before.)
__attribute__((target("ssse3"))) void expand_ssse3(const unsigned char *src, int64_t sites, char *dst)
{
	int64_t s = 0;
	for (; s + 64 <= sites; s += 64) block64_ssse3(src + s / 4, dst + s);
	if (s == sites) return;
	if (sites >= 64 && (sites & 3) == 0)
		block64_ssse3(src + (sites - 64) / 4, dst + (sites - 64));
	else
		expand_scalar(src + s / 4, sites - s, dst + s);
}

__attribute__((target("avx2"))) inline void block128_avx2(const unsigned char *src, char *dst)
{
	const __m256i first = _mm256_setr_epi8('A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A',
										   'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T');
	const __m256i second = _mm256_setr_epi8('A', 'A', 'A', 'A', 'C', 'C', 'C', 'C', 'G', 'G', 'G', 'G', 'T', 'T', 'T', 'T', 'A', 'A', 'A', 'A', 'C',
											'C', 'C', 'C', 'G', 'G', 'G', 'G', 'T', 'T', 'T', 'T');
	const __m256i low = _mm256_set1_epi8(0x0F);
	// 32 packed bytes: lane 0 = the first 64 sites, lane 1 = the next 64
	const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src));
	const __m256i lo = _mm256_and_si256(x, low), hi = _mm256_and_si256(_mm256_srli_epi16(x, 4), low);
	const __m256i c0 = _mm256_shuffle_epi8(first, lo), c1 = _mm256_shuffle_epi8(second, lo);
	const __m256i c2 = _mm256_shuffle_epi8(first, hi), c3 = _mm256_shuffle_epi8(second, hi);
	const __m256i a01 = _mm256_unpacklo_epi8(c0, c1), b01 = _mm256_unpackhi_epi8(c0, c1);
	const __m256i a23 = _mm256_unpacklo_epi8(c2, c3), b23 = _mm256_unpackhi_epi8(c2, c3);
	const __m256i q0 = _mm256_unpacklo_epi16(a01, a23), q1 = _mm256_unpackhi_epi16(a01, a23);
	const __m256i q2 = _mm256_unpacklo_epi16(b01, b23), q3 = _mm256_unpackhi_epi16(b01, b23);
	// per 128-bit lane q0..q3 are consecutive 16-byte pieces: lane 0 -> dst[0 .. 63], lane 1 -> dst[64 ..]
	_mm256_storeu_si256(reinterpret_cast<__m256i *>(dst), _mm256_permute2x128_si256(q0, q1, 0x20));
	_mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + 32), _mm256_permute2x128_si256(q2, q3, 0x20));
	_mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + 64), _mm256_permute2x128_si256(q0, q1, 0x31));
	_mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + 96), _mm256_permute2x128_si256(q2, q3, 0x31));
}

__attribute__((target("avx2"))) void expand_avx2(const unsigned char *src, int64_t sites, char *dst)
{
	int64_t s = 0;
	for (; s + 128 <= sites; s += 128) block128_avx2(src + s / 4, dst + s);
	if (s == sites) return;
	if (sites >= 128 && (sites & 3) == 0)
		block128_avx2(src + (sites - 128) / 4, dst + (sites - 128));
	else
		expand_ssse3(src + s / 4, sites - s, dst + s);
}
#endif

typedef void (*expand_fn)(const unsigned char *, int64_t, char *);

expand_fn pick()
{
#if defined(__x86_64__)
	__builtin_cpu_init();
	if (__builtin_cpu_supports("avx2")) return expand_avx2;
	if (__builtin_cpu_supports("ssse3")) return expand_ssse3;
#endif
	return expand_scalar;
}
const expand_fn g_expand = pick();

} // namespace

void expand_bases(const void *packed_row, int64_t sites, char *dst)
{
	g_expand(static_cast<const unsigned char *>(packed_row), sites, dst);
}

const char *expand_kind()
{
#if defined(__x86_64__)
	if (g_expand == expand_avx2) return "avx2";
	if (g_expand == expand_ssse3) return "ssse3";
#endif
	return "scalar";
}

} // namespace pgs
