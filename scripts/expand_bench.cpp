// development: host expander throughput against the size of the buffer the text goes to
//   g++ -O2 -std=c++17 -pthread -Ipangenomesim_b200/csrc scripts/expand_bench.cpp pangenomesim_b200/csrc/expand.cpp -o /tmp/expand_bench
//   /tmp/expand_bench <threads>
#include "expand.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

int main(int argc, char **argv)
{
	const int threads = argc > 1 ? std::atoi(argv[1]) : 16;
	const int64_t L = 1000, row_bytes = 256, rec = L + 20;
	const size_t src_bytes = (size_t)256 << 20; // per thread: streamed once per pass, as the pinned buffers are
	std::printf("expander: %s, %d threads\n", pgs::expand_kind(), threads);
	for (size_t buf_kb : {128, 256, 512, 1024, 2048, 5120, 32768}) {
		std::vector<std::thread> pool;
		std::vector<double> secs(threads);
		const size_t recs_per_buf = buf_kb * 1024 / rec;
		for (int t = 0; t < threads; t++)
			pool.emplace_back([&, t] {
				std::vector<unsigned char> src(src_bytes);
				for (size_t i = 0; i < src_bytes; i += 64) src[i] = (unsigned char)(i * 2654435761u >> 24);
				std::vector<char> buf(buf_kb * 1024 + 64);
				const size_t rows = src_bytes / row_bytes;
				auto t0 = std::chrono::steady_clock::now();
				size_t k = 0;
				for (size_t r = 0; r < rows; r++) {
					char *d = buf.data() + k * rec;
					std::memcpy(d, ">genome0001.12345\n", 18);
					pgs::expand_bases(src.data() + r * row_bytes, L, d + 18);
					d[18 + L] = '\n';
					if (++k == recs_per_buf) k = 0;
				}
				secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			});
		for (auto &th : pool) th.join();
		double worst = 0;
		for (double s : secs) worst = s > worst ? s : worst;
		const double text = (double)threads * (src_bytes / row_bytes) * rec;
		std::printf("buffer %6zu KB: %.1f GB/s of text (%.2f GB/s per thread)\n", buf_kb, text / worst / 1e9, text / worst / 1e9 / threads);
	}
	return 0;
}
