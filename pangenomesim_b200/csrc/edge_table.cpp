#include "edge_table.h"

#include <cmath>
#include <limits>

namespace pgs {

double jc_probability(double rate)
{
	if (!(rate > 0.0)) return 0.0;
	const double p = -0.75 * std::expm1(rate * (-4.0 / 3.0));
	return p > 0.75 ? 0.75 : (p < 0.0 ? 0.0 : p);
}

int build_edge_table(double rate, uint64_t H[65])
{
	constexpr double two64 = 18446744073709551616.0;
	H[0] = std::numeric_limits<uint64_t>::max();
	for (int k = 1; k <= 64; k++) H[k] = 0;
	const double p = jc_probability(rate);
	if (p <= 0.0) return 0;

	// binomial pmf by the ratio recurrence, then tails summed from the far end
	double pmf[65];
	const double odds = p / (1.0 - p);
	pmf[0] = std::exp(64.0 * std::log1p(-p));
	for (int k = 1; k <= 64; k++) {
		const double up = pmf[k - 1] * (double)(65 - k);
		pmf[k] = (up / (double)k) * odds;
	}
	double tail = 0.0;
	int kmax = 0;
	for (int k = 64; k >= 1; k--) {
		tail += pmf[k];
		const double scaled = tail * two64;
		H[k] = scaled >= 18446744073709551615.0 ? std::numeric_limits<uint64_t>::max()
												: (uint64_t)std::floor(scaled);
		if (H[k] > 0 && k > kmax) kmax = k;
	}
	return kmax;
}

} // namespace pgs
