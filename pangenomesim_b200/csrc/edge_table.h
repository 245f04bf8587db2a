// edge_table.h -- per-branch decision tables (host side, double arithmetic, computed once).
//
// The reference substitutes about rate*L sites per gene::mutate call (gene.cxx:50-58, SURVEY.md
// Q2/Q3); the north_star replaces that by Jukes-Cantor: a site differs after a branch of length
// mu*t with probability p = 3/4 (1 - exp(-4/3 mu t)).  For one 64-site block the number of
// substituted sites is K ~ Binomial(64, p); the device compares one 64-bit uniform against the
// descending tail thresholds H[k] = floor(2^64 P(K >= k)).  Integer thresholds make the device
// result independent of device floating point.
#pragma once
#include <cstdint>

namespace pgs {

double jc_probability(double rate);

// Fills H[0..64] (H[0] = 2^64-1) and returns kmax = largest k with H[k] > 0.
int build_edge_table(double rate, uint64_t H[65]);

} // namespace pgs
