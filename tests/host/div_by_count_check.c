/* div_by_count (picard-ica_b200/csrc/exact_div.h) against the hardware division: random operands over 600 binades, small integers
 * times powers of two (exact quotients and ties), signed zeros, denormals, infinities.  Prints the number of mismatching results. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../picard-ica_b200/csrc/exact_div.h"

static int same(double a, double b) { return memcmp(&a, &b, 8) == 0 || (a != a && b != b); }

int main(int argc, char** argv) {
  const long per_k = argc > 1 ? atol(argv[1]) : 1000000;
  uint64_t s = 88172645463325252ULL;
  long bad = 0, tot = 0;
  for (int k = 2; k <= 30; ++k) {
    const double kk = k, rcp = 1.0 / kk;
    for (long it = 0; it < per_k; ++it) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      uint64_t bits = s, ex = (bits >> 52) & 0x7ff;
      ex = 723 + (ex % 600);
      bits = (bits & 0x800fffffffffffffULL) | (ex << 52);
      double a; memcpy(&a, &bits, 8);
      bad += !same(div_by_count(a, kk, rcp), a / kk); ++tot;
    }
    for (long m = -20000; m <= 20000; ++m)
      for (int e = -60; e <= 60; e += 15) {
        const double a = ldexp((double)m, e);
        bad += !same(div_by_count(a, kk, rcp), a / kk); ++tot;
      }
    const double edge[] = {0.0, -0.0, 4.9e-324, -4.9e-324, 1e-300, -1e-300, 2.2250738585072014e-308, 1e300, -1e300, 1.7976931348623157e308,
                           INFINITY, -INFINITY, NAN};
    for (unsigned i = 0; i < sizeof edge / sizeof edge[0]; ++i) { bad += !same(div_by_count(edge[i], kk, rcp), edge[i] / kk); ++tot; }
  }
  printf("{\"mismatches\": %ld, \"cases\": %ld}\n", bad, tot);
  return 0;
}
