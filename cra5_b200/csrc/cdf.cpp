// cdf.cpp -- host-side construction of integer CDF tables (run once per model at update()).
// Same arithmetic as the reference helper compressai._CXX.pmf_to_quantized_cdf
// (cra5/models/compressai/cpp_exts/ops/ops.cpp:40-109): scale each probability to 2^precision with round-half-away,
// renormalise by the integer total with floor division, accumulate, force the last entry to 2^precision, then repair
// every zero-width bin by taking one count from the narrowest bin that still has more than one.
#include <cmath>
#include <cstdint>
#include <vector>

#include "host_util.h"

namespace cra5 {

void pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* cdf) {
  CRA5_CHECK(n >= 1 && precision >= 1 && precision <= 31, ERR_INVALID, "pmf_to_quantized_cdf: bad arguments");
  const uint32_t one = 1u << precision;
  uint32_t total = 0;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) {
    const float p = pmf[i];
    if (!(p >= 0.0f) || !std::isfinite(p))
      throw Error(ERR_INVALID, "Invalid `pmf`, non-finite or negative element found: " + std::to_string(p));
    cdf[i + 1] = static_cast<uint32_t>(std::round(p * static_cast<float>(one)));
    total += cdf[i + 1];
  }
  if (total == 0) throw Error(ERR_INVALID, "Invalid `pmf`: at least one element must have a non-zero probability.");
  uint32_t running = 0;
  for (int i = 1; i <= n; ++i) {
    running += static_cast<uint32_t>((static_cast<uint64_t>(one) * cdf[i]) / total);
    cdf[i] = running;
  }
  cdf[n] = one;
  for (int i = 0; i < n; ++i) {
    if (cdf[i + 1] != cdf[i]) continue;
    int donor = -1;
    uint32_t donor_width = ~0u;
    for (int j = 0; j < n; ++j) {
      const uint32_t width = cdf[j + 1] - cdf[j];
      if (width > 1 && width < donor_width) {
        donor_width = width;
        donor = j;
      }
    }
    CRA5_CHECK(donor >= 0, ERR_INVALID, "pmf_to_quantized_cdf: more symbols than probability slots");
    if (donor < i) {
      for (int j = donor + 1; j <= i; ++j) --cdf[j];
    } else {
      for (int j = i + 1; j <= donor; ++j) ++cdf[j];
    }
  }
}

}  // namespace cra5
