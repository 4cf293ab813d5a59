/* rans_oracle.c -- TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Plain-C restatement of the reference's native entropy-coding arithmetic:
 *   - pmf_to_quantized_cdf   cra5/models/compressai/cpp_exts/ops/ops.cpp:40-109
 *   - rANS encode            cra5/models/compressai/cpp_exts/rans/rans_interface.cpp:108-200
 *   - rANS decode            cra5/models/compressai/cpp_exts/rans/rans_interface.cpp:215-284
 * on top of the 64-bit rANS primitives of the un-vendored dependency ryg_rans `rans64.h`
 * (restated in oracle/rans64_restated.h / SURVEY.md Appendix B).
 *
 * Pinned against the reference's own compiled coder (oracle/_ref) by tests/test_oracle_pins.py and the KATs
 * in tests/golden/rans_kat.json.   Build: see oracle/Makefile  ->  oracle/_build/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION 16u       /* rans_interface.cpp:49 */
#define BYPASS_BITS 4u      /* rans_interface.cpp:51 */
#define BYPASS_MAX 15u      /* rans_interface.cpp:52 */
#define RANS_L (1ull << 31) /* rans64.h */

/* -------------------------------------------------------------------------------------------- pmf -> cdf */
/* returns 0 on success, 1 on invalid pmf (reference throws std::domain_error, ops.cpp:46-64) */
int oracle_pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* cdf /* n+1 */) {
  for (int i = 0; i < n; ++i)
    if (pmf[i] < 0 || !isfinite(pmf[i])) return 1;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision)); /* ops.cpp:58-59 */
  uint32_t total = 0;
  for (int i = 0; i <= n; ++i) total += cdf[i];
  if (total == 0) return 1;
  for (int i = 0; i <= n; ++i) cdf[i] = (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total); /* ops.cpp:67-70 */
  for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
  cdf[n] = 1u << precision;
  for (int i = 0; i < n; ++i) {
    if (cdf[i] == cdf[i + 1]) { /* zero-frequency symbol: steal from the smallest freq > 1, ops.cpp:75-99 */
      uint32_t best_freq = ~0u;
      int best = -1;
      for (int j = 0; j < n; ++j) {
        uint32_t f = cdf[j + 1] - cdf[j];
        if (f > 1 && f < best_freq) { best_freq = f; best = j; }
      }
      if (best < 0) return 2;
      if (best < i) {
        for (int j = best + 1; j <= i; ++j) cdf[j]--;
      } else {
        for (int j = i + 1; j <= best; ++j) cdf[j]++;
      }
    }
  }
  return 0;
}

/* -------------------------------------------------------------------------------------------- encoder */
typedef struct { uint16_t start, range; uint8_t bypass; } sym_t;

static inline void enc_put(uint64_t* x, uint32_t** pp, uint32_t start, uint32_t freq) {
  uint64_t xm = ((RANS_L >> PRECISION) << 32) * freq;
  if (*x >= xm) { *pp -= 1; **pp = (uint32_t)*x; *x >>= 32; }
  *x = ((*x / freq) << PRECISION) + (*x % freq) + start;
}
static inline void enc_put_bits(uint64_t* x, uint32_t** pp, uint32_t val, uint32_t nbits) { /* :69-87 */
  uint32_t freq = 1u << (16 - nbits);
  uint64_t xm = ((RANS_L >> 16) << 32) * freq;
  if (*x >= xm) { *pp -= 1; **pp = (uint32_t)*x; *x >>= 32; }
  *x = (*x << nbits) | val;
}

/* cdfs: [n_cdfs][cdf_stride] int32. Output: out (capacity out_cap bytes). Returns bytes written, <0 on error. */
long oracle_rans_encode(const int32_t* symbols, const int32_t* indexes, long n, const int32_t* cdfs, int cdf_stride,
                        const int32_t* cdf_sizes, const int32_t* offsets, uint8_t* out, long out_cap) {
  /* pass 1: symbols -> (start, range, bypass) list, forward order (rans_interface.cpp:117-172) */
  long cap = n + 16, cnt = 0;
  sym_t* syms = (sym_t*)malloc(sizeof(sym_t) * (size_t)cap);
  if (!syms) return -1;
#define PUSH(s_, r_, b_)                                                   \
  do {                                                                     \
    if (cnt == cap) { cap = cap * 2; syms = (sym_t*)realloc(syms, sizeof(sym_t) * (size_t)cap); } \
    syms[cnt].start = (uint16_t)(s_); syms[cnt].range = (uint16_t)(r_); syms[cnt].bypass = (b_); ++cnt; \
  } while (0)
  for (long i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    const int32_t* cdf = cdfs + (size_t)ci * cdf_stride;
    const int32_t max_value = cdf_sizes[ci] - 2;
    int32_t value = symbols[i] - offsets[ci];
    uint32_t raw = 0;
    if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; }
    else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
    PUSH(cdf[value], cdf[value + 1] - cdf[value], 0);
    if (value == max_value) {
      int32_t nb = 0;
      while ((raw >> (nb * BYPASS_BITS)) != 0) ++nb;
      int32_t v = nb;
      while (v >= (int32_t)BYPASS_MAX) { PUSH(BYPASS_MAX, BYPASS_MAX + 1, 1); v -= BYPASS_MAX; }
      PUSH(v, v + 1, 1);
      for (int32_t j = 0; j < nb; ++j) { uint32_t q = (raw >> (j * BYPASS_BITS)) & BYPASS_MAX; PUSH(q, q + 1, 1); }
    }
  }
  /* pass 2: encode in reverse, words written backwards (rans_interface.cpp:175-200) */
  uint32_t* buf = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(cnt + 4));
  uint32_t* ptr = buf + cnt + 4;
  uint64_t x = RANS_L;
  for (long i = cnt - 1; i >= 0; --i) {
    if (!syms[i].bypass) enc_put(&x, &ptr, syms[i].start, syms[i].range);
    else enc_put_bits(&x, &ptr, syms[i].start, BYPASS_BITS);
  }
  ptr -= 2; ptr[0] = (uint32_t)x; ptr[1] = (uint32_t)(x >> 32);
  long nbytes = (long)((buf + cnt + 4) - ptr) * 4;
  long rv = nbytes;
  if (nbytes > out_cap) rv = -2; else memcpy(out, ptr, (size_t)nbytes);
  free(buf); free(syms);
  return rv;
}

/* -------------------------------------------------------------------------------------------- decoder */
static inline uint32_t dec_get_bits(uint64_t* x, const uint32_t** pp, uint32_t nbits) { /* :89-105 */
  uint32_t val = (uint32_t)(*x & ((1u << nbits) - 1));
  *x >>= nbits;
  if (*x < RANS_L) { *x = (*x << 32) | **pp; *pp += 1; }
  return val;
}

/* returns 0 on success */
int oracle_rans_decode(const uint8_t* stream, long nbytes, const int32_t* indexes, long n, const int32_t* cdfs,
                       int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, int32_t* out) {
  (void)nbytes;
  const uint32_t* ptr = (const uint32_t*)stream;
  uint64_t x = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
  ptr += 2;
  for (long i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    const int32_t* cdf = cdfs + (size_t)ci * cdf_stride;
    const int32_t max_value = cdf_sizes[ci] - 2;
    const uint32_t cum = (uint32_t)(x & 0xffffu);
    int32_t s = 0; /* linear search for the first entry > cum (rans_interface.cpp:246-250) */
    while (s < cdf_sizes[ci] && (uint32_t)cdf[s] <= cum) ++s;
    s -= 1;
    const uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
    x = (uint64_t)freq * (x >> PRECISION) + (x & 0xffffu) - start;
    if (x < RANS_L) { x = (x << 32) | *ptr; ptr += 1; }
    int32_t value = s;
    if (value == max_value) { /* bypass, rans_interface.cpp:256-278 */
      int32_t val = (int32_t)dec_get_bits(&x, &ptr, BYPASS_BITS);
      int32_t nb = val;
      while (val == (int32_t)BYPASS_MAX) { val = (int32_t)dec_get_bits(&x, &ptr, BYPASS_BITS); nb += val; }
      int32_t raw = 0;
      for (int32_t j = 0; j < nb; ++j) { val = (int32_t)dec_get_bits(&x, &ptr, BYPASS_BITS); raw |= val << (j * BYPASS_BITS); }
      value = raw >> 1;
      if (raw & 1) value = -value - 1; else value += max_value;
    }
    out[i] = value + offsets[ci];
  }
  return 0;
}
