/* rans64_restated.h -- TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Restatement of the 64-bit rANS primitives that the reference's coder
 * (/root/reference/cra5/models/compressai/cpp_exts/rans/rans_interface.cpp:44,
 *  `#include "rans64.h"`) expects from the un-vendored third-party header
 * rygorous/ryg_rans `rans64.h` (public domain; the reference's setup.py:68
 * looks for it under third_party/ryg_rans, which is absent from the mount).
 * Algorithm as published by F. Giesen ("rANS with static probability
 * distributions", 64-bit state / 32-bit renormalisation variant) and as
 * specified in SURVEY.md Appendix B:
 *   state x in [L, L<<32), L = 2^31, words emitted backwards.
 * Installed as `rans64.h` into oracle/_ref/include by oracle/build_ref.py so
 * the reference's own .cpp compiles unmodified from where it lies.
 */
#ifndef RANS64_RESTATED_H
#define RANS64_RESTATED_H
#include <stdint.h>
#include <assert.h>

#define Rans64Assert(x) assert(x)
#define RANS64_L (1ull << 31)

typedef uint64_t Rans64State;

static inline void Rans64EncInit(Rans64State *r) { *r = RANS64_L; }

/* encode symbol [start, start+freq) out of 2^scale_bits */
static inline void Rans64EncPut(Rans64State *r, uint32_t **pptr, uint32_t start,
                                uint32_t freq, uint32_t scale_bits) {
  uint64_t x = *r;
  uint64_t x_max = ((RANS64_L >> scale_bits) << 32) * freq;
  if (x >= x_max) {
    *pptr -= 1;
    **pptr = (uint32_t)x;
    x >>= 32;
  }
  *r = ((x / freq) << scale_bits) + (x % freq) + start;
}

static inline void Rans64EncFlush(Rans64State *r, uint32_t **pptr) {
  uint64_t x = *r;
  *pptr -= 2;
  (*pptr)[0] = (uint32_t)(x >> 0);
  (*pptr)[1] = (uint32_t)(x >> 32);
}

static inline void Rans64DecInit(Rans64State *r, uint32_t **pptr) {
  uint64_t x;
  x = (uint64_t)((*pptr)[0]) << 0;
  x |= (uint64_t)((*pptr)[1]) << 32;
  *pptr += 2;
  *r = x;
}

static inline uint32_t Rans64DecGet(Rans64State *r, uint32_t scale_bits) {
  return (uint32_t)(*r & ((1u << scale_bits) - 1));
}

static inline void Rans64DecAdvance(Rans64State *r, uint32_t **pptr, uint32_t start,
                                    uint32_t freq, uint32_t scale_bits) {
  uint64_t mask = (1ull << scale_bits) - 1;
  uint64_t x = *r;
  x = freq * (x >> scale_bits) + (x & mask) - start;
  if (x < RANS64_L) {
    x = (x << 32) | **pptr;
    *pptr += 1;
  }
  *r = x;
}
#endif
