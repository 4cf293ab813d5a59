// coder.h -- host driver of the chunk-parallel GPU rANS coder (entropy.cu) and its container format.
//
// Container ("CR5B", little endian), one per coded tensor -- it takes the place of the single sequential stream the
// reference puts into `strings[i][0]` (vaeformer.py:348, cra5_api.py:108-116):
//     0  char[4]  magic "CR5B"
//     4  u8 version (1), u8 flags (0), u16 reserved
//     8  u32 n_channels      12  u32 L (symbols per channel)      16  u32 spc (sub-streams per channel)
//    20  u32 n_streams (= n_channels * spc)
//    24  u32 length[n_streams]   (bytes, multiples of 4)
//    ..  payload: the sub-streams back to back; sub-stream s = c * spc + k codes symbols c*L + k + i*spc, i = 0,1,..
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "model.h"

namespace cra5 {

constexpr uint32_t CR5B_HEADER = 24;
constexpr int CR5B_MAX_SPC = 64;

class RansCoder {
 public:
  RansCoder(size_t max_symbols, int max_channels);
  ~RansCoder();
  RansCoder(const RansCoder&) = delete;

  static size_t max_container_bytes(size_t n_symbols, int n_streams) {
    return CR5B_HEADER + 4 * (size_t)n_streams + 8 * n_symbols + 16 * (size_t)n_streams;
  }
  // symbols/indexes on the device -> container in host memory (pinned or pageable). Returns its size.
  size_t encode(cudaStream_t st, const int32_t* sym, const uint8_t* idx, const CdfTable& tab, int n_channels, int L,
                int spc, uint8_t* host_out, size_t host_cap);
  // Two-phase form for PINNED + MAPPED output buffers (cudaHostAlloc(.., cudaHostAllocMapped), at least
  // max_container_bytes() large): encode_begin only enqueues work -- the kernels write lengths and payload straight
  // into host memory -- so several tensors share ONE stream synchronisation; after it, encode_end checks the error
  // word, writes the header and returns the container size. slot (0 or 1) selects the metadata slot.
  // A batch of `frames` tensors ([frames][n_channels][L]) is ONE launch of every kernel; frame f's container lands at
  // host_mapped + f * frame_stride and its size in sizes[f].
  void encode_begin(cudaStream_t st, int slot, const int32_t* sym, const uint8_t* idx, const CdfTable& tab,
                    int n_channels, int L, int spc, uint8_t* host_mapped, size_t host_cap, int frames = 1,
                    size_t frame_stride = 0);
  size_t encode_end(cudaStream_t st, int slot, int n_channels, int L, int spc, uint8_t* host_mapped, size_t host_cap,
                    int frames = 1, size_t frame_stride = 0, size_t* sizes = nullptr);
  // container in host memory -> symbols and/or dequantised values on the device
  void invalidate_lut() { lut_for_ = nullptr; lut_rows_ = 0; packed_[0].key = packed_[1].key = nullptr; }
  void decode(cudaStream_t st, const uint8_t* bytes, size_t len, const uint8_t* idx, const CdfTable& tab,
              int n_channels, int L, int32_t* sym_out, const float* mu, const float* median, float* val_out);
  // the two halves of decode() for CR5B containers: enqueue only (false: nothing to decode) / one synchronisation + the
  // kernels' error word. Several containers staged at disjoint, 16-byte aligned offsets can share one decode_finish().
  // `frames` containers (same sub-stream count) decode in one launch into [frames][n_channels][L] outputs.
  bool decode_cr5b(cudaStream_t st, const uint8_t* const* bytes, const size_t* lens, int frames, const uint8_t* idx,
                   const CdfTable& tab, int n_channels, int L, int32_t* sym_out, const float* mu, const float* median,
                   float* val_out, size_t stage_off, bool sync_before, size_t mu_frame_extra);
  void decode_finish(cudaStream_t st);
  void reset_error(cudaStream_t st);
  size_t stage_capacity() const { return host_stage_cap_; }

 private:
  // CDF table repacked for the shared-memory kernels (uint16 rows back to back + coarse inverse table), cached per
  // table pointer: the model alternates between two tables (EntropyBottleneck, GaussianConditional)
  struct Packed {
    const int32_t* key = nullptr;
    int rows = 0, total = 0;
    bool usable = false, has_lut = false;
    uint16_t* data = nullptr;
    int32_t* row_off = nullptr;
    uint16_t* lut = nullptr;
    uint64_t stamp = 0;
  };
  const Packed* packed_for(cudaStream_t st, const CdfTable& tab);
  static constexpr int PACK_CAP = 96 * 1024;   // entries
  static constexpr int PACK_ROWS = 256;
  Packed packed_[2];
  uint64_t pack_clock_ = 0;
  size_t max_symbols_;
  int max_streams_;
  uint32_t *scratch_ = nullptr, *lengths_ = nullptr, *offsets_ = nullptr;
  uint8_t* payload_ = nullptr;
  size_t payload_cap_ = 0, scratch_words_ = 0;
  int* err_ = nullptr;
  uint32_t* host_meta_ = nullptr;  // pinned + mapped, two slots: lengths + total + err
  uint8_t* host_stage_ = nullptr;  // pinned + mapped staging (decode upload; encode into pageable caller memory)
  uint32_t* host_meta_dev_ = nullptr;  // device aliases of the two
  uint8_t* host_stage_dev_ = nullptr;
  size_t meta_slot_words_ = 0;
  uint16_t* lut_ = nullptr;        // coarse inverse-CDF table of the last GaussianConditional-style table seen
  const int32_t* lut_for_ = nullptr;
  int lut_rows_ = 0;
  size_t host_stage_cap_ = 0;
};

}  // namespace cra5
