#include "coder.h"

#include <string.h>

#include "host_util.h"
#include "kernels.h"

namespace cra5 {

// ---------------------------------------------------------------------------------------------------------------
// Small transfers between the device and MAPPED pinned host memory are done by the SMs (zero-copy loads / stores over
// PCIe), not by cudaMemcpyAsync: a copy-engine transfer queues FIFO behind whatever that engine is already moving, and
// in a streaming pipeline that is the 1.1 GB frame of the next / previous time step (20 ms) -- the 2-5 MB bitstream
// would wait for it and serialise encode, decode and both frame copies.
namespace {
constexpr int XFER_BLOCKS = 64, XFER_THREADS = 256;

// lengths[n_streams], total, err -> meta_host; payload[0, total) -> payload_host   (total = offsets[n_streams])
__global__ void container_to_host_kernel(const uint32_t* __restrict__ lengths, const uint32_t* __restrict__ offsets,
                                         const int* __restrict__ err, int n_streams,
                                         const uint32_t* __restrict__ payload, uint32_t* meta_host,
                                         uint32_t* payload_host) {
  const uint32_t total = offsets[n_streams];
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  for (int i = gt; i < n_streams; i += gs) meta_host[i] = lengths[i];
  if (gt == 0) {
    meta_host[n_streams] = total;
    meta_host[n_streams + 1] = (uint32_t)*err;
  }
  if (payload_host != nullptr)
    for (uint32_t i = gt; i < total / 4; i += gs) payload_host[i] = payload[i];
}
// stage_host = [lengths: ns words][payload: words] -> lengths, payload on the device
__global__ void container_from_host_kernel(const uint32_t* stage_host, int ns, uint32_t n_words, uint32_t* lengths,
                                           uint32_t* payload) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  for (uint32_t i = gt; i < n_words; i += gs) {
    const uint32_t v = stage_host[i];
    if (i < (uint32_t)ns) lengths[i] = v; else payload[i - ns] = v;
  }
}
__global__ void word_to_host_kernel(const int* src, uint32_t* dst_host) { *dst_host = (uint32_t)*src; }

void* device_alias(void* host_mapped) {
  void* d = nullptr;
  CRA5_CUDA(cudaHostGetDevicePointer(&d, host_mapped, 0));
  return d;
}
}  // namespace

RansCoder::RansCoder(size_t max_symbols, int max_channels)
    : max_symbols_(max_symbols), max_streams_(max_channels * CR5B_MAX_SPC) {
  scratch_words_ = 2 * max_symbols + 8 * (size_t)max_streams_;
  payload_cap_ = scratch_words_ * 4;
  CRA5_CUDA(cudaMalloc(&scratch_, scratch_words_ * 4));
  CRA5_CUDA(cudaMalloc(&lengths_, (size_t)max_streams_ * 4));
  CRA5_CUDA(cudaMalloc(&offsets_, ((size_t)max_streams_ + 1) * 4));
  CRA5_CUDA(cudaMalloc(&payload_, payload_cap_));
  CRA5_CUDA(cudaMalloc(&err_, 2 * sizeof(int)));   // [0] error flag, [1] scratch word of packed_for()
  CRA5_CUDA(cudaMemset(err_, 0, 2 * sizeof(int)));
  meta_slot_words_ = (size_t)max_streams_ + 4;
  CRA5_CUDA(cudaHostAlloc(&host_meta_, 2 * meta_slot_words_ * 4, cudaHostAllocMapped));
  host_stage_cap_ = payload_cap_ + (size_t)max_streams_ * 4 + 64;
  CRA5_CUDA(cudaHostAlloc(&host_stage_, host_stage_cap_, cudaHostAllocMapped));
  host_meta_dev_ = static_cast<uint32_t*>(device_alias(host_meta_));
  host_stage_dev_ = static_cast<uint8_t*>(device_alias(host_stage_));
  CRA5_CUDA(cudaMalloc(&lut_, (size_t)256 * 257 * 2));
  for (Packed& p : packed_) {
    CRA5_CUDA(cudaMalloc(&p.data, ((size_t)PACK_CAP + 8) * 2));
    CRA5_CUDA(cudaMalloc(&p.row_off, ((size_t)PACK_ROWS + 1) * 4));
    CRA5_CUDA(cudaMalloc(&p.lut, ((size_t)PACK_ROWS * 257 + 8) * 2));
  }
}

// the table in shared-memory form, or nullptr when it is not known to fit (row count unknown / too large)
const RansCoder::Packed* RansCoder::packed_for(cudaStream_t st, const CdfTable& tab) {
  if (tab.rows <= 0 || tab.rows > PACK_ROWS) return nullptr;
  Packed* slot = nullptr;
  for (Packed& p : packed_)
    if (p.key == tab.cdf && p.rows == tab.rows) slot = &p;
  if (slot == nullptr) {
    slot = (packed_[0].stamp <= packed_[1].stamp) ? &packed_[0] : &packed_[1];
    slot->key = tab.cdf;
    slot->rows = tab.rows;
    int* total_dev = err_ + 1;
    pack_cdf(st, tab.cdf, tab.cols, tab.length, tab.rows, slot->row_off, slot->data, PACK_CAP, total_dev);
    slot->has_lut = tab.cols > 64;   // wide rows (GaussianConditional): coarse inverse table for the decoder
    if (slot->has_lut) build_decode_lut(st, tab.cdf, tab.cols, tab.length, tab.rows, slot->lut);
    int total = 0;
    CRA5_CUDA(cudaMemcpyAsync(&total, total_dev, sizeof(int), cudaMemcpyDeviceToHost, st));   // once per table
    CRA5_CUDA(cudaStreamSynchronize(st));
    slot->total = total;
    slot->usable = total <= PACK_CAP && rans_tables_fit(tab.rows, total, slot->has_lut);
  }
  slot->stamp = ++pack_clock_;
  return slot->usable ? slot : nullptr;
}

RansCoder::~RansCoder() {
  cudaFree(scratch_);
  cudaFree(lengths_);
  cudaFree(offsets_);
  cudaFree(payload_);
  cudaFree(err_);
  cudaFreeHost(host_meta_);
  cudaFreeHost(host_stage_);
  cudaFree(lut_);
  for (Packed& p : packed_) {
    cudaFree(p.data);
    cudaFree(p.row_off);
    cudaFree(p.lut);
  }
}

static void put_u32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
static uint32_t get_u32(const uint8_t* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}

size_t RansCoder::encode(cudaStream_t st, const int32_t* sym, const uint8_t* idx, const CdfTable& tab, int n_channels,
                         int L, int spc, uint8_t* host_out, size_t host_cap) {
  CRA5_CHECK(tab.ready(), ERR_STATE, "Uninitialized CDFs. Run update() first");
  CRA5_CHECK(spc >= 0 && spc <= CR5B_MAX_SPC, ERR_INVALID, "streams per channel must be in [0, 64]");
  CRA5_CHECK(n_channels >= 0 && L >= 0, ERR_INVALID, "rans_encode: negative size");
  if (spc == 0) {
    // reference format: the whole tensor as ONE sequential stream, no container -- byte-identical to what
    // RansEncoder.encode_with_indexes returns (rans_interface.cpp:202-213). One GPU thread; interop, not throughput.
    const size_t n = (size_t)n_channels * L;
    CRA5_CHECK(n <= max_symbols_ && n < (size_t)1 << 30, ERR_INVALID, "rans_encode: tensor larger than the coder was sized for");
    const int cap_words = (int)(2 * n + 6);
    CRA5_CHECK((size_t)cap_words <= scratch_words_, ERR_INTERNAL, "rans_encode: scratch sizing");
    rans_encode(st, sym, idx, idx == nullptr, tab.cdf, tab.cols, tab.length, tab.offset, 1, (int)n, 1, L > 0 ? L : 1,
                scratch_, cap_words, lengths_, offsets_, payload_, err_);
    CRA5_CUDA(cudaMemcpyAsync(host_meta_, lengths_, 4, cudaMemcpyDeviceToHost, st));
    CRA5_CUDA(cudaMemcpyAsync(host_meta_ + 1, err_, 4, cudaMemcpyDeviceToHost, st));
    CRA5_CUDA(cudaStreamSynchronize(st));
    if (host_meta_[1] != 0) {
      CRA5_CUDA(cudaMemsetAsync(err_, 0, sizeof(int), st));
      throw Error(ERR_INTERNAL, "rans_encode: scratch overflow");
    }
    const uint32_t total = host_meta_[0];
    CRA5_CHECK(host_cap >= total, ERR_INVALID, "rans_encode: output buffer too small");
    CRA5_CUDA(cudaMemcpyAsync(host_out, payload_, total, cudaMemcpyDeviceToHost, st));
    CRA5_CUDA(cudaStreamSynchronize(st));
    return total;
  }
  // chunked container through caller memory that the device may not be able to reach: stage, then memcpy
  const int n_streams = n_channels * spc;
  const size_t head = CR5B_HEADER + 4 * (size_t)n_streams;
  CRA5_CHECK(host_cap >= head, ERR_INVALID, "rans_encode: output buffer too small");
  encode_begin(st, 0, sym, idx, tab, n_channels, L, spc, host_stage_, host_stage_cap_);
  CRA5_CUDA(cudaStreamSynchronize(st));
  const size_t total = encode_end(st, 0, n_channels, L, spc, host_stage_, host_stage_cap_);
  CRA5_CHECK(host_cap >= total, ERR_INVALID, "rans_encode: output buffer too small");
  memcpy(host_out, host_stage_, total);
  return total;
}

// Enqueue the encode of one tensor; its container lands in `host_mapped` (pinned + mapped, at least
// max_container_bytes() large) once the stream has been synchronised and encode_end() has written the header.
void RansCoder::encode_begin(cudaStream_t st, int slot, const int32_t* sym, const uint8_t* idx, const CdfTable& tab,
                             int n_channels, int L, int spc, uint8_t* host_mapped, size_t host_cap) {
  CRA5_CHECK(tab.ready(), ERR_STATE, "Uninitialized CDFs. Run update() first");
  CRA5_CHECK(spc >= 1 && spc <= CR5B_MAX_SPC, ERR_INVALID, "streams per channel must be in [1, 64]");
  CRA5_CHECK(n_channels >= 0 && L >= 0 && (slot == 0 || slot == 1), ERR_INVALID, "rans_encode: bad argument");
  const int n_streams = n_channels * spc;
  CRA5_CHECK((size_t)n_channels * L <= max_symbols_ && n_streams <= max_streams_, ERR_INVALID,
             "rans_encode: tensor larger than the coder was sized for");
  CRA5_CHECK(host_cap >= max_container_bytes((size_t)n_channels * L, n_streams), ERR_INVALID,
             "rans_encode: output buffer too small");
  const size_t head = CR5B_HEADER + 4 * (size_t)n_streams;
  const int count_max = (L + spc - 1) / spc;
  const int cap_words = 2 * count_max + 6;
  CRA5_CHECK((size_t)n_streams * cap_words <= scratch_words_, ERR_INTERNAL, "rans_encode: scratch sizing");
  if (n_streams == 0 || L == 0) return;
  if (const Packed* pk = packed_for(st, tab))
    rans_encode_smem(st, sym, idx, idx == nullptr, pk->data, pk->row_off, tab.length, tab.offset, pk->rows, pk->total,
                     n_channels, L, spc, L, scratch_, cap_words, lengths_, offsets_, payload_, err_);
  else
    rans_encode(st, sym, idx, idx == nullptr, tab.cdf, tab.cols, tab.length, tab.offset, n_channels, L, spc, L > 0 ? L : 1,
                scratch_, cap_words, lengths_, offsets_, payload_, err_);
  uint8_t* out_dev = static_cast<uint8_t*>(device_alias(host_mapped)) + head;   // 4-byte aligned: head is
  count_launch();
  container_to_host_kernel<<<XFER_BLOCKS, XFER_THREADS, 0, st>>>(
      lengths_, offsets_, err_, n_streams, reinterpret_cast<const uint32_t*>(payload_),
      host_meta_dev_ + slot * meta_slot_words_, reinterpret_cast<uint32_t*>(out_dev));
  CRA5_CUDA(cudaGetLastError());
}

size_t RansCoder::encode_end(cudaStream_t st, int slot, int n_channels, int L, int spc, uint8_t* host_mapped,
                             size_t host_cap) {
  const int n_streams = n_channels * spc;
  const size_t head = CR5B_HEADER + 4 * (size_t)n_streams;
  uint32_t* meta = host_meta_ + slot * meta_slot_words_;
  uint32_t total = 0;
  if (n_streams > 0 && L > 0) {
    if (meta[n_streams + 1] != 0) {
      CRA5_CUDA(cudaMemsetAsync(err_, 0, sizeof(int), st));
      throw Error(ERR_INTERNAL, "rans_encode: per-stream scratch overflow");
    }
    total = meta[n_streams];
  } else {
    for (int s = 0; s < n_streams; ++s) meta[s] = 0;
  }
  CRA5_CHECK(host_cap >= head + total, ERR_INVALID, "rans_encode: output buffer too small");
  memcpy(host_mapped, "CR5B", 4);
  host_mapped[4] = 1;
  host_mapped[5] = 0;
  host_mapped[6] = host_mapped[7] = 0;
  put_u32(host_mapped + 8, (uint32_t)n_channels);
  put_u32(host_mapped + 12, (uint32_t)L);
  put_u32(host_mapped + 16, (uint32_t)spc);
  put_u32(host_mapped + 20, (uint32_t)n_streams);
  memcpy(host_mapped + CR5B_HEADER, meta, (size_t)n_streams * 4);
  return head + total;
}

void RansCoder::decode(cudaStream_t st, const uint8_t* bytes, size_t len, const uint8_t* idx, const CdfTable& tab,
                       int n_channels, int L, int32_t* sym_out, const float* mu, const float* median, float* val_out) {
  CRA5_CHECK(tab.ready(), ERR_STATE, "Uninitialized CDFs. Run update() first");
  CRA5_CHECK(bytes != nullptr && len >= 8, ERR_BITSTREAM, "bitstream: truncated");
  if (memcmp(bytes, "CR5B", 4) != 0) {
    // no container magic: a reference-format stream (one sequential rANS stream per tensor, as CRA5 .bin archives
    // written by the PyTorch reference hold) -- decoded by a single thread
    const size_t n = (size_t)n_channels * L;
    CRA5_CHECK((len & 3) == 0, ERR_BITSTREAM, "bitstream: reference stream length must be a multiple of 4");
    CRA5_CHECK(n <= max_symbols_ && len <= payload_cap_ && len <= host_stage_cap_, ERR_BITSTREAM, "bitstream: too large");
    if (n == 0) return;
    CRA5_CUDA(cudaStreamSynchronize(st));
    memcpy(host_stage_, bytes, len);
    uint32_t* offs = reinterpret_cast<uint32_t*>(host_stage_ + ((len + 15) & ~size_t(15)));
    offs[0] = 0;
    offs[1] = (uint32_t)len;
    CRA5_CUDA(cudaMemcpyAsync(payload_, host_stage_, len, cudaMemcpyHostToDevice, st));
    CRA5_CUDA(cudaMemcpyAsync(offsets_, offs, 8, cudaMemcpyHostToDevice, st));
    rans_decode(st, payload_, offsets_, idx, idx == nullptr, tab.cdf, tab.cols, tab.length, tab.offset, nullptr, 0, 1,
                (int)n, 1, L > 0 ? L : 1, sym_out, mu, median, val_out, err_);
    CRA5_CUDA(cudaMemcpyAsync(host_meta_, err_, 4, cudaMemcpyDeviceToHost, st));
    CRA5_CUDA(cudaStreamSynchronize(st));
    if (host_meta_[0] != 0) {
      CRA5_CUDA(cudaMemsetAsync(err_, 0, sizeof(int), st));
      throw Error(ERR_BITSTREAM, "bitstream: stream exhausted while decoding (corrupt data or wrong tensor shape)");
    }
    return;
  }
  if (decode_cr5b(st, bytes, len, idx, tab, n_channels, L, sym_out, mu, median, val_out, 0, true)) decode_finish(st);
}

// CR5B container -> device: validates the header, stages lengths + payload in the pinned buffer at `stage_off` (a
// multiple of 16) and enqueues the upload, the length scan and the decode kernel. Returns false when there is nothing
// to decode (no GPU work enqueued). With sync_before the stream is drained first, because the staging buffer may still
// be in flight from a previous call; a caller that stages two containers at disjoint offsets inside one call passes
// false for the second (Model::bin_to_latent). Errors found by the kernels are collected by decode_finish().
bool RansCoder::decode_cr5b(cudaStream_t st, const uint8_t* bytes, size_t len, const uint8_t* idx, const CdfTable& tab,
                            int n_channels, int L, int32_t* sym_out, const float* mu, const float* median,
                            float* val_out, size_t stage_off, bool sync_before) {
  CRA5_CHECK(tab.ready(), ERR_STATE, "Uninitialized CDFs. Run update() first");
  CRA5_CHECK(bytes != nullptr && len >= 8 && memcmp(bytes, "CR5B", 4) == 0, ERR_BITSTREAM, "bitstream: not a CR5B container");
  CRA5_CHECK((stage_off & 15) == 0, ERR_INTERNAL, "decode: staging offset");
  CRA5_CHECK(len >= CR5B_HEADER, ERR_BITSTREAM, "bitstream: truncated header");
  CRA5_CHECK(bytes[4] == 1, ERR_BITSTREAM, "bitstream: unsupported version");
  const uint32_t nc = get_u32(bytes + 8), l = get_u32(bytes + 12), spc = get_u32(bytes + 16),
                 ns = get_u32(bytes + 20);
  CRA5_CHECK(nc == (uint32_t)n_channels && l == (uint32_t)L, ERR_BITSTREAM,
             "bitstream: tensor shape does not match the model");
  CRA5_CHECK(spc >= 1 && spc <= (uint32_t)CR5B_MAX_SPC && ns == nc * spc, ERR_BITSTREAM, "bitstream: bad stream count");
  CRA5_CHECK((size_t)nc * l <= max_symbols_ && (int)ns <= max_streams_, ERR_BITSTREAM, "bitstream: too large");
  const size_t head = CR5B_HEADER + 4 * (size_t)ns;
  CRA5_CHECK(len >= head, ERR_BITSTREAM, "bitstream: truncated length table");
  uint64_t total = 0;
  for (uint32_t s = 0; s < ns; ++s) {
    const uint32_t ls = get_u32(bytes + CR5B_HEADER + 4 * (size_t)s);
    CRA5_CHECK((ls & 3) == 0 && (l == 0 || ls >= 8), ERR_BITSTREAM, "bitstream: bad sub-stream length");
    total += ls;
  }
  CRA5_CHECK(head + total == len, ERR_BITSTREAM, "bitstream: payload size mismatch");
  if (ns == 0 || l == 0) return false;
  CRA5_CHECK(total <= payload_cap_ && stage_off + len <= host_stage_cap_, ERR_BITSTREAM, "bitstream: too large");
  if (sync_before) CRA5_CUDA(cudaStreamSynchronize(st));  // the staging buffer may still be in flight from a previous call
  memcpy(host_stage_ + stage_off, bytes + CR5B_HEADER, len - CR5B_HEADER);
  count_launch();
  container_from_host_kernel<<<XFER_BLOCKS, XFER_THREADS, 0, st>>>(
      reinterpret_cast<const uint32_t*>(host_stage_dev_ + stage_off), (int)ns, (uint32_t)(ns + total / 4), lengths_,
      reinterpret_cast<uint32_t*>(payload_));
  CRA5_CUDA(cudaGetLastError());
  scan_lengths(st, lengths_, (int)ns, offsets_);
  if (const Packed* pk = packed_for(st, tab)) {
    rans_decode_smem(st, payload_, offsets_, idx, idx == nullptr, pk->data, pk->row_off, tab.length, tab.offset,
                     pk->has_lut ? pk->lut : nullptr, pk->rows, pk->total, n_channels, L, (int)spc, L, sym_out, mu, median,
                     val_out, err_);
  } else {
    // wide tables (GaussianConditional: up to 3133 entries per row) get a coarse inverse table; the per-channel
    // EntropyBottleneck rows are a few dozen entries and are searched directly
    const uint16_t* lut = nullptr;
    int lut_rows = 0;
    if (idx != nullptr && tab.rows <= 256 && tab.cols > 64) {
      if (lut_for_ != tab.cdf || lut_rows_ != tab.rows) {
        build_decode_lut(st, tab.cdf, tab.cols, tab.length, tab.rows, lut_);
        lut_for_ = tab.cdf;
        lut_rows_ = tab.rows;
      }
      lut = lut_;
      lut_rows = tab.rows;
    }
    rans_decode(st, payload_, offsets_, idx, idx == nullptr, tab.cdf, tab.cols, tab.length, tab.offset, lut, lut_rows,
                n_channels, L, (int)spc, L > 0 ? L : 1, sym_out, mu, median, val_out, err_);
  }
  return true;
}

// one synchronisation for everything decode_cr5b() enqueued: fetches the kernels' error word
void RansCoder::decode_finish(cudaStream_t st) {
  word_to_host_kernel<<<1, 1, 0, st>>>(err_, host_meta_dev_);
  CRA5_CUDA(cudaGetLastError());
  CRA5_CUDA(cudaStreamSynchronize(st));
  if (host_meta_[0] != 0) {
    CRA5_CUDA(cudaMemsetAsync(err_, 0, sizeof(int), st));
    throw Error(ERR_BITSTREAM, "bitstream: sub-stream exhausted while decoding (corrupt data)");
  }
}

void RansCoder::reset_error(cudaStream_t st) { cudaMemsetAsync(err_, 0, sizeof(int), st); }

}  // namespace cra5
