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

// `frames` containers of ns sub-streams each, frame_words 32-bit words apart in host memory. Per frame f: the stream
// lengths go to its length table (word 6 = byte 24 of the container), its payload bytes behind that table; the payload
// size of frame f lands in meta_host[f], the coder's error word in meta_host[frames]. (The 24-byte headers are written
// by the host in encode_end.)
__global__ void container_to_host_kernel(const uint32_t* __restrict__ lengths, const uint32_t* __restrict__ offsets,
                                         const int* __restrict__ err, int ns, int frames, size_t frame_words,
                                         const uint32_t* __restrict__ payload, uint32_t* meta_host,
                                         uint32_t* containers_host) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  for (int f = 0; f < frames; ++f) {
    uint32_t* c = containers_host + (size_t)f * frame_words;
    const uint32_t base = offsets[(size_t)f * ns], end = offsets[(size_t)(f + 1) * ns];
    for (int i = gt; i < ns; i += gs) c[6 + i] = lengths[(size_t)f * ns + i];
    uint32_t* dst = c + 6 + ns;
    const uint32_t* src = payload + base / 4;
    for (uint32_t i = gt; i < (end - base) / 4; i += gs) dst[i] = src[i];
    if (gt == 0) meta_host[f] = end - base;
  }
  if (gt == 0) meta_host[frames] = (uint32_t)*err;
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
    // the kernels stage these with 16-byte copies that run into the 8-entry slack: give it defined contents
    CRA5_CUDA(cudaMemset(p.data, 0, ((size_t)PACK_CAP + 8) * 2));
    CRA5_CUDA(cudaMemset(p.lut, 0, ((size_t)PACK_ROWS * 257 + 8) * 2));
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
    rans_encode(st, sym, idx, idx == nullptr ? (n_channels > 0 ? n_channels : 1) : 0, tab.cdf, tab.cols, tab.length, tab.offset, 1,
                (int)n, 1, L > 0 ? L : 1, scratch_, cap_words, lengths_, offsets_, payload_, err_);
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
  size_t total = 0;
  encode_end(st, 0, n_channels, L, spc, host_stage_, host_stage_cap_, 1, 0, &total);
  CRA5_CHECK(host_cap >= total, ERR_INVALID, "rans_encode: output buffer too small");
  memcpy(host_out, host_stage_, total);
  return total;
}

// Enqueue the encode of `frames` tensors of n_channels x L symbols each (contiguous: [frames][n_channels][L]); their
// containers land in `host_mapped` (pinned + mapped), frame f at byte f * frame_stride (each slot at least
// max_container_bytes() large), once the stream has been synchronised and encode_end() has written the headers. The
// batch is ONE launch of every kernel: frames * n_channels * spc sub-streams.
void RansCoder::encode_begin(cudaStream_t st, int slot, const int32_t* sym, const uint8_t* idx, const CdfTable& tab,
                             int n_channels, int L, int spc, uint8_t* host_mapped, size_t host_cap, int frames,
                             size_t frame_stride) {
  CRA5_CHECK(tab.ready(), ERR_STATE, "Uninitialized CDFs. Run update() first");
  CRA5_CHECK(spc >= 1 && spc <= CR5B_MAX_SPC, ERR_INVALID, "streams per channel must be in [1, 64]");
  CRA5_CHECK(n_channels >= 0 && L >= 0 && (slot == 0 || slot == 1) && frames >= 1, ERR_INVALID, "rans_encode: bad argument");
  const int ns = n_channels * spc;                 // sub-streams per frame
  const size_t per_frame = max_container_bytes((size_t)n_channels * L, ns);
  if (frames == 1) frame_stride = host_cap;
  CRA5_CHECK((size_t)frames * n_channels * L <= max_symbols_ && (size_t)frames * ns <= (size_t)max_streams_, ERR_INVALID,
             "rans_encode: tensor larger than the coder was sized for");
  CRA5_CHECK((frame_stride & 3) == 0 && frame_stride >= per_frame && host_cap >= (frames - 1) * frame_stride + per_frame,
             ERR_INVALID, "rans_encode: output buffer too small");
  const int count_max = (L + spc - 1) / spc;
  const int cap_words = 2 * count_max + 6;
  CRA5_CHECK((size_t)frames * ns * cap_words <= scratch_words_, ERR_INTERNAL, "rans_encode: scratch sizing");
  if (ns == 0 || L == 0) return;
  const int chan_mod = (idx == nullptr) ? n_channels : 0;   // EntropyBottleneck: CDF row = channel within the frame
  if (const Packed* pk = packed_for(st, tab))
    rans_encode_smem(st, sym, idx, chan_mod, pk->data, pk->row_off, tab.length, tab.offset, pk->rows, pk->total,
                     frames * n_channels, L, spc, L, scratch_, cap_words, lengths_, offsets_, payload_, err_);
  else
    rans_encode(st, sym, idx, chan_mod, tab.cdf, tab.cols, tab.length, tab.offset, frames * n_channels, L, spc, L > 0 ? L : 1,
                scratch_, cap_words, lengths_, offsets_, payload_, err_);
  count_launch();
  container_to_host_kernel<<<XFER_BLOCKS, XFER_THREADS, 0, st>>>(
      lengths_, offsets_, err_, ns, frames, frame_stride / 4, reinterpret_cast<const uint32_t*>(payload_),
      host_meta_dev_ + slot * meta_slot_words_, static_cast<uint32_t*>(device_alias(host_mapped)));
  CRA5_CUDA(cudaGetLastError());
}

// after the stream synchronisation: checks the error word, writes the 24-byte header of every frame's container and
// stores the container sizes in sizes[frames]; returns the size of the first
size_t RansCoder::encode_end(cudaStream_t st, int slot, int n_channels, int L, int spc, uint8_t* host_mapped,
                             size_t host_cap, int frames, size_t frame_stride, size_t* sizes) {
  const int ns = n_channels * spc;
  const size_t head = CR5B_HEADER + 4 * (size_t)ns;
  uint32_t* meta = host_meta_ + slot * meta_slot_words_;
  if (frames == 1) frame_stride = host_cap;
  const bool coded = ns > 0 && L > 0;
  if (coded && meta[frames] != 0) {
    CRA5_CUDA(cudaMemsetAsync(err_, 0, sizeof(int), st));
    throw Error(ERR_INTERNAL, "rans_encode: per-stream scratch overflow");
  }
  size_t first = 0;
  for (int f = 0; f < frames; ++f) {
    uint8_t* c = host_mapped + (size_t)f * frame_stride;
    const uint32_t total = coded ? meta[f] : 0;
    CRA5_CHECK(frame_stride >= head + total, ERR_INVALID, "rans_encode: output buffer too small");
    memcpy(c, "CR5B", 4);
    c[4] = 1;
    c[5] = 0;
    c[6] = c[7] = 0;
    put_u32(c + 8, (uint32_t)n_channels);
    put_u32(c + 12, (uint32_t)L);
    put_u32(c + 16, (uint32_t)spc);
    put_u32(c + 20, (uint32_t)ns);
    if (!coded) memset(c + CR5B_HEADER, 0, (size_t)ns * 4);
    if (sizes != nullptr) sizes[f] = head + total;
    if (f == 0) first = head + total;
  }
  return first;
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
    rans_decode(st, payload_, offsets_, idx, idx == nullptr ? (n_channels > 0 ? n_channels : 1) : 0, tab.cdf, tab.cols,
                tab.length, tab.offset, nullptr, 0, 1, (int)n, 1, L > 0 ? L : 1, sym_out, mu, median, val_out, err_);
    CRA5_CUDA(cudaMemcpyAsync(host_meta_, err_, 4, cudaMemcpyDeviceToHost, st));
    CRA5_CUDA(cudaStreamSynchronize(st));
    if (host_meta_[0] != 0) {
      CRA5_CUDA(cudaMemsetAsync(err_, 0, sizeof(int), st));
      throw Error(ERR_BITSTREAM, "bitstream: stream exhausted while decoding (corrupt data or wrong tensor shape)");
    }
    return;
  }
  if (decode_cr5b(st, &bytes, &len, 1, idx, tab, n_channels, L, sym_out, mu, median, val_out, 0, true, 0)) decode_finish(st);
}

// CR5B containers of `frames` tensors (n_channels x L symbols each, all with the same sub-stream count) -> device:
// validates every header, stages the length tables of all frames followed by their payloads in the pinned buffer at
// `stage_off` (a multiple of 16) and enqueues ONE upload, ONE length scan and ONE decode launch for the whole batch.
// Outputs are [frames][n_channels][L]; with per-symbol means, frame f reads mu[pos + f * mu_frame_extra]. Returns false
// when there is nothing to decode (no GPU work enqueued). With sync_before the stream is drained first, because the
// staging buffer may still be in flight from a previous call; a caller that stages two batches at disjoint offsets
// inside one call passes false for the second (Model::bin_to_latent). Errors found by the kernels are collected by
// decode_finish().
bool RansCoder::decode_cr5b(cudaStream_t st, const uint8_t* const* bytes, const size_t* lens, int frames,
                            const uint8_t* idx, const CdfTable& tab, int n_channels, int L, int32_t* sym_out,
                            const float* mu, const float* median, float* val_out, size_t stage_off, bool sync_before,
                            size_t mu_frame_extra) {
  CRA5_CHECK(tab.ready(), ERR_STATE, "Uninitialized CDFs. Run update() first");
  CRA5_CHECK((stage_off & 15) == 0 && frames >= 1, ERR_INTERNAL, "decode: staging offset / frames");
  uint32_t spc = 0, ns = 0;
  uint64_t total_all = 0;
  for (int f = 0; f < frames; ++f) {
    const uint8_t* b = bytes[f];
    const size_t len = lens[f];
    CRA5_CHECK(b != nullptr && len >= 8 && memcmp(b, "CR5B", 4) == 0, ERR_BITSTREAM, "bitstream: not a CR5B container");
    CRA5_CHECK(len >= CR5B_HEADER, ERR_BITSTREAM, "bitstream: truncated header");
    CRA5_CHECK(b[4] == 1, ERR_BITSTREAM, "bitstream: unsupported version");
    const uint32_t nc = get_u32(b + 8), l = get_u32(b + 12), spc_f = get_u32(b + 16), ns_f = get_u32(b + 20);
    CRA5_CHECK(nc == (uint32_t)n_channels && l == (uint32_t)L, ERR_BITSTREAM,
               "bitstream: tensor shape does not match the model");
    CRA5_CHECK(spc_f >= 1 && spc_f <= (uint32_t)CR5B_MAX_SPC && ns_f == nc * spc_f, ERR_BITSTREAM, "bitstream: bad stream count");
    if (f == 0) { spc = spc_f; ns = ns_f; }
    CRA5_CHECK(spc_f == spc, ERR_BITSTREAM, "bitstream: the containers of one batch must share their sub-stream count");
    const size_t head = CR5B_HEADER + 4 * (size_t)ns;
    CRA5_CHECK(len >= head, ERR_BITSTREAM, "bitstream: truncated length table");
    uint64_t total = 0;
    for (uint32_t s = 0; s < ns; ++s) {
      const uint32_t ls = get_u32(b + CR5B_HEADER + 4 * (size_t)s);
      CRA5_CHECK((ls & 3) == 0 && (l == 0 || ls >= 8), ERR_BITSTREAM, "bitstream: bad sub-stream length");
      total += ls;
    }
    CRA5_CHECK(head + total == len, ERR_BITSTREAM, "bitstream: payload size mismatch");
    total_all += total;
  }
  CRA5_CHECK((size_t)frames * n_channels * L <= max_symbols_ && (size_t)frames * ns <= (size_t)max_streams_, ERR_BITSTREAM,
             "bitstream: too large");
  if (ns == 0 || L == 0) return false;
  const size_t table_bytes = (size_t)frames * ns * 4;
  CRA5_CHECK(total_all <= payload_cap_ && stage_off + table_bytes + total_all <= host_stage_cap_, ERR_BITSTREAM,
             "bitstream: too large");
  if (sync_before) CRA5_CUDA(cudaStreamSynchronize(st));  // the staging buffer may still be in flight from a previous call
  uint8_t* stage = host_stage_ + stage_off;
  size_t pay = table_bytes;
  for (int f = 0; f < frames; ++f) {
    const size_t head = CR5B_HEADER + 4 * (size_t)ns;
    memcpy(stage + (size_t)f * ns * 4, bytes[f] + CR5B_HEADER, (size_t)ns * 4);
    memcpy(stage + pay, bytes[f] + head, lens[f] - head);
    pay += lens[f] - head;
  }
  const int ns_all = frames * (int)ns, nc_all = frames * n_channels;
  count_launch();
  container_from_host_kernel<<<XFER_BLOCKS, XFER_THREADS, 0, st>>>(
      reinterpret_cast<const uint32_t*>(host_stage_dev_ + stage_off), ns_all, (uint32_t)(ns_all + total_all / 4), lengths_,
      reinterpret_cast<uint32_t*>(payload_));
  CRA5_CUDA(cudaGetLastError());
  scan_lengths(st, lengths_, ns_all, offsets_);
  const int chan_mod = (idx == nullptr) ? n_channels : 0;
  if (const Packed* pk = packed_for(st, tab)) {
    rans_decode_smem(st, payload_, offsets_, idx, chan_mod, pk->data, pk->row_off, tab.length, tab.offset,
                     pk->has_lut ? pk->lut : nullptr, pk->rows, pk->total, nc_all, L, (int)spc, L, sym_out, mu, median,
                     val_out, err_, n_channels, mu_frame_extra);
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
    rans_decode(st, payload_, offsets_, idx, chan_mod, tab.cdf, tab.cols, tab.length, tab.offset, lut, lut_rows, nc_all, L,
                (int)spc, L > 0 ? L : 1, sym_out, mu, median, val_out, err_, n_channels, mu_frame_extra);
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
