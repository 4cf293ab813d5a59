"""`compressai.ans` (pybind11 module of the reference, cpp_exts/rans/rans_interface.cpp:361-381) over libcra5b200.so.

Same classes, same method signatures (Python lists in, `bytes` / list out, arguments copied per call), same bytes: the
C ABI's coder is run in the reference's single-sequential-stream format (`spc = 0`), which is byte-identical to
RansEncoder.encode_with_indexes for the same symbols, indexes and tables (tests/test_gpu_entropy.py, tests/golden/
rans_kat.json). The coding itself runs on the GPU (one thread per call: this is the interoperability path, not the
fast one -- `VAEformer.compress` uses the chunk-parallel CR5B container instead).

Additive: every list argument may also be a torch tensor (device tensors are used in place, no `.tolist()` round trip).
Limits (stated, not silent): at most 256 distinct CDF rows per stream (scale indexes travel as uint8); a
BufferedRansEncoder that is fed different tables between flushes concatenates their rows up to that limit.
"""
import ctypes

import torch

from cra5_b200 import _lib


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("compressai.ans (cra5_b200 drop-in) needs a CUDA device: the coder is a GPU kernel, "
                           "there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _i32(v, dev):
    if isinstance(v, torch.Tensor):
        return v.to(device=dev, dtype=torch.int32).contiguous()
    return torch.tensor(v, dtype=torch.int32, device=dev)


def _table(cdfs, cdfs_sizes, offsets, dev):
    """list of (ragged) rows -> padded int32 [rows][cols] + lengths + offsets on the device"""
    if isinstance(cdfs, torch.Tensor):
        cdf = cdfs.to(device=dev, dtype=torch.int32).contiguous()
    else:
        cols = max(len(r) for r in cdfs)
        cdf = torch.zeros((len(cdfs), cols), dtype=torch.int32)
        for i, r in enumerate(cdfs):
            cdf[i, : len(r)] = torch.tensor(r, dtype=torch.int32)
        cdf = cdf.to(dev)
    if cdf.dim() != 2 or cdf.shape[0] > 256:
        raise ValueError(f"Invalid CDF size {tuple(cdf.shape)} (2-D, at most 256 rows)")
    return cdf, _i32(cdfs_sizes, dev).reshape(-1), _i32(offsets, dev).reshape(-1)


def _encode(symbols, indexes, cdf, sizes, offs, dev):
    sym = _i32(symbols, dev).reshape(-1)
    idx = _i32(indexes, dev).reshape(-1)
    n = sym.numel()
    if idx.numel() != n:
        raise ValueError("symbols and indexes must have the same length")
    if n == 0:          # the reference returns the flushed initial state: 2^31 as two words, low word first
        return (1 << 31).to_bytes(8, "little")
    cap = 16 * n + 64             # worst case: every symbol a bypass value with the maximum number of nibbles
    out = (ctypes.c_uint8 * cap)()
    ln = ctypes.c_uint64()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.cra5_op_rans_encode_table(
            _lib.ptr(sym), _lib.ptr(idx.to(torch.uint8)), _lib.ptr(cdf), int(cdf.shape[0]), int(cdf.shape[1]),
            _lib.ptr(sizes), _lib.ptr(offs), 1, n, 0, out, ctypes.c_uint64(cap), ctypes.byref(ln), _lib.stream_ptr()))
    return bytes(out[: ln.value])


def _decode(encoded, indexes, cdf, sizes, offs, dev):
    idx = _i32(indexes, dev).reshape(-1)
    n = idx.numel()
    if n == 0:
        return []
    sym = torch.empty(n, dtype=torch.int32, device=dev)
    encoded = bytes(encoded)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.cra5_op_rans_decode_table(
            encoded, ctypes.c_uint64(len(encoded)), _lib.ptr(idx.to(torch.uint8)), _lib.ptr(cdf), int(cdf.shape[0]),
            int(cdf.shape[1]), _lib.ptr(sizes), _lib.ptr(offs), 1, n, _lib.ptr(sym), _lib.stream_ptr()))
        torch.cuda.synchronize(dev)
    return sym.tolist()


class RansEncoder:
    """rans_interface.cpp:202-213"""

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets) -> bytes:
        dev = _dev()
        return _encode(symbols, indexes, *_table(cdfs, cdfs_sizes, offsets, dev), dev)


class _Accumulated:
    """symbols / indexes of several calls against possibly different tables, merged into one table by stacking rows"""

    def __init__(self):
        self.clear()

    def clear(self):
        self.sym, self.idx, self.tables, self.rows = [], [], [], 0

    def add(self, symbols, indexes, cdfs, cdfs_sizes, offsets, dev):
        cdf, sizes, offs = _table(cdfs, cdfs_sizes, offsets, dev)
        base = None
        for b, (c, s, o) in self.tables:       # same table as an earlier call: reuse its rows
            if c.shape == cdf.shape and torch.equal(c, cdf) and torch.equal(s, sizes) and torch.equal(o, offs):
                base = b
                break
        if base is None:
            base = self.rows
            self.tables.append((base, (cdf, sizes, offs)))
            self.rows += cdf.shape[0]
            if self.rows > 256:
                raise ValueError("more than 256 distinct CDF rows between flushes")
        self.sym.append(_i32(symbols, dev).reshape(-1))
        self.idx.append(_i32(indexes, dev).reshape(-1) + base)

    def merged(self, dev):
        cols = max(t[0].shape[1] for _, t in self.tables)
        cdf = torch.zeros((self.rows, cols), dtype=torch.int32, device=dev)
        for b, (c, _, _) in self.tables:
            cdf[b: b + c.shape[0], : c.shape[1]] = c
        sizes = torch.cat([t[1] for _, t in self.tables])
        offs = torch.cat([t[2] for _, t in self.tables])
        return torch.cat(self.sym), torch.cat(self.idx), cdf, sizes, offs


class BufferedRansEncoder:
    """rans_interface.cpp:108-200: symbols of every encode_with_indexes call since the last flush go into ONE stream"""

    def __init__(self):
        self._acc = _Accumulated()

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets) -> None:
        self._acc.add(symbols, indexes, cdfs, cdfs_sizes, offsets, _dev())

    def flush(self) -> bytes:
        dev = _dev()
        if not self._acc.sym:
            return (1 << 31).to_bytes(8, "little")
        out = _encode(*self._acc.merged(dev), dev)
        self._acc.clear()
        return out


class RansDecoder:
    """rans_interface.cpp:215-358. `set_stream` + successive `decode_stream` calls continue in one stream; the GPU
    decoder has no resumable state, so call k re-decodes the stream prefix (all earlier calls' symbols + its own) and
    returns the new tail."""

    def __init__(self):
        self._stream = None
        self._acc = _Accumulated()

    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes, offsets):
        dev = _dev()
        return _decode(encoded, indexes, *_table(cdfs, cdfs_sizes, offsets, dev), dev)

    def set_stream(self, encoded) -> None:
        self._stream = bytes(encoded)
        self._acc.clear()

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets):
        if self._stream is None:
            raise ValueError("decode_stream called before set_stream")
        dev = _dev()
        n = len(indexes) if not isinstance(indexes, torch.Tensor) else indexes.numel()
        self._acc.add([], indexes, cdfs, cdfs_sizes, offsets, dev)
        _, idx, cdf, sizes, offs = self._acc.merged(dev)
        return _decode(self._stream, idx, cdf, sizes, offs, dev)[-n:] if n else []
