"""Test-side parser of the CR5B chunk-parallel container (format: cra5_b200/csrc/coder.h)."""
import struct


def parse(b: bytes):
    assert b[:4] == b"CR5B" and b[4] == 1, "not a CR5B v1 container"
    n_channels, L, spc, n_streams = struct.unpack_from("<4I", b, 8)
    assert n_streams == n_channels * spc
    lengths = struct.unpack_from(f"<{n_streams}I", b, 24)
    off = 24 + 4 * n_streams
    streams = []
    for ln in lengths:
        streams.append(b[off:off + ln])
        off += ln
    assert off == len(b)
    return dict(n_channels=n_channels, L=L, spc=spc, streams=streams)
