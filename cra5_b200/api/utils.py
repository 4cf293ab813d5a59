"""Framing helpers of the CRA5 `.bin` container (reference: cra5/api/utils.py:10-33): big-endian uint32 fields and raw
byte strings."""
import struct
from pathlib import Path


def filesize(filepath: str) -> int:
    p = Path(filepath)
    if not p.is_file():
        raise ValueError(f'Invalid file "{filepath}".')
    return p.stat().st_size


def write_uints(fd, values) -> int:
    fd.write(struct.pack(f">{len(values)}I", *values))
    return 4 * len(values)


def write_bytes(fd, values) -> int:
    if len(values) == 0:
        return 0
    fd.write(bytes(values))
    return len(values)


def read_uints(fd, n):
    raw = fd.read(4 * n)
    if len(raw) != 4 * n:
        raise ValueError("truncated .bin container")
    return struct.unpack(f">{n}I", raw)


def read_bytes(fd, n) -> bytes:
    raw = fd.read(n)
    if len(raw) != n:
        raise ValueError("truncated .bin container")
    return raw


def write_bin(path, strings, z_shape) -> int:
    """`>I z_h, >I z_w, >I n_strings, [>I len, bytes] * n` -- first batch item only (cra5_api.py:108-116)"""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with path.open("wb") as f:
        n = write_uints(f, (int(z_shape[0]), int(z_shape[1]), len(strings)))
        for s in strings:
            n += write_uints(f, (len(s[0]),))
            n += write_bytes(f, s[0])
    return n


def read_bin(path):
    """-> (strings as [[bytes], [bytes]], (z_h, z_w))  (cra5_api.py:132-140, 161-169)"""
    with Path(path).open("rb") as f:
        shape = read_uints(f, 2)
        n_strings = read_uints(f, 1)[0]
        strings = []
        for _ in range(n_strings):
            strings.append([read_bytes(f, read_uints(f, 1)[0])])
    return strings, shape
