"""Parser / differ for the `pairs.bin` file `bin/match` writes and `bin/frog` reads.

Layout (little-endian), written at match/match.cpp:684-742 and consumed by
registration/imageGroup.cxx:1353-1417:

    u16 nImages
    nImages x { u16 nameLen; char name[nameLen]; f64 rigid[3]; u32 nPoints;
                nPoints x f32[6] (x, y, z, scale, laplacianSign, response) }
    per computed image pair (including empty ones):
                { u16 i; u16 j; u32 size; size x (u32 first, u32 second) }
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


@dataclass
class PairsFile:
    names: list = field(default_factory=list)
    rigids: list = field(default_factory=list)
    points: list = field(default_factory=list)  # per image [nPoints,6] float32
    blocks: list = field(default_factory=list)  # (i, j, [size,2] uint32) in file order
    header_bytes: bytes = b""

    def block_map(self) -> dict:
        return {(i, j): m for i, j, m in self.blocks}

    def n_matches(self) -> int:
        return int(sum(m.shape[0] for _, _, m in self.blocks))


def parse(path_or_bytes) -> PairsFile:
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        buf = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            buf = f.read()
    pf = PairsFile()
    off = 0
    (n_img,) = struct.unpack_from("<H", buf, off)
    off += 2
    for _ in range(n_img):
        (ln,) = struct.unpack_from("<H", buf, off)
        off += 2
        pf.names.append(buf[off:off + ln].decode("latin-1"))
        off += ln
        pf.rigids.append(struct.unpack_from("<3d", buf, off))
        off += 24
        (npts,) = struct.unpack_from("<I", buf, off)
        off += 4
        pf.points.append(np.frombuffer(buf, dtype="<f4", count=npts * 6, offset=off).reshape(npts, 6).copy())
        off += npts * 24
    pf.header_bytes = buf[:off]
    while off < len(buf):
        i, j, size = struct.unpack_from("<HHI", buf, off)
        off += 8
        m = np.frombuffer(buf, dtype="<u4", count=size * 2, offset=off).reshape(size, 2).copy()
        off += size * 8
        pf.blocks.append((i, j, m))
    if off != len(buf):
        raise ValueError("trailing bytes in pairs.bin")
    return pf


def diff(ref: PairsFile, new: PairsFile) -> dict:
    """Structured comparison used by the parity tests and the CLI end-to-end check."""
    out = {
        "header_equal": ref.header_bytes == new.header_bytes,
        "block_keys_equal": [(i, j) for i, j, _ in ref.blocks] == [(i, j) for i, j, _ in new.blocks],
        "blocks_sequence_equal": True,
        "blocks_set_equal": True,
        "only_ref": [],
        "only_new": [],
    }
    rmap, nmap = ref.block_map(), new.block_map()
    for key in sorted(set(rmap) | set(nmap)):
        a = rmap.get(key, np.zeros((0, 2), np.uint32))
        b = nmap.get(key, np.zeros((0, 2), np.uint32))
        if a.shape != b.shape or not np.array_equal(a, b):
            out["blocks_sequence_equal"] = False
            sa = set(map(tuple, a.tolist()))
            sb = set(map(tuple, b.tolist()))
            if sa != sb:
                out["blocks_set_equal"] = False
                out["only_ref"] += [(key, p) for p in sorted(sa - sb)]
                out["only_new"] += [(key, p) for p in sorted(sb - sa)]
    return out
