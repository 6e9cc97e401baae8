"""Python face of the host-side code `bin/match` runs (frog_b200/csrc/keypoint_io.cpp): keypoint
readers, z-filter / pruning and the pairs.bin writer, through libfmio.so.  Lets Python drivers
(FROG.py-style) and the tests reuse the exact C++ the executable uses."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfmio.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m frog_b200.build`")
        L = C.CDLL(LIB_PATH)
        L.fmio_read.restype = C.c_int64
        L.fmio_read.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32,
                                C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]
        L.fmio_filter_prune.restype = C.c_int64
        L.fmio_filter_prune.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_float, C.c_float,
                                        C.c_float, C.c_float, C.c_int]
        L.fmio_write_pairs_bin.restype = C.c_int
        L.fmio_write_pairs_bin.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
        L.fmio_write_csv.restype = C.c_int
        L.fmio_write_csv.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_int]
        L.fmio_fuzz_fmt.restype = C.c_int64
        L.fmio_fuzz_fmt.argtypes = [C.c_uint64, C.c_int64]
        _lib = L
    return _lib


def read_keypoints(path: str, d_cap: int = 64):
    """-> (head[n,6], desc[n,d]) float32, exactly as bin/match loads the file."""
    L = load()
    cap = max(os.path.getsize(path) // 100 + 16, 16)
    while True:
        head = np.zeros((cap, 6), np.float32)
        desc = np.zeros((cap, d_cap), np.float32)
        d = C.c_uint32()
        err = C.create_string_buffer(512)
        n = L.fmio_read(path.encode(), head.ctypes.data, desc.ctypes.data, cap, d_cap, C.byref(d), err, 512)
        if n == -1:
            raise IOError(err.value.decode())
        if n == -2:
            cap *= 4
            d_cap = max(d_cap, d.value)
            continue
        return head[:n].copy(), desc[:n, :d.value].copy()


def filter_prune(head, desc, zT=0.0, zmin=-1e20, zmax=1e20, sp=0.0, np_keep=1000000):
    L = load()
    head = np.ascontiguousarray(head, np.float32).copy()
    desc = np.ascontiguousarray(desc, np.float32).copy()
    n = L.fmio_filter_prune(head.ctypes.data, desc.ctypes.data, head.shape[0], desc.shape[1], zT, zmin, zmax, sp,
                            np_keep)
    return head[:n].copy(), desc[:n].copy()


def write_pairs_bin(path, filenames, rigids, heads, blocks) -> None:
    """blocks: list of (first, second, pairs[m,2] uint32) in the reference's (i, j) order."""
    L = load()
    n_img = len(filenames)
    names = (C.c_char_p * n_img)(*[f.encode() for f in filenames])
    offsets = np.concatenate([[0], np.cumsum([h.shape[0] for h in heads])]).astype(np.int64)
    head = np.ascontiguousarray(np.concatenate(heads) if n_img else np.zeros((0, 6)), np.float32)
    rg = None if rigids is None else np.ascontiguousarray(rigids, np.float64)
    first = np.array([b[0] for b in blocks], np.int32)
    second = np.array([b[1] for b in blocks], np.int32)
    counts = np.array([b[2].shape[0] for b in blocks], np.int64)
    poff = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    pairs = np.ascontiguousarray(np.concatenate([b[2].reshape(-1, 2) for b in blocks]) if blocks else np.zeros((0, 2)),
                                 np.uint32)
    rc = L.fmio_write_pairs_bin(path.encode(), n_img, names, None if rg is None else rg.ctypes.data,
                                offsets.ctypes.data, head.ctypes.data, len(blocks), first.ctypes.data,
                                second.ctypes.data, counts.ctypes.data, poff.ctypes.data, pairs.ctypes.data)
    if rc != 0:
        raise IOError(f"write error : {path}")


def write_csv(path: str, head, desc, gz_level: int = -1) -> None:
    """surf3d's text format (vtk3DSURF.cxx:451-484), plain or gzipped; the GIL is released while the C code runs."""
    L = load()
    head = np.ascontiguousarray(head, np.float32)
    desc = np.ascontiguousarray(desc, np.float32)
    rc = L.fmio_write_csv(path.encode(), head.ctypes.data, desc.ctypes.data, head.shape[0], desc.shape[1], gz_level)
    if rc != 0:
        raise IOError(f"write error {rc}: {path}")
