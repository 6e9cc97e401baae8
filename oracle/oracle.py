"""Python face of the oracle: ctypes loaders + a numpy restatement of ComputeMatches.

TEST INFRASTRUCTURE ONLY (see oracle/match_oracle.c header).  Three independent statements of the
reference hot path live here so they can pin each other:

* `RefLib`   -- oracle/_ref/libmatch_ref.so: the UNMODIFIED reference match.cpp, in-process
* `PortLib`  -- oracle/libmatch_oracle.so: the plain-C restatement (match_oracle.c)
* `compute_matches_numpy` -- float32 numpy, vectorised over columns, k kept sequential

Parity status: pinned (tests/test_oracle.py: numpy == C port == verbatim reference == golden).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "match_ref")
REF_LIB = os.path.join(HERE, "_ref", "libmatch_ref.so")
PORT_LIB = os.path.join(HERE, "libmatch_oracle.so")
FAST_LIB = os.path.join(HERE, "libfast_oracle.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")

_CM_ARGS = [_f32p, _f32p, _f32p, C.c_uint32, _f32p, _f32p, _f32p, C.c_uint32, C.c_uint32,
            C.c_float, C.c_float, C.c_int, _u32p]


def build(ref: bool = True) -> None:
    """Compile the C port (always) and, where /root/reference exists, the verbatim reference."""
    targets = ["port"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE, "-f", os.path.join(HERE, "Makefile")] + targets, check=True)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _CMLib:
    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built -- run `make -C oracle`")
        self.lib = C.CDLL(path)
        self._cm = getattr(self.lib, prefix + "compute_matches")
        self._cm.argtypes = _CM_ARGS
        self._cm.restype = C.c_int64
        self._cma = getattr(self.lib, prefix + "compute_matches_all")
        self._cma.argtypes = [_f32p, _f32p, _f32p, C.c_uint32, _f32p, _f32p, _f32p, C.c_uint32, C.c_uint32,
                              C.c_float, C.c_int, C.c_void_p, C.c_int64]
        self._cma.restype = C.c_int64
        self._norm = getattr(self.lib, prefix + "norm")
        self._norm.argtypes = [_f32p, _f32p, C.c_int]
        self._norm.restype = C.c_float

    def norm(self, a, b) -> np.float32:
        a, b = _f32(a), _f32(b)
        return np.float32(self._norm(a, b, a.shape[0]))

    def compute_matches(self, first, second, threshold: float, ratio: float, sym: bool = False) -> np.ndarray:
        """first/second: (desc[N,D], scale[N], lap[N]).  Returns [M,2] uint32 (first_idx, second_idx)
        -- or (second_idx, first_idx) columns when sym, exactly as the reference pushes them."""
        d1, s1, l1 = map(_f32, first)
        d2, s2, l2 = map(_f32, second)
        out = np.zeros((max(d2.shape[0], 1), 2), np.uint32)
        n = self._cm(d1, s1, l1, d1.shape[0], d2, s2, l2, d2.shape[0], d1.shape[1],
                     threshold, ratio, int(sym), out.reshape(-1))
        return out[:n].copy()


    def compute_matches_all(self, first, second, threshold: float, sym: bool = False) -> np.ndarray:
        """ComputeMatches with matchAll = true (`-all`, match.cpp:295-300): [M,2] uint32 as pushed."""
        d1, s1, l1 = map(_f32, first)
        d2, s2, l2 = map(_f32, second)
        args = (d1, s1, l1, d1.shape[0], d2, s2, l2, d2.shape[0], d1.shape[1], threshold, int(sym))
        n = self._cma(*args, None, 0)
        out = np.zeros((max(n, 1), 2), np.uint32)
        self._cma(*args, out.ctypes.data_as(C.c_void_p), n)
        return out[:n].copy()


class RefLib(_CMLib):
    def __init__(self):
        super().__init__(REF_LIB, "ref_")
        self.lib.ref_distances.argtypes = [_f32p, _f32p, C.c_uint32, _u32p, _u32p, C.c_int64, _f32p]
        self.lib.ref_distances.restype = None

    def distances(self, desc_first, desc_second, first_idx, second_idx) -> np.ndarray:
        a, b = _f32(desc_first), _f32(desc_second)
        fi = np.ascontiguousarray(first_idx, np.uint32)
        si = np.ascontiguousarray(second_idx, np.uint32)
        out = np.zeros(fi.shape[0], np.float32)
        self.lib.ref_distances(a, b, a.shape[1], fi, si, fi.shape[0], out)
        return out


class FastLib:
    """oracle/libfast_oracle.so: ComputeMatches vectorised across columns (fast_oracle.c) -- bit-identical to the port
    and to the verbatim reference (tests/test_oracle.py), ~100x faster: the checker for 20k x 20k / 50k x 50k blocks."""

    def __init__(self):
        if not os.path.exists(FAST_LIB):
            raise FileNotFoundError(f"{FAST_LIB} not built -- run `make -C oracle port`")
        self.lib = C.CDLL(FAST_LIB)
        self.lib.fo_compute_matches.argtypes = _CM_ARGS
        self.lib.fo_compute_matches.restype = C.c_int64

    def compute_matches(self, first, second, threshold: float, ratio: float, sym: bool = False) -> np.ndarray:
        d1, s1, l1 = map(_f32, first)
        d2, s2, l2 = map(_f32, second)
        out = np.zeros((max(d2.shape[0], 1), 2), np.uint32)
        n = self.lib.fo_compute_matches(d1, s1, l1, d1.shape[0], d2, s2, l2, d2.shape[0], d1.shape[1],
                                        threshold, ratio, int(sym), out.reshape(-1))
        if n < 0:
            raise ValueError("fast oracle refused the input (threshold >= 1.8e19 or out of memory)")
        return out[:n].copy()


class PortLib(_CMLib):
    def __init__(self):
        super().__init__(PORT_LIB, "mo_")
        L = self.lib
        L.mo_match_pairs.argtypes = [_f32p, _f32p, _f32p, _i64p, C.c_uint32, _u32p, _u32p, C.c_int64,
                                     C.c_float, C.c_float, C.c_int, _i64p, _u32p, _i64p]
        L.mo_match_pairs.restype = None
        L.mo_read_bin.argtypes = [C.c_char_p, _f32p, C.c_int64]
        L.mo_read_bin.restype = C.c_int64
        L.mo_parse_csv.argtypes = [C.c_char_p, _f32p, C.c_int64, C.c_int]
        L.mo_parse_csv.restype = C.c_int64

    def match_pairs(self, images, pair_first, pair_second, threshold, ratio, sym=False):
        """images: list of (desc, scale, lap).  Returns list of [M,2] uint32 per pair."""
        desc = _f32(np.concatenate([im[0] for im in images]))
        scale = _f32(np.concatenate([im[1] for im in images]))
        lap = _f32(np.concatenate([im[2] for im in images]))
        ns = np.array([im[1].shape[0] for im in images], np.int64)
        offsets = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
        pf = np.ascontiguousarray(pair_first, np.uint32)
        ps = np.ascontiguousarray(pair_second, np.uint32)
        cap = ns[ps] + (ns[pf] if sym else 0)
        out_off = np.concatenate([[0], np.cumsum(cap)]).astype(np.int64)
        out = np.zeros(2 * max(int(out_off[-1]), 1), np.uint32)
        counts = np.zeros(pf.shape[0], np.int64)
        self.lib.mo_match_pairs(desc, scale, lap, offsets, desc.shape[1], pf, ps, pf.shape[0],
                                threshold, ratio, int(sym), out_off, out, counts)
        return [out[2 * out_off[p]: 2 * (out_off[p] + counts[p])].reshape(-1, 2).copy()
                for p in range(pf.shape[0])]

    def match_pairs_all(self, images, pair_first, pair_second, threshold, sym=False):
        """`-all` over image pairs: list of [M,2] uint32 per pair (count pass, then fill pass)."""
        desc = _f32(np.concatenate([im[0] for im in images]))
        scale = _f32(np.concatenate([im[1] for im in images]))
        lap = _f32(np.concatenate([im[2] for im in images]))
        ns = np.array([im[1].shape[0] for im in images], np.int64)
        offsets = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
        pf = np.ascontiguousarray(pair_first, np.uint32)
        ps = np.ascontiguousarray(pair_second, np.uint32)
        counts = np.zeros(pf.shape[0], np.int64)
        fn = self.lib.mo_match_pairs_all
        fn.argtypes = [_f32p, _f32p, _f32p, _i64p, C.c_uint32, _u32p, _u32p, C.c_int64, C.c_float, C.c_int,
                       C.c_void_p, C.c_void_p, _i64p]
        fn.restype = None
        fn(desc, scale, lap, offsets, desc.shape[1], pf, ps, pf.shape[0], threshold, int(sym), None, None, counts)
        out_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        out = np.zeros(2 * max(int(out_off[-1]), 1), np.uint32)
        counts2 = np.zeros_like(counts)
        fn(desc, scale, lap, offsets, desc.shape[1], pf, ps, pf.shape[0], threshold, int(sym),
           out_off.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), counts2)
        assert np.array_equal(counts, counts2)
        return [out[2 * out_off[p]: 2 * out_off[p + 1]].reshape(-1, 2).copy() for p in range(pf.shape[0])]

    def read_bin(self, path: str) -> np.ndarray:
        cap = os.path.getsize(path) // 216 + 2
        rec = np.zeros((cap, 54), np.float32)
        n = self.lib.mo_read_bin(path.encode(), rec.reshape(-1), cap)
        if n < 0:
            raise OSError(path)
        return rec[:n].copy()

    def parse_csv(self, text: bytes, d: int = 48) -> np.ndarray:
        cap = text.count(b"\n") + 2
        rec = np.zeros((cap, 6 + d), np.float32)
        n = self.lib.mo_parse_csv(text + b"\0", rec.reshape(-1), cap, d)
        if n < 0:
            raise ValueError("ragged CSV")
        return rec[:n].copy()


# ----------------------------------------------------------------------------------------------
# numpy restatement


def norm_matrix_numpy(desc_second: np.ndarray, desc_first: np.ndarray) -> np.ndarray:
    """[n_second, n_first] float32 squared distances in the reference's operation order
    (match.cpp:246-248): r = fl(r + fl(fl(a-b) * fl(a-b))), k ascending."""
    a = _f32(desc_second)
    b = _f32(desc_first)
    r = np.zeros((a.shape[0], b.shape[0]), np.float32)
    for k in range(a.shape[1]):
        diff = a[:, k][:, None] - b[:, k][None, :]
        r += diff * diff
    return r


def gate_matrix_numpy(scale_second, lap_second, scale_first, lap_first) -> np.ndarray:
    """True where the pair survives the Laplacian (match.cpp:270) and scale (:273-275) gates."""
    s1, l1 = _f32(scale_second)[:, None], _f32(lap_second)[:, None]
    s2, l2 = _f32(scale_first)[None, :], _f32(lap_first)[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        bad = (l1 != l2) | ((s1 / s2).astype(np.float64) > 1.3) | ((s2 / s1).astype(np.float64) > 1.3)
    return ~bad


def compute_matches_numpy(first, second, threshold: float, ratio: float, sym: bool = False) -> np.ndarray:
    d1_, s1_, l1_ = first
    d2_, s2_, l2_ = second
    dist = norm_matrix_numpy(d2_, d1_)
    ok = gate_matrix_numpy(s2_, l2_, s1_, l1_)
    big = np.float32(np.finfo(np.float32).max)
    dist = np.where(ok, dist, np.float32(np.inf))
    n2 = dist.shape[0]
    if dist.shape[1] == 0:
        return np.zeros((0, 2), np.uint32)
    match = np.argmin(dist, axis=1)  # first minimum == strict '<' scan
    d1 = dist[np.arange(n2), match]
    tmp = dist.copy()
    tmp[np.arange(n2), match] = np.inf
    d2 = tmp.min(axis=1) if dist.shape[1] > 1 else np.full(n2, np.inf, np.float32)
    d1 = np.where(np.isinf(d1), big, d1).astype(np.float32)
    d2 = np.where(np.isinf(d2), big, d2).astype(np.float32)
    thr, rat = np.float32(threshold), np.float32(ratio)
    with np.errstate(divide="ignore", invalid="ignore"):
        accept = ((np.sqrt(d1 / d2) < rat) | (d2 == big)) & (np.sqrt(d1) < thr)
    rows = np.nonzero(accept)[0]
    cols = match[rows]
    pairs = np.stack([rows, cols] if sym else [cols, rows], axis=1)
    return pairs.astype(np.uint32)


def run_ref_binary(args, cwd=None, threads=None) -> subprocess.CompletedProcess:
    """Run the verbatim reference executable (oracle/_ref/match_ref)."""
    if not os.path.exists(REF_BIN):
        raise FileNotFoundError(f"{REF_BIN} not built -- run `make -C oracle ref`")
    cmd = [REF_BIN] + [str(a) for a in args]
    if threads:
        cmd += ["-nt", str(threads)]
    return subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, check=True)


# ----------------------------------------------------------------------------------------------
# consumer hand-off: the links bin/frog builds when it reads pairs.bin


def read_pairs_links(blocks, n_points):
    """ImageGroup::readPairs, registration/imageGroup.cxx:1386-1411, restated literally: blocks = [(image1, image2,
    [m,2] uint32)] in FILE order; for every entry (p1, p2) push {image2, p2} onto image1's point p1 and {image1, p1}
    onto image2's point p2.  Returns {image: list (per point) of lists of (image, point)}."""
    links = {img: [[] for _ in range(n)] for img, n in n_points.items()}
    for i, j, m in blocks:
        for p1, p2 in m.tolist():
            links[i][p1].append((j, p2))
            links[j][p2].append((i, p1))
    return links


def read_pairs_links_csr(blocks, n_points):
    """The same as CSR arrays (offsets per image, [total,2] links), through a STABLE sort of the push_back sequence by
    point -- fast enough for BASELINE-size groups; pinned to read_pairs_links in tests/test_oracle.py."""
    imgs = sorted(n_points)
    base, run = {}, 0
    for img in imgs:
        base[img] = run
        run += n_points[img]
    gids, vals = [], []
    for i, j, m in blocks:
        if m.shape[0] == 0:
            continue
        m = m.astype(np.int64)
        g = np.empty(2 * m.shape[0], np.int64)
        v = np.empty((2 * m.shape[0], 2), np.uint32)
        g[0::2] = base[i] + m[:, 0]
        g[1::2] = base[j] + m[:, 1]
        v[0::2, 0], v[0::2, 1] = j, m[:, 1]
        v[1::2, 0], v[1::2, 1] = i, m[:, 0]
        gids.append(g)
        vals.append(v)
    if gids:
        g, v = np.concatenate(gids), np.concatenate(vals)
        order = np.argsort(g, kind="stable")
        v = v[order]
        deg = np.bincount(g, minlength=run)
    else:
        v, deg = np.zeros((0, 2), np.uint32), np.zeros(run, np.int64)
    off = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint64)
    return {img: off[base[img]: base[img] + n_points[img] + 1] for img in imgs}, v
