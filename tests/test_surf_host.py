"""CPU: host side of the SURF3D producer (libfrogsurf.so's host logic, bin/surf3d's I/O) and its C ABI."""
import ctypes as C
import gzip
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, have_gpu
from frog_b200 import build, surf

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_surf_golden as mg  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "surf")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))


def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z_0-9]+)\s*\(", text)))


def test_exports_match_header(built):
    lib = C.CDLL(build.SURF_LIB)
    declared = _declared("frogsurf.h", "fs_")
    assert sorted(surf.PUBLIC_SYMBOLS) == declared
    for name in declared + _declared("frogsurf_debug.h", "fs_debug_"):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    text = open(os.path.join(ROOT, "include", "frogsurf.h")).read().lower()
    assert "torch" not in text and "at::" not in text


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built, tmp_path):
    with pytest.raises(surf.FrogSurfError, match="no CUDA device"):
        surf.Producer(0)
    vol = np.zeros((40, 40, 40), np.int16)
    mha = str(tmp_path / "v.mha")
    surf.write_metaimage(mha, vol)
    r = subprocess.run([build.SURF_BIN, mha, "-o", str(tmp_path / "pts")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    assert not os.path.exists(tmp_path / "pts.csv.gz")


def test_cli_usage_and_rejections(built, tmp_path):
    r = subprocess.run([build.SURF_BIN], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("Usage : surf3d file [options]")  # surf3d.cxx:19-42
    mha = str(tmp_path / "v.mha")
    surf.write_metaimage(mha, np.zeros((8, 8, 8), np.int16))
    for flags in (["-s", "0.75"], ["-type", "2"], ["-m", "mask.mhd"], ["-pad", "3"]):
        r = subprocess.run([build.SURF_BIN, mha] + flags, capture_output=True, text=True)
        assert r.returncode == 7 and "VTK" in r.stderr  # stated limits, not silent approximations
    r = subprocess.run([build.SURF_BIN, str(tmp_path / "missing.mha")], capture_output=True, text=True)
    assert r.returncode == 5 and "Cannot load file" in r.stderr  # vtkRobustImageReader.h:35-38


def test_expf_restatement_equals_libm(built):
    """surf.cxx:227 calls expf; the kernel's restatement must return libm's bits.  Dense sample of the descriptor's
    argument range, a coarse sweep of the whole domain, and the edges."""
    libm = C.CDLL("libm.so.6")
    libm.expf.restype = C.c_float
    libm.expf.argtypes = [C.c_float]
    rng = np.random.default_rng(3)
    xs = np.concatenate([
        -rng.random(60000, dtype=np.float32) * np.float32(6.0),
        np.arange(0x80000000, 0xc2d00000, 104729, dtype=np.uint32).view(np.float32),  # negative floats down to -104
        np.arange(0, 0x42b20000, 150001, dtype=np.uint32).view(np.float32),           # positive floats up to 89
        np.array([0.0, -0.0, -87.9, -88.1, -103.9, -103.98, -104.5, -1e30, 88.7, 88.8, float.fromhex('-0x1.f8cbb2p+5'), float.fromhex('0x1.04845ep+5')], np.float32),
    ])
    got = surf.debug_expf(xs)
    want = np.array([libm.expf(float(x)) for x in xs], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_layer_geometry_matches_reference():
    for name, m in MANIFEST.items():
        nz, ny, nx = m["shape"]
        got = surf.debug_layers(nx, ny, nz)
        want = [[l["width"], l["height"], l["depth"], l["step"], l["filter"], l["limit"]] for l in m["layers"]]
        assert got.tolist() == want
    assert len(surf.debug_layers(400, 400, 400)) == 10 and len(surf.debug_layers(60, 300, 300)) == 4


def test_select_reproduces_reference_order():
    """vtk3DSURF.cxx:209-226 on the detector's push_back order gives the reference's final order (same std::sort /
    std::partial_sort calls), including ties."""
    for name in ("small", "mid"):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        order = surf.debug_select(g["det_xyzsr"][:, 4], 20000)
        assert np.array_equal(g["det_xyzsr"][order].view(np.uint32), g["xyzsr"].view(np.uint32))
    g = np.load(os.path.join(GOLD, "small.npz"))
    order = surf.debug_select(g["det_xyzsr"][:, 4], 40)  # partial_sort branch: the raw-descriptor case kept 40
    assert np.array_equal(g["det_xyzsr"][order][:, :4].view(np.uint32), g["raw_xyzsr"][:, :4].view(np.uint32))
    r = np.array([3, 1, 3, 2, 3, 1, 2, 2, 3, 1] * 7, np.float32)  # heavy ties: a permutation, sorted, deterministic
    o1, o2 = surf.debug_select(r, 1000), surf.debug_select(r, 1000)
    assert np.array_equal(o1, o2) and sorted(o1.tolist()) == list(range(70)) and np.all(np.diff(r[o1]) <= 0)
    assert len(surf.debug_select(r, 0)) == 70 and np.array_equal(surf.debug_select(r, -1), np.arange(70))


def _golden_points():
    g = np.load(os.path.join(GOLD, "small.npz"))
    pts = np.zeros(len(g["xyzsr"]), surf.POINT_DTYPE)
    for i, k in enumerate(("x", "y", "z", "scale", "response")):
        pts[k] = g["xyzsr"][:, i]
    pts["laplacian"] = g["lap"]
    return pts, g["desc"]


def test_writers_match_reference_files(built, tmp_path):
    """csv, csv.gz (default and `-gz 9 -precision 4`) and bin files byte-identical to the ones the reference's own
    writers produced for the same keypoints (vtk3DSURF.cxx:405-525)."""
    c = mg.CASES["small"]
    pts, desc = _golden_points()
    for fmt in ("csv", "csv.gz", "bin"):
        out = str(tmp_path / ("p." + fmt))
        surf.write_points(out, fmt, pts, desc, c["spacing"], c["origin"])
        assert open(out, "rb").read() == open(os.path.join(GOLD, "small_points." + fmt), "rb").read(), fmt
    out = str(tmp_path / "p9.csv.gz")
    surf.write_points(out, "csv.gz", pts, desc, c["spacing"], c["origin"], gz_opts="9", precision=4)
    assert open(out, "rb").read() == open(os.path.join(GOLD, "small_points_l9p4.csv.gz"), "rb").read()
    assert gzip.open(out).read().count(b"\n") == len(pts)


def test_written_keypoints_feed_the_matcher_reader(built, tmp_path):
    """The producer's files are the matcher's inputs (match.cpp:51-92, 179-208): bin/match's readers parse them."""
    from frog_b200 import hostio
    c = mg.CASES["small"]
    pts, desc = _golden_points()
    heads = {}
    for fmt in ("csv.gz", "bin"):
        out = str(tmp_path / ("p." + fmt))
        surf.write_points(out, fmt, pts, desc, c["spacing"], c["origin"])
        head, d = hostio.read_keypoints(out)
        n = len(pts) + (1 if fmt == "bin" else 0)  # readBinary's phantom record
        assert d.shape == (n, 48) and head.shape[0] == n
        heads[fmt] = (head, d)
    assert np.array_equal(heads["bin"][1][:len(pts)].view(np.uint32), desc.view(np.uint32))
    assert np.abs(heads["csv.gz"][1] - desc).max() <= 5.1e-7  # "%f": six decimals


def test_metaimage_reader(built, tmp_path):
    rng = np.random.default_rng(0)
    for dt in (np.uint8, np.int16, np.uint16, np.int32, np.float32):
        vol = (rng.random((5, 6, 7)) * 200).astype(dt)
        p = str(tmp_path / f"v_{np.dtype(dt).name}.mha")
        surf.write_metaimage(p, vol, spacing=(0.5, 0.75, 2.0), origin=(1.0, -2.0, 3.5))
        got, sp, org = surf.read_metaimage(p)
        assert np.array_equal(got, vol) and sp == (0.5, 0.75, 2.0) and org == (1.0, -2.0, 3.5)
    # detached data file + a negative direction cosine: flipped like vtkRobustImageReader.h:97-113
    vol = (rng.random((4, 5, 6)) * 1000).astype(np.int16)
    vol.tofile(str(tmp_path / "d.raw"))
    open(tmp_path / "d.mhd", "w").write(
        "ObjectType = Image\nNDims = 3\nDimSize = 6 5 4\nElementSpacing = 1 2 3\nOffset = 10 20 30\n"
        "TransformMatrix = 1 0 0 0 -1 0 0 0 1\nElementType = MET_SHORT\nElementDataFile = d.raw\n")
    got, sp, org = surf.read_metaimage(str(tmp_path / "d.mhd"))
    assert np.array_equal(got, vol[:, ::-1, :]) and org == (10.0, 20.0 - 2.0 * 4, 30.0)
    with pytest.raises(surf.FrogSurfError):
        surf.read_metaimage(str(tmp_path / "nope.mhd"))


def test_select_ties_follow_libstdcxx_on_ipoint_like_objects(built, tmp_path):
    """The order of equal responses after std::partial_sort / std::sort is decided by libstdc++'s algorithm, not by the
    payload: the library's 8-byte keys must come out in the order an Ipoint-like object (with a std::vector member,
    moved around like the reference's, ipoint.h:26-66; comparator as vtk3DSURF.cxx:32) does -- ties included, both
    branches of vtk3DSURF.cxx:209-226."""
    src = tmp_path / "ipsort.cpp"
    src.write_text(r'''
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
struct Ip { float x, y, z, scale, response; int laplacian; std::vector<float> descriptor; int id; };
bool compareResponses(Ip& i, Ip& j) { return (i.response > j.response); }
int main(int argc, char** argv) {
  int keep = atoi(argv[1]);
  std::vector<Ip> points;
  float r;
  int id = 0;
  while (scanf("%f", &r) == 1) { Ip p; p.response = r; p.id = id++; p.descriptor.assign(3, r); points.push_back(p); }
  if (keep > 0) {
    if ((int)points.size() > keep) {
      std::partial_sort(points.begin(), points.begin() + keep, points.end(), compareResponses);
      points.resize(keep);
    } else {
      std::sort(points.begin(), points.end(), compareResponses);
    }
  }
  for (auto& p : points) printf("%d\n", p.id);
}
''')
    exe = str(tmp_path / "ipsort")
    subprocess.run(["g++", "-O3", "-std=c++17", str(src), "-o", exe], check=True)
    rng = np.random.default_rng(21)
    for n, keep, levels in ((500, 120, 7), (500, 800, 7), (3000, 2999, 40), (64, 16, 2), (2000, 500, 100000)):
        resp = rng.integers(0, levels, n).astype(np.float32)
        out = subprocess.run([exe, str(keep)], input="\n".join(repr(float(v)) for v in resp), capture_output=True, text=True, check=True)
        want = np.array([int(t) for t in out.stdout.split()], np.uint32)
        got = surf.debug_select(resp, keep)
        assert np.array_equal(got, want), (n, keep, levels)


def test_point_file_reader_matches_reference(built, tmp_path):
    """surf3d -p: vtk3DSURF::ReadIPoints (vtk3DSURF.cxx:34-77).  World-unit lines -> voxel-unit keypoints identical to what
    the reference converted from the same file (empty line skipped, CRLF tolerated); its quirks on short lines."""
    c = mg.CASES["small"]
    g = np.load(os.path.join(GOLD, "pfile.npz"))
    xyzs, outside = surf.read_points_file(os.path.join(GOLD, "pfile_points.csv"), c["spacing"], c["origin"], c["shape"])
    assert outside == 0 and xyzs.shape == (50, 4)
    assert np.array_equal(xyzs.view(np.uint32), g["xyzsr"][:, :4].view(np.uint32))
    # a failed getline leaves the cell string as it was: missing cells repeat the last one; a trailing comma erases it
    p = str(tmp_path / "q.csv")
    open(p, "w").write("1.5,2.5\n\n4,5,6,7,8\n  3.25 ,1e1,-2,0.5junk\n")
    xyzs, outside = surf.read_points_file(p, (1.0, 2.0, 4.0), (0.5, 0.5, 0.5), (100, 100, 100))
    ss = (1.0 * 2.0 * 4.0) ** (1.0 / 3.0)
    want = np.array([[1.0, 1.0, 0.5, np.float32(2.5) / ss], [3.5, 2.25, 1.375, np.float32(7) / ss],
                     [2.75, 4.75, -0.625, np.float32(0.5) / ss]], np.float32)
    assert np.array_equal(xyzs, want) and outside == 1  # z = -2 lies outside the image; the point is kept
    open(p, "w").write("1,2,3,\n")
    with pytest.raises(surf.FrogSurfError):
        surf.read_points_file(p, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), (10, 10, 10))  # std::stof("") throws in the reference
    open(p, "w").write("1,x,3,4\n")
    with pytest.raises(surf.FrogSurfError):
        surf.read_points_file(p, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), (10, 10, 10))


def test_extrema_ordering_is_a_stable_sort(built):
    """fs_detect restores the reference's push_back order by sorting loop-position keys: the radix sort must equal a
    stable sort for every key pattern (constant bytes skipped, duplicates, high bits, empty and single inputs)."""
    rng = np.random.default_rng(4)
    cases = [np.zeros(0, np.uint64), np.array([7], np.uint64), rng.integers(0, 5, 1000).astype(np.uint64),
             (rng.integers(0, 8, 30000).astype(np.uint64) << np.uint64(48)) | rng.integers(0, 1 << 23, 30000).astype(np.uint64),
             rng.integers(0, 1 << 62, 5000, dtype=np.int64).astype(np.uint64),
             np.full(300, 0xABCDEF0123, np.uint64), np.arange(4000, 0, -1).astype(np.uint64) * np.uint64(257)]
    for keys in cases:
        got = surf.debug_sort_keys(keys)
        assert np.array_equal(got, np.argsort(keys, kind="stable").astype(np.uint32))


def test_fixed_decimal_formatter_is_printf(built):
    """The writers' "%f" / "%.<n>f" replacement (exact 128-bit integer arithmetic, ties to even) against snprintf:
    descriptor-like floats, coordinates, responses, tiny and negative-zero values, exact binary ties, every precision
    the fast path takes, and the ranges it hands back to snprintf."""
    import ctypes as C
    L = C.CDLL(build.FSIO)
    L.fsio_debug_format_check.restype = C.c_long
    L.fsio_debug_format_check.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p]
    rng = np.random.default_rng(8)
    sets = [
        rng.uniform(-1, 1, 400000).astype(np.float32).astype(np.float64),
        rng.uniform(-2000, 2000, 200000),
        (rng.uniform(0, 1, 100000) * 10.0 ** rng.integers(-12, 13, 100000)),
        rng.uniform(0, 3e6, 100000).astype(np.float32).astype(np.float64),
        np.array([0.0, -0.0, 5e-7, -5e-7, 4.9999999e-7, 1e-300, -1e-300, 5e-324, 0.9999995, 0.99999949999, 999999.9999995,
                  8.9e12, 9.1e12, 1e20, -1e20, np.inf, -np.inf, 0.5, 1.5, 2.5, -0.5, 0.125, 0.375, 0.0625, 2.0 ** -20, 1.0, -1.0]),
        (np.arange(-2000, 2000) + 0.5) / 8.0,       # exact binary ties at one, two and three decimals
        np.arange(0, 4096) / 4096.0,
    ]
    for vals in sets:
        v = np.ascontiguousarray(vals, np.float64)
        for decimals in (6, 0, 1, 2, 3, 4, 9, 12):
            bad = C.c_long(-1)
            n_bad = L.fsio_debug_format_check(v.ctypes.data, v.size, decimals, C.byref(bad))
            assert n_bad == 0, f"decimals {decimals}: {n_bad} differ, first {v[bad.value]!r}"
