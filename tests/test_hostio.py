"""CPU: the host-side C++ of bin/match (readers, pruning, writer) against the oracle port and the
reference's documented quirks."""
import gzip
import os

import numpy as np

from frog_b200 import hostio, pairsbin, synth
from oracle import oracle as O


def test_bin_reader_phantom_record(built, tmp_path):
    """match.cpp:184-205: `while(!feof)` appends one record whose six header fields all equal the
    last response and whose descriptor is zero."""
    kp = synth.make("iid", 37, 5)
    p = str(tmp_path / "a.bin")
    synth.write_bin(kp, p)
    head, desc = hostio.read_keypoints(p)
    assert head.shape == (38, 6) and desc.shape == (38, 48)
    assert np.array_equal(head[:37], kp.records()[:, :6]) and np.array_equal(desc[:37], kp.desc)
    assert np.all(head[37] == kp.response[-1]) and not desc[37].any()
    rec = O.PortLib().read_bin(p)
    assert np.array_equal(rec[:, :6], head) and np.array_equal(rec[:, 6:], desc)


def test_bin_reader_truncated_tail(built, tmp_path):
    kp = synth.make("iid", 5, 6)
    raw = kp.records().astype("<f4").tobytes()
    p = str(tmp_path / "t.bin")
    open(p, "wb").write(raw[:-100])  # cut inside the last descriptor
    head, desc = hostio.read_keypoints(p)
    rec = O.PortLib().read_bin(p)
    assert np.array_equal(rec[:, :6], head) and np.array_equal(rec[:, 6:], desc)
    assert head.shape[0] == 5  # the short read sets EOF: no phantom after a truncated record


def test_text_readers_match_port_and_bin(built, tmp_path):
    kp = synth.make("bank", 64, 2)
    pc, pg = str(tmp_path / "a.csv"), str(tmp_path / "a.csv.gz")
    synth.write_csv(kp, pc)
    synth.write_csv_gz(kp, pg)
    hc, dc = hostio.read_keypoints(pc)
    hg, dg = hostio.read_keypoints(pg)
    assert np.array_equal(hc, hg) and np.array_equal(dc, dg)
    assert np.array_equal(hc, kp.records()[:, :6]) and np.array_equal(dc, kp.desc)
    rec = O.PortLib().parse_csv(gzip.open(pg).read())
    assert np.array_equal(rec[:, :6], hc) and np.array_equal(rec[:, 6:], dc)


def test_csv_quirks(built, golden_dir):
    """CRLF, exponent cells, leading blanks, trailing commas, short and empty lines."""
    head, desc = hostio.read_keypoints(os.path.join(golden_dir, "quirks.csv"))
    rec = O.PortLib().parse_csv(open(os.path.join(golden_dir, "quirks.csv"), "rb").read())
    assert head.shape[0] == 40
    assert np.array_equal(rec[:, :6], head) and np.array_equal(rec[:, 6:], desc)


def test_prune_and_zfilter(built):
    kp = synth.make("iid", 500, 9)
    rec = kp.records()
    head, desc = hostio.filter_prune(rec[:, :6], rec[:, 6:], zT=10.0, zmin=200.0, zmax=900.0, sp=1000.0, np_keep=50)
    z = rec[:, 2] + np.float32(10.0)
    keep = ~((z < 200.0) | (z > 900.0)) & ~(rec[:, 5] < 1000.0)
    assert head.shape[0] == 50
    # survivors are the 50 largest responses of the kept set (order is libstdc++'s, pinned by golden bin_prune)
    assert set(head[:, 5].tolist()) == set(np.sort(rec[keep, 5])[::-1][:50].tolist())


def test_writer_layout(built, tmp_path):
    heads = [np.arange(12, dtype=np.float32).reshape(2, 6), np.zeros((0, 6), np.float32)]
    blocks = [(0, 1, np.array([[1, 0]], np.uint32)), (65536 + 3, 1, np.zeros((0, 2), np.uint32))]
    p = str(tmp_path / "w.bin")
    hostio.write_pairs_bin(p, ["/x/y/a.bin", "b\\c.csv"], [[1.0, 2.0, 3.0], [0, 0, 0]], heads, blocks)
    pf = pairsbin.parse(p)
    assert pf.names == ["a.bin", "c.csv"] and pf.rigids[0] == (1.0, 2.0, 3.0)
    assert [b[:2] for b in pf.blocks] == [(0, 1), (3, 1)]  # ids truncated to u16 (match.cpp:735-736)
    assert pf.blocks[1][2].shape == (0, 2)  # empty blocks are still written (match.cpp:732-738)


def test_fast_decimal_parser_is_strtof(built):
    """keypoint_io.cpp fast_strtof: the libc-free text->float path must agree with strtof (= std::stof,
    match.cpp:69,152) bit for bit, value and consumed prefix, on surf3d-style "%f" cells, long digit
    strings, exponent forms (also broken ones), exact float midpoints and odd shapes."""
    import ctypes as C

    from frog_b200 import build
    L = C.CDLL(build.FMIO)
    L.fmio_fuzz_floats.argtypes = [C.c_uint64, C.c_int64, C.POINTER(C.c_int64), C.c_char_p, C.c_size_t]
    L.fmio_fuzz_floats.restype = C.c_int64
    n_fast, bad = C.c_int64(), C.create_string_buffer(128)
    for seed in (1, 2, 3):
        n_bad = L.fmio_fuzz_floats(seed, 400_000, C.byref(n_fast), bad, 128)
        assert n_bad == 0, f"first mismatching cell: {bad.value!r}"
        assert n_fast.value > 200_000  # the fast path really carries the common formats


def test_csv_long_descriptors_through_cli_readers(tmp_path):
    """surf3d -type 2 writes 8 r^3 = 1000 values per keypoint at the default radius (vtkOpenSURF3D/surf3d.cxx:36-39);
    the CSV readers take whatever length the rows have (match.cpp:160-168)."""
    import gzip
    rng = np.random.default_rng(3)
    n, d = 17, 1000
    rec = np.concatenate([rng.uniform(0, 100, (n, 6)), rng.standard_normal((n, d))], axis=1).astype(np.float32)
    text = "\n".join(",".join("%f" % v for v in row) for row in rec) + "\n"
    p_csv, p_gz = tmp_path / "long.csv", tmp_path / "long.csv.gz"
    p_csv.write_text(text)
    with gzip.open(p_gz, "wt") as f:
        f.write(text)
    want = np.array([[np.float32(float("%f" % v)) for v in row] for row in rec], np.float32)
    for p in (p_csv, p_gz):
        head, desc = hostio.read_keypoints(str(p))
        assert head.shape == (n, 6) and desc.shape == (n, d)
        assert np.array_equal(head, want[:, :6]) and np.array_equal(desc, want[:, 6:])
