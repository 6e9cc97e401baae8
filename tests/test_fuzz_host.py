"""CPU: mutation fuzz of the two parsers that read untrusted keypoint files -- the gzip/DEFLATE decoder
(fast_inflate.cpp) and the CSV text parser (keypoint_io.cpp) -- compiled with AddressSanitizer and
UndefinedBehaviorSanitizer.  Whatever the input: no out-of-bounds access, no UB; what the inflate decoder accepts,
zlib accepts too and agrees with byte for byte."""
import os
import subprocess

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "frog_b200", "csrc")
FUZZ = os.path.join(ROOT, "tests", "fuzz")


def _build(tmp_path, name, sources):
    exe = str(tmp_path / name)
    cmd = ["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17",
           "-I", CSRC, "-I", os.path.join(ROOT, "include")] + sources + ["-lz", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build unavailable: " + r.stderr[-300:])
    return exe


def test_fuzz_fast_inflate(tmp_path):
    exe = _build(tmp_path, "fuzz_inflate", [os.path.join(FUZZ, "fuzz_inflate.cpp"), os.path.join(CSRC, "fast_inflate.cpp")])
    r = subprocess.run([exe, "16"], capture_output=True, text=True, timeout=600)  # 16 base streams x 1500 mutations
    assert r.returncode == 0 and "fuzz ok" in r.stdout, (r.stdout + r.stderr)[-2000:]


def test_fuzz_csv_parser(tmp_path):
    exe = _build(tmp_path, "fuzz_csv", [os.path.join(FUZZ, "fuzz_csv.cpp"), os.path.join(CSRC, "keypoint_io.cpp"),
                                        os.path.join(CSRC, "fast_inflate.cpp")])
    r = subprocess.run([exe, "10"], capture_output=True, text=True, timeout=600)  # 10 base texts x 2000 mutations
    assert r.returncode == 0 and "csv fuzz ok" in r.stdout, (r.stdout + r.stderr)[-2000:]
