"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; without a GPU
the product fails loudly instead of falling back."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT, have_gpu
from frog_b200 import build, capi


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fm_[a-z_0-9]+)\s*\(", text)))


def test_exports_match_headers(built):
    lib = ctypes.CDLL(build.LIB)
    declared = _declared("frogmatch.h")
    assert sorted(capi.PUBLIC_SYMBOLS) == declared
    for name in declared + _declared("frogmatch_debug.h"):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "frogmatch.h")).read()
    assert "torch" not in text.lower() and "at::" not in text


def test_library_is_sm100a_tcgen05(built):
    sass = subprocess.run(["cuobjdump", "-sass", build.LIB], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass or "SM100a" in sass or "sm_100" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):  # tcgen05.mma, tcgen05.ld, TMA bulk copy
        assert mnemonic in sass


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built, golden_dir, tmp_path):
    with pytest.raises(capi.FrogMatchError):
        capi.Matcher(0)
    r = subprocess.run([build.BIN, os.path.join(golden_dir, "list_bin.txt"), "-o", str(tmp_path / "o.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    assert not os.path.exists(tmp_path / "o.bin")


def test_cli_usage(built):
    r = subprocess.run([build.BIN], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("Usage : match pointFiles.txt")  # match.cpp:347-350
    r = subprocess.run([build.BIN, "/nonexistent/list.txt"], capture_output=True, text=True)
    assert r.returncode == 1 and "Bad argument" in r.stderr  # match.cpp:494-498
