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


def test_cli_gpu_plan(built, golden_dir, tmp_path):
    """The executable chooses its GPU count BEFORE the first CUDA call (cuInit costs seconds per visible device on
    a multi-GPU box): -gpus G, else the G that minimises start-up(G) + pairs / (G x rate) with the pairs estimated
    from the keypoint file sizes (each further GPU costs a one-shot process ~0.65 s of CUDA bring-up), never more than
    there are image pairs.  `-plan 1` prints the plan without touching CUDA."""
    lst = os.path.join(golden_dir, "list_bin.txt")  # 4 images x 221 records: 6 image pairs

    def plan(*extra):
        r = subprocess.run([build.BIN, lst, "-plan", "1"] + list(extra), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        m = re.search(r"Planned GPUs : (\d+) \((\d+) image pairs\)", r.stdout)
        return int(m.group(1)), int(m.group(2))

    assert plan() == (1, 6)                      # tiny group: one GPU
    assert plan("-gpus", "4") == (4, 6)
    assert plan("-gpus", "16") == (6, 6)         # never more GPUs than image pairs
    assert plan("-gpus", "8", "-targ", "0") == (3, 3)
    assert plan("-gpus", "8", "-n", "2") == (1, 1)
    # file-size estimate: 40 sparse 54 MB .bin files (250k records each) = 780 pairs x 6.25e10 = 4.9e13 pairs:
    # 4.3 s on one GPU, 0.65 + 1.9 s on two, 1.3 + 1.25 s on three -> 2 or 3; the model picks the first minimum
    big = tmp_path / "big"
    big.mkdir()
    names = []
    for i in range(40):
        p = big / f"p{i}.bin"
        with open(p, "wb") as f:
            f.truncate(250000 * 216)
        names.append(str(p))
    (big / "list.txt").write_text("\n".join(names) + "\n")
    r = subprocess.run([build.BIN, str(big / "list.txt"), "-plan", "1"], capture_output=True, text=True)
    assert "Planned GPUs : 2 (780 image pairs)" in r.stdout


def test_bench_reference_arm_contract(built, tmp_path):
    """`bench.py --impl reference` (the reference's own CPU matcher on a bounded sample) prints one JSON line with the
    keys the driver reads; it needs no GPU."""
    import json
    import sys
    from oracle import oracle as O
    if not os.path.exists(O.REF_BIN):
        pytest.skip("oracle/_ref/match_ref not built here")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "descriptor pairs/s" and j["higher_is_better"] is True
    assert j["value"] > 1e6 and j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["workload"].startswith("c4:") and j["scaling"] == "strong" and j["wall"]["seconds"] > 0
