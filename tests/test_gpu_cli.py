"""GPU: the drop-in executable.  bin/match must write byte-identical pairs.bin files to the
reference binary on every golden case and on a fresh group matched by the reference here."""
import json
import os
import subprocess

import pytest

from conftest import ROOT
from frog_b200 import build, pairsbin, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
CASES = sorted(json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json"))).items())


@pytest.mark.parametrize("name,case", CASES)
def test_cli_golden_bytes(built, golden_dir, tmp_path, name, case):
    out = str(tmp_path / "pairs.bin")
    r = subprocess.run([build.BIN, os.path.join(golden_dir, case["list"]), "-o", out] + case["args"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert f"Nb Match : {case['nb_match']}" in r.stdout and "Pairing... " in r.stdout
    golden = open(os.path.join(golden_dir, name + ".pairs.bin"), "rb").read()
    mine = open(out, "rb").read()
    if mine != golden:
        pytest.fail(str(pairsbin.diff(pairsbin.parse(golden), pairsbin.parse(mine))))


def test_cli_exact_engine_and_stats(built, golden_dir, tmp_path):
    out, stats = str(tmp_path / "p.bin"), str(tmp_path / "s.json")
    r = subprocess.run([build.BIN, os.path.join(golden_dir, "list_bin.txt"), "-o", out, "-d", "1", "-exact", "1",
                        "-stats", stats, "-gpus", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "bin_runsh.pairs.bin"), "rb").read()
    s = json.load(open(stats))
    assert s["image_pairs"] == 6 and s["rows_exact"] == s["rows"]


def test_cli_default_output_name(built, golden_dir, tmp_path):
    r = subprocess.run([build.BIN, os.path.join(golden_dir, "list_bin.txt")], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(tmp_path / "out_list_bin_4.bin")  # match.cpp:671


def test_cli_vs_reference_binary_fresh_group(built, tmp_path):
    """Config 1 of BASELINE.json: 2 keypoint files (~5k points, csv.gz), defaults -d 0.22 -d2 1."""
    if not os.path.exists(O.REF_BIN):
        pytest.skip("oracle/_ref/match_ref was not shipped")
    lst = synth.write_group(str(tmp_path / "g"), "bank", 2, 5000, fmt="csv.gz")
    ref_out, out = str(tmp_path / "ref.bin"), str(tmp_path / "new.bin")
    ref = O.run_ref_binary([lst, "-o", ref_out])
    new = subprocess.run([build.BIN, lst, "-o", out], capture_output=True, text=True)
    assert new.returncode == 0, new.stderr
    assert open(out, "rb").read() == open(ref_out, "rb").read()
    nb = [l for l in ref.stdout.splitlines() if l.startswith("Nb Match")][0]
    assert nb in new.stdout


@pytest.mark.parametrize("gather", ["host", "nccl"])
def test_cli_two_gpus_gather(built, golden_dir, tmp_path, gather):
    """-gpus 2: image pairs sharded over two GPUs, lists merged on the host or gathered to GPU 0 over NCCL;
    either way the file is the reference's, byte for byte."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out, stats = str(tmp_path / "p.bin"), str(tmp_path / "s.json")
    r = subprocess.run([build.BIN, os.path.join(golden_dir, "list_bin.txt"), "-o", out, "-d", "1", "-gpus", "2",
                        "-gather", gather, "-stats", stats], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "bin_runsh.pairs.bin"), "rb").read()
    s = json.load(open(stats))
    assert s["gpus"] == 2 and s["gather"] == gather
