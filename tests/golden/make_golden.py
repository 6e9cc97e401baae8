"""Generate the golden fixtures in this directory with the UNMODIFIED reference matcher.

Run in the authoring container (needs /root/reference):  python tests/golden/make_golden.py

It writes small synthetic keypoint groups (frog_b200/synth.py, fixed seeds) in all three input
formats, runs oracle/_ref/match_ref -- the verbatim reference match.cpp compiled by
oracle/Makefile -- on them with the flag sets the reference's callers use (run.sh, FROG.py,
tools/register.py, desk UI) plus the edge-case sets, and stores every pairs.bin.
manifest.json records the command line of each case; the inputs and outputs are committed so the
tests never need /root/reference.
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from frog_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

N_IMAGES, N_POINTS = 4, 220


def write_inputs():
    kps = [synth.make("bank", N_POINTS, i) for i in range(N_IMAGES)]
    # edge cases folded into the group: an exact duplicate descriptor inside image 1 (d1 == d2 tie
    # for any row matching it), a scale sitting exactly on the 1.3f boundary, a third laplacian value
    kps[1].desc[5] = kps[1].desc[4]
    kps[1].scale[5] = kps[1].scale[4]
    kps[1].lap[5] = kps[1].lap[4]
    kps[2].scale[7] = np.float32(np.float32(1.3) * kps[0].scale[7])
    kps[2].lap[7] = kps[0].lap[7]
    kps[3].lap[10:14] = 2.0
    kps[0].lap[10:14] = 2.0
    for fmt in ("bin", "csv", "csv.gz"):
        for i, kp in enumerate(kps):
            synth.WRITERS[fmt](kp, os.path.join(HERE, f"points{i}.{fmt}"))
    # a hand-formatted CSV exercising the parser: CRLF line ends, %g / exponent cells, blanks,
    # a short line that must be dropped (<= 6 cells), a trailing comma
    rec = kps[0].records()[:40].astype(np.float64)
    lines = []
    for r, row in enumerate(rec):
        cells = ["%.9g" % v if (r + k) % 3 else " %e" % v for k, v in enumerate(row)]
        line = ",".join(cells)
        if r % 5 == 0:
            line += ","
        lines.append(line + ("\r" if r % 2 else ""))
        if r == 7:
            lines.append("1,2,3,4,5,6")
        if r == 9:
            lines.append("")
    with open(os.path.join(HERE, "quirks.csv"), "w", newline="") as f:
        f.write("\n".join(lines) + "\n")


def list_file(name, entries):
    with open(os.path.join(HERE, name), "w") as f:
        f.write("\n".join(entries) + "\n")


CASES = {
    # name: (list file, extra argv)
    "bin_default": ("list_bin.txt", []),
    "bin_runsh": ("list_bin.txt", ["-d", "1"]),
    "bin_frogpy": ("list_bin.txt", ["-d", "10000000000", "-np", "20000", "-d2", "1"]),
    "bin_ratio": ("list_bin.txt", ["-d", "1", "-d2", "0.8"]),
    "bin_desk": ("list_bin.txt", ["-d", "0.5", "-d2", "0.998"]),
    "bin_tie_accept": ("list_bin.txt", ["-d", "1", "-d2", "1.01"]),
    "bin_sym": ("list_bin.txt", ["-d", "1", "-sym"]),
    "bin_targ": ("list_bin.txt", ["-d", "1", "-targ", "1"]),
    "bin_prune": ("list_bin.txt", ["-d", "1", "-np", "100", "-sp", "500"]),
    "bin_n3": ("list_bin.txt", ["-d", "1", "-n", "3"]),
    "gz_rigid_zwin": ("list_gz.txt", ["-d", "1", "-zmin", "100", "-zmax", "1200"]),
    "gz_default": ("list_gz.txt", []),
    "csv_relative": ("list_csv.txt", ["-d", "1", "-d2", "0.9"]),
    "csv_quirks": ("list_quirks.txt", ["-d", "1"]),
    # -all (desk UI): every gated-in column under -d emits a pair naming the stale running nearest (match.cpp:295-300);
    # the key is a flag but the parser skips the token after it
    "bin_all": ("list_bin.txt", ["-d", "0.3", "-all", "1"]),
    "bin_all_sym": ("list_bin.txt", ["-all", "1", "-d", "0.45", "-sym"]),
}


def main():
    oracle.build(ref=True)
    write_inputs()
    # list files use paths relative to this directory through the "parent/NAME.csv" rule
    # (match.cpp:471-475) for csv, and absolute paths written at test time for the others; to keep
    # the fixtures relocatable the absolute lists are TEMPLATES with {DIR} expanded by the tests.
    list_file("list_bin.txt", ["{DIR}/points%d.bin" % i for i in range(N_IMAGES)])
    list_file("list_gz.txt", ["{DIR}/points%d.csv.gz,%.1f,%.1f,%.1f" % (i, 0.5 * i, -0.25 * i, 50.0 * i) for i in range(N_IMAGES)])
    list_file("list_csv.txt", ["points%d" % i for i in range(N_IMAGES)])
    list_file("list_quirks.txt", ["quirks", "points1"])
    manifest = {}
    work = os.path.join(HERE, "_work")
    for name, (lst, extra) in CASES.items():
        shutil.rmtree(work, ignore_errors=True)
        os.makedirs(work)
        # materialise the list next to the keypoint files (the relative rule needs that)
        text = open(os.path.join(HERE, lst)).read().replace("{DIR}", HERE)
        tmp_list = os.path.join(HERE, "_" + lst)
        open(tmp_list, "w").write(text)
        out = os.path.join(HERE, f"{name}.pairs.bin")
        cmd = [oracle.REF_BIN, tmp_list, "-o", out] + extra
        res = subprocess.run(cmd, capture_output=True, text=True, check=True)
        os.remove(tmp_list)
        nb = [l for l in res.stdout.splitlines() if l.startswith("Nb Match")]
        manifest[name] = {"list": lst, "args": extra, "nb_match": int(nb[0].split(":")[1])}
        print(name, manifest[name]["nb_match"], os.path.getsize(out))
    shutil.rmtree(work, ignore_errors=True)
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
