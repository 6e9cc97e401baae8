#!/usr/bin/env python
"""End-to-end check of `bin/match` at BASELINE.json scale on a GPU box.

    python scripts/scale_check.py --config c3 --gpus 1 [--sub 6] [--fmt bin] [--keep]

Generates the synthetic keypoint group of the named config (frog_b200/synth.py), runs the drop-in
`bin/match` on the whole group, runs the verbatim reference binary (oracle/_ref/match_ref) on the
first `--sub` images (`-n SUB`, the reference's own flag, so both tools read the same files), and
compares: every image-pair block of the sub-group must be identical, the per-image keypoint records
of the sub-group must be byte-identical, and all blocks of the full group must satisfy the
size-independent properties (one match per second-image keypoint, ordered by it, ids in range).
Prints one JSON line; exits non-zero on any mismatch.  Test infrastructure: the oracle is only the
checker here.
"""
import argparse, json, os, re, shutil, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frog_b200 import build, pairsbin, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CONFIGS = {"c1": (2, 5000, "bank", "csv.gz", ["-d", "0.22", "-d2", "1"]),
           "c2": (10, 20000, "iid", "bin", ["-d", "1"]),
           "c2b": (10, 20000, "bank", "bin", ["-d", "1"]),
           "c3": (50, 50000, "iid", "bin", ["-d", "1"]),
           "c4": (200, 20000, "iid", "bin", ["-d", "1", "-d2", "0.8"])}

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--sub", type=int, default=0, help="images the reference is run on (0 = all)")
ap.add_argument("--fmt", default=None)
ap.add_argument("--keep", action="store_true")
ap.add_argument("--gather", default=None, choices=["nccl", "host"], help="bin/match -gather (multi-GPU list hand-off)")
a = ap.parse_args()
n_img, n_pts, kind, fmt, flags = CONFIGS[a.config]
fmt = a.fmt or fmt
tmp = tempfile.mkdtemp(prefix=f"fm_{a.config}_")
res = {"config": a.config, "images": n_img, "keypoints": n_pts, "kind": kind, "fmt": fmt, "flags": flags, "gpus": a.gpus}
ok = True
try:
    t0 = time.time()
    lst = synth.write_group(tmp, kind, n_img, n_pts, fmt=fmt)
    res["generate_s"] = round(time.time() - t0, 2)
    out, stats = os.path.join(tmp, "pairs.bin"), os.path.join(tmp, "stats.json")
    t0 = time.time()
    r = subprocess.run([build.BIN, lst, "-o", out, "-gpus", str(a.gpus), "-stats", stats] + flags + (["-gather", a.gather] if a.gather else []), capture_output=True, text=True)
    res["match_wall_s"] = round(time.time() - t0, 3)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:]); sys.exit(2)
    secs = [float(x) for x in re.findall(r" : ([0-9.eE+-]+)s$", r.stdout, flags=re.M)]
    res["phases_s"] = dict(zip(["load", "prune", "pairing"], secs))
    res["stats"] = json.load(open(stats))
    t0 = time.time()
    mine = pairsbin.parse(out)
    res["parse_s"] = round(time.time() - t0, 2)
    res["pairs_bin_bytes"] = os.path.getsize(out)
    res["matches"] = mine.n_matches()
    # size-independent properties on every block
    npts = [p.shape[0] for p in mine.points]
    bad = 0
    for i, j, m in mine.blocks:
        if m.shape[0] == 0: continue
        if not (np.all(np.diff(m[:, 1].astype(np.int64)) > 0) and int(m[:, 0].max()) < npts[i] and int(m[:, 1].max()) < npts[j]):
            bad += 1
    res["blocks"] = len(mine.blocks); res["blocks_violating_properties"] = bad
    ok &= bad == 0 and len(mine.blocks) == n_img * (n_img - 1) // 2
    sub = a.sub or n_img
    ref_out = os.path.join(tmp, "ref.bin")
    cores = os.cpu_count() or 1
    t0 = time.time()
    rr = O.run_ref_binary([lst, "-o", ref_out, "-n", str(sub)] + flags, threads=cores)
    res["ref_wall_s"] = round(time.time() - t0, 2)
    rsecs = [float(x) for x in re.findall(r" : ([0-9.eE+-]+)s$", rr.stdout, flags=re.M)]
    res["ref_phases_s"] = dict(zip(["load", "prune", "pairing"], rsecs)); res["ref_cores"] = cores; res["ref_images"] = sub
    ref = pairsbin.parse(ref_out)
    mm = mine.block_map()
    n_cmp = n_bad = 0
    for i, j, m in ref.blocks:
        n_cmp += 1
        if not np.array_equal(mm[(i, j)], m): n_bad += 1
    head_ok = all(np.array_equal(ref.points[k].view(np.uint32), mine.points[k].view(np.uint32)) and ref.names[k] == mine.names[k]
                  for k in range(sub))
    res["ref_blocks_compared"] = n_cmp; res["ref_blocks_mismatching"] = n_bad; res["ref_matches"] = ref.n_matches()
    res["keypoint_records_identical"] = bool(head_ok)
    if sub == n_img:
        res["pairs_bin_byte_identical"] = open(out, "rb").read() == open(ref_out, "rb").read()
        ok &= res["pairs_bin_byte_identical"]
    ok &= n_bad == 0 and head_ok and n_cmp == sub * (sub - 1) // 2
    dp = res["stats"]["descriptor_pairs"]
    res["gpu_pairs_per_s_pairing_phase"] = dp / max(res["phases_s"].get("pairing", 1e-9), 1e-9)
    res["gpu_pairs_per_s_kernels"] = dp / max(res["stats"]["gpu_ms_max"] * 1e-3, 1e-9)
    nn = np.array([p.shape[0] for p in ref.points], np.float64)
    res["ref_pairs_per_s"] = ((nn.sum() ** 2 - (nn ** 2).sum()) / 2) / max(res["ref_phases_s"].get("pairing", 1e-9), 1e-9)
finally:
    if not a.keep: shutil.rmtree(tmp, ignore_errors=True)
res["ok"] = bool(ok)
print(json.dumps(res))
sys.exit(0 if ok else 1)
