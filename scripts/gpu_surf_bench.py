"""GPU: throughput of the SURF3D producer on one B200, the reference's CPU implementation timed beside
it on the box's cores, parity of the two outputs, per-kernel rooflines.

    python scripts/gpu_surf_bench.py [--size 400] [--steps 10] [--warmup 3] [--points 20000] [--no-ref] [--out FILE]

Workload: one size^3 int16 volume (eight independently seeded synthetic CT-like octants, dense in blobs so that more than
`points` keypoints are detected), threshold 0, `-n points`, SURF3D descriptors (type 0, radius 5) -- run.sh's
per-image surf3d call after resampling (run.sh:80-87).  A step = fs_set_volume (host buffer in: H2D inside the step)
+ fs_detect + fs_select + fs_describe + read-back of keypoints and descriptors.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frog_b200 import surf, synth  # noqa: E402
from oracle import surf_oracle as so  # noqa: E402


def bench_volume(size: int, seed: int = 50) -> np.ndarray:
    """size^3 volume assembled from eight independently seeded octants (a mirrored volume would hold every structure
    twice, hence pairs of exactly equal detector responses, which real images do not have)."""
    half = (size + 1) // 2
    out = np.empty((size, size, size), np.int16)
    k = 0
    for z0 in (0, half):
        for y0 in (0, half):
            for x0 in (0, half):
                v = synth.make_volume((half, half, half), seed + k, blobs_per_mvox=1500.0, texture=250.0)
                out[z0:z0 + half, y0:y0 + half, x0:x0 + half] = v[:size - z0, :size - y0, :size - x0]
                k += 1
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=400)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=20000)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--tile-sweep", action="store_true", help="also time the response map under every thread-to-voxel mapping")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch

    vol = bench_volume(a.size)
    nvox = vol.size
    host = torch.from_numpy(vol).pin_memory()
    p = surf.Producer(0)
    stage = {k: [] for k in ("ms_integral", "ms_response_map", "ms_extrema", "ms_describe")}
    calls = {k: [] for k in ("set_volume", "detect", "select", "describe", "read_back")}
    walls = []
    for it in range(a.warmup + a.steps):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        p.set_volume(host.numpy())
        t.append(time.perf_counter())
        n_det = p.detect(0.0)
        t.append(time.perf_counter())
        p.select(a.points)
        t.append(time.perf_counter())
        p.describe(0, 5, True)
        t.append(time.perf_counter())
        pts, desc = p.points()
        t.append(time.perf_counter())
        if it >= a.warmup:
            walls.append(t[-1] - t[0])
            st = p.stats()
            for k in stage:
                stage[k].append(st[k])
            for i, k in enumerate(calls):
                calls[k].append((t[i + 1] - t[i]) * 1e3)
    st = p.stats()
    med = {k: float(np.median(v)) for k, v in stage.items()}
    gpu_ms = sum(med.values())
    wall = float(np.median(walls))
    layers = p.layers()
    layer_vox = int(st["response_voxels"])
    out = {
        "metric": "voxels/sec through the SURF3D producer (integral volume, response map, extrema, 20k descriptors)",
        "workload": f"{a.size}^3 int16 volume, threshold 0, -n {a.points}, descriptor type 0 radius 5",
        "steps": a.steps, "warmup": a.warmup,
        "value": nvox / wall, "unit": "voxels/s", "ms_per_step": wall * 1e3,
        "gpu_ms": med, "gpu_ms_total": gpu_ms, "host_ms": wall * 1e3 - gpu_ms,
        "call_ms": {k: float(np.median(v)) for k, v in calls.items()},
        "h2d_bytes_per_step": int(vol.nbytes), "d2h_bytes_per_step": int(pts.nbytes + desc.nbytes),
        "n_detected": int(n_det), "n_candidates": int(st["n_candidates"]), "n_points": int(len(pts)), "n_layers": int(st["n_layers"]),
        "n_clamped": int(st["n_clamped"]),
        "roofline": {
            # algorithmic HBM bytes: integral = voxel read + 8 B write + 8 B read + 16 B write; a response layer must read
            # the 8 B integral volume once and write 6 B per layer voxel; the gathers beyond that are L1 / L2 traffic
            "integral": {"bound": "hbm", "bytes": int(nvox * (vol.itemsize + 32)), "ms": med["ms_integral"]},
            "response_map": {"bound": "hbm", "bytes": int(len(layers) * nvox * 8 + 6 * sum(l["responses"].size for l in layers)),
                             "gather_bytes_l1": int(layer_vox * 144 * 8), "ms": med["ms_response_map"]},
            "describe": {"bound": "l2 gathers", "gather_bytes_l1": int(len(pts) * 1000 * 20 * 8), "ms": med["ms_describe"]},
        },
    }
    peak = 6443.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for k, r in out["roofline"].items():
        if "bytes" in r:
            r["achieved_gbs"] = r["bytes"] / (r["ms"] * 1e-3) / 1e9
            r["peak_gbs"] = peak
            r["frac"] = r["achieved_gbs"] / peak
        if "gather_bytes_l1" in r:
            r["gather_gbs"] = r["gather_bytes_l1"] / (r["ms"] * 1e-3) / 1e9
    if a.tile_sweep:
        names = {0: "flat", 1: "32x8x1", 2: "32x4x2", 3: "32x2x4", 4: "32x1x8", 5: "32x4x4", 6: "32x4x2 <=51 regs",
                 7: "32x4x2 <=42 regs", 8: "32x4x1 (128 threads)", 9: "32x2x1 (64 threads)"}
        sweep = {}
        for v, name in names.items():
            surf.debug_set_option("response_tile", v)
            ms = []
            for _ in range(4):
                p.detect(0.0)
                ms.append(p.stats()["ms_response_map"])
            sweep[name] = float(np.median(ms[1:]))
        surf.debug_set_option("response_tile", 5)
        out["response_tile_sweep_ms"] = sweep
    if not a.no_ref:
        t0 = time.perf_counter()
        ref = so.RefSurf(vol)
        rx, rlap, rdesc = ref.update(threshold=0.0, number_of_points=a.points)
        ref_s = time.perf_counter() - t0
        g = np.stack([pts["x"], pts["y"], pts["z"], pts["scale"], pts["response"]], 1)
        same = len(pts) == len(rx)
        out["cpu_baseline"] = {"kind": "reference", "value": nvox / ref_s, "unit": "voxels/s", "seconds": ref_s,
                               "cores": os.cpu_count(), "sample": "the same volume, whole pipeline (vtk3DSURF::Update), OpenMP on all cores"}
        out["parity"] = {
            "n_points": [int(len(pts)), int(len(rx))],
            "points_bits_differ": int(np.count_nonzero(g.view(np.uint32) != rx.view(np.uint32))) if same else None,
            "laplacian_differ": int(np.count_nonzero(pts["laplacian"] != rlap)) if same else None,
            "descriptor_values_differ": int(np.count_nonzero(desc.view(np.uint32) != rdesc.view(np.uint32))) if same else None,
        }
        out["speedup_vs_reference"] = ref_s / wall
    s = json.dumps(out)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
