#!/usr/bin/env python
"""Throughput of the exact FP32 side paths (CUDA cores): -exact at d = 48, the K-chunked kernel at d = 1000 / 3000
(surf3d -type 2 / 1 sizes), and the -all mode.  Prints one JSON object."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frog_b200 import capi, synth

m = capi.Matcher(0)
out = {}


def run(tag, images, thr, ratio=1.0, **kw):
    n = len(images)
    pf = [i for i in range(n) for j in range(i + 1, n)]
    ps = [j for i in range(n) for j in range(i + 1, n)]
    m.clear()
    for i, (d, s, l) in enumerate(images):
        m.upload(i, d, s, l)
    best = None
    for _ in range(3):
        res = m.match(pf, ps, thr, ratio, device_only=True, **kw)
        st = m.stats()
        tot = res.total
        res.free()
        ms = st["ms_total"]
        best = ms if best is None else min(best, ms)
    pairs = st["descriptor_pairs"]
    out[tag] = {"descriptor_pairs": pairs, "ms": round(best, 3), "pairs_per_s": pairs / (best * 1e-3), "matches": tot,
                "d": int(images[0][0].shape[1]), "flop_per_s": 3.0 * images[0][0].shape[1] * pairs / (best * 1e-3)}
    print(tag, out[tag], flush=True)


kps = [synth.make("iid", 10000, i) for i in range(6)]
img48 = [(k.desc, k.scale, k.lap) for k in kps]
run("exact_d48 (-exact 1)", img48, 1.0, force_exact=True)
run("tensor_d48 (same group)", img48, 1.0)
run("all_d48 (-all, -d 0.3)", img48, 0.3, match_all=True)
run("all_d48 (-all, -d 1.2: most gated-in pairs emit)", img48[:3], 1.2, match_all=True)
rng = np.random.default_rng(1)
for d, n in ((1000, 4000), (3000, 2500)):
    imgs = []
    for i in range(4):
        x = rng.standard_normal((n, d)).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        imgs.append((x, kps[i].scale[:n].copy(), kps[i].lap[:n].copy()))
    run(f"generic_d{d}", imgs, 1.5)
m.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/side_paths.json", "w"), indent=1)
