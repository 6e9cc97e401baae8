#!/usr/bin/env python
"""Consumer hand-off at scale: device link build (fm_links_build) on C2 / C3, timed and compared with the stable-sort
restatement of ImageGroup::readPairs (oracle.read_pairs_links_csr)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from frog_b200 import capi, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

out = {}
for name, n_img, n_pts in (("c2", 10, 20000), ("c3", 50, 50000)):
    kps = [synth.make("iid", n_pts, i) for i in range(n_img)]
    pf = [i for i in range(n_img) for j in range(i + 1, n_img)]
    ps = [j for i in range(n_img) for j in range(i + 1, n_img)]
    m = capi.Matcher(0)
    for i, k in enumerate(kps):
        m.upload(i, k.desc, k.scale, k.lap)
    res = m.match(pf, ps, 1.0, 1.0)
    st = m.stats()
    t0 = time.perf_counter()
    offsets, data, ms = res.links(pf, ps)
    t_all = time.perf_counter() - t0
    lists = res.all_pairs()
    res.free()
    m.close()
    t0 = time.perf_counter()
    want_off, want = O.read_pairs_links_csr([(i, j, l) for i, j, l in zip(pf, ps, lists)], {i: n_pts for i in range(n_img)})
    t_cpu = time.perf_counter() - t0
    same = bool(np.array_equal(data, want)) and all(np.array_equal(offsets[i].astype(np.uint64), want_off[i]) for i in range(n_img))
    out[name] = {"matches": int(data.shape[0] // 2), "half_links": int(data.shape[0]), "device_build_ms": ms, "build_plus_fetch_s": t_all,
                 "numpy_stable_sort_restatement_s": t_cpu, "identical": same, "match_gpu_ms": st["ms_total"]}
    print(name, out[name], flush=True)
json.dump(out, open("gpurun_out/r2_links.json", "w"), indent=1)
