#!/usr/bin/env python
"""CPU model of the scoring kernel's capture path (fm_score.cuh), to size design alternatives without a GPU.

For one image pair of the C2 workload it replays, warp by warp (32 sorted rows) and step by step (32 sorted columns),
what the epilogue does: the running threshold thr = g2 - 2 eps from 16-column chunk maxima, the look-ahead visit of
the first tiles, and counts
  * slow-path entries: steps in which ANY of the warp's 32 rows has a column above its threshold,
  * captured columns per row,
for (a) no look-ahead, (b) the look-ahead the kernel uses, (c) an oracle that knows every row's final threshold from
the first column on -- the floor for any threshold-seeding scheme at this vote granularity -- and (d) the same oracle
with votes over 16 or 8 rows instead of 32 (what a finer vote granularity would buy).
Scores are computed in float32 from the float32 descriptors (the FP16 rounding moves them by < eps and does not change
the statistics).  Prints one JSON object.

    python scripts/sim_capture.py [keypoints_per_image [look-ahead depths, comma separated]]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frog_b200 import synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000  # keypoints per image (20000 = C2 / C4, 50000 = C3)
PRES = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8]  # look-ahead depths (tiles) to replay
EPS2, UNIT, TILE, STEP = 1.3e-3, 256, 64, 32
a, b = synth.make("iid", N, 0), synth.make("iid", N, 1)  # columns = image `first`, rows = image `second`


def order(k):
    return np.lexsort((k.scale, k.lap))


pa, pb = order(a), order(b)
A, B = a.desc[pa], b.desc[pb]
sa, sb, la, lb = a.scale[pa], b.scale[pb], a.lap[pa], b.lap[pb]
half_norm = 0.5 * (A.astype(np.float64) ** 2).sum(1).astype(np.float32)
rng = np.random.default_rng(0)
units = rng.choice(N // UNIT, 12, replace=False)
KEYS = ["no_lookahead"] + [f"lookahead{p}" for p in PRES] + ["oracle32", "oracle16", "oracle8"]
tot = {k: dict(entries=0, steps=0, captured=0, rows=0, hit_rows=0) for k in KEYS}

for u in units:
    r0 = u * UNIT
    rows = slice(r0, r0 + UNIT)
    gate = (lb[rows, None] == la[None, :]) & ~((sb[rows, None] / sa[None, :] > np.float32(1.3)) | (sa[None, :] / sb[rows, None] > np.float32(1.3)))
    cols = np.nonzero(gate.any(0))[0]
    if len(cols) == 0:
        continue
    t0, t1 = cols[0] // TILE, cols[-1] // TILE + 1
    c0, c1 = t0 * TILE, min(t1 * TILE, N)
    t = B[rows] @ A[c0:c1].T - half_norm[None, c0:c1]
    t = np.where(gate[:, c0:c1], t, -np.inf).astype(np.float32)
    pad = (-t.shape[1]) % STEP
    if pad:
        t = np.concatenate([t, np.full((UNIT, pad), -np.inf, np.float32)], 1)
    n_steps = t.shape[1] // STEP
    ts = t.reshape(UNIT, n_steps, STEP)
    chunk_max = ts.reshape(UNIT, n_steps, 2, 16).max(3)  # [row][step][2]
    srt = np.sort(t, 1)
    final_thr = srt[:, -2] - EPS2  # second best score of the row - 2 eps

    def replay(n_pre_tiles):
        n_pre = min(n_pre_tiles, (n_steps // 2) // 4) * 2  # look-ahead depth in steps (two steps per 64-column tile)
        g1 = np.full(UNIT, -np.inf, np.float32)
        g2 = np.full(UNIT, -np.inf, np.float32)
        entries = np.zeros(UNIT // 32, np.int64)
        captured = np.zeros(UNIT, np.int64)
        hit_rows = 0
        visits = [(s, True, False) for s in range(n_pre)] + [(s, s >= n_pre, True) for s in range(n_steps)]
        for s, upd, cap in visits:
            if upd:
                hi, lw = chunk_max[:, s].max(1), chunk_max[:, s].min(1)
                g2 = np.maximum(np.maximum(g2, lw), np.minimum(g1, hi))
                g1 = np.maximum(g1, hi)
            if cap:
                thr = g2 - np.float32(EPS2)
                unseeded = np.isneginf(g2)  # the kernel seeds these from node maxima: model as "second best of this step"
                thr = np.where(unseeded, np.sort(ts[:, s], 1)[:, -2] - np.float32(EPS2), thr)
                above = ts[:, s] > thr[:, None]
                captured += above.sum(1)
                entries += above.any(1).reshape(-1, 32).any(1)
                hit_rows += int(above.any(1).sum())
        return entries.sum(), captured.sum(), n_steps * (UNIT // 32), hit_rows

    def oracle(group):
        above = ts > final_thr[:, None, None]
        any_row = above.any(2)  # [row][step]
        ent = any_row.reshape(UNIT // group, group, n_steps).any(1).sum()
        return ent, above.sum(), n_steps * (UNIT // group), int(any_row.sum())

    results = [("no_lookahead", replay(0))] + [(f"lookahead{p}", replay(p)) for p in PRES]
    results += [("oracle32", oracle(32)), ("oracle16", oracle(16)), ("oracle8", oracle(8))]
    for key, (e, c, s, h) in results:
        tot[key]["hit_rows"] += int(h)
        tot[key]["entries"] += int(e)
        tot[key]["captured"] += int(c)
        tot[key]["steps"] += int(s)
        tot[key]["rows"] += UNIT

out = {k: {"entry_rate": round(v["entries"] / v["steps"], 3), "captured_per_row": round(v["captured"] / v["rows"], 2),
           "hitting_rows_per_entry": round(v["hit_rows"] / max(v["entries"], 1), 2),
           "captured_columns_per_entry": round(v["captured"] / max(v["entries"], 1), 2), "vote_steps": v["steps"]}
       for k, v in tot.items()}
print(json.dumps(out, indent=1))
