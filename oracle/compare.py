"""Comparison of match lists against the reference's, with the north star's classification of differences.

TEST INFRASTRUCTURE ONLY (tests/, scripts/ and bench.py's cpu_baseline leg): given the per-image-pair
(first, second) lists of the product and of the reference on the same keypoints, report

* identical blocks (same pairs, same order),
* differing pairs, each classified with the REFERENCE arithmetic (oracle.norm_matrix_numpy, match.cpp:243-251
  and the acceptance rule :320-321): a difference is a "threshold-epsilon" case when the reference's
  sqrt(d1) lies within EPS_ULP float32 ulps of -d, or its sqrt(d1/d2) within EPS_ULP ulps of -d2;
  anything else is a real failure.

The product decides in the reference's own FP32 arithmetic, so both counts are expected to be 0; the
classification exists so that a non-zero count is reported for what it is.
"""
from __future__ import annotations

import numpy as np

from . import oracle

EPS_ULP = 4  # epsilon = 4 ulp of the FP32 value


def _ulps(a: np.float32, b: np.float32) -> float:
    a, b = np.float32(a), np.float32(b)
    if not (np.isfinite(a) and np.isfinite(b)):
        return float("inf")
    return abs(float(a) - float(b)) / float(np.spacing(np.float32(max(abs(a), abs(b), np.finfo(np.float32).tiny))))


def row_top2(first, second, row: int):
    """(d1, d2, match) of row `row` of image `second` over image `first` as match.cpp:262-317 computes them."""
    d1_, s1_, l1_ = first
    d2_, s2_, l2_ = second
    dist = oracle.norm_matrix_numpy(d2_[row:row + 1], d1_)[0]
    ok = oracle.gate_matrix_numpy(s2_[row:row + 1], l2_[row:row + 1], s1_, l1_)[0]
    big = np.float32(np.finfo(np.float32).max)
    dist = np.where(ok, dist, np.float32(np.inf))
    if dist.size == 0 or not np.isfinite(dist).any():
        return big, big, -1
    m = int(np.argmin(dist))
    d1 = np.float32(dist[m])
    rest = np.delete(dist, m)
    d2 = np.float32(rest.min()) if rest.size and np.isfinite(rest.min()) else big
    return d1, d2, m


def classify_row(first, second, row: int, dist_thr: float, ratio_thr: float):
    """'eps' if the reference's decision for this row sits within EPS_ULP ulps of a threshold, else 'fail'."""
    d1, d2, _ = row_top2(first, second, row)
    thr, rat = np.float32(dist_thr), np.float32(ratio_thr)
    big = np.float32(np.finfo(np.float32).max)
    near = _ulps(np.sqrt(d1), thr) <= EPS_ULP
    if d2 != big:
        with np.errstate(divide="ignore", invalid="ignore"):
            near = near or _ulps(np.sqrt(np.float32(d1 / d2)), rat) <= EPS_ULP
    return "eps" if near else "fail"


def compare_blocks(ours: dict, ref: dict, images, dist_thr: float, ratio_thr: float, classify_limit: int = 200) -> dict:
    """ours / ref: {(i, j): [m,2] uint32}; images: list of (desc, scale, lap) in pairs.bin id order.
    Only non-sym blocks can be classified row by row (second index = outer-loop row)."""
    rep = {"blocks": 0, "identical_blocks": 0, "pairs_ref": 0, "pairs_ours": 0, "differing_pairs": 0,
           "threshold_eps_count": 0, "failures": 0, "eps_ulp": EPS_ULP, "missing_blocks": 0}
    for key, r in ref.items():
        rep["blocks"] += 1
        rep["pairs_ref"] += int(r.shape[0])
        o = ours.get(key)
        if o is None:
            rep["missing_blocks"] += 1
            rep["failures"] += int(r.shape[0]) + 1
            continue
        rep["pairs_ours"] += int(o.shape[0])
        if o.shape == r.shape and np.array_equal(o, r):
            rep["identical_blocks"] += 1
            continue
        so = {(int(a), int(b)) for a, b in o}
        sr = {(int(a), int(b)) for a, b in r}
        diff = sorted(so ^ sr)
        rep["differing_pairs"] += len(diff)
        if not diff:  # same set, different order
            rep["failures"] += 1
            continue
        rows = sorted({b for _, b in diff})
        for row in rows[:classify_limit]:
            kind = classify_row(images[key[0]], images[key[1]], row, dist_thr, ratio_thr)
            rep["threshold_eps_count" if kind == "eps" else "failures"] += 1
        rep["failures"] += max(0, len(rows) - classify_limit)
    rep["set_identical"] = rep["differing_pairs"] == 0 and rep["missing_blocks"] == 0
    return rep
