#!/usr/bin/env python
"""Stand-in for the consumer stage of config 5: reads pairs.bin the way ImageGroup::readPairs does
(registration/imageGroup.cxx:1353-1417) -- header, per-image records, pair blocks to end of file -- and fails on a
malformed file.  Prints what frog prints about its input: images, points, pairs."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from frog_b200 import pairsbin  # noqa: E402

pf = pairsbin.parse(sys.argv[1])
for i, j, m in pf.blocks:
    assert i < len(pf.points) and j < len(pf.points)
    if m.shape[0]:
        assert m[:, 0].max() < pf.points[i].shape[0] and m[:, 1].max() < pf.points[j].shape[0], "match index out of range"
print(f"read {len(pf.points)} images, {sum(p.shape[0] for p in pf.points)} points, {len(pf.blocks)} pair blocks, {pf.n_matches()} pairs")
