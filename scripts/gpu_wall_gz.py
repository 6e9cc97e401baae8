"""GPU box: wall time and load phase of `bin/match` on a 200 x 20 000 `.csv.gz` group (C4's shape; 16 distinct images
written cyclically -- the load phase does not care), one-shot and through the resident server.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from frog_b200 import synth  # noqa: E402

base = [synth.make("iid", 20000, i) for i in range(16)]
kps = [base[i % 16] for i in range(200)]
out = bench.bin_match_wall(kps, "iid", 1.0, 0.8, 1, ["csv.gz"])
out["cores"] = os.cpu_count()
print(json.dumps(out))
