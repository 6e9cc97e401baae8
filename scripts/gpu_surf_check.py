"""GPU: stage-by-stage comparison of libfrogsurf.so with the verbatim reference producer (oracle/_ref/libsurf_ref.so)
on synthetic volumes; prints what differs and where.  Usage: python scripts/gpu_surf_check.py [nz ny nx [seed]]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frog_b200 import surf, synth  # noqa: E402
from oracle import surf_oracle as so  # noqa: E402


def main():
    shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (120, 104, 112)
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    npts = 20000
    vol = synth.make_volume(shape, seed)
    out = {"shape": shape, "seed": seed}
    t = time.time()
    ref = so.RefSurf(vol)
    rx, rlap, rdesc = ref.update(threshold=0.0, number_of_points=npts)
    out["ref_update_s"] = time.time() - t
    p = surf.Producer(0)
    p.keep_cast_volume(True)
    p.set_volume(vol)
    out["cast_equal"] = bool(np.array_equal(p.cast_volume(), ref.cast_volume()))
    out["integral_equal"] = bool(np.array_equal(p.integral(), ref.integral_volume()))
    n = p.detect(0.0)
    st = p.stats()
    # layers
    rl = ref.response_layers(0.0)
    gl = p.layers()
    out["layers"] = []
    for a, b in zip(gl, rl):
        lim = so.layer_limit(b["filter"], b["step"])
        sl = (slice(lim, b["depth"] - lim), slice(lim, b["height"] - lim), slice(lim, b["width"] - lim))
        geom = all(a[k] == b[k] for k in ("width", "height", "depth", "step", "filter"))
        e = dict(filter=b["filter"], geom=geom,
                 resp_bits=int(np.count_nonzero(a["responses"][sl].view(np.uint32) != b["responses"][sl].view(np.uint32))),
                 lap=int(np.count_nonzero(a["laplacian"][sl] != b["laplacian"][sl])),
                 blob=int(np.count_nonzero(a["isblob"][sl] != b["isblob"][sl])),
                 interior=int(a["responses"][sl].size),
                 outside_nonzero=int(np.count_nonzero(a["responses"]) - np.count_nonzero(a["responses"][sl])))
        out["layers"].append(e)
    # detection, push_back order
    dx, dlap = ref.detect(0.0)
    pts, _ = p.points(with_descriptors=False)
    out["n_detect"] = [int(n), int(len(dx))]
    if n == len(dx):
        g = np.stack([pts["x"], pts["y"], pts["z"], pts["scale"], pts["response"]], 1)
        out["detect_bits_differ"] = int(np.count_nonzero(g.view(np.uint32) != dx.view(np.uint32)))
        out["detect_max_abs"] = float(np.abs(g - dx)[:, :4].max()) if n else 0.0
        out["detect_lap_differ"] = int(np.count_nonzero(pts["laplacian"] != dlap))
        out["detect_response_bits"] = int(np.count_nonzero(g[:, 4].view(np.uint32) != dx[:, 4].view(np.uint32)))
    # full pipeline
    p.select(npts)
    p.describe(0, 5, True)
    pts, desc = p.points()
    out["n_final"] = [int(len(pts)), int(len(rx))]
    if len(pts) == len(rx):
        g = np.stack([pts["x"], pts["y"], pts["z"], pts["scale"], pts["response"]], 1)
        same_pts = np.all(g.view(np.uint32) == rx.view(np.uint32), 1)
        out["final_points_identical"] = int(same_pts.sum())
        dd = desc.view(np.uint32) != rdesc.view(np.uint32)
        out["desc_rows_differ_on_identical_points"] = int(np.count_nonzero(dd.any(1) & same_pts))
        out["desc_values_differ"] = int(np.count_nonzero(dd))
        out["desc_max_abs"] = float(np.abs(desc - rdesc).max()) if len(pts) else 0.0
    # descriptors on the reference's own points: bit for bit
    p.set_points(rx[:, :4])
    p.describe(0, 5, True)
    _, desc2 = p.points()
    dd = desc2.view(np.uint32) != rdesc.view(np.uint32)
    out["desc_on_ref_points_values_differ"] = int(np.count_nonzero(dd))
    out["desc_on_ref_points_rows_differ"] = int(np.count_nonzero(dd.any(1)))
    if dd.any():
        i = int(np.argmax(dd.any(1)))
        out["first_bad_row"] = dict(i=i, got=desc2[i][:8].tolist(), want=rdesc[i][:8].tolist(), pt=rx[i].tolist())
    out["stats"] = p.stats()
    # raw Haar descriptors
    ref1 = so.RefSurf(vol)
    r1x, _, r1desc = ref1.update(threshold=0.0, number_of_points=500, descriptor_type=1, radius=3)
    p.set_points(r1x[:, :4])
    p.describe(1, 3, True)
    _, d1 = p.points()
    out["raw_desc_values_differ"] = int(np.count_nonzero(d1.view(np.uint32) != r1desc.view(np.uint32)))
    out["det_stats"] = st
    print(json.dumps(out, indent=1, default=str))


if __name__ == "__main__":
    main()
