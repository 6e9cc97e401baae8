"""Golden vectors for the SURF3D producer, written by the UNMODIFIED reference sources
(oracle/_ref/libsurf_ref.so: vtkOpenSURF3D/{integral,fasthessian,surf,vtk3DSURF}.cxx compiled against
oracle/shim_surf).  The reference ships no tests or fixtures of its own for this path.

    python tests/golden/make_surf_golden.py        # needs /root/reference (run in the authoring container)

Cases (synthetic volumes from frog_b200.synth.make_volume, regenerated from the seed at test time):
  small  120 x 104 x 112 (z, y, x) int16, seed 1, spacing (0.7, 0.8, 1.25), origin (-120.5, 33.25, 1000):
         keypoints + SURF3D descriptors, raw Haar descriptors (type 1, radius 3), the detector's push_back order,
         SHA-256 of the shifted volume, of the integral volume and of every response layer's interior, and the three
         keypoint files the reference's own writers produce (csv, csv.gz, bin)
  mid    216 x 200 x 208 int16, seed 2 (three octaves, ~1000 keypoints): keypoints + descriptors + hashes
  f32    96 x 110 x 104 float32 with a fractional minimum (cast / shift rounding), seed 3: keypoints + hashes
  pfile  the small volume described at the points of a csv file (surf3d -p), see make_point_file_case
         (python tests/golden/make_surf_golden.py --only-pfile regenerates this case alone)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from frog_b200 import synth  # noqa: E402
from oracle import oracle, surf_oracle as so  # noqa: E402

OUT = os.path.join(HERE, "surf")

CASES = {
    "small": dict(shape=(120, 104, 112), seed=1, dtype="int16", spacing=(0.7, 0.8, 1.25), origin=(-120.5, 33.25, 1000.0)),
    "mid": dict(shape=(216, 200, 208), seed=2, dtype="int16", spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)),
    "f32": dict(shape=(96, 110, 104), seed=3, dtype="float32", spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)),
}


def case_volume(name):
    c = CASES[name]
    v = synth.make_volume(c["shape"], c["seed"])
    if c["dtype"] == "float32":
        v = v.astype(np.float32) * np.float32(0.37) - np.float32(0.7)
    return v


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def layer_hashes(layers):
    out = []
    for l in layers:
        lim = so.layer_limit(l["filter"], l["step"])
        sl = (slice(lim, l["depth"] - lim), slice(lim, l["height"] - lim), slice(lim, l["width"] - lim))
        out.append(dict(filter=l["filter"], step=l["step"], width=l["width"], height=l["height"], depth=l["depth"], limit=lim,
                        responses=sha(l["responses"][sl]), laplacian=sha(l["laplacian"][sl]), isblob=sha(l["isblob"][sl])))
    return out


def make_point_file_case():
    """surf3d -p (vtk3DSURF::ReadIPoints): the small volume described at given points.  Input: 50 of the small case's
    keypoints in world units plus an empty line and a CRLF line; outputs of the reference: the converted points in
    file order with their descriptors (-n absent), and the csv.gz it writes with -n 30 (all responses are 0, so the
    order is std::partial_sort's on equal keys)."""
    c = CASES["small"]
    vol = case_volume("small")
    g = np.load(os.path.join(OUT, "small.npz"))
    sp, org = np.array(c["spacing"]), np.array(c["origin"])
    ss = float(np.prod(sp) ** (1.0 / 3.0))
    lines = []
    for i, p in enumerate(g["xyzsr"][:50]):
        w = p[:3].astype(np.float64) * sp + org
        lines.append("%.6f,%.6f,%.6f,%.6f" % (w[0], w[1], w[2], float(p[3]) * ss))
        if i == 10:
            lines.append("")
    text = "\n".join(lines[:20]) + "\r\n" + "\n".join(lines[20:]) + "\n"
    path = os.path.join(OUT, "pfile_points.csv")
    open(path, "w", newline="").write(text)
    ref = so.RefSurf(vol, c["spacing"], c["origin"])
    xyzsr, lap, desc = ref.update(threshold=0.0, number_of_points=-1, point_file=path)
    ref2 = so.RefSurf(vol, c["spacing"], c["origin"])
    x2, _, d2 = ref2.update(threshold=0.0, number_of_points=30, point_file=path)
    ref2.write(os.path.join(OUT, "pfile_points_n30.csv.gz"), "csv.gz")
    np.savez_compressed(os.path.join(OUT, "pfile.npz"), xyzsr=xyzsr, lap=lap, desc=desc, n30_xyzsr=x2, n30_desc=d2)
    print("pfile", len(xyzsr), "points;", len(x2), "kept with -n 30")


def main():
    oracle.build(ref=True)
    if "--only-pfile" in sys.argv:
        make_point_file_case()
        return
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name, c in CASES.items():
        vol = case_volume(name)
        ref = so.RefSurf(vol, c["spacing"], c["origin"])
        xyzsr, lap, desc = ref.update(threshold=0.0, number_of_points=20000)
        det_xyzsr, det_lap = ref.detect(0.0)
        m = dict(c, volume=sha(vol), cast=sha(ref.cast_volume()), integral=sha(ref.integral_volume()),
                 layers=layer_hashes(ref.response_layers(0.0)), n_points=int(len(xyzsr)), n_detected=int(len(det_xyzsr)))
        arrays = dict(xyzsr=xyzsr, lap=lap, desc=desc, det_xyzsr=det_xyzsr, det_lap=det_lap)
        if name == "small":
            for fmt in ("csv", "csv.gz", "bin"):
                ref.write(os.path.join(OUT, "small_points." + fmt), fmt)
            ref.write(os.path.join(OUT, "small_points_l9p4.csv.gz"), "csv.gz", gz_opts="9", precision=4)
            ref1 = so.RefSurf(vol, c["spacing"], c["origin"])
            x1, _, d1 = ref1.update(threshold=0.0, number_of_points=40, descriptor_type=1, radius=3)
            arrays.update(raw_xyzsr=x1, raw_desc=d1)
            ref2 = so.RefSurf(vol, c["spacing"], c["origin"])
            x2, _, d2 = ref2.update(threshold=2000.0, number_of_points=30, descriptor_type=0, radius=4, normalize=False)
            arrays.update(r4_xyzsr=x2, r4_desc=d2)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        manifest[name] = m
        print(name, m["n_points"], "keypoints,", m["n_detected"], "detected")
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1)
    make_point_file_case()


if __name__ == "__main__":
    main()
