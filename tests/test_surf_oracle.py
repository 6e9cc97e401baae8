"""CPU: the SURF3D producer's oracle.  Three statements of the reference arithmetic pin each other:
the verbatim build (oracle/_ref/libsurf_ref.so, the reference's own sources), the numpy restatement
(oracle/surf_numpy.py) and the golden vectors the verbatim build wrote (tests/golden/surf, make_surf_golden.py)."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_surf_golden as mg  # noqa: E402
from oracle import surf_numpy as sn, surf_oracle as so  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "surf")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))

needs_ref = pytest.mark.skipif(not so.available(), reason="oracle/_ref/libsurf_ref.so not built (needs /root/reference)")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name", ["small", "mid", "f32"])
def test_numpy_restatement_matches_golden_hashes(name):
    """cast / shift, integral volume and every response layer of the numpy restatement hash to what the verbatim
    reference build produced."""
    m = MANIFEST[name]
    vol = mg.case_volume(name)
    assert mg.sha(vol) == m["volume"], "synthetic volume generator drifted from the fixtures"
    cast = sn.cast_shift(vol)
    assert mg.sha(cast) == m["cast"]
    integral = sn.integral(cast)
    assert mg.sha(integral) == m["integral"]
    layers = m["layers"] if name != "mid" else m["layers"][4:]  # mid: the coarse octaves only (the fine ones take a minute)
    for l in layers:
        assert sn.layer_limit(l["filter"], l["step"]) == l["limit"]
        r, lp, ib = sn.response_layer(integral, l["width"], l["height"], l["depth"], l["step"], l["filter"])
        lim = l["limit"]
        sl = (slice(lim, l["depth"] - lim), slice(lim, l["height"] - lim), slice(lim, l["width"] - lim))
        assert mg.sha(r[sl]) == l["responses"], f"layer {l['filter']}"
        assert mg.sha(lp[sl]) == l["laplacian"] and mg.sha(ib[sl]) == l["isblob"]
        assert not r[:lim].any() and not r[:, :lim].any() and not r[:, :, :lim].any()


def test_numpy_descriptor_matches_golden():
    """Surf::getDescriptor restated with scalar loops and libm's expf: bit-identical to the reference's descriptors."""
    g = gold("small")
    integral = sn.integral(sn.cast_shift(mg.case_volume("small")))
    for i in (0, 11, 40, 64):
        d = sn.descriptor(integral, *g["xyzsr"][i, :4])
        assert np.array_equal(d.view(np.uint32), g["desc"][i].view(np.uint32))
    d = sn.descriptor(integral, *g["r4_xyzsr"][3, :4], radius=4, normalize=False)
    assert np.array_equal(d.view(np.uint32), g["r4_desc"][3].view(np.uint32))


@needs_ref
@pytest.mark.parametrize("name", ["small", "f32"])
def test_verbatim_build_reproduces_golden(name):
    m, c, g = MANIFEST[name], mg.CASES[name], gold(name)
    vol = mg.case_volume(name)
    ref = so.RefSurf(vol, c["spacing"], c["origin"])
    xyzsr, lap, desc = ref.update(threshold=0.0, number_of_points=20000)
    assert mg.sha(ref.cast_volume()) == m["cast"] and mg.sha(ref.integral_volume()) == m["integral"]
    assert np.array_equal(xyzsr.view(np.uint32), g["xyzsr"].view(np.uint32))
    assert np.array_equal(lap, g["lap"]) and np.array_equal(desc.view(np.uint32), g["desc"].view(np.uint32))
    assert mg.layer_hashes(ref.response_layers(0.0)) == m["layers"]
    det, det_lap = ref.detect(0.0)
    assert np.array_equal(det.view(np.uint32), g["det_xyzsr"].view(np.uint32)) and np.array_equal(det_lap, g["det_lap"])


@needs_ref
def test_verbatim_writers_reproduce_golden_files(tmp_path):
    c = mg.CASES["small"]
    ref = so.RefSurf(mg.case_volume("small"), c["spacing"], c["origin"])
    ref.update(threshold=0.0, number_of_points=20000)
    for fmt in ("csv", "csv.gz", "bin"):
        out = str(tmp_path / ("p." + fmt))
        ref.write(out, fmt)
        assert open(out, "rb").read() == open(os.path.join(GOLD, "small_points." + fmt), "rb").read()


def test_interpolation_solve_is_the_truncated_pseudo_inverse():
    """fasthessian.cxx:614-661 solves X = -pinv(H) dD through cv::SVD with singular values below 0.001 of the largest
    dropped.  OpenCV is absent, so this step is pinned to a TOLERANCE: the product's solve (cyclic Jacobi on the
    symmetric H, host and device statement of fs_kernels.cuh) against numpy's SVD-based truncated pseudo-inverse,
    including nearly singular H, to 1e-9.  (The oracle's own stand-in, a one-sided Jacobi SVD in
    oracle/shim_surf/opencv2/opencv.hpp, is a third method; the keypoints it leads to are bit-identical to the
    product's on every golden volume, tests/test_gpu_surf.py.)"""
    from frog_b200 import surf
    if not os.path.exists(surf.build.SURF_LIB):
        pytest.skip("libfrogsurf.so not built")
    rng = np.random.default_rng(5)
    for t in range(300):
        a = rng.standard_normal((4, 4))
        a = a + a.T
        if t % 4 == 0:
            a[:, 3] *= 1e-6
            a[3, :] *= 1e-6
        d = rng.standard_normal(4)
        u, s, vt = np.linalg.svd(a)
        sinv = np.array([1 / s[0]] + [0 if s[i] / s[0] < 0.001 else 1 / s[i] for i in range(1, 4)])
        want = -(vt.T * sinv) @ u.T @ d
        h10 = [a[0, 0], a[1, 1], a[2, 2], a[3, 3], a[0, 1], a[0, 2], a[0, 3], a[1, 2], a[1, 3], a[2, 3]]
        got = surf.debug_solve_offsets(d, h10)
        assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
