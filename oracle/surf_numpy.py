"""numpy restatement of the SURF3D producer's arithmetic.  TEST INFRASTRUCTURE ONLY.

An independent statement of what vtk3DSURF::Update computes, written from the reference sources
(cited per function), used to pin the verbatim build (oracle/_ref/libsurf_ref.so) and the device
kernels to each other: float32 numpy operations are single IEEE operations in the order written,
which is what the reference's scalar code compiles to on x86-64 without FMA contraction.
Volumes are indexed [z, y, x].  Parity status: pinned to the verbatim build by tests/test_surf_oracle.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

_f = np.float32


def cast_shift(volume: np.ndarray) -> np.ndarray:
    """vtkImageCast (clamp, static_cast<int>) then vtkImageShiftScale with shift = -min (vtk3DSURF.cxx:158-176)."""
    lo, hi = float(np.iinfo(np.int32).min), float(np.iinfo(np.int32).max)
    shift = -float(volume.min())
    c = np.trunc(np.clip(volume.astype(np.float64), lo, hi)).astype(np.int32)
    return np.trunc(np.clip((c.astype(np.float64) + shift) * 1.0, lo, hi)).astype(np.int32)


def integral(cast: np.ndarray) -> np.ndarray:
    """ComputeIntegral (integral.cxx:11-121): inclusive prefix sums along x, y, z in unsigned 64-bit."""
    out = cast.astype(np.int64).astype(np.uint64)
    for axis in (2, 1, 0):
        out = np.cumsum(out, axis=axis, dtype=np.uint64)
    return out


def _box(I, x0, y0, z0, sx, sy, sz):
    """BoxIntegralOptim (integral.h:66-115) for index arrays x0 / y0 / z0 (broadcastable)."""
    x1, y1, z1 = x0 - 1, y0 - 1, z0 - 1
    x2, y2, z2 = x0 + sx - 1, y0 + sy - 1, z0 + sz - 1
    with np.errstate(over="ignore"):  # unsigned wrap-around is the reference's arithmetic too
        return _box_terms(I, x1, y1, z1, x2, y2, z2)


def _box_terms(I, x1, y1, z1, x2, y2, z2):
    return (I[z2, y2, x2] - I[z1, y2, x2] - I[z2, y1, x2] - I[z2, y2, x1] + I[z2, y1, x1] + I[z1, y2, x1]
            + I[z1, y1, x2] - I[z1, y1, x1])


def layer_limit(filter_size: int, step: int) -> int:
    return int(np.ceil(_f(filter_size + 1) / _f(step) / _f(2))) + 1  # fasthessian.cxx:366


def response_layer(I: np.ndarray, width: int, height: int, depth: int, step: int, filter_size: int):
    """FastHessian::buildResponseLayer (fasthessian.cxx:343-481).  Returns (responses, laplacian, isblob), zero outside
    the computed interior."""
    b, l, w = (filter_size - 1) // 2, filter_size // 3, filter_size
    m = 2 * l - 1
    inv = _f(1.0) / ((_f(w) * _f(w) * _f(w)) * (_f(w) * _f(w) * _f(w)) * (_f(w) * _f(w) * _f(w)))
    lim = layer_limit(filter_size, step)
    resp = np.zeros((depth, height, width), np.float32)
    lap = np.zeros((depth, height, width), np.uint8)
    blob = np.zeros((depth, height, width), np.uint8)
    if min(width, height, depth) - 2 * lim <= 0:
        return resp, lap, blob
    ax = np.arange(lim, width - lim)[None, None, :] * step
    ay = np.arange(lim, height - lim)[None, :, None] * step
    az = np.arange(lim, depth - lim)[:, None, None] * step
    x, y, z = np.broadcast_arrays(ax, ay, az)

    def B(x0, y0, z0, sx, sy, sz):
        return _box(I, x0, y0, z0, sx, sy, sz).astype(np.float32)  # (float) of an unsigned long long

    three = _f(3)
    Dxx = B(x - b, y - l + 1, z - l + 1, w, m, m) - B(x - l // 2, y - l + 1, z - l + 1, l, m, m) * three
    Dyy = B(x - l + 1, y - b, z - l + 1, m, w, m) - B(x - l + 1, y - l // 2, z - l + 1, m, l, m) * three
    Dzz = B(x - l + 1, y - l + 1, z - b, m, m, w) - B(x - l + 1, y - l + 1, z - l // 2, m, m, l) * three
    Dxy = B(x - l, y - l, z - l + 1, l, l, m) + B(x + 1, y + 1, z - l + 1, l, l, m) - B(x - l, y + 1, z - l + 1, l, l, m) \
        - B(x + 1, y - l, z - l + 1, l, l, m)
    Dyz = B(x - l + 1, y - l, z - l, m, l, l) + B(x - l + 1, y + 1, z + 1, m, l, l) - B(x - l + 1, y - l, z + 1, m, l, l) \
        - B(x - l + 1, y + 1, z - l, m, l, l)
    Dxz = B(x - l, y - l + 1, z - l, l, m, l) + B(x + 1, y - l + 1, z + 1, l, m, l) - B(x - l, y - l + 1, z + 1, l, m, l) \
        - B(x + 1, y - l + 1, z - l, l, m, l)
    c833, c7603 = _f(0.8330), _f(0.7603)
    Sdet2p = Dyy * Dzz + Dxx * Dyy + Dxx * Dzz - c833 * (Dxy * Dxy + Dxz * Dxz + Dyz * Dyz)
    Trace = Dxx + Dyy + Dzz
    d = (Dxx * Dyy * Dzz).astype(np.float64)
    d = d + 2.0 * Dxy.astype(np.float64) * Dyz.astype(np.float64) * Dxz.astype(np.float64) * np.float64(c7603)
    d = d - (Dxx * Dyz * Dyz * c833).astype(np.float64)
    d = d - (Dyy * Dxz * Dxz * c833).astype(np.float64)
    d = d - (Dzz * Dxy * Dxy * c833).astype(np.float64)
    Det = d.astype(np.float32)
    sl = (slice(lim, depth - lim), slice(lim, height - lim), slice(lim, width - lim))
    blob[sl] = (Sdet2p > 0) & (Trace * Det > 0)
    resp[sl] = np.abs(Det * inv)
    lap[sl] = Trace >= 0
    return resp, lap, blob


def _f_round(v) -> int:
    return int(np.floor(_f(v) + _f(0.5)))  # surf.h:86-89, argument converted to float first


def descriptor(I: np.ndarray, x: float, y: float, z: float, scale: float, radius: int = 5, normalize: bool = True) -> np.ndarray:
    """Surf::getDescriptor (surf.cxx:63-156) for one keypoint, scalar loops; expf is this machine's libm expf, the very
    function the reference calls (surf.cxx:227)."""
    libm = C.CDLL("libm.so.6")
    libm.expf.restype = C.c_float
    libm.expf.argtypes = [C.c_float]
    nz, ny, nx = I.shape

    def box(x0, y0, z0, sx, sy, sz):
        return int(_box(I, np.int64(x0), np.int64(y0), np.int64(z0), sx, sy, sz))

    def haar(px, py, pz, s):
        h = s // 2
        def sgn(v):
            v &= (1 << 64) - 1
            return v - (1 << 64) if v >> 63 else v
        hx = sgn(box(px, py - h, pz - h, h, s, s)) - sgn(box(px - h, py - h, pz - h, h, s, s))
        hy = sgn(box(px - h, py, pz - h, s, h, s)) - sgn(box(px - h, py - h, pz - h, s, h, s))
        hz = sgn(box(px - h, py - h, pz, s, s, h)) - sgn(box(px - h, py - h, pz - h, s, s, h))
        return _f(hx), _f(hy), _f(hz)

    fx, fy, fz = _f(x), _f(y), _f(z)
    sc = float(_f(scale))
    ix0, iy0, iz0 = _f_round(fx), _f_round(fy), _f_round(fz)
    half = _f(float(_f(radius - 1.0)) / 2.0)
    s = 2 * _f_round(_f(sc))
    sig = _f(float(_f(2.5)) * sc)
    desc = np.zeros(48, np.float32)
    length = 0.0
    count = 0
    for i in (-radius, 0):
        for j in (-radius, 0):
            for k in (-radius, 0):
                acc = [0.0] * 6
                ixf, jxf, kxf = _f(i) + half, _f(j) + half, _f(k) + half
                xs = _f_round(float(fx) + float(ixf) * sc)
                ys = _f_round(float(fy) + float(jxf) * sc)
                zs = _f_round(float(fz) + float(kxf) * sc)
                for u in range(i, i + radius):
                    for v in range(j, j + radius):
                        for w in range(k, k + radius):
                            sx_, sy_, sz_ = _f_round(ix0 + u * sc), _f_round(iy0 + v * sc), _f_round(iz0 + w * sc)
                            gx, gy, gz = _f(xs) - _f(sx_), _f(ys) - _f(sy_), _f(zs) - _f(sz_)
                            arg = -(gx * gx + gy * gy + gz * gz) / (_f(2.0) * sig * sig)
                            g = float(_f(1.0) / (sig * sig * sig) * _f(libm.expf(float(arg))))
                            hx, hy, hz = haar(sx_, sy_, sz_, s)
                            r = (g * float(hx), g * float(hy), g * float(hz))
                            for q in range(3):
                                acc[q] += r[q]
                                acc[3 + q] += abs(r[q])
                for q in range(6):
                    desc[count] = acc[q]
                    count += 1
                length += (acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2] + acc[3] * acc[3] + acc[4] * acc[4] + acc[5] * acc[5])
    if not normalize:
        return desc
    length = float(np.sqrt(np.float64(length)))
    if length == 0:
        return desc
    return (desc.astype(np.float64) / length).astype(np.float32)
