"""Python face of the SURF3D producer oracle.  TEST INFRASTRUCTURE ONLY.

`RefSurf` drives oracle/_ref/libsurf_ref.so: the UNMODIFIED reference sources
vtkOpenSURF3D/{integral,fasthessian,surf,vtk3DSURF}.cxx compiled against oracle/shim_surf (VTK and
OpenCV are absent in this image; see the shim headers for what they stand in for).  It replays
surf3d.cxx:258-328 -- options, vtk3DSURF::Update(), the point writers -- on a volume given as a numpy
array, and exposes the intermediate products so each device stage is compared on its own.

Parity status: integral volume, response layers (responses / laplacian / isblob), extremum
detection, sort / prune, descriptors and the writers are the reference's own compiled code: PINNED.
The sub-voxel interpolation (fasthessian.cxx:614-661) calls cv::SVD from OpenCV, which is absent:
the shim's SVD restates the published one-sided Jacobi algorithm, so offsets are pinned to a
tolerance (1e-9 absolute on the offsets), not to OpenCV's bits.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SURF_REF_LIB = os.path.join(HERE, "_ref", "libsurf_ref.so")

_VTK = {np.dtype(np.int16): 4, np.dtype(np.uint16): 5, np.dtype(np.int32): 6, np.dtype(np.float32): 10,
        np.dtype(np.uint8): 3, np.dtype(np.float64): 11}


def available() -> bool:
    return os.path.exists(SURF_REF_LIB)


def _lib():
    L = C.CDLL(SURF_REF_LIB)
    L.sr_create.restype = C.c_void_p
    L.sr_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.sr_destroy.argtypes = [C.c_void_p]
    L.sr_update.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
    for f in ("sr_num_points", "sr_descriptor_size"):
        getattr(L, f).restype = C.c_int64
        getattr(L, f).argtypes = [C.c_void_p]
    L.sr_points.argtypes = [C.c_void_p] * 4
    L.sr_cast_volume.argtypes = [C.c_void_p, C.c_void_p]
    L.sr_integral_volume.argtypes = [C.c_void_p, C.c_void_p]
    for f in ("sr_write_csv", "sr_write_bin", "sr_write_json"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_char_p]
    L.sr_write_csvgz.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    L.sr_build_response_map.restype = C.c_int
    L.sr_build_response_map.argtypes = [C.c_void_p, C.c_double]
    L.sr_layer_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.sr_layer_data.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sr_detect.restype = C.c_int64
    L.sr_detect.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int64]
    return L


class RefSurf:
    """One volume through the reference producer.  `volume` is indexed [z, y, x] (x fastest, as VTK stores it)."""

    def __init__(self, volume: np.ndarray, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)):
        if not available():
            raise FileNotFoundError(f"{SURF_REF_LIB} not built -- run `make -C oracle surfref`")
        self.L = _lib()
        v = np.ascontiguousarray(volume)
        self.shape = v.shape
        dims = (C.c_int * 3)(v.shape[2], v.shape[1], v.shape[0])
        sp = (C.c_double * 3)(*spacing)
        org = (C.c_double * 3)(*origin)
        self.h = self.L.sr_create(v.ctypes.data, _VTK[v.dtype], dims, sp, org)
        self._updated = False

    def close(self):
        if self.h:
            self.L.sr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update(self, threshold=0.0, number_of_points=-1, descriptor_type=0, radius=5, normalize=True, point_file=None):
        self.L.sr_update(self.h, threshold, number_of_points, descriptor_type, radius, int(normalize),
                         point_file.encode() if point_file else None)
        self._updated = True
        return self.points()

    def points(self):
        """(xyzsr [n,5] float32 -- x, y, z, scale, response in voxel units --, lap [n] int32, desc [n,D] float32)"""
        n = self.L.sr_num_points(self.h)
        d = self.L.sr_descriptor_size(self.h)
        xyzsr = np.zeros((n, 5), np.float32)
        lap = np.zeros(n, np.int32)
        desc = np.zeros((n, d), np.float32)
        if n:
            self.L.sr_points(self.h, xyzsr.ctypes.data, lap.ctypes.data, desc.ctypes.data if d else None)
        return xyzsr, lap, desc

    def cast_volume(self) -> np.ndarray:
        out = np.zeros(self.shape, np.int32)
        self.L.sr_cast_volume(self.h, out.ctypes.data)
        return out

    def integral_volume(self) -> np.ndarray:
        out = np.zeros(self.shape, np.uint64)
        self.L.sr_integral_volume(self.h, out.ctypes.data)
        return out

    def response_layers(self, threshold=0.0):
        """list of dicts: width/height/depth/step/filter + responses, laplacian, isblob as [depth, height, width]"""
        n = self.L.sr_build_response_map(self.h, threshold)
        out = []
        for i in range(n):
            info = (C.c_int * 5)()
            self.L.sr_layer_info(self.h, i, info)
            w, h, d, step, filt = list(info)
            r = np.zeros((d, h, w), np.float32)
            lp = np.zeros((d, h, w), np.uint8)
            ib = np.zeros((d, h, w), np.uint8)
            self.L.sr_layer_data(self.h, i, r.ctypes.data, lp.ctypes.data, ib.ctypes.data)
            out.append(dict(width=w, height=h, depth=d, step=step, filter=filt, responses=r, laplacian=lp, isblob=ib))
        return out

    def detect(self, threshold=0.0, cap=1 << 22):
        xyzsr = np.zeros((cap, 5), np.float32)
        lap = np.zeros(cap, np.int32)
        n = self.L.sr_detect(self.h, threshold, xyzsr.ctypes.data, lap.ctypes.data, cap)
        return xyzsr[:n].copy(), lap[:n].copy()

    def write(self, path: str, fmt: str, gz_opts=None, precision=-1):
        p = path.encode()
        if fmt == "csv":
            self.L.sr_write_csv(self.h, p)
        elif fmt == "bin":
            self.L.sr_write_bin(self.h, p)
        elif fmt == "csv.gz":
            self.L.sr_write_csvgz(self.h, p, gz_opts.encode() if gz_opts else None, precision)
        elif fmt == "json":
            self.L.sr_write_json(self.h, p)
        else:
            raise ValueError(fmt)


def layer_limit(filter_size: int, step: int) -> int:
    """Border (in layer voxels) outside which buildResponseLayer computes nothing (fasthessian.cxx:366)."""
    return int(np.ceil(np.float32(filter_size + 1) / np.float32(step) / np.float32(2))) + 1
