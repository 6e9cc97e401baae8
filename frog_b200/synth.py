"""Synthetic SURF3D keypoint sets in the three on-disk formats `bin/match` reads.

Test/bench data only.  The record layout and value distributions follow the reference's
producer:

* record  = x, y, z, scale, laplacianSign, response, d0..d47
            (vtkOpenSURF3D/vtk3DSURF.cxx:402-525 -- the .csv / .csv.gz / .bin writers)
* descriptor = 8 sub-blocks x (dx, dy, dz, |dx|, |dy|, |dz|), L2-normalised
            (vtkOpenSURF3D/surf.cxx:135-155) => components 3,4,5 (mod 6) are >= 0
* scale   = spacing * 0.1333 * (m + u*step), filter sizes m/step per octave
            (vtkOpenSURF3D/fasthessian.cxx:144,285-289,599-605)
* laplacian in {0, 1} (fasthessian.cxx:453); response sorted descending (vtk3DSURF.cxx:209-221)

Every value is rounded to 6 decimals before being narrowed to float32 so that the "%f" text
formats and the raw-float .bin format carry bit-identical float32 values.
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass

import numpy as np

D = 48
_FILTERS = np.array([(15, 6), (21, 6), (27, 12), (39, 12), (51, 24), (75, 24), (99, 48), (147, 48)], dtype=np.float64)
_SPACING = 0.75


@dataclass
class Keypoints:
    """One image's keypoints, float32, in file order."""

    xyz: np.ndarray  # [N,3]
    scale: np.ndarray  # [N]
    lap: np.ndarray  # [N]
    response: np.ndarray  # [N]
    desc: np.ndarray  # [N,48]

    @property
    def n(self) -> int:
        return int(self.scale.shape[0])

    def records(self) -> np.ndarray:
        """[N,54] float32 rows exactly as the .bin writer lays them out."""
        return np.concatenate(
            [self.xyz, self.scale[:, None], self.lap[:, None], self.response[:, None], self.desc], axis=1
        ).astype(np.float32)


def _r6(a: np.ndarray) -> np.ndarray:
    return np.round(a.astype(np.float64), 6).astype(np.float32)


def _surf_pattern(g: np.ndarray) -> np.ndarray:
    """Impose the (dx,dy,dz,|dx|,|dy|,|dz|) sign pattern and L2-normalise rows."""
    d = g.reshape(g.shape[0], 8, 6).copy()
    d[:, :, 3:] = np.abs(d[:, :, 3:])
    d = d.reshape(g.shape[0], D)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    return d


def _scales(rng: np.random.Generator, n: int) -> np.ndarray:
    w = 8.0 ** -np.arange(4)
    octave = rng.choice(4, size=n, p=w / w.sum())
    filt = octave * 2 + rng.integers(0, 2, size=n)
    m, step = _FILTERS[filt, 0], _FILTERS[filt, 1]
    u = rng.uniform(-1.0, 1.0, size=n)
    return _SPACING * 0.1333 * (m + u * step)


def _geometry(rng: np.random.Generator, n: int):
    xyz = rng.uniform(0.0, 1.0, size=(n, 3)) * np.array([300.0, 300.0, 1500.0])
    response = np.sort(rng.uniform(0.0, 1.0e4, size=n))[::-1]
    return xyz, response


def make_iid(n: int, image: int) -> Keypoints:
    """`iid` set (throughput): independent random unit descriptors, seed = 1000 + image."""
    rng = np.random.default_rng(1000 + image)
    desc = _surf_pattern(rng.standard_normal((n, D)))
    xyz, response = _geometry(rng, n)
    lap = rng.integers(0, 2, size=n).astype(np.float64)
    return Keypoints(_r6(xyz), _r6(_scales(rng, n)), lap.astype(np.float32), _r6(response), _r6(desc))


_BANK_CACHE: dict = {}


def _bank(n: int):
    if n not in _BANK_CACHE:
        rng = np.random.default_rng(7)
        proto = _surf_pattern(rng.standard_normal((2 * n, D)))
        _BANK_CACHE[n] = (proto, _scales(rng, 2 * n), rng.integers(0, 2, size=2 * n).astype(np.float64))
    return _BANK_CACHE[n]


def make_bank(n: int, image: int, noise: float = 0.02, outliers: float = 0.10) -> Keypoints:
    """`bank` set (parity): images share 2n prototypes, so true matches exist at -d 0.22."""
    proto, pscale, plap = _bank(n)
    rng = np.random.default_rng(2000 + image)
    pick = rng.permutation(2 * n)[:n]
    desc = _surf_pattern(proto[pick] + noise * rng.standard_normal((n, D)))
    scale = pscale[pick] * (1.0 + rng.uniform(-0.03, 0.03, size=n))
    lap = plap[pick].copy()
    out = rng.uniform(size=n) < outliers
    if out.any():
        k = int(out.sum())
        desc[out] = _surf_pattern(rng.standard_normal((k, D)))
        scale[out] = _scales(rng, k)
        lap[out] = rng.integers(0, 2, size=k)
    xyz, response = _geometry(rng, n)
    return Keypoints(_r6(xyz), _r6(scale), lap.astype(np.float32), _r6(response), _r6(desc))


def make_dense(n: int, image: int) -> Keypoints:
    """`dense` set (kernel study): iid descriptors, ONE laplacian class and ONE scale, so no column is gated out
    (match.cpp:270-275 pass for every pair) and the scoring kernel executes every algorithmic pair."""
    kp = make_iid(n, image)
    return Keypoints(kp.xyz, np.full(n, 2.0, np.float32), np.zeros(n, np.float32), kp.response, kp.desc)


def make(kind: str, n: int, image: int) -> Keypoints:
    if kind == "iid":
        return make_iid(n, image)
    if kind == "dense":
        return make_dense(n, image)
    if kind == "bank":
        return make_bank(n, image)
    raise ValueError(f"unknown synthetic set {kind!r}")


# ----------------------------------------------------------------------------------------------
# writers (formats of vtk3DSURF.cxx:402-525)


def _text_lines(kp: Keypoints) -> bytes:
    rec = kp.records().astype(np.float64)
    lines = []
    for r in rec:
        head = "%f,%f,%f,%f,%d,%f," % (r[0], r[1], r[2], r[3], int(r[4]), r[5])
        lines.append(head + ",".join("%f" % v for v in r[6:]))
    return ("\n".join(lines) + "\n").encode()


def write_bin(kp: Keypoints, path: str) -> None:
    kp.records().astype("<f4").tofile(path)


def write_csv(kp: Keypoints, path: str) -> None:
    with open(path, "wb") as f:
        f.write(_text_lines(kp))


def write_csv_gz(kp: Keypoints, path: str) -> None:
    with gzip.GzipFile(path, "wb", compresslevel=1, mtime=0) as f:
        f.write(_text_lines(kp))


WRITERS = {"bin": write_bin, "csv": write_csv, "csv.gz": write_csv_gz}


def _write_fast(kp: Keypoints, path: str, fmt: str) -> None:
    """Same bytes (.bin) / same text (.csv, .csv.gz) as WRITERS, through libfmio's C writer."""
    if fmt == "bin":
        write_bin(kp, path)
        return
    from . import hostio
    rec = kp.records()
    hostio.write_csv(path, rec[:, :6], rec[:, 6:], gz_level=1 if fmt == "csv.gz" else -1)


def write_group(dirpath: str, kind: str, n_images: int, n_points: int, fmt: str = "bin",
                rigid: bool = False, first_image: int = 0, threads: int = 1, keypoints=None) -> str:
    """Write `n_images` keypoint files plus the list file `bin/match` takes; returns the list path.

    List lines are absolute paths (as run.sh:86-88 writes them), optionally followed by a rigid
    translation `,x,y,z` (match.cpp:478-490).  threads > 1 (or ready-made `keypoints`) writes through
    libfmio's C text writer from a thread pool -- big groups for the bench.
    """
    os.makedirs(dirpath, exist_ok=True)
    paths = [os.path.abspath(os.path.join(dirpath, f"points{i}.{fmt}")) for i in range(n_images)]
    if threads > 1 or keypoints is not None:
        from concurrent.futures import ThreadPoolExecutor

        def one(i):
            kp = keypoints[i] if keypoints is not None else make(kind, n_points, first_image + i)
            _write_fast(kp, paths[i], fmt)

        with ThreadPoolExecutor(max(1, threads)) as ex:
            list(ex.map(one, range(n_images)))
    else:
        for i in range(n_images):
            WRITERS[fmt](make(kind, n_points, first_image + i), paths[i])
    lines = []
    for i, p in enumerate(paths):
        if rigid:
            lines.append(f"{p},{0.5 * i:.3f},{-0.25 * i:.3f},{1.0 * i:.3f}")
        else:
            lines.append(p)
    lst = os.path.join(dirpath, "points.txt")
    with open(lst, "w") as f:
        f.write("\n".join(lines) + "\n")
    return lst


def make_volume(shape, seed: int, blobs_per_mvox: float = 900.0, dtype=np.int16, noise: float = 12.0, texture: float = 0.0) -> np.ndarray:
    """Synthetic CT-like volume [z, y, x] for the SURF3D producer: a smooth background, Gaussian
    blobs of both signs over the detector's scale range (sigma 1.2 .. 14 voxels, small ones most frequent) and
    a little noise, quantised to integers in roughly [-1000, 2500] like Hounsfield units."""
    rng = np.random.default_rng(4000 + seed)
    nz, ny, nx = shape
    vol = np.zeros(shape, np.float32)
    z, y, x = np.meshgrid(np.linspace(0, 1, nz, dtype=np.float32), np.linspace(0, 1, ny, dtype=np.float32),
                          np.linspace(0, 1, nx, dtype=np.float32), indexing="ij")
    vol += 200.0 * np.sin(3.1 * x + 0.3) * np.cos(2.3 * y) + 150.0 * z
    n_blobs = max(8, int(blobs_per_mvox * nz * ny * nx / 1e6))
    sig = 1.2 * (14.0 / 1.2) ** (rng.random(n_blobs) ** 2.2)
    cz, cy, cx = rng.random(n_blobs) * nz, rng.random(n_blobs) * ny, rng.random(n_blobs) * nx
    amp = rng.choice([-1.0, 1.0], n_blobs) * rng.uniform(300.0, 1500.0, n_blobs)
    for s, a, pz, py, px in zip(sig, amp, cz, cy, cx):
        r = int(np.ceil(3.5 * s))
        z0, z1 = max(0, int(pz) - r), min(nz, int(pz) + r + 1)
        y0, y1 = max(0, int(py) - r), min(ny, int(py) + r + 1)
        x0, x1 = max(0, int(px) - r), min(nx, int(px) + r + 1)
        if z0 >= z1 or y0 >= y1 or x0 >= x1:
            continue
        gz = np.exp(-0.5 * ((np.arange(z0, z1) - pz) / s) ** 2).astype(np.float32)
        gy = np.exp(-0.5 * ((np.arange(y0, y1) - py) / s) ** 2).astype(np.float32)
        gx = np.exp(-0.5 * ((np.arange(x0, x1) - px) / s) ** 2).astype(np.float32)
        vol[z0:z1, y0:y1, x0:x1] += a * gz[:, None, None] * gy[None, :, None] * gx[None, None, :]
    vol += rng.normal(0.0, noise, shape).astype(np.float32)
    if texture > 0:  # fine-grained tissue-like texture: white noise smoothed to a ~2 voxel correlation length
        from scipy.ndimage import gaussian_filter
        t = gaussian_filter(rng.normal(0.0, 1.0, shape).astype(np.float32), 1.4)
        vol += np.float32(texture) * t / t.std()
    return np.clip(np.rint(vol), -1000, 2500).astype(dtype)
