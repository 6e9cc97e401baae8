"""GPU: each device stage of the tensor-core path against numpy, through the debug C ABI
(include/frogmatch_debug.h): sort + class table, FP16 operand packing, gate bands, the tcgen05
score tile, candidate capture."""
import numpy as np
import pytest

from frog_b200 import capi, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def matcher(built):
    m = capi.Matcher(0)
    yield m
    m.close()


def eps_bound(desc_a, desc_b):
    """task_eps of fm_score.cuh: Cauchy-Schwarz on the actual FP16 rounding residuals."""
    def stats(d):
        d = d.astype(np.float64)
        r = d.astype(np.float16).astype(np.float64) - d
        return np.sqrt((d ** 2).sum(1).max()), np.sqrt((r ** 2).sum(1).max())
    na, da = stats(desc_a)
    nb, db = stats(desc_b)
    return 1.01 * (na * db + da * nb + da * db) + 3e-5 * max(1.0, na * na, nb * nb)


def test_prep_sort_classes_operands(matcher):
    kp = synth.make("bank", 700, 3)
    kp.lap[:5] = -0.0  # -0.0 == 0.0 for the reference's float compare: must land in the 0 class
    matcher.clear()
    matcher.upload(0, kp.desc, kp.scale, kp.lap)
    d = matcher.debug_image(0, kp.n)
    assert d["flags"] == 0 and d["n_classes"] == 2
    perm = d["perm"]
    assert sorted(perm.tolist()) == list(range(kp.n))
    lap_s, scale_s = kp.lap[perm], kp.scale[perm]
    assert np.array_equal(d["scale_sorted"], scale_s)
    assert np.all(np.diff(lap_s) >= 0)
    for c in range(2):
        b, e = d["class_begin"][c], d["class_begin"][c + 1]
        assert np.all(lap_s[b:e] == d["class_lap"][c]) and np.all(np.diff(scale_s[b:e]) >= 0)
    assert d["class_begin"][2] == kp.n
    n2 = (kp.desc.astype(np.float64) ** 2).sum(1)
    assert abs(d["max_norm2"] - n2.max()) < 1e-5
    rowop, colop = d["rowop"], d["colop"]
    assert rowop.shape == (768, 64)
    assert np.array_equal(rowop[:kp.n, :48], kp.desc[perm].astype(np.float16))
    assert np.array_equal(colop[:kp.n, :48], rowop[:kp.n, :48])
    assert np.all(rowop[:kp.n, 48:50] == 1) and not rowop[:kp.n, 50:].any() and not colop[:kp.n, 50:].any()
    assert not rowop[kp.n:].any() and not colop[kp.n:].any()
    half_norm = colop[:kp.n, 48].astype(np.float64) + colop[:kp.n, 49].astype(np.float64)
    assert np.max(np.abs(half_norm + 0.5 * n2[perm])) < 2e-6


@pytest.mark.parametrize("kind,n_first,n_second", [("bank", 900, 600), ("iid", 1500, 257)])
def test_bands_score_tile_and_candidates(matcher, kind, n_first, n_second):
    a, b = synth.make(kind, n_first, 0), synth.make(kind, n_second, 1)
    matcher.clear()
    matcher.upload(0, a.desc, a.scale, a.lap)
    matcher.upload(1, b.desc, b.scale, b.lap)
    da, db = matcher.debug_image(0, a.n), matcher.debug_image(1, b.n)
    pa, pb = da["perm"], db["perm"]
    gate = O.gate_matrix_numpy(b.scale[pb], b.lap[pb], a.scale[pa], a.lap[pa])  # sorted rows x sorted cols
    exact = O.norm_matrix_numpy(b.desc[pb], a.desc[pa]).astype(np.float64)
    n2a = (a.desc[pa].astype(np.float64) ** 2).sum(1)
    n2b = (b.desc[pb].astype(np.float64) ** 2).sum(1)
    t_exact = 0.5 * (n2b[:, None] - exact)  # = a.b - |col|^2/2 up to the reference's own rounding
    eps = eps_bound(a.desc, b.desc)
    rowop = db["rowop"].astype(np.float64)
    colop = da["colop"].astype(np.float64)
    worst = 0.0
    for rb in range((n_second + 255) // 256):
        u = matcher.debug_score_unit(0, 1, rb, a.n, b.n)
        r0 = rb * 256
        nr = u["bands"].shape[0]
        for r in range(nr):
            lo, hi = u["bands"][r]
            cols = np.nonzero(gate[r0 + r])[0]
            if len(cols) == 0:
                assert lo == hi
            else:  # both gates select one contiguous interval of the (laplacian, scale) order
                assert lo == cols[0] and hi == cols[-1] + 1 and len(cols) == hi - lo
        t = u["t"][:nr, :a.n]
        t_fp16 = rowop[r0:r0 + nr] @ colop[:a.n].T  # what the MMA should produce, in float64
        visited = ~np.isnan(t)
        assert np.all(visited[gate[r0:r0 + nr]]), "a gated-in column was not scored"
        assert np.max(np.abs(t[visited] - t_fp16[visited])) < 2e-5  # FP32 accumulation in the tensor core
        err = np.abs(t[visited] - t_exact[r0:r0 + nr][visited])
        worst = max(worst, float(err.max()))
        assert err.max() <= eps, "certified error bound violated"
        # candidate capture: every column within 2*eps of the row's 2nd best gated score is listed
        tm = np.where(gate[r0:r0 + nr], t, -np.inf)
        for r in range(nr):
            order = np.sort(tm[r])[::-1]
            a2 = order[1] if len(order) > 1 else -np.inf
            need = set(np.nonzero(np.isfinite(tm[r]) & (tm[r] >= a2 - 2 * eps))[0].tolist())
            cols_r = u["cand_col"][r] & 0x7FFFFFFF
            got = {int(c) for c, v in zip(cols_r, u["cand_t"][r]) if np.isfinite(v)}
            overflow = np.isposinf(u["cand_t"][r][0])  # marker: the row's capture list overflowed
            truncated = np.isfinite(u["cand_t"][r][0]) and bool(u["cand_col"][r][0] >> 31)
            if truncated:  # the 8 best were kept: complete iff the smallest kept one is below the band
                kept = u["cand_t"][r][np.isfinite(u["cand_t"][r])]
                assert len(kept) == 8 and kept.min() >= np.sort(tm[r])[::-1][7]
                overflow = kept.min() >= a2 - 2 * eps
            if not overflow:
                assert need <= got
                assert len(got) <= 8
            for c, v in zip(cols_r, u["cand_t"][r]):
                if np.isfinite(v):
                    assert v == t[r, c] and gate[r0 + r, c]
    print(f"max |t_fp16 - t_exact| = {worst:.2e} (bound {eps:.2e})")
