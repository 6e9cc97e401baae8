"""GPU: libfrogmatch (through the C ABI) against the oracle on seeded inputs -- bit-exact."""
import os

import numpy as np
import pytest

import helpers
from frog_b200 import capi, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(built):
    m = capi.Matcher(0)
    yield m, O.PortLib()
    m.close()


def run_both(m, port, images, pf, ps, thr, rat, sym=False, engines=(False, True)):
    want = port.match_pairs(images, pf, ps, thr, rat, sym)
    m.clear()
    for i, (d, s, l) in enumerate(images):
        m.upload(i, d, s, l)
    stats = {}
    for force_exact in engines:
        res = m.match(pf, ps, thr, rat, sym=sym, force_exact=force_exact)
        got = res.all_pairs()
        stats[force_exact] = m.stats()
        res.free()
        assert len(got) == len(want)
        for p, (g, w) in enumerate(zip(got, want)):
            assert np.array_equal(g, w), f"pair {p} engine={'exact' if force_exact else 'tensor'}: {len(g)} vs {len(w)}"
    return stats


PARAMS = [(0.22, 1.0), (1.0, 1.0), (1.0, 0.8), (1e10, 1.0), (0.5, 0.998), (1.0, 1.01)]


@pytest.mark.parametrize("kind", ["bank", "iid"])
@pytest.mark.parametrize("thr,rat", PARAMS)
def test_group_parity(env, kind, thr, rat):
    m, port = env
    images = helpers.random_group(kind, 4, 1100)
    images[2] = tuple(x[:777] for x in images[2])  # ragged sizes, not multiples of any tile
    sched = helpers.pair_schedule(4, -1)
    st = run_both(m, port, images, [s[0] for s in sched], [s[1] for s in sched], thr, rat)
    assert st[False]["score_launches"] >= 1 and st[False]["scored_pairs"] > 0  # the tensor path really ran
    assert st[False]["rows_exact"] < 0.02 * st[False]["rows"]


def test_sym_and_repeated_pairs(env):
    m, port = env
    images = helpers.random_group("bank", 3, 900)
    run_both(m, port, images, [0, 0, 2, 1, 1], [1, 1, 0, 2, 1], 1.0, 0.9, sym=True)


def test_sizes_edge(env):
    m, port = env
    base = helpers.random_group("bank", 1, 600)[0]
    images = [base, tuple(x[:1] for x in base), tuple(x[:0] for x in base), tuple(x[:256] for x in base),
              tuple(x[:257] for x in base), tuple(x[:128] for x in base)]
    sched = [(i, j) for i in range(6) for j in range(6) if i != j]
    run_both(m, port, images, [s[0] for s in sched], [s[1] for s in sched], 1.0, 1.0)


def test_known_answers_on_gpu(env):
    m, port = env
    e = np.eye(48, dtype=np.float32)
    o = np.ones
    row = (0.6 * e[:1], o(1, np.float32), np.zeros(1, np.float32))
    cases = [
        (e[:1] * 0.5, o(1), np.zeros(1)),
        (e[:1] * 0.5, o(1), o(1)),
        (e[:1] * 0.5, o(1) * 1.31, np.zeros(1)),
        (e[:1] * 0.5, o(1) * np.float32(1.3), np.zeros(1)),
        (e[:1] * 0.5, o(1) * np.nextafter(np.float32(1.3), np.float32(2)), np.zeros(1)),
        (np.stack([e[1] * 0.5, e[2] * 0.5]), o(2), np.zeros(2)),
        (np.stack([row[0][0], row[0][0]]), o(2), np.zeros(2)),
        (np.stack([row[0][0], e[3], row[0][0] * 0.999]), o(3), np.zeros(3)),
    ]
    images = [row] + [tuple(np.asarray(x, np.float32) for x in c) for c in cases]
    pf = list(range(1, len(images)))
    ps = [0] * len(pf)
    for thr, rat in [(1.0, 1.0), (1.0, 1.01), (1e10, 1.0), (0.1, 1.0), (1.0, 0.5)]:
        run_both(m, port, images, pf, ps, thr, rat)
        run_both(m, port, images, ps, pf, thr, rat)


def test_many_duplicates_overflow_to_exact_rows(env):
    """Rows with more than 4 near-tied candidates cannot be certified from the top-4 list: they
    must be redone by the exact row kernel and still come out identical."""
    m, port = env
    a, b = helpers.random_group("iid", 2, 800)
    ad = a[0].copy()
    ad[100:140] = ad[100]  # 40 identical columns
    bd = b[0].copy()
    bd[:50] = ad[100] * np.float32(0.999)
    images = [(ad, np.full(800, 1.5, np.float32), np.zeros(800, np.float32)),
              (bd, np.full(800, 1.6, np.float32), np.zeros(800, np.float32))]
    for thr, rat in [(1.0, 1.0), (1.0, 1.01)]:
        st = run_both(m, port, images, [0], [1], thr, rat)
        assert st[False]["rows_exact"] >= 50


def test_uncertified_images_route_to_exact(env):
    m, port = env
    base = helpers.random_group("bank", 4, 500)
    nan_img = (base[1][0].copy(), base[1][1].copy(), base[1][2].copy())
    nan_img[0][7, 3] = np.nan
    neg_scale = (base[2][0], base[2][1].copy(), base[2][2])
    neg_scale[1][11] = -1.0
    big = (base[3][0] * np.float32(9.0), base[3][1], base[3][2])
    laps = (base[0][0], base[0][1], (np.arange(500) % 11).astype(np.float32))
    phantom = (np.concatenate([base[0][0], np.zeros((1, 48), np.float32)]),
               np.concatenate([base[0][1], [np.float32(321.5)]]).astype(np.float32),
               np.concatenate([base[0][2], [np.float32(321.5)]]).astype(np.float32))
    images = [base[0], nan_img, neg_scale, big, laps, phantom]
    sched = helpers.pair_schedule(6, -1)
    st = run_both(m, port, images, [s[0] for s in sched], [s[1] for s in sched], 1.0, 1.0)
    assert st[False]["rows_exact"] > 0


def test_other_descriptor_lengths(env):
    """d != 48 (surf3d descriptor types 1/2: 24 r^3 or 8 r^3 values, vtkOpenSURF3D/surf3d.cxx:36-39 -- 1000 and 3000
    at the default radius 5) runs on the K-chunked exact kernel: lengths below, equal to, not a multiple of and far
    above the 64-value chunk."""
    m, port = env
    rng = np.random.default_rng(5)
    for d in (1, 24, 64, 72, 192, 1000, 3000):
        images = []
        for i in range(3):
            x = rng.standard_normal((300 + 7 * i, d)).astype(np.float32)
            x /= np.linalg.norm(x, axis=1, keepdims=True)
            images.append((x, rng.uniform(1, 2, x.shape[0]).astype(np.float32), rng.integers(0, 2, x.shape[0]).astype(np.float32)))
        run_both(m, port, images, [0, 0, 1], [1, 2, 2], 1.2, 0.95, sym=(d == 72), engines=(False,))
        want = port.match_pairs_all(images, [0, 1], [1, 2], 1.3, d == 1000)
        res = m.match([0, 1], [1, 2], 1.3, 1.0, sym=(d == 1000), match_all=True)  # images are still resident
        assert res.total == sum(len(w) for w in want)
        for g, w in zip(res.all_pairs(), want):
            assert np.array_equal(g, w)
        res.free()


def test_api_errors(env):
    m, _ = env
    m.clear()
    d, s, l = helpers.random_group("iid", 1, 10)[0]
    m.upload(0, d, s, l)
    with pytest.raises(capi.FrogMatchError):
        m.match([0], [3])  # image 3 was never uploaded
    with pytest.raises(capi.FrogMatchError):
        m.match([0], [0], dist=3e19)  # would need the reference's stale-index leak (match.cpp:320-321)
    with pytest.raises(capi.FrogMatchError):
        m.upload(1, d[:, :24], s, l)  # mixed descriptor lengths


def test_full_size_engines_agree(env):
    """BASELINE config 2 scale (10 x 20k, iid): tensor-core path == exact CUDA path on every pair,
    and sampled pairs == oracle."""
    m, port = env
    n_img, n_pts = 10, 20000
    kps = [synth.make("iid", n_pts, i) for i in range(n_img)]
    m.clear()
    for i, k in enumerate(kps):
        m.upload(i, k.desc, k.scale, k.lap)
    sched = helpers.pair_schedule(n_img, -1)
    pf, ps = [s[0] for s in sched], [s[1] for s in sched]
    for thr, rat in [(1.0, 1.0), (1.0, 0.8)]:
        fast = m.match(pf, ps, thr, rat)
        lf = fast.all_pairs()
        st = m.stats()
        fast.free()
        ex = m.match(pf[:6], ps[:6], thr, rat, force_exact=True)
        le = ex.all_pairs()
        ex.free()
        for p in range(6):
            assert np.array_equal(lf[p], le[p])
        assert st["descriptor_pairs"] == 45 * n_pts * n_pts
        assert st["rows_exact"] < 0.02 * st["rows"]
        # size-independent properties: one match per row at most, sorted by second index, ids in range
        for l in lf:
            assert np.all(np.diff(l[:, 1].astype(np.int64)) > 0) and l[:, 0].max(initial=0) < n_pts
    img = [(k.desc, k.scale, k.lap) for k in kps[:2]]
    sub = (img[1][0][:1500], img[1][1][:1500], img[1][2][:1500])
    want = port.compute_matches(img[0], sub, 1.0, 0.8)
    got = lf[0][lf[0][:, 1] < 1500]
    assert np.array_equal(got, want)


def test_async_results_in_flight(env):
    """FM_FLAG_ASYNC: several fm_match calls queued back to back on one context, completed out of
    order with fm_result_wait(); every list equals the oracle's (and the synchronous call's)."""
    m, port = env
    images = helpers.random_group("bank", 4, 1500)
    sched = helpers.pair_schedule(4, -1)
    pf, ps = [s[0] for s in sched], [s[1] for s in sched]
    m.clear()
    for i, (d, s, l) in enumerate(images):
        m.upload(i, d, s, l)
    params = [(1.0, 1.0, False), (0.22, 1.0, False), (1.0, 0.8, True), (1.0, 1.0, False)]
    want = [port.match_pairs(images, pf, ps, t, r, False) for t, r, _ in params]
    inflight = [m.match(pf, ps, t, r, device_only=dev_only, asynchronous=True) for t, r, dev_only in params]
    assert all(r.counts is None for r in inflight)
    for k in (2, 0, 3, 1):  # completion order differs from submission order
        res = inflight[k].wait()
        if params[k][2]:
            res.fetch()
        assert res.total == sum(len(w) for w in want[k])
        st = res.stats()
        assert st["score_launches"] >= 1 and st["ms_score"] > 0
        for g, w in zip(res.all_pairs(), want[k]):
            assert np.array_equal(g, w)
    for r in inflight:
        r.free()
    # the context is still good for a synchronous call afterwards
    res = m.match(pf, ps, 1.0, 1.0)
    for g, w in zip(res.all_pairs(), want[0]):
        assert np.array_equal(g, w)
    res.free()


@pytest.mark.parametrize("thr,sym", [(0.22, False), (0.45, True), (1.0, False)])
def test_match_all_mode(env, thr, sym):
    """FM_FLAG_MATCH_ALL = the reference's -all (match.cpp:295-300), stale `match` and all: lists equal the
    oracle's, pair for pair, in emission order."""
    m, port = env
    images = helpers.random_group("bank", 4, 700)
    images[1] = tuple(x[:333] for x in images[1])
    images.append(tuple(x[:0] for x in images[0]))  # an empty image on either side
    sched = helpers.pair_schedule(5, -1)
    pf, ps = [s[0] for s in sched], [s[1] for s in sched]
    want = port.match_pairs_all(images, pf, ps, thr, sym)
    m.clear()
    for i, (d, s, l) in enumerate(images):
        m.upload(i, d, s, l)
    res = m.match(pf, ps, thr, 1.0, sym=sym, match_all=True)
    assert res.total == sum(len(w) for w in want) and res.total > 0
    for p, (g, w) in enumerate(zip(res.all_pairs(), want)):
        assert np.array_equal(g, w), f"pair {p}: {len(g)} vs {len(w)}"
    res.free()


def test_match_all_generic_descriptor_length(env):
    m, port = env
    rng = np.random.default_rng(5)
    images = []
    for i in range(3):
        n = 150 + 10 * i
        d = rng.normal(size=(n, 24)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        images.append((d, rng.uniform(1.0, 1.5, n).astype(np.float32), rng.integers(0, 2, n).astype(np.float32)))
    pf, ps = [0, 0, 1], [1, 2, 2]
    want = port.match_pairs_all(images, pf, ps, 1.25, True)
    m.clear()
    for i, (d, s, l) in enumerate(images):
        m.upload(i, d, s, l)
    res = m.match(pf, ps, 1.25, 1.0, sym=True, match_all=True)
    assert res.total == sum(len(w) for w in want) and res.total > 0
    for g, w in zip(res.all_pairs(), want):
        assert np.array_equal(g, w)
    res.free()


@pytest.mark.parametrize("force_exact", [False, True])
def test_distances_side_output(env, force_exact):
    """FM_FLAG_DISTANCES: the squared distance kept for every emitted match is the value the reference's own norm()
    (match.cpp:243-251) returns for that pair -- tolerance 0 ulp (the north star allows 2)."""
    m, port = env
    ref = O.RefLib() if os.path.exists(O.REF_LIB) else None
    images = helpers.random_group("bank", 3, 1300)
    pf, ps = [0, 0, 1], [1, 2, 2]
    m.clear()
    for i, (d, s, l) in enumerate(images):
        m.upload(i, d, s, l)
    for sym in (False, True):
        res = m.match(pf, ps, 1.0, 0.95, sym=sym, force_exact=force_exact, distances=True)
        want = port.match_pairs(images, pf, ps, 1.0, 0.95, sym)
        for p in range(3):
            got, dist = res.pairs(p), res.distances(p)
            assert np.array_equal(got, want[p]) and dist.shape[0] == got.shape[0] > 0
            a, b = images[pf[p]][0], images[ps[p]][0]
            # forward part: (first_idx, second_idx); the -sym part appended after it is (second-image row.. swapped roles)
            n_fwd = port.match_pairs(images, [pf[p]], [ps[p]], 1.0, 0.95, False)[0].shape[0]
            fi, si = got[:n_fwd, 0], got[:n_fwd, 1]
            exp = O.norm_matrix_numpy(b[si], a)[np.arange(n_fwd), fi] if n_fwd else np.zeros(0, np.float32)
            assert np.array_equal(dist[:n_fwd].view(np.uint32), exp.view(np.uint32))
            if ref is not None and n_fwd:
                assert np.array_equal(ref.distances(a, b, fi, si).view(np.uint32), dist[:n_fwd].view(np.uint32))
            if sym:  # reverse pass: rows of image `first` scan columns of image `second`; emitted as (row, match)
                ri, ci = got[n_fwd:, 0], got[n_fwd:, 1]
                exp = O.norm_matrix_numpy(a[ri], b)[np.arange(ri.shape[0]), ci]
                assert np.array_equal(dist[n_fwd:].view(np.uint32), exp.view(np.uint32))
        res.free()


def _group(kind, n_images, n_points):
    kps = [synth.make(kind, n_points, i) for i in range(n_images)]
    return [(k.desc, k.scale, k.lap) for k in kps]


def _fast_oracle_lists(images, pf, ps, thr, rat):
    fast = O.FastLib()
    return [fast.compute_matches(images[i], images[j], thr, rat) for i, j in zip(pf, ps)]


@pytest.mark.parametrize("kind,thr,rat", [("iid", 1.0, 0.8), ("bank", 1.0, 0.8), ("bank", 0.22, 0.9), ("bank", 0.5, 0.998),
                                          ("iid", 1.0, 0.5), ("bank", 1e10, 0.95)])
def test_two_phase_scoring_forced(built, kind, thr, rat):
    """Two-phase scoring (reject pass from the two best chunk maxima, capture pass for the warps with surviving rows)
    forced on: lists bit-identical to the oracle whether nearly every row is rejected (iid) or most survive (bank)."""
    images = _group(kind, 6, 12000)
    images[3] = tuple(x[:7777] for x in images[3])  # ragged
    pf = [i for i in range(6) for j in range(i + 1, 6)]
    ps = [j for i in range(6) for j in range(i + 1, 6)]
    want = _fast_oracle_lists(images, pf, ps, thr, rat)
    m = capi.Matcher(0)
    try:
        capi.debug_set_option("two_phase", 1)
        for i, (d, s, l) in enumerate(images):
            m.upload(i, d, s, l)
        res = m.match(pf, ps, thr, rat)
        got, st = res.all_pairs(), m.stats()
        res.free()
    finally:
        capi.debug_set_option("two_phase", -1)
        m.close()
    assert st["two_phase_batches"] >= 1 and st["rows_exact"] < 0.01 * st["rows"]
    for p, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g, w), f"pair {p}: {g.shape[0]} vs {w.shape[0]}"
    if kind == "iid":
        assert st["rows_rejected_early"] > 0.98 * st["rows"]
    else:
        assert sum(w.shape[0] for w in want) > 1000


def test_two_phase_scoring_survivor_overflow(built):
    """More survivors than the capture pass has units for: the surplus rows are matched by the exact row kernel."""
    images = _group("bank", 4, 13000)
    pf, ps = [0, 0, 0, 1, 1, 2], [1, 2, 3, 2, 3, 3]
    want = _fast_oracle_lists(images, pf, ps, 1.0, 0.9)
    m = capi.Matcher(0)
    try:
        capi.debug_set_option("two_phase", 1)
        capi.debug_set_option("surv_cap", 7)  # 7 units = 1792 rows of ~30 000 survivors
        for i, (d, s, l) in enumerate(images):
            m.upload(i, d, s, l)
        res = m.match(pf, ps, 1.0, 0.9)
        got, st = res.all_pairs(), m.stats()
        res.free()
    finally:
        capi.debug_set_option("two_phase", -1)
        capi.debug_set_option("surv_cap", -1)
        m.close()
    assert st["two_phase_batches"] >= 1 and st["rows_exact"] > 10000
    for p, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g, w), f"pair {p}: {g.shape[0]} vs {w.shape[0]}"


def test_two_phase_scoring_is_chosen_from_what_the_data_shows(built):
    """The library switches to two-phase scoring only after a call at -d2 < 1 rejected >= 99 % of its rows (random
    descriptors), never for data with real correspondences, never at -d2 >= 1; results identical throughout."""
    pf = [i for i in range(6) for j in range(i + 1, 6)]
    ps = [j for i in range(6) for j in range(i + 1, 6)]
    for kind, expect_switch in (("iid", True), ("bank", False)):
        images = _group(kind, 6, 12000)
        want = _fast_oracle_lists(images, pf, ps, 1.0, 0.8)
        m = capi.Matcher(0)
        try:
            for i, (d, s, l) in enumerate(images):
                m.upload(i, d, s, l)
            used = []
            for call in range(3):
                res = m.match(pf, ps, 1.0, 0.8)
                got, st = res.all_pairs(), m.stats()
                res.free()
                used.append(st["two_phase_batches"])
                for g, w in zip(got, want):
                    assert np.array_equal(g, w)
            assert used[0] == 0 and (used[1] >= 1 and used[2] >= 1) == expect_switch, used
            res = m.match(pf, ps, 1.0, 1.0)  # -d2 1: the ratio test cannot be decided from approximate scores
            assert m.stats()["two_phase_batches"] == 0
            res.free()
        finally:
            m.close()


@pytest.mark.parametrize("sym", [False, True])
def test_links_build_matches_read_pairs(env, sym):
    """Consumer hand-off: the adjacency built on the device from the match lists == ImageGroup::readPairs restated
    (registration/imageGroup.cxx:1386-1411) on the same blocks in file order -- every point's links in push_back order."""
    m, port = env
    images = helpers.random_group("bank", 4, 1500)
    images[1] = tuple(x[:611] for x in images[1])
    # a submission order that is NOT the file order, with a -targ style pair (3, 0) whose block sorts between others
    pairs = [(1, 2), (0, 1), (3, 0), (0, 3), (2, 3), (0, 2)]
    pf, ps = [p[0] for p in pairs], [p[1] for p in pairs]
    order = sorted(range(len(pairs)), key=lambda k: pairs[k])  # match.cpp:727-742: row-major over (first, second)
    m.clear()
    for i, (d, s, l) in enumerate(images):
        m.upload(i, d, s, l)
    res = m.match(pf, ps, 1.0, 0.95, sym=sym)
    lists = res.all_pairs()
    offsets, data, ms = res.links(pf, ps, block_order=order)
    res.free()
    n_points = {i: images[i][1].shape[0] for i in range(4)}
    blocks = [(pairs[k][0], pairs[k][1], lists[k]) for k in order]
    want_off, want = O.read_pairs_links_csr(blocks, n_points)
    assert data.shape[0] == 2 * sum(l.shape[0] for l in lists) > 1000
    assert np.array_equal(data, want)
    for img in range(4):
        assert np.array_equal(offsets[img].astype(np.uint64), want_off[img])
    # and against the literal push_back loops on a few points
    literal = O.read_pairs_links(blocks, n_points)
    for img, p in ((0, 0), (0, 7), (1, 610), (2, 100), (3, 1499)):
        got = [tuple(x) for x in data[int(offsets[img][p]): int(offsets[img][p + 1])].tolist()]
        assert got == literal[img][p]


def test_links_build_at_scale(built):
    """C2-size group (0.9 M matches, 1.8 M half-links): device link build == the stable-sort restatement."""
    kps = [synth.make("iid", 20000, i) for i in range(10)]
    pf = [i for i in range(10) for j in range(i + 1, 10)]
    ps = [j for i in range(10) for j in range(i + 1, 10)]
    m = capi.Matcher(0)
    try:
        for i, k in enumerate(kps):
            m.upload(i, k.desc, k.scale, k.lap)
        res = m.match(pf, ps, 1.0, 1.0)
        lists = res.all_pairs()
        offsets, data, ms = res.links(pf, ps)
        res.free()
    finally:
        m.close()
    want_off, want = O.read_pairs_links_csr([(i, j, l) for i, j, l in zip(pf, ps, lists)], {i: 20000 for i in range(10)})
    assert data.shape[0] > 1_500_000 and np.array_equal(data, want)
    for img in range(10):
        assert np.array_equal(offsets[img].astype(np.uint64), want_off[img])
    assert 0 < ms < 50
