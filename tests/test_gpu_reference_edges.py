"""The reference's own edge cases as a regression gate for the CUDA path
(/root/reference/tests/test_edge_cases.py:152-172, 214-253, 325-355,
tests/test_zncc_service.py:232-255; klt.py:46, 268-273), run against the UNMODIFIED
reference modules on the same inputs (oracle/_ref, placed by build(); the files
travel to the GPU box).  Also: the one-upload device cache of the drop-in classes,
outlier filtering on the fixed-kernel-size path, non-15 block sizes through the
default corner mode, and the exchange header written on the device."""
import logging

import numpy as np
import pandas as pd
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402
from oracle import refimport, vendor_ref  # noqa: E402

DEFAULT = dict(minDistance=10, blocksize=15, maxCorners=20000, matching_winsize=25,
               qualityLevel=0.1, xStart=0, tile_size=20000, laplacian_kernel_size=7,
               outliers_filtering=False, laplacian_invert_polarity=False)


@pytest.fixture(scope="module")
def ref_mods():
    """(klt, zncc_service, configuration) of the unmodified reference."""
    import os
    if not (vendor_ref.available() or os.path.isdir("/root/reference/karios/matcher")):
        pytest.skip("no reference tree (oracle/_ref missing: run __graft_entry__.build() where /root/reference exists)")
    logging.getLogger("karios").setLevel(logging.ERROR)
    return refimport.load()


def _texture(h, w, seed, dtype=np.uint16):
    from karios_b200 import synth
    ref_t, mon_t = synth.make_pair(h, w, seed=seed)
    to_np = lambda t: t.view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    ref, mon = to_np(ref_t), to_np(mon_t)
    if dtype == np.uint16:
        return ref.copy(), mon.copy()
    return ref.astype(dtype), mon.astype(dtype)


def _frames_equal(got, want, tol=1e-3):
    assert len(got) == len(want)
    for f, g in zip(got, want):
        assert len(f) == len(g)
        assert np.array_equal(f["x0"].to_numpy(), g["x0"].to_numpy())
        assert np.array_equal(f["y0"].to_numpy(), g["y0"].to_numpy())
        if len(f):
            assert np.abs(f["dx"].to_numpy() - g["dx"].to_numpy()).max() < tol
            assert np.abs(f["dy"].to_numpy() - g["dy"].to_numpy()).max() < tol
            assert np.abs(f["score"].to_numpy() - g["score"].to_numpy()).max() < 2e-2


def test_flat_image_like_reference(ref_mods):
    """test_edge_cases.py:152-172: a flat image has no feature; both return None."""
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.matcher.klt import klt_tracker
    rklt, _, rcfg = ref_mods
    ref_data = np.ones((20, 20), np.uint8) * 128
    mask = np.ones((20, 20), np.uint8)
    kw = dict(DEFAULT, minDistance=1, blocksize=2, maxCorners=10, matching_winsize=3, qualityLevel=0.001,
              tile_size=1000, laplacian_kernel_size=3)
    want = rklt.klt_tracker(ref_data, ref_data.copy(), mask, rcfg.KLTConfiguration(**kw))
    got = klt_tracker(ref_data, ref_data.copy(), mask, KLTConfiguration(**kw))
    assert want is None and got is None


def test_tile_size_one_like_reference(ref_mods):
    """test_edge_cases.py:214-253 without the mock: tile_size 1 on a 5 x 5 image is 25
    one-pixel tiles; every stage has to cope with a 1 x 1 raster."""
    from karios_b200 import _native as N
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.core.image import ArrayRaster
    from karios_b200.matcher.klt import KLT
    rklt, _, rcfg = ref_mods
    rng = np.random.default_rng(5)
    test_ref = rng.integers(0, 255, (5, 5), dtype=np.uint8)
    test_img = rng.integers(0, 255, (5, 5), dtype=np.uint8)
    kw = dict(DEFAULT, qualityLevel=0.01, tile_size=1, laplacian_kernel_size=3, outliers_filtering=True)
    want = list(rklt.KLT(rcfg.KLTConfiguration(**kw)).match(refimport.ArrayImage(test_img),
                                                            refimport.ArrayImage(test_ref), None))
    got = list(KLT(KLTConfiguration(**kw)).match(ArrayRaster(test_img), ArrayRaster(test_ref), None))
    _frames_equal(got, want)
    # the fused tile call on a 1 x 1 context
    ctx = N.Context(1, 1, 20000)
    try:
        rows = N.RowBuffers(20000, ctx.device)
        kc = N.make_conf(KLTConfiguration(**dict(kw, outliers_filtering=False)))
        for y in range(5):
            for x in range(5):
                st = ctx.match_tile(torch.from_numpy(test_img).cuda(), torch.from_numpy(test_ref).cuda(), None,
                                    (x, y, 1, 1), kc, rows)
                assert st.n_corners == 0 and st.n_kept == 0
                assert st.valid == int(test_img[y, x] != 0 and test_ref[y, x] != 0)
    finally:
        ctx.close()
    # 2 x 3 tiles (remainder tiles of a few pixels) through the reference and the drop-in
    kw2 = dict(DEFAULT, tile_size=3, laplacian_kernel_size=3, qualityLevel=0.01)
    want = list(rklt.KLT(rcfg.KLTConfiguration(**kw2)).match(refimport.ArrayImage(test_img),
                                                             refimport.ArrayImage(test_ref), None))
    got = list(KLT(KLTConfiguration(**kw2)).match(ArrayRaster(test_img), ArrayRaster(test_ref), None))
    _frames_equal(got, want)


def test_zncc_out_of_image_rows_through_dataframe(ref_mods):
    """test_edge_cases.py:325-355: rows outside the rasters give NaN through the real
    DataFrame path -- mixed with valid rows, a non-default index, and the all-outside frame
    on objects that have no pixel data at all (Mock(spec=GdalRasterImage) in the reference)."""
    from unittest.mock import Mock
    from karios_b200.core.image import ArrayRaster
    from karios_b200.matcher.zncc_service import ZNCCService
    _, rzs, _ = ref_mods
    ref, mon = _texture(300, 340, 11)
    rows = pd.DataFrame({
        "x0": np.float32([-100.0, 10000.0, 150.0, 27.0, 28.0, 311.0, 312.0, 200.0, 100.0]),
        "y0": np.float32([-100.0, 10000.0, 120.0, 100.0, 100.0, 100.0, 100.0, 271.0, 272.0]),
        "dx": np.float32([0.0, 0.0, 0.3, 0.0, 0.49, 0.4, 0.0, -0.2, 0.0]),
        "dy": np.float32([0.0, 0.0, -0.2, 0.0, 0.0, 0.0, 0.0, 0.51, 0.0]),
    }, index=[5, 9, 2, 40, 41, 42, 43, 44, 45])
    want = rzs.ZNCCService().compute_zncc(rows, refimport.ArrayImage(mon), refimport.ArrayImage(ref))
    got = ZNCCService().compute_zncc(rows, ArrayRaster(mon), ArrayRaster(ref))
    assert list(got.index) == list(want.index) and got.dtype == np.float64
    assert np.array_equal(np.isnan(got.to_numpy()), np.isnan(want.to_numpy()))
    assert np.isnan(got.to_numpy()[:2]).all() and not np.isnan(got.to_numpy()[2])
    assert np.nanmax(np.abs(got.to_numpy() - want.to_numpy())) < 1e-5
    # a float64 column in the frame makes the reference round x0 + dx in float64
    rows64 = rows.copy()
    rows64["radial error"] = np.float64(0.1)
    rows64.loc[2, "dx"] = np.float32(0.5)              # 150 + 0.5: half-to-even either way
    want = rzs.ZNCCService().compute_zncc(rows64, refimport.ArrayImage(mon), refimport.ArrayImage(ref))
    got = ZNCCService().compute_zncc(rows64, ArrayRaster(mon), ArrayRaster(ref))
    assert np.array_equal(np.isnan(got.to_numpy()), np.isnan(want.to_numpy()))
    assert np.nanmax(np.abs(got.to_numpy() - want.to_numpy())) < 1e-5
    # the reference's own case: nothing but sizes on the image objects
    invalid = pd.DataFrame({"x0": [-100.0, 10000.0], "y0": [-100.0, 10000.0], "dx": [0.0, 0.0], "dy": [0.0, 0.0]})
    m_mon, m_ref = Mock(), Mock()
    for m in (m_mon, m_ref):
        m.x_size = m.y_size = 100
        m.device_array = None
    res = ZNCCService().compute_zncc(invalid, m_mon, m_ref)
    assert pd.isna(res.iloc[0]) and pd.isna(res.iloc[1])
    m_mon.clear_cache.assert_called_once()
    m_ref.clear_cache.assert_called_once()
    # empty frame
    assert len(ZNCCService().compute_zncc(rows.iloc[:0], ArrayRaster(mon), ArrayRaster(ref))) == 0


@pytest.mark.parametrize("ksize", [3, 7])
def test_float_raster_with_nan_and_inf(ref_mods, ksize):
    """klt.py:46 (np.nanmin / np.nanmax in _to_uint8) and :268-273 (np.isfinite in the auto
    mask): float32 rasters holding NaN, +Inf-free and zero pixels."""
    from karios_b200 import _native as N
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.core.image import ArrayRaster
    from karios_b200.matcher.klt import KLT
    rklt, _, rcfg = ref_mods
    ref, mon = _texture(400, 520, 31, np.float32)
    rng = np.random.default_rng(7)
    for a in (ref, mon):
        a[rng.integers(0, 400, 300), rng.integers(0, 520, 300)] = np.nan
        a[rng.integers(0, 400, 50), rng.integers(0, 520, 50)] = 0.0
    mon[60:90, 100:140] = np.nan
    ref[200:210, 300:330] = np.nan
    # stage level: min / max ignore NaN, the mask drops non-finite and zero pixels
    ctx = N.Context(520, 400, 20000)
    try:
        mask = ctx.minmax_mask(torch.from_numpy(mon).cuda(), torch.from_numpy(ref).cuda(), want_mask=True)
        st = ctx.read_stats()
        assert (st.min_a, st.max_a) == (float(np.nanmin(mon)), float(np.nanmax(mon)))
        assert (st.min_b, st.max_b) == (float(np.nanmin(ref)), float(np.nanmax(ref)))
        want_mask, cnt = O.auto_mask(mon, ref)
        assert np.array_equal(mask.cpu().numpy(), want_mask) and st.valid == cnt
        with np.errstate(invalid="ignore"):
            u8 = rklt._to_uint8(mon)                      # the unmodified reference function
        lap = ctx.u8_laplacian(torch.from_numpy(mon).cuda(), ksize, slot=0)
        assert np.array_equal(lap.cpu().numpy(), O.laplacian(u8, ksize))
    finally:
        ctx.close()
    # whole path against the unmodified reference
    kw = dict(DEFAULT, maxCorners=600, laplacian_kernel_size=ksize)
    with np.errstate(invalid="ignore"):
        want = list(rklt.KLT(rcfg.KLTConfiguration(**kw)).match(refimport.ArrayImage(mon),
                                                                refimport.ArrayImage(ref), None))
    got = list(KLT(KLTConfiguration(**kw)).match(ArrayRaster(mon), ArrayRaster(ref), None))
    assert len(want) == 1 and len(want[0]) > 100
    _frames_equal(got, want)
    # a raster with +Inf: nanmax is Inf, every uint8 value collapses to 0 -> no feature
    mon_inf = mon.copy()
    mon_inf[10, 10] = np.inf
    with np.errstate(invalid="ignore"):
        want = list(rklt.KLT(rcfg.KLTConfiguration(**kw)).match(refimport.ArrayImage(mon_inf),
                                                                refimport.ArrayImage(ref), None))
    got = list(KLT(KLTConfiguration(**kw)).match(ArrayRaster(mon_inf), ArrayRaster(ref), None))
    _frames_equal(got, want)


@pytest.mark.parametrize("resident", [False, True])
def test_outliers_filtering_fixed_kernel_size(ref_mods, resident):
    """klt.py:52-71, 161-163 on the fixed-kernel-size path: the filter runs on OpenCV-ordered,
    tile-local rows before offsets and the (x0, y0) sort.  A block of the monitored image is
    displaced so that the filter has rows to drop; two tiles so that offsets matter."""
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.core.image import ArrayRaster, DeviceRaster
    from karios_b200.matcher.klt import KLT
    rklt, _, rcfg = ref_mods
    ref, mon = _texture(420, 900, 17)
    mon[100:220, 500:640] = np.roll(mon[100:220, 500:640], 2, axis=1)       # local 2 px outliers
    kw = dict(DEFAULT, maxCorners=900, tile_size=500, outliers_filtering=True)
    want = list(rklt.KLT(rcfg.KLTConfiguration(**kw)).match(refimport.ArrayImage(mon),
                                                            refimport.ArrayImage(ref), None))
    mk = (lambda a: DeviceRaster(torch.from_numpy(a).cuda())) if resident else ArrayRaster
    got = list(KLT(KLTConfiguration(**kw)).match(mk(mon), mk(ref), None))
    kw_off = dict(kw, outliers_filtering=False)
    unfiltered = list(rklt.KLT(rcfg.KLTConfiguration(**kw_off)).match(refimport.ArrayImage(mon),
                                                                      refimport.ArrayImage(ref), None))
    assert sum(len(f) for f in want) < sum(len(f) for f in unfiltered)      # the filter did drop rows
    _frames_equal(got, want)


def test_scene_matcher_rejects_what_it_does_not_do():
    from karios_b200 import _native as N
    from karios_b200.api import SceneMatcher
    from karios_b200.core.configuration import KLTConfiguration
    with pytest.raises(N.KariosB200Error):
        SceneMatcher(64, 64, KLTConfiguration(outliers_filtering=True))
    with pytest.raises(N.KariosB200Error):
        SceneMatcher(64, 64, KLTConfiguration(laplacian_kernel_size="auto"))


@pytest.mark.parametrize("block", [3, 7, 21])
def test_other_block_sizes_route_to_the_exact_kernel(golden, block):
    """The two-tier corner response is built for blockSize 15 with a bounded maxCorners; any
    other block size has to take the one-tier kernel (not a wrong bound) and still give
    OpenCV's corner list, order included."""
    from karios_b200 import _native as N
    g = golden("basic")
    lap, mask = g["lap_ref"], g["mask_box"]
    h, w = lap.shape
    ctx = N.Context(w, h, 20000)
    try:
        for mc, q, md in ((400, 0.1, 10.0), (20000, 0.01, 3.0)):
            pts = ctx.good_features(torch.from_numpy(lap).cuda(), torch.from_numpy(mask).cuda(), mc, q, md, block)
            st = ctx.read_stats()
            assert st.two_tier == 0, "blockSize != 15 must not use the two-tier bound"
            want = O.good_features(lap, mask, mc, q, md, block)
            want = np.zeros((0, 2), np.float32) if want is None else want.reshape(-1, 2)
            assert np.array_equal(pts.cpu().numpy(), want), (block, mc)
        # and block 15 does use it (so the assertion above is not vacuous)
        ctx.good_features(torch.from_numpy(lap).cuda(), torch.from_numpy(mask).cuda(), 400, 0.1, 10.0, 15)
        assert ctx.read_stats().two_tier == 1
    finally:
        ctx.close()


def test_one_upload_per_raster_through_the_dropin_api():
    """KLT.match (2 tiles) + compute_zncc + compute_mutual_info + compute_mi on host rasters:
    the reference reads each raster once per consumer (api/core.py:845-907); here each raster
    is uploaded exactly once and shared."""
    from karios_b200.core import image as kimg
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.core.image import ArrayRaster
    from karios_b200.matcher.klt import KLT
    from karios_b200.matcher.mutual_info_service import MutualInfoService
    from karios_b200.matcher.zncc_service import ZNCCService
    ref, mon = _texture(400, 700, 23)
    mask = np.ones((400, 700), np.uint8)
    mask[:, :30] = 0
    mon_img, ref_img, mask_img = ArrayRaster(mon), ArrayRaster(ref), ArrayRaster(mask)
    before = dict(kimg.uploads)
    frames = list(KLT(KLTConfiguration(maxCorners=500, tile_size=400)).match(mon_img, ref_img, mask_img))
    assert len(frames) == 2
    df = pd.concat(frames)
    z = ZNCCService().compute_zncc(df, mon_img, ref_img)
    mi = MutualInfoService().compute_mutual_info(df, mon_img, ref_img)
    nmi = ZNCCService().compute_mi(df, mon_img, ref_img)
    assert len(z) == len(mi) == len(nmi) == len(df) > 100
    assert kimg.uploads["count"] - before["count"] == 3                      # mon, ref, mask: once each
    assert kimg.uploads["bytes"] - before["bytes"] == mon.nbytes + ref.nbytes + mask.nbytes
    assert (df["x0"] >= 30).all()
    # a new raster object is a new raster; releasing drops the copy
    frames2 = list(KLT(KLTConfiguration(maxCorners=500, tile_size=400)).match(ArrayRaster(mon), ref_img, mask_img))
    assert kimg.uploads["count"] - before["count"] == 4
    _frames_equal(frames2, frames, tol=1e-12)
    kimg.release_device(ref_img)
    list(KLT(KLTConfiguration(maxCorners=500)).match(mon_img, ref_img, None))
    assert kimg.uploads["count"] - before["count"] == 5


def test_exchange_header_written_on_the_device():
    """kr_unit_header_write (count + dx / dy moments of a unit, stream-ordered after
    kr_match_tile) against the torch restatement of the record (sharding.write_header), and
    the rows of match_many living inside the exchange arena."""
    from karios_b200 import sharding, synth
    from karios_b200.api import SceneMatcher
    from karios_b200.core.configuration import KLTConfiguration
    pairs = []
    for seed in (3, 4, 5):
        ref_t, mon_t = synth.make_pair(420, 610, seed=seed, device="cuda")
        pairs.append((mon_t, ref_t))
    sm = SceneMatcher(420, 610, KLTConfiguration(maxCorners=403, tile_size=350), 0.4, depth=2)
    try:
        res, total = sm.match_many(pairs)
        torch.cuda.synchronize()
        arena, cap = sm.last_arena, sm.rows.capacity
        assert arena.shape == (12, sharding.unit_words(cap)) and cap == 403
        counts, flags, mom = sharding.headers(arena[None])
        assert counts[0].cpu().tolist() == sm.last_counts and int(counts.sum()) == total
        assert int(flags.sum()) == 0
        want = arena.clone()
        for u, n in enumerate(sm.last_counts):
            sharding.write_header(want, u, cap, n)
        _, _, mom_want = sharding.headers(want[None])
        assert torch.equal(mom[..., 0], mom_want[..., 0])                      # n
        assert torch.equal(mom[..., 5:], mom_want[..., 5:])                    # min / max: exact
        assert torch.allclose(mom[..., 1:5], mom_want[..., 1:5], rtol=1e-12, atol=1e-12)
        # the per-tile result views alias the arena records
        u = 0
        for tiles in res:
            for f, z in tiles:
                _, rows, zz = sharding.unit_views(arena, u, cap)
                assert f.data_ptr() == rows.data_ptr() and z.data_ptr() == zz.data_ptr()
                u += 1
        m = sharding.moments_dict(sharding.batch_moments(arena[None]))
        alldx = torch.cat([f[2] for tiles in res for f, _ in tiles]).double()
        assert m["n"] == total and abs(m["mean_dx"] - float(alldx.mean())) < 1e-9
    finally:
        sm.close()


def test_running_cut_of_the_corner_response():
    """Tier 1 of the corner response drops whole row pieces once its running estimate of the
    selection cut-off is known (kr_corner_fast.cu).  Which rows are dropped depends on timing;
    the corners must not: default mode against OpenCV's arithmetic at every pixel (corner mode 1)
    on mid-size images of different character -- even texture, texture next to a flat half, few
    strong corners in noise, a masked scene -- several times each."""
    from karios_b200 import _native as N
    from karios_b200 import synth
    rng = np.random.default_rng(3)
    ref_t, _ = synth.make_pair(2100, 3100, seed=8)
    tex = O.laplacian(O.to_uint8(ref_t.view(torch.int16).numpy().view(np.uint16)), 7)
    half = tex.copy()
    half[:, 1500:] = 0
    sparse = rng.integers(0, 12, tex.shape).astype(np.uint8)
    for _ in range(300):
        y, x = int(rng.integers(20, 2080)), int(rng.integers(20, 3080))
        sparse[y:y + 9, x:x + 9] = 255
    mask = (rng.random(tex.shape) > 0.4).astype(np.uint8)
    mask[:, :700] = 0
    cases = [("texture", tex, None, 3000), ("half flat", half, None, 3000), ("sparse", sparse, None, 200),
             ("masked", tex, mask, 1500), ("texture, many corners", tex, None, 20000)]
    ctx = N.Context(3100, 2100, 20000)
    used = 0
    try:
        for name, img, m, mc in cases:
            d_img = torch.from_numpy(img).cuda()
            d_m = None if m is None else torch.from_numpy(m).cuda()
            ctx.set_corner_mode(1)
            want = ctx.good_features(d_img, d_m, mc, 0.1, 10, 15).cpu().numpy()
            ctx.set_corner_mode(0)
            for rep in range(3):
                got = ctx.good_features(d_img, d_m, mc, 0.1, 10, 15).cpu().numpy()
                st = ctx.read_stats()
                assert np.array_equal(got, want), (name, rep, got.shape, want.shape)
                print(f"{name}: corners {len(got)}, estimate bits {st.est_cut_bits:#x}, row pieces dropped "
                      f"{st.rows_skipped}, candidates {st.n_candidates}, two-tier {st.two_tier}, "
                      f"fallback {st.two_tier_fallback}")
                used += int(st.rows_skipped > 0)
                if name in ("texture", "texture, many corners"):
                    assert st.two_tier_fallback == 0, "an even texture must not need the exact re-run"
    finally:
        ctx.close()
    assert used > 0, "the running cut never engaged"


def test_zncc_served_from_the_matching_pass():
    """KLT.match leaves the ZNCC of the rows it yields (computed in the same launch sequence);
    compute_zncc serves them only for bit-identical rows of that tile and computes otherwise --
    the values are the same either way."""
    from karios_b200.core import image as kimg
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.core.image import ArrayRaster
    from karios_b200.matcher.klt import KLT
    from karios_b200.matcher.zncc_service import ZNCCService
    ref, mon = _texture(420, 700, 29)
    mon_img, ref_img = ArrayRaster(mon), ArrayRaster(ref)
    zs = ZNCCService()
    for df in KLT(KLTConfiguration(maxCorners=500, tile_size=400)).match(mon_img, ref_img, None):
        cand = df[df["score"] >= 0.4]
        assert kimg.recall_scores(mon_img, ref_img, cand) is not None          # known from the matching pass
        served = zs.compute_zncc(cand, mon_img, ref_img)
        fresh = zs.compute_zncc(cand, ArrayRaster(mon), ArrayRaster(ref))       # other raster objects: computed
        assert list(served.index) == list(cand.index)
        assert np.array_equal(np.isnan(served.to_numpy()), np.isnan(fresh.to_numpy()))
        assert np.array_equal(np.nan_to_num(served.to_numpy()), np.nan_to_num(fresh.to_numpy()))
        # rows that are not rows of the tile are computed, not served
        moved = cand.copy()
        moved["dx"] = moved["dx"] + np.float32(1.0)
        assert kimg.recall_scores(mon_img, ref_img, moved) is None
        renum = cand.reset_index(drop=True)
        if len(renum) != len(df):
            assert kimg.recall_scores(mon_img, ref_img, renum) is None
        z2 = zs.compute_zncc(renum, mon_img, ref_img)
        assert np.array_equal(np.nan_to_num(z2.to_numpy()), np.nan_to_num(served.to_numpy()))


def test_prefetch_gives_the_same_rows():
    """core.image.prefetch starts the upload on a copy stream; KLT.match / compute_zncc order
    themselves after it and give what they give without it (one upload per raster either way)."""
    from karios_b200.core import image as kimg
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.core.image import ArrayRaster
    from karios_b200.matcher.klt import KLT
    from karios_b200.matcher.zncc_service import ZNCCService
    pairs = [_texture(600, 900, s) for s in (41, 42, 43)]
    conf = KLTConfiguration(maxCorners=700, tile_size=500)

    def run(prefetch):
        out = []
        imgs = [(ArrayRaster(torch.from_numpy(m).pin_memory().numpy()), ArrayRaster(torch.from_numpy(r).pin_memory().numpy()))
                for r, m in pairs]
        before = kimg.uploads["count"]
        if prefetch:
            for im in imgs[0]:
                kimg.prefetch(im)
        for i, (mon_img, ref_img) in enumerate(imgs):
            if prefetch and i + 1 < len(imgs):
                for im in imgs[i + 1]:
                    kimg.prefetch(im)
            frames = list(KLT(conf).match(mon_img, ref_img, None))
            df = pd.concat(frames)
            z = ZNCCService().compute_zncc(df, mon_img, ref_img)
            out.append((df, z))
        assert kimg.uploads["count"] - before == 2 * len(pairs)
        return out

    a, b = run(False), run(True)
    for (da, za), (db, zb) in zip(a, b):
        assert len(da) == len(db) > 300
        for c in ("x0", "y0", "dx", "dy", "score"):
            assert np.array_equal(da[c].to_numpy(), db[c].to_numpy())
        assert np.array_equal(np.nan_to_num(za.to_numpy()), np.nan_to_num(zb.to_numpy()))


def test_parallel_upload_of_pageable_rasters():
    """kr_upload_pageable (host threads staging chunks through pinned buffers) delivers the bytes
    torch's own copy delivers: sizes around the chunk size, a raster-sized array, repeated use."""
    from karios_b200 import _native as N
    dev = torch.device("cuda", torch.cuda.current_device())
    rng = np.random.default_rng(1)
    for shape, dt in (((2048, 2049), np.uint16), ((4097, 1025), np.uint16), ((3000, 3001), np.float32),
                      ((9000, 9011), np.uint16), ((4096, 2048), np.uint8)):
        a = rng.integers(0, 60000, shape).astype(dt)
        assert a.nbytes >= (8 << 20)
        for _ in range(2):
            got = N.to_device(a, dev)
            torch.cuda.synchronize()
            want = torch.from_numpy(a.view(np.int16) if dt == np.uint16 else a).cuda()
            if dt == np.uint16:
                got = got.view(torch.int16)
            assert got.shape == want.shape and torch.equal(got, want), (shape, dt)
    # small arrays and pinned memory keep the plain copy
    small = rng.integers(0, 255, (100, 100)).astype(np.uint8)
    assert torch.equal(N.to_device(small, dev), torch.from_numpy(small).cuda())
