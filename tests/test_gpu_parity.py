"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the golden
vectors of the unmodified reference + cv2 4.13 (tests/golden, oracle/make_golden.py).

Bars (BASELINE.json north_star): corners bit-exact incl. order; tracked
positions within 1e-3 px; status / back-check flags identical; ZNCC within 1e-5
with an identical NaN pattern.  Integer stages (min/max, mask, uint8, Laplacian,
pyrDown) are bit-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import conf_from_golden  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = ["basic", "tiles_mask", "dict_inv"]


@pytest.fixture(scope="module")
def N():
    from karios_b200 import _native
    return _native


@pytest.fixture(scope="module")
def ctx(N):
    c = N.Context(1024, 1024, 20000)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _ks(conf):
    k = conf.laplacian_kernel_size
    return (k["mon"], k["ref"]) if isinstance(k, dict) else (k, k)


@pytest.mark.parametrize("name", CASES)
def test_minmax_mask(golden, N, ctx, name):
    g = golden(name)
    mon, ref = g["mon"], g["ref"]
    mask = ctx.minmax_mask(dev(mon), dev(ref), want_mask=True)
    st = ctx.read_stats()
    assert (st.min_a, st.max_a, st.min_b, st.max_b) == (mon.min(), mon.max(), ref.min(), ref.max())
    want, cnt = O.auto_mask(mon, ref)
    assert np.array_equal(mask.cpu().numpy(), want)
    assert st.valid == cnt
    # nodata handling (klt.py:270-273)
    nd = int(mon[5, 5])
    mask2 = ctx.minmax_mask(dev(mon), dev(ref), nodata_a=nd, want_mask=True)
    want2, cnt2 = O.auto_mask(mon, ref, nd_mon=nd)
    assert np.array_equal(mask2.cpu().numpy(), want2) and ctx.read_stats().valid == cnt2


@pytest.mark.parametrize("name", CASES)
def test_laplacian(golden, N, ctx, name):
    g = golden(name)
    conf = conf_from_golden(g, O.KLTConfiguration)
    mk, rk = _ks(conf)
    ref, mon = dev(g["ref"]), dev(g["mon"])
    assert np.array_equal(ctx.u8_laplacian(ref, rk).cpu().numpy(), g["lap_ref"])
    inv = conf.laplacian_invert_polarity is True
    assert np.array_equal(ctx.u8_laplacian(mon, mk, invert=inv).cpu().numpy(), g["lap_mon"])
    for kk in (1, 3, 5, 9, 11):
        got = ctx.u8_laplacian(ref, kk).cpu().numpy()
        assert np.array_equal(got, g[f"lap_ref_k{kk}"]), f"ksize {kk}: {(got != g[f'lap_ref_k{kk}']).sum()} px differ"
    # slots filled by a previous min/max pass give the same planes
    ctx.minmax_mask(mon, ref)
    assert np.array_equal(ctx.u8_laplacian(ref, rk, slot=1).cpu().numpy(), g["lap_ref"])
    assert np.array_equal(ctx.u8_laplacian(mon, mk, invert=inv, slot=0).cpu().numpy(), g["lap_mon"])


def test_laplacian_dtypes_and_odd_shapes(N, ctx):
    rng = np.random.default_rng(5)
    for shape in ((1, 1), (3, 5), (5, 5), (67, 131), (200, 97)):
        a16 = rng.integers(0, 65535, shape).astype(np.uint16)
        for k in (1, 3, 5, 7, 9, 11):
            want = O.laplacian(O.to_uint8(a16), k)
            got = ctx.u8_laplacian(dev(a16), k).cpu().numpy()
            assert np.array_equal(got, want), (shape, k)
    a = rng.integers(0, 255, (90, 77)).astype(np.uint8)                      # uint8: no normalisation
    assert np.array_equal(ctx.u8_laplacian(dev(a), 7).cpu().numpy(), O.laplacian(a, 7))
    assert np.array_equal(ctx.u8_laplacian(dev(a), 5, invert=True).cpu().numpy(), O.laplacian(255 - a, 5))
    ai = rng.integers(-2000, 9000, (90, 77)).astype(np.int16)               # int16 -> float64 scale
    assert np.array_equal(ctx.u8_laplacian(dev(ai), 7).cpu().numpy(), O.laplacian(O.to_uint8(ai), 7))
    af = rng.normal(100, 30, (90, 77)).astype(np.float32)                    # float32 -> float32 scale
    assert np.array_equal(ctx.u8_laplacian(dev(af), 3).cpu().numpy(), O.laplacian(O.to_uint8(af), 3))
    flat = np.full((40, 40), 7, np.uint16)                                   # max == min -> zeros
    assert not ctx.u8_laplacian(dev(flat), 7).cpu().numpy().any()


def test_laplacian_four_pixel_path(N, ctx):
    """k_laplacian4 (4 pixels per lane; vector-aligned rows, w % 4 == 0, w >= 256, h >= 64):
    _to_uint8 as one multiply-high per pixel (16-bit rasters with a value range >= 256, unsigned
    and signed, with and without inversion) or through the table (narrower ranges), uint8 rasters, mirrored first / last lanes, reflected top / bottom rows, windows of a
    larger raster, every packed kernel size -- bit-exact against the oracle."""
    rng = np.random.default_rng(11)
    for shape in ((64, 256), (70, 260), (131, 492), (300, 1000), (517, 724)):
        narrow = rng.integers(1000, 4000, shape).astype(np.uint16)
        wide = rng.integers(0, 65535, shape).astype(np.uint16)
        smooth = (np.add.outer(np.arange(shape[0]) * 7, np.arange(shape[1]) * 3) % 9000 + 500).astype(np.uint16)
        for a in (narrow, wide, smooth):
            u8 = O.to_uint8(a)
            for k in (3, 5, 7):
                got = ctx.u8_laplacian(dev(a), k).cpu().numpy()
                want = O.laplacian(u8, k)
                assert np.array_equal(got, want), (shape, k, int((got != want).sum()))
        got = ctx.u8_laplacian(dev(narrow), 7, invert=True).cpu().numpy()
        assert np.array_equal(got, O.laplacian(255 - O.to_uint8(narrow), 7)), shape
        b = rng.integers(0, 255, shape).astype(np.uint8)
        for k in (3, 5, 7):
            assert np.array_equal(ctx.u8_laplacian(dev(b), k).cpu().numpy(), O.laplacian(b, k)), (shape, k)
        assert np.array_equal(ctx.u8_laplacian(dev(b), 7, invert=True).cpu().numpy(), O.laplacian(255 - b, 7))
        ai = rng.integers(-3000, 9000, shape).astype(np.int16)
        assert np.array_equal(ctx.u8_laplacian(dev(ai), 7).cpu().numpy(), O.laplacian(O.to_uint8(ai), 7)), shape
    # value ranges either side of the switch between the table and the multiply-high (256), the
    # extremes of both 16-bit types, inverted polarity
    for lo, hi, dt in ((500, 510, np.uint16), (7, 262, np.uint16), (7, 263, np.uint16), (7, 264, np.uint16),
                       (0, 65535, np.uint16), (65000, 65535, np.uint16), (-32768, 32767, np.int16),
                       (-300, -40, np.int16), (-300, -44, np.int16), (-5, 12000, np.int16)):
        a = rng.integers(lo, hi + 1, (96, 512)).astype(dt)
        a[0, 0], a[-1, -1] = lo, hi
        for inv in (False, True):
            u8 = O.to_uint8(a)
            got = ctx.u8_laplacian(dev(a), 7, invert=inv).cpu().numpy()
            assert np.array_equal(got, O.laplacian(255 - u8 if inv else u8, 7)), (lo, hi, dt, inv)
    # aligned window of a larger raster (tile of a resident scene) and a value range that
    # straddles a 4-aligned table base
    big = rng.integers(1003, 1003 + 32000, (400, 1024)).astype(np.uint16)
    t = dev(big)
    for (y, x, h, w) in ((0, 0, 400, 1024), (16, 256, 300, 512), (100, 4, 256, 1020)):
        sub = big[y:y + h, x:x + w]
        got = ctx.u8_laplacian(t[y:y + h, x:x + w], 7).cpu().numpy()
        assert np.array_equal(got, O.laplacian(O.to_uint8(sub), 7)), (y, x, h, w)
    # the context slots (min/max of a previous pass) drive the same kernel
    a = rng.integers(2000, 2300, (200, 512)).astype(np.uint16)
    b2 = rng.integers(100, 60000, (200, 512)).astype(np.uint16)
    ctx.minmax_mask(dev(a), dev(b2))
    assert np.array_equal(ctx.u8_laplacian(dev(a), 5, slot=0).cpu().numpy(), O.laplacian(O.to_uint8(a), 5))
    assert np.array_equal(ctx.u8_laplacian(dev(b2), 7, slot=1).cpu().numpy(), O.laplacian(O.to_uint8(b2), 7))


@pytest.mark.parametrize("name", CASES)
def test_min_eigen_val(golden, N, ctx, name):
    g = golden(name)
    conf = conf_from_golden(g, O.KLTConfiguration)
    eig = ctx.corner_min_eigen_val(dev(g["lap_ref"]), conf.blocksize).cpu().numpy()
    want = O.min_eigen_val(g["lap_ref"], conf.blocksize)
    bad = eig != want
    assert not bad.any(), f"{bad.sum()} of {bad.size} eigenvalues differ, first at {np.argwhere(bad)[:5]}"
    assert eig.max() == g["eig_max"]
    for block in (3, 4, 7, 10):
        e = ctx.corner_min_eigen_val(dev(g["lap_ref"]), block).cpu().numpy()
        assert np.array_equal(e, O.min_eigen_val(g["lap_ref"], block)), block


def test_min_eigen_val_small_images(N):
    rng = np.random.default_rng(9)
    ctx = N.Context(1400, 600, 1000)
    for shape in ((1, 1), (2, 3), (5, 5), (9, 40), (33, 65), (64, 32), (16, 16), (17, 130), (300, 1301),
                  (530, 212)):
        a = rng.integers(0, 255, shape).astype(np.uint8)
        for block in (3, 15):
            e = ctx.corner_min_eigen_val(dev(a), block).cpu().numpy()
            assert np.array_equal(e, O.min_eigen_val(a, block)), (shape, block)


@pytest.mark.parametrize("name", CASES)
def test_good_features(golden, N, ctx, name):
    g = golden(name)
    conf = conf_from_golden(g, O.KLTConfiguration)
    lap, mask = dev(g["lap_ref"]), dev(g["mask_box"])
    p0 = ctx.good_features(lap, mask, conf.maxCorners, conf.qualityLevel, conf.minDistance,
                           conf.blocksize).cpu().numpy()
    want = g["p0"].reshape(-1, 2)
    assert p0.shape == want.shape, (p0.shape, want.shape, ctx.read_stats().as_dict())
    assert np.array_equal(p0, want)                      # same corners, same order
    p0b = ctx.good_features(lap, None, 150, 0.05, 4, 7).cpu().numpy()
    assert np.array_equal(p0b, g["p0_alt"].reshape(-1, 2))


def test_good_features_parameter_sweep(golden, N, ctx):
    g = golden("basic")
    lap = g["lap_ref"]
    eig = O.min_eigen_val(lap, 15)
    for mc, q, md in ((0, 0.1, 10), (50, 0.3, 25.5), (1000, 0.01, 1), (1000, 0.01, 0.5),
                      (7, 0.5, 3), (100000, 0.001, 2.5)):
        want, _, _ = O.select_corners(eig, None, mc, q, md)
        got = ctx.good_features(dev(lap), None, mc, q, md, 15).cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want), (mc, q, md, got.shape, want.shape)
    # flat image / all-zero mask: no corners
    z = np.zeros((50, 60), np.uint8)
    assert ctx.good_features(dev(z), None, 100, 0.1, 10, 15).shape[0] == 0
    assert ctx.good_features(dev(lap), dev(np.zeros_like(lap)), 100, 0.1, 10, 15).shape[0] == 0


def test_corner_modes_agree(golden, N):
    """The two-tier corner response (integer bounds + exact values where needed) and
    OpenCV's arithmetic at every pixel return the same corners in the same order:
    golden Laplacians, noise, saturated / flat images, masks, image borders."""
    from karios_b200 import synth
    rng = np.random.default_rng(17)
    imgs = []
    for name in CASES:
        g = golden(name)
        imgs.append((g["lap_ref"], g["mask_box"]))
    imgs.append((rng.integers(0, 256, (333, 517)).astype(np.uint8), None))
    flat = np.zeros((260, 300), np.uint8)
    flat[60:130, 70:200] = 255
    flat[180:, :] = rng.integers(0, 3, (80, 300))
    flat[:40, 250:] = rng.integers(0, 256, (40, 50))
    imgs.append((flat, None))
    imgs.append((np.full((100, 140), 7, np.uint8), None))                 # nothing to find
    m = (rng.random((333, 517)) > 0.3).astype(np.uint8)
    m[:, :40] = 0
    imgs.append((imgs[3][0], m))
    ref_t, _ = synth.make_pair(700, 900, seed=5)
    big = O.laplacian(O.to_uint8(ref_t.view(torch.int16).numpy().view(np.uint16)), 7)
    imgs.append((big, None))
    edge = rng.integers(0, 40, (200, 240)).astype(np.uint8)               # strongest response on the border
    edge[:3, :] = 255
    edge[:, -2:] = 255
    imgs.append((edge, None))
    c = N.Context(1024, 1024, 5000)
    try:
        for k, (img, mask) in enumerate(imgs):
            for mc, q, md in ((5000, 0.1, 10), (40, 0.01, 3), (700, 0.3, 1.5), (200, 1e-5, 6)):
                d_img = dev(img)
                d_mask = None if mask is None else dev(mask)
                c.set_corner_mode(1)
                want = c.good_features(d_img, d_mask, mc, q, md, 15).cpu().numpy()
                c.set_corner_mode(0)
                got = c.good_features(d_img, d_mask, mc, q, md, 15).cpu().numpy()
                assert got.shape == want.shape and np.array_equal(got, want), (k, mc, q, md, got.shape, want.shape)
    finally:
        c.close()


@pytest.mark.parametrize("name", CASES)
def test_pyr_down(golden, N, ctx, name):
    g = golden(name)
    assert np.array_equal(ctx.pyr_down(dev(g["lap_ref"])).cpu().numpy(), g["pyr_ref"])
    # the last shapes take the 4-outputs-per-lane kernel (w % 4 == 0, w >= 256): mirrored first /
    # last lanes, shifted last warp, odd heights, reflected top / bottom rows, pitched windows
    for shape in ((1, 1), (2, 2), (7, 9), (51, 50), (130, 257), (8, 256), (37, 260), (64, 492), (131, 1000),
                  (301, 724), (66, 1024)):
        a = np.random.default_rng(1).integers(0, 255, shape).astype(np.uint8)
        assert np.array_equal(ctx.pyr_down(dev(a)).cpu().numpy(), O.pyr_down(a)), shape
    # output planes with an aligned pitch (as the context's pyramid planes have): the 4-output kernel
    # for widths whose last full warp ends a few columns before the image edge
    for shape in ((21, 1204), (19, 1208), (33, 964), (18, 1212), (40, 260), (9, 10980)):
        a = np.random.default_rng(3).integers(0, 255, shape).astype(np.uint8)
        dh_, dw_ = (shape[0] + 1) // 2, (shape[1] + 1) // 2
        buf = torch.zeros((dh_, (dw_ + 127) // 128 * 128), dtype=torch.uint8, device="cuda")
        got = ctx.pyr_down(dev(a), out=buf[:, :dw_]).cpu().numpy() if shape[1] <= 1024 else None
        if got is None:
            c2 = N.Context(shape[1], max(shape[0], 16), 100)
            got = c2.pyr_down(dev(a), out=buf[:, :dw_]).cpu().numpy()
            c2.close()
        assert np.array_equal(got, O.pyr_down(a)), shape
    big = np.random.default_rng(2).integers(0, 255, (300, 1024)).astype(np.uint8)
    t = dev(big)
    for (y, x, h, w) in ((0, 0, 300, 1024), (10, 256, 200, 512), (3, 4, 257, 1016)):
        got = ctx.pyr_down(t[y:y + h, x:x + w]).cpu().numpy()
        assert np.array_equal(got, O.pyr_down(big[y:y + h, x:x + w])), (y, x, h, w)


@pytest.mark.parametrize("name", CASES)
def test_pyr_lk(golden, N, ctx, name):
    g = golden(name)
    conf = conf_from_golden(g, O.KLTConfiguration)
    w = conf.matching_winsize
    for p0k, p1k, stk, errk, a, b in (("p0", "lk_p1", "lk_st", "lk_err", "lap_ref", "lap_mon"),
                                      ("lk_p1", "lk_p0r", "lk_st2", "lk_err2", "lap_mon", "lap_ref"),
                                      ("lkb_p0", "lkb_p1", "lkb_st", "lkb_err", "lap_ref", "lap_mon")):
        p1, st, err = ctx.pyr_lk(dev(g[a]), dev(g[b]), dev(g[p0k].reshape(-1, 2)), win=w)
        p1, st, err = p1.cpu().numpy(), st.cpu().numpy(), err.cpu().numpy()
        # (1) against the oracle with exact integer sums: the same arithmetic, bit for bit
        o1, ost, oerr = O.pyr_lk(g[a], g[b], g[p0k], win=w, acc_mode=0)
        assert np.array_equal(st, ost.ravel()), (name, p0k)
        assert np.array_equal(p1, o1.reshape(-1, 2)), np.abs(p1 - o1.reshape(-1, 2)).max()
        ok = st == 1
        assert np.array_equal(err[ok], oerr.ravel()[ok])
        # (2) against cv2 (golden): north-star tolerances
        assert np.array_equal(st, g[stk].ravel())
        d = np.abs(p1 - g[p1k].reshape(-1, 2)).max(-1)
        assert d.max() < 1e-3
        e_ref = g[errk].ravel()[ok]
        assert (np.abs(err[ok] - e_ref) <= 2e-3 * np.maximum(1.0, e_ref)).all()


def test_pyr_lk_windows_and_small_images(N, ctx):
    rng = np.random.default_rng(2)
    base = rng.integers(0, 255, (120, 140)).astype(np.uint8)
    import scipy.ndimage as ndi
    a = ndi.gaussian_filter(base.astype(np.float32), 2.0)
    a = ((a - a.min()) / (a.max() - a.min()) * 255).astype(np.uint8)
    b = np.roll(a, (1, -1), (0, 1))
    pts = np.stack([rng.uniform(-5, 145, 200), rng.uniform(-5, 125, 200)], -1).astype(np.float32)
    for win, lvl in ((5, 1), (15, 2), (21, 0), (29, 3), (25, 1)):
        o1, ost, oerr = O.pyr_lk(a, b, pts.reshape(-1, 1, 2), win=win, max_level=lvl, acc_mode=0)
        p1, st, err = ctx.pyr_lk(dev(a), dev(b), dev(pts), win=win, max_level=lvl)
        assert np.array_equal(st.cpu().numpy(), ost.ravel()), (win, lvl)
        assert np.array_equal(p1.cpu().numpy(), o1.reshape(-1, 2)), (win, lvl)
    small_a, small_b = a[:40, :45].copy(), b[:40, :45].copy()       # no second level (<= win)
    o1, ost, _ = O.pyr_lk(small_a, small_b, pts[:50].reshape(-1, 1, 2), win=25, acc_mode=0)
    p1, st, _ = ctx.pyr_lk(dev(small_a), dev(small_b), dev(pts[:50]), win=25)
    assert np.array_equal(st.cpu().numpy(), ost.ravel()) and np.array_equal(p1.cpu().numpy(), o1.reshape(-1, 2))


@pytest.mark.parametrize("name", CASES)
def test_klt_tracker(golden, name):
    """The drop-in klt_tracker (klt.py:83-172) against the unmodified reference's output."""
    from karios_b200.matcher.klt import klt_tracker
    from karios_b200.core.configuration import KLTConfiguration
    g = golden(name)
    conf = conf_from_golden(g, KLTConfiguration)
    df, ninit = klt_tracker(g["lap_ref"], g["lap_mon"], g["mask_box"], conf)
    assert ninit == int(g["trk_ninit"])
    assert list(df.columns) == ["x0", "y0", "dx", "dy", "score"]
    assert all(str(t) == "float32" for t in df.dtypes)
    assert len(df) == len(g["trk_x0"])                     # identical back-check decisions
    assert np.array_equal(df["x0"].to_numpy(), g["trk_x0"]) and np.array_equal(df["y0"].to_numpy(), g["trk_y0"])
    assert np.abs(df["dx"].to_numpy() - g["trk_dx"]).max() < 1e-3
    assert np.abs(df["dy"].to_numpy() - g["trk_dy"]).max() < 1e-3
    assert np.abs(df["score"].to_numpy() - g["trk_score"]).max() < 2e-2
    # p0 given: mask ignored, same result (klt.py:109-120)
    df2, n2 = klt_tracker(g["lap_ref"], g["lap_mon"], None, conf, p0=g["p0"])
    assert n2 == ninit and np.array_equal(df2["x0"].to_numpy(), df["x0"].to_numpy())
    assert np.array_equal(df2["dx"].to_numpy(), df["dx"].to_numpy())


def test_klt_tracker_no_features():
    from karios_b200.matcher.klt import klt_tracker
    from karios_b200.core.configuration import KLTConfiguration
    z = np.zeros((64, 64), np.uint8)
    assert klt_tracker(z, z, np.ones_like(z), KLTConfiguration()) is None


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("resident", [False, True])
def test_klt_match(golden, name, resident):
    """KLT.match (klt.py:198-349): tiling, masks, offsets, order -- host rasters and
    rasters resident in HBM."""
    from karios_b200.matcher.klt import KLT
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.core.image import ArrayRaster, DeviceRaster
    g = golden(name)
    conf = conf_from_golden(g, KLTConfiguration)
    mk = (lambda a: DeviceRaster(dev(a))) if resident else ArrayRaster
    mask = mk(g["mask"]) if "mask" in g.files else None
    frames = list(KLT(conf).match(mk(g["mon"]), mk(g["ref"]), mask))
    assert len(frames) == int(g["match_ntiles"])
    for i, f in enumerate(frames):
        assert len(f) == len(g[f"match{i}_x0"]), (i, len(f), len(g[f"match{i}_x0"]))
        assert np.array_equal(f["x0"].to_numpy(), g[f"match{i}_x0"])
        assert np.array_equal(f["y0"].to_numpy(), g[f"match{i}_y0"])
        assert np.abs(f["dx"].to_numpy() - g[f"match{i}_dx"]).max() < 1e-3
        assert np.abs(f["dy"].to_numpy() - g[f"match{i}_dy"]).max() < 1e-3
        assert np.abs(f["score"].to_numpy() - g[f"match{i}_score"]).max() < 2e-2


@pytest.mark.parametrize("name", ["auto_modes", "auto_ksize", "auto_device"])
@pytest.mark.parametrize("batched", [True, False])
def test_klt_match_auto_modes(golden, name, batched):
    """laplacian_kernel_size="auto" / laplacian_invert_polarity="auto" (klt.py:438-545): the
    on-device search (kr_auto_ksize) and the per-pair host loop select the kernel sizes and
    polarity the oracle selects (exact-integer LK sums, bit-identical to the CUDA path), return
    its rows, and agree with the unmodified reference's golden output."""
    from karios_b200.matcher.klt import KLT
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.core.image import ArrayRaster
    g = golden(name)
    conf = conf_from_golden(g, KLTConfiguration)
    k = KLT(conf)
    k._batched_auto = batched
    frames = list(k.match(ArrayRaster(g["mon"]), ArrayRaster(g["ref"]), None))
    want = O.match(g["mon"], g["ref"], None, conf_from_golden(g, O.KLTConfiguration), acc_mode=1)
    assert len(frames) == len(want) == int(g["match_ntiles"])
    assert [tuple(x) for x in k._auto_selected_ksizes] == [tuple(t["ksize"]) for t in want]
    if conf.laplacian_invert_polarity == "auto":
        assert k._selected_polarities == [t["polarity"] for t in want]
    for f, t in zip(frames, want):
        assert np.array_equal(f["x0"].to_numpy(), t["x0"]) and np.array_equal(f["y0"].to_numpy(), t["y0"])
        assert np.abs(f["dx"].to_numpy() - t["dx"]).max() < 1e-3
        assert np.abs(f["dy"].to_numpy() - t["dy"]).max() < 1e-3
    # against the reference itself (cv2's float32 sums): same selection on these pairs
    assert [list(x) for x in k._auto_selected_ksizes] == g["tile_ksizes"].tolist()
    assert k._selected_polarities == [str(x) for x in g["tile_polarities"]]
    assert k.auto_selected_ksize == tuple(g["auto_selected_ksize"].tolist())
    for i, f in enumerate(frames):
        assert np.array_equal(f["x0"].to_numpy(), g[f"match{i}_x0"])
        assert np.array_equal(f["y0"].to_numpy(), g[f"match{i}_y0"])
        assert np.abs(f["dx"].to_numpy() - g[f"match{i}_dx"]).max() < 1e-3
        assert np.abs(f["dy"].to_numpy() - g[f"match{i}_dy"]).max() < 1e-3


@pytest.mark.parametrize("name", CASES)
def test_zncc(golden, name):
    import pandas as pd
    from karios_b200.matcher.zncc_service import ZNCCService
    from karios_b200.core.image import ArrayRaster
    g = golden(name)
    df = pd.DataFrame({c: g["all_" + c] for c in ("x0", "y0", "dx", "dy", "score")})
    df.index = df.index + 7                       # the result must carry the caller's index
    z = ZNCCService().compute_zncc(df, ArrayRaster(g["mon"]), ArrayRaster(g["ref"]))
    assert z.index.equals(df.index) and z.dtype == np.float64
    zr = g["zncc"]
    assert np.array_equal(np.isnan(z.to_numpy()), np.isnan(zr))
    ok = ~np.isnan(zr)
    assert np.abs(z.to_numpy()[ok] - zr[ok]).max() < 1e-5
    zo = O.zncc(g["all_x0"], g["all_y0"], g["all_dx"], g["all_dy"], g["mon"], g["ref"])
    assert np.array_equal(z.to_numpy()[ok], zo[ok])        # same arithmetic as the oracle


def test_zncc_known_answers(golden, N, ctx):
    """tests/test_zncc_service.py / test_zncc_zero_std_fix.py of the reference, through kr_zncc."""
    g = golden("zncc_known")
    a, b = g["a"], g["b"]
    f = lambda v: torch.tensor([v], dtype=torch.float32, device="cuda")  # noqa: E731

    def one(p, q, x=28.0, y=28.0, dx=0.0, dy=0.0):
        return float(ctx.zncc(dev(p), dev(q), f(x), f(y), f(dx), f(dy)).cpu()[0])
    assert abs(one(a, b) - float(g["z_ab"])) < 1e-12
    assert abs(one(a, a) - 1.0) < 1e-12
    assert abs(one(a, (65535 - a).astype(np.uint16)) + 1.0) < 1e-12
    assert np.isnan(one(np.full((57, 57), 1234, np.uint16), b))
    assert np.isnan(one(a, b, x=27.0)) and np.isnan(one(a, b, y=29.0))
    big = np.random.default_rng(0).integers(1, 60000, (90, 90)).astype(np.uint16)
    for dx, want in ((0.5, 40), (1.5, 42), (-0.5, 40), (2.5, 42)):            # half to even
        z = one(big, big, 40.0, 40.0, dx, 0.0)
        assert abs(z - O.zncc2(big, big, 40, 40, 40, want, 21)) < 1e-12
    # other dtypes
    a8, b8 = (a >> 8).astype(np.uint8), (b >> 8).astype(np.uint8)
    assert abs(one(a8, b8) - O.zncc2(a8, b8, 28, 28, 28, 28, 21)) < 1e-12
    af, bf = a.astype(np.float32) * 0.37, b.astype(np.float32) * 1.7 - 5
    assert abs(one(af, bf) - O.zncc2(af.astype(np.float64), bf.astype(np.float64), 28, 28, 28, 28, 21)) < 1e-9


MI_TOL = 1e-9        # float64 entropies from integer counts vs NumPy's sums of p log p


def _mi_close(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want)), (np.isnan(got).sum(), np.isnan(want).sum())
    assert (np.nan_to_num(np.abs(got - want)) <= MI_TOL).all(), np.nanmax(np.abs(got - want))


def test_mutual_info_known_answers(golden, N, ctx):
    """Patch families of the reference's known-answer tests
    (/root/reference/tests/test_mutual_info_service.py:15-62) through kr_mutual_info,
    against the values of the unmodified _mutual_info / _mutual_information."""
    g = golden("mi_known")
    f = lambda v: torch.tensor([v], dtype=torch.float32, device="cuda")  # noqa: E731
    for fam in ("ident", "corr", "indep", "unif1", "unif2"):
        for dt in ("f32", "u16", "u8"):
            a, b = g[f"{fam}_{dt}_a"], g[f"{fam}_{dt}_b"]
            out = ctx.mutual_info(dev(a), dev(b), f(28.0), f(28.0), f(0.0), f(0.0)).cpu().numpy()
            _mi_close(out[0], [float(g[f"{fam}_{dt}_studholme"])])
            _mi_close(out[1], [float(g[f"{fam}_{dt}_nmi"])])
    out = ctx.mutual_info(dev(g["ident_u16_a"]), dev(g["ident_u16_a"]), f(28.0), f(28.0), f(0.0), f(0.0))
    assert abs(float(out[0, 0]) - 2.0) < 1e-10 and abs(float(out[1, 0]) - 1.0) < 1e-10
    # int16 rasters with negative values
    rng = np.random.default_rng(8)
    a = rng.integers(-3000, 3000, (57, 57)).astype(np.int16)
    b = (a // 3 + rng.integers(-200, 200, (57, 57))).astype(np.int16)
    out = ctx.mutual_info(dev(a), dev(b), f(28.0), f(28.0), f(0.0), f(0.0)).cpu().numpy()
    _mi_close(out[0], [O.mutual_info_studholme(a, b)])
    _mi_close(out[1], [O.mutual_info_nmi(a, b)])


def test_mutual_info_services(golden):
    """MutualInfoService.compute_mutual_info / ZNCCService.compute_mi drop-ins on the
    golden rows (border rows, flat areas, NaN / Inf pixels in float rasters)."""
    import pandas as pd
    from karios_b200.core.image import ArrayRaster
    from karios_b200.matcher.mutual_info_service import MutualInfoService
    from karios_b200.matcher.zncc_service import ZNCCService
    g = golden("mi_known")
    df = pd.DataFrame({k: g["svc_" + k] for k in ("x0", "y0", "dx", "dy")}, index=np.arange(160)[::-1])
    for suffix in ("", "_f32"):
        mon, ref = ArrayRaster(g["svc_mon" + suffix]), ArrayRaster(g["svc_ref" + suffix])
        st = MutualInfoService().compute_mutual_info(df, mon, ref)
        mi = ZNCCService().compute_mi(df, mon, ref)
        assert st.index.equals(df.index) and mi.index.equals(df.index)
        _mi_close(st.to_numpy(), g["svc_studholme" + suffix])
        _mi_close(mi.to_numpy(), g["svc_nmi" + suffix])
    empty = MutualInfoService().compute_mutual_info(df.iloc[:0], mon, ref)
    assert len(empty) == 0


def test_mutual_info_fused_in_match_tile(N):
    """kr_match_tile with compute_mi: the scores of the fused launch sequence equal
    the oracle's on the rows it produced (NaN below the confidence threshold)."""
    from karios_b200 import synth
    from karios_b200.core.configuration import KLTConfiguration
    ref_t, mon_t = synth.make_pair(400, 520, seed=31)
    to_np = lambda t: t.view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    ref, mon = to_np(ref_t), to_np(mon_t)
    conf = KLTConfiguration(maxCorners=600)
    c = N.Context(520, 400, 600)
    try:
        rows = N.RowBuffers(600, torch.device("cuda"), with_zncc=True, with_mi=True)
        kconf = N.make_conf(conf, compute_zncc=True, compute_mi=True, zncc_min_score=0.4)
        st = c.match_tile(dev(mon), dev(ref), None, (0, 0, 520, 400), kconf, rows)
        n = int(st.n_kept)
        f = rows.f32[:, :n].cpu().numpy()
        got = rows.mi[:, :n].cpu().numpy()
    finally:
        c.close()
    assert n > 100
    want_st, want_mi = O.mutual_info(f[0], f[1], f[2], f[3], mon, ref)
    low = f[4] < np.float32(0.4)
    want_st[low] = np.nan
    want_mi[low] = np.nan
    _mi_close(got[0], want_st)
    _mi_close(got[1], want_mi)
    assert (~np.isnan(got[0])).sum() > 50


def test_large_offset_matcher():
    """LargeOffsetMatcher.match (whole-pixel phase correlation) against the NumPy
    restatement of skimage's algorithm: circular shifts, a zero-filled large shift of a
    uint16 scene (BASELINE config 4), odd sizes."""
    from karios_b200.core.image import ArrayRaster, DeviceRaster
    from karios_b200.matcher.large_offset import LargeOffsetMatcher, phase_cross_correlation_shift
    rng = np.random.default_rng(2)
    base = rng.random((96, 120))
    for dy, dx in ((0, 0), (3, -5), (-20, 11), (47, 59), (-48, -60)):
        moving = np.roll(base, (dy, dx), axis=(0, 1))
        got = phase_cross_correlation_shift(base, moving)
        assert got.dtype == np.float64 and np.array_equal(got, O.phase_cross_correlation_shift(base, moving))
    # uint16 scene displaced by whole pixels with zero fill (BASELINE config 4 geometry)
    ref = (rng.random((601, 777)) * 3000 + 1000).astype(np.uint16)
    mon = O.shift_image(ref, y_off=52, x_off=-37)
    want = O.phase_cross_correlation_shift(mon, ref)
    got = LargeOffsetMatcher(ArrayRaster(ref), ArrayRaster(mon)).match()
    assert np.array_equal(got, want) and np.array_equal(want, [-52, 37]), (got, want)
    got = LargeOffsetMatcher(DeviceRaster(dev(ref.view(np.int16)).view(torch.uint16)),
                             DeviceRaster(dev(mon.view(np.int16)).view(torch.uint16))).match()
    assert np.array_equal(got, want)


def test_shift_image(golden):
    from karios_b200.core.image import shift_image
    g = golden("scene_ops")
    for k, (yo, xo) in enumerate(g["offsets"]):
        for src, key in ((g["a16"], "s16"), (g["a8"], "s8"), (g["af"], "sf")):
            got = shift_image(src, y_off=yo, x_off=xo)
            assert got.dtype == src.dtype and np.array_equal(got, g[f"{key}_{k}"]), (k, key)
    t = shift_image(dev(g["a16"].view(np.int16)), y_off=-5, x_off=7)
    assert t.is_cuda and np.array_equal(t.cpu().numpy().view(np.uint16), g["s16_3"])


def test_scene_scans():
    """_check_quality percentiles, valid-pixel count, DN filter and DEM lookup
    (karios/api/core.py:500-506, 285-290, 687-728, 1050-1053) against NumPy."""
    import pandas as pd
    from karios_b200 import api
    from karios_b200.core.image import ArrayRaster
    rng = np.random.default_rng(6)
    for shape, lo, hi, dt in (((300, 411), 900, 4100, np.uint16), ((257, 300), 0, 65535, np.uint16),
                              ((200, 200), 0, 255, np.uint8), ((123, 321), -2000, 2500, np.int16),
                              ((64, 64), 1200, 1204, np.uint16)):
        a = rng.integers(lo, hi + 1, shape).astype(dt)
        got = api.percentiles_2_98(ArrayRaster(a))
        assert np.array_equal(got, O.percentiles_2_98(a)), (shape, got, O.percentiles_2_98(a))
    skew = (rng.gamma(2.0, 300.0, (500, 300))).astype(np.uint16)
    assert np.array_equal(api.percentiles_2_98(ArrayRaster(skew)), O.percentiles_2_98(skew))
    # rows of a multiple of 8 bytes take the 8-byte-load kernel (k_scan8), others the scalar one;
    # smooth rasters exercise the merged runs, the tail columns and the fused count
    for shape, dt in (((301, 4100), np.uint16), ((300, 4103), np.uint16), ((129, 4104), np.uint8),
                      ((200, 4108), np.int16)):
        yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
        smooth = (1500 + 900 * np.sin(xx / 97.0) * np.cos(yy / 41.0) + rng.integers(0, 3, shape))
        if dt == np.uint8:
            smooth = smooth / 12
        smooth = smooth.astype(dt)
        smooth[::7, ::5] = 0
        m = (rng.random(shape) > 0.3).astype(np.uint8)
        assert np.array_equal(api.percentiles_2_98(ArrayRaster(smooth)), O.percentiles_2_98(smooth)), (shape, dt)
        pct, cnt = api.scene_scan(ArrayRaster(smooth), ArrayRaster(m))
        assert np.array_equal(pct, O.percentiles_2_98(smooth)) and cnt == O.count_valid_pixels(smooth, m)
        assert api.count_valid_pixels(ArrayRaster(smooth)) == O.count_valid_pixels(smooth)
        assert api.scene_scan(ArrayRaster(smooth))[1] == O.count_valid_pixels(smooth)
    q = api.check_quality(ArrayRaster(np.full((50, 50), 7, np.uint16)), ArrayRaster(skew))
    assert q["monitored"]["low_dynamic"] and not q["reference"]["low_dynamic"]
    mon = rng.integers(0, 5, (140, 150)).astype(np.uint16)
    ref = rng.integers(0, 5, (140, 150)).astype(np.uint16)
    mask = (rng.random((140, 150)) > 0.4).astype(np.uint8)
    assert api.count_valid_pixels(ArrayRaster(mon)) == O.count_valid_pixels(mon)
    assert api.count_valid_pixels(ArrayRaster(mon), ArrayRaster(mask)) == O.count_valid_pixels(mon, mask)
    x0 = rng.uniform(0, 149.9, 500).astype(np.float32)
    y0 = rng.uniform(0, 139.9, 500).astype(np.float32)
    df = pd.DataFrame({"x0": x0, "y0": y0, "dx": 0.0, "dy": 0.0})
    out = api.filter_by_dn_values(df, ArrayRaster(mon, no_data_value=3), ArrayRaster(ref), no_values=[0])
    keep = O.filter_by_dn_values(x0, y0, mon, ref, no_values=[0], mon_nd=3)
    assert out.index.equals(df.index[keep]) and 0 < len(out) < 500
    dem = (rng.random((140, 150)) * 3000).astype(np.float32)
    alt = api.dem_altitudes(df, ArrayRaster(dem))
    assert alt.dtype == np.float32 and np.array_equal(alt, dem[y0.astype(int), x0.astype(int)])
    centres, cnt, mean, std = api.altitude_profile(x0, alt, 100)
    grp = pd.DataFrame({"val": x0.astype(np.float64), "po": np.floor_divide(alt, 100)}).groupby("po")["val"]
    assert np.array_equal(cnt, grp.count().to_numpy()) and np.allclose(mean, grp.mean().to_numpy(), rtol=1e-12)
    assert np.allclose(std, grp.std().to_numpy(), rtol=1e-9, equal_nan=True)


def test_large_shift_flow():
    """BASELINE config 4: monitored raster displaced by (+37 columns, -52 rows); the
    detected offset is undone, KLT runs on the shifted raster, dx / dy get the offset
    back (karios/api/core.py:233-252) -- against the oracle run on the same steps."""
    from karios_b200 import api, synth
    from karios_b200.core.configuration import KLTConfiguration
    ref_t, _ = synth.make_pair(500, 640, seed=12)
    to_np = lambda t: t.view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    rng = np.random.default_rng(12)
    ref = (to_np(ref_t).astype(np.int32) + rng.integers(-60, 61, (500, 640))).astype(np.uint16)
    mon = O.shift_image(ref, y_off=52, x_off=-37)               # content moves by (+37, -52)
    conf = KLTConfiguration(maxCorners=800)
    df, applied = api.match_pair_large_shift(mon, ref, None, conf, offset_threshold=15)
    off = O.phase_cross_correlation_shift(mon, ref)
    assert np.array_equal(off, [-52, 37]) and applied == (37.0, -52.0)
    shifted = O.shift_image(mon, y_off=off[0], x_off=off[1])
    want = O.match(shifted, ref, None, O.KLTConfiguration(maxCorners=800))[0]
    assert np.array_equal(df["x0"].to_numpy(), want["x0"]) and np.array_equal(df["y0"].to_numpy(), want["y0"])
    assert np.abs(df["dx"].to_numpy() - (want["dx"] + off[1])).max() < 1e-3
    assert np.abs(df["dy"].to_numpy() - (want["dy"] + off[0])).max() < 1e-3
    assert np.isnan(df["zncc_score"]).all() and len(df) > 300
    small, none = api.match_pair_large_shift(ref, ref, None, conf, offset_threshold=15)
    assert none is None and len(small) > 300 and not np.isnan(small["zncc_score"]).all()


def test_sort_and_unlimited_corners(N):
    """maxCorners = 0 exercises the multi-chunk sort and the full NMS."""
    from karios_b200 import synth
    ref_t, _ = synth.make_pair(900, 1000, seed=3)
    ref = ref_t.view(torch.int16).numpy().view(np.uint16)
    lap = O.laplacian(O.to_uint8(ref), 7)
    c = N.Context(1000, 900, 0)
    try:
        got = c.good_features(dev(lap), None, 0, 0.01, 2, 15).cpu().numpy()
    finally:
        c.close()
    want, _, _ = O.select_corners(O.min_eigen_val(lap, 15), None, 0, 0.01, 2)
    assert got.shape == want.shape and np.array_equal(got, want), (got.shape, want.shape)
    assert len(want) > 3 * 8192


@pytest.mark.parametrize("case", [
    # (h, w, dtype, ksize, tile_size, mask kind, nodata, maxCorners): widths with and without the
    # 4-pixel Laplacian (w % 4), heights that are no multiple of the 64-row segments, 8/16-bit rasters,
    # tiles with remainders, user mask, nodata value, zero blocks (auto mask)
    (1100, 1500, "u16", 7, 20000, None, None, 3000),
    (1047, 1302, "u16", 5, 20000, None, None, 2500),
    (900, 1024, "u8", 7, 20000, None, None, 2000),
    (1300, 1400, "u16", 7, 700, "user", None, 800),
    (777, 1296, "i16", 3, 20000, None, None, 1500),
    (1000, 1204, "u16", 7, 20000, "zeros", 1234, 2000),
])
def test_klt_match_shape_and_dtype_sweep(case):
    """KLT.match + ZNCCService on mid-size synthetic pairs against the oracle (exact-integer LK
    sums: corners identical incl. order, dx/dy bit-level close, ZNCC identical NaN pattern)."""
    import pandas as pd
    from karios_b200 import synth
    from karios_b200.matcher.klt import KLT
    from karios_b200.matcher.zncc_service import ZNCCService
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.core.image import ArrayRaster
    h, w, dt, ksize, tile, mask_kind, nodata, mc = case
    ref_t, mon_t = synth.make_pair(h, w, seed=h + w, shift=(0.4, -0.3))
    ref = ref_t.view(torch.int16).numpy().view(np.uint16).copy()
    mon = mon_t.view(torch.int16).numpy().view(np.uint16).copy()
    if dt == "u8":
        ref, mon = (ref >> 4).astype(np.uint8), (mon >> 4).astype(np.uint8)
    elif dt == "i16":
        ref, mon = (ref.astype(np.int32) - 2500).astype(np.int16), (mon.astype(np.int32) - 2500).astype(np.int16)
    mask = None
    if mask_kind == "user":
        mask = synth.make_mask(h, w, seed=3).numpy()
    elif mask_kind == "zeros":
        mon[200:330, 400:640] = 0
        ref[:40, :300] = 0
        mon[500:520, :] = nodata
    kw = dict(laplacian_kernel_size=ksize, tile_size=tile, maxCorners=mc)
    conf = KLTConfiguration(**kw)
    mon_img, ref_img = ArrayRaster(mon, nodata), ArrayRaster(ref, nodata)
    frames = list(KLT(conf).match(mon_img, ref_img, ArrayRaster(mask) if mask is not None else None))
    want = O.match(mon, ref, mask, O.KLTConfiguration(**kw), nd_mon=nodata, nd_ref=nodata, acc_mode=1)
    assert len(frames) == len(want) and len(frames) >= 1
    for f, t in zip(frames, want):
        assert np.array_equal(f["x0"].to_numpy(), t["x0"]) and np.array_equal(f["y0"].to_numpy(), t["y0"])
        assert np.abs(f["dx"].to_numpy() - t["dx"]).max() < 1e-3
        assert np.abs(f["dy"].to_numpy() - t["dy"]).max() < 1e-3
    df = pd.concat(frames, ignore_index=True)
    z = ZNCCService().compute_zncc(df, mon_img, ref_img).to_numpy()
    zo = O.zncc(df["x0"].to_numpy(), df["y0"].to_numpy(), df["dx"].to_numpy(), df["dy"].to_numpy(), mon, ref)
    assert np.array_equal(np.isnan(z), np.isnan(zo))
    if (~np.isnan(zo)).any():
        assert np.nanmax(np.abs(z - zo)) < 1e-5


def test_full_s2_scene_properties_and_opencv():
    """BASELINE config 2: a full 10980 x 10980 pair through SceneMatcher (one tile,
    default config).  Size-independent properties, then -- when OpenCV is importable
    on this box -- the same scene through the reference's OpenCV calls
    (oracle/cv2_path.py): corners identical up to the documented ~1e-7 class of
    one-ulp eigenvalue differences, dx/dy within 1e-3 px, ZNCC within 1e-5."""
    from karios_b200 import synth
    from karios_b200.api import SceneMatcher
    from karios_b200.core.configuration import KLTConfiguration
    size = 10980
    ref_t, mon_t = synth.make_pair(size, size, seed=1234, device="cuda")
    conf = KLTConfiguration()
    sm = SceneMatcher(size, size, conf, 0.4)
    try:
        tiles, total = sm.match_device(mon_t, ref_t, None)
        st = sm.ctx.read_stats()
        df = sm.to_frame(tiles)
    finally:
        sm.close()
    assert total == len(df) and 0 < total <= 20000 and st.n_corners == 20000
    x0, y0 = df["x0"].to_numpy(), df["y0"].to_numpy()
    key = x0.astype(np.int64) * 65536 + y0.astype(np.int64)
    assert (np.diff(key) > 0).all()                                   # sorted by (x0, y0), unique
    assert x0.min() >= 1 and y0.min() >= 1 and x0.max() <= size - 2 and y0.max() <= size - 2
    # min-distance contract: no two corners closer than minDistance (grid check)
    cell = {}
    for x, y in zip(x0.astype(int), y0.astype(int)):
        cell.setdefault((x // 10, y // 10), []).append((x, y))
    for (cx, cy), pts in cell.items():
        for ddx in (-1, 0, 1):
            for ddy in (-1, 0, 1):
                for (x, y) in pts:
                    for (u, v) in cell.get((cx + ddx, cy + ddy), []):
                        assert (u, v) == (x, y) or (x - u) ** 2 + (y - v) ** 2 >= 100
    assert (df["score"] > 0).all() and (df["score"] <= 1).all()
    z = df["zncc_score"].to_numpy()
    inner = (x0 >= 30) & (y0 >= 30) & (x0 < size - 30) & (y0 < size - 30) & (df["score"].to_numpy() >= 0.4)
    assert not np.isnan(z[inner]).any() and np.nanmax(np.abs(z)) <= 1 + 1e-9
    assert abs(df["dx"].mean() - 0.30) < 0.1 and abs(df["dy"].mean() + 0.20) < 0.1     # the synthetic shift

    to_np = lambda t: t.cpu().view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    ref, mon = to_np(ref_t), to_np(mon_t)
    c, how = _reference_rows(mon, ref, None, {})
    key_cv = c["x0"].astype(np.int64) * 65536 + c["y0"].astype(np.int64)
    common, ia, ib = np.intersect1d(key, key_cv, return_indices=True)
    only_gpu, only_ref = np.setdiff1d(key, key_cv), np.setdiff1d(key_cv, key)
    # rows = corners that passed the back-check (status): a row on one side only is either a corner
    # difference (the documented one-ulp eigenvalue tie class, SURVEY A.3: ~1e-7 of the pixels) or a
    # differing status / back-check decision
    agreement = 1.0 - (len(only_gpu) + len(only_ref)) / max(int(st.n_corners), 1)
    print(f"reference: {how}")
    print(f"rows: gpu {len(key)} ref {len(key_cv)} common {len(common)}; only gpu {_xy(only_gpu)}, only ref {_xy(only_ref)}; "
          f"status agreement {agreement:.6f}")
    assert len(only_gpu) + len(only_ref) <= 4, (_xy(only_gpu), _xy(only_ref))
    assert agreement > 0.999
    ddx = np.abs(df["dx"].to_numpy()[ia] - c["dx"][ib])
    ddy = np.abs(df["dy"].to_numpy()[ia] - c["dy"][ib])
    print(f"max |ddx| {ddx.max():.2e} max |ddy| {ddy.max():.2e} identical {(np.maximum(ddx, ddy) == 0).mean():.4f}")
    assert ddx.max() < 1e-3 and ddy.max() < 1e-3
    zc = c["zncc"][ib]
    zg = z[ia]
    assert np.array_equal(np.isnan(zc), np.isnan(zg))
    assert np.nanmax(np.abs(zc - zg)) < 1e-5
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        json.dump({"reference": how, "n_init_gpu": int(st.n_corners), "rows_gpu": int(len(key)),
                   "rows_reference": int(len(key_cv)), "common": int(len(common)),
                   "only_gpu_xy": _xy(only_gpu), "only_reference_xy": _xy(only_ref),
                   "status_agreement": agreement,
                   "max_abs_ddx": float(ddx.max()), "max_abs_ddy": float(ddy.max()),
                   "bit_identical_fraction": float((np.maximum(ddx, ddy) == 0).mean()),
                   "zncc_max_abs_diff": float(np.nanmax(np.abs(zc - zg)))},
                  open(os.path.join(out, "full_scene_parity.json"), "w"))


def _xy(keys):
    return [(int(k) // 65536, int(k) % 65536) for k in keys]


def _reference_rows(mon, ref, mask, conf_kw):
    """The scene through the UNMODIFIED reference (oracle/_ref: KLT.match + compute_zncc, as
    bench.py's reference arm runs it) when its files are present, else through the restated
    glue around the same cv2 calls (oracle/cv2_path.py).  -> (list / dict of columns, what ran)"""
    import logging
    import os
    from oracle import cv2_path as P
    from oracle import ref_run, vendor_ref
    if vendor_ref.available() or os.path.isdir("/root/reference/karios/matcher"):
        logging.getLogger("karios").setLevel(logging.ERROR)
        rdf, _, how = ref_run.run_pair(mon, ref, mask, threshold=0.4, **conf_kw)
        cols = {k: rdf[k].to_numpy() for k in ("x0", "y0", "dx", "dy", "score")}
        cols["zncc"] = rdf["zncc_score"].to_numpy()
        return cols, how
    if not P.HAVE_CV2:
        pytest.skip("neither the reference files nor cv2 are available here")
    tiles_cv, _ = P.match_scene(mon, ref, mask, O.KLTConfiguration(**conf_kw))
    cols = {k: np.concatenate([t[k] for t in tiles_cv]) for k in ("x0", "y0", "dx", "dy", "score", "zncc")}
    return cols, "oracle/cv2_path.py"


def test_match_many_equals_sequential():
    """SceneMatcher.match_many (units of several pairs in flight on their own contexts /
    streams) returns, pair by pair and tile by tile, exactly what match_device returns
    for each pair on its own -- rows, ZNCC and both mutual-information scores."""
    from karios_b200 import synth
    from karios_b200.api import SceneMatcher
    from karios_b200.core.configuration import KLTConfiguration
    pairs = []
    for seed in (3, 4, 5, 6, 7):
        ref_t, mon_t = synth.make_pair(420, 610, seed=seed, device="cuda")
        pairs.append((mon_t, ref_t))
    conf = KLTConfiguration(maxCorners=400, tile_size=350)
    seq = SceneMatcher(420, 610, conf, 0.4, depth=1, with_mi=True)
    par = SceneMatcher(420, 610, conf, 0.4, depth=3, with_mi=True)
    try:
        assert len(par.windows) == 4
        got, total = par.match_many(pairs)
        want_total = 0
        for (mon_t, ref_t), tiles in zip(pairs, got):
            want, n = seq.match_device(mon_t, ref_t)
            want_total += n
            assert len(want) == len(tiles) == 4
            for a, b in zip(tiles, want):
                for x, y in zip(a, b):
                    assert torch.equal(torch.nan_to_num(x, nan=-7.0), torch.nan_to_num(y, nan=-7.0))
        assert total == want_total > 1000
        df = par.to_frame(got[0])
        assert list(df.columns) == ["x0", "y0", "dx", "dy", "score", "zncc_score", "mutual_info_score", "mi_score"]
    finally:
        seq.close()
        par.close()


def test_full_s2_mask_tiles_dem_vs_opencv():
    """BASELINE configs 2 ("full tiling", tile_size 6000 -> 4 tiles) and 3 (user mask
    zeroing ~30 % of the scene, DEM altitudes): a 10980 x 10980 pair through
    SceneMatcher.match_many with two pairs' worth of tiles in flight, against the
    reference's OpenCV calls on the same arrays (oracle/cv2_path.py)."""
    from karios_b200 import api, synth
    from karios_b200.api import SceneMatcher
    from karios_b200.core.configuration import KLTConfiguration
    from karios_b200.core.image import DeviceRaster
    size = 10980
    ref_t, mon_t = synth.make_pair(size, size, seed=1235, device="cuda")
    mask_t = synth.make_mask(size, size, seed=99, device="cuda")
    conf = KLTConfiguration(tile_size=6000)
    sm = SceneMatcher(size, size, conf, 0.4, depth=2)
    try:
        assert len(sm.windows) == 4
        res, total = sm.match_many([(mon_t, ref_t)], mask=mask_t)
        df = sm.to_frame(res[0])
    finally:
        sm.close()
    to_np = lambda t: t.cpu().view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    ref, mon, mask = to_np(ref_t), to_np(mon_t), mask_t.cpu().numpy()
    c, how = _reference_rows(mon, ref, mask, {"tile_size": 6000})
    assert len(res[0]) == 4
    f = torch.cat([t[0] for t in res[0]], dim=1).cpu().numpy()
    z = torch.cat([t[1] for t in res[0]]).cpu().numpy()
    # rows of the reference: tiles in x-outer / y-inner order, each sorted by (x0, y0) -- the
    # concatenation is compared in that order, row by row, after aligning on the keys
    key = f[0].astype(np.int64) * 65536 + f[1].astype(np.int64)
    key_cv = c["x0"].astype(np.int64) * 65536 + c["y0"].astype(np.int64)
    common, ia, ib = np.intersect1d(key, key_cv, return_indices=True)
    only_gpu, only_ref = np.setdiff1d(key, key_cv), np.setdiff1d(key_cv, key)
    print(f"reference: {how}; rows gpu {len(key)} ref {len(key_cv)} common {len(common)}; "
          f"only gpu {_xy(only_gpu)}, only ref {_xy(only_ref)}")
    assert len(only_gpu) + len(only_ref) <= 4, (_xy(only_gpu), _xy(only_ref))
    if not len(only_gpu) and not len(only_ref):
        assert np.array_equal(key, key_cv)                # same rows in the same (tile, x0, y0) order
    worst = max(np.abs(f[2][ia] - c["dx"][ib]).max(), np.abs(f[3][ia] - c["dy"][ib]).max())
    assert np.array_equal(np.isnan(z[ia]), np.isnan(c["zncc"][ib]))
    zworst = np.nanmax(np.abs(z[ia] - c["zncc"][ib]))
    m = mask[f[1].astype(int), f[0].astype(int)]
    assert (m != 0).all()                                 # no corner in a masked-out pixel
    print(f"tiles+mask: max |d| {worst:.2e}, zncc {zworst:.2e}")
    assert worst < 1e-3 and zworst < 1e-5
    # config 3: DEM altitudes of the key points and valid-pixel count under the mask
    dem_t = (torch.arange(size, device="cuda", dtype=torch.float32)[:, None] * 0.2 +
             torch.arange(size, device="cuda", dtype=torch.float32)[None, :] * 0.07)
    alt = api.dem_altitudes(df, DeviceRaster(dem_t))
    want = dem_t.cpu().numpy()[df["y0"].to_numpy().astype(int), df["x0"].to_numpy().astype(int)]
    assert np.array_equal(alt, want)
    assert api.count_valid_pixels(DeviceRaster(mon_t), DeviceRaster(mask_t)) == O.count_valid_pixels(mon, mask)


def test_full_s2_large_shift_config4():
    """BASELINE config 4 at full size: a 10980 x 10980 pair whose monitored raster is displaced
    by whole pixels (+37 columns, -52 rows, zero fill).  The detection has to recover exactly that
    offset (known answer; float64 FFTs of 120 Mpx on the device), and the rows have to equal the
    unmodified reference's KLT on the shifted raster with the offsets added back
    (karios/api/core.py:233-252, 746-786), ZNCC skipped (:876)."""
    from karios_b200 import _native as N
    from karios_b200 import api, synth
    from karios_b200.core.configuration import KLTConfiguration
    size = 10980
    ref_t, mon_t = synth.make_pair(size, size, seed=1236, device="cuda")
    mon_far = N.shift_image(mon_t, 52, -37)
    torch.cuda.reset_peak_memory_stats()
    df, applied = api.match_pair_large_shift(mon_far, ref_t, None, KLTConfiguration(), offset_threshold=10)
    peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    assert applied == (37.0, -52.0)
    assert np.isnan(df["zncc_score"]).all() and len(df) > 15000
    to_np = lambda t: t.cpu().view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    ref, mon = to_np(ref_t), to_np(mon_t)
    far = O.shift_image(mon, y_off=52, x_off=-37)
    assert np.array_equal(to_np(mon_far), far)                          # kr_shift_image, full size
    shifted = O.shift_image(far, y_off=-52, x_off=37)
    del far, mon_far
    c, how = _reference_rows(shifted, ref, None, {})
    key = df["x0"].to_numpy().astype(np.int64) * 65536 + df["y0"].to_numpy().astype(np.int64)
    key_cv = c["x0"].astype(np.int64) * 65536 + c["y0"].astype(np.int64)
    common, ia, ib = np.intersect1d(key, key_cv, return_indices=True)
    only_gpu, only_ref = np.setdiff1d(key, key_cv), np.setdiff1d(key_cv, key)
    print(f"config 4: {how}; rows gpu {len(key)} ref {len(key_cv)}; only gpu {_xy(only_gpu)}, only ref {_xy(only_ref)}; "
          f"peak device memory {peak_gb:.2f} GB")
    assert len(only_gpu) + len(only_ref) <= 4
    assert np.abs(df["dx"].to_numpy()[ia] - (c["dx"][ib] + np.float32(37))).max() < 1e-3
    assert np.abs(df["dy"].to_numpy()[ia] - (c["dy"][ib] + np.float32(-52))).max() < 1e-3
    assert peak_gb < 20


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()
