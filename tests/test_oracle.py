"""Pin the CPU oracle (oracle/klt_oracle.c + oracle/oracle.py) against the golden
vectors produced by the unmodified reference + cv2 4.13 (oracle/make_golden.py),
and against the reference's own ZNCC known-answer tests
(/root/reference/tests/test_zncc_service.py:26-34,63-70,93-104,107-125,
tests/test_zncc_zero_std_fix.py:65-69)."""
import numpy as np
import pytest

from oracle import oracle as O
from conftest import conf_from_golden

CASES = ["basic", "tiles_mask", "dict_inv"]


def _ks(conf):
    k = conf.laplacian_kernel_size
    return (k["mon"], k["ref"]) if isinstance(k, dict) else (k, k)


@pytest.mark.parametrize("name", CASES)
def test_to_uint8_and_laplacian(golden, name):
    g = golden(name)
    conf = conf_from_golden(g, O.KLTConfiguration)
    mk, rk = _ks(conf)
    ref_u8, mon_u8 = O.to_uint8(g["ref"]), O.to_uint8(g["mon"])
    assert np.array_equal(ref_u8, g["ref_u8"]) and np.array_equal(mon_u8, g["mon_u8"])
    mn, mx = float(g["ref"].min()), float(g["ref"].max())
    assert np.array_equal(O.to_uint8_lut(mn, mx)[g["ref"]], g["ref_u8"])
    assert np.array_equal(O.laplacian(ref_u8, rk), g["lap_ref"])
    m = (255 - mon_u8) if conf.laplacian_invert_polarity is True else mon_u8
    assert np.array_equal(O.laplacian(m, mk), g["lap_mon"])
    for kk in (1, 3, 5, 9, 11):
        assert np.array_equal(O.laplacian(ref_u8, kk), g[f"lap_ref_k{kk}"]), kk


def test_auto_mask(golden):
    g = golden("basic")
    mask, cnt = O.auto_mask(g["mon"], g["ref"])
    assert np.array_equal(mask, g["mask_box"]) and cnt == int(g["mask_box"].sum())
    assert cnt < mask.size        # the case has zero blocks


@pytest.mark.parametrize("name", CASES)
def test_min_eigen_val_and_corners(golden, name):
    g = golden(name)
    conf = conf_from_golden(g, O.KLTConfiguration)
    eig = O.min_eigen_val(g["lap_ref"], conf.blocksize)
    h, w = eig.shape
    if "eig" in g.files:
        assert np.array_equal(eig, g["eig"])
    assert eig.max() == g["eig_max"]
    assert np.array_equal(eig[[0, 1, h // 2, h - 2, h - 1]], g["eig_rows"])
    assert np.array_equal(eig[:, [0, 1, w // 2, w - 5, w - 4, w - 2, w - 1]], g["eig_cols"])
    p0 = O.good_features(g["lap_ref"], g["mask_box"], conf.maxCorners, conf.qualityLevel,
                         conf.minDistance, conf.blocksize)
    assert np.array_equal(p0, g["p0"])          # identical corners, identical order
    p0b = O.good_features(g["lap_ref"], None, 150, 0.05, 4, 7)
    assert np.array_equal(p0b, g["p0_alt"])


@pytest.mark.parametrize("name", CASES)
def test_pyr_down(golden, name):
    g = golden(name)
    assert np.array_equal(O.pyr_down(g["lap_ref"]), g["pyr_ref"])


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("acc_mode", [0, 1])
def test_pyr_lk(golden, name, acc_mode):
    """Positions within 1e-3 px (north star), status identical, err within 1e-2
    on tracked points.  acc_mode 0 = exact integer sums (what the CUDA kernel
    does), 1 = cv2's float32 SIMD accumulation order."""
    g = golden(name)
    conf = conf_from_golden(g, O.KLTConfiguration)
    w = conf.matching_winsize
    for p0k, p1k, stk, errk, a, b in (("p0", "lk_p1", "lk_st", "lk_err", "lap_ref", "lap_mon"),
                                      ("lk_p1", "lk_p0r", "lk_st2", "lk_err2", "lap_mon", "lap_ref"),
                                      ("lkb_p0", "lkb_p1", "lkb_st", "lkb_err", "lap_ref", "lap_mon")):
        p1, st, err = O.pyr_lk(g[a], g[b], g[p0k], win=w, acc_mode=acc_mode)
        assert np.array_equal(st, g[stk]), (name, p0k)
        d = np.abs(p1 - g[p1k]).reshape(-1, 2).max(-1)
        assert d.max() < 1e-3
        assert (d == 0).mean() > (0.97 if acc_mode else 0.85)
        ok = g[stk].ravel() == 1
        e_ref = g[errk].ravel()[ok]
        assert (np.abs(err.ravel()[ok] - e_ref) <= 2e-3 * np.maximum(1.0, e_ref)).all()


@pytest.mark.parametrize("name", CASES)
def test_klt_tracker_and_match(golden, name):
    g = golden(name)
    conf = conf_from_golden(g, O.KLTConfiguration)
    cols, ninit = O.klt_tracker(g["lap_ref"], g["lap_mon"], g["mask_box"], conf, acc_mode=1)
    assert ninit == int(g["trk_ninit"])
    assert np.array_equal(cols["x0"], g["trk_x0"]) and np.array_equal(cols["y0"], g["trk_y0"])
    for c in ("dx", "dy"):
        assert np.abs(cols[c] - g["trk_" + c]).max() < 1e-3
    assert np.abs(cols["score"] - g["trk_score"]).max() < 1e-2
    mask = g["mask"] if "mask" in g.files else None
    tiles = O.match(g["mon"], g["ref"], mask, conf, acc_mode=1)
    assert len(tiles) == int(g["match_ntiles"])
    for i, t in enumerate(tiles):
        assert np.array_equal(t["x0"], g[f"match{i}_x0"]) and np.array_equal(t["y0"], g[f"match{i}_y0"])
        assert np.abs(t["dx"] - g[f"match{i}_dx"]).max() < 1e-3
        assert np.abs(t["dy"] - g[f"match{i}_dy"]).max() < 1e-3


@pytest.mark.parametrize("name", CASES)
def test_zncc(golden, name):
    g = golden(name)
    z = O.zncc(g["all_x0"], g["all_y0"], g["all_dx"], g["all_dy"], g["mon"], g["ref"])
    zr = g["zncc"]
    assert np.array_equal(np.isnan(z), np.isnan(zr))
    assert np.isnan(zr).any() and (~np.isnan(zr)).any()
    ok = ~np.isnan(zr)
    assert np.abs(z[ok] - zr[ok]).max() < 1e-12


def test_zncc_known_answers(golden):
    g = golden("zncc_known")
    a, b = g["a"], g["b"]
    c = np.float32(28)
    z = np.float32(0)

    def one(p, q):
        return O.zncc([c], [c], [z], [z], q, p)[0]      # zncc(..., monitored, reference)
    assert abs(one(a, b) - float(g["z_ab"])) < 1e-12
    assert abs(one(a, a) - 1.0) < 1e-12
    assert abs(one(a, (65535 - a).astype(np.uint16)) + 1.0) < 1e-12
    assert np.isnan(one(np.full((57, 57), 1234, np.uint16), b))      # zero std -> NaN
    # chip geometry / border rule: one pixel off-centre on a 57x57 image -> NaN
    assert np.isnan(O.zncc([np.float32(27)], [c], [z], [z], b, a)[0])
    assert np.isnan(O.zncc([c], [np.float32(29)], [z], [z], b, a)[0])
    # _zncc2 argument checks (tests/test_zncc_service.py:93-104)
    with pytest.raises(IndexError):
        O.zncc2(a, b, 2, 2, 2, 2, 5)
    with pytest.raises(ValueError):
        O.zncc2(a, b, 28, 28, 28, 28, -1)
    # identical 3x3 patches -> 1.0 (tests/test_zncc_service.py:26-34)
    p = np.arange(9, dtype=np.float64).reshape(3, 3)
    assert abs(O.zncc2(p, p, 1, 1, 1, 1, 1) - 1.0) < 1e-12


def test_round_half_even_of_float32_sum():
    """zncc_service.py:197-198: round(np.float32 + np.float32) is half-to-even."""
    img = np.random.default_rng(0).integers(1, 60000, (80, 80)).astype(np.uint16)
    x0 = np.float32(40)
    for dx, want in ((np.float32(0.5), 40), (np.float32(1.5), 42), (np.float32(-0.5), 40)):
        z = O.zncc([x0], [x0], [dx], [np.float32(0)], img, img)[0]
        zr = O.zncc2(img, img, 40, 40, 40, want, 21)
        assert abs(z - zr) < 1e-12


@pytest.mark.parametrize("name", CASES)
def test_cv2_path_matches_golden(golden, name):
    """oracle/cv2_path.py (the cv2-based restatement timed as the CPU baseline)
    reproduces the unmodified reference's KLT.match + compute_zncc output."""
    from oracle import cv2_path as P
    if not P.HAVE_CV2:
        pytest.skip("cv2 not importable")
    g = golden(name)
    conf = conf_from_golden(g, O.KLTConfiguration)
    mask = g["mask"] if "mask" in g.files else None
    tiles, total = P.match_scene(g["mon"], g["ref"], mask, conf, zncc_threshold=-1.0)
    assert len(tiles) == int(g["match_ntiles"]) and total == len(g["zncc"])
    for i, t in enumerate(tiles):
        for c in ("x0", "y0", "dx", "dy", "score"):
            assert np.array_equal(t[c], g[f"match{i}_{c}"]), (i, c)
    z = np.concatenate([t["zncc"] for t in tiles])
    assert np.array_equal(np.isnan(z), np.isnan(g["zncc"]))
    assert np.nanmax(np.abs(z - g["zncc"])) < 1e-12


# ------------------------------------------------------------------ f1: mutual information
MI_FAMILIES = ["ident", "corr", "indep", "unif1", "unif2"]


def _same(a, b, tol=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.array_equal(np.isnan(a), np.isnan(b)) and (np.nan_to_num(np.abs(a - b)) <= tol).all()


@pytest.mark.parametrize("fam", MI_FAMILIES)
@pytest.mark.parametrize("dt", ["f32", "u16", "u8"])
def test_mutual_info_patches_match_reference(golden, fam, dt):
    """oracle restatement vs the unmodified _mutual_info / _mutual_information on the
    patch families of /root/reference/tests/test_mutual_info_service.py:15-62."""
    g = golden("mi_known")
    a, b = g[f"{fam}_{dt}_a"], g[f"{fam}_{dt}_b"]
    assert _same(O.mutual_info_studholme(a, b), g[f"{fam}_{dt}_studholme"])
    assert _same(O.mutual_info_nmi(a, b), g[f"{fam}_{dt}_nmi"])


def test_mutual_info_known_answers(golden):
    """The reference's own assertions (test_mutual_info_service.py:15-62)."""
    g = golden("mi_known")
    assert abs(float(g["ident_u16_studholme"]) - 2.0) < 1e-10
    assert 1.0 <= float(g["corr_f32_studholme"]) <= 2.0
    assert 1.0 <= float(g["indep_f32_studholme"]) < 1.2
    assert abs(float(g["unif1_f32_studholme"]) - 1.0) < 1e-10
    assert np.isnan(g["unif2_f32_studholme"]) and np.isnan(g["unif2_u16_nmi"])


def test_mutual_info_service_rows_match_reference(golden):
    g = golden("mi_known")
    cols = [g[k] for k in ("svc_x0", "svc_y0", "svc_dx", "svc_dy")]
    st, mi = O.mutual_info(*cols, g["svc_mon"], g["svc_ref"])
    assert _same(st, g["svc_studholme"]) and _same(mi, g["svc_nmi"])
    assert 0 < np.isnan(st).sum() < len(st)
    st, mi = O.mutual_info(*cols, g["svc_mon_f32"], g["svc_ref_f32"])
    assert _same(st, g["svc_studholme_f32"]) and _same(mi, g["svc_nmi_f32"])
    assert np.isnan(st).sum() > np.isnan(g["svc_studholme"]).sum()     # the NaN / Inf pixels


def test_integer_joint_histogram_equals_numpy():
    """The float-free binning the CUDA kernel uses is np.histogram2d's."""
    rng = np.random.default_rng(5)
    for trial in range(60):
        hi = int(rng.choice([1, 2, 31, 32, 33, 255, 1000, 4095, 65535]))
        dt = np.uint16 if hi > 255 else np.uint8
        a = rng.integers(0, hi + 1, (57, 57)).astype(dt)
        b = rng.integers(0, hi + 1, (57, 57)).astype(dt)
        if trial % 7 == 0:
            b[:] = b[0, 0]
        want = np.histogram2d(a.ravel(), b.ravel(), bins=32)[0]
        assert np.array_equal(O.joint_histogram_int(a, b), want), (trial, hi)
    a = rng.integers(-3000, 3000, (57, 57)).astype(np.int16)
    b = rng.integers(-32768, 32767, (57, 57)).astype(np.int16)
    assert np.array_equal(O.joint_histogram_int(a, b), np.histogram2d(a.ravel(), b.ravel(), bins=32)[0])


# ------------------------------------------------------------------ f2 / f3: scene-level passes
def test_shift_image_matches_reference(golden):
    g = golden("scene_ops")
    for k, (yo, xo) in enumerate(g["offsets"]):
        for src, key in ((g["a16"], "s16"), (g["a8"], "s8"), (g["af"], "sf")):
            assert np.array_equal(O.shift_image(src, y_off=yo, x_off=xo), g[f"{key}_{k}"]), (k, key)


def test_phase_correlation_known_shifts():
    """skimage is absent (parity unpinned by reference vectors): the restatement must
    recover circular and zero-filled whole-pixel shifts, with skimage's sign."""
    rng = np.random.default_rng(2)
    base = rng.random((96, 120))
    for dy, dx in ((0, 0), (3, -5), (-20, 11), (47, 59), (-48, -60)):
        moving = np.roll(base, (dy, dx), axis=(0, 1))
        assert np.array_equal(O.phase_cross_correlation_shift(base, moving), [-dy, -dx]), (dy, dx)
    ref = (rng.random((160, 200)) * 3000 + 1000).astype(np.uint16)
    mon = O.shift_image(ref, y_off=-13, x_off=21)              # config 4 style: zero fill
    # LargeOffsetMatcher calls (mon, ref); shifting mon by the result re-aligns it
    off = O.phase_cross_correlation_shift(mon, ref)
    back = O.shift_image(mon, y_off=off[0], x_off=off[1])
    inner = (slice(30, 130), slice(30, 170))
    assert np.array_equal(back[inner], ref[inner])


def test_filter_and_count_restatements():
    rng = np.random.default_rng(4)
    mon = rng.integers(0, 5, (40, 50)).astype(np.uint16)
    ref = rng.integers(0, 5, (40, 50)).astype(np.uint16)
    mask = (rng.random((40, 50)) > 0.4).astype(np.uint8)
    assert O.count_valid_pixels(mon) == int((mon != 0).sum())
    assert O.count_valid_pixels(mon, mask) == int(((mon != 0) & (mask != 0)).sum())
    x0 = rng.uniform(0, 49.9, 200).astype(np.float32)
    y0 = rng.uniform(0, 39.9, 200).astype(np.float32)
    keep = O.filter_by_dn_values(x0, y0, mon, ref, no_values=[0], mon_nd=3)
    xi, yi = x0.astype(int), y0.astype(int)
    want = (mon[yi, xi] != 0) & (ref[yi, xi] != 0) & (mon[yi, xi] != 3)
    assert np.array_equal(keep, want) and 0 < keep.sum() < 200


# ------------------------------------------------------------------ a14: automatic modes
@pytest.mark.parametrize("name", ["auto_modes", "auto_ksize", "auto_device"])
def test_auto_ksize_and_polarity_match_reference(golden, name):
    """klt.py:438-545 restated (oracle.auto_ksize_search / track_once / match_tile) against the
    unmodified reference run by oracle/make_golden.py: the kernel sizes and polarities selected
    per tile, the inlier ratio of each of the 25 kernel-size pairs, and the winning rows.
    (OpenCV's float32 accumulation order, acc_mode 0: a back-check decision on a chaotic track of
    the losing polarity can flip with the exact-integer sums the CUDA path uses.)"""
    g = golden(name)
    conf = conf_from_golden(g, O.KLTConfiguration)
    tiles = O.match(g["mon"], g["ref"], None, conf, acc_mode=0)
    assert len(tiles) == int(g["match_ntiles"])
    assert [list(t["ksize"]) for t in tiles] == g["tile_ksizes"].tolist()
    if conf.laplacian_invert_polarity == "auto":
        assert [t["polarity"] for t in tiles] == [str(x) for x in g["tile_polarities"]]
    for i, t in enumerate(tiles):
        assert np.array_equal(t["x0"], g[f"match{i}_x0"]) and np.array_equal(t["y0"], g[f"match{i}_y0"])
        assert np.abs(t["dx"] - g[f"match{i}_dx"]).max() < 1e-3
        assert np.abs(t["dy"] - g[f"match{i}_dy"]).max() < 1e-3
    # the 25 ratios of the first tile, both polarities where recorded
    ts = conf.tile_size
    mon, ref = g["mon"][:ts, :ts], g["ref"][:ts, :ts]
    mask, _ = O.auto_mask(mon, ref)
    for label in ("normal", "inverted"):
        if f"scores_{label}" not in g.files:
            continue
        u8 = O.to_uint8(mon)
        _, scores, best = O.auto_ksize_search(255 - u8 if label == "inverted" else u8, ref, mask, conf, acc_mode=0)
        want = {(int(a), int(b)): r for a, b, r in g[f"scores_{label}"]}
        assert scores.keys() == want.keys()
        # a back-check decision on a chaotic track (wrong polarity) may flip with the last bit of
        # the float32 sums: at most one pair off by a point or two, the selection unchanged
        off = [k for k in want if abs(scores[k] - want[k]) >= 1e-12]
        assert len(off) <= 1 and all(abs(scores[k] - want[k]) <= 0.02 for k in off), (label, off)
        assert list(best) == g[f"best_{label}"].tolist()


# ---------------------------------------------------------------------------
# The vendored reference (oracle/_ref, placed by oracle/vendor_ref.py during build()) that
# bench.py's reference arm and cpu_baseline leg run on the GPU box.
def test_vendored_reference_is_unmodified():
    import hashlib
    import os
    from oracle import vendor_ref
    if not vendor_ref.available():
        if not os.path.isdir(os.path.join(vendor_ref.REF_SRC, "karios", "matcher")):
            pytest.skip("no reference tree and no vendored copy here")
        vendor_ref.vendor()
    assert vendor_ref.available() and vendor_ref.verify() == []
    if os.path.isdir(os.path.join(vendor_ref.REF_SRC, "karios", "matcher")):
        for rel in vendor_ref.FILES:
            src = os.path.join(vendor_ref.REF_SRC, rel)
            if os.path.exists(src):
                a = hashlib.sha256(open(src, "rb").read()).hexdigest()
                b = hashlib.sha256(open(os.path.join(vendor_ref.DEST, rel), "rb").read()).hexdigest()
                assert a == b, rel


def test_reference_run_matches_restated_glue():
    """oracle/ref_run.py (the unmodified KLT.match + compute_zncc, the bench's reference arm)
    and oracle/cv2_path.py (the restated glue) give the same rows on a small pair with a
    remainder tile."""
    import logging
    import os
    import torch
    from karios_b200 import synth
    from oracle import cv2_path as P
    from oracle import ref_run, vendor_ref
    if not P.HAVE_CV2:
        pytest.skip("cv2 not importable")
    if not (vendor_ref.available() or os.path.isdir("/root/reference/karios/matcher")):
        pytest.skip("no reference files")
    logging.getLogger("karios").setLevel(logging.ERROR)
    ref_t, mon_t = synth.make_pair(360, 500, seed=9)
    to_np = lambda t: t.view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    ref, mon = to_np(ref_t), to_np(mon_t)
    df, secs, how = ref_run.run_pair(mon, ref, None, maxCorners=300, tile_size=320)
    tiles, total = P.match_scene(mon, ref, None, O.KLTConfiguration(maxCorners=300, tile_size=320))
    assert total == len(df) > 300 and "unmodified" in how and secs > 0
    for col, key in (("x0", "x0"), ("y0", "y0"), ("dx", "dx"), ("dy", "dy"), ("score", "score"), ("zncc_score", "zncc")):
        want = np.concatenate([t[key] for t in tiles])
        got = df[col].to_numpy()
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.allclose(got, want, rtol=0, atol=1e-12, equal_nan=True), col


def test_to_uint8_is_an_integer_division():
    """The arithmetic _to_uint8 of the 4-pixel Laplacian kernel (kr_prep.cu: lap4_body, ARITH):
    ((v - mn) / (mx - mn) * 255).astype(uint8) in float64 == (v - mn) * 255 // (mx - mn), and for
    ranges >= 256 == ((v - mn) * ceil(2^32 * 255 / R)) >> 32 with a multiplier below 2^32.  A sample
    of ranges here (all of them: tools/verify_to_uint8_integer.py), unsigned and signed."""
    rng = np.random.default_rng(5)
    ranges = sorted(set([1, 2, 3, 5, 15, 17, 51, 85, 254, 255, 256, 257, 510, 765, 1020, 3000, 4095, 4096,
                         32767, 32768, 65025, 65534, 65535] + [int(r) for r in rng.integers(1, 65536, 300)]))
    for R in ranges:
        for mn in {0, int(rng.integers(0, 65536 - R)), 65535 - R}:
            v = np.arange(mn, mn + R + 1, dtype=np.uint16)
            want = ((v - float(mn)) / (float(mn + R) - float(mn)) * 255).astype(np.uint8)
            n = v.astype(np.int64) - mn
            assert np.array_equal(want, (n * 255) // R), (R, mn)
            assert np.array_equal(want, O.to_uint8(v)), (R, mn)
            if R >= 256:
                M = -(-(255 << 32) // R)
                assert M < 2 ** 32 and np.array_equal(want, (n * M) >> 32), (R, mn)
    for R in (300, 5000, 40000, 65535):                                   # int16 rasters
        mn = -32768 if R == 65535 else int(rng.integers(-32768, 32767 - R))
        v = np.arange(mn, mn + R + 1, dtype=np.int32).astype(np.int16)
        want = ((v - float(mn)) / (float(mn + R) - float(mn)) * 255).astype(np.uint8)
        n = v.astype(np.int64) - mn
        assert np.array_equal(want, (n * (-(-(255 << 32) // R))) >> 32)
