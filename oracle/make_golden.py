"""Generate tests/golden/*.npz by running the UNMODIFIED reference hot path.

Runs only in the build container (needs /root/reference and cv2): imports
karios/matcher/klt.py, karios/matcher/zncc_service.py and
karios/core/configuration.py through oracle/refimport.py
and records, for small synthetic pairs, the outputs of every reference call
site on the path (klt.py:42-49 _to_uint8, :433-434 cv2.Laplacian, :120
goodFeaturesToTrack, :134-140 calcOpticalFlowPyrLK, :83-172 klt_tracker,
:198-349 KLT.match, zncc_service.py:162-238 compute_zncc).  The fixtures pin
oracle/klt_oracle.c (tests/test_oracle.py) and the CUDA path (tests -m gpu).

    python oracle/make_golden.py        # rewrites tests/golden/
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != HERE]
sys.path.insert(0, ROOT)

import cv2  # noqa: E402
import pandas as pd  # noqa: E402
import torch  # noqa: E402

from karios_b200 import synth  # noqa: E402
from oracle import refimport  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np16(t):
    return t.view(torch.int16).numpy().view(np.uint16).copy()


def conf_of(cfg, **kw):
    base = dict(minDistance=10, blocksize=15, maxCorners=20000, matching_winsize=25,
                qualityLevel=0.1, xStart=0, tile_size=20000, laplacian_kernel_size=7,
                outliers_filtering=False, laplacian_invert_polarity=False)
    base.update(kw)
    return cfg.KLTConfiguration(**base)


def df_cols(df, prefix):
    return {f"{prefix}_{c}": df[c].to_numpy() for c in df.columns}


def case(name, h, w, seed, shift, conf_kw, mask_mode=None, zero_block=False, extra_lk=True,
         negate_mon=False, keep_eig=False):
    klt, zs, cfg = refimport.load()
    ref_t, mon_t = synth.make_pair(h, w, seed=seed, shift=shift)
    ref, mon = _np16(ref_t), _np16(mon_t)
    if negate_mon:                      # a sensor with inverted contrast (polarity test)
        mon = (5000 - mon.astype(np.int32)).astype(np.uint16)
    if zero_block:                      # exercise the auto mask (klt.py:268-273)
        mon[h // 3: h // 3 + 40, w // 4: w // 4 + 60] = 0
        ref[: 25, : 70] = 0
    mask = None
    if mask_mode == "user":
        mask = synth.make_mask(h, w, seed=7).numpy()
        mask[h // 2:, : w // 3] = 0
    conf = conf_of(cfg, **conf_kw)
    out = dict(ref=ref, mon=mon, h=h, w=w, seed=seed, shift=np.asarray(shift, np.float64))
    out["conf_json"] = np.array(repr(conf_kw))
    if mask is not None:
        out["mask"] = mask

    # --- whole-array stages (single tile view) -------------------------------
    k = conf.laplacian_kernel_size
    mk, rk = (k.get("mon"), k.get("ref")) if isinstance(k, dict) else (k, k)
    ref_u8, mon_u8 = klt._to_uint8(ref), klt._to_uint8(mon)
    out["ref_u8"], out["mon_u8"] = ref_u8, mon_u8
    mon_for_lap = (255 - mon_u8) if conf.laplacian_invert_polarity is True else mon_u8
    lap_ref = cv2.Laplacian(ref_u8, cv2.CV_8U, ksize=rk)
    lap_mon = cv2.Laplacian(mon_for_lap, cv2.CV_8U, ksize=mk)
    out["lap_ref"], out["lap_mon"] = lap_ref, lap_mon
    for kk in (1, 3, 5, 9, 11):
        out[f"lap_ref_k{kk}"] = cv2.Laplacian(ref_u8, cv2.CV_8U, ksize=kk)
    if mask is None:
        m = (mon != 0) & (ref != 0) & np.isfinite(ref) & np.isfinite(mon)
        mask_box = m.astype(np.uint8)
    else:
        mask_box = mask
    out["mask_box"] = mask_box
    eig = cv2.cornerMinEigenVal(lap_ref, conf.blocksize, ksize=3)
    if keep_eig:
        out["eig"] = eig
    out["eig_max"] = eig.max()
    out["eig_rows"] = eig[[0, 1, h // 2, h - 2, h - 1]]
    out["eig_cols"] = eig[:, [0, 1, w // 2, w - 5, w - 4, w - 2, w - 1]]
    p0 = cv2.goodFeaturesToTrack(lap_ref, mask=mask_box, maxCorners=conf.maxCorners,
                                 qualityLevel=conf.qualityLevel, minDistance=conf.minDistance,
                                 blockSize=conf.blocksize)
    out["p0"] = p0 if p0 is not None else np.zeros((0, 1, 2), np.float32)
    # a second parameter set on the same image (small maxCorners / other distance)
    p0b = cv2.goodFeaturesToTrack(lap_ref, mask=None, maxCorners=150, qualityLevel=0.05,
                                  minDistance=4, blockSize=7)
    out["p0_alt"] = p0b
    out["pyr_ref"] = cv2.pyrDown(lap_ref)
    if p0 is not None and extra_lk:
        wsz = conf.matching_winsize
        lk = dict(winSize=(wsz, wsz), maxLevel=1,
                  criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.03))
        p1, st, err = cv2.calcOpticalFlowPyrLK(lap_ref, lap_mon, p0, None, **lk)
        p0r, st2, err2 = cv2.calcOpticalFlowPyrLK(lap_mon, lap_ref, p1, None, **lk)
        out.update(lk_p1=p1, lk_st=st, lk_err=np.where(st == 1, err, 0),
                   lk_p0r=p0r, lk_st2=st2, lk_err2=np.where(st2 == 1, err2, 0))
        # points close to / beyond the border, to pin the status rules
        rng = np.random.default_rng(seed)
        pb = np.stack([rng.uniform(-30, w + 30, 300), rng.uniform(-30, h + 30, 300)], -1)
        pb = pb.astype(np.float32).reshape(-1, 1, 2)
        q1, qs, qe = cv2.calcOpticalFlowPyrLK(lap_ref, lap_mon, pb, None, **lk)
        out.update(lkb_p0=pb, lkb_p1=q1, lkb_st=qs, lkb_err=np.where(qs == 1, qe, 0))
    res = klt.klt_tracker(lap_ref, lap_mon, mask_box, conf)
    if res is not None:
        df, ninit = res
        out.update(df_cols(df, "trk"))
        out["trk_ninit"] = ninit

    # --- the reference KLT.match generator (tiling, offsets, sort) ------------
    mon_img, ref_img = refimport.ArrayImage(mon), refimport.ArrayImage(ref)
    mask_img = refimport.ArrayImage(mask) if mask is not None else None
    frames = list(klt.KLT(conf).match(mon_img, ref_img, mask_img))
    out["match_ntiles"] = len(frames)
    if frames:
        for i, f in enumerate(frames):
            out.update(df_cols(f, f"match{i}"))
        full = pd.concat(frames, ignore_index=True)
        # --- reference ZNCC on all rows (zncc_service.py:162-184) -------------
        z = zs.ZNCCService().compute_zncc(full, mon_img, ref_img)
        out["zncc"] = z.to_numpy(np.float64)
        out.update(df_cols(full, "all"))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()
                 if k in ("p0", "trk_x0", "zncc", "match_ntiles")})


def auto_case(name, h, w, seed, shift, conf_kw, negate_mon=False):
    """KLT.match with laplacian_kernel_size="auto" and / or laplacian_invert_polarity="auto"
    (klt.py:438-545, 286-304) through the unmodified reference: the per-tile frames, the
    kernel sizes and polarities it selected per tile, and -- for the first tile -- the
    inlier ratio of every (mon_ksize, ref_ksize) pair of both polarities."""
    klt, _, cfg = refimport.load()
    ref_t, mon_t = synth.make_pair(h, w, seed=seed, shift=shift)
    ref, mon = _np16(ref_t), _np16(mon_t)
    if negate_mon:
        mon = (5000 - mon.astype(np.int32)).astype(np.uint16)
    conf = conf_of(cfg, **conf_kw)
    out = dict(ref=ref, mon=mon, h=h, w=w, seed=seed, shift=np.asarray(shift, np.float64))
    out["conf_json"] = np.array(repr(conf_kw))
    mon_img, ref_img = refimport.ArrayImage(mon), refimport.ArrayImage(ref)
    k = klt.KLT(conf)
    frames = list(k.match(mon_img, ref_img, None))
    out["match_ntiles"] = len(frames)
    for i, f in enumerate(frames):
        out.update(df_cols(f, f"match{i}"))
    out["tile_ksizes"] = np.array(k._auto_selected_ksizes, np.int64).reshape(-1, 2)
    out["tile_polarities"] = np.array(k._selected_polarities)
    out["auto_selected_ksize"] = np.array(k.auto_selected_ksize if k.auto_selected_ksize else (0, 0), np.int64)
    out["auto_selected_polarity"] = np.array(k.auto_selected_polarity or "")
    if conf.laplacian_kernel_size == "auto":
        ts = conf.tile_size
        mb, rb = mon[:ts, :ts], ref[:ts, :ts]
        mask_box = ((mb != 0) & (rb != 0)).astype(np.uint8)
        for label, box in (("normal", mb), ("inverted", 255 - klt._to_uint8(mb))):
            _, scores, best = klt.KLT(conf)._match_tile_auto_ksize(box, rb, mask_box)
            out[f"scores_{label}"] = np.array([[a, b, r] for (a, b), r in scores.items()], np.float64)
            out[f"best_{label}"] = np.array(best if best else (0, 0), np.int64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "tiles", len(frames), "ksizes", out["tile_ksizes"].tolist(), "polarities",
          out["tile_polarities"].tolist(), "rows", [len(f) for f in frames])


def zncc_known_answers():
    """The reference's own ZNCC known-answer cases (tests/test_zncc_service.py,
    tests/test_zncc_zero_std_fix.py) evaluated through the unmodified _zncc2."""
    _, zs, _ = refimport.load()
    rng = np.random.default_rng(3)
    a = rng.integers(0, 65535, (57, 57)).astype(np.uint16)
    b = rng.integers(0, 65535, (57, 57)).astype(np.uint16)
    out = dict(a=a, b=b)
    out["z_ab"] = zs._zncc2(a, b, 28, 28, 28, 28, 21)
    out["z_aa"] = zs._zncc2(a, a, 28, 28, 28, 28, 21)
    out["z_anti"] = zs._zncc2(a, (65535 - a).astype(np.uint16), 28, 28, 28, 28, 21)
    flat = np.full((57, 57), 1234, np.uint16)
    out["z_flat"] = zs._zncc2(flat, b, 28, 28, 28, 28, 21)
    np.savez_compressed(os.path.join(OUT, "zncc_known.npz"), **out)
    print("zncc_known", {k: float(v) for k, v in out.items() if k.startswith("z_")})


def mi_golden():
    """Mutual-information scores through the UNMODIFIED reference:
    _mutual_info (mutual_info_service.py:32-63), _mutual_information
    (zncc_service.py:129-151) on the patch families of the reference's own
    known-answer tests (tests/test_mutual_info_service.py: identical, correlated,
    independent, one uniform, both uniform) in the raster dtypes the CUDA path
    takes, and MutualInfoService.compute_mutual_info / ZNCCService.compute_mi on
    the rows of a synthetic pair, border rows included."""
    _, zs, _ = refimport.load()
    mis = refimport.load_mutual_info()
    out = {}
    rng = np.random.default_rng(42)
    base = rng.random((57, 57))
    fam = {
        "ident": (np.arange(57 * 57, dtype=np.float64).reshape(57, 57),) * 2,
        "corr": (base, base + rng.random((57, 57)) * 0.1),
        "indep": (rng.random((57, 57)), rng.random((57, 57))),
        "unif1": (np.ones((57, 57)), rng.random((57, 57))),
        "unif2": (np.ones((57, 57)), np.ones((57, 57))),
    }
    for name, (a, b) in fam.items():
        for dt, scale in (("f32", 1.0), ("u16", 6000.0), ("u8", 255.0)):
            if dt == "f32":
                pa, pb = a.astype(np.float32), b.astype(np.float32)
            else:
                t = np.uint16 if dt == "u16" else np.uint8
                sa = scale / max(1.0, a.max())
                sb = scale / max(1.0, b.max())
                pa, pb = (a * sa).astype(t), (b * sb).astype(t)
            out[f"{name}_{dt}_a"], out[f"{name}_{dt}_b"] = pa, pb
            out[f"{name}_{dt}_studholme"] = mis._mutual_info(pa, pb)
            out[f"{name}_{dt}_nmi"] = zs._mutual_information(pa, pb)
    # service level, uint16 rasters
    ref_t, mon_t = synth.make_pair(300, 420, seed=11, shift=(0.30, -0.20))
    ref, mon = _np16(ref_t), _np16(mon_t)
    ref[100:170, 200:280] = 1500            # a flat area: one-bin chips -> NaN / 1.0 / 0.0
    mon[100:170, 200:280] = 1500
    mon[20:60, 300:330] = 0
    n = 160
    x0 = rng.uniform(0, 420, n).astype(np.float32)
    y0 = rng.uniform(0, 300, n).astype(np.float32)
    x0[:6] = [28.0, 27.9, 391.0, 392.0, 240.0, 238.5]
    y0[:6] = [28.0, 60.0, 271.0, 100.0, 135.0, 133.5]
    dx = rng.uniform(-1.5, 1.5, n).astype(np.float32)
    dy = rng.uniform(-1.5, 1.5, n).astype(np.float32)
    dx[4:6] = 0.25
    dy[4:6] = -0.25
    df = pd.DataFrame({"x0": x0, "y0": y0, "dx": dx, "dy": dy})
    mon_img, ref_img = refimport.ArrayImage(mon), refimport.ArrayImage(ref)
    out.update(svc_ref=ref, svc_mon=mon, svc_x0=x0, svc_y0=y0, svc_dx=dx, svc_dy=dy)
    out["svc_studholme"] = mis.MutualInfoService().compute_mutual_info(df, mon_img, ref_img).to_numpy(np.float64)
    out["svc_nmi"] = zs.ZNCCService().compute_mi(df, mon_img, ref_img).to_numpy(np.float64)
    # float32 rasters with a NaN and an Inf pixel (np.histogram2d raises -> NaN)
    reff, monf = ref.astype(np.float32) / 7, mon.astype(np.float32) / 7
    reff[50, 50] = np.nan
    monf[250, 100] = np.inf
    out.update(svc_ref_f32=reff, svc_mon_f32=monf)
    import logging
    logging.disable(logging.CRITICAL)
    out["svc_studholme_f32"] = mis.MutualInfoService().compute_mutual_info(
        df, refimport.ArrayImage(monf), refimport.ArrayImage(reff)).to_numpy(np.float64)
    out["svc_nmi_f32"] = zs.ZNCCService().compute_mi(
        df, refimport.ArrayImage(monf), refimport.ArrayImage(reff)).to_numpy(np.float64)
    logging.disable(logging.NOTSET)
    np.savez_compressed(os.path.join(OUT, "mi_known.npz"), **out)
    print("mi_known", {k: float(v) for k, v in out.items() if k.endswith("u16_studholme") or k.endswith("u16_nmi")},
          "svc NaN", int(np.isnan(out["svc_studholme"]).sum()), int(np.isnan(out["svc_nmi"]).sum()),
          int(np.isnan(out["svc_studholme_f32"]).sum()))


def scene_golden():
    """shift_image (karios/core/image.py:70-101) through the unmodified reference, for
    the offset signs / magnitudes of its four branches and the rounding of float offsets."""
    im = refimport.load_core_image()
    rng = np.random.default_rng(9)
    a16 = rng.integers(1, 60000, (37, 53)).astype(np.uint16)
    a8 = rng.integers(1, 255, (20, 31)).astype(np.uint8)
    af = rng.random((16, 18)).astype(np.float32)
    out = dict(a16=a16, a8=a8, af=af)
    offs = [(0, 0), (3, 0), (0, -4), (-5, 7), (6, -2), (2.5, -3.5), (1.4999, 0.5), (-36, 52), (40, 0)]
    out["offsets"] = np.array(offs, np.float64)
    for k, (yo, xo) in enumerate(offs):
        out[f"s16_{k}"] = im.shift_image(a16, y_off=yo, x_off=xo)
        out[f"s8_{k}"] = im.shift_image(a8, y_off=yo, x_off=xo)
        out[f"sf_{k}"] = im.shift_image(af, y_off=yo, x_off=xo)
    np.savez_compressed(os.path.join(OUT, "scene_ops.npz"), **out)
    print("scene_ops", len(offs), "offsets")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    cv2.setNumThreads(1)
    # W mod 32 = 4 like an S2 band; default CLI config; auto mask with zero blocks
    case("basic", 300, 420, 11, (0.30, -0.20), {}, zero_block=True, keep_eig=True)
    # user mask, 4 tiles with remainders, ksize 5, capped corners
    case("tiles_mask", 430, 500, 12, (-0.45, 0.35),
         dict(tile_size=256, laplacian_kernel_size=5, maxCorners=300, minDistance=7),
         mask_mode="user")
    # dict ksize, inverted polarity, small window, outlier filter on
    case("dict_inv", 280, 333, 13, (0.8, 0.6),
         dict(laplacian_kernel_size={"mon": 5, "ref": 7}, laplacian_invert_polarity=True,
              matching_winsize=15, outliers_filtering=True, qualityLevel=0.02, minDistance=5,
              blocksize=7, maxCorners=2000), negate_mon=True)
    # automatic kernel-size and polarity search (klt.py:438-545): two tiles, a contrast-
    # inverted monitored image (the inverted polarity must win), outlier filter on
    auto_case("auto_modes", 200, 330, 14, (0.35, -0.25),
              dict(laplacian_kernel_size="auto", laplacian_invert_polarity="auto", maxCorners=250,
                   tile_size=200, outliers_filtering=True), negate_mon=True)
    # automatic kernel size with a fixed polarity, single tile
    auto_case("auto_ksize", 180, 260, 15, (-0.6, 0.4),
              dict(laplacian_kernel_size="auto", maxCorners=300, minDistance=6))
    # both searches, no outlier filter (the configuration the on-device search covers), contrast-
    # inverted monitored image, user-sized tiles with a remainder
    auto_case("auto_device", 210, 300, 16, (0.25, 0.45),
              dict(laplacian_kernel_size="auto", laplacian_invert_polarity="auto", maxCorners=200,
                   tile_size=180, minDistance=8), negate_mon=True)
    zncc_known_answers()
    mi_golden()
    scene_golden()
