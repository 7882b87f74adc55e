"""CPU oracle for the KARIOS KLT matching hot path (numpy + oracle/klt_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by karios_b200/.

Mirrors the reference call structure so that parity tests read like the
reference's own:
  to_uint8            <- karios/matcher/klt.py:42-49   (_to_uint8)
  laplacian           <- klt.py:433-434                 (cv2.Laplacian(u8, CV_8U, k))
  auto_mask           <- klt.py:268-276
  good_features       <- klt.py:120                     (cv2.goodFeaturesToTrack)
  pyr_lk              <- klt.py:134-140                 (cv2.calcOpticalFlowPyrLK)
  klt_tracker         <- klt.py:83-172
  match_tile / match  <- klt.py:198-349 (default + dict ksize, fixed polarity)
  zncc                <- karios/matcher/zncc_service.py:162-238, 45-126

Parity pinning: tests/golden/*.npz were produced by the UNMODIFIED reference
modules running against cv2 4.13.0 (oracle/make_golden.py); tests/test_oracle.py
checks every function here against them (and against cv2 directly when cv2 is
importable).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libklt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/klt_oracle.c (gcc, see oracle/Makefile)."""
    src = os.path.join(_HERE, "klt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libklt_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        c = ctypes
        vp, i32, i64, f64 = c.c_void_p, c.c_int, c.c_int64, c.c_double
        L.orc_minmax_u16.argtypes = [vp, i64, vp, vp]
        L.orc_minmax_u16.restype = None
        L.orc_to_uint8_u16.argtypes = [vp, i64, f64, f64, vp]
        L.orc_to_uint8_u16.restype = None
        L.orc_auto_mask_u16.argtypes = [vp, vp, i64, i32, f64, i32, f64, vp]
        L.orc_auto_mask_u16.restype = i64
        L.orc_laplacian_u8.argtypes = [vp, i32, i32, i32, vp]
        L.orc_laplacian_u8.restype = i32
        L.orc_min_eigen_val.argtypes = [vp, i32, i32, i32, i32, vp]
        L.orc_min_eigen_val.restype = i32
        L.orc_select_corners.argtypes = [vp, vp, i32, i32, i32, f64, f64, vp, i64, vp, vp]
        L.orc_select_corners.restype = i64
        L.orc_pyr_down_u8.argtypes = [vp, i32, i32, vp]
        L.orc_pyr_down_u8.restype = None
        L.orc_pyr_lk.argtypes = [vp, vp, i32, i32, vp, i64, i32, i32, i32, f64, f64, i32, vp, vp, vp]
        L.orc_pyr_lk.restype = i32
        L.orc_zncc_u16.argtypes = [vp, i32, i32, vp, i32, i32, vp, vp, vp, vp, i64, vp]
        L.orc_zncc_u16.restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# --------------------------------------------------------------------------- a4
def to_uint8(arr: np.ndarray) -> np.ndarray:
    """klt.py:42-49.  uint8 is a no-op; uint16 uses the C path; anything else the
    literal numpy expression of the reference."""
    if arr.dtype == np.uint8:
        return arr
    if arr.dtype == np.uint16:
        a = _c(arr, np.uint16)
        mn, mx = ctypes.c_double(), ctypes.c_double()
        lib().orc_minmax_u16(_p(a), a.size, ctypes.byref(mn), ctypes.byref(mx))
        out = np.empty(a.shape, np.uint8)
        lib().orc_to_uint8_u16(_p(a), a.size, mn.value, mx.value, _p(out))
        return out
    arr_min, arr_max = float(np.nanmin(arr)), float(np.nanmax(arr))
    if arr_max > arr_min:
        return ((arr - arr_min) / (arr_max - arr_min) * 255).astype(np.uint8)
    return np.zeros_like(arr, dtype=np.uint8)


def to_uint8_lut(mn: float, mx: float) -> np.ndarray:
    """The 65 536-entry table the uint16 path is equivalent to."""
    v = np.arange(65536, dtype=np.uint16)
    if not mx > mn:
        return np.zeros(65536, np.uint8)
    with np.errstate(invalid="ignore"):
        q = (v - float(mn)) / (float(mx) - float(mn)) * 255
    out = np.zeros(65536, np.uint8)
    ok = (q >= 0) & (q < 256)
    out[ok] = q[ok].astype(np.uint8)
    return out


# --------------------------------------------------------------------------- a3
def auto_mask(mon: np.ndarray, ref: np.ndarray, nd_mon=None, nd_ref=None):
    """klt.py:268-276 -> (uint8 mask, valid pixel count)."""
    if mon.dtype == np.uint16 and ref.dtype == np.uint16:
        m, r = _c(mon, np.uint16), _c(ref, np.uint16)
        mask = np.empty(m.shape, np.uint8)
        cnt = lib().orc_auto_mask_u16(_p(m), _p(r), m.size,
                                      int(nd_mon is not None), float(nd_mon or 0),
                                      int(nd_ref is not None), float(nd_ref or 0), _p(mask))
        return mask, int(cnt)
    mask = (mon != 0) & (ref != 0) & np.isfinite(ref) & np.isfinite(mon)
    if nd_mon is not None:
        mask &= mon != nd_mon
    if nd_ref is not None:
        mask &= ref != nd_ref
    mask = mask.astype(np.uint8)
    return mask, int(np.count_nonzero(mask))


# --------------------------------------------------------------------------- a6
def laplacian(u8: np.ndarray, ksize: int) -> np.ndarray:
    """cv2.Laplacian(u8, cv2.CV_8U, ksize=ksize) (klt.py:433-434)."""
    src = _c(u8, np.uint8)
    h, w = src.shape
    dst = np.empty_like(src)
    rc = lib().orc_laplacian_u8(_p(src), w, h, int(ksize), _p(dst))
    if rc:
        raise ValueError(f"unsupported Laplacian ksize {ksize}")
    return dst


# --------------------------------------------------------------------------- a7
def cv_tail_start(w: int) -> int:
    """First column of the non-FMA Sobel row-filter tail in the cv2 4.13 AVX-512
    build."""
    return 32 * (w // 32)


def min_eigen_val(u8: np.ndarray, block: int, tail_start: Optional[int] = None) -> np.ndarray:
    src = _c(u8, np.uint8)
    h, w = src.shape
    eig = np.empty((h, w), np.float32)
    ts = cv_tail_start(w) if tail_start is None else tail_start
    rc = lib().orc_min_eigen_val(_p(src), w, h, int(block), int(ts), _p(eig))
    if rc:
        raise MemoryError("orc_min_eigen_val")
    return eig


def select_corners(eig: np.ndarray, mask, max_corners: int, quality: float, min_distance: float):
    e = _c(eig, np.float32)
    h, w = e.shape
    m = None if mask is None else _c(mask, np.uint8)
    cap = (w * h) if max_corners <= 0 else max_corners
    cap = min(cap, max(1, (w * h)))
    out = np.empty((cap, 2), np.float32)
    ncand = ctypes.c_int64()
    mval = ctypes.c_float()
    n = lib().orc_select_corners(_p(e), None if m is None else _p(m), w, h, int(max_corners),
                                 float(quality), float(min_distance), _p(out), cap,
                                 ctypes.byref(ncand), ctypes.byref(mval))
    if n < 0:
        raise MemoryError("orc_select_corners")
    return out[:n].copy(), int(ncand.value), float(mval.value)


def good_features(u8, mask, max_corners, quality, min_distance, block, tail_start=None):
    """cv2.goodFeaturesToTrack(...) -> [N,1,2] float32 or None (klt.py:120)."""
    eig = min_eigen_val(u8, block, tail_start)
    pts, _, _ = select_corners(eig, mask, max_corners, quality, min_distance)
    if len(pts) == 0:
        return None
    return pts.reshape(-1, 1, 2)


# --------------------------------------------------------------------------- a8
def pyr_down(u8: np.ndarray) -> np.ndarray:
    src = _c(u8, np.uint8)
    h, w = src.shape
    dst = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyr_down_u8(_p(src), w, h, _p(dst))
    return dst


def pyr_lk(prev, nxt, p0, win=25, max_level=1, max_count=30, eps=0.03, min_eig=1e-4,
           acc_mode=0):
    """cv2.calcOpticalFlowPyrLK(prev, next, p0, None, winSize=(win,win), maxLevel,
    criteria=(EPS|COUNT, max_count, eps)) -> (p1 [N,1,2], status [N,1] u8, err [N,1])."""
    a, b = _c(prev, np.uint8), _c(nxt, np.uint8)
    h, w = a.shape
    pts = _c(np.asarray(p0).reshape(-1, 2), np.float32)
    n = len(pts)
    out = np.empty((n, 2), np.float32)
    st = np.empty(n, np.uint8)
    err = np.empty(n, np.float32)
    lib().orc_pyr_lk(_p(a), _p(b), w, h, _p(pts), n, int(win), int(max_level), int(max_count),
                     float(eps), float(min_eig), int(acc_mode), _p(out), _p(st), _p(err))
    return out.reshape(-1, 1, 2), st.reshape(-1, 1), err.reshape(-1, 1)


# --------------------------------------------------------------------------- a13
@dataclass
class KLTConfiguration:
    """Field-for-field mirror of karios/core/configuration.py:36-50."""
    minDistance: int = 10
    blocksize: int = 15
    maxCorners: int = 20000
    matching_winsize: int = 25
    qualityLevel: float = 0.1
    xStart: int = 0
    tile_size: int = 20000
    laplacian_kernel_size: object = 7
    outliers_filtering: bool = False
    laplacian_invert_polarity: object = False


def filter_outliers(x0, y0, x1, y1, score):
    """klt.py:52-71."""
    dx = x1 - x0
    dy = y1 - y0
    while True:
        ind = ((np.abs(dx - dx.mean()) < 3 * dx.std()) & (np.abs(dy - dy.mean()) < 3 * dy.std())
               & (np.abs(dx - dx.mean()) < 20) & (np.abs(dy - dy.mean()) < 20))
        if ind.sum() == len(dx):
            break
        dx, dy, x0, x1, y0, y1, score = dx[ind], dy[ind], x0[ind], x1[ind], y0[ind], y1[ind], score[ind]
    return x0, y0, x1, y1, score


def klt_tracker(ref_data, image_data, mask, conf, p0=None, acc_mode=0):
    """klt.py:83-172 -> (dict of float32 columns x0,y0,dx,dy,score, Ninit) or None."""
    if p0 is None:
        p0 = good_features(ref_data, mask, conf.maxCorners, conf.qualityLevel, conf.minDistance,
                           conf.blocksize)
    if p0 is None:
        return None
    w = conf.matching_winsize
    p1, _, _ = pyr_lk(ref_data, image_data, p0, win=w, acc_mode=acc_mode)
    p0r, _, _ = pyr_lk(image_data, ref_data, p1, win=w, acc_mode=acc_mode)
    d = np.abs(p0 - p0r).reshape(-1, 2).max(-1)
    st = d < 0.1
    ninit = len(p0)
    p0k, p1k, dk = p0[st], p1[st], d[st]
    score = 1 - dk / 0.1
    x0, y0 = p0k[:, 0, 0], p0k[:, 0, 1]
    x1, y1 = p1k[:, 0, 0], p1k[:, 0, 1]
    if conf.outliers_filtering and len(x0):
        x0, y0, x1, y1, score = filter_outliers(x0, y0, x1, y1, score)
    cols = {"x0": x0, "y0": y0, "dx": x1 - x0, "dy": y1 - y0, "score": score.astype(np.float32)}
    return cols, ninit


def _ksizes(conf):
    k = conf.laplacian_kernel_size
    if isinstance(k, dict):
        return k.get("mon", k.get("ref", 1)), k.get("ref", k.get("mon", 1))
    return k, k


LAPLACIAN_AUTO_CANDIDATES = [3, 5, 7, 9, 11]        # klt.py:39


def auto_ksize_search(mon_u8, ref_box, mask_box, conf, acc_mode=0):
    """KLT._match_tile_auto_ksize (klt.py:465-545): every (mon_ksize, ref_ksize) pair of
    LAPLACIAN_AUTO_CANDIDATES, corners once per reference kernel size, winner = highest
    inlier ratio len(points) / Ninit, the first maximum wins (strict >, product order:
    mon outer, ref inner).  -> ((cols, ninit) | None, scores dict, (mk, rk) | None)."""
    import itertools
    ref_u8 = to_uint8(ref_box)
    mon_l = {k: laplacian(mon_u8, k) for k in LAPLACIAN_AUTO_CANDIDATES}
    ref_l = {k: laplacian(ref_u8, k) for k in LAPLACIAN_AUTO_CANDIDATES}
    p0s = {k: good_features(lap, mask_box, conf.maxCorners, conf.qualityLevel, conf.minDistance,
                            conf.blocksize) for k, lap in ref_l.items()}
    scores, best, best_ratio, best_k = {}, None, -1.0, None
    for mk, rk in itertools.product(LAPLACIAN_AUTO_CANDIDATES, repeat=2):
        p0 = p0s[rk]
        res = None if p0 is None else klt_tracker(ref_l[rk], mon_l[mk], mask_box, conf, p0=p0,
                                                  acc_mode=acc_mode)
        ratio = 0.0
        if res is not None:
            cols, ninit = res
            ratio = len(cols["x0"]) / ninit if ninit > 0 else 0.0
        scores[(mk, rk)] = ratio
        if res is not None and ratio > best_ratio:
            best, best_ratio, best_k = res, ratio, (mk, rk)
    return best, scores, best_k


def track_once(mon_box, ref_box, mask_box, conf, invert_mon, acc_mode=0):
    """KLT._laplacian_track_once (klt.py:407-436) -> ((cols, ninit) | None, (mk, rk) | None)."""
    mon_u8 = to_uint8(mon_box)
    if invert_mon:
        mon_u8 = 255 - mon_u8
    if conf.laplacian_kernel_size == "auto":
        res, _, best_k = auto_ksize_search(mon_u8, ref_box, mask_box, conf, acc_mode)
        return res, best_k
    mk, rk = _ksizes(conf)
    res = klt_tracker(laplacian(to_uint8(ref_box), rk), laplacian(mon_u8, mk), mask_box, conf,
                      acc_mode=acc_mode)
    return res, (mk, rk)


def match_tile(mon_box, ref_box, mask_box, conf, x_off=0, y_off=0, nd_mon=None, nd_ref=None,
               acc_mode=0):
    """klt.py:236-349, every kernel-size / polarity mode.  Returns the dict of columns
    sorted by (x0, y0) with tile offsets applied (plus ninit, ksize, polarity), or None."""
    if mask_box is None:
        mask_box, valid = auto_mask(mon_box, ref_box, nd_mon, nd_ref)
    else:
        valid = int(np.count_nonzero(mask_box > 0))
    if valid == 0:
        return None
    polarity = None
    if conf.laplacian_invert_polarity == "auto":
        # klt.py:286-296, _select_best_polarity :438-463: stable sort on the ratio, the
        # normal polarity first on ties
        cands = []
        for label, inv in (("normal", False), ("inverted", True)):
            r, ks_ = track_once(mon_box, ref_box, mask_box, conf, inv, acc_mode)
            if r is None:
                continue
            ratio = len(r[0]["x0"]) / r[1] if r[1] > 0 else 0.0
            cands.append((label, ratio, r, ks_))
        if not cands:
            return None
        cands.sort(key=lambda c: c[1], reverse=True)
        polarity, _, res, ks = cands[0]
    else:
        res, ks = track_once(mon_box, ref_box, mask_box, conf, bool(conf.laplacian_invert_polarity),
                             acc_mode)
    if res is None:
        return None
    cols, ninit = res
    cols = dict(cols)
    cols["x0"] = cols["x0"] + np.float32(x_off)
    cols["y0"] = cols["y0"] + np.float32(y_off)
    order = np.lexsort((cols["y0"], cols["x0"]))
    cols = {k: v[order] for k, v in cols.items()}
    cols["ninit"] = ninit
    cols["ksize"] = ks
    cols["polarity"] = polarity
    return cols


def tile_boxes(x_size, y_size, conf):
    """Tile enumeration of KLT.match (klt.py:221-249): x outer, y inner."""
    out = []
    for x_off in range(0, x_size, conf.tile_size):
        if x_off < conf.xStart:
            continue
        for y_off in range(0, y_size, conf.tile_size):
            xs = conf.tile_size if x_off + conf.tile_size < x_size else x_size - x_off
            ys = conf.tile_size if y_off + conf.tile_size < y_size else y_size - y_off
            out.append((x_off, y_off, xs, ys))
    return out


def match(mon, ref, mask, conf, nd_mon=None, nd_ref=None, acc_mode=0):
    """KLT.match over whole arrays: list of per-tile column dicts."""
    h, w = mon.shape
    res = []
    for (xo, yo, xs, ys) in tile_boxes(w, h, conf):
        mb = None if mask is None else mask[yo:yo + ys, xo:xo + xs]
        r = match_tile(mon[yo:yo + ys, xo:xo + xs], ref[yo:yo + ys, xo:xo + xs], mb, conf, xo, yo,
                       nd_mon, nd_ref, acc_mode)
        if r is not None:
            res.append(r)
    return res


# --------------------------------------------------------------------------- a12
def zncc(x0, y0, dx, dy, monitored: np.ndarray, reference: np.ndarray) -> np.ndarray:
    """ZNCCService.compute_zncc over float32 columns (zncc_service.py:162-238):
    float64 scores, NaN where the reference yields NaN."""
    n = len(x0)
    out = np.empty(n, np.float64)
    if monitored.dtype == np.uint16 and reference.dtype == np.uint16:
        m, r = _c(monitored, np.uint16), _c(reference, np.uint16)
        a = [_c(v, np.float32) for v in (x0, y0, dx, dy)]
        lib().orc_zncc_u16(_p(r), r.shape[1], r.shape[0], _p(m), m.shape[1], m.shape[0],
                           _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), n, _p(out))
        return out
    # generic dtype: literal restatement of _compute_zncc / _zncc2
    for i in range(n):
        out[i] = _zncc_one(np.float32(x0[i]), np.float32(y0[i]), np.float32(dx[i]),
                           np.float32(dy[i]), monitored, reference)
    return out


def zncc2(img1, img2, u1, v1, u2, v2, n):
    """_zncc2 (zncc_service.py:45-126)."""
    if n < 0:
        raise ValueError("Window half-size n must be non-negative")
    h1, w1 = img1.shape
    h2, w2 = img2.shape
    if (u1 - n < 0 or u1 + n >= h1 or v1 - n < 0 or v1 + n >= w1
            or u2 - n < 0 or u2 + n >= h2 or v2 - n < 0 or v2 + n >= w2):
        raise IndexError("Patch window extends beyond image boundaries")
    p1 = img1[u1 - n:u1 + n + 1, v1 - n:v1 + n + 1]
    p2 = img2[u2 - n:u2 + n + 1, v2 - n:v2 + n + 1]
    s1, s2 = float(np.std(p1)), float(np.std(p2))
    if s1 == 0 or s2 == 0:
        return np.nan
    return float(np.mean(((p1 - np.mean(p1)) / s1) * ((p2 - np.mean(p2)) / s2)))


def _zncc_one(x0f, y0f, dxf, dyf, monitored, reference):
    m = 28
    x0, y0 = int(x0f), int(y0f)
    x1, y1 = round(x0f + dxf), round(y0f + dyf)
    if x0 - m < 0 or y0 - m < 0 or x1 - m < 0 or y1 - m < 0:
        return np.nan
    if (x0 >= reference.shape[1] - m or y0 >= reference.shape[0] - m
            or x1 >= monitored.shape[1] - m or y1 >= monitored.shape[0] - m):
        return np.nan
    cr = reference[y0 - m:y0 + m + 1, x0 - m:x0 + m + 1]
    cm = monitored[y1 - m:y1 + m + 1, x1 - m:x1 + m + 1]
    try:
        return zncc2(cr, cm, 28, 28, 28, 28, 21)
    except Exception:  # noqa: BLE001 - mirrors zncc_service.py:232-238
        return np.nan


# --------------------------------------------------------------------------- f1
def mutual_info_studholme(patch1, patch2, bins=32):
    """_mutual_info (mutual_info_service.py:32-63): (H(X) + H(Y)) / H(X, Y),
    natural logarithm, np.histogram2d with per-patch min/max edges."""
    hist_2d, _, _ = np.histogram2d(patch1.ravel(), patch2.ravel(), bins=bins)
    n = hist_2d.sum()
    if n == 0:
        return np.nan
    pxy = hist_2d / n
    px, py = pxy.sum(axis=1), pxy.sum(axis=0)
    hx = -np.sum(px[px > 0] * np.log(px[px > 0]))
    hy = -np.sum(py[py > 0] * np.log(py[py > 0]))
    hxy = -np.sum(pxy[pxy > 0] * np.log(pxy[pxy > 0]))
    if hxy == 0:
        return np.nan
    return float((hx + hy) / hxy)


def mutual_info_nmi(patch1, patch2, bins=32):
    """_mutual_information (zncc_service.py:129-151): 2 MI / (H(X) + H(Y)), log2."""
    p1 = patch1.ravel().astype(np.float64)
    p2 = patch2.ravel().astype(np.float64)
    joint_hist, _, _ = np.histogram2d(p1, p2, bins=bins)
    joint_prob = joint_hist / joint_hist.sum()

    def entropy(p):
        p = p[p > 0]
        return float(-np.sum(p * np.log2(p)))

    h_x, h_y = entropy(joint_prob.sum(axis=1)), entropy(joint_prob.sum(axis=0))
    h_xy = entropy(joint_prob.ravel())
    denom = h_x + h_y
    if denom == 0:
        return np.nan
    return float(2.0 * (h_x + h_y - h_xy) / denom)


def joint_histogram_int(patch1, patch2, bins=32):
    """The same joint histogram for INTEGER patches without floating point: the
    linspace edges are the exact rationals min + k (max - min) / bins, so the bin
    is ((v - min) * bins) // (max - min), capped at bins - 1; a constant patch
    falls into bin bins // 2 (edges min - 0.5 ... max + 0.5).  This is the form
    the CUDA kernel uses; tests check it against np.histogram2d."""
    def bin_of(p):
        v = p.ravel().astype(np.int64)
        mn, mx = int(v.min()), int(v.max())
        if mx == mn:
            return np.full(v.shape, bins // 2, np.int64)
        return np.minimum(bins - 1, ((v - mn) * bins) // (mx - mn))
    ka, kb = bin_of(patch1), bin_of(patch2)
    return np.bincount(ka * bins + kb, minlength=bins * bins).reshape(bins, bins)


def mutual_info(x0, y0, dx, dy, monitored: np.ndarray, reference: np.ndarray):
    """MutualInfoService.compute_mutual_info (mutual_info_service.py:73-138) and
    ZNCCService.compute_mi (zncc_service.py:240-287) over float32 columns ->
    (studholme [n], nmi [n]) float64; NaN where the reference returns NaN."""
    n = len(x0)
    st, mi = np.full(n, np.nan), np.full(n, np.nan)
    m = 28
    for i in range(n):
        x0f, y0f = np.float32(x0[i]), np.float32(y0[i])
        ax, ay = int(x0f), int(y0f)
        bx, by = round(x0f + np.float32(dx[i])), round(y0f + np.float32(dy[i]))
        if ax - m < 0 or ay - m < 0 or bx - m < 0 or by - m < 0:
            continue
        if (ax >= reference.shape[1] - m or ay >= reference.shape[0] - m
                or bx >= monitored.shape[1] - m or by >= monitored.shape[0] - m):
            continue
        cr = reference[ay - m:ay + m + 1, ax - m:ax + m + 1]
        cm = monitored[by - m:by + m + 1, bx - m:bx + m + 1]
        try:
            st[i] = mutual_info_studholme(cr, cm)
        except Exception:  # noqa: BLE001 - mutual_info_service.py:122-126
            pass
        try:
            mi[i] = mutual_info_nmi(cr, cm)
        except Exception:  # noqa: BLE001 - zncc_service.py:283-287
            pass
    return st, mi


# --------------------------------------------------------------------------- f2
def phase_cross_correlation_shift(reference_image, moving_image):
    """Whole-pixel part of skimage.registration.phase_cross_correlation(reference,
    moving) as LargeOffsetMatcher.match uses it (karios/matcher/large_offset.py:32-41,
    called with (mon, ref)).  scikit-image is pinned by the reference
    (environment.yml: scikit-image=0.24.*) but is not installed here nor vendored
    under /root/reference: this restates the published algorithm of
    skimage/registration/_phase_cross_correlation.py (0.24) for upsample_factor = 1,
    space = "real", normalization = "phase".  PARITY UNPINNED by reference vectors
    (the reference's tests mock this call, tests/test_large_offset_matcher.py:26-48);
    pinned by known-answer shifts in tests/test_oracle.py."""
    src_freq = np.fft.fftn(np.asarray(reference_image, np.float64))
    target_freq = np.fft.fftn(np.asarray(moving_image, np.float64))
    shape = src_freq.shape
    image_product = src_freq * target_freq.conj()
    eps = np.finfo(image_product.real.dtype).eps
    image_product /= np.maximum(np.abs(image_product), 100 * eps)
    cross_correlation = np.fft.ifftn(image_product)
    maxima = np.unravel_index(np.argmax(np.abs(cross_correlation)), cross_correlation.shape)
    midpoint = np.array([np.fix(axis_size / 2) for axis_size in shape])
    shift = np.stack(maxima).astype(np.float64, copy=False)
    shift[shift > midpoint] -= np.array(shape)[shift > midpoint]
    for dim in range(src_freq.ndim):
        if shape[dim] == 1:
            shift[dim] = 0
    return shift


def shift_image(img, y_off=0, x_off=0):
    """shift_image (karios/core/image.py:70-101)."""
    y_off, x_off = int(round(y_off)), int(round(x_off))
    h, w = img.shape
    out = np.zeros(img.shape, img.dtype)
    ys, xs = np.arange(h) + y_off, np.arange(w) + x_off
    yv, xv = (ys >= 0) & (ys < h), (xs >= 0) & (xs < w)
    out[np.ix_(yv, xv)] = img[np.ix_(ys[yv], xs[xv])]
    return out


# --------------------------------------------------------------------------- f3
def percentiles_2_98(arr):
    """_check_quality (karios/api/core.py:500-506)."""
    return np.nanpercentile(arr, [2, 98])


def count_valid_pixels(mon, mask=None):
    """analyze_accuracy (karios/api/core.py:285-290)."""
    masked = mon
    if mask is not None:
        masked = np.copy(mon)
        masked[mask == 0] = 0
    return int(np.count_nonzero(masked))


def filter_by_dn_values(x0, y0, mon, ref, no_values=None, mon_nd=None, ref_nd=None):
    """_filter_by_dn_values (karios/api/core.py:687-728) -> boolean keep mask."""
    xi, yi = np.asarray(x0).astype(int), np.asarray(y0).astype(int)
    rv, mv = ref[yi, xi], mon[yi, xi]
    keep = np.ones(len(xi), dtype=bool)
    for nv in no_values or []:
        keep &= ~((rv == nv) | (mv == nv))
    for nd, vals in ((ref_nd, rv), (mon_nd, mv)):
        if nd is not None:
            keep &= ~(vals == nd)
    return keep
