"""The reference's CPU path restated on top of the SAME third-party routines it
calls (OpenCV + NumPy), for the cpu_baseline / --impl reference timing legs.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by karios_b200/).

/root/reference does not travel to the GPU box, but opencv-python does (same
image), so this module re-states the few lines of glue of
karios/matcher/klt.py:42-49, 83-172, 236-349 and
karios/matcher/zncc_service.py:186-238, 45-126 around cv2.Laplacian,
cv2.goodFeaturesToTrack and cv2.calcOpticalFlowPyrLK.  Pinned against the
unmodified reference by tests/test_oracle.py::test_cv2_path_matches_golden.
"""
from __future__ import annotations

import numpy as np

try:
    import cv2
    HAVE_CV2 = True
except Exception:  # noqa: BLE001
    cv2 = None
    HAVE_CV2 = False


def to_uint8(arr):
    """klt.py:42-49."""
    if arr.dtype == np.uint8:
        return arr
    mn, mx = float(np.nanmin(arr)), float(np.nanmax(arr))
    if mx > mn:
        return ((arr - mn) / (mx - mn) * 255).astype(np.uint8)
    return np.zeros_like(arr, dtype=np.uint8)


def filter_outliers(x0, y0, x1, y1, score):
    """klt.py:52-71."""
    dx, dy = x1 - x0, y1 - y0
    while True:
        ind = ((np.abs(dx - dx.mean()) < 3 * dx.std()) & (np.abs(dy - dy.mean()) < 3 * dy.std())
               & (np.abs(dx - dx.mean()) < 20) & (np.abs(dy - dy.mean()) < 20))
        if ind.sum() == len(dx):
            break
        dx, dy, x0, x1, y0, y1, score = dx[ind], dy[ind], x0[ind], x1[ind], y0[ind], y1[ind], score[ind]
    return x0, y0, x1, y1, score


def klt_tracker(ref_data, image_data, mask, conf):
    """klt.py:83-172 without the DataFrame: dict of float32 columns, Ninit."""
    p0 = cv2.goodFeaturesToTrack(ref_data, mask=mask, maxCorners=conf.maxCorners,
                                 qualityLevel=conf.qualityLevel, minDistance=conf.minDistance,
                                 blockSize=conf.blocksize)
    if p0 is None:
        return None
    lk = dict(winSize=(conf.matching_winsize, conf.matching_winsize), maxLevel=1,
              criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.03))
    p1, _, _ = cv2.calcOpticalFlowPyrLK(ref_data, image_data, p0, None, **lk)
    p0r, _, _ = cv2.calcOpticalFlowPyrLK(image_data, ref_data, p1, None, **lk)
    d = abs(p0 - p0r).reshape(-1, 2).max(-1)
    st = d < 0.1
    ninit = len(p0)
    p0, p1, d = p0[st], p1[st], d[st]
    score = 1 - d / 0.1
    x0, y0 = p0[:, 0, 0], p0[:, 0, 1]
    x1, y1 = p1[:, 0, 0], p1[:, 0, 1]
    if conf.outliers_filtering and len(x0):
        x0, y0, x1, y1, score = filter_outliers(x0, y0, x1, y1, score)
    return {"x0": x0, "y0": y0, "dx": x1 - x0, "dy": y1 - y0, "score": score}, ninit


def match_tile(mon_box, ref_box, mask_box, conf, x_off=0, y_off=0):
    """klt.py:236-349, fixed kernel size and polarity."""
    if mask_box is None:
        mask_box = ((mon_box != 0) & (ref_box != 0) & np.isfinite(ref_box)
                    & np.isfinite(mon_box)).astype(np.uint8)
    if len(mask_box[mask_box > 0]) == 0:
        return None
    k = conf.laplacian_kernel_size
    mk, rk = (k.get("mon", k.get("ref", 1)), k.get("ref", k.get("mon", 1))) if isinstance(k, dict) else (k, k)
    mon_src = (255 - to_uint8(mon_box)) if conf.laplacian_invert_polarity is True else mon_box
    lap_mon = cv2.Laplacian(to_uint8(mon_src), cv2.CV_8U, ksize=mk)
    lap_ref = cv2.Laplacian(to_uint8(ref_box), cv2.CV_8U, ksize=rk)
    res = klt_tracker(lap_ref, lap_mon, mask_box, conf)
    if res is None:
        return None
    cols, ninit = res
    cols["x0"] = cols["x0"] + np.float32(x_off)
    cols["y0"] = cols["y0"] + np.float32(y_off)
    order = np.lexsort((cols["y0"], cols["x0"]))
    cols = {c: v[order] for c, v in cols.items()}
    cols["ninit"] = ninit
    return cols


def zncc_rows(x0, y0, dx, dy, monitored, reference):
    """zncc_service.py:186-238 + _zncc2 (:111-126): one Python iteration per row,
    like the reference's df.apply."""
    out = np.full(len(x0), np.nan)
    m = 28
    for i in range(len(x0)):
        ax, ay = int(x0[i]), int(y0[i])
        bx, by = round(x0[i] + dx[i]), round(y0[i] + dy[i])
        if ax - m < 0 or ay - m < 0 or bx - m < 0 or by - m < 0:
            continue
        if (ax >= reference.shape[1] - m or ay >= reference.shape[0] - m
                or bx >= monitored.shape[1] - m or by >= monitored.shape[0] - m):
            continue
        c1 = reference[ay - m:ay + m + 1, ax - m:ax + m + 1]
        c2 = monitored[by - m:by + m + 1, bx - m:bx + m + 1]
        p1, p2 = c1[7:50, 7:50], c2[7:50, 7:50]
        s1, s2 = np.std(p1), np.std(p2)
        if s1 == 0 or s2 == 0:
            continue
        out[i] = np.mean(((p1 - np.mean(p1)) / s1) * ((p2 - np.mean(p2)) / s2))
    return out


def match_scene(mon, ref, mask, conf, zncc_threshold=0.4):
    """KLT.match + the ZNCC part of _handle_klt_results (api/core.py:870-891) over
    whole arrays -> (list of per-tile column dicts with 'zncc', total rows)."""
    h, w = mon.shape
    tiles, total = [], 0
    for x_off in range(0, w, conf.tile_size):
        if x_off < conf.xStart:
            continue
        for y_off in range(0, h, conf.tile_size):
            xs = conf.tile_size if x_off + conf.tile_size < w else w - x_off
            ys = conf.tile_size if y_off + conf.tile_size < h else h - y_off
            mb = None if mask is None else mask[y_off:y_off + ys, x_off:x_off + xs]
            t = match_tile(mon[y_off:y_off + ys, x_off:x_off + xs],
                           ref[y_off:y_off + ys, x_off:x_off + xs], mb, conf, x_off, y_off)
            if t is None:
                continue
            z = np.full(len(t["x0"]), np.nan)
            sel = t["score"] >= np.float32(zncc_threshold)
            z[sel] = zncc_rows(t["x0"][sel], t["y0"][sel], t["dx"][sel], t["dy"][sel], mon, ref)
            t["zncc"] = z
            tiles.append(t)
            total += len(z)
    return tiles, total
