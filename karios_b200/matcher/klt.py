"""GPU drop-in for karios/matcher/klt.py: same names, arguments and results.

    klt_tracker(ref_data, image_data, mask, conf, p0=None) -> (DataFrame, Ninit) | None
    KLT(conf, gen_laplacian=False, out_dir=None).match(mon_img, ref_img, mask) -> Iterator[DataFrame]

Every array operation of the reference (np.nanmin/nanmax, _to_uint8,
cv2.Laplacian, cv2.goodFeaturesToTrack, cv2.calcOpticalFlowPyrLK, the
back-check, the (x0, y0) sort) runs in libkarios_b200.so on the current CUDA
device; this module only mirrors the control flow of klt.py:83-172 and
klt.py:198-545 (tiling, mask choice, polarity / kernel-size search).  There is
no CPU fallback.
"""
from __future__ import annotations

import itertools
import logging
import os
import threading
from collections import Counter
from typing import Iterator

import numpy as np
import torch
from pandas import DataFrame

from karios_b200 import _native as N
from karios_b200.core.image import device_full, remember_scores

logger = logging.getLogger(__name__)

LAPLACIAN_AUTO_CANDIDATES = [3, 5, 7, 9, 11]          # klt.py:40

_tls = threading.local()


def get_context(w: int, h: int, max_corners: int) -> N.Context:
    """Per-thread workspace, grown on demand (klt_tracker is called from a
    thread pool by the reference in auto-ksize mode, klt.py:526-527)."""
    ctx = getattr(_tls, "ctx", None)
    dev = torch.cuda.current_device() if torch.cuda.is_available() else None
    if ctx is None or not ctx.fits(w, h, max_corners) or ctx.device.index != dev:
        if ctx is not None:
            w, h = max(w, ctx.max_w), max(h, ctx.max_h)
            if ctx.max_corners <= 0 or max_corners <= 0:
                max_corners = 0
            else:
                max_corners = max(max_corners, ctx.max_corners)
            ctx.close()
        ctx = N.Context(w, h, max_corners)
        _tls.ctx = ctx
    return ctx


def _outlier_keep(dx, dy):
    """Rows kept by klt.py:52-71 (iterated 3-sigma / 20 px test on dx, dy), as a boolean
    mask over the input rows.  O(N) on <= maxCorners rows: stays in NumPy, float32 like the
    reference; the rows must be in the reference's (OpenCV) order so that the float32
    means and deviations are summed in the same order."""
    keep = np.arange(len(dx))
    while True:
        ind = ((np.abs(dx - dx.mean()) < 3 * dx.std()) & (np.abs(dy - dy.mean()) < 3 * dy.std())
               & (np.abs(dx - dx.mean()) < 20) & (np.abs(dy - dy.mean()) < 20))
        if ind.sum() == len(dx):
            break
        dx, dy, keep = dx[ind], dy[ind], keep[ind]
    return keep


def _frame(cols: np.ndarray, conf) -> DataFrame:
    """[5, n] float32 (x0, y0, dx, dy, score) -> the DataFrame of klt.py:166-168.  dx / dy are
    the kernel's float32 differences p1 - p0 of tile-local coordinates, i.e. the reference's
    `x1 - x0`; they are never rebuilt from offset coordinates."""
    x0, y0, dx, dy, score = (cols[i] for i in range(5))
    if conf.outliers_filtering and len(x0):
        keep = _outlier_keep(dx, dy)
        x0, y0, dx, dy, score = x0[keep], y0[keep], dx[keep], dy[keep], score[keep]
    return DataFrame.from_dict({"x0": x0, "y0": y0, "dx": dx, "dy": dy, "score": score})


def klt_tracker(ref_data, image_data, mask, conf, p0=None, tail_mode=N.KR_TAIL_AVX512):
    """Run KLT (klt.py:83-172).  ref_data / image_data: uint8 planes (NumPy or
    CUDA tensors); mask: uint8 region of interest or None; p0: optional
    [N,1,2] float32 features.  Returns (DataFrame x0,y0,dx,dy,score, Ninit) or
    None when no feature was found."""
    logger.info("Start tracking")
    dev = torch.device("cuda", torch.cuda.current_device())
    ref = N.to_device(ref_data, dev)
    mon = N.to_device(image_data, dev)
    if ref.dtype != torch.uint8 or mon.dtype != torch.uint8:
        raise N.KariosB200Error("klt_tracker expects uint8 planes (cv2 would reject others)")
    m = None if (mask is None or p0 is not None) else N.to_device(mask, dev)
    h, w = ref.shape
    n_p0 = 0
    p0_t = None
    if p0 is not None:
        p0_t = N.to_device(np.asarray(p0, np.float32).reshape(-1, 2), dev) \
            if not isinstance(p0, torch.Tensor) else p0.reshape(-1, 2).to(dev, torch.float32)
        n_p0 = p0_t.shape[0]
        if n_p0 == 0:
            return None
    cap = n_p0 if p0 is not None else int(conf.maxCorners)
    ctx = get_context(w, h, max(cap, 0) if p0 is None else max(cap, int(conf.maxCorners), 1))
    if cap <= 0:
        cap = ctx._cap_unlimited(w, h)
    rows = N.RowBuffers(cap, dev, with_zncc=False)
    kconf = N.make_conf(conf, ksize_mon=1, ksize_ref=1, invert_mon=False, tail_mode=tail_mode)
    n_init, n_kept = ctx.klt_track(ref, mon, m, kconf, rows, p0_t)
    if n_init == 0:
        logger.info("No features extracted")
        return None
    cols = rows.f32[:, :n_kept].cpu().numpy()
    logger.info("Tracking finished")
    return _frame(cols, conf), n_init


class KLT:
    # pylint: disable=too-few-public-methods
    """Class to execute the KLT matcher (klt.py:175-545) on the GPU."""

    def __init__(self, conf, gen_laplacian: bool = False, out_dir: str | None = None,
                 tail_mode: int = N.KR_TAIL_AVX512):
        self._conf = conf
        self._gen_laplacian = gen_laplacian
        self._out_dir = out_dir
        self._tail_mode = tail_mode
        self._auto_selected_ksizes: list[tuple[int, int]] = []
        self._selected_polarities: list[str] = []
        # "auto" kernel size: the whole search on the device (False: one klt_track per pair)
        self._batched_auto = True

    # ---------------------------------------------------------------- match
    def match(self, mon_img, ref_img, mask) -> Iterator[DataFrame]:
        """Run KLT on the image to monitor against a reference image, tile by
        tile, x outer / y inner (klt.py:198-234)."""
        logger.info("KLT...")
        logger.info("%s %s", mon_img.x_size, mon_img.y_size)
        for x_off in range(0, mon_img.x_size, self._conf.tile_size):
            if x_off < self._conf.xStart:
                continue
            for y_off in range(0, mon_img.y_size, self._conf.tile_size):
                points = self._match_tile(x_off, y_off, mon_img, ref_img, mask)
                if points is None:
                    continue
                yield points

    @staticmethod
    def _box(img, dev, x_off, y_off, x_size, y_size):
        """(tensor holding the tile, window inside it).  A host raster is uploaded once
        as a whole (core.image.device_full: shared with the scoring services, which need
        the whole raster anyway, zncc_service.py:209-210) and tiles are windows of it; a
        raster without `.array` is read tile by tile like the reference does (klt.py:251)."""
        if getattr(img, "device_array", None) is not None or hasattr(img, "array"):
            return device_full(img, dev), (x_off, y_off, x_size, y_size)
        box = N.to_device(img.read(1, x_off, y_off, x_size, y_size), dev)
        return box, (0, 0, x_size, y_size)

    def _match_tile(self, x_off, y_off, mon_img, ref_img, mask) -> DataFrame | None:
        conf = self._conf
        logger.info("Tile: %s %s (%s %s)", x_off, y_off, mon_img.x_size, mon_img.y_size)
        x_size = conf.tile_size if x_off + conf.tile_size < mon_img.x_size else mon_img.x_size - x_off
        y_size = conf.tile_size if y_off + conf.tile_size < mon_img.y_size else mon_img.y_size - y_off
        dev = torch.device("cuda", torch.cuda.current_device())
        mon_t, win = self._box(mon_img, dev, x_off, y_off, x_size, y_size)
        ref_t, win_r = self._box(ref_img, dev, x_off, y_off, x_size, y_size)
        if win_r != win or mon_t.shape != ref_t.shape or mon_t.dtype != ref_t.dtype:
            # mixed host / device or mixed dtype rasters: bring both to plain boxes
            mon_t = mon_t[win[1]:win[1] + y_size, win[0]:win[0] + x_size].contiguous()
            ref_t = ref_t[win_r[1]:win_r[1] + y_size, win_r[0]:win_r[0] + x_size].contiguous()
            if mon_t.dtype != ref_t.dtype:
                raise N.KariosB200Error("monitored and reference rasters must share a dtype")
            win = (0, 0, x_size, y_size)
        mask_t = None
        if mask:
            logger.info("Read mask at offset x %s, y %s, with tile size %s, %s", x_off, y_off,
                        x_size, y_size)
            whole = getattr(mask, "device_array", None) is not None or hasattr(mask, "array")
            if whole and win[0] == x_off and win[1] == y_off:
                mask_t = device_full(mask, dev, as_mask=True)      # uint8 0/1, same frame as the rasters
                if mask_t.shape != mon_t.shape:
                    raise N.KariosB200Error("mask and rasters differ in size")
            else:
                mb = device_full(mask, dev, as_mask=True)[y_off:y_off + y_size, x_off:x_off + x_size] \
                    if whole else N.to_device(mask.read(1, x_off, y_off, x_size, y_size), dev)
                mask_t = mb if mb.dtype == torch.uint8 else (mb > 0).to(torch.uint8)
                if win[0] or win[1]:       # rasters as whole frames, mask as a tile: pad to raster frame
                    full = torch.zeros(mon_t.shape, dtype=torch.uint8, device=dev)
                    full[y_off:y_off + y_size, x_off:x_off + x_size] = mask_t
                    mask_t = full
                elif not mask_t.is_contiguous():
                    mask_t = mask_t.contiguous()
        ctx = get_context(x_size, y_size, int(conf.maxCorners))
        cap = int(conf.maxCorners) if conf.maxCorners > 0 else ctx._cap_unlimited(x_size, y_size)
        rows = N.RowBuffers(cap, dev, with_zncc=True)
        nd = (mon_img.no_data_value, ref_img.no_data_value)

        polarity = conf.laplacian_invert_polarity
        if polarity == "auto":
            cands = []
            for label, inv in (("normal", False), ("inverted", True)):
                res = self._track_once(ctx, mon_t, ref_t, mask_t, win, rows, nd, inv)
                if res is not None:
                    cands.append((label,) + res)
            if not cands:
                logger.info("Auto polarity: no candidate produced a result")
                results = None
            else:
                # highest inlier ratio, first wins on ties (stable sort, klt.py:459)
                cands.sort(key=lambda c: (len(c[1]) / c[2]) if c[2] > 0 else 0.0, reverse=True)
                label, pts, ninit, ks = cands[0]
                self._selected_polarities.append(label)
                results = (pts, ninit, ks, label == "inverted")
        else:
            res = self._track_once(ctx, mon_t, ref_t, mask_t, win, rows, nd, bool(polarity))
            results = None if res is None else res + (bool(polarity),)

        if results is None:
            logger.warning("No result for tile %s %s (%s %s)", x_off, y_off, mon_img.x_size,
                           mon_img.y_size)
            return None
        points, ninit, (mk, rk), inverted = results
        if conf.laplacian_kernel_size == "auto":
            self._auto_selected_ksizes.append((mk, rk))
        if self._gen_laplacian:
            self._dump_laplacians(ctx, mon_t, ref_t, win, mk, rk, inverted, x_off, y_off)
        # klt.py:341-348.  "final": offsets added and rows sorted on the device;
        # "sorted": tile-local but already in (x0, y0) order; "raw": OpenCV order.
        state = points.attrs.pop("kr_state", "raw")
        zncc = points.attrs.pop("kr_zncc", None)
        if state != "final":
            points["x0"] = points["x0"] + x_off
            points["y0"] = points["y0"] + y_off
        if state == "raw":
            points.sort_values(by=["x0", "y0"], inplace=True)
        elif zncc is not None and polarity != "auto":
            remember_scores(mon_img, ref_img,
                            np.stack([points[c].to_numpy() for c in ("x0", "y0", "dx", "dy")]), zncc)
        logger.info("NbPoints(init/final): %s / %s", ninit, len(points.dx))
        return points

    # ------------------------------------------------------------- helpers
    def _track_once(self, ctx, mon_t, ref_t, mask_t, win, rows, nd, invert):
        """klt.py:407-436 -> (DataFrame, Ninit, (mon_ksize, ref_ksize)) | None."""
        conf = self._conf
        if conf.laplacian_kernel_size == "auto":
            return self._auto_ksize(ctx, mon_t, ref_t, mask_t, win, rows, nd, invert)
        kconf = N.make_conf(conf, invert_mon=invert, tail_mode=self._tail_mode)
        if conf.outliers_filtering:
            return self._track_once_unsorted(ctx, mon_t, ref_t, mask_t, win, rows, nd, invert, kconf)
        # ZNCC of every row in the same launch sequence (the caller asks for it next,
        # api/core.py:884-891; see core.image.remember_scores)
        fused = rows.zncc is not None and mon_t.shape == ref_t.shape
        if fused:
            kconf.compute_zncc, kconf.zncc_min_score = 1, -3.0e38
        st = ctx.match_tile(mon_t, ref_t, mask_t, win, kconf, rows, nd[0], nd[1])
        if mask_t is None and st.valid == 0:
            logger.info("-- No valid pixels, skipping this tile")
            return None
        if st.n_corners == 0:
            return None
        cols = rows.f32[:, : st.n_kept].cpu().numpy()
        df = _frame(cols, conf)
        # the library added the window offset and sorted by (x0, y0); a host box has
        # window offset 0, so the tile offset is still to be added by the caller
        df.attrs["kr_state"] = "final" if (win[0] or win[1]) else "sorted"
        if fused:
            df.attrs["kr_zncc"] = rows.zncc[: st.n_kept].cpu().numpy()
        return df, int(st.n_corners), (kconf.ksize_mon, kconf.ksize_ref)

    def _track_once_unsorted(self, ctx, mon_t, ref_t, mask_t, win, rows, nd, invert, kconf):
        """Fixed kernel size with outliers_filtering: the filter (klt.py:161-163) runs on the
        rows in OpenCV order and tile-local coordinates, before offsets and the (x0, y0) sort,
        so the stages are called one by one (kr_minmax_mask, kr_u8_laplacian, kr_klt_track)
        instead of the fused kr_match_tile, which sorts on the device."""
        conf = self._conf
        mk, rk = int(kconf.ksize_mon), int(kconf.ksize_ref)
        planes = self._planes(ctx, mon_t, ref_t, mask_t, win, nd, invert, None, pair=(mk, rk))
        if planes is None:
            if mask_t is None:
                logger.info("-- No valid pixels, skipping this tile")
            return None
        m, mon_l, ref_l = planes
        tconf = N.make_conf(conf, ksize_mon=1, ksize_ref=1, invert_mon=False, tail_mode=self._tail_mode)
        n_init, n_kept = ctx.klt_track(ref_l[rk], mon_l[mk], m, tconf, rows, None)
        if n_init == 0:
            return None
        df = _frame(rows.f32[:, :n_kept].cpu().numpy(), conf)
        df.attrs["kr_state"] = "raw"
        return df, int(n_init), (mk, rk)

    def _planes(self, ctx, mon_t, ref_t, mask_t, win, nd, invert, ksizes, pair=None):
        """Auto mask (when needed), min/max and Laplacians for a set of sizes."""
        x, y, w, h = win
        mon_b = mon_t[y:y + h, x:x + w]
        ref_b = ref_t[y:y + h, x:x + w]
        if mask_t is None:
            m = ctx.minmax_mask(mon_b, ref_b, nd[0], nd[1], want_mask=True)
            valid = ctx.read_stats().valid
        else:
            ctx.minmax_mask(mon_b, ref_b, nd[0], nd[1], want_mask=False)
            m = mask_t[y:y + h, x:x + w] if mask_t.shape != (h, w) else mask_t
            valid = int(torch.count_nonzero(m).item())
        if valid == 0:
            return None
        mon_l = {k: ctx.u8_laplacian(mon_b, k, invert=invert, slot=0) for k in (ksizes or [pair[0]])}
        ref_l = {k: ctx.u8_laplacian(ref_b, k, invert=False, slot=1) for k in (ksizes or [pair[1]])}
        return m, mon_l, ref_l

    def _auto_ksize_device(self, ctx, mon_t, ref_t, mask_t, win, rows, nd, invert):
        """The whole search as one device launch sequence (kr_auto_ksize): 10 Laplacians and
        pyramids, 5 corner sets, 25 LK round trips, ratios and winner on the device; one
        synchronisation.  -> (DataFrame, Ninit, (mk, rk)) | None | "redo" (host loop needed)."""
        conf = self._conf
        x, y, w, h = win
        mon_b, ref_b = mon_t[y:y + h, x:x + w], ref_t[y:y + h, x:x + w]
        m = None
        if mask_t is not None:
            m = mask_t[y:y + h, x:x + w] if mask_t.shape != (h, w) else mask_t
            if int(torch.count_nonzero(m).item()) == 0:
                logger.info("-- No valid pixels, skipping this tile")
                return None
        kconf = N.make_conf(conf, ksize_mon=1, ksize_ref=1, invert_mon=invert, tail_mode=self._tail_mode)
        r = ctx.auto_ksize(mon_b, ref_b, m, kconf, LAPLACIAN_AUTO_CANDIDATES, rows, nd[0], nd[1])
        if m is None and r.valid == 0:
            logger.info("-- No valid pixels, skipping this tile")
            return None
        if r.redo:
            return "redo"
        nk = len(LAPLACIAN_AUTO_CANDIDATES)
        for i, (mk, rk) in enumerate(itertools.product(LAPLACIAN_AUTO_CANDIDATES, repeat=2)):
            ni, kept = r.counts[i][0], r.counts[i][1]
            if ni > 0:
                logger.info("Auto laplacian: mon_ksize=%s ref_ksize=%s -> inlier ratio=%.3f (%d/%d)",
                            mk, rk, kept / ni, kept, ni)
        assert r.n_k == nk
        if r.best_mon == 0:
            return None
        df = _frame(rows.f32[:, : r.n_kept].cpu().numpy(), conf)
        df.attrs["kr_state"] = "raw"
        return df, int(r.n_init), (int(r.best_mon), int(r.best_ref))

    def _auto_ksize(self, ctx, mon_t, ref_t, mask_t, win, rows, nd, invert):
        """klt.py:465-545: every (mon_ksize, ref_ksize) pair, best inlier ratio,
        first maximum wins (strict >)."""
        conf = self._conf
        if conf.maxCorners > 0 and not conf.outliers_filtering and self._batched_auto:
            res = self._auto_ksize_device(ctx, mon_t, ref_t, mask_t, win, rows, nd, invert)
            if res != "redo":
                return res
        planes = self._planes(ctx, mon_t, ref_t, mask_t, win, nd, invert, LAPLACIAN_AUTO_CANDIDATES)
        if planes is None:
            logger.info("-- No valid pixels, skipping this tile")
            return None
        m, mon_l, ref_l = planes
        p0s = {k: ctx.good_features(lap, m, conf.maxCorners, conf.qualityLevel, conf.minDistance,
                                    conf.blocksize, self._tail_mode) for k, lap in ref_l.items()}
        best, best_ratio, best_k = None, -1.0, None
        kconf = N.make_conf(conf, ksize_mon=1, ksize_ref=1, invert_mon=False, tail_mode=self._tail_mode)
        for mk, rk in itertools.product(LAPLACIAN_AUTO_CANDIDATES, repeat=2):
            p0 = p0s[rk]
            if p0.shape[0] == 0:
                continue
            if rows.capacity < p0.shape[0]:
                rows = N.RowBuffers(p0.shape[0], p0.device, with_zncc=False)
            n_init, n_kept = ctx.klt_track(ref_l[rk], mon_l[mk], None, kconf, rows, p0)
            df = _frame(rows.f32[:, :n_kept].cpu().numpy(), conf)
            ratio = len(df) / n_init if n_init > 0 else 0.0
            logger.info("Auto laplacian: mon_ksize=%s ref_ksize=%s -> inlier ratio=%.3f (%d/%d)",
                        mk, rk, ratio, len(df), n_init)
            if ratio > best_ratio:
                best, best_ratio, best_k = (df, n_init), ratio, (mk, rk)
        if best is None:
            return None
        df, n_init = best
        df.attrs["kr_state"] = "raw"
        return df, n_init, best_k

    def _dump_laplacians(self, ctx, mon_t, ref_t, win, mk, rk, inverted, x_off, y_off):
        """klt.py:307-322 (debug dump; needs scikit-image like the reference)."""
        from skimage import io  # noqa: PLC0415
        x, y, w, h = win
        ctx.minmax_mask(mon_t[y:y + h, x:x + w], ref_t[y:y + h, x:x + w])
        lm = ctx.u8_laplacian(mon_t[y:y + h, x:x + w], mk, invert=inverted, slot=0).cpu().numpy()
        lr = ctx.u8_laplacian(ref_t[y:y + h, x:x + w], rk, slot=1).cpu().numpy()
        suffix = "_inv" if inverted else ""
        io.imsave(os.path.join(self._out_dir, f"mon_laplacian{suffix}_k{mk}_{x_off}_{y_off}_{w}_{h}.tif"), lm)
        io.imsave(os.path.join(self._out_dir, f"ref_laplacian_k{rk}_{x_off}_{y_off}_{w}_{h}.tif"), lr)

    @property
    def auto_selected_ksize(self):
        """Most common (mon_ksize, ref_ksize) over the auto-mode tiles (klt.py:351-356)."""
        if not self._auto_selected_ksizes:
            return None
        return Counter(self._auto_selected_ksizes).most_common(1)[0][0]

    @property
    def auto_selected_polarity(self):
        """Most common polarity over the tiles in 'auto' mode (klt.py:399-405)."""
        if not self._selected_polarities:
            return None
        return Counter(self._selected_polarities).most_common(1)[0][0]
