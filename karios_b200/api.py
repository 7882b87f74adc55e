"""Scene-level glue of the hot path: what KariosAPI._compute_matches +
_handle_klt_results do around the matcher (karios/api/core.py:845-921), minus
CSV: per tile KLT._match_tile, then ZNCC (and, with_mi, the two mutual-information
scores) for rows with score >= confidence_threshold, all in one stream-ordered
launch sequence per tile.

    match_pair(mon, ref, mask, conf) -> DataFrame[x0, y0, dx, dy, score, zncc_score]

`mon` / `ref` / `mask` may be NumPy arrays, host (pinned) tensors or CUDA
tensors.  ScenePipeline keeps one set of device buffers and overlaps the upload
of the next pair with the matching of the current one.
"""
from __future__ import annotations

import threading
import time

import numpy as np
import torch
from pandas import DataFrame

from karios_b200 import _native as N


def tile_windows(x_size: int, y_size: int, conf):
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


class SceneMatcher:
    """Workspace + row buffers for scenes up to (h, w); reusable across pairs."""

    def __init__(self, h: int, w: int, conf, confidence_threshold: float = 0.4,
                 tail_mode: int = N.KR_TAIL_AVX512, device=None, depth: int = 1,
                 with_mi: bool = False):
        if isinstance(conf.laplacian_kernel_size, str) or conf.laplacian_invert_polarity == "auto":
            raise N.KariosB200Error("SceneMatcher handles fixed kernel size / polarity; "
                                    "use matcher.klt.KLT for the 'auto' searches")
        if conf.outliers_filtering:
            # klt.py:161-163 filters the rows of a tile on the host, in OpenCV order, before the
            # frame is built and before ZNCC; the fused device sequence here has no such step
            raise N.KariosB200Error("SceneMatcher does not apply outliers_filtering; "
                                    "use matcher.klt.KLT (it filters per tile like the reference)")
        self.conf = conf
        self.h, self.w = h, w
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        tw, th = min(w, conf.tile_size), min(h, conf.tile_size)
        self.ctx = N.Context(tw, th, int(conf.maxCorners), self.device)
        cap = int(conf.maxCorners) if conf.maxCorners > 0 else self.ctx._cap_unlimited(tw, th)
        self.with_mi = bool(with_mi)
        self.rows = N.RowBuffers(cap, self.device, with_zncc=True, with_mi=self.with_mi)
        self.kconf = N.make_conf(conf, tail_mode=tail_mode, compute_zncc=True,
                                 zncc_min_score=confidence_threshold, compute_mi=self.with_mi)
        self.windows = tile_windows(w, h, conf)
        # extra (context, rows, stream) slots for match_many: tiles of successive
        # pairs are independent, so `depth` of them are in flight at once and the
        # latency-bound stages of one (selection, NMS, sorts) overlap the
        # bandwidth / issue-bound stages of the others
        self.depth = max(1, int(depth))
        # hook of the multi-GPU exchange: called with (arena, unit streams) right after the last
        # unit of a match_many call has been enqueued (sharding.exchange on a side stream)
        self.on_last_enqueued = None
        self._slots = [(self.ctx, self.rows, torch.cuda.Stream(device=self.device))]
        for _ in range(self.depth - 1):
            self._slots.append((N.Context(tw, th, int(conf.maxCorners), self.device),
                                N.RowBuffers(cap, self.device, with_zncc=True, with_mi=self.with_mi),
                                torch.cuda.Stream(device=self.device)))

    def _check_device(self):
        """Kernels are enqueued on the current stream of the current device: it has to be the
        device the workspace lives on."""
        if torch.cuda.current_device() != (self.device.index or 0):
            raise N.KariosB200Error(f"current CUDA device {torch.cuda.current_device()} is not the matcher's "
                                    f"{self.device}; wrap the call in torch.cuda.device(...)")

    def match_device(self, mon: torch.Tensor, ref: torch.Tensor, mask=None, nodata=(None, None),
                     collect=True):
        """Both rasters resident in HBM.  -> (list of per-tile [6, n] float64-free
        device tensors (x0,y0,dx,dy,score) + zncc, total rows)."""
        tiles, total = [], 0
        self._check_device()
        for win in self.windows:
            st = self.ctx.match_tile(mon, ref, mask, win, self.kconf, self.rows, nodata[0], nodata[1])
            n = int(st.n_kept)
            if (mask is None and st.valid == 0) or st.n_corners == 0:
                continue
            total += n
            if collect:
                t = (self.rows.f32[:, :n].clone(), self.rows.zncc[:n].clone())
                if self.with_mi:
                    t = t + (self.rows.mi[:, :n].clone(),)
                tiles.append(t)
        return tiles, total

    def match_many(self, pairs, mask=None, nodata=(None, None), collect=True):
        """pairs: iterable of (mon, ref) CUDA tensors.  Every tile of every pair is
        one unit of work; unit k runs on slot k mod depth (own context and stream) and
        is finalised -- counts read -- only when its slot is needed again, so up to
        `depth` units overlap on the device.  The kernels write the rows of unit k and
        its exchange header straight into record k of one arena (karios_b200/sharding.py:
        the arena is the payload of the multi-GPU exchange; no copy, no packing).
        -> (list per pair of per-tile (rows, zncc[, mi]) views into the arena, total rows)."""
        from karios_b200 import sharding
        results, total = [], 0
        pairs = list(pairs)
        self._check_device()
        # One result arena for the whole call, allocated before any unit is in flight: nothing
        # inside the pipelined loop touches the allocator (a cudaMalloc of the caching allocator
        # while four units are in flight was measured to stall the host for 10-40 ms).
        n_units = len(pairs) * len(self.windows)
        cap = self.rows.capacity
        arena = sharding.new_arena(n_units, cap, self.device)
        arena_mi = torch.empty((max(n_units, 1), 2, cap), dtype=torch.float64, device=self.device) \
            if self.with_mi else None
        self.n_redo, self.redo_flags, self.unit_done_t = 0, [], []     # diagnostics of the last call
        self.unit_events = []
        self.trace_units = getattr(self, "trace_units", False)
        self.last_arena, self.last_counts = arena, [0] * n_units
        inflight = [None] * self.depth
        cur = torch.cuda.current_stream(self.device)

        def unit_rows(unit):
            hdr, f32, z = sharding.unit_views(arena, unit, cap)
            return hdr, N.RowBuffers.from_views(cap, f32, z, arena_mi[unit] if self.with_mi else None)

        def finalise(slot):
            nonlocal total
            job = inflight[slot]
            if job is None:
                return
            inflight[slot] = None
            ctx, _, stream = self._slots[slot]
            pair_idx, mon, ref, win, unit, hdr, rows = job
            with torch.cuda.stream(stream):
                st = ctx.read_stats()
                self.unit_done_t.append(time.perf_counter())
                if st.select_incomplete:          # rare: redo with every candidate
                    self.n_redo += 1
                    self.redo_flags.append((int(st.two_tier_fallback), int(st.overflow), int(st.n_corners),
                                            int(st.n_sorted), int(st.nms_rounds)))
                    st = ctx.match_tile(mon, ref, mask, win, self.kconf, rows, nodata[0], nodata[1])
                    ctx.unit_header(rows, hdr)
                n = int(st.n_kept)
                if (mask is None and st.valid == 0) or st.n_corners == 0:
                    if n:                          # tile skipped by the reference: no rows
                        hdr.zero_()
                    return
                total += n
                self.last_counts[unit] = n
                if collect:
                    t = (rows.f32[:, :n], rows.zncc[:n])
                    if self.with_mi:
                        t = t + (rows.mi[:, :n],)
                    results[pair_idx].append(t)

        k = 0
        for pair_idx, (mon, ref) in enumerate(pairs):
            results.append([])
            for win in self.windows:
                slot = k % self.depth
                finalise(slot)
                ctx, _, stream = self._slots[slot]
                hdr, rows = unit_rows(k)
                stream.wait_stream(cur)
                with torch.cuda.stream(stream):
                    if self.trace_units:
                        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                        ev[0].record()
                        t_enq = time.perf_counter()
                    ctx.match_tile_async(mon, ref, mask, win, self.kconf, rows, nodata[0], nodata[1])
                    ctx.unit_header(rows, hdr)
                    if self.trace_units:
                        ev[1].record()
                        self.unit_events.append(ev + (t_enq, time.perf_counter()))
                inflight[slot] = (pair_idx, mon, ref, win, k, hdr, rows)
                k += 1
                if k == n_units and self.on_last_enqueued is not None:
                    self.on_last_enqueued(arena, [s for _, _, s in self._slots])
        for j in range(self.depth):
            finalise((k + j) % self.depth)
        for _, _, stream in self._slots:
            cur.wait_stream(stream)
        return results, total

    @staticmethod
    def to_frame(tiles) -> DataFrame:
        if not tiles:
            return DataFrame({c: np.empty(0, np.float32) for c in ("x0", "y0", "dx", "dy", "score")}
                             | {"zncc_score": np.empty(0, np.float64)})
        f = torch.cat([t[0] for t in tiles], dim=1).cpu().numpy()
        z = torch.cat([t[1] for t in tiles]).cpu().numpy()
        df = DataFrame({"x0": f[0], "y0": f[1], "dx": f[2], "dy": f[3], "score": f[4]})
        df["zncc_score"] = z
        if tiles and len(tiles[0]) > 2:
            mi = torch.cat([t[2] for t in tiles], dim=1).cpu().numpy()
            df["mutual_info_score"] = mi[0]         # api/core.py:894-897
            df["mi_score"] = mi[1]                  # api/core.py:902-907
        return df

    def close(self):
        for ctx, _, _ in self._slots:
            ctx.close()


def match_pair(mon, ref, mask, conf, confidence_threshold: float = 0.4, nodata=(None, None),
               tail_mode: int = N.KR_TAIL_AVX512) -> DataFrame:
    """One scene pair end to end (upload if needed, every tile, ZNCC)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    mon_t, ref_t = N.to_device(mon, dev), N.to_device(ref, dev)
    mask_t = None if mask is None else N.to_device(mask, dev)
    h, w = mon_t.shape
    sm = _cached_matcher(h, w, conf, confidence_threshold, tail_mode, dev)
    tiles, _ = sm.match_device(mon_t, ref_t, mask_t, nodata)
    return sm.to_frame(tiles)


# match_pair keeps its workspace (a cudaMalloc of ~1.3 GB for an S2 tile costs ~80 ms, forty times
# the matching): one SceneMatcher per thread, replaced when the scene size or the configuration changes
_pair_tls = threading.local()


def _cached_matcher(h, w, conf, confidence_threshold, tail_mode, dev):
    key = (h, w, dev.index, float(confidence_threshold), int(tail_mode),
           tuple(sorted((k, repr(v)) for k, v in vars(conf).items())))
    ent = getattr(_pair_tls, "ent", None)
    if ent is not None and ent[0] == key:
        return ent[1]
    if ent is not None:
        ent[1].close()
        _pair_tls.ent = None
    sm = SceneMatcher(h, w, conf, confidence_threshold, tail_mode, dev)
    _pair_tls.ent = (key, sm)
    return sm


def release_workspaces() -> None:
    """Free the workspace match_pair keeps on this thread."""
    ent = getattr(_pair_tls, "ent", None)
    if ent is not None:
        ent[1].close()
        _pair_tls.ent = None


class ScenePipeline:
    """Host-resident scene pairs -> matches, double-buffered: the H2D copy of pair
    i+1 (copy stream) overlaps the matching of pair i (compute stream)."""

    def __init__(self, h, w, dtype, conf, confidence_threshold=0.4, device=None):
        self.sm = SceneMatcher(h, w, conf, confidence_threshold, device=device)
        dev = self.sm.device
        self.bufs = [(torch.empty((h, w), dtype=dtype, device=dev), torch.empty((h, w), dtype=dtype, device=dev))
                     for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        cap = self.sm.rows.capacity
        self.host_rows = torch.empty((5, cap), dtype=torch.float32).pin_memory()
        self.host_zncc = torch.empty(cap, dtype=torch.float64).pin_memory()

    def _upload(self, slot, mon_h, ref_h):
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            self.bufs[slot][0].copy_(mon_h, non_blocking=True)
            self.bufs[slot][1].copy_(ref_h, non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        del cur

    def copy_only(self, pairs):
        """The uploads of run() without any kernel: the host-to-device ceiling of this
        pipeline on this box (bench.py reports run() as a fraction of it)."""
        if not pairs:
            return
        cur = torch.cuda.current_stream()
        for s in range(2):
            self.free[s].record(cur)
        self._upload(0, *pairs[0])
        for i in range(len(pairs)):
            slot = i & 1
            if i + 1 < len(pairs):
                self._upload(slot ^ 1, *pairs[i + 1])
            cur.wait_event(self.ready[slot])
            self.free[slot].record(cur)
        cur.synchronize()

    def run(self, pairs):
        """pairs: list of (mon_host, ref_host) pinned tensors.  Returns the total
        number of matches; the rows of the last pair stay in host_rows/host_zncc."""
        total = 0
        if not pairs:
            return 0
        self.sm._check_device()
        for s in range(2):
            self.free[s].record(torch.cuda.current_stream())
        self._upload(0, *pairs[0])
        for i in range(len(pairs)):
            slot = i & 1
            if i + 1 < len(pairs):
                self._upload(slot ^ 1, *pairs[i + 1])
            torch.cuda.current_stream().wait_event(self.ready[slot])
            mon_d, ref_d = self.bufs[slot]
            for win in self.sm.windows:
                st = self.sm.ctx.match_tile(mon_d, ref_d, None, win, self.sm.kconf, self.sm.rows)
                n = int(st.n_kept)
                self.host_rows[:, :n].copy_(self.sm.rows.f32[:, :n], non_blocking=True)
                self.host_zncc[:n].copy_(self.sm.rows.zncc[:n], non_blocking=True)
                total += n
            self.free[slot].record(torch.cuda.current_stream())
        torch.cuda.current_stream().synchronize()
        return total

    def close(self):
        self.sm.close()


# ---------------------------------------------------------------------------
# Scene-level passes of KariosAPI around the matcher
def _dev_raster(img, dev):
    full = getattr(img, "device_array", None)
    if full is not None:
        return full
    return N.to_device(img.array if hasattr(img, "array") else img, dev)


def _percentiles_from_hist(t, coarse_hist, lo, shift):
    """NumPy's 'linear' 2nd / 98th percentile from the coarse histogram of raster `t` plus one
    refinement pass (two when the coarse bins of the wanted order statistics are more than 8192
    values apart) at full resolution."""
    coarse = torch.cumsum(coarse_hist, 0).cpu().numpy()
    n = int(coarse[-1])
    wanted = []
    for q in (np.float64(2) / 100, np.float64(98) / 100):
        virtual = n * q + (1 + q * (1 - 1 - 1)) - 1          # numpy _compute_virtual_index, alpha = beta = 1
        prev = np.floor(virtual)
        k0 = int(prev)
        wanted.append((k0, min(k0 + 1, n - 1), virtual - prev))
    ks = sorted({k for k0, k1, _ in wanted for k in (k0, k1)})
    cbs = {k: int(np.searchsorted(coarse, k + 1, side="left")) for k in ks}
    fine = {}
    if shift:
        need = sorted(set(cbs.values()))
        groups, cur = [], [need[0]]
        for cb in need[1:]:                                    # coarse bins close enough share a pass
            if ((cb - cur[0] + 1) << shift) <= 8192:
                cur.append(cb)
            else:
                groups.append(cur)
                cur = [cb]
        groups.append(cur)
        for g in groups:
            base = lo + (g[0] << shift)
            hist = N.histogram(t, base, lo + (g[-1] << shift) + (1 << shift) - 1).cpu().numpy()
            for cb in g:
                off = (cb - g[0]) << shift
                fine[cb] = np.cumsum(hist[off: off + (1 << shift)])

    def value_at(k):                                   # k-th smallest value (0-based)
        cb = cbs[k]
        if shift == 0:
            return lo + cb
        below = int(coarse[cb - 1]) if cb > 0 else 0
        return lo + (cb << shift) + int(np.searchsorted(fine[cb], k + 1 - below, side="left"))

    out = []
    for k0, k1, gamma in wanted:
        a, b = np.float64(value_at(k0)), np.float64(value_at(k1))
        diff = b - a
        out.append(a + diff * gamma if gamma < 0.5 else b - diff * (1 - gamma))   # numpy _lerp
    return np.array(out)


_INT_RANGES = {torch.uint8: (0, 255), torch.uint16: (0, 65535), torch.int16: (-32768, 32767)}


def percentiles_2_98(raster) -> np.ndarray:
    """np.nanpercentile(image.array, [2, 98]) of an integer raster
    (KariosAPI._check_quality, karios/api/core.py:500-506): one coarse device histogram (8 values
    per bin for 16 bits), one refinement pass over the coarse bins that hold the wanted order
    statistics, then NumPy's 'linear' rule."""
    dev = torch.device("cuda", torch.cuda.current_device())
    t = _dev_raster(raster, dev)
    if t.dtype not in _INT_RANGES:
        raise N.KariosB200Error("percentiles need an integer raster (uint8, uint16, int16)")
    lo, hi = _INT_RANGES[t.dtype]
    shift = 0 if t.dtype == torch.uint8 else 3
    return _percentiles_from_hist(t, N.histogram(t, lo, hi, shift), lo, shift)


def scene_scan(raster, mask=None):
    """(2nd / 98th percentiles, valid-pixel count) of one raster: the coarse histogram of
    _check_quality (api/core.py:500-506) and np.count_nonzero under the mask (:285-290) come
    from ONE pass over the raster (kr_histogram_count), the percentile refinement is a second."""
    dev = torch.device("cuda", torch.cuda.current_device())
    t = _dev_raster(raster, dev)
    if t.dtype not in _INT_RANGES:
        raise N.KariosB200Error("scene_scan needs an integer raster (uint8, uint16, int16)")
    m = None if mask is None else _dev_raster(mask, dev)
    if m is not None and m.dtype != torch.uint8:
        m = (m != 0).to(torch.uint8)
    lo, hi = _INT_RANGES[t.dtype]
    shift = 0 if t.dtype == torch.uint8 else 3
    hist, count = N.histogram_count(t, lo, hi, shift, m)
    return _percentiles_from_hist(t, hist, lo, shift), count


def check_quality(monitored_image, reference_image) -> dict:
    """KariosAPI._check_quality (api/core.py:491-506): dynamic range between the 2nd
    and 98th percentile, flagged low when <= 10."""
    res = {}
    for name, img in (("monitored", monitored_image), ("reference", reference_image)):
        mm = percentiles_2_98(img)
        res[name] = {"p2": float(mm[0]), "p98": float(mm[1]), "low_dynamic": bool(mm[1] - mm[0] <= 10)}
    return res


def count_valid_pixels(monitored_image, mask=None) -> int:
    """np.count_nonzero of the monitored image with masked-out pixels zeroed
    (KariosAPI.analyze_accuracy, api/core.py:285-290)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    m = None if mask is None else _dev_raster(mask, dev)
    if m is not None and m.dtype != torch.uint8:
        m = (m != 0).to(torch.uint8)
    return N.count_valid(_dev_raster(monitored_image, dev), m)


def _point_values(points: DataFrame, raster):
    dev = torch.device("cuda", torch.cuda.current_device())
    x = torch.from_numpy(np.array(points["x0"].to_numpy(np.float32), copy=True)).to(dev)
    y = torch.from_numpy(np.array(points["y0"].to_numpy(np.float32), copy=True)).to(dev)
    return N.gather_points(_dev_raster(raster, dev), x, y).cpu().numpy()


def filter_by_dn_values(points: DataFrame, monitored_image, reference_image, no_values=None) -> DataFrame:
    """KariosAPI._filter_by_dn_values (api/core.py:655-744): drop key points whose
    pixel (int(y0), int(x0)) holds an excluded DN in either image or an image's own
    no-data value."""
    ref_nd = getattr(reference_image, "no_data_value", None)
    mon_nd = getattr(monitored_image, "no_data_value", None)
    if not no_values and ref_nd is None and mon_nd is None:
        return points
    if len(points) == 0:
        return points
    ref_values = _point_values(points, reference_image)
    mon_values = _point_values(points, monitored_image)
    keep = np.ones(len(points), dtype=bool)
    for no_value in no_values or []:
        keep &= ~((ref_values == no_value) | (mon_values == no_value))
    if ref_nd is not None:
        keep &= ~(ref_values == ref_nd)
    if mon_nd is not None:
        keep &= ~(mon_values == mon_nd)
    return points[keep].copy()


def dem_altitudes(points: DataFrame, dem) -> np.ndarray:
    """dem.array[int(y0), int(x0)] per key point (api/core.py:1050-1053)."""
    vals = _point_values(points, dem)
    src = getattr(dem, "device_array", None)
    dt = src.dtype if src is not None else getattr(getattr(dem, "array", dem), "dtype", None)
    if dt in (torch.float32, np.dtype(np.float32)):
        return vals.astype(np.float32)
    return vals


def altitude_profile(values, altitudes, bin_size: int = 100):
    """mean_profile (karios/report/commons.py:49-71) of `values` grouped by altitude
    bins: (bin centres int32, count, mean, std with ddof = 1) -- O(N) on <= maxCorners
    rows, NumPy on the host like the reference's pandas groupby."""
    values = np.asarray(values, np.float64)
    group = np.floor_divide(np.asarray(altitudes), bin_size)
    keys = np.unique(group)
    centres = (keys * bin_size + bin_size // 2).astype(np.int32)
    cnt = np.array([np.sum(group == k) for k in keys])
    mean = np.array([values[group == k].mean() for k in keys])
    std = np.array([values[group == k].std(ddof=1) if np.sum(group == k) > 1 else np.nan for k in keys])
    return centres, cnt, mean, std


def match_pair_large_shift(mon, ref, mask, conf, offset_threshold: float, confidence_threshold: float = 0.4):
    """KariosAPI.match_images with enable_large_shift_detection (api/core.py:233-252,
    746-786): whole-pixel offsets from the phase correlation, each axis applied only
    when |offset| >= threshold, KLT on the shifted monitored raster, offsets added back
    to dx / dy, no ZNCC once a shift was applied (api/core.py:876).
    -> (DataFrame, (x_offset, y_offset) applied or None)"""
    from karios_b200.core.image import shift_image
    from karios_b200.matcher.large_offset import phase_cross_correlation_shift
    dev = torch.device("cuda", torch.cuda.current_device())
    mon_t, ref_t = N.to_device(mon, dev), N.to_device(ref, dev)
    offsets = phase_cross_correlation_shift(mon_t, ref_t)
    if abs(offsets[1]) < offset_threshold:
        offsets[1] = 0
    if abs(offsets[0]) < offset_threshold:
        offsets[0] = 0
    if offsets[0] == 0 and offsets[1] == 0:
        return match_pair(mon_t, ref_t, mask, conf, confidence_threshold), None
    shifted = shift_image(mon_t, x_off=offsets[1], y_off=offsets[0])
    df = match_pair(shifted, ref_t, mask, conf, confidence_threshold)
    df["dx"] = df["dx"] + offsets[1]
    df["dy"] = df["dy"] + offsets[0]
    df["zncc_score"] = np.nan
    return df, (offsets[1], offsets[0])
