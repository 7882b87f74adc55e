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
        self._slots = [(self.ctx, self.rows, torch.cuda.Stream(device=self.device))]
        for _ in range(self.depth - 1):
            self._slots.append((N.Context(tw, th, int(conf.maxCorners), self.device),
                                N.RowBuffers(cap, self.device, with_zncc=True, with_mi=self.with_mi),
                                torch.cuda.Stream(device=self.device)))

    def match_device(self, mon: torch.Tensor, ref: torch.Tensor, mask=None, nodata=(None, None),
                     collect=True):
        """Both rasters resident in HBM.  -> (list of per-tile [6, n] float64-free
        device tensors (x0,y0,dx,dy,score) + zncc, total rows)."""
        tiles, total = [], 0
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
        one unit of work; unit k runs on slot k mod depth (own context, row buffers
        and stream) and is finalised -- counts read, rows cloned -- only when its
        slot is needed again, so up to `depth` units overlap on the device.
        -> (list per pair of per-tile (rows, zncc[, mi]) tuples, total rows)."""
        results, total = [], 0
        inflight = [None] * self.depth
        cur = torch.cuda.current_stream(self.device)

        def finalise(slot):
            nonlocal total
            job = inflight[slot]
            if job is None:
                return
            inflight[slot] = None
            ctx, rows, stream = self._slots[slot]
            pair_idx, mon, ref, win = job
            with torch.cuda.stream(stream):
                st = ctx.read_stats()
                if st.select_incomplete:          # rare: redo with every candidate
                    st = ctx.match_tile(mon, ref, mask, win, self.kconf, rows, nodata[0], nodata[1])
                n = int(st.n_kept)
                if (mask is None and st.valid == 0) or st.n_corners == 0:
                    return
                total += n
                if collect:
                    t = (rows.f32[:, :n].clone(), rows.zncc[:n].clone())
                    if self.with_mi:
                        t = t + (rows.mi[:, :n].clone(),)
                    for x in t:
                        x.record_stream(cur)
                    results[pair_idx].append(t)

        k = 0
        for pair_idx, (mon, ref) in enumerate(pairs):
            results.append([])
            for win in self.windows:
                slot = k % self.depth
                finalise(slot)
                ctx, rows, stream = self._slots[slot]
                stream.wait_stream(cur)
                with torch.cuda.stream(stream):
                    ctx.match_tile_async(mon, ref, mask, win, self.kconf, rows, nodata[0], nodata[1])
                inflight[slot] = (pair_idx, mon, ref, win)
                k += 1
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
    sm = SceneMatcher(h, w, conf, confidence_threshold, tail_mode, dev)
    try:
        tiles, _ = sm.match_device(mon_t, ref_t, mask_t, nodata)
        return sm.to_frame(tiles)
    finally:
        sm.close()


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

    def run(self, pairs):
        """pairs: list of (mon_host, ref_host) pinned tensors.  Returns the total
        number of matches; the rows of the last pair stay in host_rows/host_zncc."""
        total = 0
        if not pairs:
            return 0
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
