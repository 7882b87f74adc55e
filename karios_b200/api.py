"""Scene-level glue of the hot path: what KariosAPI._compute_matches +
_handle_klt_results do around the matcher (karios/api/core.py:845-921), minus
CSV / NMI: per tile KLT._match_tile, then ZNCC for rows with
score >= confidence_threshold, all in one stream-ordered launch sequence per tile.

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
                 tail_mode: int = N.KR_TAIL_AVX512, device=None):
        if isinstance(conf.laplacian_kernel_size, str) or conf.laplacian_invert_polarity == "auto":
            raise N.KariosB200Error("SceneMatcher handles fixed kernel size / polarity; "
                                    "use matcher.klt.KLT for the 'auto' searches")
        self.conf = conf
        self.h, self.w = h, w
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        tw, th = min(w, conf.tile_size), min(h, conf.tile_size)
        self.ctx = N.Context(tw, th, int(conf.maxCorners), self.device)
        cap = int(conf.maxCorners) if conf.maxCorners > 0 else self.ctx._cap_unlimited(tw, th)
        self.rows = N.RowBuffers(cap, self.device, with_zncc=True)
        self.kconf = N.make_conf(conf, tail_mode=tail_mode, compute_zncc=True,
                                 zncc_min_score=confidence_threshold)
        self.windows = tile_windows(w, h, conf)

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
                tiles.append((self.rows.f32[:, :n].clone(), self.rows.zncc[:n].clone()))
        return tiles, total

    @staticmethod
    def to_frame(tiles) -> DataFrame:
        if not tiles:
            return DataFrame({c: np.empty(0, np.float32) for c in ("x0", "y0", "dx", "dy", "score")}
                             | {"zncc_score": np.empty(0, np.float64)})
        f = torch.cat([t[0] for t in tiles], dim=1).cpu().numpy()
        z = torch.cat([t[1] for t in tiles]).cpu().numpy()
        df = DataFrame({"x0": f[0], "y0": f[1], "dx": f[2], "dy": f[3], "score": f[4]})
        df["zncc_score"] = z
        return df

    def close(self):
        self.ctx.close()


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
