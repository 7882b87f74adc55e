"""GPU drop-in for ZNCCService (karios/matcher/zncc_service.py:154-238, 289-297).

    ZNCCService().compute_zncc(df, monitored, reference) -> Series (float64, NaN
    where the reference yields NaN: chip outside the raster, zero variance)

The per-row Python loop of the reference (`df.apply`, zncc_service.py:177) is
one kernel launch here (kr_zncc).  The module-private helper `_zncc2` of the
reference (zncc_service.py:45-126) has no counterpart in this package: nothing
computes on the host here (its restatement lives in oracle/oracle.py, for the
tests).  Host rasters are uploaded once per raster object and shared with
KLT.match and the mutual-information service (core.image.device_full)."""
from __future__ import annotations

import logging

import numpy as np
import torch
from pandas import DataFrame, Series

from karios_b200 import _native as N
from karios_b200.core.image import device_full, recall_mi, recall_scores, remember_mi
from karios_b200.matcher.klt import get_context

logger = logging.getLogger(__name__)


def _raster_tensor(img, dev) -> torch.Tensor:
    return device_full(img, dev)


def _score_inputs(df, monitored, reference, margin=28):
    """(ref tensor, mon tensor, [x0, y0, dx, dy] float32 device columns), or None when no row
    of `df` has both chips inside the rasters (the reference returns NaN for those rows
    before it touches the pixel data, zncc_service.py:208-215 -- so the rasters are not
    read, or uploaded, either).

    The reference evaluates `round(series["x0"] + series["dx"])` on the row Series of
    `df.apply(axis=1)`, whose dtype is the common dtype of the frame's columns: float32 for
    the frames KLT.match yields, float64 as soon as the frame holds a float64 column.  The
    kernel adds in float32; for a float64 frame the rounded monitored position is therefore
    formed here in float64 and handed over as an exact integer displacement."""
    cols = {c: df[c].to_numpy() for c in ("x0", "y0", "dx", "dy")}
    host = np.empty((4, len(df)), np.float32)
    common = np.result_type(*[df[c].dtype for c in df.columns]) if len(df.columns) else np.float32
    for i, c in enumerate(("x0", "y0", "dx", "dy")):
        host[i] = cols[c]
    with np.errstate(invalid="ignore"):
        if common != np.float32:
            for i, (p, d) in enumerate((("x0", "dx"), ("y0", "dy"))):
                p64 = cols[p].astype(np.float64)
                x1 = np.rint(p64 + cols[d].astype(np.float64))
                host[2 + i] = x1 - np.trunc(p64)
                host[i] = np.trunc(p64)
        ax, ay = np.trunc(host[0]), np.trunc(host[1])
        bx, by = np.rint(host[0] + host[2]), np.rint(host[1] + host[3])
        inside = ((ax - margin >= 0) & (ay - margin >= 0) & (bx - margin >= 0) & (by - margin >= 0)
                  & (ax < reference.x_size - margin) & (ay < reference.y_size - margin)
                  & (bx < monitored.x_size - margin) & (by < monitored.y_size - margin))
    if not inside.any():
        return None
    dev = torch.device("cuda", torch.cuda.current_device())
    mon = _raster_tensor(monitored, dev)
    ref = _raster_tensor(reference, dev)
    if mon.dtype != ref.dtype:
        raise N.KariosB200Error("monitored and reference rasters must share a dtype")
    dcols = torch.from_numpy(host).to(dev)             # one H2D copy for the four columns
    return ref, mon, [dcols[i] for i in range(4)], host


def mutual_info_pair(df, monitored, reference):
    """[2, n] float64 -- row 0: Studholme NMI (MutualInfoService), row 1: 2 MI / (Hx + Hy)
    (ZNCCService.compute_mi) -- of the rows of `df`; None when no row has its chips inside the
    rasters.  One kr_mutual_info launch gives both; the second of the reference's two calls
    (api/core.py:894-907) is served from the first."""
    cols = np.empty((4, len(df)), np.float32)
    for i, c in enumerate(("x0", "y0", "dx", "dy")):
        cols[i] = df[c].to_numpy()
    known = recall_mi(monitored, reference, cols)
    if known is not None:
        return known
    inp = _score_inputs(df, monitored, reference)
    if inp is None:
        return None
    pair = get_context(64, 64, 1024).mutual_info(*inp[:2], *inp[2]).cpu().numpy()
    remember_mi(monitored, reference, cols, pair)
    return pair


class ZNCCService:
    """Zero-mean normalised cross-correlation between chips of two rasters."""

    def __init__(self):
        self._chip_size = 57
        self._chip_margin = int((self._chip_size - 1) / 2)

    def compute_zncc(self, df: DataFrame, monitored, reference) -> Series:
        """ZNCC for each key point of `df` (columns x0, y0, dx, dy): Series with
        the index of `df`, NaN where not computable."""
        logger.info("Compute ZNCC for %s points", len(df))
        known = recall_scores(monitored, reference, df) if len(df) else None
        if len(df) == 0:
            score = Series(np.empty(0, np.float64), index=df.index, dtype=np.float64)
        elif known is not None:            # computed along with the matching of this tile
            score = Series(known, index=df.index, dtype=np.float64)
        else:
            inp = _score_inputs(df, monitored, reference)
            if inp is None:
                score = Series(np.full(len(df), np.nan), index=df.index, dtype=np.float64)
            else:
                z = get_context(64, 64, 1024).zncc(*inp[:2], *inp[2])
                score = Series(z.cpu().numpy(), index=df.index, dtype=np.float64)
        monitored.clear_cache()
        reference.clear_cache()
        logger.info("ZNCC computation finish")
        return score

    def compute_mi(self, df: DataFrame, monitored, reference) -> Series:
        """NMI = 2 MI / (H(X) + H(Y)) of the 57x57 chips of each key point
        (zncc_service.py:240-287, _mutual_information :129-151): one launch of
        kr_mutual_info instead of the per-row np.histogram2d loop."""
        logger.info("Compute NMI for %s points", len(df))
        if len(df) == 0:
            score = Series(np.empty(0, np.float64), index=df.index, dtype=np.float64)
        else:
            pair = mutual_info_pair(df, monitored, reference)
            score = Series(np.full(len(df), np.nan) if pair is None else pair[1], index=df.index, dtype=np.float64)
        monitored.clear_cache()
        reference.clear_cache()
        logger.info("NMI computation finish")
        return score
