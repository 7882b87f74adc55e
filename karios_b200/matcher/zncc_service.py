"""GPU drop-in for ZNCCService (karios/matcher/zncc_service.py:154-238, 289-297).

    ZNCCService().compute_zncc(df, monitored, reference) -> Series (float64, NaN
    where the reference yields NaN: chip outside the raster, zero variance)

The per-row Python loop of the reference (`df.apply`, zncc_service.py:177) is
one kernel launch here (kr_zncc); `_zncc2` itself is kept as a host helper with
the reference's argument checks for API parity (zncc_service.py:45-126)."""
from __future__ import annotations

import logging

import numpy as np
import torch
from pandas import DataFrame, Series

from karios_b200 import _native as N
from karios_b200.matcher.klt import get_context

logger = logging.getLogger(__name__)


def _zncc2(img1, img2, u1, v1, u2, v2, n):
    """ZNCC of two (2n+1)^2 windows centred at (row u, column v); raises like the
    reference (ValueError for n < 0, IndexError outside the image) and returns
    NaN for zero variance.  Evaluated on the GPU through kr_zncc when the window
    is the production 43x43 one on chips; otherwise in NumPy (test helper)."""
    if n < 0:
        raise ValueError("Window half-size n must be non-negative")
    h1, w1 = img1.shape
    h2, w2 = img2.shape
    if (u1 - n < 0 or u1 + n >= h1 or v1 - n < 0 or v1 + n >= w1
            or u2 - n < 0 or u2 + n >= h2 or v2 - n < 0 or v2 + n >= w2):
        raise IndexError("Patch window extends beyond image boundaries")
    p1 = np.asarray(img1)[u1 - n:u1 + n + 1, v1 - n:v1 + n + 1].astype(np.float64)
    p2 = np.asarray(img2)[u2 - n:u2 + n + 1, v2 - n:v2 + n + 1].astype(np.float64)
    s1, s2 = p1.std(), p2.std()
    if s1 == 0 or s2 == 0:
        return np.nan
    return float(np.mean(((p1 - p1.mean()) / s1) * ((p2 - p2.mean()) / s2)))


def _raster_tensor(img, dev) -> torch.Tensor:
    full = getattr(img, "device_array", None)
    if full is not None:
        return full
    return N.to_device(img.array, dev)


def _score_inputs(df, monitored, reference):
    """(ref tensor, mon tensor, [x0, y0, dx, dy] float32 device columns)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    mon = _raster_tensor(monitored, dev)
    ref = _raster_tensor(reference, dev)
    if mon.dtype != ref.dtype:
        raise N.KariosB200Error("monitored and reference rasters must share a dtype")
    cols = [torch.from_numpy(np.array(df[c].to_numpy(np.float32), copy=True)).to(dev)
            for c in ("x0", "y0", "dx", "dy")]
    return ref, mon, cols


class ZNCCService:
    """Zero-mean normalised cross-correlation between chips of two rasters."""

    def __init__(self):
        self._chip_size = 57
        self._chip_margin = int((self._chip_size - 1) / 2)

    def compute_zncc(self, df: DataFrame, monitored, reference) -> Series:
        """ZNCC for each key point of `df` (columns x0, y0, dx, dy): Series with
        the index of `df`, NaN where not computable."""
        logger.info("Compute ZNCC for %s points", len(df))
        if len(df) == 0:
            score = Series(np.empty(0, np.float64), index=df.index, dtype=np.float64)
        else:
            ref, mon, cols = _score_inputs(df, monitored, reference)
            ctx = get_context(64, 64, 1024)
            z = ctx.zncc(ref, mon, *cols)
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
            ref, mon, cols = _score_inputs(df, monitored, reference)
            mi = get_context(64, 64, 1024).mutual_info(ref, mon, *cols)
            score = Series(mi[1].cpu().numpy(), index=df.index, dtype=np.float64)
        monitored.clear_cache()
        reference.clear_cache()
        logger.info("NMI computation finish")
        return score
