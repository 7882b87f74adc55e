"""GPU drop-in for MutualInfoService (karios/matcher/mutual_info_service.py:66-138).

    MutualInfoService().compute_mutual_info(df, monitored, reference) -> Series

Studholme's normalised mutual information (H(X) + H(Y)) / H(X, Y) of the two
57x57 chips of every key point, natural logarithm, 32 x 32 np.histogram2d bins
(mutual_info_service.py:32-63).  The per-row `df.apply` of the reference is one
launch of kr_mutual_info; NaN where the reference returns NaN (chip outside the
raster, H(X, Y) == 0).  No CPU fallback."""
from __future__ import annotations

import logging

import numpy as np
from pandas import DataFrame, Series

from karios_b200.matcher.zncc_service import mutual_info_pair

logger = logging.getLogger(__name__)


class MutualInfoService:
    """Normalised mutual information between chips of two rasters."""

    def __init__(self):
        self._chip_size = 57
        self._chip_margin = int((self._chip_size - 1) / 2)

    def compute_mutual_info(self, df: DataFrame, monitored, reference) -> Series:
        """NMI for each key point of `df` (columns x0, y0, dx, dy): Series with the
        index of `df`, NaN where not computed."""
        logger.info("Compute mutual information for %s points", len(df))
        if len(df) == 0:
            score = Series(np.empty(0, np.float64), index=df.index, dtype=np.float64)
        else:
            pair = mutual_info_pair(df, monitored, reference)
            score = Series(np.full(len(df), np.nan) if pair is None else pair[0], index=df.index, dtype=np.float64)
        monitored.clear_cache()
        reference.clear_cache()
        logger.info("Mutual information computation finish")
        return score
