"""Drive the UNMODIFIED reference over one scene pair, the way its own API does.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by karios_b200/).

    run_pair(mon, ref, mask=None, **klt_conf) -> (DataFrame, seconds, description)

The call sequence is that of KariosAPI._compute_matches + _handle_klt_results
(karios/api/core.py:845-891): `KLT(conf).match(monitored, reference, mask)`
yields one DataFrame per tile (karios/matcher/klt.py:198-349); for each of them
`ZNCCService().compute_zncc(df[score >= threshold], monitored, reference)`
(karios/matcher/zncc_service.py:162-184, the per-row `df.apply`) fills
`zncc_score`.  The radial-error / angle columns, the two mutual-information
scores and the CSV append are left out on both arms of the benchmark (they are
not part of the headline path).  Every line that computes is the
reference's own file, loaded by oracle/refimport.py from oracle/_ref (or
/root/reference); OpenCV is the installed opencv-python, with all host threads.
"""
from __future__ import annotations

import os
import time

import numpy as np

from oracle import refimport

# karios/configuration/processing_configuration.json:8-19 (the CLI default)
DEFAULT_KLT = dict(minDistance=10, blocksize=15, maxCorners=20000, matching_winsize=25,
                   qualityLevel=0.1, xStart=0, tile_size=20000, laplacian_kernel_size=7,
                   outliers_filtering=False, laplacian_invert_polarity=False)


def describe() -> str:
    import cv2
    import pandas
    return (f"unmodified karios/matcher/klt.py + zncc_service.py from {os.path.relpath(refimport.REF_ROOT)}"
            f" (KLT.match + ZNCCService.compute_zncc), opencv-{cv2.__version__}, pandas-{pandas.__version__}")


def run_pair(mon: np.ndarray, ref: np.ndarray, mask=None, threshold: float = 0.4, threads: int | None = None,
             no_data=(None, None), **klt_conf):
    """One full pass of the reference path over host arrays.  Returns the
    concatenated DataFrame (x0, y0, dx, dy, score, zncc_score), the wall time of
    the pass and a description of what ran."""
    import cv2
    import pandas as pd
    klt, zs, cfg = refimport.load()
    cv2.setNumThreads(threads if threads is not None else (os.cpu_count() or 1))
    kw = dict(DEFAULT_KLT)
    kw.update(klt_conf)
    conf = cfg.KLTConfiguration(**kw)
    mon_img = refimport.ArrayImage(mon, no_data[0])
    ref_img = refimport.ArrayImage(ref, no_data[1])
    mask_img = None if mask is None else refimport.ArrayImage(mask)
    service = zs.ZNCCService()
    t0 = time.perf_counter()
    all_frame = pd.DataFrame()
    for dataframe in klt.KLT(conf).match(mon_img, ref_img, mask_img):        # api/core.py:845
        zncc_candidates = dataframe[dataframe["score"] >= threshold]       # :884
        dataframe["zncc_score"] = np.nan                                    # :887
        zncc_scores = service.compute_zncc(zncc_candidates, mon_img, ref_img)   # :891
        dataframe.loc[zncc_candidates.index, "zncc_score"] = zncc_scores    # :899
        all_frame = pd.concat([all_frame, dataframe])                       # :919
    secs = time.perf_counter() - t0
    return all_frame, secs, describe()
