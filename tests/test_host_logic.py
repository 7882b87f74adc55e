"""Host-side logic of the drop-in classes that needs no GPU: the outlier filter's keep mask, the
memo of the scores computed with the matching pass, the exchange record layout helpers used by
bench.py, and the two bench arms agreeing on the workload description."""
import importlib
import os
import sys

import numpy as np
import pandas as pd
import pytest

from oracle import oracle as O


def test_outlier_keep_equals_reference_filter():
    from karios_b200.matcher.klt import _outlier_keep
    rng = np.random.default_rng(0)
    for n, n_out in ((400, 25), (50, 0), (3000, 300), (7, 2)):
        dx = rng.normal(0.3, 0.05, n).astype(np.float32)
        dy = rng.normal(-0.2, 0.05, n).astype(np.float32)
        bad = rng.choice(n, n_out, replace=False)
        dx[bad] += rng.normal(0, 8, n_out).astype(np.float32)
        dy[bad] += rng.normal(0, 30, n_out).astype(np.float32)
        x0 = rng.integers(0, 5000, n).astype(np.float32)
        y0 = rng.integers(0, 5000, n).astype(np.float32)
        score = rng.random(n).astype(np.float32)
        keep = _outlier_keep(dx, dy)
        wx0, wy0, wx1, wy1, wscore = O.filter_outliers(x0, y0, x0 + dx, y0 + dy, score)
        assert np.array_equal(x0[keep], wx0) and np.array_equal(y0[keep], wy0)
        assert np.array_equal(score[keep], wscore)
        try:
            from oracle import refimport
            if refimport.available() or os.path.isdir(os.path.join(refimport.VENDORED, "karios", "matcher")):
                klt, _, _ = refimport.load()
                f = getattr(klt, "__filter_outliers")             # module-level name with two underscores
                rx0, ry0, rx1, ry1, rscore = f(x0, y0, x0 + dx, y0 + dy, score)
                assert np.array_equal(x0[keep], rx0) and np.array_equal(score[keep], rscore)
        except ImportError:
            pass


def test_score_memo_serves_only_identical_rows():
    from karios_b200.core import image as kimg

    class R:
        pass
    mon, ref, other = R(), R(), R()
    rng = np.random.default_rng(1)
    cols = rng.random((4, 50)).astype(np.float32)
    z = rng.random(50)
    z[3] = np.nan
    kimg.remember_scores(mon, ref, cols, z)
    df = pd.DataFrame({"x0": cols[0], "y0": cols[1], "dx": cols[2], "dy": cols[3], "score": np.float32(1)})
    sub = df[df["x0"] > 0.5]
    got = kimg.recall_scores(mon, ref, sub)
    assert got is not None and np.array_equal(np.nan_to_num(got, nan=-1), np.nan_to_num(z[sub.index.to_numpy()], nan=-1))
    assert kimg.recall_scores(mon, other, sub) is None                    # another reference raster
    assert kimg.recall_scores(other, ref, sub) is None                    # another monitored raster
    moved = sub.copy()
    moved["dy"] = moved["dy"] + np.float32(0.25)
    assert kimg.recall_scores(mon, ref, moved) is None                    # not the rows of the tile
    assert kimg.recall_scores(mon, ref, sub.reset_index(drop=True)) is None or len(sub) == len(df)
    wide = sub.astype({"dx": np.float64})
    assert kimg.recall_scores(mon, ref, wide) is None                     # not the yielded float32 columns
    assert kimg.recall_scores(mon, ref, df.iloc[:0]) is None              # empty
    out = sub.copy()
    out.index = out.index + 1000
    assert kimg.recall_scores(mon, ref, out) is None                      # labels outside the tile
    kimg.release_device(mon)
    assert kimg.recall_scores(mon, ref, sub) is None


def test_bench_arms_describe_the_same_workload():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    for world in (1, 2, 8):
        a, b = bench.bench_config(10980, world), bench.bench_config(10980, world)
        assert a == b and a["workload"] == bench.WORKLOAD and str(world) in a["parallelism"]
    # the reference arm takes its configuration from the same defaults the CUDA arm uses
    from oracle import ref_run
    from karios_b200.core.configuration import KLTConfiguration
    assert vars(bench.default_conf(KLTConfiguration)) == ref_run.DEFAULT_KLT


def test_scene_matcher_rejects_before_touching_the_gpu():
    from karios_b200 import _native as N
    from karios_b200.api import SceneMatcher
    from karios_b200.core.configuration import KLTConfiguration
    with pytest.raises(N.KariosB200Error, match="outliers_filtering"):
        SceneMatcher(64, 64, KLTConfiguration(outliers_filtering=True))
    with pytest.raises(N.KariosB200Error, match="auto"):
        SceneMatcher(64, 64, KLTConfiguration(laplacian_invert_polarity="auto"))
