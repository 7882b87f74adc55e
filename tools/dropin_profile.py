"""Where the time of one drop-in step goes (KLT.match + compute_zncc on NumPy rasters)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from karios_b200 import synth
from karios_b200.core import image as kimg
from karios_b200.core.configuration import KLTConfiguration
from karios_b200.matcher.klt import KLT
from karios_b200.matcher.zncc_service import ZNCCService
import cProfile, pstats

size = 10980
ref, mon = synth.make_pair(size, size, seed=1234, device="cuda")
hp = [t.cpu().pin_memory() for t in (mon, ref)]
npv = [t.view(torch.int16).numpy().view(np.uint16) for t in hp]
conf = KLTConfiguration()
zs = ZNCCService()

def step(sync=False):
    T = {}
    t0 = time.perf_counter()
    mon_img, ref_img = kimg.ArrayRaster(npv[0]), kimg.ArrayRaster(npv[1])
    gen = KLT(conf).match(mon_img, ref_img, None)
    df = next(gen)
    T["match"] = time.perf_counter() - t0; t0 = time.perf_counter()
    cand = df[df["score"] >= 0.4]
    df["zncc_score"] = np.nan
    T["filter"] = time.perf_counter() - t0; t0 = time.perf_counter()
    z = zs.compute_zncc(cand, mon_img, ref_img)
    T["zncc"] = time.perf_counter() - t0; t0 = time.perf_counter()
    df.loc[cand.index, "zncc_score"] = z
    allf = pd.concat([pd.DataFrame(), df])
    T["assign+concat"] = time.perf_counter() - t0
    return T

for _ in range(3): step()
acc = {}
for _ in range(10):
    for k, v in step().items(): acc[k] = acc.get(k, 0) + v / 10
print({k: round(1e3 * v, 3) for k, v in acc.items()}, "total ms", round(1e3 * sum(acc.values()), 3))
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
