"""Does the running cut of the corner response engage every time?  N matchings of one S2 pair,
alone and with other pairs in flight; prints the spread of dropped row pieces and candidates."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from karios_b200 import synth
from karios_b200.api import SceneMatcher
from karios_b200.core.configuration import KLTConfiguration

size = 10980
ref, mon = synth.make_pair(size, size, seed=1234, device="cuda")
sm = SceneMatcher(size, size, KLTConfiguration(), 0.4, depth=6)
sk, cand, ms = [], [], []
for i in range(40):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st = sm.ctx.match_tile(mon, ref, None, sm.windows[0], sm.kconf, sm.rows)
    b.record(); torch.cuda.synchronize()
    sk.append(int(st.rows_skipped)); cand.append(int(st.n_candidates)); ms.append(a.elapsed_time(b))
print("alone: dropped row pieces min/median/max", min(sk), int(np.median(sk)), max(sk), "candidates max", max(cand),
      "no-cut runs", sum(1 for s in sk if s == 0), "ms median/max", round(float(np.median(ms)), 3), round(max(ms), 3))
# pipelined: stats of each slot's context after a batch
sk2 = []
for rep in range(6):
    sm.match_many([(mon, ref)] * 12)
    torch.cuda.synchronize()
    for ctx, _, stream in sm._slots:
        with torch.cuda.stream(stream):
            sk2.append(int(ctx.read_stats().rows_skipped))
print("in flight: dropped row pieces min/median/max", min(sk2), int(np.median(sk2)), max(sk2), "no-cut units", sum(1 for s in sk2 if s == 0), "of", len(sk2))
