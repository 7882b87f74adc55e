import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from karios_b200 import synth, _native as N
size = 10980
ref, mon = synth.make_pair(size, size, seed=77, device="cuda")
def variants():
    a = ref.clone().to(torch.int32)
    yield "uniform", ref
    b = a.clone(); b[size//2:] = (b[size//2:] - 2500) * 3 // 10 + 2500          # weak bottom half
    yield "weak bottom half", b.to(torch.uint16)
    c = a.clone(); c[:, :size//3] = (c[:, :size//3] - 2500) // 8 + 2500; c[3000:5000, 6000:9000] = 0   # weak left third + nodata block
    yield "weak left third + zeros", c.to(torch.uint16)
    d = a.clone(); d[:9000] = (d[:9000] - 2500) // 10 + 2500                      # strong content only at the bottom
    yield "strong bottom only", d.to(torch.uint16)
ctx = N.Context(size, size, 20000)
for name, img in variants():
    ctx.minmax_mask(img, img, want_mask=False)
    lap = ctx.u8_laplacian(img, 7, slot=0)
    mask = (img != 0).to(torch.uint8)
    ctx.set_corner_mode(1)
    want = ctx.good_features(lap, mask, 20000, 0.1, 10, 15)
    ctx.set_corner_mode(0)
    for rep in range(3):
        got = ctx.good_features(lap, mask, 20000, 0.1, 10, 15)
        st = ctx.read_stats()
        print(name, rep, "equal", bool(torch.equal(got, want)), "corners", got.shape[0], "two_tier", st.two_tier, "fallback", st.two_tier_fallback,
              "dropped", st.rows_skipped, "cand", st.n_candidates, "est", hex(st.est_cut_bits))
