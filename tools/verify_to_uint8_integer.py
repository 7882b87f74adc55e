#!/usr/bin/env python
"""Exhaustive check behind the arithmetic _to_uint8 of k_laplacian4 (kr_prep.cu, lap4_body): for
EVERY 16-bit value range R = max - min and every value n = v - min in it, the reference's float64
expression ((v - mn) / (mx - mn) * 255).astype(uint8) (karios/matcher/klt.py:47-48) equals
floor(n * 255 / R), and for R >= 256 also (n * ceil(2^32 * 255 / R)) >> 32 with a 32-bit
multiplier.  Takes about five minutes; last run: "done bad 0".  tests/test_oracle.py runs a sample."""
import numpy as np
bad=0
for R in range(1,65536):
    v=np.arange(0,R+1,dtype=np.uint16)
    ref=((v-0.0)/(float(R)-0.0)*255).astype(np.uint8)
    n=v.astype(np.int64)
    exact=(n*255)//R
    ok=np.array_equal(ref,exact)
    if R>=256:
        M=-(-(255<<32)//R)
        ok = ok and np.array_equal(exact,(n*M)>>32) and M<2**32
    if not ok:
        bad+=1; print('MISMATCH R',R, flush=True)
print('done bad',bad, flush=True)
