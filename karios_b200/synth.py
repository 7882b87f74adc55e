"""Synthetic Sentinel-2-shaped scene pairs.

Texture t = blur(U[0,1), sigma 2) + 0.5 * blur(U[0,1), sigma 8), normalised to
[0,1]; ref = uint16(t*3000 + 1000); mon = the same texture translated by a
sub-pixel shift (separable cubic interpolation, reflect border) and mapped the
same way.  Written with torch so the same generator runs on the host and on
the device (bench scenes are generated in HBM); no cv2 dependency.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _gauss_kernel(sigma: float, device) -> torch.Tensor:
    r = max(1, int(math.ceil(4.0 * sigma)))
    x = torch.arange(-r, r + 1, dtype=torch.float32, device=device)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _sep_filter(img: torch.Tensor, kx: torch.Tensor, ky: torch.Tensor) -> torch.Tensor:
    """img [H,W] float32; separable correlation with reflect padding."""
    x = img[None, None]
    rx, ry = (len(kx) - 1) // 2, (len(ky) - 1) // 2
    x = F.conv2d(F.pad(x, (rx, rx, 0, 0), mode="reflect"), kx.view(1, 1, 1, -1))
    x = F.conv2d(F.pad(x, (0, 0, ry, ry), mode="reflect"), ky.view(1, 1, -1, 1))
    return x[0, 0]


def _cubic_weights(frac: float, device) -> torch.Tensor:
    """Keys cubic (a = -0.75) weights for sampling at offset `frac` in [0,1)."""
    a = -0.75
    t = frac
    w0 = ((a * (t + 1) - 5 * a) * (t + 1) + 8 * a) * (t + 1) - 4 * a
    w1 = ((a + 2) * t - (a + 3)) * t * t + 1
    w2 = ((a + 2) * (1 - t) - (a + 3)) * (1 - t) * (1 - t) + 1
    w3 = 1.0 - w0 - w1 - w2
    return torch.tensor([w0, w1, w2, w3], dtype=torch.float32, device=device)


def _translate(img: torch.Tensor, sx: float, sy: float) -> torch.Tensor:
    """out(x, y) = img(x - sx, y - sy): features move by (+sx, +sy)."""
    def taps(s):
        # sample position p = x - s = (x + i0) + frac ; 4 taps at i0-1 .. i0+2
        i0 = math.floor(-s)
        frac = -s - i0
        return i0, _cubic_weights(frac, img.device)
    ix, wx = taps(sx)
    iy, wy = taps(sy)
    pad = 4 + max(abs(ix), abs(iy))
    x = F.pad(img[None, None], (pad, pad, pad, pad), mode="reflect")
    h, w = img.shape
    # horizontal
    x = F.conv2d(x, wx.view(1, 1, 1, 4))
    x = x[..., pad + ix - 1: pad + ix - 1 + w]
    x = F.conv2d(x, wy.view(1, 1, 4, 1))
    x = x[..., pad + iy - 1: pad + iy - 1 + h, :]
    return x[0, 0]


def make_texture(h: int, w: int, seed: int, device="cpu") -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    a = torch.rand((h, w), generator=g, device=device, dtype=torch.float32)
    k2 = _gauss_kernel(2.0, device)
    t = _sep_filter(a, k2, k2)
    del a
    b = torch.rand((h, w), generator=g, device=device, dtype=torch.float32)
    k8 = _gauss_kernel(8.0, device)
    t += 0.5 * _sep_filter(b, k8, k8)
    del b
    t -= t.min()
    t /= t.max()
    return t


def to_u16(t: torch.Tensor) -> torch.Tensor:
    """uint16(t*3000 + 1000) stored as torch.uint16."""
    return (t * 3000.0 + 1000.0).clamp_(0, 65535).to(torch.int32).to(torch.uint16)


def make_pair(h: int, w: int, seed: int = 1234, shift=(0.30, -0.20), device="cpu",
              int_shift=(0, 0)):
    """-> (ref, mon) torch.uint16 [h, w] on `device`.  `int_shift` = extra whole
    pixel offset (columns, rows) with zero fill, for the large-shift config."""
    t = make_texture(h, w, seed, device)
    ref = to_u16(t)
    mon = to_u16(_translate(t, float(shift[0]), float(shift[1])).clamp_(0, 1))
    del t
    cx, cy = int(int_shift[0]), int(int_shift[1])
    if cx or cy:
        out = torch.zeros_like(mon)
        ys, yd = (slice(0, h - cy), slice(cy, h)) if cy >= 0 else (slice(-cy, h), slice(0, h + cy))
        xs, xd = (slice(0, w - cx), slice(cx, w)) if cx >= 0 else (slice(-cx, w), slice(0, w + cx))
        out[yd, xd] = mon[ys, xs]
        mon = out
    return ref, mon


def make_mask(h: int, w: int, seed: int = 99, device="cpu") -> torch.Tensor:
    """uint8 validity mask of config 3: ~30 % zeroed (left 40 columns, two large
    rectangles, 200 random 64x64 blocks)."""
    m = torch.ones((h, w), dtype=torch.uint8, device=device)
    m[:, :min(40, w)] = 0
    m[h // 8: h // 8 + h // 3, w // 6: w // 6 + w // 3] = 0
    m[(5 * h) // 8: (5 * h) // 8 + h // 4, w // 2: w // 2 + (2 * w) // 5] = 0
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    ys = torch.randint(0, max(1, h - 64), (200,), generator=g).tolist()
    xs = torch.randint(0, max(1, w - 64), (200,), generator=g).tolist()
    for y, x in zip(ys, xs):
        m[y:y + 64, x:x + 64] = 0
    return m
