"""GPU drop-in for LargeOffsetMatcher (karios/matcher/large_offset.py:24-41).

    LargeOffsetMatcher(reference_image, monitored_image).match() -> array([dy, dx])

The reference calls skimage.registration.phase_cross_correlation(mon.array,
ref.array) with its defaults (upsample_factor 1, normalization "phase", arguments
deliberately swapped, large_offset.py:38-41) and keeps element [0], the whole-pixel
shift.  Here: both rasters in float64 on the device, two real FFTs (cuFFT through
torch.fft), the cross-power spectrum normalised in place (kr_cross_power), one
inverse FFT and the first maximum of |correlation| (kr_argmax_abs).  The half
spectrum of a real FFT is enough because the correlation of two real images is real.
No CPU fallback."""
from __future__ import annotations

import numpy as np
import torch

from karios_b200 import _native as N


def phase_cross_correlation_shift(reference_image, moving_image) -> np.ndarray:
    """Whole-pixel shift (row, col), float64, as skimage 0.24
    registration/_phase_cross_correlation.py computes it for upsample_factor = 1."""
    dev = torch.device("cuda", torch.cuda.current_device())
    ref = N.to_device(reference_image, dev)
    mov = N.to_device(moving_image, dev)
    if ref.shape != mov.shape:
        raise ValueError("images must be same shape")
    shape = tuple(ref.shape)
    src_freq = torch.fft.rfft2(ref.to(torch.float64)).contiguous()
    target_freq = torch.fft.rfft2(mov.to(torch.float64)).contiguous()
    N.cross_power_(src_freq, target_freq)                  # src * conj(target) / max(|.|, 100 eps)
    del target_freq
    cc = torch.fft.irfft2(src_freq, s=shape)
    del src_freq
    flat = int(N.argmax_abs(cc.reshape(-1)).item())
    maxima = np.array(np.unravel_index(flat, shape), dtype=np.float64)
    midpoint = np.array([np.fix(axis_size / 2) for axis_size in shape])
    shift = maxima.copy()
    over = shift > midpoint
    shift[over] -= np.array(shape, dtype=np.float64)[over]
    for dim in range(2):
        if shape[dim] == 1:
            shift[dim] = 0
    return shift


def _raster(img):
    full = getattr(img, "device_array", None)
    return full if full is not None else img.array


class LargeOffsetMatcher:
    """Class to compute row/col offset between 2 images"""

    def __init__(self, reference_image, monitored_image):
        self._ref = reference_image
        self._mon = monitored_image

    def match(self):
        """-> [row (y), col (x)] offset; ref and mon are deliberately inverted, as in
        the reference."""
        return phase_cross_correlation_shift(_raster(self._mon), _raster(self._ref))
