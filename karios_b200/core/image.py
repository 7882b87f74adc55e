"""In-memory stand-ins for GdalRasterImage (karios/core/image.py:255): the
matcher only needs .read/.array/.x_size/.y_size/.no_data_value/.clear_cache.
DeviceRaster keeps the raster in HBM so tiles are windows, not copies."""
from __future__ import annotations

import numpy as np
import torch


class ArrayRaster:
    """NumPy-backed raster (what GdalRasterImage.array / .read return)."""

    def __init__(self, arr: np.ndarray, no_data_value=None, filepath="memory"):
        self._a = arr
        self.no_data_value = no_data_value
        self.y_size, self.x_size = arr.shape
        self.filepath = filepath

    @property
    def array(self):
        return self._a

    def read(self, band, x_off, y_off, x_size, y_size):  # noqa: ARG002
        return self._a[y_off:y_off + y_size, x_off:x_off + x_size]

    def clear_cache(self):
        pass


class DeviceRaster(ArrayRaster):
    """Raster resident on the GPU (torch CUDA tensor, 2-D).  `device_array`
    marks it for the zero-copy path of KLT.match / ZNCCService."""

    def __init__(self, t: torch.Tensor, no_data_value=None, filepath="device"):
        if t.dim() != 2 or t.device.type != "cuda":
            raise ValueError("DeviceRaster needs a 2-D CUDA tensor")
        self._a = t
        self.device_array = t
        self.no_data_value = no_data_value
        self.y_size, self.x_size = t.shape
        self.filepath = filepath


def shift_image(img, y_off=0, x_off=0):
    """shift_image (karios/core/image.py:70-101): whole-pixel shift, shape kept, zeros
    where the source leaves the raster.  A CUDA tensor gives a CUDA tensor; a NumPy
    array is uploaded, shifted on the device (kr_shift_image) and returned as NumPy."""
    from karios_b200 import _native as N
    if isinstance(img, torch.Tensor):
        return N.shift_image(img, y_off, x_off)
    dev = torch.device("cuda", torch.cuda.current_device())
    arr = np.asarray(img)
    out = N.shift_image(N.to_device(arr, dev), y_off, x_off).cpu()
    if arr.dtype == np.uint16:
        return out.view(torch.int16).numpy().view(np.uint16)
    return out.numpy()
