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
