"""KLT parameter contract, field for field the reference's dataclass
(karios/core/configuration.py:36-50) so a KARIOS `KLTConfiguration` object and
this one are interchangeable (the matcher only reads attributes)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Literal, Union


@dataclass
class KLTConfiguration:
    # pylint: disable=invalid-name, too-many-instance-attributes
    minDistance: int = 10
    blocksize: int = 15
    maxCorners: int = 20000
    matching_winsize: int = 25
    qualityLevel: float = 0.1
    xStart: int = 0
    tile_size: int = 20000
    laplacian_kernel_size: Union[int, Dict[str, int], Literal["auto"]] = 7
    outliers_filtering: bool = False
    laplacian_invert_polarity: Union[bool, Literal["auto"]] = False
