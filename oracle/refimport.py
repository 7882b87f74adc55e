"""Import the UNMODIFIED reference hot-path modules.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by karios_b200/).

Source tree, in this order: $KARIOS_REFERENCE, /root/reference (build container
only), oracle/_ref (byte-identical copies placed by oracle/vendor_ref.py during
build(); they travel to the GPU box).  Used by oracle/make_golden.py, by the
tests that pin the oracle, and by bench.py's CPU legs (`--impl reference`,
`cpu_baseline`).  Stubs follow SURVEY.md Appendix B: skimage / osgeo are absent,
and karios/__init__.py (-> matplotlib) must not execute.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
VENDORED = os.path.join(_HERE, "_ref")


def _pick_root() -> str:
    env = os.environ.get("KARIOS_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/karios/matcher"):
        return "/root/reference"
    return VENDORED


REF_ROOT = _pick_root()


def use_vendored() -> None:
    """Point the loader at oracle/_ref (what the GPU box has) even when
    /root/reference exists: bench.py runs the same files here and there."""
    global REF_ROOT
    if "karios.matcher.klt" in sys.modules:
        return
    REF_ROOT = VENDORED


def available() -> bool:
    try:
        import cv2  # noqa: F401
        import pandas  # noqa: F401
    except Exception:  # noqa: BLE001
        return False
    return os.path.isdir(os.path.join(REF_ROOT, "karios", "matcher"))


def load():
    """-> (klt module, zncc_service module, configuration module)"""
    if "karios.matcher.klt" in sys.modules and getattr(sys.modules["karios"], "_b200_stub", False):
        m = sys.modules
        return m["karios.matcher.klt"], m["karios.matcher.zncc_service"], m["karios.core.configuration"]
    for name in ("skimage", "skimage.io", "osgeo", "osgeo.gdal", "osgeo.osr", "osgeo.ogr"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["osgeo.gdal"].GDT_Byte = 1
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    sys.modules["osgeo"].gdal = sys.modules["osgeo.gdal"]
    sys.modules["osgeo"].osr = sys.modules["osgeo.osr"]
    sys.modules["osgeo"].ogr = sys.modules["osgeo.ogr"]
    for pkg, sub in (("karios", ""), ("karios.core", "core"), ("karios.matcher", "matcher")):
        mod = types.ModuleType(pkg)
        mod.__path__ = [os.path.join(REF_ROOT, "karios", sub)]
        mod._b200_stub = True
        sys.modules[pkg] = mod
    klt = importlib.import_module("karios.matcher.klt")
    zs = importlib.import_module("karios.matcher.zncc_service")
    cfg = importlib.import_module("karios.core.configuration")
    return klt, zs, cfg


def load_mutual_info():
    """-> the unmodified karios/matcher/mutual_info_service.py module."""
    load()
    return importlib.import_module("karios.matcher.mutual_info_service")


def load_core_image():
    """-> the unmodified karios/core/image.py module (shift_image)."""
    load()
    return importlib.import_module("karios.core.image")


class ArrayImage:
    """Duck-typed stand-in for GdalRasterImage (karios/core/image.py:255) over an
    in-memory array: .read/.array/.x_size/.y_size/.no_data_value/.clear_cache."""

    def __init__(self, arr, no_data_value=None):
        self._a = arr
        self.no_data_value = no_data_value
        self.x_size = arr.shape[1]
        self.y_size = arr.shape[0]
        self.filepath = "memory"

    @property
    def array(self):
        return self._a

    def read(self, band, x_off, y_off, x_size, y_size):  # noqa: ARG002
        return self._a[y_off:y_off + y_size, x_off:x_off + x_size]

    def clear_cache(self):
        pass
