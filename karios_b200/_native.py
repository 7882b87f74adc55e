"""ctypes binding of libkarios_b200.so (C ABI: include/karios_b200.h).

There is NO fallback: if the CUDA library is missing or no CUDA device is
present, importing / using this module raises.  PyTorch is used only for device
memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# KR_LIB: another build of the same library (kernel A/B runs on one box, tools/build_variant.sh)
LIB_PATH = os.environ.get("KR_LIB") or os.path.join(_HERE, "_lib", "libkarios_b200.so")

KR_U8, KR_U16, KR_I16, KR_F32 = 0, 1, 2, 3
KR_TAIL_NONE, KR_TAIL_AVX512 = 0, 32

EXPORTS = [
    "kr_version", "kr_last_error", "kr_launch_count", "kr_ctx_create", "kr_ctx_destroy", "kr_read_stats",
    "kr_set_select_all", "kr_set_corner_mode", "kr_minmax_mask", "kr_u8_laplacian", "kr_corner_min_eigen_val",
    "kr_good_features", "kr_pyr_down", "kr_pyr_lk", "kr_klt_track", "kr_zncc", "kr_mutual_info",
    "kr_match_tile", "kr_auto_ksize", "kr_auto_ksize_scratch_bytes",
    "kr_set_profiling", "kr_read_stage_ms",
    "kr_unit_header_write", "kr_upload_pageable", "kr_cross_power", "kr_argmax_abs", "kr_shift_image", "kr_histogram", "kr_count_valid", "kr_histogram_count",
    "kr_gather_points",
]
NUM_STAGES = 12
STAGE_NAMES = ["minmax_mask", "laplacian_mon", "laplacian_ref", "corner_response", "select",
               "nms", "corner_sort", "pyramids", "lk_roundtrip", "rows", "zncc", "mutual_info"]


class KltConf(C.Structure):
    _fields_ = [("max_corners", C.c_int32), ("block_size", C.c_int32), ("win_size", C.c_int32),
                ("max_level", C.c_int32), ("max_count", C.c_int32), ("ksize_mon", C.c_int32),
                ("ksize_ref", C.c_int32), ("invert_mon", C.c_int32), ("tail_mode", C.c_int32),
                ("compute_zncc", C.c_int32), ("compute_mi", C.c_int32), ("quality_level", C.c_double),
                ("min_distance", C.c_double), ("eps", C.c_double),
                ("min_eig_threshold", C.c_double), ("back_threshold", C.c_double),
                ("zncc_min_score", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("min_a", C.c_double), ("max_a", C.c_double), ("min_b", C.c_double),
                ("max_b", C.c_double), ("valid", C.c_uint64), ("eig_max", C.c_float),
                ("n_candidates", C.c_uint32), ("n_above_threshold", C.c_uint32),
                ("n_sorted", C.c_uint32), ("n_corners", C.c_uint32), ("n_kept", C.c_uint32),
                ("nms_rounds", C.c_uint32), ("overflow", C.c_uint32),
                ("select_incomplete", C.c_uint32), ("two_tier", C.c_uint32),
                ("two_tier_fallback", C.c_uint32), ("n_border_maxima", C.c_uint32),
                ("n_exact", C.c_uint32), ("est_cut_bits", C.c_uint32), ("rows_skipped", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


KR_AUTO_MAX_K = 8


class AutoResult(C.Structure):
    """kr_auto_result: outcome of the on-device kernel-size search."""
    _fields_ = [("best_mon", C.c_int32), ("best_ref", C.c_int32), ("n_init", C.c_int32),
                ("n_kept", C.c_int32), ("redo", C.c_int32), ("n_k", C.c_int32), ("valid", C.c_uint64),
                ("counts", (C.c_int32 * 2) * (KR_AUTO_MAX_K * KR_AUTO_MAX_K))]


class Rows(C.Structure):
    _fields_ = [("x0", C.c_void_p), ("y0", C.c_void_p), ("dx", C.c_void_p), ("dy", C.c_void_p),
                ("score", C.c_void_p), ("zncc", C.c_void_p), ("mutual_info", C.c_void_p),
                ("mi", C.c_void_p), ("capacity", C.c_int32)]


class UnitHeader(C.Structure):
    """kr_unit_header: exchange record header of one unit (128 bytes)."""
    _fields_ = [("n_rows", C.c_int32), ("flags", C.c_int32), ("n", C.c_double), ("sum_dx", C.c_double),
                ("sum_dy", C.c_double), ("sum_dx2", C.c_double), ("sum_dy2", C.c_double),
                ("min_dx", C.c_double), ("min_dy", C.c_double), ("max_dx", C.c_double),
                ("max_dy", C.c_double), ("reserved", C.c_double * 6)]


_lib = None
_lib_lock = threading.Lock()


class KariosB200Error(RuntimeError):
    pass


def load_library(path: str = LIB_PATH):
    """dlopen the C-ABI library and declare every prototype."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(path):
            raise KariosB200Error(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  karios_b200 has no CPU fallback.")
        L = C.CDLL(path)
        vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.kr_version.restype = i32
        L.kr_last_error.restype = C.c_char_p
        L.kr_launch_count.restype = C.c_uint64
        L.kr_ctx_create.argtypes = [i32, i32, i32, i32, C.POINTER(vp)]
        L.kr_ctx_destroy.argtypes = [vp]
        L.kr_ctx_destroy.restype = None
        L.kr_read_stats.argtypes = [vp, vp, C.POINTER(Stats)]
        L.kr_set_select_all.argtypes = [vp, i32]
        L.kr_set_corner_mode.argtypes = [vp, i32]
        L.kr_set_profiling.argtypes = [vp, i32]
        L.kr_read_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.kr_minmax_mask.argtypes = [vp, vp, i64, vp, i64, i32, i32, i32, i32, f64, i32, f64, vp, i64, vp]
        L.kr_u8_laplacian.argtypes = [vp, vp, i64, i32, i32, i32, i32, i32, i32, vp, i64, vp]
        L.kr_corner_min_eigen_val.argtypes = [vp, vp, i64, i32, i32, i32, i32, vp, i64, vp]
        L.kr_good_features.argtypes = [vp, vp, i64, vp, i64, i32, i32, i32, f64, f64, i32, i32, vp,
                                       i32, vp, vp]
        L.kr_pyr_down.argtypes = [vp, vp, i64, i32, i32, vp, i64, vp]
        L.kr_pyr_lk.argtypes = [vp, vp, i64, vp, i64, i32, i32, vp, i32, vp, i32, i32, i32, f64, f64,
                                vp, vp, vp, vp]
        L.kr_klt_track.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, i32, C.POINTER(KltConf), vp,
                                   i32, Rows, vp]
        L.kr_zncc.argtypes = [vp, vp, i64, i32, i32, vp, i64, i32, i32, i32, vp, vp, vp, vp, i32, vp,
                              vp, vp]
        L.kr_mutual_info.argtypes = [vp, vp, i64, i32, i32, vp, i64, i32, i32, i32, vp, vp, vp, vp, i32,
                                     vp, vp, vp, vp]
        L.kr_match_tile.argtypes = [vp, vp, i64, vp, i64, i32, i32, i32, vp, i64, i32, i32, i32, i32,
                                    i32, f64, i32, f64, C.POINTER(KltConf), Rows, vp]
        L.kr_auto_ksize_scratch_bytes.argtypes = [i32, i32, i32, i32, i32, i32]
        L.kr_auto_ksize.argtypes = [vp, vp, i64, vp, i64, i32, i32, i32, vp, i64, i32, f64, i32, f64,
                                    C.POINTER(KltConf), C.POINTER(C.c_int32), i32, vp, i64, Rows, vp, vp]
        L.kr_unit_header_write.argtypes = [vp, Rows, vp, vp]
        L.kr_upload_pageable.argtypes = [vp, vp, i64, i32, vp]
        L.kr_cross_power.argtypes = [vp, vp, i64, i32, vp]
        L.kr_argmax_abs.argtypes = [vp, i64, i32, vp, vp, vp]
        L.kr_shift_image.argtypes = [vp, i64, vp, i64, i32, i32, i32, i32, i32, vp]
        L.kr_histogram.argtypes = [vp, i64, i32, i32, i32, i32, i32, i32, vp, vp]
        L.kr_count_valid.argtypes = [vp, i64, i32, i32, i32, vp, i64, vp, vp]
        L.kr_histogram_count.argtypes = [vp, i64, i32, i32, i32, i32, i32, i32, vp, vp, i64, vp, vp]
        L.kr_gather_points.argtypes = [vp, i64, i32, i32, i32, vp, vp, i32, vp, vp]
        for name in EXPORTS:
            if name not in ("kr_last_error", "kr_ctx_destroy", "kr_version", "kr_launch_count",
                            "kr_auto_ksize_scratch_bytes"):
                getattr(L, name).restype = i32
        L.kr_auto_ksize_scratch_bytes.restype = i64
        _lib = L
        return L


def _check(rc: int):
    if rc != 0:
        msg = load_library().kr_last_error().decode("utf-8", "replace")
        raise KariosB200Error(f"karios_b200 error {rc}: {msg}")


_DTYPES = {torch.uint8: KR_U8, torch.uint16: KR_U16, torch.int16: KR_I16, torch.float32: KR_F32}


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise KariosB200Error(f"raster dtype {t.dtype} not supported (uint8, uint16, int16, float32)")


def to_device(a, device) -> torch.Tensor:
    """numpy array / torch tensor -> 2-D CUDA tensor with unit column stride."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.asarray(a)
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.device.type != "cuda":
        dev = torch.device(device)
        big = t.numel() * t.element_size() >= (8 << 20)
        if big and t.is_contiguous() and dev.type == "cuda" and not t.is_pinned():
            # ordinary host memory: parallel staged upload (kr_upload_pageable)
            out = torch.empty(t.shape, dtype=t.dtype, device=dev)
            _check(load_library().kr_upload_pageable(out.data_ptr(), t.data_ptr(), t.numel() * t.element_size(),
                                                     dev.index if dev.index is not None else torch.cuda.current_device(),
                                                     torch.cuda.current_stream(dev).cuda_stream))
            t = out
        else:
            t = t.to(device, non_blocking=True)
    if t.dim() != 2:
        raise KariosB200Error(f"expected a 2-D raster, got shape {tuple(t.shape)}")
    if t.stride(1) != 1:
        t = t.contiguous()
    return t


def _pitch(t: torch.Tensor) -> int:
    return t.stride(0) * t.element_size()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def make_conf(conf, ksize_mon=None, ksize_ref=None, invert_mon=None, tail_mode=KR_TAIL_AVX512,
              compute_zncc=False, zncc_min_score=0.4, compute_mi=False) -> KltConf:
    """KLTConfiguration (karios/core/configuration.py:36-50) + the constants of
    klt.py:128-132,143 -> kr_klt_conf."""
    k = conf.laplacian_kernel_size
    if isinstance(k, dict):
        km = k.get("mon", k.get("ref", 1))
        kr = k.get("ref", k.get("mon", 1))
    elif isinstance(k, str):
        km = kr = 0          # "auto": the caller supplies explicit sizes
    else:
        km = kr = int(k)
    inv = conf.laplacian_invert_polarity
    c = KltConf()
    c.max_corners = int(conf.maxCorners)
    c.block_size = int(conf.blocksize)
    c.win_size = int(conf.matching_winsize)
    c.max_level = 1
    c.max_count = 30
    c.ksize_mon = int(km if ksize_mon is None else ksize_mon)
    c.ksize_ref = int(kr if ksize_ref is None else ksize_ref)
    c.invert_mon = int(bool(inv) if invert_mon is None and inv != "auto" else bool(invert_mon))
    c.tail_mode = int(tail_mode)
    c.compute_zncc = int(bool(compute_zncc))
    c.compute_mi = int(bool(compute_mi))
    c.quality_level = float(conf.qualityLevel)
    c.min_distance = float(conf.minDistance)
    c.eps = 0.03
    c.min_eig_threshold = 1e-4
    c.back_threshold = 0.1
    c.zncc_min_score = float(zncc_min_score)
    return c


class RowBuffers:
    """Device SoA for the rows of one tile (kr_rows)."""

    def __init__(self, capacity: int, device, with_zncc=True, with_mi=False):
        self.capacity = int(capacity)
        self.f32 = torch.empty((5, self.capacity), dtype=torch.float32, device=device)
        self.zncc = torch.empty(self.capacity, dtype=torch.float64, device=device) if with_zncc else None
        # [0] = mutual_info_score (Studholme), [1] = mi_score (api/core.py:894-907)
        self.mi = torch.empty((2, self.capacity), dtype=torch.float64, device=device) if with_mi else None

    @classmethod
    def from_views(cls, capacity: int, f32: torch.Tensor, zncc=None, mi=None):
        """Row buffers over existing device memory: f32 [5, >= capacity] (unit column
        stride, any row stride), zncc [>= capacity] float64, mi [2, >= capacity] float64."""
        self = cls.__new__(cls)
        self.capacity, self.f32, self.zncc, self.mi = int(capacity), f32, zncc, mi
        return self

    def struct(self) -> Rows:
        r = Rows()
        base, step = self.f32.data_ptr(), self.f32.stride(0) * 4
        r.x0, r.y0, r.dx, r.dy, r.score = (base + i * step for i in range(5))
        r.zncc = self.zncc.data_ptr() if self.zncc is not None else None
        r.mutual_info = self.mi[0].data_ptr() if self.mi is not None else None
        r.mi = self.mi[1].data_ptr() if self.mi is not None else None
        r.capacity = self.capacity
        return r


class Context:
    """One kr_ctx: scratch for tiles up to max_w x max_h on one device.  Not
    thread-safe (one per host thread)."""

    def __init__(self, max_w: int, max_h: int, max_corners: int, device=None):
        if not torch.cuda.is_available():
            raise KariosB200Error("no CUDA device: karios_b200 has no CPU fallback")
        self.lib = load_library()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.max_w, self.max_h, self.max_corners = int(max_w), int(max_h), int(max_corners)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _check(self.lib.kr_ctx_create(self.device.index or 0, self.max_w, self.max_h,
                                          self.max_corners, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.kr_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def fits(self, w, h, max_corners):
        return w <= self.max_w and h <= self.max_h and (
            (max_corners <= 0 and self.max_corners <= 0) or
            (0 < max_corners <= self.max_corners) or (max_corners > 0 and self.max_corners <= 0))

    # ---- thin wrappers, one per C entry point --------------------------------
    def read_stats(self) -> Stats:
        s = Stats()
        _check(self.lib.kr_read_stats(self._h, _stream(), C.byref(s)))
        if s.overflow:
            raise KariosB200Error("candidate list overflowed the context capacity "
                                  "(create the Context for a larger tile)")
        return s

    def set_profiling(self, on: bool):
        _check(self.lib.kr_set_profiling(self._h, int(bool(on))))

    def stage_ms(self):
        """Durations (ms) of the stages of the last match_tile (after a synchronise)."""
        buf = (C.c_float * NUM_STAGES)()
        _check(self.lib.kr_read_stage_ms(self._h, buf))
        return dict(zip(STAGE_NAMES, list(buf)))

    def set_select_all(self, on: bool):
        _check(self.lib.kr_set_select_all(self._h, int(bool(on))))

    def set_corner_mode(self, mode: int):
        """0: two-tier corner response (default); 1: OpenCV's arithmetic at every pixel."""
        _check(self.lib.kr_set_corner_mode(self._h, int(mode)))

    def minmax_mask(self, a, b=None, nodata_a=None, nodata_b=None, want_mask=False):
        h, w = a.shape
        mask = torch.empty((h, w), dtype=torch.uint8, device=a.device) if want_mask else None
        _check(self.lib.kr_minmax_mask(
            self._h, a.data_ptr(), _pitch(a), b.data_ptr() if b is not None else None,
            _pitch(b) if b is not None else 0, dtype_code(a), w, h,
            int(nodata_a is not None), float(nodata_a or 0), int(nodata_b is not None),
            float(nodata_b or 0), mask.data_ptr() if want_mask else None, w if want_mask else 0,
            _stream()))
        return mask

    def u8_laplacian(self, img, ksize, invert=False, slot=-1, out=None):
        h, w = img.shape
        if out is None:
            out = torch.empty((h, w), dtype=torch.uint8, device=img.device)
        _check(self.lib.kr_u8_laplacian(self._h, img.data_ptr(), _pitch(img), dtype_code(img), w, h,
                                        int(slot), int(ksize), int(bool(invert)), out.data_ptr(),
                                        _pitch(out), _stream()))
        return out

    def corner_min_eigen_val(self, u8, block, tail_mode=KR_TAIL_AVX512):
        h, w = u8.shape
        eig = torch.empty((h, w), dtype=torch.float32, device=u8.device)
        _check(self.lib.kr_corner_min_eigen_val(self._h, u8.data_ptr(), _pitch(u8), w, h, int(block),
                                                int(tail_mode), eig.data_ptr(), _pitch(eig), _stream()))
        return eig

    def good_features(self, u8, mask, max_corners, quality, min_distance, block,
                      tail_mode=KR_TAIL_AVX512):
        """-> CUDA tensor [N, 2] float32 in OpenCV order (N may be 0)."""
        h, w = u8.shape
        cap = int(max_corners) if max_corners > 0 else max(1, min(w * h, self._cap_unlimited(w, h)))
        out = torch.empty((cap, 2), dtype=torch.float32, device=u8.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=u8.device)
        for attempt in range(2):
            _check(self.lib.kr_good_features(
                self._h, u8.data_ptr(), _pitch(u8), mask.data_ptr() if mask is not None else None,
                _pitch(mask) if mask is not None else 0, w, h, int(max_corners), float(quality),
                float(min_distance), int(block), int(tail_mode), out.data_ptr(), cap, cnt.data_ptr(),
                _stream()))
            st = self.read_stats()
            if not st.select_incomplete or attempt == 1:
                break
            self.set_select_all(True)
        self.set_select_all(False)
        return out[: int(st.n_corners)]

    def _cap_unlimited(self, w, h):
        p = self.max_w * self.max_h
        return max(p // 8, 65536)

    def pyr_down(self, u8, out=None):
        h, w = u8.shape
        if out is None:
            out = torch.empty(((h + 1) // 2, (w + 1) // 2), dtype=torch.uint8, device=u8.device)
        _check(self.lib.kr_pyr_down(self._h, u8.data_ptr(), _pitch(u8), w, h, out.data_ptr(),
                                    _pitch(out), _stream()))
        return out

    def pyr_lk(self, prev, nxt, p0, win=25, max_level=1, max_count=30, eps=0.03, min_eig=1e-4):
        h, w = prev.shape
        p0 = p0.reshape(-1, 2).contiguous()
        n = p0.shape[0]
        p1 = torch.empty_like(p0)
        st = torch.empty(n, dtype=torch.uint8, device=p0.device)
        err = torch.empty(n, dtype=torch.float32, device=p0.device)
        if n:
            _check(self.lib.kr_pyr_lk(self._h, prev.data_ptr(), _pitch(prev), nxt.data_ptr(),
                                      _pitch(nxt), w, h, p0.data_ptr(), n, None, int(win),
                                      int(max_level), int(max_count), float(eps), float(min_eig),
                                      p1.data_ptr(), st.data_ptr(), err.data_ptr(), _stream()))
        return p1, st, err

    def klt_track(self, ref_u8, mon_u8, mask, kconf: KltConf, rows: RowBuffers, p0=None):
        """-> (n_init, n_kept) after one synchronisation."""
        h, w = ref_u8.shape
        n_p0 = 0
        if p0 is not None:
            p0 = p0.reshape(-1, 2).contiguous()
            n_p0 = p0.shape[0]
        for attempt in range(2):
            _check(self.lib.kr_klt_track(
                self._h, ref_u8.data_ptr(), _pitch(ref_u8), mon_u8.data_ptr(), _pitch(mon_u8),
                mask.data_ptr() if mask is not None else None,
                _pitch(mask) if mask is not None else 0, w, h, C.byref(kconf),
                p0.data_ptr() if p0 is not None else None, n_p0, rows.struct(), _stream()))
            st = self.read_stats()
            if p0 is not None or not st.select_incomplete or attempt == 1:
                break
            self.set_select_all(True)
        self.set_select_all(False)
        return int(st.n_corners), int(st.n_kept)

    def auto_ksize(self, mon, ref, mask, kconf: KltConf, ksizes, rows: RowBuffers,
                   nodata_mon=None, nodata_ref=None) -> AutoResult:
        """KLT._match_tile_auto_ksize (klt.py:465-545) on the device: one launch sequence,
        one synchronisation (the 536-byte result record).  mon / ref: raw tile windows."""
        h, w = mon.shape
        ks = (C.c_int32 * len(ksizes))(*[int(k) for k in ksizes])
        need = int(self.lib.kr_auto_ksize_scratch_bytes(w, h, len(ksizes), int(kconf.max_corners),
                                                        int(kconf.win_size), int(kconf.max_level)))
        if need <= 0:
            raise KariosB200Error("kr_auto_ksize: unsupported configuration")
        scratch = getattr(self, "_auto_scratch", None)
        if scratch is None or scratch.numel() < need:
            scratch = None
            self._auto_scratch = None
            scratch = torch.empty(need, dtype=torch.uint8, device=mon.device)
            self._auto_scratch = scratch
        rec = torch.zeros(C.sizeof(AutoResult), dtype=torch.uint8, device=mon.device)
        _check(self.lib.kr_auto_ksize(
            self._h, mon.data_ptr(), _pitch(mon), ref.data_ptr(), _pitch(ref), dtype_code(mon), w, h,
            mask.data_ptr() if mask is not None else None, _pitch(mask) if mask is not None else 0,
            int(nodata_mon is not None), float(nodata_mon or 0.0), int(nodata_ref is not None),
            float(nodata_ref or 0.0), C.byref(kconf), ks, len(ksizes), scratch.data_ptr(), scratch.numel(),
            rows.struct(), rec.data_ptr(), _stream()))
        out = AutoResult.from_buffer_copy(rec.cpu().numpy().tobytes())
        return out

    def zncc(self, ref, mon, x0, y0, dx, dy):
        n = x0.shape[0]
        out = torch.empty(n, dtype=torch.float64, device=ref.device)
        if n:
            _check(self.lib.kr_zncc(self._h, ref.data_ptr(), _pitch(ref), ref.shape[1], ref.shape[0],
                                    mon.data_ptr(), _pitch(mon), mon.shape[1], mon.shape[0],
                                    dtype_code(ref), x0.data_ptr(), y0.data_ptr(), dx.data_ptr(),
                                    dy.data_ptr(), n, None, out.data_ptr(), _stream()))
        return out

    def mutual_info(self, ref, mon, x0, y0, dx, dy):
        """-> [2, n] float64: row 0 = Studholme NMI (MutualInfoService), row 1 =
        2 MI / (Hx + Hy) (ZNCCService.compute_mi)."""
        n = x0.shape[0]
        out = torch.empty((2, n), dtype=torch.float64, device=ref.device)
        if n:
            _check(self.lib.kr_mutual_info(self._h, ref.data_ptr(), _pitch(ref), ref.shape[1],
                                           ref.shape[0], mon.data_ptr(), _pitch(mon), mon.shape[1],
                                           mon.shape[0], dtype_code(ref), x0.data_ptr(), y0.data_ptr(),
                                           dx.data_ptr(), dy.data_ptr(), n, None, out[0].data_ptr(),
                                           out[1].data_ptr(), _stream()))
        return out

    def match_tile_async(self, mon, ref, mask, window, kconf: KltConf, rows: RowBuffers,
                         nodata_mon=None, nodata_ref=None):
        """Enqueue KLT._match_tile for window (x_off, y_off, w, h) of full rasters;
        no synchronisation (read the counts later with read_stats)."""
        x_off, y_off, tw, th = window
        ih, iw = mon.shape
        _check(self.lib.kr_match_tile(
            self._h, mon.data_ptr(), _pitch(mon), ref.data_ptr(), _pitch(ref), dtype_code(mon), iw, ih,
            mask.data_ptr() if mask is not None else None, _pitch(mask) if mask is not None else 0,
            int(x_off), int(y_off), int(tw), int(th), int(nodata_mon is not None),
            float(nodata_mon or 0), int(nodata_ref is not None), float(nodata_ref or 0),
            C.byref(kconf), rows.struct(), _stream()))

    def unit_header(self, rows: RowBuffers, header: torch.Tensor):
        """Enqueue kr_unit_header_write: count + dx / dy moments of the unit just matched into
        `header` (32 float32 words of an exchange record, karios_b200/sharding.py)."""
        _check(self.lib.kr_unit_header_write(self._h, rows.struct(), header.data_ptr(), _stream()))

    def match_tile(self, mon, ref, mask, window, kconf, rows, nodata_mon=None, nodata_ref=None):
        """match_tile_async + stats; re-runs once with every candidate when the
        corner pre-selection was too small.  -> Stats"""
        for attempt in range(2):
            self.match_tile_async(mon, ref, mask, window, kconf, rows, nodata_mon, nodata_ref)
            st = self.read_stats()
            if not st.select_incomplete or attempt == 1:
                break
            self.set_select_all(True)
        self.set_select_all(False)
        return st


# ---- context-free entry points (full-frame passes around the matching path) ------
def _require_cuda():
    if not torch.cuda.is_available():
        raise KariosB200Error("no CUDA device: karios_b200 has no CPU fallback")
    return load_library()


def cross_power_(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """In place: a <- a conj(b) / max(|a conj(b)|, 100 eps) on complex CUDA tensors."""
    lib = _require_cuda()
    if a.dtype != b.dtype or a.shape != b.shape or a.dtype not in (torch.complex64, torch.complex128):
        raise KariosB200Error("cross_power_ needs two complex tensors of one dtype and shape")
    if not (a.is_contiguous() and b.is_contiguous()):
        raise KariosB200Error("cross_power_ needs contiguous tensors")
    _check(lib.kr_cross_power(a.data_ptr(), b.data_ptr(), a.numel(), int(a.dtype == torch.complex128),
                              _stream()))
    return a


def argmax_abs(x: torch.Tensor) -> torch.Tensor:
    """First index of max |x| (np.argmax(np.abs(x))) -> device int64 tensor [1]."""
    lib = _require_cuda()
    if x.dtype not in (torch.float32, torch.float64) or not x.is_contiguous():
        raise KariosB200Error("argmax_abs needs a contiguous float32 / float64 tensor")
    scratch = torch.empty(2, dtype=torch.int64, device=x.device)
    out = torch.empty(1, dtype=torch.int64, device=x.device)
    _check(lib.kr_argmax_abs(x.data_ptr(), x.numel(), int(x.dtype == torch.float64), scratch.data_ptr(),
                             out.data_ptr(), _stream()))
    return out


def shift_image(img: torch.Tensor, y_off=0, x_off=0) -> torch.Tensor:
    """shift_image (karios/core/image.py:70-101) on a CUDA raster: a new tensor."""
    lib = _require_cuda()
    y_off, x_off = int(round(y_off)), int(round(x_off))
    h, w = img.shape
    out = torch.empty_like(img)
    _check(lib.kr_shift_image(img.data_ptr(), _pitch(img), out.data_ptr(), _pitch(out), dtype_code(img),
                              w, h, x_off, y_off, _stream()))
    return out


def histogram(img: torch.Tensor, lo: int, hi: int, shift: int = 0) -> torch.Tensor:
    """Counts of (value - lo) >> shift for the integer values lo..hi (inclusive)
    -> device int64 tensor [((hi - lo) >> shift) + 1]."""
    lib = _require_cuda()
    h, w = img.shape
    n = ((int(hi) - int(lo)) >> shift) + 1
    hist = torch.zeros(n, dtype=torch.int64, device=img.device)
    for start in range(0, n, 8192):
        nb = min(8192, n - start)
        _check(lib.kr_histogram(img.data_ptr(), _pitch(img), dtype_code(img), w, h,
                                int(lo) + (start << shift), int(shift), nb, hist[start:].data_ptr(),
                                _stream()))
    return hist


def histogram_count(img: torch.Tensor, lo: int, hi: int, shift: int = 0, mask=None):
    """histogram() and count_valid() in one pass (at most 8192 bins) -> (hist int64 tensor, count int)."""
    lib = _require_cuda()
    h, w = img.shape
    n = ((int(hi) - int(lo)) >> shift) + 1
    if n > 8192:
        raise KariosB200Error("histogram_count: at most 8192 bins in one pass")
    hist = torch.zeros(n, dtype=torch.int64, device=img.device)
    out = torch.zeros(1, dtype=torch.int64, device=img.device)
    _check(lib.kr_histogram_count(img.data_ptr(), _pitch(img), dtype_code(img), w, h, int(lo), int(shift), n,
                                  hist.data_ptr(), mask.data_ptr() if mask is not None else None,
                                  _pitch(mask) if mask is not None else 0, out.data_ptr(), _stream()))
    return hist, int(out.item())


def count_valid(img: torch.Tensor, mask=None) -> int:
    lib = _require_cuda()
    h, w = img.shape
    out = torch.zeros(1, dtype=torch.int64, device=img.device)
    _check(lib.kr_count_valid(img.data_ptr(), _pitch(img), dtype_code(img), w, h,
                              mask.data_ptr() if mask is not None else None,
                              _pitch(mask) if mask is not None else 0, out.data_ptr(), _stream()))
    return int(out.item())


def gather_points(img: torch.Tensor, x0: torch.Tensor, y0: torch.Tensor) -> torch.Tensor:
    """img[int(y0), int(x0)] as float64 (NaN outside the raster)."""
    lib = _require_cuda()
    h, w = img.shape
    n = x0.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=img.device)
    if n:
        _check(lib.kr_gather_points(img.data_ptr(), _pitch(img), dtype_code(img), w, h, x0.data_ptr(),
                                    y0.data_ptr(), n, out.data_ptr(), _stream()))
    return out
