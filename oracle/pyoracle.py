"""ctypes front-end for the TEST-ONLY checkers in oracle/ (see frac_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  `restatement()` loads our C restatement
(libfrac_oracle.so, prefix fo_); `reference(fma)` loads the REAL reference compiled
from /root/reference into oracle/_ref/ (prefix fr_) or returns None when absent.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

GRID_ITEM = np.dtype([("x", "<u4"), ("y", "<u4"), ("w", "<u4"), ("h", "<u4"), ("bin", "<i4")])
ENCODE_ITEM = np.dtype(
    [
        ("x", "<u4"), ("y", "<u4"), ("w", "<u4"), ("h", "<u4"),
        ("distance", "<f8"), ("contrast", "<f8"), ("brightness", "<f8"),
        ("transform", "<i4"), ("pad", "<i4"),
        ("match_x", "<u4"), ("match_y", "<u4"), ("src_w", "<u4"), ("src_h", "<u4"),
    ]
)
assert GRID_ITEM.itemsize == 20 and ENCODE_ITEM.itemsize == 64


class Plane(C.Structure):
    _fields_ = [("px", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("stride", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("rms_threshold", C.c_double), ("s_max", C.c_double), ("use_classifier", C.c_int), ("fma", C.c_int),
                ("isometries", C.c_int), ("reserved_", C.c_int)]


def build(ref: bool = True) -> None:
    """Compile the checkers (gcc; and the reference when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)


def _plane(img: np.ndarray) -> Plane:
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    return Plane(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0])


class Oracle:
    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.path = path
        u32, i32, dbl, sz, vp = C.c_uint32, C.c_int, C.c_double, C.c_size_t, C.c_void_p
        PP, PR = C.POINTER(Plane), C.POINTER(Params)
        self._sig("sample_sum4", i32, [PP, u32, u32, u32, u32, u32, u32, i32])
        self._sig("block_sum", u32, [PP, u32, u32, u32, u32])
        self._sig("distance", dbl, [PP, PP] + [u32] * 8 + [i32])
        self._sig("category4", i32, [dbl] * 4)
        self._sig("category", i32, [PP, u32, u32, u32, u32])
        self._sig("create_uniform_grid", sz, [u32] * 6 + [vp, sz])
        self._sig("preclassify", None, [PP, vp, sz])
        self._sig("match", None, [PP, vp, PP, vp, PR, vp])
        self._sig("estimate", None, [PP, PP, vp, sz, vp, PR, vp])
        self._sig("encode_level", None, [PP, PP, vp, sz, vp, sz, PR, i32, sz, vp])
        self._sig("encode_quadtree", sz, [PP, u32, u32, PR, i32, vp, sz, vp])
        self._sig("decode", None, [vp, sz, vp, u32, u32, u32, i32, dbl, i32, C.POINTER(C.c_int), C.POINTER(C.c_double)])
        self._sig("quantize", C.c_uint64, [dbl, dbl, dbl, i32])
        self._sig("dequantize", dbl, [C.c_uint64, dbl, dbl, i32])
        self._sig("rgb2yuv", None, [vp, u32, u32, u32, vp, u32, vp, u32, vp, u32, i32])
        self._sig("yuv2rgb", None, [vp, u32, u32, u32, vp, u32, vp, u32, vp, u32, i32])
        self._sig("hardware_threads", i32, [])
        self._sig("version", C.c_char_p, [])
        if prefix == "fo_":
            self._sig("synth_image", None, [vp, u32, u32, u32, C.c_uint64, i32])
        if prefix == "fr_":
            self._sig("load_luma", i32, [C.c_char_p, vp, sz, C.POINTER(u32), C.POINTER(u32)])

    def _sig(self, name, res, args):
        f = getattr(self.lib, self.prefix + name)
        f.restype = res
        f.argtypes = args
        setattr(self, "_" + name, f)

    # ---- thin pythonic wrappers -------------------------------------------------
    def version(self) -> str:
        return self._version().decode()

    def hardware_threads(self) -> int:
        return self._hardware_threads()

    def sample_sum4(self, img, px, py, pw, ph, lx, ly, t):
        return self._sample_sum4(C.byref(_plane(img)), px, py, pw, ph, lx, ly, t)

    def block_sum(self, img, x, y, w, h):
        return self._block_sum(C.byref(_plane(img)), x, y, w, h)

    def distance(self, a, b, sa, sb, t):
        return self._distance(C.byref(_plane(a)), C.byref(_plane(b)), *sa, *sb, t)

    def category4(self, a1, a2, a3, a4):
        return self._category4(a1, a2, a3, a4)

    def category(self, img, x, y, w, h):
        return self._category(C.byref(_plane(img)), x, y, w, h)

    def uniform_grid(self, W, H, size, step) -> np.ndarray:
        n = self._create_uniform_grid(W, H, size, size, step, step, None, 0)
        out = np.zeros(n, GRID_ITEM)
        if n:
            self._create_uniform_grid(W, H, size, size, step, step, out.ctypes.data, n)
        return out

    def uniform_grid_xy(self, W, H, sx, sy, ox, oy) -> np.ndarray:
        n = self._create_uniform_grid(W, H, sx, sy, ox, oy, None, 0)
        out = np.zeros(n, GRID_ITEM)
        if n:
            self._create_uniform_grid(W, H, sx, sy, ox, oy, out.ctypes.data, n)
        return out

    def preclassify(self, img, items: np.ndarray) -> np.ndarray:
        items = np.ascontiguousarray(items, GRID_ITEM).copy()
        self._preclassify(C.byref(_plane(img)), items.ctypes.data, len(items))
        return items

    @staticmethod
    def params(thr=0.0, smax=-1.0, classifier=False, fma=False, isometries=4) -> Params:
        return Params(float(thr), float(smax), int(bool(classifier)), int(bool(fma)), int(isometries), 0)

    def match(self, src, dom, tgt, rng, p: Params) -> np.ndarray:
        d = np.array([tuple(dom) + (-1,)], GRID_ITEM)
        r = np.array([tuple(rng) + (-1,)], GRID_ITEM)
        out = np.zeros(1, ENCODE_ITEM)
        self._match(C.byref(_plane(src)), d.ctypes.data, C.byref(_plane(tgt)), r.ctypes.data, C.byref(p), out.ctypes.data)
        return out[0]

    def encode_level(self, src, tgt, domains, ranges, p: Params, nthreads=0, sample_stride=1) -> np.ndarray:
        domains = np.ascontiguousarray(domains, GRID_ITEM)
        ranges = np.ascontiguousarray(ranges, GRID_ITEM)
        out = np.zeros(len(ranges), ENCODE_ITEM)
        self._encode_level(C.byref(_plane(src)), C.byref(_plane(tgt)), domains.ctypes.data, len(domains),
                           ranges.ctypes.data, len(ranges), C.byref(p), nthreads, sample_stride, out.ctypes.data)
        return out

    def encode_quadtree(self, img, t_max, t_min, p: Params, nthreads=0):
        cap = (img.shape[0] // t_min) * (img.shape[1] // t_min)
        out = np.zeros(cap, ENCODE_ITEM)
        counts = np.zeros(16, np.uint64)
        n = self._encode_quadtree(C.byref(_plane(img)), t_max, t_min, C.byref(p), nthreads, out.ctypes.data, cap, counts.ctypes.data)
        if n == C.c_size_t(-1).value:
            raise RuntimeError("oracle quadtree failed (alignment or capacity)")
        nlev = int(np.log2(t_max // t_min)) + 1
        return out[:n].copy(), [int(c) for c in counts[:nlev]]

    def decode(self, items, W, H, stride=None, max_iters=-1, eps=1e-5, fma=False, init=0):
        stride = stride or W
        items = np.ascontiguousarray(items, ENCODE_ITEM)
        tgt = np.full((H, stride), init, np.uint8)
        it, rms = C.c_int(0), C.c_double(0)
        self._decode(items.ctypes.data, len(items), tgt.ctypes.data, W, H, stride, max_iters, eps, int(fma), C.byref(it), C.byref(rms))
        return tgt[:, :W], it.value, rms.value

    def rgb2yuv(self, rgb: np.ndarray, fma=False):
        """[H, W, 3] uint8 -> (Y [H, W], U, V [ceil(H/2), ceil(W/2)]) like ImageIO::rgb2yuv (tightly packed planes)."""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        H, W, _ = rgb.shape
        y = np.zeros((H, W), np.uint8)
        u = np.zeros(((H + 1) // 2, (W + 1) // 2), np.uint8)
        v = np.zeros_like(u)
        self._rgb2yuv(rgb.ctypes.data, W, H, W * 3, y.ctypes.data, W, u.ctypes.data, u.shape[1], v.ctypes.data, v.shape[1], int(fma))
        return y, u, v

    def yuv2rgb(self, y: np.ndarray, u: np.ndarray, v: np.ndarray, fma=False) -> np.ndarray:
        H, W = y.shape
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        rgb = np.zeros((H, W, 3), np.uint8)
        self._yuv2rgb(y.ctypes.data, W, H, W, u.ctypes.data, u.shape[1], v.ctypes.data, v.shape[1], rgb.ctypes.data, W, int(fma))
        return rgb

    def quantize(self, v, vmin, vmax, bits):
        return self._quantize(v, vmin, vmax, bits)

    def dequantize(self, q, vmin, vmax, bits):
        return self._dequantize(q, vmin, vmax, bits)

    def synth_image(self, W, H, seed=1234, kind=0, stride=None) -> np.ndarray:
        stride = stride or W
        buf = np.zeros((H, stride), np.uint8)
        self._synth_image(buf.ctypes.data, W, H, stride, seed, kind)
        return buf[:, :W]

    def load_luma(self, path: str) -> np.ndarray:
        w, h = C.c_uint32(0), C.c_uint32(0)
        buf = np.zeros(1 << 24, np.uint8)
        rc = self._load_luma(path.encode(), buf.ctypes.data, buf.size, C.byref(w), C.byref(h))
        assert rc == 0
        return buf[: w.value * h.value].reshape(h.value, w.value).copy()


def restatement() -> Oracle:
    path = os.path.join(HERE, "libfrac_oracle.so")
    if not os.path.exists(path):
        build(ref=False)
    return Oracle(path, "fo_")


def reference(fma: bool = False):
    """The real reference (oracle/_ref), or None when it was never built here."""
    path = os.path.join(HERE, "_ref", "libfracref_fma.so" if fma else "libfracref_nofma.so")
    if not os.path.exists(path):
        return None
    return Oracle(path, "fr_")


def sort_items(items: np.ndarray) -> np.ndarray:
    """Canonical order for comparing transform lists: by (y, x, w, h)."""
    return items[np.lexsort((items["h"], items["w"], items["x"], items["y"]))]


def dump_lines(items: np.ndarray) -> str:
    """SURVEY 8c dump format, one line per item, sorted by (y, x)."""
    out = []
    for e in sort_items(items):
        bits = lambda v: np.float64(v).view(np.uint64)
        out.append("%u %u %u %u | %u %u %u %u | t=%d d=%016x s=%016x o=%016x" % (
            e["x"], e["y"], e["w"], e["h"], e["match_x"], e["match_y"], e["src_w"], e["src_h"],
            e["transform"], bits(e["distance"]), bits(e["contrast"]), bits(e["brightness"])))
    return "\n".join(out) + "\n"
