"""ctypes binding of include/fractencode_b200.h (the same entry points a cgo/JNI/C++ host binds)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# Frac2::UniformGridItem (image/partition2.hpp:93-99) and Frac::encode_item_t (encode/datatypes.h:8-23)
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

SEARCH_AUTO, SEARCH_EXACT, SEARCH_UMMA = 0, 1, 2


class Params(C.Structure):
    """fe_params: TransformMatcher(rmsThreshold, sMax) + classifier choice + FMA mode + engine."""
    _fields_ = [("rms_threshold", C.c_double), ("s_max", C.c_double), ("use_classifier", C.c_int32),
                ("fma", C.c_int32), ("search_impl", C.c_int32), ("isometries", C.c_int32)]

    def __init__(self, rms_threshold=0.0, s_max=-1.0, use_classifier=False, fma=False, search_impl=SEARCH_AUTO, isometries=4):
        super().__init__(float(rms_threshold), float(s_max), int(bool(use_classifier)), int(bool(fma)), int(search_impl), int(isometries))


class Stats(C.Structure):
    _fields_ = [("matches", C.c_uint64), ("kernel_launches", C.c_uint64), ("fp32_regime_items", C.c_uint64),
                ("umma_levels", C.c_uint64), ("exact_levels", C.c_uint64),
                ("level_items", C.c_uint64 * 8), ("level_ranges", C.c_uint64 * 8), ("level_matches", C.c_uint64 * 8),
                ("level_search_ms", C.c_float * 8), ("level_prep_ms", C.c_float * 8), ("last_decode_ms", C.c_float), ("reserved_", C.c_uint32),
                ("evaluated", C.c_uint64), ("level_evaluated", C.c_uint64 * 8), ("level_passes", C.c_uint64 * 8),
                ("prefiltered", C.c_uint64), ("level_prefiltered", C.c_uint64 * 8)]


class ThresholdPlan(C.Structure):
    """fe_threshold_plan: how a threshold search is pruned (host-only planning, no GPU needed)."""
    _fields_ = [("use_threshold", C.c_int32), ("thr16", C.c_uint32), ("radius", C.c_uint64), ("bin_width", C.c_uint32),
                ("n_bins", C.c_uint32), ("bin_span", C.c_uint32)]


def plan_threshold(rms_threshold: float, S: int, T: int) -> ThresholdPlan:
    pl = ThresholdPlan()
    rc = load_library().fe_plan_threshold(float(rms_threshold), int(S), int(T), C.byref(pl))
    if rc != 0:
        raise FractencodeError(rc, "fe_plan_threshold(%g, %d, %d)" % (rms_threshold, S, T))
    return pl


class FractencodeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fractencode_b200 error %d: %s" % (code, msg))
        self.code = code


def library_path() -> str:
    return os.path.join(HERE, "libfractencode_b200.so")


def build_library() -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-j4", "-C", os.path.join(HERE, "csrc")], check=True)
    return library_path()


_LIB = None
EXPORTS = [
    "fe_abi_version", "fe_create", "fe_destroy", "fe_last_error", "fe_set_image", "fe_set_images", "fe_set_image_device",
    "fe_classify", "fe_encode_level", "fe_encode_quadtree", "fe_encode_quadtree_device", "fe_encode_quadtree_slice_device", "fe_fetch_items", "fe_device_items", "fe_encode_batch", "fe_encode_planes", "fe_rgb_to_yuv420", "fe_yuv420_to_rgb",
    "fe_decode", "fe_copy_items", "fe_quantize", "fe_pack_items", "fe_unpack_items", "fe_items_minmax_device", "fe_pack_items_device", "fe_pack_errors", "fe_get_stats", "fe_stats_reset", "fe_synchronize", "fe_set_synthetic_image", "fe_get_image", "fe_plan_threshold",
]


def load_library():
    """Load the C-ABI library.  Fails loudly when it has not been built: there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise FractencodeError(-4, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(the CUDA extension is the product; there is no CPU path)" % path)
    lib = C.CDLL(path)
    vp, u32, i32, sz, dbl = C.c_void_p, C.c_uint32, C.c_int, C.c_size_t, C.c_double
    sig = {
        "fe_abi_version": (i32, []),
        "fe_create": (i32, [C.POINTER(vp), i32, vp]),
        "fe_destroy": (None, [vp]),
        "fe_last_error": (C.c_char_p, [vp]),
        "fe_set_image": (i32, [vp, vp, u32, u32, u32]),
        "fe_set_images": (i32, [vp, vp, u32, u32, u32, vp, u32, u32, u32]),
        "fe_set_image_device": (i32, [vp, vp, u32, u32, u32]),
        "fe_classify": (i32, [vp, i32, vp, sz, vp]),
        "fe_encode_level": (i32, [vp, vp, sz, vp, sz, C.POINTER(Params), vp]),
        "fe_encode_quadtree": (i32, [vp, u32, u32, C.POINTER(Params), vp, sz, C.POINTER(sz), vp]),
        "fe_encode_quadtree_device": (i32, [vp, u32, u32, C.POINTER(Params), C.POINTER(sz)]),
        "fe_encode_quadtree_slice_device": (i32, [vp, u32, u32, C.POINTER(Params), sz, sz, C.POINTER(sz)]),
        "fe_fetch_items": (i32, [vp, vp, sz, C.POINTER(sz)]),
        "fe_encode_batch": (i32, [vp, vp, sz, u32, u32, u32, u32, u32, C.POINTER(Params), vp, sz, vp]),
        "fe_encode_planes": (i32, [vp, vp, sz, vp, vp, vp, u32, u32, C.POINTER(Params), vp, vp, vp, vp]),
        "fe_rgb_to_yuv420": (i32, [vp, vp, u32, u32, u32, vp, u32, vp, u32, vp, u32, i32]),
        "fe_yuv420_to_rgb": (i32, [vp, vp, u32, u32, u32, vp, u32, vp, u32, vp, u32, i32]),
        "fe_device_items": (vp, [vp, C.POINTER(sz)]),
        "fe_decode": (i32, [vp, vp, sz, vp, u32, u32, u32, i32, dbl, i32, C.POINTER(C.c_int), C.POINTER(dbl)]),
        "fe_copy_items": (i32, [vp, vp, vp, u32, u32, u32, vp, sz, i32]),
        "fe_quantize": (i32, [vp, vp, sz, i32, i32, vp, vp, vp]),
        "fe_pack_items": (i32, [vp, vp, sz, u32, i32, i32, vp, vp]),
        "fe_unpack_items": (i32, [vp, vp, sz, u32, i32, i32, vp, i32, vp]),
        "fe_items_minmax_device": (i32, [vp, vp]),
        "fe_pack_items_device": (i32, [vp, u32, i32, i32, vp, vp, sz, C.POINTER(sz)]),
        "fe_pack_errors": (i32, [vp, C.POINTER(u32)]),
        "fe_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "fe_stats_reset": (i32, [vp]),
        "fe_synchronize": (i32, [vp]),
        "fe_set_synthetic_image": (i32, [vp, u32, u32, C.c_uint64, i32]),
        "fe_get_image": (i32, [vp, vp, u32]),
        "fe_plan_threshold": (i32, [C.c_double, u32, u32, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    _LIB = lib
    return lib


def uniform_grid(W: int, H: int, size: int, step: int) -> np.ndarray:
    """Host mirror of Frac2::createUniformGrid (image/partition2.hpp:110-135) for square items."""
    if size <= 0 or step <= 0 or W % size or H % size or W % step or H % step:
        raise ValueError("can't create grid partition on unaligned image")
    xs = np.arange(0, W - size + 1, step, dtype=np.uint32)
    ys = np.arange(0, H - size + 1, step, dtype=np.uint32)
    out = np.zeros(len(xs) * len(ys), GRID_ITEM)
    out["x"] = np.tile(xs, len(ys))
    out["y"] = np.repeat(ys, len(xs))
    out["w"] = size
    out["h"] = size
    out["bin"] = -1
    return out


class Context:
    """One fe_ctx: one device, one host thread (the reference's thread-per-engine model)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.fe_create(C.byref(h), device, stream)
        if rc != 0:
            raise FractencodeError(rc, self.lib.fe_last_error(None).decode())
        self.h = h
        self.size = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.fe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise FractencodeError(rc, self.lib.fe_last_error(self.h).decode())

    # ---- images ----
    def set_image(self, img: np.ndarray):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        self._check(self.lib.fe_set_image(self.h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0]))
        self.size = (img.shape[1], img.shape[0])

    def set_images(self, src: np.ndarray, tgt: np.ndarray):
        for a in (src, tgt):
            assert a.dtype == np.uint8 and a.ndim == 2 and a.strides[1] == 1
        self._check(self.lib.fe_set_images(self.h, src.ctypes.data, src.shape[1], src.shape[0], src.strides[0],
                                           tgt.ctypes.data, tgt.shape[1], tgt.shape[0], tgt.strides[0]))
        self.size = (src.shape[1], src.shape[0])

    def set_image_device(self, dev_ptr: int, W: int, H: int, stride: int):
        self._check(self.lib.fe_set_image_device(self.h, dev_ptr, W, H, stride))
        self.size = (W, H)

    def set_synthetic_image(self, W: int, H: int, seed: int = 1234, kind: int = 0):
        self._check(self.lib.fe_set_synthetic_image(self.h, W, H, seed, kind))
        self.size = (W, H)

    def get_image(self) -> np.ndarray:
        W, H = self.size
        out = np.zeros((H, W), np.uint8)
        self._check(self.lib.fe_get_image(self.h, out.ctypes.data, W))
        return out

    # ---- search ----
    def classify(self, items: np.ndarray, which: int = 0) -> np.ndarray:
        items = np.ascontiguousarray(items, GRID_ITEM)
        bins = np.zeros(len(items), np.int32)
        self._check(self.lib.fe_classify(self.h, which, items.ctypes.data, len(items), bins.ctypes.data))
        return bins

    def encode_level(self, domains: np.ndarray, ranges: np.ndarray, params: Params) -> np.ndarray:
        domains = np.ascontiguousarray(domains, GRID_ITEM)
        ranges = np.ascontiguousarray(ranges, GRID_ITEM)
        out = np.zeros(len(ranges), ENCODE_ITEM)
        self._check(self.lib.fe_encode_level(self.h, domains.ctypes.data, len(domains), ranges.ctypes.data, len(ranges),
                                             C.byref(params), out.ctypes.data))
        return out

    def encode_quadtree(self, t_max: int, t_min: int, params: Params):
        W, H = self.size
        cap = (W // t_min) * (H // t_min)
        out = np.zeros(cap, ENCODE_ITEM)
        n = C.c_size_t(0)
        counts = (C.c_size_t * 8)()
        self._check(self.lib.fe_encode_quadtree(self.h, t_max, t_min, C.byref(params), out.ctypes.data, cap, C.byref(n), counts))
        nlev = int(np.log2(t_max // t_min)) + 1
        return out[: n.value].copy(), [int(c) for c in counts[:nlev]]

    def encode_quadtree_device(self, t_max: int, t_min: int, params: Params) -> int:
        n = C.c_size_t(0)
        self._check(self.lib.fe_encode_quadtree_device(self.h, t_max, t_min, C.byref(params), C.byref(n)))
        return n.value

    def encode_quadtree_slice_device(self, t_max: int, t_min: int, params: Params, first_block: int, n_blocks: int) -> int:
        """One shard of the quadtree encode: top-level range blocks first_block..+n_blocks of the whole image's grid."""
        n = C.c_size_t(0)
        self._check(self.lib.fe_encode_quadtree_slice_device(self.h, t_max, t_min, C.byref(params), first_block, n_blocks, C.byref(n)))
        return n.value

    def encode_batch(self, images, t_max: int, t_min: int, params: Params, out: np.ndarray | None = None):
        """Quadtree-encode a batch of equally sized planes (list of 2-D uint8 arrays, or one [n, H, W] array), pipelined.
        Returns one transform list per image."""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        n = len(imgs)
        H, W = imgs[0].shape
        assert all(im.shape == (H, W) for im in imgs)
        cap = (W // t_min) * (H // t_min)
        if out is None:
            out = np.zeros(n * cap, ENCODE_ITEM)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        counts = (C.c_size_t * n)()
        self._check(self.lib.fe_encode_batch(self.h, ptrs, n, W, H, W, t_max, t_min, C.byref(params), out.ctypes.data, cap, counts))
        return [out[i * cap: i * cap + counts[i]] for i in range(n)]

    def encode_planes(self, planes, t_max: int, t_min: int, params: Params):
        """Quadtree-encode planes of different sizes (the colour path: Y and the two half-size chroma planes)."""
        pl = [np.ascontiguousarray(p, np.uint8) for p in planes]
        n = len(pl)
        caps = [(p.shape[1] // t_min) * (p.shape[0] // t_min) for p in pl]
        offs = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
        out = np.zeros(int(sum(caps)), ENCODE_ITEM)
        ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pl])
        ws = np.array([p.shape[1] for p in pl], np.uint32)
        hs = np.array([p.shape[0] for p in pl], np.uint32)
        capsa = np.array(caps, np.uint64)
        counts = (C.c_size_t * n)()
        self._check(self.lib.fe_encode_planes(self.h, ptrs, n, ws.ctypes.data, hs.ctypes.data, ws.ctypes.data, t_max, t_min, C.byref(params),
                                              out.ctypes.data, offs.ctypes.data, capsa.ctypes.data, counts))
        return [out[int(offs[i]): int(offs[i]) + counts[i]] for i in range(n)]

    def rgb_to_yuv420(self, rgb: np.ndarray, fma: bool = False):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        H, W, _ = rgb.shape
        y = np.zeros((H, W), np.uint8)
        u = np.zeros((H // 2, W // 2), np.uint8)
        v = np.zeros_like(u)
        self._check(self.lib.fe_rgb_to_yuv420(self.h, rgb.ctypes.data, W, H, 3 * W, y.ctypes.data, W, u.ctypes.data, W // 2, v.ctypes.data, W // 2, int(fma)))
        return y, u, v

    def yuv420_to_rgb(self, y: np.ndarray, u: np.ndarray, v: np.ndarray, fma: bool = False) -> np.ndarray:
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        H, W = y.shape
        rgb = np.zeros((H, W, 3), np.uint8)
        self._check(self.lib.fe_yuv420_to_rgb(self.h, y.ctypes.data, W, H, W, u.ctypes.data, W // 2, v.ctypes.data, W // 2, rgb.ctypes.data, W, int(fma)))
        return rgb

    def fetch_items(self, out: np.ndarray | None = None) -> np.ndarray:
        n = C.c_size_t(0)
        self.lib.fe_device_items(self.h, C.byref(n))
        if out is None:
            out = np.zeros(n.value, ENCODE_ITEM)
        self._check(self.lib.fe_fetch_items(self.h, out.ctypes.data, len(out), C.byref(n)))
        return out[: n.value]

    # ---- decode / quantize ----
    def decode(self, items: np.ndarray, W: int, H: int, stride: int | None = None, max_iters: int = -1, eps: float = 1e-5,
               fma: bool = False, init: int = 0):
        stride = stride or W
        items = np.ascontiguousarray(items, ENCODE_ITEM)
        tgt = np.full((H, stride), init, np.uint8)
        it, rms = C.c_int(0), C.c_double(0)
        self._check(self.lib.fe_decode(self.h, items.ctypes.data, len(items), tgt.ctypes.data, W, H, stride, max_iters, eps,
                                       int(fma), C.byref(it), C.byref(rms)))
        return tgt[:, :W], it.value, rms.value

    def copy_items(self, source: np.ndarray, target: np.ndarray, items: np.ndarray, fma: bool = False) -> np.ndarray:
        """One Decoder2::decodeStep: target[item] = clamp(trunc(s * sample(source) + o)); returns the new target."""
        items = np.ascontiguousarray(items, ENCODE_ITEM)
        source = np.ascontiguousarray(source, np.uint8)
        out = np.ascontiguousarray(target, np.uint8).copy()
        H, W = out.shape
        self._check(self.lib.fe_copy_items(self.h, source.ctypes.data, out.ctypes.data, W, H, W, items.ctypes.data, len(items), int(fma)))
        return out

    def quantize(self, items: np.ndarray, bits_s: int = 5, bits_o: int = 7):
        items = np.ascontiguousarray(items, ENCODE_ITEM)
        qs = np.zeros(len(items), np.uint32)
        qo = np.zeros(len(items), np.uint32)
        mm = np.zeros(4, np.float64)
        self._check(self.lib.fe_quantize(self.h, items.ctypes.data, len(items), bits_s, bits_o, qs.ctypes.data, qo.ctypes.data, mm.ctypes.data))
        return qs, qo, mm

    def pack_items(self, items: np.ndarray, t_max: int, bits_s: int = 5, bits_o: int = 7):
        """64-bit packed quantised records + the (min_s, max_s, min_o, max_o) header."""
        items = np.ascontiguousarray(items, ENCODE_ITEM)
        packed = np.zeros(len(items), np.uint64)
        mm = np.zeros(4, np.float64)
        self._check(self.lib.fe_pack_items(self.h, items.ctypes.data, len(items), t_max, bits_s, bits_o, packed.ctypes.data, mm.ctypes.data))
        return packed, mm

    def unpack_items(self, packed: np.ndarray, t_max: int, mm: np.ndarray, bits_s: int = 5, bits_o: int = 7, fma: bool = False) -> np.ndarray:
        packed = np.ascontiguousarray(packed, np.uint64)
        mm = np.ascontiguousarray(mm, np.float64)
        out = np.zeros(len(packed), ENCODE_ITEM)
        self._check(self.lib.fe_unpack_items(self.h, packed.ctypes.data, len(packed), t_max, bits_s, bits_o, mm.ctypes.data, int(fma), out.ctypes.data))
        return out

    # ---- device-resident post-pass (no synchronisation): minmax and packed records of the last quadtree result ----
    def items_minmax_device(self, minmax_dev_ptr: int):
        self._check(self.lib.fe_items_minmax_device(self.h, minmax_dev_ptr))

    def pack_items_device(self, t_max: int, minmax_dev_ptr: int, packed_dev_ptr: int, cap: int, bits_s: int = 5, bits_o: int = 7) -> int:
        n = C.c_size_t(0)
        self._check(self.lib.fe_pack_items_device(self.h, t_max, bits_s, bits_o, minmax_dev_ptr, packed_dev_ptr, cap, C.byref(n)))
        return n.value

    def pack_errors(self) -> int:
        n = C.c_uint32(0)
        self._check(self.lib.fe_pack_errors(self.h, C.byref(n)))
        return n.value

    # ---- misc ----
    def stats(self) -> Stats:
        s = Stats()
        self._check(self.lib.fe_get_stats(self.h, C.byref(s)))
        return s

    def stats_reset(self):
        self._check(self.lib.fe_stats_reset(self.h))

    def synchronize(self):
        self._check(self.lib.fe_synchronize(self.h))
