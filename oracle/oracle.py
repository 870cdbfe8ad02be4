"""ctypes binding of the CPU oracle (oracle/ttmlblend_ref.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see ttmlblend_ref.h). Importable
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs only; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FORMATS = {"I420": 0, "NV12": 1, "AYUV": 2, "RGBA": 3, "BGRA": 4,
           "YV12": 5, "NV21": 6, "ARGB": 7, "ABGR": 8,
           # x formats use their alpha twins' pack/unpack in GStreamer (PACK_RGBA ...)
           "RGBx": 3, "BGRx": 4, "xRGB": 7, "xBGR": 8,
           "Y42B": 13, "Y444": 14, "YUY2": 15, "UYVY": 16, "GRAY8": 17, "NV16": 18, "NV24": 19,
           "NV61": 20, "YVYU": 21, "VYUY": 22, "v308": 23, "IYU2": 24, "RGB": 25, "BGR": 26}
FLAG_PREMULTIPLIED_ALPHA = 1


class RefFrame(C.Structure):
    _fields_ = [("format", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("flags", C.c_uint32), ("data", C.c_void_p * 3), ("stride", C.c_int32 * 3)]


class RefRectangle(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32),
                ("stride", C.c_int32), ("x", C.c_int32), ("y", C.c_int32),
                ("global_alpha", C.c_float), ("flags", C.c_uint32),
                ("render_width", C.c_int32), ("render_height", C.c_int32)]


class RefRegion(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("background_color", C.c_uint32), ("opacity", C.c_double),
                ("layer", C.c_void_p), ("layer_stride", C.c_int32)]


def build(native: bool = False, out_dir: str = None) -> str:
    """Compiles the oracle. native=True adds -march=native (CPU baseline on the box it runs on)."""
    out_dir = out_dir or _HERE
    name = "libttmlblend_ref_native.so" if native else "libttmlblend_ref.so"
    out = os.path.join(out_dir, name)
    src = os.path.join(_HERE, "ttmlblend_ref.c")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src) and not native:
        return out
    flags = ["-O3", "-fPIC", "-shared", "-std=c99", "-D_POSIX_C_SOURCE=200809L"]
    if native:
        flags.append("-march=native")
    subprocess.check_call(["gcc", *flags, "-o", out, src, "-lpthread", "-lm"])
    return out


_libs = {}


def load(native: bool = False, out_dir: str = None):
    key = (native, out_dir)
    if key in _libs:
        return _libs[key]
    lib = C.CDLL(build(native, out_dir))
    lib.tbref_video_blend.restype = C.c_int
    lib.tbref_video_blend.argtypes = [C.POINTER(RefFrame), C.POINTER(RefRectangle)]
    lib.tbref_composition_blend.restype = C.c_int
    lib.tbref_composition_blend.argtypes = [C.POINTER(RefFrame), C.POINTER(RefRectangle), C.c_uint32]
    lib.tbref_blend_many.restype = C.c_double
    lib.tbref_blend_many.argtypes = [C.POINTER(RefFrame), C.c_uint32, C.POINTER(RefRectangle),
                                     C.c_uint32, C.c_uint32]
    lib.tbref_gaussian_kernel.restype = C.c_int32
    lib.tbref_gaussian_kernel.argtypes = [C.c_int32, C.c_double, C.POINTER(C.c_int32)]
    lib.tbref_blur_argb32.restype = None
    lib.tbref_blur_argb32.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_double, C.c_void_p, C.c_int32]
    lib.tbref_scale_linear_rgba.restype = None
    lib.tbref_scale_linear_rgba.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_void_p]
    for n in ("tbref_matrix_prea_rgb_to_yuv", "tbref_matrix_rgb_to_yuv", "tbref_matrix_yuv_to_rgb"):
        getattr(lib, n).restype = None
        getattr(lib, n).argtypes = [C.c_void_p, C.c_uint32]
    _libs[key] = lib
    return lib


def make_frame(fmt: str, width: int, height: int, planes: Sequence[np.ndarray],
               premultiplied: bool = False) -> RefFrame:
    f = RefFrame()
    f.format = FORMATS[fmt]
    f.width, f.height = width, height
    f.flags = FLAG_PREMULTIPLIED_ALPHA if premultiplied else 0
    for i, p in enumerate(planes):
        assert p.dtype == np.uint8 and p.ndim == 2 and p.strides[1] == 1
        f.data[i] = p.ctypes.data
        f.stride[i] = p.strides[0]
    return f


def make_rectangles(rectangles: Sequence[dict]):
    """Same dicts as TtmlBlend.overlay_set_rectangles."""
    arr = (RefRectangle * max(1, len(rectangles)))()
    for i, r in enumerate(rectangles):
        px = r["pixels"]
        assert px.dtype == np.uint8 and px.ndim == 3 and px.shape[2] == 4 and px.strides[2] == 1
        arr[i] = RefRectangle(px.ctypes.data, px.shape[1], px.shape[0], px.strides[0],
                              int(r.get("x", 0)), int(r.get("y", 0)),
                              float(r.get("global_alpha", 1.0)),
                              FLAG_PREMULTIPLIED_ALPHA if r.get("premultiplied", True) else 0,
                              int(r.get("render_width", 0)), int(r.get("render_height", 0)))
    return arr


def composition_blend(fmt: str, width: int, height: int, planes: Sequence[np.ndarray],
                      rectangles: Sequence[dict], premultiplied_dest: bool = False, lib=None):
    """gst_video_overlay_composition_blend on `planes`, in place. Returns planes."""
    lib = lib or load()
    f = make_frame(fmt, width, height, planes, premultiplied_dest)
    arr = make_rectangles(rectangles)
    ok = lib.tbref_composition_blend(C.byref(f), arr, len(rectangles))
    if not ok and rectangles:
        raise RuntimeError("tbref_composition_blend returned FALSE")
    return planes


def ttmlrender_rectangles(bgra: np.ndarray, rects: Sequence[Sequence[int]] = ()):
    """What the reference pipeline blends: ttmlrender's frame-sized image as ONE
    premultiplied rectangle at (0,0); `rects` is ignored on purpose (the region boxes only
    tell the GPU path where non-transparent pixels can be)."""
    return [dict(pixels=bgra, x=0, y=0, global_alpha=1.0, premultiplied=True)]


def scale_linear_rgba(img: np.ndarray, dest_width: int, dest_height: int, lib=None) -> np.ndarray:
    """gst_video_blend_scale_linear_RGBA on an h x w x 4 uint8 image (h, w >= 2)."""
    lib = lib or load()
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 4 and img.strides[2] == 1
    assert img.shape[0] >= 2 and img.shape[1] >= 2
    out = np.empty((dest_height, dest_width, 4), np.uint8)
    lib.tbref_scale_linear_rgba(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                                dest_width, dest_height, out.ctypes.data)
    return out


REF_PATH = os.path.join(_HERE, "_ref", "libttmlblur_ref.so")
_ref = None


def load_ref():
    """oracle/_ref/libttmlblur_ref.so: the reference's own gstttmlblur.c compiled from
    /root/reference against the stand-in headers of oracle/refstub/ (`make -C oracle ref`,
    done by __graft_entry__.build() where /root/reference exists). None if it is not there."""
    global _ref
    if _ref is None and os.path.exists(REF_PATH):
        lib = C.CDLL(REF_PATH)
        lib.ttmlref_blur_argb32.restype = None
        lib.ttmlref_blur_argb32.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_double, C.c_void_p, C.c_int32]
        lib.ttmlref_last_filter_params.restype = C.c_int32
        lib.ttmlref_last_filter_params.argtypes = [C.POINTER(C.c_int32), C.c_int32]
        _ref = lib
    return _ref


def ref_blur_argb32(img: np.ndarray, radius: int, sigma: float):
    """(blurred image, filter parameters) from the reference's gst_ttml_blur_image_surface: the
    parameters are what it handed to pixman_image_set_filter (2 sizes + (2r+1)^2 taps, 16.16)."""
    lib = load_ref()
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 4 and img.strides[2] == 1
    out = np.zeros_like(img)
    lib.ttmlref_blur_argb32(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], radius, sigma,
                            out.ctypes.data, out.strides[0])
    n = (2 * radius + 1) ** 2 + 2
    buf = (C.c_int32 * n)()
    got = lib.ttmlref_last_filter_params(buf, n)
    assert got == n, (got, n)
    return out, np.array(buf[:], dtype=np.int64)


def blur_argb32(img: np.ndarray, radius: int, sigma: float, lib=None) -> np.ndarray:
    """gst_ttml_blur_image_surface (surface, radius, sigma) on an h x w x 4 uint8 image."""
    lib = lib or load()
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 4 and img.strides[2] == 1
    out = np.zeros_like(img)
    lib.tbref_blur_argb32(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], radius,
                          float(sigma), out.ctypes.data, out.strides[0])
    return out


def gaussian_kernel(radius: int, sigma: float, lib=None) -> np.ndarray:
    lib = lib or load()
    n = (2 * radius + 1) ** 2
    taps = (C.c_int32 * n)()
    lib.tbref_gaussian_kernel(radius, float(sigma), taps)
    return np.array(taps, dtype=np.int32).reshape(2 * radius + 1, 2 * radius + 1)


def compose_regions(regions: Sequence[dict], width: int, height: int, lib=None) -> np.ndarray:
    """gst_ttmlrender_show_regions without the text rasterisation. Each dict: x, y, w, h,
    background_color (0xRRGGBBAA), opacity, layer (h x w x 4 uint8 premultiplied BGRA or None).
    Returns the frame-sized premultiplied BGRA overlay."""
    lib = lib or load()
    lib.tbref_compose_regions.restype = None
    lib.tbref_compose_regions.argtypes = [C.POINTER(RefRegion), C.c_uint32, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_int32]
    arr = (RefRegion * max(1, len(regions)))()
    for i, r in enumerate(regions):
        layer = r.get("layer")
        arr[i] = RefRegion(r["x"], r["y"], r["w"], r["h"], r.get("background_color", 0),
                           float(r.get("opacity", 1.0)),
                           layer.ctypes.data if layer is not None else None,
                           layer.strides[0] if layer is not None else 0)
    out = np.zeros((height, width, 4), dtype=np.uint8)
    lib.tbref_compose_regions(arr, len(regions), width, height, out.ctypes.data, out.strides[0])
    return out
