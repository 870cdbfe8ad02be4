"""ctypes binding of host/libfluc_videooverlay.so: the C mirror of
gst_video_overlay_rectangle_new_raw / gst_video_overlay_composition_new /
gst_video_overlay_composition_blend (host/fluc_videooverlay.h)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

from . import ttmlblend as tb

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "host", "libfluc_videooverlay.so")

FLAG_NONE = 0
FLAG_PREMULTIPLIED_ALPHA = 1
FLAG_GLOBAL_ALPHA = 2


class VideoFrame(C.Structure):
    _fields_ = [("format", C.c_int), ("width", C.c_int32), ("height", C.c_int32),
                ("flags", C.c_uint32), ("data", C.c_void_p * 3), ("stride", C.c_int32 * 3)]


PROTOTYPES = {
    "fluc_video_overlay_rectangle_new_raw": (C.c_void_p, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                          C.c_int32, C.c_int32, C.c_uint32, C.c_uint32,
                                                          C.c_uint32]),
    "fluc_video_overlay_rectangle_ref": (C.c_void_p, [C.c_void_p]),
    "fluc_video_overlay_rectangle_unref": (None, [C.c_void_p]),
    "fluc_video_overlay_rectangle_set_global_alpha": (None, [C.c_void_p, C.c_float]),
    "fluc_video_overlay_rectangle_get_global_alpha": (C.c_float, [C.c_void_p]),
    "fluc_video_overlay_rectangle_set_render_rectangle": (None, [C.c_void_p, C.c_int32, C.c_int32,
                                                                 C.c_uint32, C.c_uint32]),
    "fluc_video_overlay_composition_new": (C.c_void_p, [C.c_void_p]),
    "fluc_video_overlay_composition_add_rectangle": (None, [C.c_void_p, C.c_void_p]),
    "fluc_video_overlay_composition_n_rectangles": (C.c_uint32, [C.c_void_p]),
    "fluc_video_overlay_composition_ref": (C.c_void_p, [C.c_void_p]),
    "fluc_video_overlay_composition_unref": (None, [C.c_void_p]),
    "fluc_video_overlay_composition_blend": (C.c_int, [C.c_void_p, C.POINTER(VideoFrame)]),
    "fluc_video_overlay_set_device": (C.c_int, [C.c_int]),
    "fluc_video_overlay_get_context": (C.c_void_p, []),
    "fluc_video_overlay_deinit": (None, []),
}

_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        tb.load_library()                      # the C ABI it links against
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class Rectangle:
    def __init__(self, pixels: np.ndarray, x: int, y: int, flags: int = FLAG_PREMULTIPLIED_ALPHA,
                 render_width: int = 0, render_height: int = 0):
        assert pixels.dtype == np.uint8 and pixels.ndim == 3 and pixels.shape[2] == 4
        self.lib = load_library()
        self.h = self.lib.fluc_video_overlay_rectangle_new_raw(
            pixels.ctypes.data, pixels.shape[1], pixels.shape[0], pixels.strides[0], x, y,
            render_width, render_height, flags)
        if not self.h:
            raise ValueError("fluc_video_overlay_rectangle_new_raw returned NULL")

    def set_global_alpha(self, a: float):
        self.lib.fluc_video_overlay_rectangle_set_global_alpha(self.h, a)

    def get_global_alpha(self) -> float:
        return self.lib.fluc_video_overlay_rectangle_get_global_alpha(self.h)

    def set_render_rectangle(self, x: int, y: int, render_width: int = 0, render_height: int = 0):
        self.lib.fluc_video_overlay_rectangle_set_render_rectangle(self.h, x, y, render_width, render_height)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.fluc_video_overlay_rectangle_unref(self.h)
            self.h = None


class Composition:
    def __init__(self, rect: Rectangle = None):
        self.lib = load_library()
        self.h = self.lib.fluc_video_overlay_composition_new(rect.h if rect else None)

    def add_rectangle(self, rect: Rectangle):
        self.lib.fluc_video_overlay_composition_add_rectangle(self.h, rect.h)

    def n_rectangles(self) -> int:
        return self.lib.fluc_video_overlay_composition_n_rectangles(self.h)

    def blend(self, fmt: str, width: int, height: int, planes: Sequence[np.ndarray],
              premultiplied_dest: bool = False) -> bool:
        """gst_video_overlay_composition_blend (comp, frame): in place on host planes."""
        f = VideoFrame()
        f.format = tb.FORMATS[fmt]
        f.width, f.height = width, height
        f.flags = tb.FLAG_PREMULTIPLIED_ALPHA if premultiplied_dest else 0
        for i, p in enumerate(planes):
            assert p.dtype == np.uint8 and p.ndim == 2 and p.strides[1] == 1
            f.data[i] = p.ctypes.data
            f.stride[i] = p.strides[0]
        return bool(self.lib.fluc_video_overlay_composition_blend(self.h, C.byref(f)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.fluc_video_overlay_composition_unref(self.h)
            self.h = None
