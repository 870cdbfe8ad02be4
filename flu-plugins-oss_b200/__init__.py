"""flu-plugins-oss_b200: the B200-native TTML overlay-blend path.

Scope: compositing ttmlrender's rasterised, premultiplied
BGRA cue image onto raw video frames (I420 / NV12 / AYUV / RGBA / BGRA and
their plane/byte-order siblings), bit-exact with gst-plugins-base's
gst_video_overlay_composition_blend. The product is csrc/ (hand-written
sm_100a CUDA + the C ABI of include/fluc_ttmlblend.h) and host/ (C mirror of the
GStreamer call it replaces); the Python modules are a ctypes binding and the
synthetic workloads shared by tests/ and bench.py.

The directory name is not a Python identifier; import it through
__graft_entry__.load_package() (registers it as `flu_plugins_oss_b200`).
"""
from . import sharding, ttmlblend, videooverlay, workloads  # noqa: F401
from .ttmlblend import TtmlBlend, TtmlBlendError, TtmlBlendMulti, load_library  # noqa: F401

__all__ = ["sharding", "ttmlblend", "videooverlay", "workloads", "TtmlBlend", "TtmlBlendError", "TtmlBlendMulti",
           "load_library"]
