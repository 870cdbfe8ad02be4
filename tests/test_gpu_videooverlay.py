"""GPU tests through the host-side C mirror of the GStreamer API
(fluc_video_overlay_composition_blend == gst_video_overlay_composition_blend): these read
like gst-plugins-base's own overlay-composition checks -- build rectangles, build a
composition, blend onto a mapped frame, compare bytes -- with the CPU oracle as the
expected value."""
import numpy as np
import pytest

from helpers import ALL_FORMATS, assert_planes_equal, copy_planes, oracle_blend, pkg, random_frame, random_overlay

pytestmark = pytest.mark.gpu
vo = pkg.videooverlay


@pytest.mark.parametrize("fmt", ALL_FORMATS)
def test_composition_blend(ctx, fmt):
    w, h = 320, 180
    px0, px1 = random_overlay(200, 40, 1), random_overlay(64, 64, 2, premultiplied=False)
    r0 = vo.Rectangle(px0, 60, 120, vo.FLAG_PREMULTIPLIED_ALPHA)
    r1 = vo.Rectangle(px1, -10, 100, vo.FLAG_NONE)
    r1.set_global_alpha(0.5)
    comp = vo.Composition(r0)
    comp.add_rectangle(r1)
    assert comp.n_rectangles() == 2
    want_rects = [dict(pixels=px0, x=60, y=120, premultiplied=True),
                  dict(pixels=px1, x=-10, y=100, premultiplied=False, global_alpha=0.5)]
    for k in range(3):                        # the cached overlay serves every later frame
        planes = random_frame(fmt, w, h, 10 + k)
        want = oracle_blend(fmt, w, h, copy_planes(planes), want_rects)
        got = copy_planes(planes)
        assert comp.blend(fmt, w, h, got) is True
        assert_planes_equal(got, want, f"{fmt} frame {k}")


def test_empty_composition_and_bad_format(ctx):
    planes = random_frame("NV12", 64, 32, 3)
    before = copy_planes(planes)
    comp = vo.Composition()
    assert comp.blend("NV12", 64, 32, planes) is True
    assert_planes_equal(planes, before, "empty composition")
    f = vo.VideoFrame()
    f.format, f.width, f.height = 99, 64, 32
    import ctypes as C
    assert vo.load_library().fluc_video_overlay_composition_blend(comp.h, C.byref(f)) == 0
