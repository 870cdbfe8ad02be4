"""GPU parity of the region form: overlays composed on the GPU
from region descriptors (background colour, tts:opacity, optional host-rasterised text
layer) against the oracle's restatement of gst_ttmlrender_show_regions
(/root/reference/plugins/ttml/gstttmlrender.c:1250-1268,1375-1381) with Cairo's colour
conversion and pixman's 8-bit OVER / IN. Neither Cairo nor pixman is installed: parity
unpinned for that arithmetic (docs/BLENDSPEC.md section 9)."""
import numpy as np
import pytest

from helpers import ALL_FORMATS, assert_planes_equal, copy_planes, gpu_blend, oracle_blend, pkg, random_frame, random_overlay
from oracle import oracle

pytestmark = pytest.mark.gpu


def layer_for(w, h, seed):
    """A text-like premultiplied layer: a few glyph boxes on a cleared surface."""
    px = random_overlay(w, h, seed, premultiplied=True, density=0.9)
    mask = np.zeros((h, w), dtype=bool)
    for k in range(0, w - 6, 11):
        mask[h // 4: 3 * h // 4, k:k + 6] = True
    px[~mask] = 0
    return px


REGION_SETS = [
    # background only, opaque and translucent, one with opacity
    [dict(x=10, y=20, w=200, h=40, background_color=0x000000FF, opacity=1.0),
     dict(x=30, y=100, w=150, h=30, background_color=0x2040C080, opacity=1.0),
     dict(x=100, y=70, w=120, h=50, background_color=0xFF8000FF, opacity=0.4)],
    # overlapping regions in z order, hanging over the frame
    [dict(x=-20, y=-10, w=150, h=80, background_color=0x00FF00C0, opacity=0.75),
     dict(x=60, y=30, w=300, h=100, background_color=0x0000FFFF, opacity=1.0),
     dict(x=200, y=90, w=200, h=120, background_color=0xFFFFFF40, opacity=0.9)],
    # with text layers, with and without background / opacity
    [dict(x=16, y=120, w=288, h=40, background_color=0x000000C0, opacity=1.0, layer=7),
     dict(x=40, y=10, w=100, h=32, background_color=0, opacity=1.0, layer=8),
     dict(x=150, y=40, w=130, h=48, background_color=0x80000080, opacity=0.6, layer=9)],
    # nothing to draw
    [dict(x=5, y=5, w=50, h=50, background_color=0, opacity=1.0),
     dict(x=400, y=400, w=50, h=50, background_color=0xFFFFFFFF, opacity=1.0)],
]


def materialise(regions):
    out = []
    for r in regions:
        r = dict(r)
        if isinstance(r.get("layer"), int):
            r["layer"] = layer_for(r["w"], r["h"], r["layer"])
        out.append(r)
    return out


@pytest.mark.parametrize("fmt", ALL_FORMATS)
@pytest.mark.parametrize("k", range(len(REGION_SETS)))
def test_regions_match_oracle(ctx, fmt, k):
    w, h = 320, 180
    regions = materialise(REGION_SETS[k])
    ov = oracle.compose_regions(regions, w, h)
    assert (ov[:, :, :3].max(axis=2) <= ov[:, :, 3]).all()          # still valid premultiplied
    planes = random_frame(fmt, w, h, 60 + k)
    want = oracle_blend(fmt, w, h, copy_planes(planes), oracle.ttmlrender_rectangles(ov))
    ctx.overlay_set_regions(70 + k, w, h, regions)
    for mode in ("out", "inplace", "host"):
        got = gpu_blend(ctx, fmt, w, h, planes, mode=mode, stream=70 + k, set_overlay=False)
        assert_planes_equal(got, want, f"{fmt} regions {k} {mode}")


def test_region_known_answers(ctx):
    """Hand-computed: translucent red box, opaque green box at opacity 0.5 over it."""
    regions = [dict(x=2, y=1, w=4, h=2, background_color=0xFF000080, opacity=1.0),
               dict(x=4, y=2, w=4, h=2, background_color=0x00FF00FF, opacity=0.5)]
    ov = oracle.compose_regions(regions, 10, 5)
    assert list(ov[1, 2]) == [0, 0, 128, 128]             # B, G, R, A
    assert list(ov[3, 7]) == [0, 128, 0, 128]             # 0.5 * 65535 + 0.5 -> 0x8000 >> 8 = 128
    assert list(ov[2, 5]) == [0, 128, 64, 192]            # green IN 128, OVER the red
    frame = [np.zeros((5, 40), dtype=np.uint8)]
    frame[0][:, 3::4] = 255
    ctx.overlay_set_regions(99, 10, 5, regions)
    got = gpu_blend(ctx, "BGRA", 10, 5, frame, mode="out", stream=99, set_overlay=False)
    px = got[0].reshape(5, 10, 4)
    assert list(px[1, 2]) == [0, 0, 128, 255] and list(px[3, 7]) == [0, 128, 0, 255]
    assert list(px[2, 5]) == [0, 128, 64, 255] and list(px[0, 0]) == [0, 0, 0, 255]


def test_region_argument_checks(ctx):
    tb = pkg.ttmlblend
    bad = (tb.Region * 1)(tb.Region(0, 0, 10, 10, 0xFF, 1.5, None, 0))
    assert ctx.lib.fluc_ttmlblend_overlay_set_regions(ctx.h, 1, 64, 64, bad, 1) == tb.ERROR_INVALID_ARGUMENT
    assert ctx.lib.fluc_ttmlblend_overlay_set_regions(ctx.h, 1, 0, 64, bad, 0) == tb.ERROR_INVALID_ARGUMENT
    assert ctx.lib.fluc_ttmlblend_overlay_set_regions(ctx.h, 1, 64, 64, None, 0) == 0     # empty cue
