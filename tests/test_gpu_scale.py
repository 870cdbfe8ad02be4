"""GPU parity for rectangles whose render size differs from their pixel size: the composition
scales them first (gst_video_blend_scale_linear_RGBA; docs/BLENDSPEC.md section 10) -- here on
the GPU, once per cue -- and then blends the scaled image. Bit-exact against the oracle."""
import numpy as np
import pytest

from helpers import (ALL_FORMATS, assert_planes_equal, copy_planes, gpu_blend, oracle_blend, pkg, random_frame,
                     random_overlay)
from oracle import oracle

pytestmark = pytest.mark.gpu
vo = pkg.videooverlay
tb = pkg.ttmlblend

RECTS = [
    # src w, h, x, y, render w, h, global alpha, premultiplied
    (50, 20, 10, 50, 120, 33, 1.0, True),           # up
    (90, 60, -7, -5, 40, 25, 0.6, False),           # down > 2x, hanging over the top-left corner
    (30, 30, 140, 70, 61, 30, 1.0, True),           # over the bottom-right corner, height kept
    (64, 48, 20, 10, 64, 11, 1.0, True),            # width kept
    (2, 2, 0, 0, 33, 17, 1.0, True),                # the smallest source
    (40, 40, 100, 5, 1, 1, 1.0, True),              # down to one pixel
]


def rect_dict(i, spec):
    sw, sh, x, y, rw, rh, ga, pm = spec
    return dict(pixels=random_overlay(sw, sh, 500 + i, premultiplied=pm), x=x, y=y, render_width=rw,
                render_height=rh, global_alpha=ga, premultiplied=pm)


@pytest.mark.parametrize("mode", ("out", "inplace", "host"))
@pytest.mark.parametrize("fmt", ALL_FORMATS)
def test_scaled_rectangles_match_oracle(ctx, fmt, mode):
    w, h = 160, 90
    rects = [rect_dict(i, s) for i, s in enumerate(RECTS)]
    planes = random_frame(fmt, w, h, 78)
    want = oracle_blend(fmt, w, h, copy_planes(planes), rects)
    got = gpu_blend(ctx, fmt, w, h, planes, rects, mode=mode)
    assert_planes_equal(got, want, f"{fmt} {mode}")


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_scaled_geometry(ctx, seed):
    r = np.random.default_rng(9000 + seed)
    fmt = ALL_FORMATS[seed % len(ALL_FORMATS)]
    w, h = int(r.integers(16, 300)), int(r.integers(16, 200))
    rects = []
    for i in range(int(r.integers(1, 4))):
        sw, sh = int(r.integers(2, 120)), int(r.integers(2, 90))
        rw, rh = int(r.integers(1, 2 * w)), int(r.integers(1, 2 * h))
        pm = bool(r.integers(0, 2))
        rects.append(dict(pixels=random_overlay(sw, sh, 100 * seed + i, premultiplied=pm),
                          x=int(r.integers(-rw, w)), y=int(r.integers(-rh, h)), render_width=rw, render_height=rh,
                          global_alpha=float(r.choice([1.0, 1.0, 0.5])), premultiplied=pm))
    planes = random_frame(fmt, w, h, seed)
    want = oracle_blend(fmt, w, h, copy_planes(planes), rects)
    got = gpu_blend(ctx, fmt, w, h, planes, rects, mode="out")
    assert_planes_equal(got, want, f"{fmt} {w}x{h} seed {seed}")


def test_line_cache_quirk_rows_on_the_gpu(ctx):
    """A 4377-row source scaled to 3340 rows: rows 512, 1024, ... copy a stale cache line
    upstream; the runtime's row plan reproduces that (tests/test_host_logic.py)."""
    r = np.random.default_rng(11)
    px = r.integers(0, 256, (4377, 6, 4), dtype=np.uint8)
    px[..., :3] = (px[..., :3].astype(np.int32) * px[..., 3:] // 255).astype(np.uint8)
    rects = [dict(pixels=px, x=3, y=0, render_width=9, render_height=3340)]
    w, h = 32, 3340
    planes = random_frame("BGRA", w, h, 5)
    want = oracle_blend("BGRA", w, h, copy_planes(planes), rects)
    got = gpu_blend(ctx, "BGRA", w, h, planes, rects, mode="out")
    assert_planes_equal(got, want, "quirk")


def test_subtitle_rendered_at_half_size_then_scaled_to_the_frame(ctx):
    """A use the composition API allows: the cue image is produced at 960x540 and rendered
    at 1920x1080 over an NV12 frame."""
    ov = pkg.workloads.overlay_for(pkg.workloads.CONFIGS[2])[::2, ::2].copy()
    w, h = 1920, 1080
    rects = [dict(pixels=ov, x=0, y=0, render_width=w, render_height=h)]
    planes = random_frame("NV12", w, h, 6)
    want = oracle_blend("NV12", w, h, copy_planes(planes), rects)
    got = gpu_blend(ctx, "NV12", w, h, planes, rects, mode="host")
    assert_planes_equal(got, want, "half-size cue")


def test_host_mirror_render_size(ctx):
    """gst_video_overlay_rectangle_new_raw (.., render_width, render_height, ..) and
    _set_render_rectangle through the C mirror."""
    w, h = 200, 120
    px = random_overlay(40, 30, 3)
    r0 = vo.Rectangle(px, 5, 7, vo.FLAG_PREMULTIPLIED_ALPHA, render_width=100, render_height=45)
    r1 = vo.Rectangle(px, 0, 0, vo.FLAG_PREMULTIPLIED_ALPHA)
    r1.set_render_rectangle(120, 60, 70, 55)
    comp = vo.Composition(r0)
    comp.add_rectangle(r1)
    rects = [dict(pixels=px, x=5, y=7, render_width=100, render_height=45),
             dict(pixels=px, x=120, y=60, render_width=70, render_height=55)]
    planes = random_frame("I420", w, h, 4)
    want = oracle_blend("I420", w, h, copy_planes(planes), rects)
    got = copy_planes(planes)
    assert comp.blend("I420", w, h, got) is True
    assert_planes_equal(got, want, "mirror")


def test_scaling_a_one_pixel_wide_source_is_refused(ctx):
    """Upstream's increment is -1 there and it reads outside the image: an error here."""
    px = np.zeros((8, 1, 4), np.uint8)
    with pytest.raises(tb.TtmlBlendError) as e:
        ctx.overlay_set_rectangles(3, [dict(pixels=px, x=0, y=0, render_width=4, render_height=8)])
    assert e.value.code == tb.ERROR_INVALID_ARGUMENT
    ctx.overlay_set_rectangles(3, [dict(pixels=px, x=0, y=0, render_width=1, render_height=8)])   # not scaled
