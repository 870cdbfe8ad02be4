"""fluc_ttmlblend_overlay_update: the next state of a cue (a <set> animation step, a roll-up
line) patches the stream's overlay -- untouched region boxes keep their device pixels and
prepared planes. Whatever is kept or replaced, the frames must come out exactly as after a
full overlay_set of the new image (and as the oracle says)."""
import numpy as np
import pytest

from helpers import assert_planes_equal, copy_planes, oracle_blend, pkg, random_frame, wl
from oracle import oracle

pytestmark = pytest.mark.gpu

W, H = 640, 360
REGIONS = [wl.Region(32, 16, 576, 48, (0, 0, 0, 255), 0.8),
           wl.Region(16, 150, 240, 60, (0, 0, 128, 255), 0.5, ((255, 255, 0, 255),)),
           wl.Region(64, 280, 512, 64, (32, 32, 32, 255), 1.0)]
BOXES = [(r.x, r.y, r.w, r.h) for r in REGIONS]


def image(seeds):
    """ttmlrender's frame-sized image with region k drawn from seeds[k] (None: region absent)."""
    img = np.zeros((H, W, 4), dtype=np.uint8)
    for k, (reg, seed) in enumerate(zip(REGIONS, seeds)):
        if seed is None:
            continue
        one = wl.make_overlay(W, H, [reg], 4000 + 10 * seed + k)
        img[reg.y:reg.y + reg.h, reg.x:reg.x + reg.w] = one[reg.y:reg.y + reg.h, reg.x:reg.x + reg.w]
    return img


def blend(ctx, stream, fmt, frame):
    src, dst = ctx.acquire(fmt, W, H), ctx.acquire(fmt, W, H)
    src.upload(frame)
    ctx.wait(ctx.submit(stream, fmt, W, H, src.c, dst.c))
    out = dst.download()
    src.release()
    dst.release()
    return out


def want(fmt, frame, img):
    return oracle_blend(fmt, W, H, copy_planes(frame), oracle.ttmlrender_rectangles(img))


@pytest.mark.parametrize("fmt", ["NV12", "I420", "BGRA", "AYUV"])
def test_update_equals_full_set(ctx, fmt):
    frame = random_frame(fmt, W, H, 61)
    a = image([1, 1, 1])
    ctx.overlay_set(800, a, BOXES)
    assert_planes_equal(blend(ctx, 800, fmt, frame), want(fmt, frame, a), "before the update")
    ctx.sync()
    st0 = ctx.stats()
    b = image([1, 2, 1])                               # only the middle region differs
    ctx.overlay_update(800, b, [BOXES[1]])
    st1 = ctx.stats()
    assert st1["overlays_updated"] - st0["overlays_updated"] == 1
    assert st1["h2d_bytes"] - st0["h2d_bytes"] == BOXES[1][2] * BOXES[1][3] * 4     # nothing else went up
    got = blend(ctx, 800, fmt, frame)
    assert_planes_equal(got, want(fmt, frame, b), "after the update")
    ctx.overlay_set(801, b, BOXES)                     # the same image installed whole
    assert_planes_equal(blend(ctx, 801, fmt, frame), got, "update == full overlay_set")


def test_partial_spanning_and_outside_changes(ctx):
    fmt = "NV12"
    frame = random_frame(fmt, W, H, 62)
    cur = image([3, 3, 3])
    ctx.overlay_set(810, cur, BOXES)
    blend(ctx, 810, fmt, frame)
    # a change inside one box, smaller than the box
    nxt = cur.copy()
    nxt[290:310, 100:300] = image([3, 3, 4])[290:310, 100:300]
    st0 = ctx.stats()
    ctx.overlay_update(810, nxt, [(100, 290, 200, 20)])
    assert ctx.stats()["overlays_updated"] - st0["overlays_updated"] == 1
    assert_planes_equal(blend(ctx, 810, fmt, frame), want(fmt, frame, nxt), "change inside a box")
    # two rectangles, two boxes
    cur, nxt = nxt, image([5, 3, 5])
    nxt[150:210, 16:256] = cur[150:210, 16:256]
    ctx.overlay_update(810, nxt, [BOXES[0], BOXES[2]])
    assert_planes_equal(blend(ctx, 810, fmt, frame), want(fmt, frame, nxt), "two boxes changed")
    # no change at all
    st0 = ctx.stats()
    ctx.overlay_update(810, nxt, [])
    st1 = ctx.stats()
    assert st1["overlays_set"] == st0["overlays_set"] and st1["h2d_bytes"] == st0["h2d_bytes"]
    assert_planes_equal(blend(ctx, 810, fmt, frame), want(fmt, frame, nxt), "nothing changed")
    # something appears outside every region box: nothing to patch, the image is installed whole
    cur = nxt
    nxt = cur.copy()
    nxt[100:120, 300:500] = (40, 80, 120, 200)                     # premultiplied BGRA, colour <= alpha
    st0 = ctx.stats()
    ctx.overlay_update(810, nxt, [(300, 100, 200, 20)])
    st1 = ctx.stats()
    assert st1["overlays_updated"] == st0["overlays_updated"] and st1["overlays_set"] == st0["overlays_set"] + 1
    assert_planes_equal(blend(ctx, 810, fmt, frame), want(fmt, frame, nxt), "change outside the boxes")
    # ... and that whole-image overlay can be patched again (one box: the image)
    nxt2 = nxt.copy()
    nxt2[100:120, 300:500] = (10, 20, 30, 90)
    ctx.overlay_update(810, nxt2, [(300, 100, 200, 20)])
    assert_planes_equal(blend(ctx, 810, fmt, frame), want(fmt, frame, nxt2), "whole-image overlay patched")


def test_update_on_a_stream_without_an_overlay_installs_the_image(ctx):
    fmt = "I420"
    frame = random_frame(fmt, W, H, 63)
    ctx.overlay_clear(820)
    img = image([6, None, 6])
    ctx.overlay_update(820, img, [BOXES[0]])
    assert_planes_equal(blend(ctx, 820, fmt, frame), want(fmt, frame, img), "update without a previous overlay")


def test_queued_frames_keep_the_cue_they_were_submitted_under(ctx):
    fmt = "NV12"
    a, b = image([7, 7, 7]), image([7, 7, 8])
    ctx.overlay_set(830, a, BOXES)
    f1, f2 = random_frame(fmt, W, H, 64), random_frame(fmt, W, H, 65)
    ctx.set_batch(64, 0)
    try:
        bufs = [ctx.acquire(fmt, W, H) for _ in range(4)]
        bufs[0].upload(f1)
        bufs[2].upload(f2)
        t1 = ctx.submit(830, fmt, W, H, bufs[0].c, bufs[1].c)       # queued, not launched
        ctx.overlay_update(830, b, [BOXES[2]])
        t2 = ctx.submit(830, fmt, W, H, bufs[2].c, bufs[3].c)
        ctx.wait(t1)
        ctx.wait(t2)
        assert_planes_equal(bufs[1].download(), want(fmt, f1, a), "frame queued before the update")
        assert_planes_equal(bufs[3].download(), want(fmt, f2, b), "frame queued after the update")
        for x in bufs:
            x.release()
    finally:
        ctx.set_batch(32, 200)


def test_many_updates_neither_drift_nor_leak():
    """A <set> animation: one region changes again and again. Every state blends like a fresh
    overlay_set, host frames included, and the cache does not grow with the number of updates."""
    fmt = "NV12"
    c = pkg.TtmlBlend(0)
    try:
        frame = random_frame(fmt, W, H, 66)
        seeds = [9, 9, 9]
        c.overlay_set(1, image(seeds), BOXES)
        blend(c, 1, fmt, frame)
        sizes = []
        for step in range(24):
            k = step % 3
            seeds[k] = 10 + step
            img = image(seeds)
            c.overlay_update(1, img, [BOXES[k]])
            if step % 4 == 0:
                assert_planes_equal(blend(c, 1, fmt, frame), want(fmt, frame, img), f"step {step}")
                host = copy_planes(frame)
                c.wait(c.blend_host(1, fmt, W, H, host))
                assert_planes_equal(host, want(fmt, frame, img), f"step {step}, host frame")
            c.sync()
            sizes.append(c.stats()["cache_bytes"])
        assert c.stats()["overlays_updated"] == 24
        assert max(sizes[12:]) <= max(sizes[:12]), sizes       # steady: blocks of replaced boxes are freed
    finally:
        c.close()
