"""The GStreamer glue (flu-plugins-oss_b200/gst/gstttmlblend.c + gstflucallocator.c), compiled
against the functional fake of the GStreamer API in tests/gst_stub/ and RUN: subtitle buffers
with PTS / duration, ttmlrender's all-zero clear buffers, gap events, flushes and video frames go
through the element's pads the way upstream elements would push them, and every frame that
comes out is compared with the oracle (blended with the cue whose interval [PTS, PTS+duration)
holds the frame's running time -- /root/reference/plugins/ttml/gstttmlbase.c:180-181 -- or
untouched where no cue is showing, /root/reference/plugins/ttml/gstttmlevent.c:221-224)."""
import ctypes as C
import os
import subprocess
import threading
import time

import numpy as np
import pytest

import __graft_entry__ as graft
from helpers import assert_planes_equal, copy_planes, oracle_blend, pkg, random_frame, random_overlay
from oracle import oracle

pytestmark = pytest.mark.gpu

STUB = os.path.join(graft.ROOT, "tests", "gst_stub")
SEC = 1_000_000_000
NONE = (1 << 64) - 1
EV_FLUSH_START, EV_FLUSH_STOP, EV_EOS, EV_GAP = 1, 2, 5, 6
W, H = 640, 360


@pytest.fixture(scope="module")
def glue():
    subprocess.check_call(["make", "-s", "-C", STUB])
    lib = C.CDLL(os.path.join(STUB, "libgstglue_test.so"))
    lib.th_new.restype = C.c_void_p
    lib.th_new.argtypes = [C.c_int, C.c_int]
    lib.th_start.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    lib.th_has_allocator.argtypes = [C.c_void_p]
    lib.th_segment.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]
    lib.th_push_subtitle.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
    lib.th_subtitle_event.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
    lib.th_push_video.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    lib.th_frame_size.restype = C.c_uint64
    lib.th_frame_size.argtypes = [C.c_void_p]
    lib.th_layout.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    lib.th_stats.argtypes = [C.c_void_p, C.POINTER(pkg.ttmlblend.Stats), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.th_errors.argtypes = [C.c_void_p]
    lib.th_free.argtypes = [C.c_void_p]
    return lib


class Pipeline:
    """One ttmlblend element with a video branch (format, W x H) and a subtitle branch."""

    def __init__(self, lib, fmt="NV12", auto_register=False, video_segment=(0, NONE, 0), subtitle_segment=(0, NONE, 0)):
        self.lib, self.fmt = lib, fmt
        self.h = lib.th_new(0, 1 if auto_register else 0)
        assert lib.th_start(self.h, fmt.encode(), W, H) == 0
        assert lib.th_segment(self.h, 0, *video_segment) == 0
        assert lib.th_segment(self.h, 1, *subtitle_segment) == 0
        n = len(pkg.workloads.plane_shapes(fmt, W, H))
        off, st = (C.c_uint64 * 4)(), (C.c_int32 * 4)()
        lib.th_layout(self.h, off, st)
        self.layout = [(off[p], st[p]) for p in range(n)]
        self.size = lib.th_frame_size(self.h)

    def close(self):
        if self.h:
            self.lib.th_free(self.h)
            self.h = None

    def cue(self, img, pts, duration=NONE):
        img = np.ascontiguousarray(img)
        return self.lib.th_push_subtitle(self.h, img.ctypes.data, img.shape[1], img.shape[0], pts, duration)

    def event(self, kind, ts=0, duration=0):
        assert self.lib.th_subtitle_event(self.h, kind, ts, duration) == 0

    def frame(self, planes, pts, use_allocator=False):
        """Pushes one frame (planes in, planes out) through the video pad."""
        buf = np.zeros(self.size, dtype=np.uint8)
        views = []
        for (off, stride), p in zip(self.layout, planes):
            v = buf[off:off + stride * p.shape[0]].reshape(p.shape[0], stride)[:, :p.shape[1]]
            v[...] = p
            views.append(v)
        ret = self.lib.th_push_video(self.h, buf.ctypes.data, pts, 1 if use_allocator else 0)
        assert ret == 0, f"flow return {ret}"
        return [v.copy() for v in views]

    def stats(self):
        st, b, p = pkg.ttmlblend.Stats(), C.c_uint64(), C.c_uint64()
        self.lib.th_stats(self.h, C.byref(st), C.byref(b), C.byref(p))
        d = {k: getattr(st, k) for k, _ in pkg.ttmlblend.Stats._fields_}
        d["element_blended"], d["element_passed"] = b.value, p.value
        return d


def want(fmt, frame, img):
    if img is None:
        return copy_planes(frame)
    return oracle_blend(fmt, W, H, copy_planes(frame), oracle.ttmlrender_rectangles(img))


def images(n, seed):
    out = []
    for i in range(n):
        img = np.zeros((H, W, 4), dtype=np.uint8)
        img[240:330, 60:580] = random_overlay(520, 90, seed + i, density=0.8)      # a cue near the bottom
        out.append(img)
    return out


@pytest.mark.parametrize("fmt,use_allocator", [("NV12", False), ("NV12", True), ("I420", False), ("BGRA", True)])
def test_cue_intervals(glue, fmt, use_allocator):
    """Cues arrive ahead of the video (ttmlrender parses faster than real time); each one shows
    for [PTS, PTS+duration), a clear buffer shows nothing, a cue without duration lasts until the
    next one starts."""
    p = Pipeline(glue, fmt)
    try:
        assert glue.th_has_allocator(p.h)
        a, b, c = images(3, 100)
        clear = np.zeros((H, W, 4), dtype=np.uint8)
        assert p.cue(a, 1 * SEC, 2 * SEC) == 0
        assert p.cue(clear, 3 * SEC, 1 * SEC) == 0         # gstttmlevent.c:221-224
        assert p.cue(b, 4 * SEC) == 0                      # no duration
        assert p.cue(c, 6 * SEC, 1 * SEC) == 0
        showing = lambda t: a if 1 <= t < 3 else b if 4 <= t < 6 else c if 6 <= t < 7 else None
        before = p.stats()
        for k in range(17):                                # 0 .. 8 s in half seconds
            t = k * 0.5
            frame = random_frame(fmt, W, H, 200 + k)
            got = p.frame(frame, int(t * SEC), use_allocator)
            assert_planes_equal(got, want(fmt, frame, showing(t)), f"{fmt} frame at {t} s")
        after = p.stats()
        # the clear buffer is a cue like any other to the element (an image without a single
        # non-transparent pixel): 2 more frames go to the library, which finds nothing to do
        assert after["element_blended"] - before["element_blended"] == 4 + 2 + 4 + 2
        assert after["element_passed"] - before["element_passed"] == 17 - 12
        staged = after["staged_frames"] - before["staged_frames"]
        assert staged == (0 if use_allocator else 10)      # pool frames are reached over PCIe directly
        assert glue.th_errors(p.h) == 0
    finally:
        p.close()


def test_running_time_of_both_branches(glue):
    """The two pads have segments of their own: a video frame at PTS 10.5 s of a segment starting
    at 10 s and a cue at PTS 100.5 s of a segment starting at 100 s meet at running time 0.5 s."""
    fmt = "NV12"
    p = Pipeline(glue, fmt, video_segment=(10 * SEC, NONE, 0), subtitle_segment=(100 * SEC, NONE, 0))
    try:
        a, = images(1, 110)
        assert p.cue(a, 100 * SEC + SEC // 2, SEC) == 0
        frame = random_frame(fmt, W, H, 220)
        for pts, img in ((10.0, None), (10.5, a), (11.25, a), (11.5, None), (12.0, None)):
            assert_planes_equal(p.frame(frame, int(pts * SEC)), want(fmt, frame, img), f"video PTS {pts}")
        # a base: the second segment of the video continues the running time where the first stopped
        assert glue.th_segment(p.h, 0, 50 * SEC, NONE, 2 * SEC) == 0
        b, = images(1, 111)
        assert p.cue(b, 103 * SEC, SEC) == 0               # running time 3 s
        for pts, img in ((50.5, None), (51.0, b), (51.9, b), (52.0, None)):
            assert_planes_equal(p.frame(frame, int(pts * SEC)), want(fmt, frame, img), f"video PTS {pts} (base 2 s)")
    finally:
        p.close()


def test_gap_event_and_flush(glue):
    fmt = "NV12"
    p = Pipeline(glue, fmt)
    try:
        a, b = images(2, 120)
        frame = random_frame(fmt, W, H, 230)
        assert p.cue(a, 1 * SEC) == 0                      # open ended ...
        p.event(EV_GAP, 2 * SEC, 5 * SEC)                  # ... until the gap
        for t, img in ((0.5, None), (1.0, a), (1.9, a), (2.0, None), (3.0, None)):
            assert_planes_equal(p.frame(frame, int(t * SEC)), want(fmt, frame, img), f"gap, {t} s")
        # a seek: queued cues are dropped, and so is the one showing
        assert p.cue(a, 4 * SEC, 10 * SEC) == 0
        assert p.cue(b, 20 * SEC, 10 * SEC) == 0
        assert_planes_equal(p.frame(frame, int(4.5 * SEC)), want(fmt, frame, a), "before the flush")
        p.event(EV_FLUSH_START)
        assert p.cue(b, 5 * SEC, SEC) == -2                # GST_FLOW_FLUSHING
        p.event(EV_FLUSH_STOP)
        assert glue.th_segment(p.h, 1, 0, NONE, 0) == 0
        assert_planes_equal(p.frame(frame, int(5 * SEC)), want(fmt, frame, None), "after the flush")
        assert_planes_equal(p.frame(frame, int(21 * SEC)), want(fmt, frame, None), "dropped cue never shows")
        assert p.cue(b, 22 * SEC, SEC) == 0
        assert_planes_equal(p.frame(frame, int(22.5 * SEC)), want(fmt, frame, b), "new cue after the flush")
    finally:
        p.close()


def test_subtitle_thread_blocks_until_the_video_catches_up(glue):
    """More cues than slots: the subtitle streaming thread is held back (as textoverlay holds its
    text pad) until the video position has used up the earlier cues."""
    fmt = "NV12"
    p = Pipeline(glue, fmt)
    try:
        imgs = images(4, 130)
        n_cues = 14
        done = []

        def subtitle_thread():
            for k in range(n_cues):
                assert p.cue(imgs[k % 4], (k + 1) * SEC, SEC) == 0
                done.append(k)

        th = threading.Thread(target=subtitle_thread)
        th.start()
        time.sleep(0.5)
        assert th.is_alive() and len(done) < n_cues          # blocked: the video has not moved
        frame = random_frame(fmt, W, H, 240)
        for k in range(n_cues + 1):
            t = k + 0.5
            got = p.frame(frame, int(t * SEC))
            assert_planes_equal(got, want(fmt, frame, imgs[(k - 1) % 4] if k >= 1 else None), f"{t} s")
            time.sleep(0.02)
        th.join(timeout=20)
        assert not th.is_alive() and len(done) == n_cues
    finally:
        p.close()


def test_auto_register_pins_and_forgets(glue):
    """auto-register=true: upstream's (pageable) memory is pinned on first sight and blended zero
    copy; when the GstMemory is finalised the registration goes with it."""
    fmt = "NV12"
    p = Pipeline(glue, fmt, auto_register=True)
    try:
        a, = images(1, 140)
        assert p.cue(a, 0, 100 * SEC) == 0
        before = p.stats()
        for k in range(4):
            frame = random_frame(fmt, W, H, 250 + k)
            assert_planes_equal(p.frame(frame, int((k + 0.5) * SEC)), want(fmt, frame, a), f"frame {k}")
        after = p.stats()
        assert after["staged_frames"] == before["staged_frames"]        # none staged: all pinned
        assert glue.th_errors(p.h) == 0
    finally:
        p.close()


def test_two_elements_share_the_process_wide_context(glue):
    fmt = "NV12"
    p1, p2 = Pipeline(glue, fmt), Pipeline(glue, fmt)
    try:
        a, b = images(2, 150)
        assert p1.cue(a, 0, 10 * SEC) == 0
        assert p2.cue(b, 0, 10 * SEC) == 0
        frame = random_frame(fmt, W, H, 260)
        assert_planes_equal(p1.frame(frame, SEC), want(fmt, frame, a), "element 1")
        assert_planes_equal(p2.frame(frame, SEC), want(fmt, frame, b), "element 2")
        assert_planes_equal(p1.frame(frame, 2 * SEC, True), want(fmt, frame, a), "element 1, pool frame")
    finally:
        p1.close()
        p2.close()
