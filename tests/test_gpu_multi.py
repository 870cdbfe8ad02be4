"""Several GPUs in ONE process through FlucTtmlBlendMulti (stream s on context s % n), the way
a GStreamer process with many ttmlblend elements would use a box. With one GPU visible the
same code runs with n = 1; `gpurun --gpus 2` exercises the real thing."""
import numpy as np
import pytest

from helpers import assert_planes_equal, copy_planes, oracle_blend, pkg, random_frame, random_overlay

pytestmark = pytest.mark.gpu


def test_streams_are_dealt_round_robin_and_blend_on_their_gpu():
    m = pkg.TtmlBlendMulti()
    try:
        n = m.size()
        assert n == pkg.load_library().fluc_ttmlblend_device_count() and n >= 1
        fmt, w, h, n_streams = "NV12", 320, 180, 12
        work = []
        for s in range(n_streams):
            assert m.device(s) == s % n
            ctx = m.context(s)
            assert ctx is m.context(s + n)                       # same context for s and s + n
            rects = [dict(pixels=random_overlay(120 + 8 * s, 30, 40 + s), x=10 + 3 * s, y=100 + s)]
            ctx.overlay_set_rectangles(s, rects)
            planes = random_frame(fmt, w, h, 50 + s)
            src, dst = ctx.acquire(fmt, w, h), ctx.acquire(fmt, w, h)
            src.upload(planes)
            ticket = ctx.submit(s, fmt, w, h, src.c, dst.c)      # queued; launched by the sync below
            work.append((s, ctx, rects, planes, src, dst, ticket))
        m.sync()
        for s, ctx, rects, planes, src, dst, ticket in work:
            ctx.wait(ticket)
            want = oracle_blend(fmt, w, h, copy_planes(planes), rects)
            assert_planes_equal(dst.download(), want, f"stream {s} on device {m.device(s)}")
            # the host-frame path on the same context
            host = copy_planes(planes)
            ctx.wait(ctx.blend_host(s, fmt, w, h, host))
            assert_planes_equal(host, want, f"host frame, stream {s}")
            src.release()
            dst.release()
        st = m.stats()
        assert st["frames_blended"] == 2 * n_streams and st["overlays_set"] == n_streams
        assert st["launches"] >= n
    finally:
        m.close()


def test_explicit_device_list_and_errors():
    lib = pkg.load_library()
    n = lib.fluc_ttmlblend_device_count()
    m = pkg.TtmlBlendMulti([n - 1, 0])                           # any order, repeats allowed
    assert m.size() == 2 and m.device(0) == n - 1 and m.device(1) == 0 and m.device(2) == n - 1
    m.close()
    with pytest.raises(pkg.TtmlBlendError) as e:
        pkg.TtmlBlendMulti([n + 5])
    assert e.value.code == pkg.ttmlblend.ERROR_NO_DEVICE
    assert lib.fluc_ttmlblend_multi_size(None) == 0
    assert lib.fluc_ttmlblend_multi_context(None, 3) is None
