"""Probe: 256 streams of PAL SD I420 frames, pool strides against GStreamer's default strides
(720 / 360 / 360: the chroma rows are only 8-byte aligned and take the FAST=false kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft
pkg = graft.load_package(); wl = pkg.workloads
W, H, fmt, n = 720, 576, "I420", 256
ctx = pkg.TtmlBlend(0)
ov = np.zeros((H, W, 4), np.uint8); ov[470:540, 60:660] = (8, 8, 8, 200)
for s in range(n):
    ctx.overlay_set(s, ov, [(60, 470, 600, 70)])
ctx.set_batch(n, 0)
srcs = [ctx.acquire(fmt, W, H) for _ in range(n)]; dsts = [ctx.acquire(fmt, W, H) for _ in range(n)]
# pool frames have 256-byte strides; GStreamer's default for 720 px I420 is 720 / 360 / 360
tb = pkg.ttmlblend
def gst_view(fr):
    f = tb.Frame()
    for pl, st in enumerate((720, 360, 360)):
        f.plane[pl] = fr.c.plane[pl]; f.stride[pl] = st
    return f
for label, S, D in (("pool strides (768/512/512)", [s.c for s in srcs], [d.c for d in dsts]),
                    ("GStreamer strides (720/360/360)", [gst_view(s) for s in srcs], [gst_view(d) for d in dsts])):
    batch = ctx.Batch(list(range(n)), fmt, W, H, S, D)
    for _ in range(5): ctx.submit_many(batch)
    ctx.sync(); ctx.stats_reset(); k = 100
    ctx.timer_begin()
    for _ in range(k): ctx.submit_many(batch)
    ms = ctx.timer_end(); st = ctx.stats()
    print(f"{label}: {n*k/(ms*1e-3):,.0f} frames/s, {st['launches']/k:.1f} launches per step")
