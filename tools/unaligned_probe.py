"""Probe: frames whose strides are multiples of 4 but not of 16 (GStreamer's default for widths
like 1366 or 854) take the byte-granular table kernel. How fast is it against the same frames
with 16-byte aligned strides?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
tb = pkg.ttmlblend
wl = pkg.workloads
W, H, fmt, n = 1366, 768, "NV12", 64
ctx = pkg.TtmlBlend(0)
rng = np.random.default_rng(0)
ov = np.zeros((H, W, 4), np.uint8)
ov[600:720, 100:1266] = (10, 10, 10, 160)
ctx.overlay_set(1, ov, [(100, 600, 1166, 120)])
ctx.set_batch(n, 0)
for label, stride in (("stride 1376 (16-byte aligned)", 1376), ("stride 1368 (4-byte aligned)", 1368)):
    big_s = [ctx.acquire(fmt, 2048, H) for _ in range(n)]
    big_d = [ctx.acquire(fmt, 2048, H) for _ in range(n)]
    srcs, dsts = [], []
    for s_, d_ in zip(big_s, big_d):
        sf, df = tb.Frame(), tb.Frame()
        for pl in range(2):
            sf.plane[pl], df.plane[pl] = s_.c.plane[pl], d_.c.plane[pl]
            sf.stride[pl] = df.stride[pl] = stride
        srcs.append(sf)
        dsts.append(df)
    batch = ctx.Batch([1] * n, fmt, W, H, srcs, dsts)
    for _ in range(5):
        ctx.submit_many(batch)
    ctx.sync()
    ctx.stats_reset()
    k = 50
    ctx.timer_begin()
    for _ in range(k):
        ctx.submit_many(batch)
    ms = ctx.timer_end()
    st = ctx.stats()
    fb = W * H * 3 // 2
    print(f"{label}: {n * k / (ms * 1e-3):,.0f} frames/s, {2 * fb * n * k / (ms * 1e-3) / 1e9:,.0f} GB/s of frame traffic, "
          f"{st['launches'] / k:.1f} launches per step ({st['group_launches'] / k:.1f} group)")
    for f in big_s + big_d:
        f.release()
ctx.close()
