"""Probe of the in-place "look at the overlay first" variant (LAZY group kernel): host frames
blended over PCIe and device frames blended in place, for config 3's layout with and without a
background box behind the text. Run twice, FLUC_TTMLBLEND_LAZY=0 and =1 (read once per process):

    FLUC_TTMLBLEND_LAZY=0 python tools/lazy_probe.py; FLUC_TTMLBLEND_LAZY=1 python tools/lazy_probe.py
"""
import dataclasses
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
wl = pkg.workloads
from oracle import oracle  # noqa: E402

base_cfg = wl.CONFIGS[3]
n = 32
for label, cfg in (("background box", base_cfg),
                   ("glyphs only", dataclasses.replace(base_cfg, regions=[dataclasses.replace(r, bg=(0, 0, 0, 0))
                                                                       for r in base_cfg.regions])),
                   ("opaque box", dataclasses.replace(base_cfg, regions=[dataclasses.replace(r, opacity=1.0)
                                                                      for r in base_cfg.regions]))):
    ctx = pkg.TtmlBlend(0)
    ov = wl.overlay_for(cfg)
    zero_vec = float((ov[..., 3].reshape(ov.shape[0], -1, 16).max(axis=2) == 0)[
        np.r_[72:216, 1728:2088]].mean())
    ctx.overlay_set(1, ov, wl.region_rects(cfg))
    ctx.set_batch(n, 0)
    base = wl.frame_for(cfg, 0)
    hosts = [ctx.acquire(cfg.fmt, cfg.width, cfg.height, on_host=True) for _ in range(n)]
    devs = [ctx.acquire(cfg.fmt, cfg.width, cfg.height) for _ in range(n)]
    for hf, df in zip(hosts, devs):
        for d, s in zip(hf.host_planes(), base):
            d[...] = s
        df.upload(base)

    def step_host():
        t = [ctx.blend_host_frame(1, cfg.fmt, cfg.width, cfg.height, hf.c) for hf in hosts]
        ctx.wait(t[-1])

    def step_dev():
        t = [ctx.submit(1, cfg.fmt, cfg.width, cfg.height, df.c, df.c) for df in devs]
        ctx.wait(t[-1])

    res = {}
    for name, fn, k in (("host frames over PCIe", step_host, 30), ("device frames in place", step_dev, 200)):
        for _ in range(3):
            fn()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        ctx.sync()
        res[name] = n * k / (time.perf_counter() - t0)
    # exactness of what the probe just timed: first host frame was blended 33 times in place, so
    # check a fresh one
    for d, s in zip(hosts[0].host_planes(), base):
        d[...] = s
    ctx.wait(ctx.blend_host_frame(1, cfg.fmt, cfg.width, cfg.height, hosts[0].c))
    want = oracle.composition_blend(cfg.fmt, cfg.width, cfg.height, [p.copy() for p in base],
                                    oracle.ttmlrender_rectangles(ov))
    ok = all(np.array_equal(a, b) for a, b in zip(hosts[0].host_planes(), want))
    print(f"LAZY={os.environ.get('FLUC_TTMLBLEND_LAZY', '1')} {label:15s} ({100 * zero_vec:.0f} % of the luma vectors under "
          f"the cue are transparent): " + ", ".join(f"{k} {v:,.0f} frames/s" for k, v in res.items()) +
          f"; bit-exact {ok}")
    ctx.close()
