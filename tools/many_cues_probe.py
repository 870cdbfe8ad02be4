"""Probe: config 5 (256 x 1080p I420 streams, one frame each per step) when every stream shows a
DIFFERENT cue without a background box -- after the auto-crop every stream has its own band
list, so frames cannot share group launches. Compares with config 5 as benched (boxed cue,
identical bands for every stream)."""
import dataclasses
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
wl = pkg.workloads
cfg5 = wl.CONFIGS[5]
n = int(os.environ.get("STREAMS", "256"))
for label, cfg, distinct in (("boxed cue, same layout", cfg5, 4),
                             ("glyph-only cues, 64 distinct layouts",
                              dataclasses.replace(cfg5, regions=[dataclasses.replace(cfg5.regions[0], bg=(0, 0, 0, 0))]), 64)):
    ctx = pkg.TtmlBlend(0)
    rng = np.random.default_rng(1)
    for s in range(n):
        ov = wl.overlay_for(cfg, stream=s % distinct)
        if distinct > 4:
            # different line lengths per stream, as real text has
            r = cfg.regions[0]
            cut = int(rng.integers(r.w // 4, r.w))
            ov[r.y:r.y + r.h // 2, r.x + cut:] = 0
            cut = int(rng.integers(r.w // 4, r.w))
            ov[r.y + r.h // 2:r.y + r.h, r.x + cut:] = 0
        ctx.overlay_set(s, ov, wl.region_rects(cfg))
    ctx.set_batch(min(n, 1024), 0)
    base = wl.frame_for(cfg, 0)
    srcs = [ctx.acquire(cfg.fmt, cfg.width, cfg.height) for _ in range(n)]
    dsts = [ctx.acquire(cfg.fmt, cfg.width, cfg.height) for _ in range(n)]
    for f in srcs:
        f.upload(base)
    batch = ctx.Batch(list(range(n)), cfg.fmt, cfg.width, cfg.height, [s.c for s in srcs], [d.c for d in dsts])
    for _ in range(5):
        ctx.submit_many(batch)
    ctx.sync()
    ctx.stats_reset()
    ctx.set_profiling(1)
    k = 100
    ctx.timer_begin()
    t0 = time.perf_counter()
    for _ in range(k):
        ctx.submit_many(batch)
    ms = ctx.timer_end()
    wall = time.perf_counter() - t0
    st = ctx.stats()
    ctx.set_profiling(0)
    print(f"  launches alone (event pair around each batch): {st['kernel_ms'] / max(1, st['kernel_ms_launches']):.3f} ms per step")
    print(f"{label}: {n * k / (ms * 1e-3):,.0f} frames/s on the device ({ms / k:.3f} ms per {n} frames), "
          f"{n * k / wall:,.0f} frames/s wall; {st['launches'] / k:.1f} launches per step "
          f"({st['group_launches'] / k:.1f} group, {st['multi_launches'] / k:.1f} multi-layout)")
    ctx.close()
