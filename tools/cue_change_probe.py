"""Scratch probe: cost of a cue change (overlay_set + first frame = upload, row-span scan,
prepare) for the BASELINE configs, with and without region boxes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_package(); wl = pkg.workloads
ctx = pkg.TtmlBlend(0)
for cid in (1, 2, 3):
    cfg = wl.CONFIGS[cid]
    ov = wl.overlay_for(cfg)
    src, dst = ctx.acquire(cfg.fmt, cfg.width, cfg.height), ctx.acquire(cfg.fmt, cfg.width, cfg.height)
    src.upload(wl.frame_for(cfg, 0))
    for name, regions in (("region boxes", wl.region_rects(cfg)), ("whole image", ())):
        ts, tf = [], []
        for i in range(6):
            ctx.sync(); t0 = time.perf_counter()
            ctx.overlay_set(5, ov, regions)
            t1 = time.perf_counter()
            ctx.wait(ctx.submit(5, cfg.fmt, cfg.width, cfg.height, src.c, dst.c))
            t2 = time.perf_counter()
            ts.append(t1 - t0); tf.append(t2 - t1)
        print(f"cfg {cid} {cfg.width}x{cfg.height} {name:12s}: overlay_set {min(ts)*1e3:.2f} ms, first frame (prepare + blend) {min(tf)*1e3:.3f} ms")

# The next state of a cue: one of three regions changes (a <set> animation step, a roll-up line).
# overlay_set of the whole new image against overlay_update of the changed region's box.
import dataclasses
import numpy as np
for name, W, H, fmt, regs in (
        ("4K NV12", 3840, 2160, "NV12", [wl.Region(0, 72, 3840, 144, (0, 0, 0, 255), 0.75),
                                         wl.Region(0, 900, 3840, 144, (0, 0, 64, 255), 0.75),
                                         wl.Region(0, 1728, 3840, 144, (0, 0, 0, 255), 0.75)]),
        ("4K NV12, cfg 3 layout + a third region", 3840, 2160, "NV12",
         wl.CONFIGS[3].regions + [wl.Region(960, 1000, 1920, 144, (0, 0, 64, 255), 0.75)]),
        ("1080p NV12, cfg 2 regions", 1920, 1080, "NV12", wl.CONFIGS[2].regions)):
    boxes = [(r.x, r.y, r.w, r.h) for r in regs]
    imgs = [wl.make_overlay(W, H, regs, 700 + k) for k in range(2)]
    # state B = state A with only the last region re-rendered
    r = regs[-1]
    b = imgs[0].copy()
    b[r.y:r.y + r.h, r.x:r.x + r.w] = imgs[1][r.y:r.y + r.h, r.x:r.x + r.w]
    states = [imgs[0], b]
    src, dst = ctx.acquire(fmt, W, H), ctx.acquire(fmt, W, H)
    src.upload(wl.make_frame(fmt, W, H, 5))
    ctx.overlay_set(6, states[0], boxes)
    ctx.wait(ctx.submit(6, fmt, W, H, src.c, dst.c))
    res = {}
    for how in ("overlay_set, region boxes", "overlay_set, whole image", "overlay_update, changed box"):
        ts, tf = [], []
        for i in range(8):
            img = states[(i + 1) & 1]
            ctx.sync(); t0 = time.perf_counter()
            if how.startswith("overlay_update"):
                ctx.overlay_update(6, img, [boxes[-1]])
            else:
                ctx.overlay_set(6, img, boxes if "boxes" in how else ())
            t1 = time.perf_counter()
            ctx.wait(ctx.submit(6, fmt, W, H, src.c, dst.c))
            t2 = time.perf_counter()
            ts.append(t1 - t0); tf.append(t2 - t1)
        res[how] = min(ts)
        print(f"{name}: {how:30s} {min(ts)*1e3:.3f} ms, first frame {min(tf)*1e3:.3f} ms")
        if how == "overlay_set, whole image":
            ctx.overlay_set(6, states[0], boxes); ctx.wait(ctx.submit(6, fmt, W, H, src.c, dst.c))   # boxes back for the update runs
    print(f"{name}: update is {res['overlay_set, region boxes'] / res['overlay_update, changed box']:.1f}x cheaper than overlay_set with boxes, "
          f"{res['overlay_set, whole image'] / res['overlay_update, changed box']:.1f}x cheaper than handing over the whole image")
