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
