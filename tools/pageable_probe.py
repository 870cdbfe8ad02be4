"""Scratch probe: blend_host on PAGEABLE host frames (what a pipeline without the pinned
allocator hands over) vs pinned pool frames vs host_register'ed memory. Config 3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft
pkg = graft.load_package(); wl = pkg.workloads
cfg = wl.CONFIGS[3]
ctx = pkg.TtmlBlend(0)
ctx.overlay_set(1, wl.overlay_for(cfg), wl.region_rects(cfg))
n = 16
base = wl.frame_for(cfg, 0)
def run(name, frames_c, k=20):
    def step():
        t = [ctx.blend_host_frame(1, cfg.fmt, cfg.width, cfg.height, f) for f in frames_c]
        for x in t: ctx.wait(x)
    for _ in range(2): step()
    ctx.sync(); t0 = time.perf_counter()
    for _ in range(k): step()
    ctx.sync(); dt = time.perf_counter() - t0
    print(f"{name:28s}: {n*k/dt:8.0f} frames/s")
tb = pkg.ttmlblend
pageable = [[p.copy() for p in base] for _ in range(n)]
run("pageable numpy", [tb._frame_from_arrays(f) for f in pageable])
for f in pageable:
    for p in f: ctx.host_register(p)
run("host_register'ed numpy", [tb._frame_from_arrays(f) for f in pageable])
pinned = [ctx.acquire(cfg.fmt, cfg.width, cfg.height, on_host=True) for _ in range(n)]
run("pinned pool frames", [p.c for p in pinned])
