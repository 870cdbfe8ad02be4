"""Scratch: the DMA-batch pipeline inside the library -- the bare pipeline shape on the context's own
streams (pcie_probe modes 2 / 4) before and after real DMA batches have run in the same context."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); wl = pkg.workloads
c = pkg.TtmlBlend(0)
import os
if os.environ.get('PROBE_FIRST'): print('bare pipeline, fresh context: mode 4', round(c.pcie_probe(4, 1 << 20, 0.4), 1), 'GB/s each way')
cfg = wl.CONFIGS[3]
c.overlay_set(1, wl.overlay_for(cfg), wl.region_rects(cfg))
c.set_batch(32, 0)
NS = int(os.environ.get('HOST_SETS', '2'))
sets = [[c.acquire('NV12', 3840, 2160, on_host=True) for _ in range(32)] for _ in range(NS)]
bs = [c.Batch([1] * 32, 'NV12', 3840, 2160, [f.c for f in s], [f.c for f in s]) for s in sets]
def loop(n):
    prev = None
    for i in range(n):
        t = c.blend_host_many(bs[i % NS])
        if prev is not None: c.wait(prev)
        prev = t[31]
    c.wait(prev)
loop(4); c.sync(); c.stats_reset()
t0 = time.perf_counter(); loop(40); c.sync(); dt = time.perf_counter() - t0
st = c.stats()
print('real path:', round(32 * 40 / dt), 'frames/s,', round(st['h2d_bytes'] / dt / 1e9, 1), 'GB/s each way, dma batches', st['host_dma_batches'])
print('bare pipeline, same context afterwards: mode 4', round(c.pcie_probe(4, 1 << 20, 0.4), 1), 'GB/s each way')
