#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: frames/s and % of the HBM roofline for the 4K NV12
TTML overlay blend (config 3: 3840x2160 NV12, two full-width cue regions, 32 frames per
launch), with the reference's CPU blend timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one launch of the blend over one batch (32 device-resident 4K frames,
src -> dst, whole frame read and written; 398 MB in + 398 MB out per step, i.e. inputs
larger than the 126 MB L2). Per-GPU work is fixed as N grows (weak scaling, each rank its
own frame batches, no collective on the data path); torch.distributed is only used for the
barrier, the MAX over ranks of the device time and a gather of the result records.

Numbers in the JSON line:
  value        frames/s, frames resident in HBM, CUDA-event time on the blend stream.
  e2e          frames/s through fluc_ttmlblend_blend_host (the drop-in for
               gst_video_overlay_composition_blend): pinned HOST frames, the rows the
               overlay touches go host->device, blend, device->host inside the timed region.
  roofline     algorithmic bytes per launch (frame read + frame write +
               4 B/px overlay) / mean per-launch CUDA-event time, against MEASURED_PEAKS.json.
  cpu_baseline the CPU oracle (a port of gst_video_blend: GStreamer is not installable here)
               on this box's host cores, bounded sample, rank 0 at N=1 only.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import __graft_entry__ as graft  # noqa: E402

METRIC = "frames/sec, 4K NV12 TTML overlay blend"
UNIT = "frames/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:   # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0: float, t1: float):
        rows = [s for t, s in self.samples if t0 <= t <= t1]
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    device = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's version / debug lines go to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            device = torch.device("cuda", local)
            dist_mod.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
        else:
            dist_mod.init_process_group("gloo", rank=rank, world_size=world)
        dist = dist_mod
    return dist, device, world, rank, local


def barrier(dist, device):
    if dist is not None:
        if device is not None:
            import torch
            torch.cuda.synchronize(device)
        dist.barrier()


# ---------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)

def cpu_reference_setup(cfg, wl, n_frames):
    from oracle import oracle
    lib = oracle.load(native=True, out_dir=tempfile.mkdtemp(prefix="ttmlblend_oracle_"))
    ov = wl.overlay_for(cfg)
    base = wl.frame_for(cfg, 0)
    frames, keep = [], []
    import ctypes as C
    arr = (oracle.RefFrame * n_frames)()
    for i in range(n_frames):
        planes = [np.roll(p, i * 16, axis=1).copy() for p in base]
        keep.append(planes)
        arr[i] = oracle.make_frame(cfg.fmt, cfg.width, cfg.height, planes)
    rects = oracle.make_rectangles(oracle.ttmlrender_rectangles(ov))
    return lib, arr, rects, keep, ov, C


def cpu_blend_fps(lib, arr, rects, n_frames, threads):
    secs = lib.tbref_blend_many(arr, n_frames, rects, 1, threads)
    return n_frames / secs, secs


def run_cpu_baseline(cfg, wl, target_s=12.0):
    cores = os.cpu_count() or 1
    probe_n = cores
    lib, arr, rects, keep, ov, C = cpu_reference_setup(cfg, wl, probe_n)
    fps, secs = cpu_blend_fps(lib, arr, rects, probe_n, cores)       # probe (also warms caches)
    rounds = max(1, int(target_s / max(secs, 1e-3)))
    t = 0.0
    for _ in range(rounds):
        _, s = cpu_blend_fps(lib, arr, rects, probe_n, cores)
        t += s
    n = rounds * probe_n
    # for transparency: the same CPU code if someone cropped the image to the region boxes first
    # (the reference does not: ttmlrender pushes the whole frame-sized image downstream)
    from oracle import oracle
    crop = [dict(pixels=ov[r.y:r.y + r.h, r.x:r.x + r.w], x=r.x, y=r.y) for r in cfg.regions]
    crop_rects = oracle.make_rectangles(crop)
    t_crop = lib.tbref_blend_many(arr, probe_n, crop_rects, len(crop), cores)
    t_crop = min(t_crop, lib.tbref_blend_many(arr, probe_n, crop_rects, len(crop), cores))
    return {"value": n / t, "unit": UNIT, "cores": cores, "kind": "port",
            "value_if_cropped_to_regions": probe_n / t_crop,
            "sample": f"{n} frames of {cfg.width}x{cfg.height} {cfg.fmt} ({probe_n} buffers re-blended "
                      f"{rounds}x), ttmlrender's frame-sized premultiplied BGRA image as one "
                      f"rectangle (what the reference pipeline blends), oracle/ttmlblend_ref.c "
                      f"-O3 -march=native, {cores} pthreads, {t:.1f} s"}


def run_reference_arm(args, cfg, wl, dist, device, world, rank, real_stdout):
    """--impl reference: the reference's CPU implementation of the path (oracle port; GStreamer
    cannot be installed here) on all host threads. Rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = cores                       # one frame per thread per step: a bounded sample
    lib, arr, rects, keep, ov, C = cpu_reference_setup(cfg, wl, per_step)
    for _ in range(args.warmup):
        cpu_blend_fps(lib, arr, rects, per_step, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_blend_fps(lib, arr, rects, per_step, cores)[1]
    fps = per_step * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": cfg.name, "format": cfg.fmt, "width": cfg.width, "height": cfg.height,
                   "frames_per_step": per_step,
                   "note": "CPU port of gst_video_overlay_composition_blend (oracle/ttmlblend_ref.c); "
                           "GStreamer itself is not installed in this image"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} frames per step, frame-sized overlay rectangle"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=real_stdout, flush=True)


# ---------------------------------------------------------------------------

def protect_stdout():
    """stdout carries exactly one JSON line. Native libraries (NCCL prints its version there) write
    to file descriptor 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and
    the line goes to a private duplicate of the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--format", default=None, help="override the config's frame format (config 4: RGBA/BGRA/AYUV)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-inplace", dest="inplace", action="store_false",
                    help="skip the in-place variant (reported separately, SURVEY 8d: B_inplace)")
    ap.add_argument("--inplace", dest="inplace", action="store_true", default=True)
    ap.add_argument("--profile-every", type=int, default=16,
                    help="CUDA-event pair around every n-th launch of the timed region (roofline)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    pkg = graft.load_package()
    wl = pkg.workloads
    sh = pkg.sharding
    cfg = wl.CONFIGS[args.config]
    dist, device, world, rank, local = dist_setup(args.gpus)

    if args.impl == "reference":
        run_reference_arm(args, cfg, wl, dist, device, world, rank, real_stdout)
        if dist is not None:
            dist.destroy_process_group()
        return

    ctx = pkg.TtmlBlend(local)             # raises without a GPU: there is no CPU fallback
    fmt, W, H = (args.format or cfg.fmt), cfg.width, cfg.height
    B = wl.algorithmic_bytes(cfg, fmt)
    if cfg.streams > 1:
        # config 5: independent streams, one frame each per step, sharded stream % world
        my_streams = sh.shard_streams(cfg.streams, world, rank)
        batch = len(my_streams)
        scaling = "strong"
        ovs = [wl.overlay_for(cfg, stream=k) for k in range(4)]      # 4 distinct cue images, reused
        for s_id in my_streams:
            ctx.overlay_set(s_id, ovs[s_id % len(ovs)], wl.region_rects(cfg))
        stream_ids = my_streams
    else:
        batch = cfg.batch
        scaling = "weak"
        ov = wl.overlay_for(cfg)
        ctx.overlay_set(1, ov, wl.region_rects(cfg))
        stream_ids = [1] * batch
    ctx.set_batch(min(batch, 1024), 0)     # one launch per `batch` frames, no linger timer

    base = wl.frame_for(cfg, rank, fmt)
    srcs = [ctx.acquire(fmt, W, H) for _ in range(batch)]
    dsts = [ctx.acquire(fmt, W, H) for _ in range(batch)]
    for i, s in enumerate(srcs):
        s.upload([np.roll(p, i * 16, axis=1) for p in base])
    stream_id = stream_ids[0]

    tb_batch = ctx.Batch(stream_ids, fmt, W, H, [s.c for s in srcs], [d.c for d in dsts])

    def step():
        ctx.submit_many(tb_batch)          # one C call; the batch limit launches it
        if batch > 1024:
            ctx.flush()

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        step()
    ctx.sync()
    ctx.stats_reset()
    ctx.set_profiling(args.profile_every)

    barrier(dist, device)
    t_wall0 = time.time()
    ctx.timer_begin()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_end()
    ctx.sync()
    t_wall1 = time.time()
    barrier(dist, device)
    st = ctx.stats()
    ctx.set_profiling(0)

    # clocks: if the timed region was too short for nvidia-smi to sample, keep the same
    # workload running (untimed) until there are samples
    clocks = sampler.summary(t_wall0, t_wall1)
    clocks["window"] = "timed region"
    if clocks["samples"] < 3:
        t0 = time.time()
        while time.time() - t0 < 1.5:
            for _ in range(50):
                step()
            ctx.sync()
        clocks = sampler.summary(t0, time.time())
        clocks["window"] = "same workload re-run for 1.5 s right after the timed region"

    worst_ms = sh.reduce_max(ms, dist, device)
    records = sh.gather_records((batch * args.steps, ms, st["kernel_ms"], st["kernel_ms_launches"]),
                                dist, device)
    value = sh.aggregate_fps(records)

    # roofline of the dominant (only) kernel, this rank
    peak, peak_src = load_peaks()
    launch_ms = st["kernel_ms"] / max(1, st["kernel_ms_launches"])
    achieved = (B * batch) / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None,
                "kernel": "ttmlblend_group_kernel<%s>" % ("PLANE8" if fmt in ("I420", "NV12", "YV12", "NV21")
                                                          else "PACKED"), "launch_ms": launch_ms,
                "launches_timed": int(st["kernel_ms_launches"]),
                "timing": f"CUDA-event pair around every {args.profile_every}th launch inside the timed region",
                "algorithmic_bytes_per_launch": B * batch, "peak_source": peak_src,
                "frac_of_8000_nominal": achieved / 8000.0}
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof) and args.config == 3 and fmt == cfg.fmt:
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
            if roofline["traffic"] and launch_ms > 0:
                # the same launch time against the bytes DRAM really moved (ncu): below the
                # algorithmic figure because the prepared overlay is 3 B/px and comes from L2
                roofline["traffic_gbs"] = roofline["traffic"] / (launch_ms * 1e-3) / 1e9
                roofline["frac_of_traffic"] = roofline["traffic_gbs"] / peak
        except Exception:   # noqa: BLE001
            pass

    inplace = None
    if args.inplace:
        ip_batch = ctx.Batch(stream_ids, fmt, W, H, [s.c for s in srcs], [s.c for s in srcs])
        ctx.stats_reset()
        ctx.timer_begin()
        for _ in range(args.steps):
            ctx.submit_many(ip_batch)
        ms_ip = ctx.timer_end()
        st_ip = ctx.stats()
        inplace = {"value": batch * args.steps / (ms_ip * 1e-3), "unit": UNIT,
                   "bytes_per_frame": st_ip["algorithmic_bytes"] // max(1, st_ip["frames_blended"]),
                   "algorithmic_gbs": st_ip["algorithmic_bytes"] / (ms_ip * 1e-3) / 1e9,
                   "note": "dst == src: only the rows under the cue regions are read and written "
                           "(B_inplace = 2 x touched frame bytes + 4 B/px overlay); this rank only"}

    # e2e: pinned host frames through the drop-in call, PCIe copies inside the timed region
    e2e = None
    if not args.no_e2e:
        for f in dsts:
            f.release()
        # two sets of host frames: a set is handed over while the previous one is still crossing
        # PCIe, as the buffers of a running pipeline are (one set would serialise host and bus)
        host_sets = [[ctx.acquire(fmt, W, H, on_host=True) for _ in range(batch)] for _ in range(2)]
        hosts = host_sets[0]
        for hs in host_sets:
            for i, hf in enumerate(hs):
                for dstp, srcp in zip(hf.host_planes(), base):
                    dstp[...] = np.roll(srcp, i * 16, axis=1)
        e2e_steps = max(4, min(args.steps, 100))
        host_batches = [ctx.Batch(stream_ids, fmt, W, H, [hf.c for hf in hs], [hf.c for hf in hs])
                        for hs in host_sets]

        def e2e_run(n_steps):
            prev = None
            for i in range(n_steps):
                tickets = ctx.blend_host_many(host_batches[i & 1])   # one C call per batch of host frames
                if prev is not None:
                    ctx.wait(prev)                                   # the set submitted one step earlier
                prev = tickets[len(tickets) - 1]                     # tickets complete in order
            ctx.wait(prev)

        e2e_run(4)
        ctx.sync()
        ctx.stats_reset()
        barrier(dist, device)
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        ctx.sync()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        barrier(dist, device)
        st2 = ctx.stats()
        rec2 = sh.gather_records((batch * e2e_steps, e2e_ms), dist, device)
        # the same layout with OPAQUE region boxes (opacity 1.0, the common broadcast style): the
        # result under an opaque vector does not depend on the frame, so in place the frame is
        # written without being read and crosses PCIe in one direction only. Reported beside the
        # headline, not instead of it (the headline cue is translucent: both directions).
        opaque_fps = None
        if cfg.streams == 1:
            try:
                import dataclasses
                ocfg = dataclasses.replace(cfg, regions=[dataclasses.replace(r, opacity=1.0) for r in cfg.regions])
                ctx.overlay_set(2, wl.overlay_for(ocfg), wl.region_rects(ocfg))
                ob = [ctx.Batch([2] * batch, fmt, W, H, [hf.c for hf in hs], [hf.c for hf in hs]) for hs in host_sets]

                def opaque_run(n_steps):
                    prev = None
                    for i in range(n_steps):
                        tickets = ctx.blend_host_many(ob[i & 1])
                        if prev is not None:
                            ctx.wait(prev)
                        prev = tickets[len(tickets) - 1]
                    ctx.wait(prev)

                opaque_run(4)
                ctx.sync()
                n_op = max(4, min(e2e_steps, 40))
                t0 = time.perf_counter()
                opaque_run(n_op)
                ctx.sync()
                opaque_fps = batch * n_op / (time.perf_counter() - t0)
            except Exception as e:      # noqa: BLE001
                opaque_fps = f"failed: {e!r}"
        # the same frames, one synchronous call per frame through the C mirror of the GStreamer
        # call (fluc_video_overlay_composition_blend == gst_video_overlay_composition_blend):
        # what a single streaming thread sees; not batched, so latency-bound
        sync_fps = None
        if cfg.streams == 1:
            try:
                vo = pkg.videooverlay
                if vo.load_library().fluc_video_overlay_set_device(local) == 0:
                    comp = vo.Composition(vo.Rectangle(ov, 0, 0, vo.FLAG_PREMULTIPLIED_ALPHA))
                    views = [hf.host_planes() for hf in hosts]
                    for v in views[:4]:
                        comp.blend(fmt, W, H, v)
                    n_sync = 0
                    t0 = time.perf_counter()
                    while time.perf_counter() - t0 < 1.0:
                        comp.blend(fmt, W, H, views[n_sync % len(views)])
                        n_sync += 1
                    sync_fps = n_sync / (time.perf_counter() - t0)
                    del comp
            except Exception as e:      # noqa: BLE001
                sync_fps = f"failed: {e!r}"
        e2e = {"value": sh.aggregate_fps(rec2), "unit": UNIT,
               "one_synchronous_call_per_frame": sync_fps,
               "same_layout_with_opaque_boxes": opaque_fps,
               "h2d_bytes_per_step": st2["h2d_bytes"] // e2e_steps,
               "d2h_bytes_per_step": st2["d2h_bytes"] // e2e_steps,
               "steps": e2e_steps, "launches": st2["launches"], "host_frames_numa_node": ctx.numa_node(),
               "api": "fluc_ttmlblend_blend_host_many on pinned host frames, in place: the kernel reads the rows "
                      "under the cue regions from host memory and writes them back over PCIe (zero copy), "
                      "one launch per batch; two sets of host frames alternate so that a batch is "
                      "submitted while the previous one is on the bus"}
    sampler.stop()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_baseline(cfg, wl)

    if rank == 0:
        line = {
            "metric": METRIC if args.config == 3 else f"frames/sec, TTML overlay blend ({cfg.name})",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": worst_ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg.name, "format": fmt, "width": W, "height": H,
                       "frames_per_launch": batch, "streams": cfg.streams,
                       "regions": wl.region_rects(cfg),
                       "mode": "out-of-place (whole frame read + written)",
                       "bytes_per_frame": B,
                       "l2": f"inputs larger than L2 ({batch * wl.frame_bytes(fmt, W, H) / 1e6:.0f} MB read + as "
                             "much written per step, L2 is 126 MB)"
                             if batch * wl.frame_bytes(fmt, W, H) > 126e6 else
                             "L2 flushed between steps is NOT done: inputs fit in L2 (small config)",
                       "parallelism": (f"{cfg.streams} streams sharded stream % {world}, no collective"
                                       if cfg.streams > 1 else
                                       f"{world} x independent frame batches, no collective")},
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "clocks": clocks,
            "gstreamer": ("present: " + shutil.which("gst-launch-1.0")) if shutil.which("gst-launch-1.0")
            else "absent on this box (no gst-launch-1.0): the reference pipeline itself cannot be timed",
            "gpu_launches": int(st["launches"]),
            "per_rank": [{"frames": r[0], "ms": r[1], "kernel_ms": r[2]} for r in records],
        }
        if inplace:
            line["inplace"] = inplace
        print(json.dumps(line), file=real_stdout, flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
