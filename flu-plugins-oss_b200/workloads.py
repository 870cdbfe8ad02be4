"""Synthetic frames and TTML cue layouts for the BASELINE.json configs.

Pure numpy, deterministic (splitmix64, seed 0x74746d6c + config index; SURVEY.md
section 8d). Frames are uniform random samples with alpha 255. Overlays look
like what ttmlrender emits (/root/reference/plugins/ttml/gstttmlrender.c:1427-1478):
a frame-sized, cleared, PREMULTIPLIED BGRA image with, per region
(gstttmlrender.c:1235-1385), a background box (colour x opacity) and pseudo
glyphs (opaque cores with an anti-aliased 1 px rim) in a per-line text colour.
Every overlay pixel is valid premultiplied data (colour <= alpha).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

SEED_BASE = 0x74746D6C
_M64 = (1 << 64) - 1


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n 64-bit outputs of splitmix64 started at `seed` (vectorised)."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _M64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def random_bytes(seed: int, n: int) -> np.ndarray:
    words = splitmix64(seed, (n + 7) // 8)
    return words.view(np.uint8)[:n].copy()


@dataclass
class Region:
    x: int
    y: int
    w: int
    h: int
    bg: Tuple[int, int, int, int] = (0, 0, 0, 0)      # straight RGBA
    opacity: float = 1.0
    colours: Tuple[Tuple[int, int, int, int], ...] = ((255, 255, 255, 255),)


@dataclass
class Config:
    index: int
    name: str
    fmt: str
    width: int
    height: int
    regions: List[Region]
    batch: int = 1
    streams: int = 1
    extra_formats: Tuple[str, ...] = field(default_factory=tuple)


_SPAN_COLOURS = ((255, 255, 255, 255), (255, 255, 0, 255), (0, 255, 255, 0xC0), (255, 64, 64, 0x80))

CONFIGS = {
    1: Config(1, "720p_i420_single_cue", "I420", 1280, 720,
              [Region(128, 576, 1024, 108)]),
    2: Config(2, "1080p_nv12_3_regions", "NV12", 1920, 1080,
              [Region(192, 54, 1536, 108, (0, 0, 0, 255), 0.8),
               Region(96, 486, 672, 162, (0, 0, 128, 255), 0.5, ((255, 255, 0, 255),)),
               Region(192, 864, 1536, 162, (32, 32, 32, 255), 1.0)]),
    3: Config(3, "4k_nv12_fullwidth_batch32", "NV12", 3840, 2160,
              [Region(0, 1728, 3840, 360, (0, 0, 0, 255), 0.75),
               Region(0, 72, 3840, 144, (0, 0, 0, 255), 0.75, ((255, 255, 0, 255),))],
              batch=32),
    4: Config(4, "4k_packed_per_span_colours", "RGBA", 3840, 2160,
              [Region(0, 1728, 3840, 360, (0, 0, 0, 255), 0.75, _SPAN_COLOURS),
               Region(0, 72, 3840, 144, (16, 16, 64, 0xA0), 1.0, _SPAN_COLOURS)],
              batch=8, extra_formats=("BGRA", "AYUV")),
    5: Config(5, "256x1080p_i420_streams", "I420", 1920, 1080,
              [Region(192, 864, 1536, 162, (0, 0, 0, 255), 0.6)],
              batch=1, streams=256),
}


def region_rects(cfg: Config) -> List[Tuple[int, int, int, int]]:
    return [(r.x, r.y, r.w, r.h) for r in cfg.regions]


def plane_shapes(fmt: str, width: int, height: int):
    f = fmt.upper()
    cw, ch = (width + 1) // 2, (height + 1) // 2
    if f in ("I420", "YV12"):
        return [(height, width), (ch, cw), (ch, cw)]
    if f in ("NV12", "NV21"):
        return [(height, width), (ch, 2 * cw)]
    if f in ("NV16", "NV61"):
        return [(height, width), (height, 2 * cw)]
    if f == "NV24":
        return [(height, width), (height, 2 * width)]
    if f == "Y42B":
        return [(height, width), (height, cw), (height, cw)]
    if f == "Y444":
        return [(height, width)] * 3
    if f in ("YUY2", "UYVY", "YVYU", "VYUY"):
        return [(height, 4 * cw)]
    if f in ("V308", "IYU2", "RGB", "BGR"):
        return [(height, 3 * width)]
    if f == "GRAY8":
        return [(height, width)]
    return [(height, 4 * width)]


def alpha_byte_index(fmt: str):
    f = fmt.upper()
    if f in ("AYUV", "ARGB", "ABGR", "XRGB", "XBGR"):
        return 0
    if f in ("RGBA", "BGRA", "RGBX", "BGRX"):
        return 3
    return None


def make_frame(fmt: str, width: int, height: int, seed: int, opaque: bool = True):
    """Planes (2-D uint8 arrays) of one synthetic frame."""
    planes = []
    for i, (rows, rb) in enumerate(plane_shapes(fmt, width, height)):
        p = random_bytes(seed * 7919 + i, rows * rb).reshape(rows, rb)
        planes.append(p)
    ai = alpha_byte_index(fmt)
    if ai is not None and opaque:
        planes[0][:, ai::4] = 255
    return planes


def make_overlay(width: int, height: int, regions: List[Region], seed: int,
                 cell: Tuple[int, int] = (16, 32)) -> np.ndarray:
    """Frame-sized premultiplied BGRA overlay (H x W x 4 uint8)."""
    img = np.zeros((height, width, 4), dtype=np.uint32)
    cw, chh = cell
    for ri, reg in enumerate(regions):
        x0, y0 = max(reg.x, 0), max(reg.y, 0)
        x1, y1 = min(reg.x + reg.w, width), min(reg.y + reg.h, height)
        if x1 <= x0 or y1 <= y0:
            continue
        w, h = x1 - x0, y1 - y0
        layer = np.zeros((h, w, 4), dtype=np.uint32)      # premultiplied R,G,B,A
        r, g, b, a = reg.bg
        if a:
            layer[:, :, 3] = a
            layer[:, :, 0] = (r * a + 127) // 255
            layer[:, :, 1] = (g * a + 127) // 255
            layer[:, :, 2] = (b * a + 127) // 255
        # pseudo glyphs: coverage mask per cell
        ncx, ncy = (w + cw - 1) // cw, (h + chh - 1) // chh
        rnd = splitmix64(seed * 1315423911 + ri, ncx * ncy * 2)
        on = (rnd[: ncx * ncy] & np.uint64(3)) != 0           # 75 % of the cells hold a glyph
        rim = (rnd[ncx * ncy:] % np.uint64(254) + np.uint64(1)).astype(np.uint32)
        cov = np.zeros((h, w), dtype=np.uint32)
        yy, xx = np.mgrid[0:h, 0:w]
        cy, cx = yy // chh, xx // cw
        ly, lx = yy % chh, xx % cw
        ci = cy * ncx + cx
        inner = (lx >= 3) & (lx < cw - 3) & (ly >= 5) & (ly < chh - 5)
        edge = (lx >= 2) & (lx < cw - 2) & (ly >= 4) & (ly < chh - 4) & ~inner
        cov[inner & on[ci]] = 255
        sel = edge & on[ci]
        cov[sel] = rim[ci][sel]
        # text colour per line of cells
        col = np.array(reg.colours, dtype=np.uint32)[cy % len(reg.colours)]   # h x w x 4 (RGBA straight)
        ga = (col[:, :, 3] * cov + 127) // 255                                # glyph alpha
        glyph = np.zeros((h, w, 4), dtype=np.uint32)
        glyph[:, :, 3] = ga
        for k in range(3):
            glyph[:, :, k] = (col[:, :, k] * ga + 127) // 255
        # glyph OVER background, premultiplied
        inv = 255 - glyph[:, :, 3]
        for k in range(4):
            layer[:, :, k] = glyph[:, :, k] + (layer[:, :, k] * inv + 127) // 255
        # group opacity (cairo_paint_with_alpha)
        if reg.opacity < 1.0:
            o = int(round(reg.opacity * 255))
            layer = (layer * o + 127) // 255
        layer[:, :, :3] = np.minimum(layer[:, :, :3], layer[:, :, 3:4])
        # region OVER image
        inv = 255 - layer[:, :, 3:4]
        dst = img[y0:y1, x0:x1, :]
        img[y0:y1, x0:x1, :] = layer + (dst * inv + 127) // 255
    img[:, :, :3] = np.minimum(img[:, :, :3], img[:, :, 3:4])
    out = np.empty((height, width, 4), dtype=np.uint8)
    out[:, :, 0] = img[:, :, 2]       # B
    out[:, :, 1] = img[:, :, 1]       # G
    out[:, :, 2] = img[:, :, 0]       # R
    out[:, :, 3] = img[:, :, 3]       # A
    return out


def overlay_for(cfg: Config, stream: int = 0) -> np.ndarray:
    return make_overlay(cfg.width, cfg.height, cfg.regions, SEED_BASE + cfg.index + 1000 * stream)


def frame_for(cfg: Config, i: int, fmt: str = None):
    return make_frame(fmt or cfg.fmt, cfg.width, cfg.height, SEED_BASE + cfg.index + 31 * (i + 1))


def frame_bytes(fmt: str, width: int, height: int) -> int:
    return sum(r * c for r, c in plane_shapes(fmt, width, height))


def algorithmic_bytes(cfg: Config, fmt: str = None) -> int:
    """B = frame read + frame write + overlay read at 4 B/px."""
    px = 0
    for r in cfg.regions:
        x0, y0 = max(r.x, 0), max(r.y, 0)
        x1, y1 = min(r.x + r.w, cfg.width), min(r.y + r.h, cfg.height)
        px += max(0, x1 - x0) * max(0, y1 - y0)
    return 2 * frame_bytes(fmt or cfg.fmt, cfg.width, cfg.height) + 4 * px
