"""Shared test helpers: a numpy model of the blend's NET semantics (independent of the
oracle's line-by-line structure), random inputs, and GPU runners through the C ABI."""
from __future__ import annotations

import numpy as np

from oracle import oracle
import __graft_entry__ as graft

pkg = graft.load_package()
wl = pkg.workloads

PLANAR_420 = ("I420", "YV12", "NV12", "NV21")
# byte-plane formats beyond 4:2:0 (GStreamer spells v308 in lower case)
MORE_YUV = ("Y42B", "Y444", "YUY2", "UYVY", "GRAY8", "NV16", "NV24", "NV61", "YVYU", "VYUY", "v308", "IYU2")
YUV_FORMATS = tuple(f.upper() for f in PLANAR_420 + ("AYUV",) + MORE_YUV)
PACKED = ("AYUV", "ARGB", "ABGR", "RGBA", "BGRA")
RGB24 = ("RGB", "BGR")                     # no alpha byte: the destination is opaque
ALL_FORMATS = PLANAR_420 + PACKED + MORE_YUV + RGB24


def yuv_views(fmt, planes, w):
    """(Y, U, V, sub_x, sub_y): writable strided views of the luma / chroma samples of a
    byte-plane YUV frame (U, V None for GRAY8)."""
    if fmt in ("I420", "Y42B", "Y444"):
        return planes[0], planes[1], planes[2], (1 if fmt == "Y444" else 2), (2 if fmt == "I420" else 1)
    if fmt == "YV12":
        return planes[0], planes[2], planes[1], 2, 2
    if fmt == "NV12":
        return planes[0], planes[1][:, 0::2], planes[1][:, 1::2], 2, 2
    if fmt == "NV21":
        return planes[0], planes[1][:, 1::2], planes[1][:, 0::2], 2, 2
    if fmt == "NV16":
        return planes[0], planes[1][:, 0::2], planes[1][:, 1::2], 2, 1
    if fmt == "NV61":
        return planes[0], planes[1][:, 1::2], planes[1][:, 0::2], 2, 1
    if fmt == "YVYU":
        return planes[0][:, 0::2][:, :w], planes[0][:, 3::4], planes[0][:, 1::4], 2, 1
    if fmt == "VYUY":
        return planes[0][:, 1::2][:, :w], planes[0][:, 2::4], planes[0][:, 0::4], 2, 1
    if fmt == "V308":
        return planes[0][:, 0::3], planes[0][:, 1::3], planes[0][:, 2::3], 1, 1
    if fmt == "IYU2":
        return planes[0][:, 1::3], planes[0][:, 0::3], planes[0][:, 2::3], 1, 1
    if fmt == "NV24":
        return planes[0], planes[1][:, 0::2], planes[1][:, 1::2], 1, 1
    if fmt == "YUY2":
        return planes[0][:, 0::2][:, :w], planes[0][:, 1::4], planes[0][:, 3::4], 2, 1
    if fmt == "UYVY":
        return planes[0][:, 1::2][:, :w], planes[0][:, 0::4], planes[0][:, 2::4], 2, 1
    if fmt == "GRAY8":
        return planes[0], None, None, 1, 1
    raise KeyError(fmt)

# byte positions of (A, c1, c2, c3) inside a packed pixel, c = (Y,U,V) or (R,G,B)
PACKED_ORDER = {"AYUV": (0, 1, 2, 3), "ARGB": (0, 1, 2, 3), "ABGR": (0, 3, 2, 1),
                "RGBA": (3, 0, 1, 2), "BGRA": (3, 2, 1, 0),
                "xRGB": (0, 1, 2, 3), "xBGR": (0, 3, 2, 1), "RGBx": (3, 0, 1, 2), "BGRx": (3, 2, 1, 0)}


def rng(seed):
    return np.random.default_rng(seed)


def random_frame(fmt, w, h, seed, opaque=True, pad=0):
    """Planes with optional row padding (stride = row bytes + pad)."""
    r = rng(seed)
    planes = []
    for rows, rb in wl.plane_shapes(fmt, w, h):
        buf = r.integers(0, 256, size=(rows, rb + pad), dtype=np.uint8)
        planes.append(buf[:, :rb])
    ai = wl.alpha_byte_index(fmt)
    if ai is not None and opaque:
        planes[0][:, ai::4] = 255
    return planes


def random_overlay(w, h, seed, premultiplied=True, density=0.7, pad=0):
    """h x w x 4 BGRA with a mix of transparent, opaque and partial pixels."""
    r = rng(seed)
    buf = np.zeros((h, w + pad, 4), dtype=np.uint8)
    img = buf[:, :w, :]
    a = r.integers(0, 256, size=(h, w), dtype=np.uint8)
    kind = r.random((h, w))
    a[kind > density] = 0
    a[kind < 0.15] = 255
    a[(kind > 0.15) & (kind < 0.2)] = 1
    a[(kind > 0.2) & (kind < 0.25)] = 254
    c = r.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    if premultiplied:
        c = ((c.astype(np.uint32) * a[:, :, None].astype(np.uint32) + 127) // 255).astype(np.uint8)
    img[:, :, :3] = c
    img[:, :, 3] = a
    return img


def copy_planes(planes):
    return [np.ascontiguousarray(p).copy() for p in planes]


# ---------------------------------------------------------------------------
# numpy model: per-plane net semantics (SURVEY.md Appendix A.4 / docs/BLENDSPEC.md)

def _matrix_rgb_to_yuv(r, g, b):
    y = np.clip((47 * r + 157 * g + 16 * b + 4096) >> 8, 0, 255)
    u = np.clip((-26 * r - 87 * g + 112 * b + 32768) >> 8, 0, 255)
    v = np.clip((112 * r - 102 * g - 10 * b + 32768) >> 8, 0, 255)
    return y, u, v


def _over(cs, cd, asrc, adst, ga, sp, dp):
    """The four OVERxy operators on int64 arrays; returns (colour, alpha) for asrc > 0."""
    fa = asrc + adst * (255 - asrc) // 255
    fa1 = np.where(fa == 0, 1, fa)
    num_s = cs * ga if sp else cs * asrc
    if not dp:
        v = (num_s + cd * adst * (255 - asrc) // 255) // fa1
    else:
        v = (num_s + cd * (255 - asrc)) // 255
    return np.minimum(v, 255), fa


def model_scale_rows(sh, dh):
    """Which two horizontally resampled source rows, and which 8-bit weight, each destination
    row of gst_video_blend_scale_linear_RGBA ends up merging: a simulation of its two-line
    cache (slot = row & 1, `y1` bookkeeping) -- docs/BLENDSPEC.md section 10."""
    y_inc = 0 if dh == 1 else ((sh - 1) << 16) // (dh - 1) - 1
    slot = {0: 0}
    y1, acc, plan = 0, 0, []
    for _ in range(dh):
        j, x = acc >> 16, acc & 0xffff
        if x == 0:
            plan.append((slot[j & 1], slot[j & 1], 0))
        else:
            if j > y1:
                slot[j & 1] = j
                y1 += 1
            if j >= y1:
                slot[(j + 1) & 1] = j + 1
                y1 += 1
            plan.append((slot[j & 1], slot[(j + 1) & 1], x >> 8))
        acc += y_inc
    return plan


def model_scale(img, dw, dh):
    """Independent numpy model of gst_video_blend_scale_linear_RGBA (h x w x 4 uint8 in/out)."""
    sh, sw = img.shape[:2]
    x_inc = 0 if dw == 1 else ((sw - 1) << 16) // (dw - 1) - 1
    tmp = np.arange(dw, dtype=np.int64) * x_inc
    sx, f = tmp >> 16, ((tmp >> 8) & 0xff)[None, :, None]
    src = img.astype(np.int64)
    lines = (src[:, sx] * (256 - f) + src[:, sx + 1] * f) >> 8          # every source row resampled
    plan = np.array(model_scale_rows(sh, dh), dtype=np.int64)
    a, b, p = lines[plan[:, 0]], lines[plan[:, 1]], plan[:, 2][:, None, None]
    return (a + (((b - a) * p + 128) >> 8)).astype(np.uint8)


def model_blend(fmt, w, h, planes, rectangles, dest_premul=False, chroma_average=False):
    """Blends in place on `planes` and returns them. chroma_average=True models the library's
    NON-PARITY 2x2 chroma option (fluc_ttmlblend_set_chroma_mode), not GStreamer."""
    fmt = fmt.upper()
    for rc in rectangles:
        px = rc["pixels"]
        rw_, rh_ = int(rc.get("render_width", 0)) or px.shape[1], int(rc.get("render_height", 0)) or px.shape[0]
        if (rw_, rh_) != (px.shape[1], px.shape[0]):
            px = model_scale(px, rw_, rh_)
        px = px.astype(np.int64)
        ga = int(255.0 * float(np.float32(rc.get("global_alpha", 1.0))))
        sp = bool(rc.get("premultiplied", True))
        x, y = int(rc.get("x", 0)), int(rc.get("y", 0))
        rh, rw = px.shape[:2]
        x0, y0, x1, y1 = max(x, 0), max(y, 0), min(x + rw, w), min(y + rh, h)
        if x1 <= x0 or y1 <= y0:
            continue
        sub = px[y0 - y:y1 - y, x0 - x:x1 - x]
        b, g, r, a = sub[..., 0], sub[..., 1], sub[..., 2], sub[..., 3]
        if fmt in YUV_FORMATS:
            if sp:
                an = np.where(a == 0, 1, a)
                r = np.where(a > 0, (r * 255 + a // 2) // an, r)
                g = np.where(a > 0, (g * 255 + a // 2) // an, g)
                b = np.where(a > 0, (b * 255 + a // 2) // an, b)
                sp = False
            c1, c2, c3 = _matrix_rgb_to_yuv(r, g, b)
        else:
            c1, c2, c3 = r, g, b
        asrc = a * ga // 255
        m = asrc > 0
        if fmt not in PACKED_ORDER and fmt not in RGB24:
            Y, U, V, sx, sy = yuv_views(fmt, planes, w)
            yd = Y[y0:y1, x0:x1].astype(np.int64)
            v, _ = _over(c1, yd, asrc, 255, ga, sp, dest_premul)
            Y[y0:y1, x0:x1] = np.where(m, v, yd).astype(np.uint8)
            if U is None:
                continue
            if chroma_average and sx == 2 and sy == 2:
                # every chroma sample with a covered pixel: alpha = mean of the 4 alphas (outside
                # the rectangle counts as 0), colour = alpha-weighted mean
                bx0, bx1, by0, by1 = x0 // 2, (x1 + 1) // 2, y0 // 2, (y1 + 1) // 2
                pa = np.zeros((2 * (by1 - by0), 2 * (bx1 - bx0)), dtype=np.int64)
                pu, pv = pa.copy(), pa.copy()
                sl = (slice(y0 - 2 * by0, y0 - 2 * by0 + (y1 - y0)), slice(x0 - 2 * bx0, x0 - 2 * bx0 + (x1 - x0)))
                pa[sl], pu[sl], pv[sl] = asrc, asrc * c2, asrc * c3
                blk = lambda t: t[0::2, 0::2] + t[0::2, 1::2] + t[1::2, 0::2] + t[1::2, 1::2]
                sa, su, sv = blk(pa), blk(pu), blk(pv)
                ca = (sa + 2) >> 2
                den = np.where(sa == 0, 1, sa)
                cu, cv = (su + sa // 2) // den, (sv + sa // 2) // den
                cm = ca > 0
                nby, nbx = ca.shape
                for plane, cc in ((U, cu), (V, cv)):
                    view = plane[by0:by0 + nby, bx0:bx0 + nbx]
                    d = view.astype(np.int64)
                    view[...] = np.where(cm, (cc * ca + d * (255 - ca)) // 255, d).astype(np.uint8)
                continue
            # chroma sample (bx, by) <- overlay pixel at frame (sx*bx, sy*by) only
            ex0, ey0 = -(-x0 // sx) * sx, -(-y0 // sy) * sy
            if ex0 < x1 and ey0 < y1:
                sl = (slice(ey0 - y0, None, sy), slice(ex0 - x0, None, sx))
                cu, cv, ca, cm = c2[sl], c3[sl], asrc[sl], m[sl]
                by0, bx0 = ey0 // sy, ex0 // sx
                nby, nbx = cu.shape
                for plane, cc in ((U, cu), (V, cv)):
                    view = plane[by0:by0 + nby, bx0:bx0 + nbx]
                    d = view.astype(np.int64)
                    vv, _ = _over(cc, d, ca, 255, ga, sp, dest_premul)
                    view[...] = np.where(cm, vv, d).astype(np.uint8)
        elif fmt in RGB24:
            # three bytes per pixel, adst = 255, the source keeps its premultiplied flag
            P = planes[0]
            rows = P[y0:y1, 3 * x0:3 * x1]
            order = (0, 1, 2) if fmt == "RGB" else (2, 1, 0)
            for idx, cs in zip(order, (c1, c2, c3)):
                cd = rows[:, idx::3].astype(np.int64)
                v, _ = _over(cs, cd, asrc, 255, ga, sp, dest_premul)
                rows[:, idx::3] = np.where(m, v, cd).astype(np.uint8)
        else:
            ia, i1, i2, i3 = PACKED_ORDER[fmt]
            P = planes[0]
            rows = P[y0:y1, 4 * x0:4 * x1]
            adst = rows[:, ia::4].astype(np.int64)
            outs = {}
            for idx, cs in ((i1, c1), (i2, c2), (i3, c3)):
                cd = rows[:, idx::4].astype(np.int64)
                v, fa = _over(cs, cd, asrc, adst, ga, sp, dest_premul)
                outs[idx] = np.where(m, v, cd).astype(np.uint8)
            outs[ia] = np.where(m, fa, adst).astype(np.uint8)
            for idx, val in outs.items():
                rows[:, idx::4] = val
    return planes


def oracle_blend(fmt, w, h, planes, rectangles, dest_premul=False):
    return oracle.composition_blend(fmt, w, h, planes, rectangles, dest_premul)


# ---------------------------------------------------------------------------
# GPU runners (C ABI). mode: "out" (src -> dst), "inplace" (device frame), "host"

def gpu_blend(ctx, fmt, w, h, planes, rectangles=None, mode="out", stream=1, dest_premul=False,
              overlay=None, regions=(), set_overlay=True):
    flags = pkg.ttmlblend.FLAG_PREMULTIPLIED_ALPHA if dest_premul else 0
    if set_overlay:
        if overlay is not None:
            ctx.overlay_set(stream, overlay, regions)
        elif rectangles is not None:
            ctx.overlay_set_rectangles(stream, rectangles)
    if mode == "host":
        out = copy_planes(planes)
        ctx.wait(ctx.blend_host(stream, fmt, w, h, out, flags))
        return out
    src = ctx.acquire(fmt, w, h)
    try:
        src.upload(planes)
        if mode == "inplace":
            ctx.wait(ctx.submit(stream, fmt, w, h, src.c, src.c, flags))
            return src.download()
        dst = ctx.acquire(fmt, w, h)
        try:
            ctx.wait(ctx.submit(stream, fmt, w, h, src.c, dst.c, flags))
            return dst.download()
        finally:
            dst.release()
    finally:
        src.release()


def assert_planes_equal(got, want, what=""):
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(got, want)):
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            y, x = bad[0]
            raise AssertionError(
                f"{what} plane {i}: {len(bad)} bytes differ; first at row {y} byte {x}: "
                f"got {int(a[y, x])} want {int(b[y, x])}")
