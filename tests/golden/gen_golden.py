"""Writes the golden fixtures of tests/golden/ from the CPU oracle.

PARITY UNPINNED / self-referential: the reference holds no vector for the blend path and
GStreamer (the library whose arithmetic is restated) is not installed here, so these files
freeze what oracle/ttmlblend_ref.c computes today. They catch drift of the oracle and give
the GPU tests inputs whose expected output does not depend on building the oracle at test
time. Re-run only on purpose:  python tests/golden/gen_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import copy_planes, oracle_blend, random_frame, random_overlay  # noqa: E402

VECTORS = [
    # name, fmt, w, h, rects (rw, rh, x, y, ga, premul), opaque dest, premult dest
    ("i420_cue", "I420", 96, 64, [(64, 20, 16, 40, 1.0, True)], True, False),
    ("nv12_regions", "NV12", 96, 64, [(70, 12, 13, 3, 1.0, True), (33, 17, 5, 21, 1.0, True),
                                      (80, 14, 9, 47, 1.0, True)], True, False),
    ("nv12_odd", "NV12", 95, 63, [(50, 30, -5, 40, 1.0, True), (31, 9, 70, -3, 1.0, True)], True, False),
    ("yv12_nv21", "YV12", 64, 48, [(40, 20, 11, 13, 1.0, True)], True, False),
    ("nv21", "NV21", 64, 48, [(40, 20, 11, 13, 1.0, True)], True, False),
    ("ayuv_spans", "AYUV", 80, 48, [(60, 20, 10, 20, 1.0, True), (30, 30, 40, 10, 1.0, True)], True, False),
    ("rgba_spans", "RGBA", 80, 48, [(60, 20, 10, 20, 1.0, True), (30, 30, 40, 10, 1.0, True)], True, False),
    ("bgra_spans", "BGRA", 80, 48, [(60, 20, 10, 20, 1.0, True), (30, 30, 40, 10, 1.0, True)], True, False),
    ("bgra_alpha_dest", "BGRA", 80, 48, [(60, 20, 10, 20, 0.6, True), (30, 30, 40, 10, 1.0, False)], False, False),
    ("argb_premul_dest", "ARGB", 80, 48, [(60, 20, 10, 20, 1.0, True)], False, True),
    ("abgr", "ABGR", 40, 24, [(20, 10, 3, 5, 1.0, True)], True, False),
    ("y42b", "Y42B", 63, 40, [(40, 20, 11, 13, 1.0, True), (9, 9, 57, 35, 1.0, True)], True, False),
    ("y444", "Y444", 64, 40, [(40, 20, 11, 13, 1.0, True)], True, False),
    ("yuy2", "YUY2", 63, 40, [(40, 20, 11, 13, 1.0, True), (9, 9, 57, 35, 1.0, False)], True, False),
    ("uyvy", "UYVY", 64, 40, [(41, 20, 10, 13, 1.0, True)], True, False),
    ("gray8", "GRAY8", 50, 30, [(30, 10, 7, 9, 1.0, True)], True, False),
    ("rgbx", "RGBx", 40, 24, [(20, 10, 3, 5, 1.0, True)], False, False),
    ("bgrx", "BGRx", 40, 24, [(20, 10, 3, 5, 1.0, True)], True, False),
    ("xrgb", "xRGB", 40, 24, [(20, 10, 3, 5, 1.0, True)], True, False),
    ("xbgr", "xBGR", 40, 24, [(20, 10, 3, 5, 1.0, True)], True, False),
    # append only: a vector's seed is its position in this list
    ("nv16", "NV16", 63, 40, [(40, 20, 11, 13, 1.0, True)], True, False),
    ("nv24", "NV24", 61, 40, [(40, 20, 11, 13, 1.0, True)], True, False),
    # rectangles with a render size (.., render_w, render_h): scaled before the blend
    ("nv12_scaled", "NV12", 96, 64, [(30, 10, 8, 40, 1.0, True, 80, 21), (40, 40, -6, -4, 1.0, True, 17, 23)],
     True, False),
    ("bgra_scaled", "BGRA", 80, 48, [(25, 12, 10, 20, 0.7, False, 60, 20), (50, 40, 50, 30, 1.0, True, 45, 13)],
     False, False),
    ("nv61", "NV61", 63, 40, [(40, 20, 11, 13, 1.0, True)], True, False),
    ("yvyu", "YVYU", 63, 40, [(40, 20, 11, 13, 1.0, True), (9, 9, 57, 35, 1.0, False)], True, False),
    ("vyuy", "VYUY", 64, 40, [(41, 20, 10, 13, 1.0, True)], True, False),
    ("v308", "v308", 61, 40, [(40, 20, 11, 13, 1.0, True), (9, 9, 55, 35, 0.5, True)], True, False),
    ("iyu2", "IYU2", 62, 40, [(40, 20, 11, 13, 1.0, True)], True, False),
    ("rgb", "RGB", 61, 40, [(40, 20, 11, 13, 1.0, True), (9, 9, 55, 35, 0.5, True), (20, 8, 3, 2, 0.8, False)],
     True, False),
    ("bgr", "BGR", 64, 40, [(41, 20, 10, 13, 1.0, True), (17, 9, 30, 3, 1.0, False)], True, False),
]


def main():
    manifest = {"parity": "unpinned",
                "generator": "tests/golden/gen_golden.py (oracle/ttmlblend_ref.c)",
                "vectors": []}
    for k, (name, fmt, w, h, rects, opaque, dprem) in enumerate(VECTORS):
        planes = random_frame(fmt, w, h, 7000 + k, opaque=opaque)
        planes = copy_planes(planes)
        rectangles = [dict(pixels=random_overlay(rw, rh, 7100 + 10 * k + i, premultiplied=pm),
                           x=x, y=y, global_alpha=ga, premultiplied=pm,
                           render_width=(rs + [0, 0])[0], render_height=(rs + [0, 0])[1])
                      for i, (rw, rh, x, y, ga, pm, *rs) in enumerate(rects)]
        out = oracle_blend(fmt, w, h, copy_planes(planes), rectangles, dprem)
        data = {"width": w, "height": h, "n_planes": len(planes), "n_rects": len(rectangles),
                "dest_premul": dprem,
                "pos": np.array([[r["x"], r["y"]] for r in rectangles], dtype=np.int32),
                "ga": np.array([r["global_alpha"] for r in rectangles], dtype=np.float32),
                "premul": np.array([r["premultiplied"] for r in rectangles], dtype=np.bool_)}
        if any(r["render_width"] for r in rectangles):
            data["render"] = np.array([[r["render_width"], r["render_height"]] for r in rectangles], dtype=np.int32)
        for i, p in enumerate(planes):
            data[f"in{i}"] = p
            data[f"out{i}"] = out[i]
        for i, r in enumerate(rectangles):
            data[f"rect{i}"] = r["pixels"]
        fn = f"{name}.npz"
        np.savez_compressed(os.path.join(HERE, fn), **data)
        h256 = hashlib.sha256()
        for p in out:
            h256.update(np.ascontiguousarray(p).tobytes())
        manifest["vectors"].append({"file": fn, "format": fmt, "sha256": h256.hexdigest()})
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", len(VECTORS), "vectors")


if __name__ == "__main__":
    main()
