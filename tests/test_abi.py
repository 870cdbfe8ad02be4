"""CPU tests of the drop-in boundary: the shared library loads without a GPU, exports every
symbol include/fluc_ttmlblend.h declares, and fails loudly (no CPU fallback) when no CUDA
device exists. No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

import __graft_entry__ as graft

pkg = graft.load_package()
tb = pkg.ttmlblend

HEADER = os.path.join(graft.ROOT, "include", "fluc_ttmlblend.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"FLUC_EXPORT[^;]*?\b(fluc_ttmlblend_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    for must in ("fluc_ttmlblend_new", "fluc_ttmlblend_free", "fluc_ttmlblend_overlay_set",
                 "fluc_ttmlblend_overlay_set_rectangles", "fluc_ttmlblend_overlay_clear",
                 "fluc_ttmlblend_submit", "fluc_ttmlblend_flush", "fluc_ttmlblend_wait",
                 "fluc_ttmlblend_blend_host", "fluc_ttmlblend_frame_pool_acquire",
                 "fluc_ttmlblend_frame_pool_release", "fluc_ttmlblend_stats_copy",
                 "fluc_ttmlblend_strerror"):
        assert must in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", tb.LIB_PATH], text=True)
    exported = set(re.findall(r" T (fluc_ttmlblend_\w+)", out))
    declared = set(header_symbols())
    assert declared <= exported, declared - exported
    assert exported <= declared, exported - declared     # nothing undeclared leaks out
    for name in declared:
        assert getattr(lib, name) is not None


def test_python_binding_covers_the_header():
    assert set(tb.PROTOTYPES) == set(header_symbols())


def test_header_is_plain_c(tmp_path):
    """The boundary must compile as C99 with no CUDA / GLib / torch headers."""
    src = tmp_path / "t.c"
    src.write_text('#include "fluc_ttmlblend.h"\nint main(void){FlucTtmlBlendFrame f; FlucTtmlBlendStats s;'
                   '(void)f;(void)s;return FLUC_TTMLBLEND_FORMAT_COUNT==27?0:1;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I",
                           os.path.join(graft.ROOT, "include"), "-c", str(src), "-o",
                           str(tmp_path / "t.o")])


def test_library_has_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", tb.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, out


def test_strerror_and_geometry(lib):
    assert b"no CPU fallback" in lib.fluc_ttmlblend_strerror(tb.ERROR_NO_DEVICE)
    assert lib.fluc_ttmlblend_strerror(0) == b"ok"
    assert lib.fluc_ttmlblend_format_planes(tb.FORMATS["I420"]) == 3
    assert lib.fluc_ttmlblend_format_planes(tb.FORMATS["NV12"]) == 2
    assert lib.fluc_ttmlblend_format_planes(tb.FORMATS["BGRA"]) == 1
    assert lib.fluc_ttmlblend_format_planes(99) == 0
    assert lib.fluc_ttmlblend_plane_row_bytes(tb.FORMATS["NV12"], 1, 1279) == 1280
    assert lib.fluc_ttmlblend_plane_row_bytes(tb.FORMATS["I420"], 2, 1279) == 640
    assert lib.fluc_ttmlblend_plane_rows(tb.FORMATS["I420"], 1, 719) == 360
    assert lib.fluc_ttmlblend_plane_row_bytes(tb.FORMATS["RGBA"], 0, 3840) == 15360
    for fmt in tb.FORMATS:
        lay = tb.plane_layout(fmt, 1279, 719)
        for pl, (rb, rows) in enumerate(lay):
            assert lib.fluc_ttmlblend_plane_row_bytes(tb.FORMATS[fmt], pl, 1279) == rb
            assert lib.fluc_ttmlblend_plane_rows(tb.FORMATS[fmt], pl, 719) == rows


def test_null_context_is_rejected(lib):
    assert lib.fluc_ttmlblend_flush(None) == tb.ERROR_INVALID_ARGUMENT
    assert lib.fluc_ttmlblend_wait(None, 1) == tb.ERROR_INVALID_ARGUMENT
    assert lib.fluc_ttmlblend_new(0, None) == tb.ERROR_INVALID_ARGUMENT
    lib.fluc_ttmlblend_free(None)      # no-op, must not crash


def test_no_gpu_means_loud_failure_not_cpu_fallback(lib):
    if lib.fluc_ttmlblend_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert lib.fluc_ttmlblend_new(0, C.byref(h)) == tb.ERROR_NO_DEVICE
    assert not h.value
    with pytest.raises(pkg.TtmlBlendError):
        pkg.TtmlBlend(0)


def test_product_never_touches_the_oracle():
    """Nothing under the package, include/ or the entry points' product path may reference
    oracle/ (only tests/, smoke() and bench.py's CPU legs may)."""
    bad = []
    for base in (graft.PKG_DIR, os.path.join(graft.ROOT, "include")):
        for dp, _, fns in os.walk(base):
            for fn in fns:
                if fn.endswith((".so", ".o", ".pyc")):
                    continue
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"ttmlblend_ref|tbref_|from oracle|import oracle|oracle/", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_host_mirror_exports_its_header(lib):
    """host/fluc_videooverlay.h (the C mirror of gst_video_overlay_composition_*)."""
    vo = pkg.videooverlay
    if not os.path.exists(vo.LIB_PATH):
        graft.build()
    src = open(os.path.join(graft.PKG_DIR, "host", "fluc_videooverlay.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = set(re.findall(r"FLUC_EXPORT[^;]*?\b(fluc_video_overlay_\w+)\s*\(", src))
    out = subprocess.check_output(["nm", "-D", "--defined-only", vo.LIB_PATH], text=True)
    exported = set(re.findall(r" T (fluc_video_overlay_\w+)", out))
    assert declared == exported == set(vo.PROTOTYPES)
    vlib = vo.load_library()
    # object model without a GPU: refcounts, n_rectangles, global alpha, NULL handling
    import numpy as np
    px = np.zeros((4, 6, 4), dtype=np.uint8)
    r = vo.Rectangle(px, 1, 2)
    assert r.get_global_alpha() == 1.0
    r.set_global_alpha(0.5)
    assert r.get_global_alpha() == 0.5
    r.set_global_alpha(7.0)                      # out of range: ignored, like the g_return_if_fail
    assert r.get_global_alpha() == 0.5
    c = vo.Composition(r)
    c.add_rectangle(r)
    assert c.n_rectangles() == 2
    assert vlib.fluc_video_overlay_rectangle_new_raw(None, 4, 4, 16, 0, 0, 4, 4, 0) is None
    assert vlib.fluc_video_overlay_composition_blend(None, None) == 0
    if lib.fluc_ttmlblend_device_count() == 0:
        planes = [np.zeros((8, 8 * 4), dtype=np.uint8)]
        assert c.blend("BGRA", 8, 8, planes) is False     # FALSE, not a CPU fallback


@pytest.mark.parametrize("src", ["flu-plugins-oss_b200/gst/gstttmlblend.c",
                                 "flu-plugins-oss_b200/gst/gstflucallocator.c", "oracle/xcheck_gst.c"])
def test_gstreamer_glue_is_syntactically_sound(src):
    """GStreamer is not installed here, so the element / allocator / cross-check sources cannot
    be built; they are at least parsed and type-checked against declaration-only stand-ins of
    the GStreamer headers (tests/gst_stub/README.md). Proves nothing about linking."""
    r = subprocess.run(["gcc", "-fsyntax-only", "-std=gnu99", "-Wall", "-Werror",
                        "-I", os.path.join(graft.ROOT, "tests", "gst_stub"),
                        "-I", os.path.join(graft.ROOT, "include"), "-I", os.path.join(graft.ROOT, "oracle"),
                        os.path.join(graft.ROOT, src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_gstreamer_glue_builds_against_the_functional_fake():
    """The element and the allocator are compiled (-Wall -Werror) and linked with the functional
    fake of the GStreamer API (tests/gst_stub/gstfake.c) and the test harness; the library loads
    and exports the harness entry points. What it does on a GPU: tests/test_gpu_gstglue.py."""
    import ctypes
    stub = os.path.join(graft.ROOT, "tests", "gst_stub")
    r = subprocess.run(["make", "-s", "-C", stub], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(os.path.join(stub, "libgstglue_test.so"))
    for sym in ("th_new", "th_start", "th_segment", "th_push_subtitle", "th_subtitle_event", "th_push_video",
                "th_stats", "th_free"):
        assert hasattr(lib, sym), sym
    # no GPU here: the element is created, refuses to start without a CUDA device, and goes away cleanly
    lib.th_new.restype = ctypes.c_void_p
    lib.th_new.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.th_start.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    lib.th_free.argtypes = [ctypes.c_void_p]
    lib.th_errors.argtypes = [ctypes.c_void_p]
    import flu_plugins_oss_b200 as pkg_
    if pkg_.load_library().fluc_ttmlblend_device_count() == 0:
        h = lib.th_new(0, 0)
        assert h
        assert lib.th_start(h, b"NV12", 640, 360) == -1          # start () fails: no usable CUDA device
        assert lib.th_errors(h) == 1                              # and says so through GST_ELEMENT_ERROR
        lib.th_free(h)
