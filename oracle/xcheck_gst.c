/*
 * xcheck_gst.c -- pins the oracle against a REAL libgstvideo-1.0.
 *
 * TEST INFRASTRUCTURE. Cannot be built in the graft image (no GLib/GStreamer);
 * `make -C oracle xcheck_gst` works wherever `pkg-config gstreamer-video-1.0`
 * does. It runs gst_video_overlay_composition_blend () -- the call the
 * reference pipeline ends up in for ttmlrender's BGRA buffers
 * (/root/reference/plugins/ttml/gstttmlrender.c:78-84,1427-1478) -- and
 * tbref_composition_blend () on the same seeded inputs for every supported
 * destination format and prints the number of differing bytes. Record the
 * printed GStreamer version in DESIGN.md next to the result; until someone
 * has done that, parity is "unpinned".
 */
#include <gst/gst.h>
#include <gst/video/video.h>
#include <gst/video/video-overlay-composition.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ttmlblend_ref.h"

static guint64 sm_state;

static guint64
splitmix64 (void)
{
  guint64 z = (sm_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static const struct {
  GstVideoFormat gst;
  int ref;
  const char *name;
} formats[] = {
  { GST_VIDEO_FORMAT_I420, TBREF_FORMAT_I420, "I420" },
  { GST_VIDEO_FORMAT_YV12, TBREF_FORMAT_YV12, "YV12" },
  { GST_VIDEO_FORMAT_NV12, TBREF_FORMAT_NV12, "NV12" },
  { GST_VIDEO_FORMAT_NV21, TBREF_FORMAT_NV21, "NV21" },
  { GST_VIDEO_FORMAT_AYUV, TBREF_FORMAT_AYUV, "AYUV" },
  { GST_VIDEO_FORMAT_ARGB, TBREF_FORMAT_ARGB, "ARGB" },
  { GST_VIDEO_FORMAT_ABGR, TBREF_FORMAT_ABGR, "ABGR" },
  { GST_VIDEO_FORMAT_RGBA, TBREF_FORMAT_RGBA, "RGBA" },
  { GST_VIDEO_FORMAT_BGRA, TBREF_FORMAT_BGRA, "BGRA" },
  /* padded RGB: expected to behave like their alpha twins (same pack/unpack) */
  { GST_VIDEO_FORMAT_RGBx, TBREF_FORMAT_RGBA, "RGBx" },
  { GST_VIDEO_FORMAT_BGRx, TBREF_FORMAT_BGRA, "BGRx" },
  { GST_VIDEO_FORMAT_xRGB, TBREF_FORMAT_ARGB, "xRGB" },
  { GST_VIDEO_FORMAT_xBGR, TBREF_FORMAT_ABGR, "xBGR" },
  { GST_VIDEO_FORMAT_Y42B, TBREF_FORMAT_Y42B, "Y42B" },
  { GST_VIDEO_FORMAT_Y444, TBREF_FORMAT_Y444, "Y444" },
  { GST_VIDEO_FORMAT_YUY2, TBREF_FORMAT_YUY2, "YUY2" },
  { GST_VIDEO_FORMAT_UYVY, TBREF_FORMAT_UYVY, "UYVY" },
  { GST_VIDEO_FORMAT_GRAY8, TBREF_FORMAT_GRAY8, "GRAY8" },
  { GST_VIDEO_FORMAT_NV16, TBREF_FORMAT_NV16, "NV16" },
  { GST_VIDEO_FORMAT_NV24, TBREF_FORMAT_NV24, "NV24" },
  { GST_VIDEO_FORMAT_NV61, TBREF_FORMAT_NV61, "NV61" },
  { GST_VIDEO_FORMAT_YVYU, TBREF_FORMAT_YVYU, "YVYU" },
  { GST_VIDEO_FORMAT_VYUY, TBREF_FORMAT_VYUY, "VYUY" },
  { GST_VIDEO_FORMAT_v308, TBREF_FORMAT_v308, "v308" },
  { GST_VIDEO_FORMAT_IYU2, TBREF_FORMAT_IYU2, "IYU2" },
  { GST_VIDEO_FORMAT_RGB, TBREF_FORMAT_RGB, "RGB" },
  { GST_VIDEO_FORMAT_BGR, TBREF_FORMAT_BGR, "BGR" },
};

/* geometry / flag cases: frame size, rectangle size and position (hanging over every
 * border, odd origins and sizes), premultiplied or straight source, global alpha */
static const struct {
  int W, H, RW, RH, RX, RY;
  gboolean premul;
  gfloat ga;
  int render_w, render_h;       /* 0 = pixel size; else the composition scales the rectangle first */
} cases[] = {
  { 321, 181, 200, 90, 37, 51, TRUE, 1.0f },
  { 321, 181, 200, 90, -33, -17, TRUE, 1.0f },
  { 321, 181, 200, 90, 250, 150, TRUE, 1.0f },
  { 320, 180, 320, 180, 0, 0, TRUE, 1.0f },       /* ttmlrender: frame-sized image at (0,0) */
  { 63, 47, 31, 15, 7, 9, TRUE, 1.0f },
  { 64, 48, 1, 1, 5, 5, TRUE, 1.0f },
  { 321, 181, 200, 90, 37, 51, FALSE, 1.0f },
  { 321, 181, 200, 90, 36, 50, TRUE, 0.5f },
  { 321, 181, 200, 90, 37, 51, FALSE, 0.8f },
  /* gst_video_blend_scale_linear_RGBA: up, down (> 2x: the line cache jumps), mixed */
  { 321, 181, 100, 45, 37, 51, TRUE, 1.0f, 200, 90 },
  { 321, 181, 200, 90, 37, 51, TRUE, 1.0f, 61, 29 },
  { 321, 181, 64, 64, -9, 100, TRUE, 1.0f, 257, 33 },
  { 321, 181, 97, 53, 10, 10, FALSE, 0.5f, 150, 70 },
};

int
main (int argc, char **argv)
{
  guint f, k, total_bad = 0;
  gst_init (&argc, &argv);
  printf ("GStreamer %s\n", gst_version_string ());

  for (k = 0; k < G_N_ELEMENTS (cases); k++)
  for (f = 0; f < G_N_ELEMENTS (formats); f++) {
    const int W = cases[k].W, H = cases[k].H, RW = cases[k].RW, RH = cases[k].RH;
    const int RX = cases[k].RX, RY = cases[k].RY;
    int opaque;
    for (opaque = 1; opaque >= 0; opaque--) {
      GstVideoInfo info, rinfo;
      GstBuffer *fbuf, *rbuf;
      GstVideoFrame frame;
      GstVideoOverlayRectangle *rect;
      GstVideoOverlayComposition *comp;
      GstMapInfo map;
      guint8 *copy, *rpix;
      TbRefFrame rf;
      TbRefRectangle rr;
      gsize i, bad = 0;
      guint p;

      sm_state = 0x74746d6c + (k * 64 + f) * 2 + opaque;
      gst_video_info_set_format (&info, formats[f].gst, W, H);
      fbuf = gst_buffer_new_allocate (NULL, info.size, NULL);
      gst_buffer_map (fbuf, &map, GST_MAP_WRITE);
      for (i = 0; i < map.size; i++)
        map.data[i] = (guint8) splitmix64 ();
      if (opaque && GST_VIDEO_INFO_HAS_ALPHA (&info)) {
        guint aoff = GST_VIDEO_INFO_COMP_POFFSET (&info, GST_VIDEO_COMP_A);
        for (i = aoff; i < map.size; i += 4)
          map.data[i] = 255;
      }
      copy = g_memdup2 (map.data, map.size);
      gst_buffer_unmap (fbuf, &map);

      /* BGRA rectangle; premultiplied like Cairo ARGB32, or straight */
      gst_video_info_set_format (&rinfo, GST_VIDEO_OVERLAY_COMPOSITION_FORMAT_RGB, RW, RH);
      rpix = g_malloc (RW * RH * 4);
      for (i = 0; i < (gsize) RW * RH; i++) {
        guint a = splitmix64 () & 0xff, c;
        if ((splitmix64 () & 7) == 0) a = 0;
        if ((splitmix64 () & 7) == 1) a = 255;
        for (c = 0; c < 3; c++) {
          guint v = splitmix64 () & 0xff;
          rpix[4 * i + c] = (guint8) (cases[k].premul ? (v * a + 127) / 255 : v);
        }
        rpix[4 * i + 3] = (guint8) a;
      }
      rbuf = gst_buffer_new_wrapped (g_memdup2 (rpix, RW * RH * 4), RW * RH * 4);
      gst_buffer_add_video_meta (rbuf, GST_VIDEO_FRAME_FLAG_NONE,
          GST_VIDEO_OVERLAY_COMPOSITION_FORMAT_RGB, RW, RH);
      rect = gst_video_overlay_rectangle_new_raw (rbuf, RX, RY,
          cases[k].render_w ? cases[k].render_w : RW, cases[k].render_h ? cases[k].render_h : RH,
          cases[k].premul ? GST_VIDEO_OVERLAY_FORMAT_FLAG_PREMULTIPLIED_ALPHA :
          GST_VIDEO_OVERLAY_FORMAT_FLAG_NONE);
      gst_video_overlay_rectangle_set_global_alpha (rect, cases[k].ga);
      comp = gst_video_overlay_composition_new (rect);

      gst_video_frame_map (&frame, &info, fbuf, GST_MAP_READWRITE);
      gst_video_overlay_composition_blend (comp, &frame);
      gst_video_frame_unmap (&frame);

      memset (&rf, 0, sizeof rf);
      rf.format = formats[f].ref;
      rf.width = W;
      rf.height = H;
      for (p = 0; p < GST_VIDEO_INFO_N_PLANES (&info); p++) {
        rf.data[p] = copy + GST_VIDEO_INFO_PLANE_OFFSET (&info, p);
        rf.stride[p] = GST_VIDEO_INFO_PLANE_STRIDE (&info, p);
      }
      memset (&rr, 0, sizeof rr);
      rr.pixels = rpix;
      rr.width = RW;
      rr.height = RH;
      rr.stride = RW * 4;
      rr.x = RX;
      rr.y = RY;
      rr.global_alpha = cases[k].ga;
      rr.flags = cases[k].premul ? TBREF_FLAG_PREMULTIPLIED_ALPHA : 0;
      rr.render_width = cases[k].render_w;
      rr.render_height = cases[k].render_h;
      tbref_composition_blend (&rf, &rr, 1);

      gst_buffer_map (fbuf, &map, GST_MAP_READ);
      for (i = 0; i < map.size; i++)
        bad += map.data[i] != copy[i];
      gst_buffer_unmap (fbuf, &map);
      printf ("case %u %-5s dest alpha %-7s: %" G_GSIZE_FORMAT " differing bytes of %" G_GSIZE_FORMAT "\n",
          k, formats[f].name, opaque ? "opaque" : "random", bad, (gsize) info.size);
      total_bad += bad;

      gst_video_overlay_composition_unref (comp);
      gst_video_overlay_rectangle_unref (rect);
      gst_buffer_unref (rbuf);
      gst_buffer_unref (fbuf);
      g_free (copy);
      g_free (rpix);
    }
  }
  printf ("%s\n", total_bad ? "MISMATCH" : "oracle == libgstvideo on all vectors");
  return total_bad ? 1 : 0;
}
