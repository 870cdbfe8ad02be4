/*
 * ttmlblend_ref.h -- CPU ORACLE for the TTML overlay-blend hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path (the package under
 * flu-plugins-oss_b200/, include/) may include, link or call this. Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker or the timed CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic restated here lives in GStreamer's
 * gst-plugins-base (libgstvideo-1.0: gst-libs/gst/video/video-blend.c,
 * video-overlay-composition.c, video-format.c), which the reference only
 * requires as `gstreamer-video-1.0 >= 1.19` (/root/reference/meson.build:13-17)
 * and does not vendor. The reference has no call site into it and no test
 * that pins a pixel. This file restates the published
 * algorithm (docs/BLENDSPEC.md); oracle/xcheck_gst.c diffs it against a real
 * libgstvideo wherever one is installed. One part is pinned by the reference's
 * own code: tbref_gaussian_kernel against plugins/ttml/gstttmlblur.c compiled
 * into oracle/_ref (oracle/refstub/README.md, tests/test_oracle.py).
 *
 * What the overlay contents are (premultiplied, native-endian ARGB32 = bytes
 * B,G,R,A; cleared to 0) is fixed by the reference itself:
 * /root/reference/plugins/ttml/gstttmlrender.c:1442-1452 (buffer W*H*4,
 * CAIRO_FORMAT_ARGB32, stride W*4, CLEAR paint) and :78-84 (src caps BGRA).
 */
#ifndef TTMLBLEND_REF_H
#define TTMLBLEND_REF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Destination frame formats (same numbering as include/fluc_ttmlblend.h). */
enum {
  TBREF_FORMAT_I420 = 0,
  TBREF_FORMAT_NV12 = 1,
  TBREF_FORMAT_AYUV = 2,
  TBREF_FORMAT_RGBA = 3,
  TBREF_FORMAT_BGRA = 4,
  TBREF_FORMAT_YV12 = 5,
  TBREF_FORMAT_NV21 = 6,
  TBREF_FORMAT_ARGB = 7,
  TBREF_FORMAT_ABGR = 8,
  /* 9..12 (RGBx, BGRx, xRGB, xBGR) are aliases of 3, 4, 7, 8 and are mapped by the caller */
  TBREF_FORMAT_Y42B = 13,
  TBREF_FORMAT_Y444 = 14,
  TBREF_FORMAT_YUY2 = 15,
  TBREF_FORMAT_UYVY = 16,
  TBREF_FORMAT_GRAY8 = 17,
  TBREF_FORMAT_NV16 = 18,
  TBREF_FORMAT_NV24 = 19,
  TBREF_FORMAT_NV61 = 20,
  TBREF_FORMAT_YVYU = 21,
  TBREF_FORMAT_VYUY = 22,
  TBREF_FORMAT_v308 = 23,
  TBREF_FORMAT_IYU2 = 24,
  TBREF_FORMAT_RGB = 25,
  TBREF_FORMAT_BGR = 26
};

#define TBREF_FLAG_PREMULTIPLIED_ALPHA 1u

/* Shape of a mapped GstVideoFrame, reduced to what gst_video_blend reads. */
typedef struct {
  int32_t format;
  int32_t width, height;
  uint32_t flags;          /* TBREF_FLAG_PREMULTIPLIED_ALPHA on the DEST */
  uint8_t *data[3];
  int32_t stride[3];
} TbRefFrame;

/* Shape of a GstVideoOverlayRectangle. */
typedef struct {
  const uint8_t *pixels;   /* BGRA byte order (ARGB32 little endian) */
  int32_t width, height, stride;
  int32_t x, y;            /* position in the frame; may be negative */
  float global_alpha;      /* 1.0 in every ttmlrender use */
  uint32_t flags;          /* TBREF_FLAG_PREMULTIPLIED_ALPHA: Cairo data */
  int32_t render_width, render_height;   /* 0 = pixel size (ttmlrender); else scaled first */
} TbRefRectangle;

/* gst_video_blend (dest, src, x, y, global_alpha): 1 = TRUE, 0 = FALSE. */
int tbref_video_blend (TbRefFrame *dest, const TbRefRectangle *src);

/* gst_video_overlay_composition_blend (comp, frame): rectangles in order. */
int tbref_composition_blend (TbRefFrame *dest, const TbRefRectangle *rects,
    uint32_t n_rects);

/* gst_video_blend_scale_linear_RGBA: src (>= 2x2) to a tightly packed dest_width x
 * dest_height BGRA image; what composition_blend applies when render size != pixel size. */
void tbref_scale_linear_rgba (const uint8_t *src_pixels, int32_t src_width, int32_t src_height,
    int32_t src_stride, int32_t dest_width, int32_t dest_height, uint8_t *dest_pixels);

/* The three colour matrices of video-blend.c, on an (A,c1,c2,c3) line. */
void tbref_matrix_prea_rgb_to_yuv (uint8_t *line, uint32_t width);
void tbref_matrix_rgb_to_yuv (uint8_t *line, uint32_t width);
void tbref_matrix_yuv_to_rgb (uint8_t *line, uint32_t width);

/* Outline blur. The Gaussian kernel is the
 * reference's own gst_ttml_blur_create_gaussian_kernel
 * (/root/reference/plugins/ttml/gstttmlblur.c:28-67) in pixman 16.16 fixed
 * point; the convolution restates pixman's PIXMAN_FILTER_CONVOLUTION on an
 * a8r8g8b8 image with PIXMAN_REPEAT_NONE, operator SRC (gstttmlblur.c:72-110;
 * pixman is not installed here: parity unpinned for the convolution part).
 * `taps` receives (2*radius+1)^2 values; returns that count. */
int32_t tbref_gaussian_kernel (int32_t radius, double sigma, int32_t *taps);
void tbref_blur_argb32 (const uint8_t *src, int32_t width, int32_t height, int32_t stride,
    int32_t radius, double sigma, uint8_t *dst, int32_t dst_stride);
/* the convolution alone: size x size taps in 16.16 (what pixman gets from the reference) */
void tbref_convolve_argb32 (const uint8_t *src, int32_t width, int32_t height, int32_t stride,
    int32_t size, const int32_t *taps, uint8_t *dst, int32_t dst_stride);

/* Region composition: what gst_ttmlrender_show_regions does
 * around the text (/root/reference/plugins/ttml/gstttmlrender.c:1250-1268,1375-1381) with
 * Cairo's colour conversion and pixman's 8-bit premultiplied OVER / IN restated
 * (docs/BLENDSPEC.md section 9; neither library is installed: parity unpinned). `out` is a
 * cleared-then-drawn frame_w * frame_h premultiplied BGRA image. */
typedef struct {
  int32_t x, y, w, h;
  uint32_t background_color;   /* 0xRRGGBBAA */
  double opacity;
  const uint8_t *layer;        /* optional premultiplied BGRA, w*h */
  int32_t layer_stride;
} TbRefRegion;
void tbref_compose_regions (const TbRefRegion *regions, uint32_t n, int32_t frame_w, int32_t frame_h,
    uint8_t *out, int32_t out_stride);

/* Plane geometry helpers shared by the tests and the bench. */
int32_t tbref_n_planes (int32_t format);
int32_t tbref_plane_row_bytes (int32_t format, int32_t plane, int32_t width);
int32_t tbref_plane_rows (int32_t format, int32_t plane, int32_t height);

/* CPU baseline driver: blends `n_frames` frames (one composition each,
 * frames[i] in place) on `n_threads` pthreads, frames dealt round-robin.
 * Returns wall seconds of the blending only. */
double tbref_blend_many (TbRefFrame *frames, uint32_t n_frames,
    const TbRefRectangle *rects, uint32_t n_rects, uint32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* TTMLBLEND_REF_H */
