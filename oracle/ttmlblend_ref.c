/*
 * ttmlblend_ref.c -- CPU ORACLE (test infrastructure, never shipped).
 *
 * PARITY UNPINNED -- see ttmlblend_ref.h. A line-structured restatement of
 * gst-plugins-base `gst_video_overlay_composition_blend` /
 * `gst_video_blend` (gst-libs/gst/video/video-overlay-composition.c,
 * video-blend.c) and the pack/unpack routines of video-format.c for the
 * destination formats ttmlrender's overlay can meet. The structure is kept
 * like upstream on purpose (unpack whole dest line -> unpack src segment ->
 * matrix -> per-pixel OVER -> pack whole dest line) so that it can be
 * diffed against the real file on a box that has it; docs/BLENDSPEC.md holds
 * the frozen spec, SURVEY.md Appendix A the derivation.
 *
 * Overlay contents come from the reference:
 *   /root/reference/plugins/ttml/gstttmlrender.c:1427-1478  gen_buffer():
 *     W*H*4 bytes, stride W*4, Cairo ARGB32 (premultiplied, bytes B,G,R,A)
 *   /root/reference/plugins/ttml/gstttmlrender.c:1235-1385  show_regions():
 *     background colour / opacity are already baked into those pixels, so
 *     the blend applies nothing but the per-pixel alpha.
 */
#include "ttmlblend_ref.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define TB_MIN(a, b) ((a) < (b) ? (a) : (b))
#define TB_CLAMP(v, lo, hi) ((v) < (lo) ? (lo) : ((v) > (hi) ? (hi) : (v)))

/* ---------------------------------------------------------------------- */
/* geometry                                                               */

static int
is_yuv (int32_t f)
{
  return f == TBREF_FORMAT_I420 || f == TBREF_FORMAT_NV12 ||
      f == TBREF_FORMAT_AYUV || f == TBREF_FORMAT_YV12 ||
      f == TBREF_FORMAT_NV21 || f == TBREF_FORMAT_Y42B ||
      f == TBREF_FORMAT_Y444 || f == TBREF_FORMAT_YUY2 ||
      f == TBREF_FORMAT_UYVY || f == TBREF_FORMAT_GRAY8 ||
      f == TBREF_FORMAT_NV16 || f == TBREF_FORMAT_NV24 ||
      f == TBREF_FORMAT_NV61 || f == TBREF_FORMAT_YVYU ||
      f == TBREF_FORMAT_VYUY || f == TBREF_FORMAT_v308 ||
      f == TBREF_FORMAT_IYU2;
}

/* packed 4:2:2: byte positions of the first luma (the second is 2 further), U and V inside
 * a macropixel -- unpack_/pack_ YUY2, UYVY, YVYU, VYUY of video-format.c */
static void
packed_422_order (int32_t f, int *oy, int *ou, int *ov)
{
  switch (f) {
    case TBREF_FORMAT_YUY2: *oy = 0; *ou = 1; *ov = 3; break;
    case TBREF_FORMAT_UYVY: *oy = 1; *ou = 0; *ov = 2; break;
    case TBREF_FORMAT_YVYU: *oy = 0; *ou = 3; *ov = 1; break;
    default: *oy = 1; *ou = 2; *ov = 0; break;      /* VYUY */
  }
}

int32_t
tbref_n_planes (int32_t f)
{
  switch (f) {
    case TBREF_FORMAT_I420:
    case TBREF_FORMAT_YV12:
    case TBREF_FORMAT_Y42B:
    case TBREF_FORMAT_Y444:
      return 3;
    case TBREF_FORMAT_NV12:
    case TBREF_FORMAT_NV21:
    case TBREF_FORMAT_NV16:
    case TBREF_FORMAT_NV24:
    case TBREF_FORMAT_NV61:
      return 2;
    default:
      return 1;
  }
}

int32_t
tbref_plane_row_bytes (int32_t f, int32_t plane, int32_t w)
{
  switch (f) {
    case TBREF_FORMAT_I420:
    case TBREF_FORMAT_YV12:
    case TBREF_FORMAT_Y42B:
      return plane == 0 ? w : (w + 1) / 2;
    case TBREF_FORMAT_NV12:
    case TBREF_FORMAT_NV21:
    case TBREF_FORMAT_NV16:
    case TBREF_FORMAT_NV61:
      return plane == 0 ? w : 2 * ((w + 1) / 2);
    case TBREF_FORMAT_NV24:
      return plane == 0 ? w : 2 * w;
    case TBREF_FORMAT_Y444:
    case TBREF_FORMAT_GRAY8:
      return w;
    case TBREF_FORMAT_YUY2:
    case TBREF_FORMAT_UYVY:
    case TBREF_FORMAT_YVYU:
    case TBREF_FORMAT_VYUY:
      return 4 * ((w + 1) / 2);
    case TBREF_FORMAT_v308:
    case TBREF_FORMAT_IYU2:
    case TBREF_FORMAT_RGB:
    case TBREF_FORMAT_BGR:
      return 3 * w;
    default:
      return 4 * w;
  }
}

int32_t
tbref_plane_rows (int32_t f, int32_t plane, int32_t h)
{
  if (plane > 0 && (f == TBREF_FORMAT_I420 || f == TBREF_FORMAT_YV12 ||
          f == TBREF_FORMAT_NV12 || f == TBREF_FORMAT_NV21))
    return (h + 1) / 2;
  return h;
}

/* ---------------------------------------------------------------------- */
/* video-format.c: unpack a whole line to (A, c1, c2, c3) 8-bit           */
/* 4:2:0 -- unpack_planar_420 / unpack_NV12: chroma of row y>>1 replicated */
/* over both pixels of a pair (ORC loadupdb); A = 0xff.                    */

static void
unpack_line (const TbRefFrame *f, int y, uint8_t *d, int width)
{
  int x;
  switch (f->format) {
    case TBREF_FORMAT_I420:
    case TBREF_FORMAT_YV12:{
      int pu = f->format == TBREF_FORMAT_I420 ? 1 : 2;
      int pv = f->format == TBREF_FORMAT_I420 ? 2 : 1;
      const uint8_t *sy = f->data[0] + (size_t) f->stride[0] * y;
      const uint8_t *su = f->data[pu] + (size_t) f->stride[pu] * (y >> 1);
      const uint8_t *sv = f->data[pv] + (size_t) f->stride[pv] * (y >> 1);
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = 0xff;
        d[4 * x + 1] = sy[x];
        d[4 * x + 2] = su[x >> 1];
        d[4 * x + 3] = sv[x >> 1];
      }
      break;
    }
    case TBREF_FORMAT_NV12:
    case TBREF_FORMAT_NV21:{
      int ou = f->format == TBREF_FORMAT_NV12 ? 0 : 1;
      const uint8_t *sy = f->data[0] + (size_t) f->stride[0] * y;
      const uint8_t *suv = f->data[1] + (size_t) f->stride[1] * (y >> 1);
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = 0xff;
        d[4 * x + 1] = sy[x];
        d[4 * x + 2] = suv[(x >> 1) * 2 + ou];
        d[4 * x + 3] = suv[(x >> 1) * 2 + (1 - ou)];
      }
      break;
    }
    case TBREF_FORMAT_NV16:
    case TBREF_FORMAT_NV61:
    case TBREF_FORMAT_NV24:{       /* unpack_NV16 / _NV61 / _NV24: UV (VU) interleaved, same line */
      const int sh = f->format == TBREF_FORMAT_NV24 ? 0 : 1;
      const int ou = f->format == TBREF_FORMAT_NV61 ? 1 : 0;
      const uint8_t *sy = f->data[0] + (size_t) f->stride[0] * y;
      const uint8_t *suv = f->data[1] + (size_t) f->stride[1] * y;
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = 0xff;
        d[4 * x + 1] = sy[x];
        d[4 * x + 2] = suv[(x >> sh) * 2 + ou];
        d[4 * x + 3] = suv[(x >> sh) * 2 + (1 - ou)];
      }
      break;
    }
    case TBREF_FORMAT_v308:
    case TBREF_FORMAT_IYU2:{       /* unpack_v308 (Y U V) / unpack_IYU2 (U Y V): 3 bytes per pixel */
      const uint8_t *s = f->data[0] + (size_t) f->stride[0] * y;
      const int oy = f->format == TBREF_FORMAT_v308 ? 0 : 1, ou = f->format == TBREF_FORMAT_v308 ? 1 : 0;
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = 0xff;
        d[4 * x + 1] = s[3 * x + oy];
        d[4 * x + 2] = s[3 * x + ou];
        d[4 * x + 3] = s[3 * x + 2];
      }
      break;
    }
    case TBREF_FORMAT_Y42B:{       /* unpack_Y42B: chroma of the pair, same line */
      const uint8_t *sy = f->data[0] + (size_t) f->stride[0] * y;
      const uint8_t *su = f->data[1] + (size_t) f->stride[1] * y;
      const uint8_t *sv = f->data[2] + (size_t) f->stride[2] * y;
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = 0xff;
        d[4 * x + 1] = sy[x];
        d[4 * x + 2] = su[x >> 1];
        d[4 * x + 3] = sv[x >> 1];
      }
      break;
    }
    case TBREF_FORMAT_Y444:{
      const uint8_t *sy = f->data[0] + (size_t) f->stride[0] * y;
      const uint8_t *su = f->data[1] + (size_t) f->stride[1] * y;
      const uint8_t *sv = f->data[2] + (size_t) f->stride[2] * y;
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = 0xff;
        d[4 * x + 1] = sy[x];
        d[4 * x + 2] = su[x];
        d[4 * x + 3] = sv[x];
      }
      break;
    }
    case TBREF_FORMAT_YUY2:
    case TBREF_FORMAT_UYVY:
    case TBREF_FORMAT_YVYU:
    case TBREF_FORMAT_VYUY:{       /* unpack_YUY2 / _UYVY / _YVYU / _VYUY: macropixels of two pixels */
      const uint8_t *s = f->data[0] + (size_t) f->stride[0] * y;
      int oy, ou, ov;                                          /* Y at oy and oy + 2 */
      packed_422_order (f->format, &oy, &ou, &ov);
      for (x = 0; x < width; x++) {
        const uint8_t *m = s + 4 * (x >> 1);
        d[4 * x + 0] = 0xff;
        d[4 * x + 1] = m[oy + 2 * (x & 1)];
        d[4 * x + 2] = m[ou];
        d[4 * x + 3] = m[ov];
      }
      break;
    }
    case TBREF_FORMAT_RGB:
    case TBREF_FORMAT_BGR:{        /* unpack_RGB / unpack_BGR: A = 0xff */
      const uint8_t *s = f->data[0] + (size_t) f->stride[0] * y;
      const int orr = f->format == TBREF_FORMAT_RGB ? 0 : 2;
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = 0xff;
        d[4 * x + 1] = s[3 * x + orr];
        d[4 * x + 2] = s[3 * x + 1];
        d[4 * x + 3] = s[3 * x + (2 - orr)];
      }
      break;
    }
    case TBREF_FORMAT_GRAY8:{      /* unpack_GRAY8: U = V = 0x80 */
      const uint8_t *sy = f->data[0] + (size_t) f->stride[0] * y;
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = 0xff;
        d[4 * x + 1] = sy[x];
        d[4 * x + 2] = 0x80;
        d[4 * x + 3] = 0x80;
      }
      break;
    }
    case TBREF_FORMAT_AYUV:
    case TBREF_FORMAT_ARGB:
      memcpy (d, f->data[0] + (size_t) f->stride[0] * y, (size_t) width * 4);
      break;
    case TBREF_FORMAT_RGBA:{
      const uint8_t *s = f->data[0] + (size_t) f->stride[0] * y;
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = s[4 * x + 3];
        d[4 * x + 1] = s[4 * x + 0];
        d[4 * x + 2] = s[4 * x + 1];
        d[4 * x + 3] = s[4 * x + 2];
      }
      break;
    }
    case TBREF_FORMAT_BGRA:{
      const uint8_t *s = f->data[0] + (size_t) f->stride[0] * y;
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = s[4 * x + 3];
        d[4 * x + 1] = s[4 * x + 2];
        d[4 * x + 2] = s[4 * x + 1];
        d[4 * x + 3] = s[4 * x + 0];
      }
      break;
    }
    case TBREF_FORMAT_ABGR:{
      const uint8_t *s = f->data[0] + (size_t) f->stride[0] * y;
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = s[4 * x + 0];
        d[4 * x + 1] = s[4 * x + 3];
        d[4 * x + 2] = s[4 * x + 2];
        d[4 * x + 3] = s[4 * x + 1];
      }
      break;
    }
  }
}

/* video-format.c: pack a whole (A,c1,c2,c3) line back.
 * pack_planar_420 / pack_NV12: luma always; chroma ONLY on chroma lines
 * (IS_CHROMA_LINE_420: progressive => even y) and ONLY from the even-x pixel
 * of each pair (video_orc_pack_I420/NV12: select0wb); an odd trailing pixel
 * writes its own chroma. No averaging anywhere. */
static void
pack_line (TbRefFrame *f, int y, const uint8_t *s, int width)
{
  int i, x;
  switch (f->format) {
    case TBREF_FORMAT_I420:
    case TBREF_FORMAT_YV12:{
      int pu = f->format == TBREF_FORMAT_I420 ? 1 : 2;
      int pv = f->format == TBREF_FORMAT_I420 ? 2 : 1;
      uint8_t *dy = f->data[0] + (size_t) f->stride[0] * y;
      uint8_t *du = f->data[pu] + (size_t) f->stride[pu] * (y >> 1);
      uint8_t *dv = f->data[pv] + (size_t) f->stride[pv] * (y >> 1);
      if (!(y & 1)) {
        for (i = 0; i < width / 2; i++) {
          dy[i * 2 + 0] = s[i * 8 + 1];
          dy[i * 2 + 1] = s[i * 8 + 5];
          du[i] = s[i * 8 + 2];
          dv[i] = s[i * 8 + 3];
        }
        if (width & 1) {
          i = width - 1;
          dy[i] = s[i * 4 + 1];
          du[i >> 1] = s[i * 4 + 2];
          dv[i >> 1] = s[i * 4 + 3];
        }
      } else {
        for (x = 0; x < width; x++)
          dy[x] = s[4 * x + 1];
      }
      break;
    }
    case TBREF_FORMAT_NV12:
    case TBREF_FORMAT_NV21:{
      int ou = f->format == TBREF_FORMAT_NV12 ? 0 : 1;
      uint8_t *dy = f->data[0] + (size_t) f->stride[0] * y;
      uint8_t *duv = f->data[1] + (size_t) f->stride[1] * (y >> 1);
      if (!(y & 1)) {
        for (i = 0; i < width / 2; i++) {
          dy[i * 2 + 0] = s[i * 8 + 1];
          dy[i * 2 + 1] = s[i * 8 + 5];
          duv[i * 2 + ou] = s[i * 8 + 2];
          duv[i * 2 + (1 - ou)] = s[i * 8 + 3];
        }
        if (width & 1) {
          i = width - 1;
          dy[i] = s[i * 4 + 1];
          duv[i + ou] = s[i * 4 + 2];
          duv[i + (1 - ou)] = s[i * 4 + 3];
        }
      } else {
        for (x = 0; x < width; x++)
          dy[x] = s[4 * x + 1];
      }
      break;
    }
    case TBREF_FORMAT_NV16:
    case TBREF_FORMAT_NV61:{       /* pack_NV16 / pack_NV61: every line, UV (VU) from the even pixel */
      const int ou = f->format == TBREF_FORMAT_NV61 ? 1 : 0;
      uint8_t *dy = f->data[0] + (size_t) f->stride[0] * y;
      uint8_t *duv = f->data[1] + (size_t) f->stride[1] * y;
      for (i = 0; i < width / 2; i++) {
        dy[i * 2 + 0] = s[i * 8 + 1];
        dy[i * 2 + 1] = s[i * 8 + 5];
        duv[i * 2 + ou] = s[i * 8 + 2];
        duv[i * 2 + (1 - ou)] = s[i * 8 + 3];
      }
      if (width & 1) {
        i = width - 1;
        dy[i] = s[i * 4 + 1];
        duv[i + ou] = s[i * 4 + 2];
        duv[i + (1 - ou)] = s[i * 4 + 3];
      }
      break;
    }
    case TBREF_FORMAT_v308:
    case TBREF_FORMAT_IYU2:{       /* pack_v308 / pack_IYU2: the alpha is dropped */
      uint8_t *d = f->data[0] + (size_t) f->stride[0] * y;
      const int oy = f->format == TBREF_FORMAT_v308 ? 0 : 1, ou = f->format == TBREF_FORMAT_v308 ? 1 : 0;
      for (x = 0; x < width; x++) {
        d[3 * x + oy] = s[4 * x + 1];
        d[3 * x + ou] = s[4 * x + 2];
        d[3 * x + 2] = s[4 * x + 3];
      }
      break;
    }
    case TBREF_FORMAT_NV24:{
      uint8_t *dy = f->data[0] + (size_t) f->stride[0] * y;
      uint8_t *duv = f->data[1] + (size_t) f->stride[1] * y;
      for (x = 0; x < width; x++) {
        dy[x] = s[4 * x + 1];
        duv[2 * x + 0] = s[4 * x + 2];
        duv[2 * x + 1] = s[4 * x + 3];
      }
      break;
    }
    case TBREF_FORMAT_Y42B:{       /* pack_Y42B: every line, chroma from the even pixel */
      uint8_t *dy = f->data[0] + (size_t) f->stride[0] * y;
      uint8_t *du = f->data[1] + (size_t) f->stride[1] * y;
      uint8_t *dv = f->data[2] + (size_t) f->stride[2] * y;
      for (i = 0; i < width / 2; i++) {
        dy[i * 2 + 0] = s[i * 8 + 1];
        dy[i * 2 + 1] = s[i * 8 + 5];
        du[i] = s[i * 8 + 2];
        dv[i] = s[i * 8 + 3];
      }
      if (width & 1) {
        i = width - 1;
        dy[i] = s[i * 4 + 1];
        du[i >> 1] = s[i * 4 + 2];
        dv[i >> 1] = s[i * 4 + 3];
      }
      break;
    }
    case TBREF_FORMAT_Y444:{
      uint8_t *dy = f->data[0] + (size_t) f->stride[0] * y;
      uint8_t *du = f->data[1] + (size_t) f->stride[1] * y;
      uint8_t *dv = f->data[2] + (size_t) f->stride[2] * y;
      for (x = 0; x < width; x++) {
        dy[x] = s[4 * x + 1];
        du[x] = s[4 * x + 2];
        dv[x] = s[4 * x + 3];
      }
      break;
    }
    case TBREF_FORMAT_YUY2:
    case TBREF_FORMAT_UYVY:
    case TBREF_FORMAT_YVYU:
    case TBREF_FORMAT_VYUY:{       /* pack_YUY2 / _UYVY / _YVYU / _VYUY: chroma from the even pixel; an
                                    * odd last pixel writes its Y, U and V, not the second luma */
      uint8_t *d = f->data[0] + (size_t) f->stride[0] * y;
      int oy, ou, ov;
      packed_422_order (f->format, &oy, &ou, &ov);
      for (i = 0; i < width / 2; i++) {
        d[i * 4 + oy] = s[i * 8 + 1];
        d[i * 4 + oy + 2] = s[i * 8 + 5];
        d[i * 4 + ou] = s[i * 8 + 2];
        d[i * 4 + ov] = s[i * 8 + 3];
      }
      if (width & 1) {
        i = width - 1;
        d[i * 2 + oy] = s[i * 4 + 1];
        d[i * 2 + ou] = s[i * 4 + 2];
        d[i * 2 + ov] = s[i * 4 + 3];
      }
      break;
    }
    case TBREF_FORMAT_GRAY8:{
      uint8_t *dy = f->data[0] + (size_t) f->stride[0] * y;
      for (x = 0; x < width; x++)
        dy[x] = s[4 * x + 1];
      break;
    }
    case TBREF_FORMAT_RGB:
    case TBREF_FORMAT_BGR:{        /* pack_RGB / pack_BGR: the alpha is dropped */
      uint8_t *d = f->data[0] + (size_t) f->stride[0] * y;
      const int orr = f->format == TBREF_FORMAT_RGB ? 0 : 2;
      for (x = 0; x < width; x++) {
        d[3 * x + orr] = s[4 * x + 1];
        d[3 * x + 1] = s[4 * x + 2];
        d[3 * x + (2 - orr)] = s[4 * x + 3];
      }
      break;
    }
    case TBREF_FORMAT_AYUV:
    case TBREF_FORMAT_ARGB:
      memcpy (f->data[0] + (size_t) f->stride[0] * y, s, (size_t) width * 4);
      break;
    case TBREF_FORMAT_RGBA:{
      uint8_t *d = f->data[0] + (size_t) f->stride[0] * y;
      for (x = 0; x < width; x++) {
        d[4 * x + 3] = s[4 * x + 0];
        d[4 * x + 0] = s[4 * x + 1];
        d[4 * x + 1] = s[4 * x + 2];
        d[4 * x + 2] = s[4 * x + 3];
      }
      break;
    }
    case TBREF_FORMAT_BGRA:{
      uint8_t *d = f->data[0] + (size_t) f->stride[0] * y;
      for (x = 0; x < width; x++) {
        d[4 * x + 3] = s[4 * x + 0];
        d[4 * x + 2] = s[4 * x + 1];
        d[4 * x + 1] = s[4 * x + 2];
        d[4 * x + 0] = s[4 * x + 3];
      }
      break;
    }
    case TBREF_FORMAT_ABGR:{
      uint8_t *d = f->data[0] + (size_t) f->stride[0] * y;
      for (x = 0; x < width; x++) {
        d[4 * x + 0] = s[4 * x + 0];
        d[4 * x + 3] = s[4 * x + 1];
        d[4 * x + 2] = s[4 * x + 2];
        d[4 * x + 1] = s[4 * x + 3];
      }
      break;
    }
  }
}

/* unpack_BGRA for the overlay rectangle: bytes (B,G,R,A) -> (A,R,G,B). */
static void
unpack_src_bgra (const TbRefRectangle *r, int xoff, int yoff, uint8_t *d,
    int width)
{
  const uint8_t *s = r->pixels + (size_t) r->stride * yoff + 4 * (size_t) xoff;
  int x;
  for (x = 0; x < width; x++) {
    d[4 * x + 0] = s[4 * x + 3];
    d[4 * x + 1] = s[4 * x + 2];
    d[4 * x + 2] = s[4 * x + 1];
    d[4 * x + 3] = s[4 * x + 0];
  }
}

/* ---------------------------------------------------------------------- */
/* video-blend.c colour matrices: BT.709, 8-bit, limited range            */

void
tbref_matrix_prea_rgb_to_yuv (uint8_t *t, uint32_t width)
{
  uint32_t i;
  int a, r, g, b, y, u, v;
  for (i = 0; i < width; i++) {
    a = t[i * 4 + 0];
    r = t[i * 4 + 1];
    g = t[i * 4 + 2];
    b = t[i * 4 + 3];
    if (a) {
      r = (r * 255 + a / 2) / a;
      g = (g * 255 + a / 2) / a;
      b = (b * 255 + a / 2) / a;
    }
    y = (47 * r + 157 * g + 16 * b + 4096) >> 8;
    u = (-26 * r - 87 * g + 112 * b + 32768) >> 8;
    v = (112 * r - 102 * g - 10 * b + 32768) >> 8;
    t[i * 4 + 1] = TB_CLAMP (y, 0, 255);
    t[i * 4 + 2] = TB_CLAMP (u, 0, 255);
    t[i * 4 + 3] = TB_CLAMP (v, 0, 255);
  }
}

void
tbref_matrix_rgb_to_yuv (uint8_t *t, uint32_t width)
{
  uint32_t i;
  int r, g, b, y, u, v;
  for (i = 0; i < width; i++) {
    r = t[i * 4 + 1];
    g = t[i * 4 + 2];
    b = t[i * 4 + 3];
    y = (47 * r + 157 * g + 16 * b + 4096) >> 8;
    u = (-26 * r - 87 * g + 112 * b + 32768) >> 8;
    v = (112 * r - 102 * g - 10 * b + 32768) >> 8;
    t[i * 4 + 1] = TB_CLAMP (y, 0, 255);
    t[i * 4 + 2] = TB_CLAMP (u, 0, 255);
    t[i * 4 + 3] = TB_CLAMP (v, 0, 255);
  }
}

void
tbref_matrix_yuv_to_rgb (uint8_t *t, uint32_t width)
{
  uint32_t i;
  int y, u, v, r, g, b;
  for (i = 0; i < width; i++) {
    y = t[i * 4 + 1];
    u = t[i * 4 + 2];
    v = t[i * 4 + 3];
    r = (298 * y + 459 * v - 63514) >> 8;
    g = (298 * y - 55 * u - 136 * v + 19681) >> 8;
    b = (298 * y + 541 * u - 73988) >> 8;
    t[i * 4 + 1] = TB_CLAMP (r, 0, 255);
    t[i * 4 + 2] = TB_CLAMP (g, 0, 255);
    t[i * 4 + 3] = TB_CLAMP (b, 0, 255);
  }
}

/* ---------------------------------------------------------------------- */
/* video-blend.c: A OVER B, 8 bit. alphaG = global alpha (premultiplied    */
/* source only), alphaA/colorA = source, alphaB/colorB = dest, alphaD =    */
/* blended alpha (non-premultiplied dest only). Integer, truncating.       */

#define OVER00(aG, aA, cA, aB, cB, aD) \
  (((cA) * (aA) + (cB) * (aB) * (255 - (aA)) / 255) / (aD))
#define OVER10(aG, aA, cA, aB, cB, aD) \
  (((cA) * (aG) + (cB) * (aB) * (255 - (aA)) / 255) / (aD))
#define OVER01(aG, aA, cA, aB, cB, aD) \
  (((cA) * (aA) + (cB) * (255 - (aA))) / 255)
#define OVER11(aG, aA, cA, aB, cB, aD) \
  (((cA) * (aG) + (cB) * (255 - (aA))) / 255)

#define BLENDC(op, aG, aA, cA, aB, cB, aD) do {                \
    unsigned int c_ = op ((unsigned int) (aG), (aA),           \
        (unsigned int) (cA), (aB), (unsigned int) (cB),        \
        (unsigned int) (aD));                                  \
    (cB) = (uint8_t) TB_MIN (c_, 255u);                        \
  } while (0)

#define BLENDLOOP(op) do {                                                  \
    for (j = 0; j < src_width * 4; j += 4) {                                \
      unsigned int asrc, adst;                                              \
      int final_alpha;                                                      \
      asrc = ((unsigned int) tmpsrc[j]) * (unsigned int) ga / 255u;         \
      if (!asrc)                                                            \
        continue;                                                           \
      adst = dl[j];                                                         \
      final_alpha = (int) (asrc + adst * (255u - asrc) / 255u);             \
      dl[j] = (uint8_t) final_alpha;                                        \
      if (final_alpha == 0)                                                 \
        final_alpha = 1;                                                    \
      BLENDC (op, ga, asrc, tmpsrc[j + 1], adst, dl[j + 1], final_alpha);   \
      BLENDC (op, ga, asrc, tmpsrc[j + 2], adst, dl[j + 2], final_alpha);   \
      BLENDC (op, ga, asrc, tmpsrc[j + 3], adst, dl[j + 3], final_alpha);   \
    }                                                                       \
  } while (0)

int
tbref_video_blend (TbRefFrame *dest, const TbRefRectangle *src)
{
  int i, j, ga, src_width, src_height, dest_width, dest_height;
  int src_xoff = 0, src_yoff = 0, x = src->x, y = src->y;
  int src_premul, dest_premul, mode_matrix = 0;
  uint8_t *tmpdest, *tmpsrc;

  if (!dest || !src || !src->pixels)
    return 0;
  if (dest->format < TBREF_FORMAT_I420 || dest->format > TBREF_FORMAT_BGR ||
      (dest->format > TBREF_FORMAT_ABGR && dest->format < TBREF_FORMAT_Y42B))
    return 0;

  ga = (int) (255.0 * src->global_alpha);
  dest_premul = (dest->flags & TBREF_FLAG_PREMULTIPLIED_ALPHA) != 0;
  src_premul = (src->flags & TBREF_FLAG_PREMULTIPLIED_ALPHA) != 0;

  src_width = src->width;
  src_height = src->height;
  dest_width = dest->width;
  dest_height = dest->height;

  /* completely outside the video: nothing to do, still TRUE */
  if (x + src_width <= 0 || y + src_height <= 0 || x >= dest_width ||
      y >= dest_height)
    return 1;

  /* overlay is RGB; a YUV dest needs the matrix. A premultiplied source is
   * un-premultiplied by the matrix and treated as straight from then on. */
  if (is_yuv (dest->format)) {
    if (src_premul) {
      mode_matrix = 2;
      src_premul = 0;
    } else {
      mode_matrix = 1;
    }
  }

  if (x < 0) {
    src_xoff = -x;
    src_width -= src_xoff;
    x = 0;
  }
  if (y < 0) {
    src_yoff = -y;
    src_height -= src_yoff;
    y = 0;
  }
  if (x + src_width > dest_width)
    src_width = dest_width - x;
  if (y + src_height > dest_height)
    src_height = dest_height - y;

  tmpdest = (uint8_t *) malloc ((size_t) (dest_width + 8) * 4);
  tmpsrc = (uint8_t *) malloc ((size_t) (src_width + 8) * 4);
  if (!tmpdest || !tmpsrc) {
    free (tmpdest);
    free (tmpsrc);
    return 0;
  }

  for (i = y; i < y + src_height; i++, src_yoff++) {
    uint8_t *dl;
    unpack_line (dest, i, tmpdest, dest_width);
    unpack_src_bgra (src, src_xoff, src_yoff, tmpsrc, src_width);
    dl = tmpdest + 4 * x;

    if (mode_matrix == 2)
      tbref_matrix_prea_rgb_to_yuv (tmpsrc, (uint32_t) src_width);
    else if (mode_matrix == 1)
      tbref_matrix_rgb_to_yuv (tmpsrc, (uint32_t) src_width);

    if (src_premul && dest_premul)
      BLENDLOOP (OVER11);
    else if (!src_premul && dest_premul)
      BLENDLOOP (OVER01);
    else if (src_premul && !dest_premul)
      BLENDLOOP (OVER10);
    else
      BLENDLOOP (OVER00);

    pack_line (dest, i, tmpdest, dest_width);
  }

  free (tmpdest);
  free (tmpsrc);
  return 1;
}

/* ---------------------------------------------------------------------- */
/* rectangle scaling: gst_video_blend_scale_linear_RGBA (video-blend.c) with
 * the two ORC programs it calls (video-orc.orc) [UPSTREAM-RECALL]            */

/* video_orc_resample_bilinear_u32 (d, s, p1, p2, n) = ORC `ldreslinl`: for every
 * destination pixel tmp = p1 + i*p2; the two source pixels tmp>>16 and (tmp>>16)+1 are
 * mixed per byte with the 8-bit fraction (tmp>>8)&255: (a*(256-f) + b*f) >> 8. The second
 * pixel is read even when f == 0. */
static void
orc_resample_bilinear_u32 (uint8_t *d, const uint8_t *s, int p1, int p2, int n)
{
  int i, k;
  for (i = 0; i < n; i++) {
    const int tmp = p1 + i * p2;
    const uint8_t *a = s + 4 * (tmp >> 16), *b = a + 4;
    const int f = (tmp >> 8) & 0xff;
    for (k = 0; k < 4; k++)
      d[4 * i + k] = (uint8_t) ((a[k] * (256 - f) + b[k] * f) >> 8);
  }
}

/* video_orc_merge_linear_u8 (d, s1, s2, p1, n): convubw both, subw, mullw by p1, addw 128,
 * convhwb (high byte of the 16-bit word), addb s1 -- all wrapping, which for 8-bit inputs
 * equals s1 + floor (((s2 - s1)*p1 + 128) / 256). */
static void
orc_merge_linear_u8 (uint8_t *d, const uint8_t *s1, const uint8_t *s2, int p1, int n)
{
  int i;
  for (i = 0; i < n; i++) {
    const uint16_t t = (uint16_t) ((uint16_t) ((uint16_t) s2[i] - (uint16_t) s1[i]) * (uint16_t) p1 + 128u);
    d[i] = (uint8_t) (s1[i] + (uint8_t) (t >> 8));
  }
}

/* gst_video_blend_scale_linear_RGBA: horizontally resampled source lines are kept in a
 * two-line cache (LINE (n) = slot n & 1) and merged vertically. The cache bookkeeping (y1)
 * is restated as upstream has it, including what it does when the row index jumps. A
 * source that is 1 pixel wide or high makes upstream read outside the image (increment -1);
 * the callers here reject that case. dst is tightly packed (stride dest_width*4). */
void
tbref_scale_linear_rgba (const uint8_t *src_pixels, int32_t src_width, int32_t src_height,
    int32_t src_stride, int32_t dest_width, int32_t dest_height, uint8_t *dest_pixels)
{
  int acc = 0, y_increment, x_increment, y1 = 0, i, j, x;
  const int dest_size = dest_width * 4, dest_stride = dest_width * 4;
  uint8_t *tmpbuf = (uint8_t *) malloc ((size_t) dest_width * 8 * 4);
#define LINE(n) (tmpbuf + (size_t) dest_size * ((n) & 1))

  y_increment = dest_height == 1 ? 0 : ((src_height - 1) << 16) / (dest_height - 1) - 1;
  x_increment = dest_width == 1 ? 0 : ((src_width - 1) << 16) / (dest_width - 1) - 1;

  orc_resample_bilinear_u32 (LINE (0), src_pixels, 0, x_increment, dest_width);
  for (i = 0; i < dest_height; i++) {
    j = acc >> 16;
    x = acc & 0xffff;
    if (x == 0) {
      memcpy (dest_pixels + (size_t) i * dest_stride, LINE (j), (size_t) dest_size);
    } else {
      if (j > y1) {
        orc_resample_bilinear_u32 (LINE (j), src_pixels + (size_t) j * src_stride, 0, x_increment, dest_width);
        y1++;
      }
      if (j >= y1) {
        orc_resample_bilinear_u32 (LINE (j + 1), src_pixels + (size_t) (j + 1) * src_stride, 0, x_increment,
            dest_width);
        y1++;
      }
      orc_merge_linear_u8 (dest_pixels + (size_t) i * dest_stride, LINE (j), LINE (j + 1), x >> 8, dest_width * 4);
    }
    acc += y_increment;
  }
#undef LINE
  free (tmpbuf);
}

/* gst_video_overlay_composition_blend: rectangles in list order; one whose render size
 * differs from its pixel size (gst_video_overlay_rectangle_needs_scaling) is scaled first. */
int
tbref_composition_blend (TbRefFrame *dest, const TbRefRectangle *rects,
    uint32_t n_rects)
{
  uint32_t n;
  int ret = 1;
  for (n = 0; n < n_rects; n++) {
    const TbRefRectangle *r = &rects[n];
    const int32_t rw = r->render_width > 0 ? r->render_width : r->width;
    const int32_t rh = r->render_height > 0 ? r->render_height : r->height;
    if (rw != r->width || rh != r->height) {
      TbRefRectangle scaled = *r;
      uint8_t *px;
      if (r->width < 2 || r->height < 2)
        return 0;                   /* upstream reads out of bounds here */
      px = (uint8_t *) malloc ((size_t) rw * 4 * (size_t) rh);
      tbref_scale_linear_rgba (r->pixels, r->width, r->height, r->stride, rw, rh, px);
      scaled.pixels = px;
      scaled.width = rw;
      scaled.height = rh;
      scaled.stride = rw * 4;
      scaled.render_width = scaled.render_height = 0;
      ret = tbref_video_blend (dest, &scaled);
      free (px);
    } else {
      ret = tbref_video_blend (dest, r);
    }
  }
  return ret;
}

/* ---------------------------------------------------------------------- */
/* outline blur: gstttmlblur.c + pixman's convolution filter               */

int32_t
tbref_gaussian_kernel (int32_t radius, double sigma, int32_t *taps)
{
  /* gst_ttml_blur_create_gaussian_kernel, /root/reference/plugins/ttml/gstttmlblur.c:28-67 */
  const double scale2 = 2.0 * sigma * sigma;
  const double scale1 = 1.0 / (3.14159265358979323846 * scale2);
  const int size = 2 * radius + 1;
  const int n = size * size;
  double *tmp = (double *) malloc (sizeof (double) * (size_t) n);
  double sum = 0;
  int x, y, i = 0;
  for (x = -radius; x <= radius; ++x)
    for (y = -radius; y <= radius; ++y, ++i) {
      const double u = x * x, v = y * y;
      tmp[i] = scale1 * exp (-(u + v) / scale2);
      sum += tmp[i];
    }
  for (i = 0; i < n; ++i)
    taps[i] = (int32_t) ((tmp[i] / sum) * 65536.0);      /* pixman_double_to_fixed */
  free (tmp);
  return n;
}

void
tbref_blur_argb32 (const uint8_t *src, int32_t width, int32_t height, int32_t stride,
    int32_t radius, double sigma, uint8_t *dst, int32_t dst_stride)
{
  const int size = 2 * radius + 1;
  int32_t *taps = (int32_t *) malloc (sizeof (int32_t) * (size_t) size * size);
  tbref_gaussian_kernel (radius, sigma, taps);
  tbref_convolve_argb32 (src, width, height, stride, size, taps, dst, dst_stride);
  free (taps);
}

void
tbref_convolve_argb32 (const uint8_t *src, int32_t width, int32_t height, int32_t stride,
    int32_t size, const int32_t *taps, uint8_t *dst, int32_t dst_stride)
{
  /* bits_image_fetch_pixel_convolution: window [x-r, x+r] x [y-r, y+r], pixels outside the
   * image are transparent black (REPEAT_NONE), 16.16 taps, (sum + 0x8000) >> 16, CLIP 0..255 */
  const int radius = (size - 1) / 2;
  int x, y, i, j, c;
  for (y = 0; y < height; y++)
    for (x = 0; x < width; x++) {
      int tot[4] = { 0, 0, 0, 0 };
      const int32_t *t = taps;
      for (i = y - radius; i <= y + radius; i++)
        for (j = x - radius; j <= x + radius; j++, t++) {
          if (!*t || i < 0 || i >= height || j < 0 || j >= width)
            continue;
          for (c = 0; c < 4; c++)
            tot[c] += (int) src[(size_t) i * stride + 4 * (size_t) j + c] * *t;
        }
      for (c = 0; c < 4; c++) {
        int v = (tot[c] + 0x8000) >> 16;
        dst[(size_t) y * dst_stride + 4 * (size_t) x + c] = (uint8_t) TB_CLAMP (v, 0, 255);
      }
    }
}

/* ---------------------------------------------------------------------- */
/* region composition: Cairo colour conversion + pixman 8-bit OVER / IN    */

static uint32_t
mul_un8 (uint32_t a, uint32_t b)
{
  /* pixman-combine32.h MUL_UN8 */
  uint32_t t = a * b + 0x80;
  return ((t >> 8) + t) >> 8;
}

static uint32_t
over_px (uint32_t s, uint32_t d)
{
  /* combine_over_u: d = s + d * (255 - alpha (s)) / 255, saturating per channel */
  uint32_t ia = 255 - (s >> 24), out = 0;
  int k;
  for (k = 0; k < 4; k++) {
    uint32_t v = ((s >> (8 * k)) & 0xff) + mul_un8 ((d >> (8 * k)) & 0xff, ia);
    out |= TB_MIN (v, 255u) << (8 * k);
  }
  return out;
}

static uint32_t
in_px (uint32_t s, uint32_t m)
{
  uint32_t out = 0;
  int k;
  for (k = 0; k < 4; k++)
    out |= mul_un8 ((s >> (8 * k)) & 0xff, m) << (8 * k);
  return out;
}

static uint32_t
solid_px (uint32_t c)
{
  /* GET_CAIRO_COMP (gstttmlrender.c:1178) -> cairo_set_source_rgba ->
   * _cairo_color_compute_shorts (d * 65535.0 + 0.5, colour premultiplied) ->
   * pixman solid fill (short >> 8) */
  double r = ((c >> 24) & 255) / 255.0, g = ((c >> 16) & 255) / 255.0;
  double b = ((c >> 8) & 255) / 255.0, a = (c & 255) / 255.0;
  uint32_t as = (uint16_t) (a * 65535.0 + 0.5), rs = (uint16_t) (r * a * 65535.0 + 0.5);
  uint32_t gs = (uint16_t) (g * a * 65535.0 + 0.5), bs = (uint16_t) (b * a * 65535.0 + 0.5);
  return ((as >> 8) << 24) | ((rs >> 8) << 16) | ((gs >> 8) << 8) | (bs >> 8);
}

static uint32_t
load32 (const uint8_t *p)
{
  return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24);
}

void
tbref_compose_regions (const TbRefRegion *regions, uint32_t n, int32_t W, int32_t H,
    uint8_t *out, int32_t out_stride)
{
  uint32_t i;
  int x, y;
  for (y = 0; y < H; y++)
    memset (out + (size_t) y * out_stride, 0, (size_t) W * 4);     /* CAIRO_OPERATOR_CLEAR */
  for (i = 0; i < n; i++) {
    const TbRefRegion *r = &regions[i];
    uint32_t bg = r->background_color ? solid_px (r->background_color) : 0;
    uint32_t m8 = r->opacity < 1.0 ? ((uint32_t) (uint16_t) (r->opacity * 65535.0 + 0.5)) >> 8 : 255;
    if (!bg && !r->layer)
      continue;
    for (y = r->y; y < r->y + r->h; y++) {
      if (y < 0 || y >= H)
        continue;
      for (x = r->x; x < r->x + r->w; x++) {
        uint8_t *p;
        uint32_t d, layer = 0;
        if (x < 0 || x >= W)
          continue;
        p = out + (size_t) y * out_stride + 4 * (size_t) x;
        d = load32 (p);
        if (r->layer)
          layer = load32 (r->layer + (size_t) (y - r->y) * r->layer_stride + 4 * (size_t) (x - r->x));
        if (m8 == 255) {
          /* drawn straight onto the frame-sized surface */
          if (bg)
            d = over_px (bg, d);
          if (r->layer)
            d = over_px (layer, d);
        } else {
          /* group surface, then cairo_paint_with_alpha (opacity) */
          uint32_t g = bg;
          if (r->layer)
            g = over_px (layer, g);
          d = over_px (in_px (g, m8), d);
        }
        p[0] = d & 0xff;
        p[1] = (d >> 8) & 0xff;
        p[2] = (d >> 16) & 0xff;
        p[3] = d >> 24;
      }
    }
  }
}

/* ---------------------------------------------------------------------- */
/* CPU baseline driver (bench.py cpu_baseline / --impl reference)         */

typedef struct {
  TbRefFrame *frames;
  uint32_t n_frames, first, step;
  const TbRefRectangle *rects;
  uint32_t n_rects;
} TbRefWork;

static void *
blend_worker (void *arg)
{
  TbRefWork *w = (TbRefWork *) arg;
  uint32_t i;
  for (i = w->first; i < w->n_frames; i += w->step)
    tbref_composition_blend (&w->frames[i], w->rects, w->n_rects);
  return NULL;
}

double
tbref_blend_many (TbRefFrame *frames, uint32_t n_frames,
    const TbRefRectangle *rects, uint32_t n_rects, uint32_t n_threads)
{
  struct timespec t0, t1;
  pthread_t *tids;
  TbRefWork *work;
  uint32_t t;

  if (n_threads < 1)
    n_threads = 1;
  tids = (pthread_t *) calloc (n_threads, sizeof (pthread_t));
  work = (TbRefWork *) calloc (n_threads, sizeof (TbRefWork));
  clock_gettime (CLOCK_MONOTONIC, &t0);
  for (t = 0; t < n_threads; t++) {
    work[t].frames = frames;
    work[t].n_frames = n_frames;
    work[t].first = t;
    work[t].step = n_threads;
    work[t].rects = rects;
    work[t].n_rects = n_rects;
    if (t + 1 < n_threads)
      pthread_create (&tids[t], NULL, blend_worker, &work[t]);
  }
  blend_worker (&work[n_threads - 1]);
  for (t = 0; t + 1 < n_threads; t++)
    pthread_join (tids[t], NULL);
  clock_gettime (CLOCK_MONOTONIC, &t1);
  free (tids);
  free (work);
  return (double) (t1.tv_sec - t0.tv_sec) +
      1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}
