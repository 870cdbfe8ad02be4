/*
 * selftest.c -- runs the oracle over awkward geometry under ASan + UBSan
 * (tests/test_oracle.py builds it with -fsanitize=address,undefined). Test
 * infrastructure only. Exit code 0 = no sanitizer report and the trivial
 * invariants hold.
 */
#include "ttmlblend_ref.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t state = 0x74746d6c;
static uint8_t
rnd (void)
{
  uint64_t z = (state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint8_t) ((z ^ (z >> 31)) & 0xff);
}

int
main (void)
{
  static const int sizes[][2] = { {1, 1}, {2, 2}, {3, 5}, {17, 9}, {64, 48}, {63, 47}, {129, 3} };
  int fmt, s, k, fails = 0;
  for (fmt = TBREF_FORMAT_I420; fmt <= TBREF_FORMAT_BGR; fmt++) {
    if (fmt > TBREF_FORMAT_ABGR && fmt < TBREF_FORMAT_Y42B)
      continue;
    for (s = 0; s < (int) (sizeof sizes / sizeof sizes[0]); s++) {
      const int w = sizes[s][0], h = sizes[s][1];
      TbRefFrame f;
      uint8_t *planes[3] = { NULL, NULL, NULL }, *copy0;
      size_t bytes0 = 0;
      int p, n = tbref_n_planes (fmt);
      memset (&f, 0, sizeof f);
      f.format = fmt;
      f.width = w;
      f.height = h;
      for (p = 0; p < n; p++) {
        const int rb = tbref_plane_row_bytes (fmt, p, w), rows = tbref_plane_rows (fmt, p, h);
        size_t i;
        planes[p] = (uint8_t *) malloc ((size_t) rb * rows);    /* exact size: overruns trip ASan */
        for (i = 0; i < (size_t) rb * rows; i++)
          planes[p][i] = rnd ();
        f.data[p] = planes[p];
        f.stride[p] = rb;
        if (p == 0)
          bytes0 = (size_t) rb * rows;
      }
      copy0 = (uint8_t *) malloc (bytes0);
      memcpy (copy0, planes[0], bytes0);
      for (k = 0; k < 12; k++) {
        const int rw = 1 + rnd () % (w + 9), rh = 1 + rnd () % (h + 9);
        TbRefRectangle r;
        uint8_t *px = (uint8_t *) malloc ((size_t) rw * rh * 4);
        size_t i;
        for (i = 0; i < (size_t) rw * rh * 4; i++)
          px[i] = rnd ();
        memset (&r, 0, sizeof r);
        r.pixels = px;
        r.width = rw;
        r.height = rh;
        r.stride = rw * 4;
        r.x = (int) (rnd () % (w + rw + 2)) - rw - 1;
        r.y = (int) (rnd () % (h + rh + 2)) - rh - 1;
        r.global_alpha = (k % 3 == 0) ? 0.5f : 1.0f;
        r.flags = (k & 1) ? TBREF_FLAG_PREMULTIPLIED_ALPHA : 0;
        f.flags = (k % 5 == 0) ? TBREF_FLAG_PREMULTIPLIED_ALPHA : 0;
        if (!tbref_video_blend (&f, &r))
          fails++;
        /* a rectangle entirely outside must not touch anything */
        memcpy (copy0, planes[0], bytes0);
        r.x = w;
        if (!tbref_video_blend (&f, &r) || memcmp (copy0, planes[0], bytes0))
          fails++;
        free (px);
      }
      free (copy0);
      for (p = 0; p < n; p++)
        free (planes[p]);
    }
  }
  {
    int32_t taps[49];
    uint8_t img[5 * 7 * 4], out[5 * 7 * 4];
    size_t i;
    for (i = 0; i < sizeof img; i++)
      img[i] = rnd ();
    tbref_gaussian_kernel (3, 1.5, taps);
    tbref_blur_argb32 (img, 7, 5, 28, 3, 1.5, out, 28);
  }
  {
    /* rectangle scaling: exact-size buffers in and out, sizes that make the line cache jump,
     * and a composition that scales before it blends */
    static const int cases[][4] = { {2, 2, 1, 1}, {2, 2, 3, 3}, {7, 5, 20, 3}, {33, 40, 5, 90}, {64, 9, 1, 30},
      {3, 1200, 4, 37}, {5, 37, 2, 1200} };
    size_t c, i;
    for (c = 0; c < sizeof cases / sizeof cases[0]; c++) {
      const int sw = cases[c][0], sh = cases[c][1], dw = cases[c][2], dh = cases[c][3];
      uint8_t *src = (uint8_t *) malloc ((size_t) sw * sh * 4), *dst = (uint8_t *) malloc ((size_t) dw * dh * 4);
      uint8_t *y = (uint8_t *) malloc (64 * 48), *uv = (uint8_t *) malloc (64 * 24);
      TbRefFrame f;
      TbRefRectangle r;
      for (i = 0; i < (size_t) sw * sh * 4; i++)
        src[i] = rnd ();
      tbref_scale_linear_rgba (src, sw, sh, sw * 4, dw, dh, dst);
      memset (&f, 0, sizeof f);
      f.format = TBREF_FORMAT_NV12;
      f.width = 64;
      f.height = 48;
      f.data[0] = y;
      f.data[1] = uv;
      f.stride[0] = f.stride[1] = 64;
      memset (y, 100, 64 * 48);
      memset (uv, 128, 64 * 24);
      memset (&r, 0, sizeof r);
      r.pixels = src;
      r.width = sw;
      r.height = sh;
      r.stride = sw * 4;
      r.x = -3;
      r.y = 5;
      r.global_alpha = 1.0f;
      r.flags = TBREF_FLAG_PREMULTIPLIED_ALPHA;
      r.render_width = dw;
      r.render_height = dh;
      if (!tbref_composition_blend (&f, &r, 1))
        fails++;
      free (src);
      free (dst);
      free (y);
      free (uv);
    }
  }
  printf ("oracle selftest: %d failure(s)\n", fails);
  return fails ? 1 : 0;
}
