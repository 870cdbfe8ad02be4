/*
 * shim.c -- the cairo / pixman calls /root/reference/plugins/ttml/gstttmlblur.c makes, reduced
 * to what that file needs, and two entry points for the tests. See README.md. Our code, not the
 * reference's: the reference source is compiled next to it from where it lies.
 */
#include <stdlib.h>
#include <string.h>

#include <pango/pangocairo.h>
#include <pixman.h>

#include "../ttmlblend_ref.h"

struct _cairo_surface {
  unsigned char *data;
  int width, height, stride;
  void *user_data;
  cairo_destroy_func_t destroy;
};

struct _pixman_image {
  uint32_t *bits;
  int width, height, stride;
  pixman_fixed_t *params;
  int n_params;
};

/* what the reference handed to pixman_image_set_filter last (for the tests) */
static pixman_fixed_t last_params[2 + 129 * 129];
static int last_n_params;

int cairo_image_surface_get_width (cairo_surface_t *s) { return s->width; }
int cairo_image_surface_get_height (cairo_surface_t *s) { return s->height; }
int cairo_image_surface_get_stride (cairo_surface_t *s) { return s->stride; }
unsigned char *cairo_image_surface_get_data (cairo_surface_t *s) { return s->data; }

cairo_surface_t *
cairo_image_surface_create_for_data (unsigned char *data, cairo_format_t format, int width, int height, int stride)
{
  cairo_surface_t *s = (cairo_surface_t *) calloc (1, sizeof (*s));
  (void) format;
  s->data = data;
  s->width = width;
  s->height = height;
  s->stride = stride;
  return s;
}

int
cairo_surface_set_user_data (cairo_surface_t *s, const cairo_user_data_key_t *key, void *user_data,
    cairo_destroy_func_t destroy)
{
  (void) key;
  s->user_data = user_data;
  s->destroy = destroy;
  return 0;
}

pixman_image_t *
pixman_image_create_bits (pixman_format_code_t format, int width, int height, uint32_t *bits, int stride)
{
  pixman_image_t *im = (pixman_image_t *) calloc (1, sizeof (*im));
  (void) format;
  im->bits = bits;
  im->width = width;
  im->height = height;
  im->stride = stride;
  return im;
}

pixman_bool_t
pixman_image_set_filter (pixman_image_t *im, pixman_filter_t filter, const pixman_fixed_t *params, int n)
{
  if (filter != PIXMAN_FILTER_CONVOLUTION || n < 2 || n > (int) (sizeof last_params / sizeof last_params[0]))
    return 0;
  free (im->params);
  im->params = (pixman_fixed_t *) malloc (sizeof (pixman_fixed_t) * (size_t) n);   /* pixman copies them too */
  memcpy (im->params, params, sizeof (pixman_fixed_t) * (size_t) n);
  im->n_params = n;
  memcpy (last_params, params, sizeof (pixman_fixed_t) * (size_t) n);
  last_n_params = n;
  return 1;
}

/* PIXMAN_OP_SRC of a convolution-filtered a8r8g8b8 source onto dest: the oracle's restatement
 * of pixman's convolution with the taps the reference set (this part is NOT the reference). */
void
pixman_image_composite (pixman_op_t op, pixman_image_t *src, pixman_image_t *mask, pixman_image_t *dest,
    int16_t src_x, int16_t src_y, int16_t mask_x, int16_t mask_y, int16_t dest_x, int16_t dest_y,
    uint16_t width, uint16_t height)
{
  (void) op; (void) mask; (void) src_x; (void) src_y; (void) mask_x; (void) mask_y; (void) dest_x; (void) dest_y;
  if (!src->params || src->n_params < 2)
    return;
  tbref_convolve_argb32 ((const uint8_t *) src->bits, width, height, src->stride,
      src->params[0] >> 16, src->params + 2, (uint8_t *) dest->bits, dest->stride);
}

pixman_bool_t
pixman_image_unref (pixman_image_t *im)
{
  free (im->params);
  free (im);
  return 1;
}

/* ---- entry points for the tests -------------------------------------- */

cairo_surface_t *gst_ttml_blur_image_surface (cairo_surface_t *surface, int radius, double sigma);

/* Runs the reference's gst_ttml_blur_image_surface on an ARGB32 image. */
void
ttmlref_blur_argb32 (const uint8_t *src, int32_t width, int32_t height, int32_t stride, int32_t radius,
    double sigma, uint8_t *dst, int32_t dst_stride)
{
  cairo_surface_t in, *out;
  int y;
  memset (&in, 0, sizeof in);
  in.data = (unsigned char *) src;
  in.width = width;
  in.height = height;
  in.stride = stride;
  out = gst_ttml_blur_image_surface (&in, radius, sigma);
  for (y = 0; y < height; y++)
    memcpy (dst + (size_t) y * dst_stride, out->data + (size_t) y * out->stride, (size_t) width * 4);
  if (out->destroy)
    out->destroy (out->user_data);
  free (out);
}

/* The filter parameters the reference gave pixman in the last call: 2 sizes + taps, 16.16. */
int32_t
ttmlref_last_filter_params (int32_t *out, int32_t max)
{
  int i;
  for (i = 0; i < last_n_params && i < max; i++)
    out[i] = last_params[i];
  return last_n_params;
}
