/*
 * fluc_videooverlay.c -- host side in C above the C ABI (see the header).
 * Object model after the reference's helpers: refcounted opaque structs and a
 * process-wide singleton guarded by a mutex, like the bwmeter singleton
 * (/root/reference/libs/fluc/flu-codec-sdk/fluc/bwmeter/fluc_bwmeter.c:17-44).
 * pthreads instead of GLib: GLib is not available to this build.
 */
#include "fluc_videooverlay.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

struct _FlucVideoOverlayRectangle {
  int refcount;
  uint8_t *pixels;
  int32_t width, height, stride;
  int32_t x, y;
  int32_t render_width, render_height;
  float global_alpha;
  uint32_t flags;
};

struct _FlucVideoOverlayComposition {
  int refcount;
  FlucVideoOverlayRectangle **rects;
  uint32_t n_rects;
  uint32_t stream;             /* overlay-cache key in the context */
  int uploaded;
};

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static FlucTtmlBlend *g_ctx = NULL;
static int g_device = -1;
static uint32_t g_next_stream = 0x40000000u;   /* keep clear of element stream ids */

static FlucTtmlBlend *
context_locked (void)
{
  if (!g_ctx) {
    if (fluc_ttmlblend_new (g_device, &g_ctx) != FLUC_TTMLBLEND_OK)
      g_ctx = NULL;
  }
  return g_ctx;
}

int
fluc_video_overlay_set_device (int device)
{
  int ret = 0;
  pthread_mutex_lock (&g_lock);
  if (g_ctx && device != g_device)
    ret = FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;   /* already bound */
  else
    g_device = device;
  pthread_mutex_unlock (&g_lock);
  return ret;
}

FlucTtmlBlend *
fluc_video_overlay_get_context (void)
{
  FlucTtmlBlend *c;
  pthread_mutex_lock (&g_lock);
  c = context_locked ();
  pthread_mutex_unlock (&g_lock);
  return c;
}

void
fluc_video_overlay_deinit (void)
{
  pthread_mutex_lock (&g_lock);
  if (g_ctx)
    fluc_ttmlblend_free (g_ctx);
  g_ctx = NULL;
  pthread_mutex_unlock (&g_lock);
}

FlucVideoOverlayRectangle *
fluc_video_overlay_rectangle_new_raw (const uint8_t *pixels, int32_t width, int32_t height,
    int32_t stride, int32_t render_x, int32_t render_y, uint32_t render_width, uint32_t render_height,
    uint32_t flags)
{
  FlucVideoOverlayRectangle *r;
  int32_t y;
  if (!pixels || width <= 0 || height <= 0 || stride < width * 4 || render_width > 32768 ||
      render_height > 32768)
    return NULL;
  r = (FlucVideoOverlayRectangle *) calloc (1, sizeof (*r));
  if (!r)
    return NULL;
  r->pixels = (uint8_t *) malloc ((size_t) width * 4 * (size_t) height);
  if (!r->pixels) {
    free (r);
    return NULL;
  }
  for (y = 0; y < height; y++)
    memcpy (r->pixels + (size_t) y * width * 4, pixels + (size_t) y * stride, (size_t) width * 4);
  r->refcount = 1;
  r->width = width;
  r->height = height;
  r->stride = width * 4;
  r->x = render_x;
  r->y = render_y;
  r->render_width = render_width ? (int32_t) render_width : width;
  r->render_height = render_height ? (int32_t) render_height : height;
  r->global_alpha = 1.0f;
  r->flags = flags;
  return r;
}

FlucVideoOverlayRectangle *
fluc_video_overlay_rectangle_ref (FlucVideoOverlayRectangle *rect)
{
  if (rect)
    __atomic_add_fetch (&rect->refcount, 1, __ATOMIC_SEQ_CST);
  return rect;
}

void
fluc_video_overlay_rectangle_unref (FlucVideoOverlayRectangle *rect)
{
  if (rect && __atomic_sub_fetch (&rect->refcount, 1, __ATOMIC_SEQ_CST) == 0) {
    free (rect->pixels);
    free (rect);
  }
}

void
fluc_video_overlay_rectangle_set_global_alpha (FlucVideoOverlayRectangle *rect, float global_alpha)
{
  if (rect && global_alpha >= 0.0f && global_alpha <= 1.0f)
    rect->global_alpha = global_alpha;
}

float
fluc_video_overlay_rectangle_get_global_alpha (FlucVideoOverlayRectangle *rect)
{
  return rect ? rect->global_alpha : 0.0f;
}

void
fluc_video_overlay_rectangle_set_render_rectangle (FlucVideoOverlayRectangle *rect, int32_t render_x,
    int32_t render_y, uint32_t render_width, uint32_t render_height)
{
  if (rect && render_width <= 32768 && render_height <= 32768) {
    rect->x = render_x;
    rect->y = render_y;
    rect->render_width = render_width ? (int32_t) render_width : rect->width;
    rect->render_height = render_height ? (int32_t) render_height : rect->height;
  }
}

FlucVideoOverlayComposition *
fluc_video_overlay_composition_new (FlucVideoOverlayRectangle *rect)
{
  FlucVideoOverlayComposition *c = (FlucVideoOverlayComposition *) calloc (1, sizeof (*c));
  if (!c)
    return NULL;
  c->refcount = 1;
  pthread_mutex_lock (&g_lock);
  c->stream = g_next_stream++;
  pthread_mutex_unlock (&g_lock);
  if (rect)
    fluc_video_overlay_composition_add_rectangle (c, rect);
  return c;
}

void
fluc_video_overlay_composition_add_rectangle (FlucVideoOverlayComposition *comp,
    FlucVideoOverlayRectangle *rect)
{
  FlucVideoOverlayRectangle **n;
  if (!comp || !rect || comp->uploaded || comp->n_rects >= FLUC_TTMLBLEND_MAX_RECTANGLES)
    return;
  n = (FlucVideoOverlayRectangle **) realloc (comp->rects, (comp->n_rects + 1) * sizeof (*n));
  if (!n)
    return;
  comp->rects = n;
  comp->rects[comp->n_rects++] = fluc_video_overlay_rectangle_ref (rect);
}

uint32_t
fluc_video_overlay_composition_n_rectangles (FlucVideoOverlayComposition *comp)
{
  return comp ? comp->n_rects : 0;
}

FlucVideoOverlayComposition *
fluc_video_overlay_composition_ref (FlucVideoOverlayComposition *comp)
{
  if (comp)
    __atomic_add_fetch (&comp->refcount, 1, __ATOMIC_SEQ_CST);
  return comp;
}

void
fluc_video_overlay_composition_unref (FlucVideoOverlayComposition *comp)
{
  uint32_t i;
  if (!comp || __atomic_sub_fetch (&comp->refcount, 1, __ATOMIC_SEQ_CST) != 0)
    return;
  pthread_mutex_lock (&g_lock);
  if (comp->uploaded && g_ctx)
    fluc_ttmlblend_overlay_clear (g_ctx, comp->stream);
  pthread_mutex_unlock (&g_lock);
  for (i = 0; i < comp->n_rects; i++)
    fluc_video_overlay_rectangle_unref (comp->rects[i]);
  free (comp->rects);
  free (comp);
}

int
fluc_video_overlay_composition_blend (FlucVideoOverlayComposition *comp, FlucVideoFrame *frame)
{
  FlucTtmlBlend *ctx;
  FlucTtmlBlendFrame f;
  uint64_t ticket = 0;
  int rc, i;

  if (!comp || !frame)
    return 0;
  pthread_mutex_lock (&g_lock);
  ctx = context_locked ();
  if (ctx && !comp->uploaded) {
    FlucTtmlBlendRectangle *rr =
        (FlucTtmlBlendRectangle *) calloc (comp->n_rects ? comp->n_rects : 1, sizeof (*rr));
    uint32_t n;
    for (n = 0; rr && n < comp->n_rects; n++) {
      const FlucVideoOverlayRectangle *r = comp->rects[n];
      rr[n].pixels = r->pixels;
      rr[n].width = r->width;
      rr[n].height = r->height;
      rr[n].stride = r->stride;
      rr[n].x = r->x;
      rr[n].y = r->y;
      rr[n].render_width = r->render_width;      /* != pixel size: scaled on the GPU first */
      rr[n].render_height = r->render_height;
      rr[n].global_alpha = r->global_alpha;
      rr[n].flags = (r->flags & FLUC_VIDEO_OVERLAY_FORMAT_FLAG_PREMULTIPLIED_ALPHA) ?
          FLUC_TTMLBLEND_FLAG_PREMULTIPLIED_ALPHA : 0;
    }
    rc = rr ? fluc_ttmlblend_overlay_set_rectangles (ctx, comp->stream, rr, comp->n_rects) :
        FLUC_TTMLBLEND_ERROR_OUT_OF_MEMORY;
    free (rr);
    if (rc == FLUC_TTMLBLEND_OK)
      comp->uploaded = 1;
    else
      ctx = NULL;
  }
  pthread_mutex_unlock (&g_lock);
  if (!ctx)
    return 0;

  memset (&f, 0, sizeof f);
  for (i = 0; i < 3; i++) {
    f.plane[i] = frame->data[i];
    f.stride[i] = frame->stride[i];
  }
  rc = fluc_ttmlblend_blend_host (ctx, comp->stream, frame->format, frame->width, frame->height,
      frame->flags, &f, &ticket);
  if (rc != FLUC_TTMLBLEND_OK)
    return 0;
  return fluc_ttmlblend_wait (ctx, ticket) == FLUC_TTMLBLEND_OK;
}
