/*
 * fluc_videooverlay.h -- GLib-free C mirror of the GStreamer calls the GPU path
 * replaces, layered on the C ABI of include/fluc_ttmlblend.h.
 *
 * Same names (gst_ -> fluc_), argument meaning and error behaviour as
 * gst-plugins-base's video-overlay-composition.h, restricted to what the
 * ttmlrender overlay needs:
 *   gst_video_overlay_rectangle_new_raw (pixels, x, y, render_w, render_h, flags)
 *   gst_video_overlay_rectangle_set_global_alpha (rect, alpha)
 *   gst_video_overlay_composition_new (rect) / _add_rectangle (comp, rect)
 *   gboolean gst_video_overlay_composition_blend (comp, GstVideoFrame *video_buf)
 * A composition is immutable once blended (as in GStreamer, where adding a
 * rectangle needs a writable = unshared composition): its rectangles are
 * uploaded and prepared on the GPU at the first blend and cached in HBM until
 * the last unref. blend () modifies the HOST frame in place, returns 1 (TRUE)
 * on success and 0 (FALSE) on unsupported format or any device error; there is
 * no CPU fallback.
 *
 * Producer side in the reference: gst_ttmlrender_gen_buffer () emits the BGRA
 * pixels (/root/reference/plugins/ttml/gstttmlrender.c:1427-1478); an element
 * wraps them with rectangle_new_raw (..., PREMULTIPLIED_ALPHA) once per cue.
 */
#ifndef _FLUC_VIDEOOVERLAY_H_
#define _FLUC_VIDEOOVERLAY_H_

#include "../../include/fluc_ttmlblend.h"

#ifdef __cplusplus
extern "C" {
#endif

#define FLUC_VIDEO_OVERLAY_FORMAT_FLAG_NONE 0u
#define FLUC_VIDEO_OVERLAY_FORMAT_FLAG_PREMULTIPLIED_ALPHA 1u
#define FLUC_VIDEO_OVERLAY_FORMAT_FLAG_GLOBAL_ALPHA 2u

typedef struct _FlucVideoOverlayRectangle FlucVideoOverlayRectangle;
typedef struct _FlucVideoOverlayComposition FlucVideoOverlayComposition;

/* A mapped GstVideoFrame, reduced to what the blend reads. */
typedef struct {
  FlucTtmlBlendFormat format;
  int32_t width, height;
  uint32_t flags;              /* FLUC_TTMLBLEND_FLAG_PREMULTIPLIED_ALPHA */
  void *data[3];
  int32_t stride[3];
} FlucVideoFrame;

/* The pixels are copied (GStreamer takes a GstBuffer + video meta: pixels, width, height,
 * stride stand in for it). render_width / render_height as in GStreamer; 0 = pixel size.
 * A render size that differs from the pixel size makes the composition scale the rectangle
 * (gst_video_blend_scale_linear_RGBA semantics, on the GPU, once). */
FLUC_EXPORT FlucVideoOverlayRectangle *fluc_video_overlay_rectangle_new_raw (
    const uint8_t *bgra_pixels, int32_t width, int32_t height, int32_t stride,
    int32_t render_x, int32_t render_y, uint32_t render_width, uint32_t render_height, uint32_t flags);
FLUC_EXPORT FlucVideoOverlayRectangle *fluc_video_overlay_rectangle_ref (FlucVideoOverlayRectangle *rect);
FLUC_EXPORT void fluc_video_overlay_rectangle_unref (FlucVideoOverlayRectangle *rect);
FLUC_EXPORT void fluc_video_overlay_rectangle_set_global_alpha (FlucVideoOverlayRectangle *rect, float global_alpha);
FLUC_EXPORT float fluc_video_overlay_rectangle_get_global_alpha (FlucVideoOverlayRectangle *rect);
FLUC_EXPORT void fluc_video_overlay_rectangle_set_render_rectangle (FlucVideoOverlayRectangle *rect,
    int32_t render_x, int32_t render_y, uint32_t render_width, uint32_t render_height);

FLUC_EXPORT FlucVideoOverlayComposition *fluc_video_overlay_composition_new (FlucVideoOverlayRectangle *rect);
FLUC_EXPORT void fluc_video_overlay_composition_add_rectangle (FlucVideoOverlayComposition *comp,
    FlucVideoOverlayRectangle *rect);
FLUC_EXPORT uint32_t fluc_video_overlay_composition_n_rectangles (FlucVideoOverlayComposition *comp);
FLUC_EXPORT FlucVideoOverlayComposition *fluc_video_overlay_composition_ref (FlucVideoOverlayComposition *comp);
FLUC_EXPORT void fluc_video_overlay_composition_unref (FlucVideoOverlayComposition *comp);

/* gboolean gst_video_overlay_composition_blend (comp, video_buf) */
FLUC_EXPORT int fluc_video_overlay_composition_blend (FlucVideoOverlayComposition *comp, FlucVideoFrame *video_buf);

/* Device used by blend () (default: env FLUC_TTMLBLEND_DEVICE or 0), and the
 * shared context behind it for callers that also want the batched API. */
FLUC_EXPORT int fluc_video_overlay_set_device (int device);
FLUC_EXPORT FlucTtmlBlend *fluc_video_overlay_get_context (void);
FLUC_EXPORT void fluc_video_overlay_deinit (void);

#ifdef __cplusplus
}
#endif
#endif
