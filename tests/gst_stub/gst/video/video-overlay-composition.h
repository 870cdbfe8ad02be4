/* stub: see ../../README.md */
#ifndef GST_STUB_VIDEO_OVERLAY_H
#define GST_STUB_VIDEO_OVERLAY_H
#include <gst/video/video.h>
#define GST_VIDEO_OVERLAY_COMPOSITION_FORMAT_RGB GST_VIDEO_FORMAT_BGRA
typedef enum { GST_VIDEO_OVERLAY_FORMAT_FLAG_NONE = 0, GST_VIDEO_OVERLAY_FORMAT_FLAG_PREMULTIPLIED_ALPHA = 1,
  GST_VIDEO_OVERLAY_FORMAT_FLAG_GLOBAL_ALPHA = 2 } GstVideoOverlayFormatFlags;
typedef struct _GstVideoOverlayRectangle GstVideoOverlayRectangle;
typedef struct _GstVideoOverlayComposition GstVideoOverlayComposition;
GstVideoOverlayRectangle *gst_video_overlay_rectangle_new_raw (GstBuffer *, gint, gint, guint, guint, GstVideoOverlayFormatFlags);
void gst_video_overlay_rectangle_unref (GstVideoOverlayRectangle *);
void gst_video_overlay_rectangle_set_global_alpha (GstVideoOverlayRectangle *, gfloat);
GstVideoOverlayComposition *gst_video_overlay_composition_new (GstVideoOverlayRectangle *);
void gst_video_overlay_composition_unref (GstVideoOverlayComposition *);
gboolean gst_video_overlay_composition_blend (GstVideoOverlayComposition *, GstVideoFrame *);
#endif
