/* functional fake: see ../../README.md */
#ifndef GST_STUB_VIDEO_H
#define GST_STUB_VIDEO_H
#include <gst/gst.h>
#define GST_VIDEO_MAX_PLANES 4
typedef enum { GST_VIDEO_FORMAT_UNKNOWN, GST_VIDEO_FORMAT_I420, GST_VIDEO_FORMAT_YV12, GST_VIDEO_FORMAT_NV12,
  GST_VIDEO_FORMAT_NV21, GST_VIDEO_FORMAT_AYUV, GST_VIDEO_FORMAT_ARGB, GST_VIDEO_FORMAT_ABGR,
  GST_VIDEO_FORMAT_RGBA, GST_VIDEO_FORMAT_BGRA, GST_VIDEO_FORMAT_RGBx, GST_VIDEO_FORMAT_BGRx,
  GST_VIDEO_FORMAT_xRGB, GST_VIDEO_FORMAT_xBGR, GST_VIDEO_FORMAT_Y42B, GST_VIDEO_FORMAT_Y444,
  GST_VIDEO_FORMAT_YUY2, GST_VIDEO_FORMAT_UYVY, GST_VIDEO_FORMAT_GRAY8, GST_VIDEO_FORMAT_NV16,
  GST_VIDEO_FORMAT_NV24, GST_VIDEO_FORMAT_NV61, GST_VIDEO_FORMAT_YVYU, GST_VIDEO_FORMAT_VYUY,
  GST_VIDEO_FORMAT_v308, GST_VIDEO_FORMAT_IYU2, GST_VIDEO_FORMAT_RGB, GST_VIDEO_FORMAT_BGR } GstVideoFormat;
typedef enum { GST_VIDEO_FLAG_NONE = 0, GST_VIDEO_FLAG_PREMULTIPLIED_ALPHA = 2 } GstVideoFlags;
typedef enum { GST_VIDEO_FRAME_FLAG_NONE = 0 } GstVideoFrameFlags;
typedef enum { GST_VIDEO_COMP_A = 3 } GstVideoCompStub;
typedef struct _GstVideoInfo { GstVideoFormat format; GstVideoFlags flags; gint width, height; gsize size;
  gsize offset[GST_VIDEO_MAX_PLANES]; gint stride[GST_VIDEO_MAX_PLANES]; guint n_planes; gboolean has_alpha; gint a_poffset; } GstVideoInfo;
#define GST_VIDEO_INFO_WIDTH(i) ((i)->width)
#define GST_VIDEO_INFO_HEIGHT(i) ((i)->height)
#define GST_VIDEO_INFO_FLAGS(i) ((i)->flags)
#define GST_VIDEO_INFO_N_PLANES(i) ((i)->n_planes)
#define GST_VIDEO_INFO_PLANE_OFFSET(i, p) ((i)->offset[p])
#define GST_VIDEO_INFO_PLANE_STRIDE(i, p) ((i)->stride[p])
#define GST_VIDEO_INFO_HAS_ALPHA(i) ((i)->has_alpha)
#define GST_VIDEO_INFO_COMP_POFFSET(i, c) ((i)->a_poffset + 0 * (c))
gboolean gst_video_info_from_caps (GstVideoInfo *, const GstCaps *);
gboolean gst_video_info_set_format (GstVideoInfo *, GstVideoFormat, guint, guint);
typedef struct _GstVideoFrame { GstVideoInfo info; GstVideoFrameFlags flags; GstBuffer *buffer; gpointer meta; gint id;
  gpointer data[GST_VIDEO_MAX_PLANES]; GstMapInfo map[GST_VIDEO_MAX_PLANES]; } GstVideoFrame;
gboolean gst_video_frame_map (GstVideoFrame *, const GstVideoInfo *, GstBuffer *, GstMapFlags);
void gst_video_frame_unmap (GstVideoFrame *);
#define GST_VIDEO_FRAME_N_PLANES(f) ((f)->info.n_planes)
#define GST_VIDEO_FRAME_PLANE_DATA(f, p) ((f)->data[p])
#define GST_VIDEO_FRAME_PLANE_STRIDE(f, p) ((f)->info.stride[p])
#define GST_VIDEO_FRAME_FORMAT(f) ((f)->info.format)
#define GST_VIDEO_FRAME_WIDTH(f) ((f)->info.width)
#define GST_VIDEO_FRAME_HEIGHT(f) ((f)->info.height)
#define GST_VIDEO_CAPS_MAKE(fmts) "video/x-raw, format = (string) " fmts
typedef struct _GstVideoMeta { GstBuffer *buffer; GstVideoFrameFlags flags; GstVideoFormat format; gint id; guint width, height;
  guint n_planes; gsize offset[GST_VIDEO_MAX_PLANES]; gint stride[GST_VIDEO_MAX_PLANES]; } GstVideoMeta;
struct _GstVideoMetaFake { GstVideoMeta meta; };
GstVideoMeta *gst_buffer_add_video_meta (GstBuffer *, GstVideoFrameFlags, GstVideoFormat, guint, guint);
GstVideoMeta *gst_buffer_add_video_meta_full (GstBuffer *, GstVideoFrameFlags, GstVideoFormat, guint, guint, guint n_planes,
    gsize offset[GST_VIDEO_MAX_PLANES], gint stride[GST_VIDEO_MAX_PLANES]);
GstVideoMeta *gst_buffer_get_video_meta (GstBuffer *);
GType gst_video_meta_api_get_type (void);
#define GST_VIDEO_META_API_TYPE (gst_video_meta_api_get_type ())
GstVideoFormat gst_video_format_from_string (const gchar *);
const gchar *gst_video_format_to_string (GstVideoFormat);
#define GST_VIDEO_INFO_FORMAT(i) ((i)->format)
#define GST_VIDEO_INFO_SIZE(i) ((i)->size)
#endif
