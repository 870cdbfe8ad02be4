/*
 * gstttmlblend.c -- GStreamer glue: a `ttmlblend` element that composites
 * ttmlrender's BGRA cue images onto raw video with the B200 path.
 *
 * NOT BUILT IN THE GRAFT IMAGE (no GLib / GStreamer headers there); it is the
 * element a maintainer compiles inside the reference tree with
 *   meson -Dttml_cuda=enabled     
 * next to plugins/ttml/gstttmlrender.c. It adds nothing to and changes nothing
 * in ttmlrender / ttmlparse: their pads, caps and properties stay as they are
 * (/root/reference/plugins/ttml/gstttmlrender.c:78-84,1673-1704,
 *  /root/reference/plugins/ttml/gstttmlbase.c:45-46,1625-1629).
 *
 *   filesrc ! ttmlrender ! video/x-raw,format=BGRA,width=W,height=H ! blend.subtitle_sink
 *   videotestsrc ! video/x-raw,format=NV12,width=W,height=H ! ttmlblend name=blend ! fakesink
 *
 * replaces the README pipeline's `compositor`
 * (/root/reference/plugins/ttml/README.md:45-48).
 *
 * Pads:  sink / src   video/x-raw { I420, YV12, NV12, NV21, AYUV, ARGB, ABGR, RGBA, BGRA,
 *                                   RGBx, BGRx, xRGB, xBGR, Y42B, Y444, YUY2, UYVY, GRAY8,
 *                                   NV16, NV24 }
 *        subtitle_sink  video/x-raw, format=BGRA   (= GST_TTMLRENDER_SRC_CAPS)
 * Each subtitle buffer is valid for [PTS, PTS+duration) (gst_ttmlbase_gen_buffer,
 * /root/reference/plugins/ttml/gstttmlbase.c:180-181); an all-zero "clear"
 * buffer for gaps (/root/reference/plugins/ttml/gstttmlevent.c:221-224) simply
 * blends nothing. The overlay is uploaded ONCE per subtitle buffer
 * (fluc_ttmlblend_overlay_set) and every video frame inside its interval runs
 * fluc_ttmlblend_blend_host in place: only the rows under the cue cross PCIe.
 * Errors surface as GST_FLOW_ERROR; there is no CPU fallback.
 */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif

#include <gst/gst.h>
#include <gst/base/gstbasetransform.h>
#include <gst/video/video.h>

#include "fluc_ttmlblend.h"

GST_DEBUG_CATEGORY_STATIC (ttmlblend_debug);
#define GST_CAT_DEFAULT ttmlblend_debug

#define GST_TYPE_TTMLBLEND (gst_ttmlblend_get_type ())
G_DECLARE_FINAL_TYPE (GstTTMLBlend, gst_ttmlblend, GST, TTMLBLEND, GstBaseTransform)

struct _GstTTMLBlend
{
  GstBaseTransform parent;
  GstPad *subtitle_sink;
  GstVideoInfo vinfo;

  GMutex lock;                  /* protects the fields below */
  FlucTtmlBlend *ctx;
  guint32 stream_id;
  GstClockTime ov_start, ov_stop;       /* validity of the cached overlay */
  gboolean have_overlay;
  gint device;
};

enum
{
  PROP_0,
  PROP_DEVICE
};

#define VIDEO_FORMATS "{ I420, YV12, NV12, NV21, AYUV, ARGB, ABGR, RGBA, BGRA, RGBx, BGRx, xRGB, xBGR, Y42B, Y444, YUY2, UYVY, GRAY8, NV16, NV24, NV61, YVYU, VYUY, v308, IYU2, RGB, BGR }"

static GstStaticPadTemplate video_sink_template = GST_STATIC_PAD_TEMPLATE ("sink",
    GST_PAD_SINK, GST_PAD_ALWAYS,
    GST_STATIC_CAPS (GST_VIDEO_CAPS_MAKE (VIDEO_FORMATS)));
static GstStaticPadTemplate video_src_template = GST_STATIC_PAD_TEMPLATE ("src",
    GST_PAD_SRC, GST_PAD_ALWAYS,
    GST_STATIC_CAPS (GST_VIDEO_CAPS_MAKE (VIDEO_FORMATS)));
static GstStaticPadTemplate subtitle_sink_template =
GST_STATIC_PAD_TEMPLATE ("subtitle_sink", GST_PAD_SINK, GST_PAD_ALWAYS,
    GST_STATIC_CAPS ("video/x-raw, format=BGRA, width=(int)[1,MAX], "
        "height=(int)[1,MAX], framerate=(fraction)0/1"));

G_DEFINE_TYPE (GstTTMLBlend, gst_ttmlblend, GST_TYPE_BASE_TRANSFORM);

static guint32 next_stream_id = 1;

/* One set of contexts for the whole process, shared by every ttmlblend element: frames of
 * different elements (streams) then meet in the same batch scheduler and share launches, and
 * the elements spread over the box's GPUs (device = -1: stream id % number of GPUs). */
static GMutex shared_lock;
static FlucTtmlBlendMulti *shared_multi = NULL;
static guint shared_refs = 0;

static FlucTtmlBlend *
shared_context_acquire (gint device, guint32 stream_id, int *rc)
{
  FlucTtmlBlend *ctx = NULL;
  g_mutex_lock (&shared_lock);
  *rc = FLUC_TTMLBLEND_OK;
  if (!shared_multi)
    *rc = fluc_ttmlblend_multi_new (NULL, 0, &shared_multi);
  if (*rc == FLUC_TTMLBLEND_OK) {
    if (device >= (gint) fluc_ttmlblend_multi_size (shared_multi))
      *rc = FLUC_TTMLBLEND_ERROR_NO_DEVICE;
    else
      /* every device of the box is in the set, in order: context i is device i */
      ctx = fluc_ttmlblend_multi_context (shared_multi, device < 0 ? stream_id : (guint32) device);
  }
  if (ctx)
    shared_refs++;
  g_mutex_unlock (&shared_lock);
  return ctx;
}

static void
shared_context_release (void)
{
  g_mutex_lock (&shared_lock);
  if (shared_refs > 0 && --shared_refs == 0) {
    fluc_ttmlblend_multi_free (shared_multi);
    shared_multi = NULL;
  }
  g_mutex_unlock (&shared_lock);
}

static FlucTtmlBlendFormat
to_fluc_format (GstVideoFormat f)
{
  switch (f) {
    case GST_VIDEO_FORMAT_I420: return FLUC_TTMLBLEND_FORMAT_I420;
    case GST_VIDEO_FORMAT_YV12: return FLUC_TTMLBLEND_FORMAT_YV12;
    case GST_VIDEO_FORMAT_NV12: return FLUC_TTMLBLEND_FORMAT_NV12;
    case GST_VIDEO_FORMAT_NV21: return FLUC_TTMLBLEND_FORMAT_NV21;
    case GST_VIDEO_FORMAT_AYUV: return FLUC_TTMLBLEND_FORMAT_AYUV;
    case GST_VIDEO_FORMAT_ARGB: return FLUC_TTMLBLEND_FORMAT_ARGB;
    case GST_VIDEO_FORMAT_ABGR: return FLUC_TTMLBLEND_FORMAT_ABGR;
    case GST_VIDEO_FORMAT_RGBA: return FLUC_TTMLBLEND_FORMAT_RGBA;
    case GST_VIDEO_FORMAT_BGRA: return FLUC_TTMLBLEND_FORMAT_BGRA;
    case GST_VIDEO_FORMAT_RGBx: return FLUC_TTMLBLEND_FORMAT_RGBx;
    case GST_VIDEO_FORMAT_BGRx: return FLUC_TTMLBLEND_FORMAT_BGRx;
    case GST_VIDEO_FORMAT_xRGB: return FLUC_TTMLBLEND_FORMAT_xRGB;
    case GST_VIDEO_FORMAT_xBGR: return FLUC_TTMLBLEND_FORMAT_xBGR;
    case GST_VIDEO_FORMAT_Y42B: return FLUC_TTMLBLEND_FORMAT_Y42B;
    case GST_VIDEO_FORMAT_Y444: return FLUC_TTMLBLEND_FORMAT_Y444;
    case GST_VIDEO_FORMAT_YUY2: return FLUC_TTMLBLEND_FORMAT_YUY2;
    case GST_VIDEO_FORMAT_UYVY: return FLUC_TTMLBLEND_FORMAT_UYVY;
    case GST_VIDEO_FORMAT_GRAY8: return FLUC_TTMLBLEND_FORMAT_GRAY8;
    case GST_VIDEO_FORMAT_NV16: return FLUC_TTMLBLEND_FORMAT_NV16;
    case GST_VIDEO_FORMAT_NV24: return FLUC_TTMLBLEND_FORMAT_NV24;
    case GST_VIDEO_FORMAT_NV61: return FLUC_TTMLBLEND_FORMAT_NV61;
    case GST_VIDEO_FORMAT_YVYU: return FLUC_TTMLBLEND_FORMAT_YVYU;
    case GST_VIDEO_FORMAT_VYUY: return FLUC_TTMLBLEND_FORMAT_VYUY;
    case GST_VIDEO_FORMAT_v308: return FLUC_TTMLBLEND_FORMAT_v308;
    case GST_VIDEO_FORMAT_IYU2: return FLUC_TTMLBLEND_FORMAT_IYU2;
    case GST_VIDEO_FORMAT_RGB: return FLUC_TTMLBLEND_FORMAT_RGB;
    case GST_VIDEO_FORMAT_BGR: return FLUC_TTMLBLEND_FORMAT_BGR;
    default: return FLUC_TTMLBLEND_FORMAT_COUNT;
  }
}

/* subtitle streaming thread: one BGRA buffer per timeline interval */
static GstFlowReturn
gst_ttmlblend_subtitle_chain (GstPad * pad, GstObject * parent, GstBuffer * buf)
{
  GstTTMLBlend *self = GST_TTMLBLEND (parent);
  GstCaps *caps = gst_pad_get_current_caps (pad);
  GstVideoInfo oinfo;
  GstMapInfo map;
  GstFlowReturn ret = GST_FLOW_OK;
  int rc;

  if (!caps || !gst_video_info_from_caps (&oinfo, caps)) {
    if (caps)
      gst_caps_unref (caps);
    gst_buffer_unref (buf);
    return GST_FLOW_NOT_NEGOTIATED;
  }
  gst_caps_unref (caps);

  if (!gst_buffer_map (buf, &map, GST_MAP_READ)) {
    gst_buffer_unref (buf);
    return GST_FLOW_ERROR;
  }
  g_mutex_lock (&self->lock);
  /* W*H*4 bytes, stride W*4, premultiplied (gstttmlrender.c:1442-1448) */
  rc = fluc_ttmlblend_overlay_set (self->ctx, self->stream_id, map.data,
      GST_VIDEO_INFO_WIDTH (&oinfo), GST_VIDEO_INFO_HEIGHT (&oinfo),
      GST_VIDEO_INFO_WIDTH (&oinfo) * 4, NULL, 0);
  if (rc == FLUC_TTMLBLEND_OK) {
    self->ov_start = GST_BUFFER_PTS (buf);
    self->ov_stop = GST_BUFFER_DURATION_IS_VALID (buf) ?
        GST_BUFFER_PTS (buf) + GST_BUFFER_DURATION (buf) : GST_CLOCK_TIME_NONE;
    self->have_overlay = TRUE;
  } else {
    GST_ELEMENT_ERROR (self, LIBRARY, FAILED, ("ttmlblend overlay upload failed"),
        ("%s: %s", fluc_ttmlblend_strerror (rc), fluc_ttmlblend_last_cuda_error (self->ctx)));
    ret = GST_FLOW_ERROR;
  }
  g_mutex_unlock (&self->lock);
  gst_buffer_unmap (buf, &map);
  gst_buffer_unref (buf);
  return ret;
}

static gboolean
gst_ttmlblend_subtitle_event (GstPad * pad, GstObject * parent, GstEvent * event)
{
  GstTTMLBlend *self = GST_TTMLBLEND (parent);
  switch (GST_EVENT_TYPE (event)) {
    case GST_EVENT_FLUSH_STOP:
    case GST_EVENT_EOS:
      g_mutex_lock (&self->lock);
      if (self->ctx)
        fluc_ttmlblend_overlay_clear (self->ctx, self->stream_id);
      self->have_overlay = FALSE;
      g_mutex_unlock (&self->lock);
      break;
    default:
      break;
  }
  /* subtitle events stop here except caps, which the pad stores itself */
  gst_event_unref (event);
  return TRUE;
}

static gboolean
gst_ttmlblend_set_caps (GstBaseTransform * trans, GstCaps * incaps, GstCaps * outcaps)
{
  GstTTMLBlend *self = GST_TTMLBLEND (trans);
  return gst_video_info_from_caps (&self->vinfo, incaps);
}

/* video streaming thread: gst_video_overlay_composition_blend (comp, frame) */
static GstFlowReturn
gst_ttmlblend_transform_ip (GstBaseTransform * trans, GstBuffer * buf)
{
  GstTTMLBlend *self = GST_TTMLBLEND (trans);
  GstClockTime ts = GST_BUFFER_PTS (buf);
  GstVideoFrame frame;
  FlucTtmlBlendFrame f = { {NULL, NULL, NULL}, {0, 0, 0} };
  guint64 ticket = 0;
  gboolean active;
  guint p;
  int rc;

  g_mutex_lock (&self->lock);
  active = self->have_overlay && (!GST_CLOCK_TIME_IS_VALID (ts) ||
      (ts >= self->ov_start && (!GST_CLOCK_TIME_IS_VALID (self->ov_stop) || ts < self->ov_stop)));
  g_mutex_unlock (&self->lock);
  if (!active)
    return GST_FLOW_OK;

  if (!gst_video_frame_map (&frame, &self->vinfo, buf, GST_MAP_READWRITE))
    return GST_FLOW_ERROR;
  for (p = 0; p < GST_VIDEO_FRAME_N_PLANES (&frame); p++) {
    f.plane[p] = GST_VIDEO_FRAME_PLANE_DATA (&frame, p);
    f.stride[p] = GST_VIDEO_FRAME_PLANE_STRIDE (&frame, p);
  }
  rc = fluc_ttmlblend_blend_host (self->ctx, self->stream_id,
      to_fluc_format (GST_VIDEO_FRAME_FORMAT (&frame)), GST_VIDEO_FRAME_WIDTH (&frame),
      GST_VIDEO_FRAME_HEIGHT (&frame),
      (GST_VIDEO_INFO_FLAGS (&frame.info) & GST_VIDEO_FLAG_PREMULTIPLIED_ALPHA) ?
      FLUC_TTMLBLEND_FLAG_PREMULTIPLIED_ALPHA : 0, &f, &ticket);
  if (rc == FLUC_TTMLBLEND_OK)
    rc = fluc_ttmlblend_wait (self->ctx, ticket);
  gst_video_frame_unmap (&frame);
  if (rc != FLUC_TTMLBLEND_OK) {
    GST_ELEMENT_ERROR (self, LIBRARY, FAILED, ("ttmlblend failed"),
        ("%s: %s", fluc_ttmlblend_strerror (rc), fluc_ttmlblend_last_cuda_error (self->ctx)));
    return GST_FLOW_ERROR;
  }
  return GST_FLOW_OK;
}

static gboolean
gst_ttmlblend_start (GstBaseTransform * trans)
{
  GstTTMLBlend *self = GST_TTMLBLEND (trans);
  int rc;
  self->ctx = shared_context_acquire (self->device, self->stream_id, &rc);
  if (!self->ctx) {
    GST_ELEMENT_ERROR (self, LIBRARY, INIT, ("no usable CUDA device"),
        ("%s", fluc_ttmlblend_strerror (rc)));
    return FALSE;
  }
  /* frames that do not come from gstflucallocator.c are pageable: pin the buffers of the
   * upstream pool as they show up, so that they are blended zero-copy from then on */
  fluc_ttmlblend_set_auto_register (self->ctx, 1);
  self->have_overlay = FALSE;
  return TRUE;
}

static gboolean
gst_ttmlblend_stop (GstBaseTransform * trans)
{
  GstTTMLBlend *self = GST_TTMLBLEND (trans);
  g_mutex_lock (&self->lock);
  if (self->ctx) {
    /* the context lives on for the other elements: drop only this stream's overlay */
    fluc_ttmlblend_overlay_clear (self->ctx, self->stream_id);
    shared_context_release ();
  }
  self->ctx = NULL;
  self->have_overlay = FALSE;
  g_mutex_unlock (&self->lock);
  return TRUE;
}

static void
gst_ttmlblend_set_property (GObject * object, guint prop_id, const GValue * value,
    GParamSpec * pspec)
{
  GstTTMLBlend *self = GST_TTMLBLEND (object);
  switch (prop_id) {
    case PROP_DEVICE:
      self->device = g_value_get_int (value);
      break;
    default:
      G_OBJECT_WARN_INVALID_PROPERTY_ID (object, prop_id, pspec);
      break;
  }
}

static void
gst_ttmlblend_get_property (GObject * object, guint prop_id, GValue * value, GParamSpec * pspec)
{
  GstTTMLBlend *self = GST_TTMLBLEND (object);
  switch (prop_id) {
    case PROP_DEVICE:
      g_value_set_int (value, self->device);
      break;
    default:
      G_OBJECT_WARN_INVALID_PROPERTY_ID (object, prop_id, pspec);
      break;
  }
}

static void
gst_ttmlblend_finalize (GObject * object)
{
  GstTTMLBlend *self = GST_TTMLBLEND (object);
  g_mutex_clear (&self->lock);
  G_OBJECT_CLASS (gst_ttmlblend_parent_class)->finalize (object);
}

static void
gst_ttmlblend_class_init (GstTTMLBlendClass * klass)
{
  GObjectClass *gobject_class = G_OBJECT_CLASS (klass);
  GstElementClass *element_class = GST_ELEMENT_CLASS (klass);
  GstBaseTransformClass *bt_class = GST_BASE_TRANSFORM_CLASS (klass);

  gobject_class->set_property = gst_ttmlblend_set_property;
  gobject_class->get_property = gst_ttmlblend_get_property;
  gobject_class->finalize = gst_ttmlblend_finalize;
  g_object_class_install_property (gobject_class, PROP_DEVICE,
      g_param_spec_int ("device", "CUDA device",
          "CUDA device index (-1: spread the elements over all GPUs, stream id % n)", -1, 64, -1,
          G_PARAM_READWRITE | G_PARAM_STATIC_STRINGS));

  gst_element_class_add_static_pad_template (element_class, &video_sink_template);
  gst_element_class_add_static_pad_template (element_class, &video_src_template);
  gst_element_class_add_static_pad_template (element_class, &subtitle_sink_template);
  gst_element_class_set_static_metadata (element_class, "TTML blender (CUDA)",
      "Filter/Editor/Video/Overlay/Subtitle",
      "Blends ttmlrender's BGRA cue images onto raw video on an NVIDIA B200",
      "flu-plugins-oss_b200");

  bt_class->set_caps = GST_DEBUG_FUNCPTR (gst_ttmlblend_set_caps);
  bt_class->transform_ip = GST_DEBUG_FUNCPTR (gst_ttmlblend_transform_ip);
  bt_class->start = GST_DEBUG_FUNCPTR (gst_ttmlblend_start);
  bt_class->stop = GST_DEBUG_FUNCPTR (gst_ttmlblend_stop);
  GST_DEBUG_CATEGORY_INIT (ttmlblend_debug, "ttmlblend", 0, "TTML CUDA blender");
}

static void
gst_ttmlblend_init (GstTTMLBlend * self)
{
  g_mutex_init (&self->lock);
  self->device = -1;
  self->stream_id = g_atomic_int_add ((gint *) & next_stream_id, 1);
  self->subtitle_sink = gst_pad_new_from_static_template (&subtitle_sink_template, "subtitle_sink");
  gst_pad_set_chain_function (self->subtitle_sink, GST_DEBUG_FUNCPTR (gst_ttmlblend_subtitle_chain));
  gst_pad_set_event_function (self->subtitle_sink, GST_DEBUG_FUNCPTR (gst_ttmlblend_subtitle_event));
  gst_element_add_pad (GST_ELEMENT (self), self->subtitle_sink);
  gst_base_transform_set_in_place (GST_BASE_TRANSFORM (self), TRUE);
}

/* registered from plugin_init of gstfluttml.c next to ttmlparse / ttmlrender
 * (/root/reference/plugins/ttml/gstfluttml.c:41-59):
 *   gst_element_register (plugin, "ttmlblend", GST_RANK_NONE, GST_TYPE_TTMLBLEND)   */
gboolean
gst_ttmlblend_register (GstPlugin * plugin)
{
  return gst_element_register (plugin, "ttmlblend", GST_RANK_NONE, GST_TYPE_TTMLBLEND);
}
