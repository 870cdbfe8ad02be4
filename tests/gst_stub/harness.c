/*
 * harness.c -- what the rest of a pipeline does to the ttmlblend element, for the tests:
 * creates it, negotiates caps, asks it for an allocation proposal, and then pushes segments,
 * subtitle buffers (ttmlrender's BGRA images with PTS / duration), gaps, flushes and video
 * frames through its pads -- the calls upstream elements would make. Built together with
 * gstfake.c, gstttmlblend.c and gstflucallocator.c into libgstglue_test.so and driven from
 * tests/test_gpu_gstglue.py over ctypes. Test infrastructure only.
 */
#include <gst/gst.h>
#include <gst/base/gstbasetransform.h>
#include <gst/video/video.h>

#include "fluc_ttmlblend.h"
#include "gstflucallocator.h"

GType gst_ttmlblend_get_type (void);
void gst_ttmlblend_get_stats (GstElement * element, FlucTtmlBlendStats * stats, guint64 * frames_blended,
    guint64 * frames_passed);

typedef struct
{
  GstElement *element;
  GstBaseTransform *trans;
  GstPad *video_sink, *subtitle_sink;
  GstVideoInfo vinfo;
  GstAllocator *allocator;      /* what the element proposed */
  gboolean started;
} Harness;

#define VIS __attribute__ ((visibility ("default")))

VIS void *
th_new (int device, int auto_register)
{
  Harness *h = g_new0 (Harness, 1);
  gst_init (NULL, NULL);
  h->element = g_object_new (gst_ttmlblend_get_type (), NULL);
  h->trans = GST_BASE_TRANSFORM (h->element);
  g_fake_object_set_int (h->element, "device", device);
  g_fake_object_set_int (h->element, "auto-register", auto_register);
  h->video_sink = gst_element_get_static_pad (h->element, "sink");
  h->subtitle_sink = gst_element_get_static_pad (h->element, "subtitle_sink");
  return h;
}

/* READY -> PAUSED: start, caps on the video pad, allocation query. 0 on success. */
VIS int
th_start (void *hp, const char *format, int width, int height)
{
  Harness *h = hp;
  GstBaseTransformClass *k = GST_BASE_TRANSFORM_GET_CLASS (h->trans);
  GstCaps *caps = gst_fake_video_caps_new (format, width, height);
  GstQuery q = { caps, NULL, FALSE };
  if (!k->start (h->trans))
    return -1;
  h->started = TRUE;
  gst_fake_pad_set_caps (h->video_sink, caps);
  if (!k->set_caps (h->trans, caps, caps))
    return -2;
  if (!gst_video_info_from_caps (&h->vinfo, caps))
    return -3;
  if (k->propose_allocation && k->propose_allocation (h->trans, NULL, &q) && q.allocator && q.has_video_meta)
    h->allocator = q.allocator;
  gst_caps_unref (caps);
  return 0;
}

VIS int
th_has_allocator (void *hp)
{
  return ((Harness *) hp)->allocator != NULL;
}

VIS int
th_segment (void *hp, int subtitle, uint64_t start, uint64_t stop, uint64_t base)
{
  Harness *h = hp;
  GstSegment s;
  gst_segment_init (&s, GST_FORMAT_TIME);
  s.start = s.time = s.position = start;
  s.stop = stop;
  s.base = base;
  return gst_fake_pad_send_event (subtitle ? h->subtitle_sink : h->video_sink, gst_event_new_segment (&s)) ? 0 : -1;
}

/* ttmlrender's output: a w*h premultiplied BGRA image, stride w*4. Returns the GstFlowReturn. */
VIS int
th_push_subtitle (void *hp, const uint8_t * bgra, int w, int h_, uint64_t pts, uint64_t duration)
{
  Harness *h = hp;
  GstCaps *caps = gst_fake_video_caps_new ("BGRA", w, h_);
  GstBuffer *buf;
  gst_fake_pad_send_event (h->subtitle_sink, gst_event_new_caps (caps));
  gst_caps_unref (caps);
  buf = gst_buffer_new_wrapped (g_memdup2 (bgra, (gsize) w * h_ * 4), (gsize) w * h_ * 4);
  buf->pts = pts;
  buf->duration = duration;
  return gst_fake_pad_chain (h->subtitle_sink, buf);
}

VIS int
th_subtitle_event (void *hp, int type, uint64_t ts, uint64_t duration)
{
  Harness *h = hp;
  GstEvent *e;
  switch (type) {
    case GST_EVENT_GAP: e = gst_event_new_gap (ts, duration); break;
    case GST_EVENT_FLUSH_START: e = gst_event_new_flush_start (); break;
    case GST_EVENT_FLUSH_STOP: e = gst_event_new_flush_stop (TRUE); break;
    case GST_EVENT_EOS: e = gst_event_new_eos (); break;
    default: return -1;
  }
  return gst_fake_pad_send_event (h->subtitle_sink, e) ? 0 : -1;
}

/* One video frame in GStreamer's default layout (`data`, h->vinfo.size bytes), blended in place.
 * use_allocator: the frame travels in a buffer from the allocator the element proposed (the
 * bytes are copied in with the pool frame's strides and copied back out afterwards, as a
 * decoder writing into that buffer and a sink reading it would); otherwise `data` itself is the
 * buffer's memory (ordinary pageable memory). Returns the GstFlowReturn. */
VIS int
th_push_video (void *hp, uint8_t * data, uint64_t pts, int use_allocator)
{
  Harness *h = hp;
  GstBuffer *buf;
  GstFlowReturn ret;
  guint p;
  if (use_allocator) {
    GstVideoFrame f;
    if (!h->allocator)
      return -100;
    buf = gst_fluc_allocator_alloc_video_buffer (h->allocator);
    if (!buf)
      return -101;
    if (!gst_video_frame_map (&f, &h->vinfo, buf, GST_MAP_WRITE))
      return -102;
    for (p = 0; p < h->vinfo.n_planes; p++) {
      const gint rows = (p > 0 && h->vinfo.n_planes > 1 && h->vinfo.format != GST_VIDEO_FORMAT_UNKNOWN &&
          (h->vinfo.format == GST_VIDEO_FORMAT_I420 || h->vinfo.format == GST_VIDEO_FORMAT_YV12 ||
              h->vinfo.format == GST_VIDEO_FORMAT_NV12 || h->vinfo.format == GST_VIDEO_FORMAT_NV21)) ?
          (h->vinfo.height + 1) / 2 : h->vinfo.height;
      for (gint r = 0; r < rows; r++)
        memcpy ((guint8 *) f.data[p] + (gsize) r * f.info.stride[p],
            data + h->vinfo.offset[p] + (gsize) r * h->vinfo.stride[p], (gsize) h->vinfo.stride[p]);
    }
    gst_video_frame_unmap (&f);
  } else {
    buf = gst_buffer_new_wrapped_full (0, data, h->vinfo.size, 0, h->vinfo.size, NULL, NULL);
  }
  buf->pts = pts;
  gst_buffer_ref (buf);         /* ours; the chain call consumes the other reference */
  ret = gst_fake_pad_chain (h->video_sink, buf);
  if (use_allocator) {
    GstVideoFrame f;
    if (gst_video_frame_map (&f, &h->vinfo, buf, GST_MAP_READ)) {
      for (p = 0; p < h->vinfo.n_planes; p++) {
        const gint rows = (p > 0 && (h->vinfo.format == GST_VIDEO_FORMAT_I420 || h->vinfo.format == GST_VIDEO_FORMAT_YV12 ||
                h->vinfo.format == GST_VIDEO_FORMAT_NV12 || h->vinfo.format == GST_VIDEO_FORMAT_NV21)) ?
            (h->vinfo.height + 1) / 2 : h->vinfo.height;
        for (gint r = 0; r < rows; r++)
          memcpy (data + h->vinfo.offset[p] + (gsize) r * h->vinfo.stride[p],
              (guint8 *) f.data[p] + (gsize) r * f.info.stride[p], (gsize) h->vinfo.stride[p]);
      }
      gst_video_frame_unmap (&f);
    }
  }
  gst_buffer_unref (buf);
  return ret;
}

VIS uint64_t
th_frame_size (void *hp)
{
  return ((Harness *) hp)->vinfo.size;
}

VIS void
th_layout (void *hp, uint64_t * offsets, int32_t * strides)
{
  Harness *h = hp;
  for (guint p = 0; p < h->vinfo.n_planes; p++) {
    offsets[p] = h->vinfo.offset[p];
    strides[p] = h->vinfo.stride[p];
  }
}

VIS void
th_stats (void *hp, FlucTtmlBlendStats * st, uint64_t * blended, uint64_t * passed)
{
  Harness *h = hp;
  guint64 b = 0, p = 0;
  gst_ttmlblend_get_stats (h->element, st, &b, &p);
  *blended = b;
  *passed = p;
}

VIS int
th_errors (void *hp)
{
  return ((Harness *) hp)->element->fake_errors;
}

VIS void
th_free (void *hp)
{
  Harness *h = hp;
  if (h->allocator)
    gst_object_unref (h->allocator);
  if (h->started)
    GST_BASE_TRANSFORM_GET_CLASS (h->trans)->stop (h->trans);
  g_object_unref (h->element);
  g_free (h);
}
