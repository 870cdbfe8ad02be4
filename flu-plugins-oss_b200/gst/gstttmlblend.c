/*
 * gstttmlblend.c -- GStreamer glue: a `ttmlblend` element that composites
 * ttmlrender's BGRA cue images onto raw video with the B200 path.
 *
 * It is the element a maintainer compiles inside the reference tree with
 *   meson -Dttml_cuda=enabled     
 * next to plugins/ttml/gstttmlrender.c. It adds nothing to and changes nothing
 * in ttmlrender / ttmlparse: their pads, caps and properties stay as they are
 * (/root/reference/plugins/ttml/gstttmlrender.c:78-84,1673-1704,
 *  /root/reference/plugins/ttml/gstttmlbase.c:45-46,1625-1629). The graft image has no
 * GLib / GStreamer; there this file is compiled and RUN against the functional fake in
 * tests/gst_stub/ (tests/test_gpu_gstglue.py pushes cues, gaps and frames through it).
 *
 *   filesrc ! ttmlrender ! video/x-raw,format=BGRA,width=W,height=H ! blend.subtitle_sink
 *   videotestsrc ! video/x-raw,format=NV12,width=W,height=H ! ttmlblend name=blend ! fakesink
 *
 * replaces the README pipeline's `compositor`
 * (/root/reference/plugins/ttml/README.md:45-48).
 *
 * Pads:  sink / src   video/x-raw { I420, YV12, NV12, NV21, AYUV, ARGB, ABGR, RGBA, BGRA, ... }
 *        subtitle_sink  video/x-raw, format=BGRA   (= GST_TTMLRENDER_SRC_CAPS)
 *
 * Timing. ttmlrender pushes one buffer per timeline interval, valid for [PTS, PTS+duration)
 * (gst_ttmlbase_gen_buffer, /root/reference/plugins/ttml/gstttmlbase.c:180-181), an all-zero
 * "clear" buffer for gaps (/root/reference/plugins/ttml/gstttmlevent.c:221-224) -- and it
 * pushes them as fast as it parses, i.e. AHEAD of the video. So cues are queued: each one is
 * uploaded when it arrives (fluc_ttmlblend_overlay_set on the subtitle streaming thread, into
 * one of CUE_SLOTS overlay slots of this element), and becomes the active cue only when the
 * video's running time reaches its start; the subtitle thread blocks while every slot is taken
 * (the back-pressure textoverlay / overlaycomposition apply). Both sides are compared in
 * running time (gst_segment_to_running_time with each pad's own segment). Every video frame
 * inside the active cue's interval runs fluc_ttmlblend_blend_host in place: only the rows under
 * the cue cross PCIe. Errors surface as GST_FLOW_ERROR; there is no CPU fallback.
 *
 * Memory. Frames from the pinned allocator this element proposes upstream
 * (gstflucallocator.c, propose_allocation) are blended zero copy. Any other frame is ordinary
 * pageable memory: the library stages its cue rows through pinned frames on worker threads.
 * With auto-register=true such memory is pinned the first time it is seen instead (buffer pools
 * recycle it) and unpinned again when the GstMemory is finalised.
 */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif

#include <gst/gst.h>
#include <gst/base/gstbasetransform.h>
#include <gst/video/video.h>

#include "fluc_ttmlblend.h"
#include "gstflucallocator.h"

GST_DEBUG_CATEGORY_STATIC (ttmlblend_debug);
#define GST_CAT_DEFAULT ttmlblend_debug

#define GST_TYPE_TTMLBLEND (gst_ttmlblend_get_type ())
G_DECLARE_FINAL_TYPE (GstTTMLBlend, gst_ttmlblend, GST, TTMLBLEND, GstBaseTransform)

#define CUE_SLOTS 8

typedef struct
{
  GstClockTime start, stop;     /* running time; stop may be GST_CLOCK_TIME_NONE */
  guint slot;                   /* overlay slot the image was uploaded into */
  gboolean clear;               /* a gap: nothing is shown from `start` on */
} GstTTMLBlendCue;

struct _GstTTMLBlend
{
  GstBaseTransform parent;
  GstPad *subtitle_sink;
  GstVideoInfo vinfo;
  gboolean have_vinfo;

  GMutex lock;                  /* protects the fields below */
  GCond cond;                   /* a slot became free / flushing */
  FlucTtmlBlend *ctx;
  guint32 element_id;
  GstSegment video_segment, subtitle_segment;
  GstTTMLBlendCue queue[CUE_SLOTS];     /* cues that have not started yet, by start time */
  guint n_queued;
  GstTTMLBlendCue current;      /* the cue the video position is in */
  gboolean have_current;
  guint next_slot;
  gboolean slot_busy[CUE_SLOTS];
  gboolean subtitle_flushing;
  GstAllocator *allocator;      /* proposed upstream (pinned pool frames) */
  gint device;
  gboolean auto_register;
  guint64 frames_blended, frames_passed;
};

enum
{
  PROP_0,
  PROP_DEVICE,
  PROP_AUTO_REGISTER
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

static gint next_element_id = 1;

/* One set of contexts for the whole process, shared by every ttmlblend element: frames of
 * different elements (streams) then meet in the same batch scheduler and share launches, and
 * the elements spread over the box's GPUs (device = -1: element id % number of GPUs). */
static GMutex shared_lock;
static FlucTtmlBlendMulti *shared_multi = NULL;
static guint shared_refs = 0;
static guint shared_generation = 0;     /* bumped whenever the set is torn down */

static FlucTtmlBlend *
shared_context_acquire (gint device, guint32 element_id, int *rc)
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
      ctx = fluc_ttmlblend_multi_context (shared_multi, device < 0 ? element_id : (guint32) device);
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
    fluc_ttmlblend_multi_free (shared_multi);   /* unpins whatever was registered automatically */
    shared_multi = NULL;
    shared_generation++;
  }
  g_mutex_unlock (&shared_lock);
}

/* the overlay stream id of one cue slot: unique per element inside the process-wide context */
static guint32
slot_stream (GstTTMLBlend * self, guint slot)
{
  return self->element_id * CUE_SLOTS + slot;
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

/* ---- cue queue (lock held) ------------------------------------------------ */

static void
release_slot (GstTTMLBlend * self, guint slot)
{
  if (self->ctx)
    fluc_ttmlblend_overlay_clear (self->ctx, slot_stream (self, slot));
  self->slot_busy[slot] = FALSE;
  g_cond_broadcast (&self->cond);
}

static void
drop_all_cues (GstTTMLBlend * self)
{
  guint i;
  for (i = 0; i < self->n_queued; i++)
    if (!self->queue[i].clear)
      release_slot (self, self->queue[i].slot);
  self->n_queued = 0;
  if (self->have_current && !self->current.clear)
    release_slot (self, self->current.slot);
  self->have_current = FALSE;
  g_cond_broadcast (&self->cond);
}

/* inserts by start time; a cue that starts where (or before) an already queued one starts
 * comes after it, so that the later arrival wins when both have started */
static void
enqueue_cue (GstTTMLBlend * self, const GstTTMLBlendCue * cue)
{
  guint pos = self->n_queued, i;
  while (pos > 0 && self->queue[pos - 1].start > cue->start)
    pos--;
  for (i = self->n_queued; i > pos; i--)
    self->queue[i] = self->queue[i - 1];
  self->queue[pos] = *cue;
  self->n_queued++;
}

/* the video is at running time `rt`: every queued cue that has started takes over in turn */
static void
advance_cues (GstTTMLBlend * self, GstClockTime rt)
{
  while (self->n_queued > 0 && self->queue[0].start <= rt) {
    guint i;
    if (self->have_current && !self->current.clear)
      release_slot (self, self->current.slot);
    self->current = self->queue[0];
    self->have_current = TRUE;
    for (i = 1; i < self->n_queued; i++)
      self->queue[i - 1] = self->queue[i];
    self->n_queued--;
    g_cond_broadcast (&self->cond);     /* room in the queue */
  }
}

static gint
find_free_slot (GstTTMLBlend * self)
{
  guint k;
  for (k = 0; k < CUE_SLOTS; k++) {
    const guint s = (self->next_slot + k) % CUE_SLOTS;
    if (!self->slot_busy[s]) {
      self->next_slot = (s + 1) % CUE_SLOTS;
      return (gint) s;
    }
  }
  return -1;
}

/* ---- subtitle pad ----------------------------------------------------------- */

/* subtitle streaming thread: one BGRA buffer per timeline interval */
static GstFlowReturn
gst_ttmlblend_subtitle_chain (GstPad * pad, GstObject * parent, GstBuffer * buf)
{
  GstTTMLBlend *self = GST_TTMLBLEND (parent);
  GstCaps *caps = gst_pad_get_current_caps (pad);
  GstVideoInfo oinfo;
  GstMapInfo map;
  GstFlowReturn ret = GST_FLOW_OK;
  GstTTMLBlendCue cue;
  gint slot = -1;
  int rc;

  if (!caps || !gst_video_info_from_caps (&oinfo, caps)) {
    if (caps)
      gst_caps_unref (caps);
    gst_buffer_unref (buf);
    return GST_FLOW_NOT_NEGOTIATED;
  }
  gst_caps_unref (caps);
  if (!GST_BUFFER_PTS_IS_VALID (buf)) {
    gst_buffer_unref (buf);     /* ttmlrender always stamps its buffers (gstttmlbase.c:180) */
    return GST_FLOW_OK;
  }

  g_mutex_lock (&self->lock);
  /* both sides in running time: the subtitle branch has a segment of its own */
  cue.start = gst_segment_to_running_time (&self->subtitle_segment, GST_FORMAT_TIME, GST_BUFFER_PTS (buf));
  cue.stop = GST_CLOCK_TIME_NONE;
  if (GST_BUFFER_DURATION_IS_VALID (buf))
    cue.stop = gst_segment_to_running_time (&self->subtitle_segment, GST_FORMAT_TIME,
        GST_BUFFER_PTS (buf) + GST_BUFFER_DURATION (buf));
  cue.clear = FALSE;
  if (!GST_CLOCK_TIME_IS_VALID (cue.start)) {
    g_mutex_unlock (&self->lock);       /* outside the segment: never shown */
    gst_buffer_unref (buf);
    return GST_FLOW_OK;
  }
  /* ttmlrender runs ahead of the video: wait for a free slot (the video thread frees them as
   * cues start and end) */
  while (!self->subtitle_flushing && self->ctx &&
      (self->n_queued >= CUE_SLOTS - 1 || (slot = find_free_slot (self)) < 0))
    g_cond_wait (&self->cond, &self->lock);
  if (self->subtitle_flushing || !self->ctx) {
    g_mutex_unlock (&self->lock);
    gst_buffer_unref (buf);
    return GST_FLOW_FLUSHING;
  }
  self->slot_busy[slot] = TRUE;
  cue.slot = (guint) slot;
  g_mutex_unlock (&self->lock);

  /* the upload runs unlocked: video frames keep flowing under the cue that is showing */
  if (!gst_buffer_map (buf, &map, GST_MAP_READ)) {
    g_mutex_lock (&self->lock);
    release_slot (self, cue.slot);
    g_mutex_unlock (&self->lock);
    gst_buffer_unref (buf);
    return GST_FLOW_ERROR;
  }
  /* W*H*4 bytes, stride W*4, premultiplied (gstttmlrender.c:1442-1448) */
  rc = fluc_ttmlblend_overlay_set (self->ctx, slot_stream (self, cue.slot), map.data,
      GST_VIDEO_INFO_WIDTH (&oinfo), GST_VIDEO_INFO_HEIGHT (&oinfo),
      GST_VIDEO_INFO_WIDTH (&oinfo) * 4, NULL, 0);
  gst_buffer_unmap (buf, &map);
  g_mutex_lock (&self->lock);
  if (rc != FLUC_TTMLBLEND_OK) {
    release_slot (self, cue.slot);
    GST_ELEMENT_ERROR (self, LIBRARY, FAILED, ("ttmlblend overlay upload failed"),
        ("%s: %s", fluc_ttmlblend_strerror (rc), fluc_ttmlblend_last_cuda_error (self->ctx)));
    ret = GST_FLOW_ERROR;
  } else if (self->subtitle_flushing) {
    release_slot (self, cue.slot);
    ret = GST_FLOW_FLUSHING;
  } else {
    enqueue_cue (self, &cue);
  }
  g_mutex_unlock (&self->lock);
  gst_buffer_unref (buf);
  return ret;
}

static gboolean
gst_ttmlblend_subtitle_event (GstPad * pad, GstObject * parent, GstEvent * event)
{
  GstTTMLBlend *self = GST_TTMLBLEND (parent);
  (void) pad;
  switch (GST_EVENT_TYPE (event)) {
    case GST_EVENT_SEGMENT:
      g_mutex_lock (&self->lock);
      gst_event_copy_segment (event, &self->subtitle_segment);
      g_mutex_unlock (&self->lock);
      break;
    case GST_EVENT_GAP:{
      /* nothing is shown from the gap's start on (until the next cue starts) */
      GstClockTime ts = GST_CLOCK_TIME_NONE, dur = GST_CLOCK_TIME_NONE;
      GstTTMLBlendCue cue;
      gst_event_parse_gap (event, &ts, &dur);
      g_mutex_lock (&self->lock);
      cue.start = gst_segment_to_running_time (&self->subtitle_segment, GST_FORMAT_TIME, ts);
      cue.stop = GST_CLOCK_TIME_NONE;
      cue.slot = 0;
      cue.clear = TRUE;
      while (!self->subtitle_flushing && self->ctx && self->n_queued >= CUE_SLOTS - 1)
        g_cond_wait (&self->cond, &self->lock);
      if (GST_CLOCK_TIME_IS_VALID (cue.start) && !self->subtitle_flushing && self->ctx)
        enqueue_cue (self, &cue);
      g_mutex_unlock (&self->lock);
      break;
    }
    case GST_EVENT_FLUSH_START:
      g_mutex_lock (&self->lock);
      self->subtitle_flushing = TRUE;
      g_cond_broadcast (&self->cond);
      g_mutex_unlock (&self->lock);
      break;
    case GST_EVENT_FLUSH_STOP:
      g_mutex_lock (&self->lock);
      self->subtitle_flushing = FALSE;
      drop_all_cues (self);
      gst_segment_init (&self->subtitle_segment, GST_FORMAT_TIME);
      g_mutex_unlock (&self->lock);
      break;
    case GST_EVENT_EOS:
      /* no more cues: what is queued still plays out */
      break;
    default:
      break;
  }
  /* subtitle events stop here except caps, which the pad stores itself */
  gst_event_unref (event);
  return TRUE;
}

/* ---- video side --------------------------------------------------------------- */

static gboolean
gst_ttmlblend_set_caps (GstBaseTransform * trans, GstCaps * incaps, GstCaps * outcaps)
{
  GstTTMLBlend *self = GST_TTMLBLEND (trans);
  (void) outcaps;
  self->have_vinfo = gst_video_info_from_caps (&self->vinfo, incaps);
  return self->have_vinfo;
}

static gboolean
gst_ttmlblend_sink_event (GstBaseTransform * trans, GstEvent * event)
{
  GstTTMLBlend *self = GST_TTMLBLEND (trans);
  switch (GST_EVENT_TYPE (event)) {
    case GST_EVENT_SEGMENT:
      g_mutex_lock (&self->lock);
      gst_event_copy_segment (event, &self->video_segment);
      g_mutex_unlock (&self->lock);
      break;
    case GST_EVENT_FLUSH_STOP:
      g_mutex_lock (&self->lock);
      gst_segment_init (&self->video_segment, GST_FORMAT_TIME);
      g_mutex_unlock (&self->lock);
      break;
    default:
      break;
  }
  return GST_BASE_TRANSFORM_CLASS (gst_ttmlblend_parent_class)->sink_event (trans, event);
}

/* upstream asks what memory we would like: pinned pool frames, which the kernel reaches over
 * PCIe without a copy */
static gboolean
gst_ttmlblend_propose_allocation (GstBaseTransform * trans, GstQuery * decide_query, GstQuery * query)
{
  GstTTMLBlend *self = GST_TTMLBLEND (trans);
  GstCaps *caps = NULL;
  GstVideoInfo info;
  FlucTtmlBlendFormat fmt;
  (void) decide_query;

  gst_query_parse_allocation (query, &caps, NULL);
  if (!caps || !gst_video_info_from_caps (&info, caps) || !self->ctx)
    return FALSE;
  fmt = to_fluc_format (GST_VIDEO_INFO_FORMAT (&info));
  if (fmt == FLUC_TTMLBLEND_FORMAT_COUNT)
    return FALSE;
  g_mutex_lock (&self->lock);
  if (self->allocator)
    gst_object_unref (self->allocator);
  self->allocator = gst_fluc_allocator_new (self->ctx, &info, fmt);
  gst_query_add_allocation_param (query, self->allocator, NULL);
  g_mutex_unlock (&self->lock);
  /* pool frames have 256-byte multiple strides: buffers must carry GstVideoMeta */
  gst_query_add_allocation_meta (query, GST_VIDEO_META_API_TYPE, NULL);
  return TRUE;
}

/* auto-register=true: memory we pinned behind upstream's back must be unpinned before upstream
 * frees it -- a later allocation at the same address would otherwise be taken for the pinned
 * one. The GstMemory tells us when it goes. */
typedef struct
{
  gpointer data;
  gsize size;
  guint generation;
  FlucTtmlBlend *ctx;
} PinnedNote;

static void
pinned_memory_gone (gpointer user_data, GstMiniObject * where_the_object_was)
{
  PinnedNote *n = user_data;
  (void) where_the_object_was;
  g_mutex_lock (&shared_lock);
  if (shared_multi && n->generation == shared_generation)
    fluc_ttmlblend_host_forget (n->ctx, n->data, n->size);
  g_mutex_unlock (&shared_lock);
  g_free (n);
}

static void
watch_pinned_memory (GstTTMLBlend * self, GstBuffer * buf, GstVideoFrame * frame)
{
  static GQuark quark = 0;
  guint i;
  if (!quark)
    quark = g_quark_from_static_string ("fluc-ttmlblend-pinned");
  for (i = 0; i < gst_buffer_n_memory (buf); i++) {
    GstMemory *mem = gst_buffer_peek_memory (buf, i);
    PinnedNote *n;
    if (gst_is_fluc_memory (mem) || gst_mini_object_get_qdata (GST_MINI_OBJECT_CAST (mem), quark))
      continue;
    n = g_new0 (PinnedNote, 1);
    /* one memory per frame is what raw video buffers are; the mapped frame tells where it is */
    n->data = GST_VIDEO_FRAME_PLANE_DATA (frame, 0);
    n->size = mem->size;
    n->ctx = self->ctx;
    g_mutex_lock (&shared_lock);
    n->generation = shared_generation;
    g_mutex_unlock (&shared_lock);
    gst_mini_object_set_qdata (GST_MINI_OBJECT_CAST (mem), quark, n, NULL);
    gst_mini_object_weak_ref (GST_MINI_OBJECT_CAST (mem), pinned_memory_gone, n);
  }
}

/* video streaming thread: gst_video_overlay_composition_blend (comp, frame) */
static GstFlowReturn
gst_ttmlblend_transform_ip (GstBaseTransform * trans, GstBuffer * buf)
{
  GstTTMLBlend *self = GST_TTMLBLEND (trans);
  GstClockTime rt;
  GstVideoFrame frame;
  FlucTtmlBlendFrame f = { {NULL, NULL, NULL}, {0, 0, 0} };
  guint64 ticket = 0;
  gboolean active;
  guint32 stream = 0;
  guint p;
  int rc;

  if (!self->have_vinfo || !self->ctx)
    return GST_FLOW_NOT_NEGOTIATED;

  g_mutex_lock (&self->lock);
  rt = GST_BUFFER_PTS_IS_VALID (buf) ?
      gst_segment_to_running_time (&self->video_segment, GST_FORMAT_TIME, GST_BUFFER_PTS (buf)) : GST_CLOCK_TIME_NONE;
  if (GST_CLOCK_TIME_IS_VALID (rt))
    advance_cues (self, rt);
  /* a frame without a usable timestamp shows whatever cue is current */
  active = self->have_current && !self->current.clear &&
      (!GST_CLOCK_TIME_IS_VALID (rt) || (rt >= self->current.start &&
          (!GST_CLOCK_TIME_IS_VALID (self->current.stop) || rt < self->current.stop)));
  if (self->have_current && !self->current.clear && GST_CLOCK_TIME_IS_VALID (rt) &&
      GST_CLOCK_TIME_IS_VALID (self->current.stop) && rt >= self->current.stop) {
    /* over: its overlay can go (frames already queued in the library keep their reference) */
    release_slot (self, self->current.slot);
    self->have_current = FALSE;
  }
  if (active)
    stream = slot_stream (self, self->current.slot);
  if (active)
    self->frames_blended++;
  else
    self->frames_passed++;
  g_mutex_unlock (&self->lock);
  if (!active)
    return GST_FLOW_OK;

  if (!gst_video_frame_map (&frame, &self->vinfo, buf, GST_MAP_READWRITE))
    return GST_FLOW_ERROR;
  for (p = 0; p < GST_VIDEO_FRAME_N_PLANES (&frame); p++) {
    f.plane[p] = GST_VIDEO_FRAME_PLANE_DATA (&frame, p);
    f.stride[p] = GST_VIDEO_FRAME_PLANE_STRIDE (&frame, p);
  }
  if (self->auto_register)
    watch_pinned_memory (self, buf, &frame);
  rc = fluc_ttmlblend_blend_host (self->ctx, stream,
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
  self->ctx = shared_context_acquire (self->device, self->element_id, &rc);
  if (!self->ctx) {
    GST_ELEMENT_ERROR (self, LIBRARY, INIT, ("no usable CUDA device"),
        ("%s", fluc_ttmlblend_strerror (rc)));
    return FALSE;
  }
  /* off unless asked for: pinning memory we do not own is the owner's call (auto-register) */
  if (self->auto_register)
    fluc_ttmlblend_set_auto_register (self->ctx, 1);
  g_mutex_lock (&self->lock);
  self->n_queued = 0;
  self->have_current = FALSE;
  self->subtitle_flushing = FALSE;
  gst_segment_init (&self->video_segment, GST_FORMAT_TIME);
  gst_segment_init (&self->subtitle_segment, GST_FORMAT_TIME);
  g_mutex_unlock (&self->lock);
  return TRUE;
}

static gboolean
gst_ttmlblend_stop (GstBaseTransform * trans)
{
  GstTTMLBlend *self = GST_TTMLBLEND (trans);
  GstAllocator *allocator;
  g_mutex_lock (&self->lock);
  self->subtitle_flushing = TRUE;       /* a blocked subtitle thread lets go */
  if (self->ctx) {
    /* the context lives on for the other elements: drop only this element's overlays */
    drop_all_cues (self);
    fluc_ttmlblend_sync (self->ctx);
  }
  allocator = self->allocator;
  self->allocator = NULL;
  g_cond_broadcast (&self->cond);
  g_mutex_unlock (&self->lock);
  /* memories handed out keep the allocator (and through it the pool frames) alive; the
   * context itself must outlive them: an element is stopped after upstream has released its
   * buffers (pool deactivation on the way to READY) */
  if (allocator)
    gst_object_unref (allocator);
  g_mutex_lock (&self->lock);
  if (self->ctx)
    shared_context_release ();
  self->ctx = NULL;
  g_mutex_unlock (&self->lock);
  return TRUE;
}

/* for tests and monitoring: the context's counters plus the element's own frame counts */
void
gst_ttmlblend_get_stats (GstElement * element, FlucTtmlBlendStats * stats, guint64 * frames_blended,
    guint64 * frames_passed)
{
  GstTTMLBlend *self = GST_TTMLBLEND (element);
  g_mutex_lock (&self->lock);
  if (stats && self->ctx)
    fluc_ttmlblend_stats_copy (self->ctx, stats);
  if (frames_blended)
    *frames_blended = self->frames_blended;
  if (frames_passed)
    *frames_passed = self->frames_passed;
  g_mutex_unlock (&self->lock);
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
    case PROP_AUTO_REGISTER:
      self->auto_register = g_value_get_boolean (value);
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
    case PROP_AUTO_REGISTER:
      g_value_set_boolean (value, self->auto_register);
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
  g_cond_clear (&self->cond);
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
          "CUDA device index (-1: spread the elements over all GPUs, element id % n)", -1, 64, -1,
          G_PARAM_READWRITE | G_PARAM_STATIC_STRINGS));
  g_object_class_install_property (gobject_class, PROP_AUTO_REGISTER,
      g_param_spec_boolean ("auto-register", "Pin upstream memory",
          "Pin video memory that does not come from the proposed allocator the first time it is "
          "seen (unpinned when the GstMemory is finalised); off: such frames are staged",
          FALSE, G_PARAM_READWRITE | G_PARAM_STATIC_STRINGS));

  gst_element_class_add_static_pad_template (element_class, &video_sink_template);
  gst_element_class_add_static_pad_template (element_class, &video_src_template);
  gst_element_class_add_static_pad_template (element_class, &subtitle_sink_template);
  gst_element_class_set_static_metadata (element_class, "TTML blender (CUDA)",
      "Filter/Editor/Video/Overlay/Subtitle",
      "Blends ttmlrender's BGRA cue images onto raw video on an NVIDIA B200",
      "flu-plugins-oss_b200");

  bt_class->set_caps = GST_DEBUG_FUNCPTR (gst_ttmlblend_set_caps);
  bt_class->sink_event = GST_DEBUG_FUNCPTR (gst_ttmlblend_sink_event);
  bt_class->propose_allocation = GST_DEBUG_FUNCPTR (gst_ttmlblend_propose_allocation);
  bt_class->transform_ip = GST_DEBUG_FUNCPTR (gst_ttmlblend_transform_ip);
  bt_class->start = GST_DEBUG_FUNCPTR (gst_ttmlblend_start);
  bt_class->stop = GST_DEBUG_FUNCPTR (gst_ttmlblend_stop);
  GST_DEBUG_CATEGORY_INIT (ttmlblend_debug, "ttmlblend", 0, "TTML CUDA blender");
}

static void
gst_ttmlblend_init (GstTTMLBlend * self)
{
  g_mutex_init (&self->lock);
  g_cond_init (&self->cond);
  self->device = -1;
  self->auto_register = FALSE;
  self->element_id = (guint32) g_atomic_int_add (&next_element_id, 1);
  gst_segment_init (&self->video_segment, GST_FORMAT_TIME);
  gst_segment_init (&self->subtitle_segment, GST_FORMAT_TIME);
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
