/*
 * gstfake.c -- the behaviour behind tests/gst_stub/gst/: a minimal functional fake of the
 * GLib / GObject / GStreamer calls flu-plugins-oss_b200/gst/gstttmlblend.c and
 * gstflucallocator.c make, so that the glue is COMPILED AND RUN in an image that has no
 * GStreamer. Test infrastructure: type registration with class inheritance, properties,
 * mini objects with weak references, buffers / memories / allocators, GstVideoInfo layouts as
 * gst_video_info_set_format computes them, segments and running time, pads that call the chain
 * and event functions. Nothing here is product code and nothing is copied from GStreamer.
 */
#include <gst/gst.h>
#include <gst/base/gstbasetransform.h>
#include <gst/video/video.h>

/* ---- GLib ---------------------------------------------------------------- */
void g_mutex_init (GMutex *m) { pthread_mutex_init (&m->m, NULL); }
void g_mutex_clear (GMutex *m) { pthread_mutex_destroy (&m->m); }
void g_mutex_lock (GMutex *m) { pthread_mutex_lock (&m->m); }
void g_mutex_unlock (GMutex *m) { pthread_mutex_unlock (&m->m); }
void g_cond_init (GCond *c) { pthread_cond_init (&c->c, NULL); }
void g_cond_clear (GCond *c) { pthread_cond_destroy (&c->c); }
void g_cond_wait (GCond *c, GMutex *m) { pthread_cond_wait (&c->c, &m->m); }
void g_cond_signal (GCond *c) { pthread_cond_signal (&c->c); }
void g_cond_broadcast (GCond *c) { pthread_cond_broadcast (&c->c); }
gpointer g_malloc (gsize n) { gpointer p = malloc (n ? n : 1); if (!p) abort (); return p; }
gpointer g_malloc0 (gsize n) { gpointer p = calloc (1, n ? n : 1); if (!p) abort (); return p; }
void g_free (gpointer p) { free (p); }
gpointer g_memdup2 (gconstpointer p, gsize n) { gpointer q = g_malloc (n); memcpy (q, p, n); return q; }
gint g_atomic_int_add (volatile gint *atomic, gint val) { return __sync_fetch_and_add (atomic, val); }

GQuark
g_quark_from_static_string (const gchar *s)
{
  static const gchar *known[64];
  static guint n = 0;
  for (guint i = 0; i < n; i++)
    if (strcmp (known[i], s) == 0)
      return i + 1;
  known[n++] = s;
  return n;
}

/* ---- types ------------------------------------------------------------------ */
typedef struct {
  const gchar *name;
  GType parent;
  gsize class_size, instance_size;
  void (*instance_init) (gpointer);
  gpointer klass;
} FakeType;
static FakeType types[64];
static guint n_types = 1;       /* GType 0 = invalid */

GType
g_fake_type_register (const gchar *name, GType parent, gsize class_size, gsize instance_size,
    void (*class_init) (gpointer), void (*instance_init) (gpointer), gpointer *parent_class)
{
  FakeType *t = &types[n_types];
  t->name = name;
  t->parent = parent;
  t->class_size = class_size;
  t->instance_size = instance_size;
  t->instance_init = instance_init;
  t->klass = g_malloc0 (class_size);
  if (parent) {
    /* a class starts as a copy of its parent's: inherited virtual functions and properties */
    memcpy (t->klass, types[parent].klass, types[parent].class_size);
    if (parent_class)
      *parent_class = types[parent].klass;
  }
  ((GTypeClass *) t->klass)->g_type = n_types;
  const GType id = n_types++;
  if (class_init)
    class_init (t->klass);
  return id;
}

gpointer g_fake_type_class (GType t) { return types[t].klass; }

static void
init_chain (GType t, gpointer obj)
{
  if (!t)
    return;
  init_chain (types[t].parent, obj);
  if (types[t].instance_init)
    types[t].instance_init (obj);
}

GType
g_object_get_type (void)
{
  static GType t = 0;
  if (!t)
    t = g_fake_type_register ("GObject", 0, sizeof (GObjectClass), sizeof (GObject), NULL, NULL, NULL);
  return t;
}

gpointer
g_object_new (GType type, const gchar *first, ...)
{
  (void) first;                 /* the glue only uses g_object_new (TYPE, NULL) */
  GObject *o = g_malloc0 (types[type].instance_size);
  o->g_type_instance.g_class = types[type].klass;
  o->ref_count = 1;
  init_chain (type, o);
  return o;
}

gpointer g_object_ref (gpointer o) { __sync_fetch_and_add (&((GObject *) o)->ref_count, 1); return o; }

void
g_object_unref (gpointer o)
{
  GObject *obj = o;
  if (__sync_sub_and_fetch (&obj->ref_count, 1) == 0) {
    GObjectClass *k = G_OBJECT_GET_CLASS (obj);
    if (k->dispose)
      k->dispose (obj);
    if (k->finalize)
      k->finalize (obj);
    g_free (obj);
  }
}

static void
base_finalize (GObject *o)
{
  (void) o;
}

GParamSpec *
g_param_spec_int (const gchar *name, const gchar *nick, const gchar *blurb, gint mn, gint mx, gint def, GParamFlags f)
{
  (void) nick; (void) blurb; (void) f;
  GParamSpec *p = g_new0 (GParamSpec, 1);
  p->name = name; p->minimum = mn; p->maximum = mx; p->def = def;
  return p;
}

GParamSpec *
g_param_spec_boolean (const gchar *name, const gchar *nick, const gchar *blurb, gboolean def, GParamFlags f)
{
  GParamSpec *p = g_param_spec_int (name, nick, blurb, 0, 1, def, f);
  p->is_bool = TRUE;
  return p;
}

void
g_object_class_install_property (GObjectClass *k, guint id, GParamSpec *p)
{
  p->id = id;
  p->next = k->pspecs;
  k->pspecs = p;
}

gint g_value_get_int (const GValue *v) { return v->v_int; }
void g_value_set_int (GValue *v, gint i) { v->v_int = i; }
gboolean g_value_get_boolean (const GValue *v) { return v->v_bool; }
void g_value_set_boolean (GValue *v, gboolean b) { v->v_bool = b; }

static GParamSpec *
find_pspec (gpointer o, const gchar *name)
{
  for (GParamSpec *p = G_OBJECT_GET_CLASS (o)->pspecs; p; p = p->next)
    if (strcmp (p->name, name) == 0)
      return p;
  return NULL;
}

gboolean
g_fake_object_set_int (gpointer o, const gchar *name, gint v)
{
  GParamSpec *p = find_pspec (o, name);
  if (!p || v < p->minimum || v > p->maximum)
    return FALSE;
  GValue val = { v, v != 0 };
  G_OBJECT_GET_CLASS (o)->set_property (o, p->id, &val, p);
  return TRUE;
}

gboolean
g_fake_object_get_int (gpointer o, const gchar *name, gint *v)
{
  GParamSpec *p = find_pspec (o, name);
  if (!p)
    return FALSE;
  GValue val = { 0, 0 };
  G_OBJECT_GET_CLASS (o)->get_property (o, p->id, &val, p);
  *v = p->is_bool ? val.v_bool : val.v_int;
  return TRUE;
}

/* ---- mini objects ------------------------------------------------------------ */
void
gst_mini_object_weak_ref (GstMiniObject *o, GstMiniObjectNotify notify, gpointer data)
{
  g_assert (o->n_weak < G_N_ELEMENTS (o->weak));
  o->weak[o->n_weak].notify = notify;
  o->weak[o->n_weak].data = data;
  o->n_weak++;
}

void
gst_mini_object_set_qdata (GstMiniObject *o, GQuark q, gpointer data, GDestroyNotify destroy)
{
  for (guint i = 0; i < o->n_qdata; i++)
    if (o->qdata[i].quark == q) {
      if (o->qdata[i].destroy)
        o->qdata[i].destroy (o->qdata[i].data);
      o->qdata[i].data = data;
      o->qdata[i].destroy = destroy;
      return;
    }
  g_assert (o->n_qdata < G_N_ELEMENTS (o->qdata));
  o->qdata[o->n_qdata].quark = q;
  o->qdata[o->n_qdata].data = data;
  o->qdata[o->n_qdata].destroy = destroy;
  o->n_qdata++;
}

gpointer
gst_mini_object_get_qdata (GstMiniObject *o, GQuark q)
{
  for (guint i = 0; i < o->n_qdata; i++)
    if (o->qdata[i].quark == q)
      return o->qdata[i].data;
  return NULL;
}

GstMiniObject *gst_mini_object_ref (GstMiniObject *o) { __sync_fetch_and_add (&o->refcount, 1); return o; }

void
gst_mini_object_unref (GstMiniObject *o)
{
  if (__sync_sub_and_fetch (&o->refcount, 1) != 0)
    return;
  /* weak references and qdata go first, while the object can still be looked at */
  for (guint i = 0; i < o->n_weak; i++)
    o->weak[i].notify (o->weak[i].data, o);
  for (guint i = 0; i < o->n_qdata; i++)
    if (o->qdata[i].destroy)
      o->qdata[i].destroy (o->qdata[i].data);
  if (o->free)
    o->free (o);
}

static void
mini_init (GstMiniObject *o, void (*free_fn) (GstMiniObject *))
{
  memset (o, 0, sizeof *o);
  o->refcount = 1;
  o->free = free_fn;
}

/* ---- GstObject ------------------------------------------------------------- */
static void
gst_object_class_init (gpointer k)
{
  ((GObjectClass *) k)->finalize = base_finalize;
}

GType
gst_object_get_type (void)
{
  static GType t = 0;
  if (!t)
    t = g_fake_type_register ("GstObject", G_TYPE_OBJECT, sizeof (GstObjectClass), sizeof (GstObject),
        gst_object_class_init, NULL, NULL);
  return t;
}

gpointer gst_object_ref_sink (gpointer o) { return o; }     /* floating reference becomes ours */
gpointer gst_object_ref (gpointer o) { return g_object_ref (o); }
void gst_object_unref (gpointer o) { g_object_unref (o); }

/* ---- caps --------------------------------------------------------------------- */
static void caps_free (GstMiniObject *o) { g_free (o); }

GstCaps *
gst_fake_video_caps_new (const gchar *format, gint width, gint height)
{
  GstCaps *c = g_new0 (GstCaps, 1);
  mini_init (&c->mini, caps_free);
  snprintf (c->format, sizeof c->format, "%s", format);
  c->width = width;
  c->height = height;
  return c;
}

GstCaps *gst_caps_ref (GstCaps *c) { gst_mini_object_ref (&c->mini); return c; }
void gst_caps_unref (GstCaps *c) { gst_mini_object_unref (&c->mini); }

/* ---- segments, events ------------------------------------------------------------ */
void
gst_segment_init (GstSegment *s, GstFormat f)
{
  memset (s, 0, sizeof *s);
  s->rate = s->applied_rate = 1.0;
  s->format = f;
  s->stop = s->position = s->duration = (guint64) -1;
}

/* rate 1.0 only (what the harness sends): running time = position - start + base - offset,
 * -1 outside the segment */
guint64
gst_segment_to_running_time (const GstSegment *s, GstFormat f, guint64 position)
{
  if (f != s->format || position == (guint64) -1)
    return (guint64) -1;
  if (position < s->start || (s->stop != (guint64) -1 && position > s->stop))
    return (guint64) -1;
  const guint64 r = position - s->start;
  if (r + s->base < s->offset)
    return (guint64) -1;
  return r + s->base - s->offset;
}

static void
event_free (GstMiniObject *o)
{
  GstEvent *e = (GstEvent *) o;
  if (e->caps)
    gst_caps_unref (e->caps);
  g_free (e);
}

static GstEvent *
event_new (GstEventType t)
{
  GstEvent *e = g_new0 (GstEvent, 1);
  mini_init (&e->mini, event_free);
  e->type = t;
  return e;
}

GstEvent *gst_event_new_segment (const GstSegment *s) { GstEvent *e = event_new (GST_EVENT_SEGMENT); e->segment = *s; return e; }
GstEvent *gst_event_new_gap (GstClockTime ts, GstClockTime d) { GstEvent *e = event_new (GST_EVENT_GAP); e->gap_ts = ts; e->gap_duration = d; return e; }
GstEvent *gst_event_new_eos (void) { return event_new (GST_EVENT_EOS); }
GstEvent *gst_event_new_flush_start (void) { return event_new (GST_EVENT_FLUSH_START); }
GstEvent *gst_event_new_flush_stop (gboolean reset) { (void) reset; return event_new (GST_EVENT_FLUSH_STOP); }
GstEvent *gst_event_new_caps (GstCaps *c) { GstEvent *e = event_new (GST_EVENT_CAPS); e->caps = gst_caps_ref (c); return e; }
void gst_event_parse_segment (GstEvent *e, const GstSegment **s) { *s = &e->segment; }
void gst_event_copy_segment (GstEvent *e, GstSegment *s) { *s = e->segment; }
void gst_event_parse_gap (GstEvent *e, GstClockTime *ts, GstClockTime *d) { if (ts) *ts = e->gap_ts; if (d) *d = e->gap_duration; }
void gst_event_parse_caps (GstEvent *e, GstCaps **c) { *c = e->caps; }
void gst_event_unref (GstEvent *e) { gst_mini_object_unref (&e->mini); }

/* ---- memory / allocator / buffer --------------------------------------------------- */
static void
memory_free (GstMiniObject *o)
{
  GstMemory *m = (GstMemory *) o;
  if (m->allocator) {
    GstAllocator *a = m->allocator;
    GST_ALLOCATOR_GET_CLASS (a)->free (a, m);   /* the allocator frees the structure it allocated */
    gst_object_unref (a);
    return;
  }
  if (m->fake_owned)
    g_free (m->fake_data);
  g_free (m);
}

void
gst_memory_init (GstMemory *m, guint flags, GstAllocator *allocator, GstMemory *parent, gsize maxsize, gsize align,
    gsize offset, gsize size)
{
  (void) flags;
  mini_init (&m->mini_object, memory_free);
  m->allocator = allocator ? gst_object_ref (allocator) : NULL;
  m->parent = parent;
  m->maxsize = maxsize;
  m->align = align;
  m->offset = offset;
  m->size = size;
}

GType
gst_allocator_get_type (void)
{
  static GType t = 0;
  if (!t)
    t = g_fake_type_register ("GstAllocator", GST_TYPE_OBJECT, sizeof (GstAllocatorClass), sizeof (GstAllocator),
        NULL, NULL, NULL);
  return t;
}

GstMemory *
gst_allocator_alloc (GstAllocator *a, gsize size, GstAllocationParams *params)
{
  if (a)
    return GST_ALLOCATOR_GET_CLASS (a)->alloc (a, size, params);
  GstMemory *m = g_new0 (GstMemory, 1);
  gst_memory_init (m, 0, NULL, NULL, size, 0, 0, size);
  m->fake_data = g_malloc0 (size);
  m->fake_owned = TRUE;
  return m;
}

GstMemory *gst_memory_ref (GstMemory *m) { gst_mini_object_ref (&m->mini_object); return m; }
void gst_memory_unref (GstMemory *m) { gst_mini_object_unref (&m->mini_object); }

static void
buffer_free (GstMiniObject *o)
{
  GstBuffer *b = (GstBuffer *) o;
  for (guint i = 0; i < b->n_mem; i++)
    gst_memory_unref (b->mem[i]);
  g_free (b->video_meta);
  g_free (b);
}

GstBuffer *
gst_buffer_new (void)
{
  GstBuffer *b = g_new0 (GstBuffer, 1);
  mini_init (&b->mini_object, buffer_free);
  b->pts = b->dts = b->duration = GST_CLOCK_TIME_NONE;
  return b;
}

void gst_buffer_append_memory (GstBuffer *b, GstMemory *m) { g_assert (b->n_mem < 4); b->mem[b->n_mem++] = m; }

GstBuffer *
gst_buffer_new_allocate (GstAllocator *a, gsize size, GstAllocationParams *params)
{
  GstMemory *m = gst_allocator_alloc (a, size, params);
  if (!m)
    return NULL;
  GstBuffer *b = gst_buffer_new ();
  gst_buffer_append_memory (b, m);
  return b;
}

typedef struct { gpointer user_data; GDestroyNotify notify; } WrapNotify;

GstBuffer *
gst_buffer_new_wrapped_full (guint flags, gpointer data, gsize maxsize, gsize offset, gsize size, gpointer user_data,
    GDestroyNotify notify)
{
  (void) flags;
  GstMemory *m = g_new0 (GstMemory, 1);
  gst_memory_init (m, 0, NULL, NULL, maxsize, 0, offset, size);
  m->fake_data = data;
  m->fake_owned = FALSE;
  if (notify)
    gst_mini_object_set_qdata (&m->mini_object, g_quark_from_static_string ("fake-wrap-notify"), user_data, notify);
  GstBuffer *b = gst_buffer_new ();
  gst_buffer_append_memory (b, m);
  return b;
}

GstBuffer *
gst_buffer_new_wrapped (gpointer data, gsize size)
{
  return gst_buffer_new_wrapped_full (0, data, size, 0, size, data, g_free);
}

guint gst_buffer_n_memory (GstBuffer *b) { return b->n_mem; }
GstMemory *gst_buffer_peek_memory (GstBuffer *b, guint idx) { return idx < b->n_mem ? b->mem[idx] : NULL; }

gboolean
gst_buffer_map (GstBuffer *b, GstMapInfo *info, GstMapFlags flags)
{
  if (b->n_mem != 1)
    return FALSE;               /* the fake never merges memories */
  GstMemory *m = b->mem[0];
  guint8 *p = m->allocator ? m->allocator->mem_map (m, m->maxsize, flags) : m->fake_data;
  if (!p)
    return FALSE;
  info->memory = m;
  info->flags = flags;
  info->data = p + m->offset;
  info->size = m->size;
  info->maxsize = m->maxsize - m->offset;
  return TRUE;
}

void
gst_buffer_unmap (GstBuffer *b, GstMapInfo *info)
{
  (void) b;
  if (info->memory && info->memory->allocator)
    info->memory->allocator->mem_unmap (info->memory);
}

GstBuffer *gst_buffer_ref (GstBuffer *b) { gst_mini_object_ref (&b->mini_object); return b; }
void gst_buffer_unref (GstBuffer *b) { gst_mini_object_unref (&b->mini_object); }

/* ---- queries ----------------------------------------------------------------------- */
void
gst_query_parse_allocation (GstQuery *q, GstCaps **caps, gboolean *need_pool)
{
  if (caps)
    *caps = q->caps;
  if (need_pool)
    *need_pool = FALSE;
}

void
gst_query_add_allocation_param (GstQuery *q, GstAllocator *a, const GstAllocationParams *p)
{
  (void) p;
  if (!q->allocator && a)
    q->allocator = gst_object_ref (a);
}

void
gst_query_add_allocation_meta (GstQuery *q, GType api, const GstStructure *s)
{
  (void) s;
  if (api == GST_VIDEO_META_API_TYPE)
    q->has_video_meta = TRUE;
}

/* ---- video ----------------------------------------------------------------------------- */
static const struct { GstVideoFormat f; const gchar *name; } format_names[] = {
  { GST_VIDEO_FORMAT_I420, "I420" }, { GST_VIDEO_FORMAT_YV12, "YV12" }, { GST_VIDEO_FORMAT_NV12, "NV12" },
  { GST_VIDEO_FORMAT_NV21, "NV21" }, { GST_VIDEO_FORMAT_AYUV, "AYUV" }, { GST_VIDEO_FORMAT_ARGB, "ARGB" },
  { GST_VIDEO_FORMAT_ABGR, "ABGR" }, { GST_VIDEO_FORMAT_RGBA, "RGBA" }, { GST_VIDEO_FORMAT_BGRA, "BGRA" },
  { GST_VIDEO_FORMAT_RGBx, "RGBx" }, { GST_VIDEO_FORMAT_BGRx, "BGRx" }, { GST_VIDEO_FORMAT_xRGB, "xRGB" },
  { GST_VIDEO_FORMAT_xBGR, "xBGR" }, { GST_VIDEO_FORMAT_Y42B, "Y42B" }, { GST_VIDEO_FORMAT_Y444, "Y444" },
  { GST_VIDEO_FORMAT_YUY2, "YUY2" }, { GST_VIDEO_FORMAT_UYVY, "UYVY" }, { GST_VIDEO_FORMAT_GRAY8, "GRAY8" },
  { GST_VIDEO_FORMAT_NV16, "NV16" }, { GST_VIDEO_FORMAT_NV24, "NV24" }, { GST_VIDEO_FORMAT_NV61, "NV61" },
  { GST_VIDEO_FORMAT_YVYU, "YVYU" }, { GST_VIDEO_FORMAT_VYUY, "VYUY" }, { GST_VIDEO_FORMAT_v308, "v308" },
  { GST_VIDEO_FORMAT_IYU2, "IYU2" }, { GST_VIDEO_FORMAT_RGB, "RGB" }, { GST_VIDEO_FORMAT_BGR, "BGR" },
};

GstVideoFormat
gst_video_format_from_string (const gchar *s)
{
  for (guint i = 0; i < G_N_ELEMENTS (format_names); i++)
    if (strcmp (format_names[i].name, s) == 0)
      return format_names[i].f;
  return GST_VIDEO_FORMAT_UNKNOWN;
}

const gchar *
gst_video_format_to_string (GstVideoFormat f)
{
  for (guint i = 0; i < G_N_ELEMENTS (format_names); i++)
    if (format_names[i].f == f)
      return format_names[i].name;
  return "UNKNOWN";
}

#define RUP2(x) (((x) + 1) & ~1)
#define RUP4(x) (((x) + 3) & ~3)

/* plane strides / offsets of GStreamer's default layouts (fill_planes of video-info.c, restated):
 * 4-byte aligned strides, chroma planes of 4:2:0 with ROUND_UP_2 (height) / 2 rows */
gboolean
gst_video_info_set_format (GstVideoInfo *i, GstVideoFormat f, guint w, guint h)
{
  memset (i, 0, sizeof *i);
  i->format = f;
  i->width = (gint) w;
  i->height = (gint) h;
  switch (f) {
    case GST_VIDEO_FORMAT_I420:
    case GST_VIDEO_FORMAT_YV12:
      i->n_planes = 3;
      i->stride[0] = RUP4 (w);
      i->stride[1] = i->stride[2] = RUP4 (RUP2 (w) / 2);
      i->offset[1] = (gsize) i->stride[0] * RUP2 (h);
      i->offset[2] = i->offset[1] + (gsize) i->stride[1] * (RUP2 (h) / 2);
      i->size = i->offset[2] + (gsize) i->stride[2] * (RUP2 (h) / 2);
      break;
    case GST_VIDEO_FORMAT_NV12:
    case GST_VIDEO_FORMAT_NV21:
      i->n_planes = 2;
      i->stride[0] = i->stride[1] = RUP4 (w);
      i->offset[1] = (gsize) i->stride[0] * RUP2 (h);
      i->size = i->offset[1] + (gsize) i->stride[0] * (RUP2 (h) / 2);
      break;
    case GST_VIDEO_FORMAT_GRAY8:
      i->n_planes = 1;
      i->stride[0] = RUP4 (w);
      i->size = (gsize) i->stride[0] * h;
      break;
    case GST_VIDEO_FORMAT_AYUV: case GST_VIDEO_FORMAT_ARGB: case GST_VIDEO_FORMAT_ABGR: case GST_VIDEO_FORMAT_RGBA:
    case GST_VIDEO_FORMAT_BGRA: case GST_VIDEO_FORMAT_RGBx: case GST_VIDEO_FORMAT_BGRx: case GST_VIDEO_FORMAT_xRGB:
    case GST_VIDEO_FORMAT_xBGR:
      i->n_planes = 1;
      i->stride[0] = (gint) w * 4;
      i->size = (gsize) i->stride[0] * h;
      i->has_alpha = f == GST_VIDEO_FORMAT_AYUV || f == GST_VIDEO_FORMAT_ARGB || f == GST_VIDEO_FORMAT_ABGR ||
          f == GST_VIDEO_FORMAT_RGBA || f == GST_VIDEO_FORMAT_BGRA;
      i->a_poffset = (f == GST_VIDEO_FORMAT_RGBA || f == GST_VIDEO_FORMAT_BGRA) ? 3 : 0;
      break;
    default:
      return FALSE;             /* the fake lays out only what the harness feeds */
  }
  return TRUE;
}

gboolean
gst_video_info_from_caps (GstVideoInfo *i, const GstCaps *c)
{
  const GstVideoFormat f = gst_video_format_from_string (c->format);
  if (f == GST_VIDEO_FORMAT_UNKNOWN || c->width <= 0 || c->height <= 0)
    return FALSE;
  return gst_video_info_set_format (i, f, (guint) c->width, (guint) c->height);
}

GType gst_video_meta_api_get_type (void) { return 63; }

GstVideoMeta *
gst_buffer_add_video_meta_full (GstBuffer *b, GstVideoFrameFlags flags, GstVideoFormat f, guint w, guint h, guint n_planes,
    gsize offset[GST_VIDEO_MAX_PLANES], gint stride[GST_VIDEO_MAX_PLANES])
{
  g_free (b->video_meta);
  b->video_meta = g_new0 (struct _GstVideoMetaFake, 1);
  GstVideoMeta *m = &b->video_meta->meta;
  m->buffer = b; m->flags = flags; m->format = f; m->width = w; m->height = h; m->n_planes = n_planes;
  for (guint p = 0; p < n_planes; p++) {
    m->offset[p] = offset[p];
    m->stride[p] = stride[p];
  }
  return m;
}

GstVideoMeta *
gst_buffer_add_video_meta (GstBuffer *b, GstVideoFrameFlags flags, GstVideoFormat f, guint w, guint h)
{
  GstVideoInfo i;
  if (!gst_video_info_set_format (&i, f, w, h))
    return NULL;
  return gst_buffer_add_video_meta_full (b, flags, f, w, h, i.n_planes, i.offset, i.stride);
}

GstVideoMeta *gst_buffer_get_video_meta (GstBuffer *b) { return b->video_meta ? &b->video_meta->meta : NULL; }

/* like the real one: the buffer's GstVideoMeta, when there is one, overrides the default
 * strides and offsets of `info` */
gboolean
gst_video_frame_map (GstVideoFrame *frame, const GstVideoInfo *info, GstBuffer *buffer, GstMapFlags flags)
{
  memset (frame, 0, sizeof *frame);
  frame->info = *info;
  frame->buffer = buffer;
  GstVideoMeta *meta = gst_buffer_get_video_meta (buffer);
  if (meta) {
    if (meta->format != info->format || (gint) meta->width != info->width || (gint) meta->height != info->height)
      return FALSE;
    for (guint p = 0; p < meta->n_planes; p++) {
      frame->info.offset[p] = meta->offset[p];
      frame->info.stride[p] = meta->stride[p];
    }
  }
  if (!gst_buffer_map (buffer, &frame->map[0], flags))
    return FALSE;
  for (guint p = 0; p < frame->info.n_planes; p++)
    frame->data[p] = frame->map[0].data + frame->info.offset[p];
  return TRUE;
}

void
gst_video_frame_unmap (GstVideoFrame *frame)
{
  gst_buffer_unmap (frame->buffer, &frame->map[0]);
}

/* ---- pads, elements, base transform ---------------------------------------------------- */
GstPad *
gst_pad_new_from_static_template (GstStaticPadTemplate *t, const gchar *name)
{
  GstPad *p = g_new0 (GstPad, 1);
  p->object.object.ref_count = 1;
  p->object.name = name;
  p->direction = t->direction;
  return p;
}

void gst_pad_set_chain_function (GstPad *p, GstPadChainFunction f) { p->chainfunc = f; }
void gst_pad_set_event_function (GstPad *p, GstPadEventFunction f) { p->eventfunc = f; }
GstCaps *gst_pad_get_current_caps (GstPad *p) { return p->current_caps ? gst_caps_ref (p->current_caps) : NULL; }

void
gst_fake_pad_set_caps (GstPad *p, GstCaps *c)
{
  if (p->current_caps)
    gst_caps_unref (p->current_caps);
  p->current_caps = gst_caps_ref (c);
}

GstFlowReturn
gst_fake_pad_chain (GstPad *p, GstBuffer *b)
{
  if (!p->chainfunc) {
    gst_buffer_unref (b);
    return GST_FLOW_ERROR;
  }
  return p->chainfunc (p, p->parent, b);      /* takes ownership of b, like the real call */
}

gboolean
gst_fake_pad_send_event (GstPad *p, GstEvent *e)
{
  if (GST_EVENT_TYPE (e) == GST_EVENT_CAPS) {
    GstCaps *c;
    gst_event_parse_caps (e, &c);
    gst_fake_pad_set_caps (p, c);            /* the pad stores its caps itself */
  }
  if (!p->eventfunc) {
    gst_event_unref (e);
    return FALSE;
  }
  return p->eventfunc (p, p->parent, e);
}

GType
gst_element_get_type (void)
{
  static GType t = 0;
  if (!t)
    t = g_fake_type_register ("GstElement", GST_TYPE_OBJECT, sizeof (GstElementClass), sizeof (GstElement), NULL, NULL, NULL);
  return t;
}

gboolean gst_element_register (GstPlugin *p, const gchar *n, guint r, GType t) { (void) p; (void) n; (void) r; return t != 0; }

gboolean
gst_element_add_pad (GstElement *e, GstPad *p)
{
  g_assert (e->n_pads < G_N_ELEMENTS (e->pads));
  p->parent = GST_OBJECT (e);
  e->pads[e->n_pads++] = p;
  return TRUE;
}

GstPad *
gst_element_get_static_pad (GstElement *e, const gchar *name)
{
  for (guint i = 0; i < e->n_pads; i++)
    if (strcmp (e->pads[i]->object.name, name) == 0)
      return e->pads[i];
  return NULL;
}

void gst_element_class_add_static_pad_template (GstElementClass *k, GstStaticPadTemplate *t) { (void) k; (void) t; }
void gst_element_class_set_static_metadata (GstElementClass *k, const gchar *a, const gchar *b, const gchar *c, const gchar *d)
{ (void) k; (void) a; (void) b; (void) c; (void) d; }

/* the base class's own sink pad: buffers go to transform_ip, events to sink_event */
static GstFlowReturn
base_transform_chain (GstPad *pad, GstObject *parent, GstBuffer *buf)
{
  (void) pad;
  GstBaseTransform *t = GST_BASE_TRANSFORM (parent);
  GstFlowReturn ret = GST_BASE_TRANSFORM_GET_CLASS (t)->transform_ip (t, buf);
  gst_buffer_unref (buf);       /* the harness holds its own reference and looks at the result */
  return ret;
}

static gboolean
base_transform_default_sink_event (GstBaseTransform *t, GstEvent *e)
{
  if (GST_EVENT_TYPE (e) == GST_EVENT_SEGMENT) {
    gst_event_copy_segment (e, &t->segment);
    t->have_segment = TRUE;
  }
  gst_event_unref (e);          /* "pushed downstream" */
  return TRUE;
}

static gboolean
base_transform_event (GstPad *pad, GstObject *parent, GstEvent *e)
{
  (void) pad;
  GstBaseTransform *t = GST_BASE_TRANSFORM (parent);
  return GST_BASE_TRANSFORM_GET_CLASS (t)->sink_event (t, e);
}

static void
base_transform_class_init (gpointer k)
{
  ((GstBaseTransformClass *) k)->sink_event = base_transform_default_sink_event;
}

static void
base_transform_init (gpointer o)
{
  static GstStaticPadTemplate sink_t = GST_STATIC_PAD_TEMPLATE ("sink", GST_PAD_SINK, GST_PAD_ALWAYS, GST_STATIC_CAPS ("ANY"));
  static GstStaticPadTemplate src_t = GST_STATIC_PAD_TEMPLATE ("src", GST_PAD_SRC, GST_PAD_ALWAYS, GST_STATIC_CAPS ("ANY"));
  GstBaseTransform *t = o;
  t->sinkpad = gst_pad_new_from_static_template (&sink_t, "sink");
  gst_pad_set_chain_function (t->sinkpad, base_transform_chain);
  gst_pad_set_event_function (t->sinkpad, base_transform_event);
  gst_element_add_pad (GST_ELEMENT (t), t->sinkpad);
  t->srcpad = gst_pad_new_from_static_template (&src_t, "src");
  gst_element_add_pad (GST_ELEMENT (t), t->srcpad);
  gst_segment_init (&t->segment, GST_FORMAT_TIME);
}

GType
gst_base_transform_get_type (void)
{
  static GType t = 0;
  if (!t)
    t = g_fake_type_register ("GstBaseTransform", GST_TYPE_ELEMENT, sizeof (GstBaseTransformClass), sizeof (GstBaseTransform),
        base_transform_class_init, base_transform_init, NULL);
  return t;
}

void gst_base_transform_set_in_place (GstBaseTransform *t, gboolean v) { t->fake_in_place = v; }

void gst_init (int *argc, char ***argv) { (void) argc; (void) argv; }
const gchar *gst_version_string (void) { return "fake GStreamer (tests/gst_stub)"; }
