/* A minimal FUNCTIONAL fake of the GLib / GObject / GStreamer core API that
 * flu-plugins-oss_b200/gst/ uses (see ../README.md). Signatures follow GStreamer 1.x; the
 * behaviour behind them (gstfake.c) is just enough to run the element and the allocator. */
#ifndef GST_STUB_GST_H
#define GST_STUB_GST_H
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <pthread.h>

/* ---- GLib -------------------------------------------------------------- */
typedef int gboolean; typedef int gint; typedef unsigned int guint; typedef char gchar;
typedef unsigned char guint8; typedef uint32_t guint32; typedef uint64_t guint64; typedef int64_t gint64;
typedef size_t gsize; typedef void *gpointer; typedef const void *gconstpointer; typedef float gfloat;
typedef double gdouble; typedef unsigned long GType; typedef guint64 GstClockTime; typedef guint32 GQuark;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define G_N_ELEMENTS(a) (sizeof (a) / sizeof ((a)[0]))
#define G_GSIZE_FORMAT "zu"
#define G_GUINT64_FORMAT "lu"
#define G_MAXINT 2147483647
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#define G_UNLIKELY(x) (x)
#define G_LIKELY(x) (x)
#define GST_CLOCK_TIME_NONE ((GstClockTime) -1)
#define GST_CLOCK_TIME_IS_VALID(t) (((GstClockTime) (t)) != GST_CLOCK_TIME_NONE)
#define GST_SECOND ((GstClockTime) 1000000000)
#define GST_MSECOND ((GstClockTime) 1000000)
#define GST_TIME_FORMAT "lu"
#define GST_TIME_ARGS(t) ((unsigned long) (t))

typedef struct { pthread_mutex_t m; } GMutex;
typedef struct { pthread_cond_t c; } GCond;
void g_mutex_init (GMutex *m); void g_mutex_clear (GMutex *m); void g_mutex_lock (GMutex *m); void g_mutex_unlock (GMutex *m);
void g_cond_init (GCond *c); void g_cond_clear (GCond *c); void g_cond_wait (GCond *c, GMutex *m);
void g_cond_signal (GCond *c); void g_cond_broadcast (GCond *c);
gpointer g_malloc (gsize n); gpointer g_malloc0 (gsize n); void g_free (gpointer p); gpointer g_memdup2 (gconstpointer p, gsize n);
#define g_new0(type, n) ((type *) g_malloc0 (sizeof (type) * (n)))
#define g_new(type, n) ((type *) g_malloc (sizeof (type) * (n)))
gint g_atomic_int_add (volatile gint *atomic, gint val);
GQuark g_quark_from_static_string (const gchar *s);
typedef void (*GDestroyNotify) (gpointer data);
#define g_return_val_if_fail(expr, val) do { if (!(expr)) return (val); } while (0)
#define g_return_if_fail(expr) do { if (!(expr)) return; } while (0)
#define g_assert(expr) do { if (!(expr)) { fprintf (stderr, "assertion failed: %s\n", #expr); abort (); } } while (0)
#include <stdlib.h>

/* ---- GObject ------------------------------------------------------------ */
typedef struct _GTypeClass { GType g_type; } GTypeClass;
typedef struct _GTypeInstance { GTypeClass *g_class; } GTypeInstance;
typedef struct _GValue { gint v_int; gboolean v_bool; } GValue;
typedef struct _GParamSpec { const gchar *name; guint id; gint minimum, maximum, def; gboolean is_bool; struct _GParamSpec *next; } GParamSpec;
typedef struct _GObject { GTypeInstance g_type_instance; volatile gint ref_count; gpointer qdata; } GObject;
typedef struct _GObjectClass {
  GTypeClass g_type_class;
  void (*set_property) (GObject *, guint, const GValue *, GParamSpec *);
  void (*get_property) (GObject *, guint, GValue *, GParamSpec *);
  void (*dispose) (GObject *);
  void (*finalize) (GObject *);
  GParamSpec *pspecs;           /* fake: the installed properties */
} GObjectClass;
#define G_TYPE_OBJECT (g_object_get_type ())
GType g_object_get_type (void);
#define G_OBJECT(o) ((GObject *) (o))
#define G_OBJECT_CLASS(k) ((GObjectClass *) (k))
#define G_OBJECT_GET_CLASS(o) ((GObjectClass *) ((GTypeInstance *) (o))->g_class)
typedef enum { G_PARAM_READABLE = 1, G_PARAM_WRITABLE = 2, G_PARAM_READWRITE = 3, G_PARAM_STATIC_STRINGS = 0xe0 } GParamFlags;
GParamSpec *g_param_spec_int (const gchar *, const gchar *, const gchar *, gint, gint, gint, GParamFlags);
GParamSpec *g_param_spec_boolean (const gchar *, const gchar *, const gchar *, gboolean, GParamFlags);
void g_object_class_install_property (GObjectClass *, guint, GParamSpec *);
gint g_value_get_int (const GValue *); void g_value_set_int (GValue *, gint);
gboolean g_value_get_boolean (const GValue *); void g_value_set_boolean (GValue *, gboolean);
gpointer g_object_new (GType type, const gchar *first, ...);
gpointer g_object_ref (gpointer o); void g_object_unref (gpointer o);
/* fake-only helpers for the test harness: set / get an int or boolean property by name */
gboolean g_fake_object_set_int (gpointer o, const gchar *name, gint v);
gboolean g_fake_object_get_int (gpointer o, const gchar *name, gint *v);
#define G_OBJECT_WARN_INVALID_PROPERTY_ID(o, id, p) ((void) (o), (void) (id), (void) (p))
GType g_fake_type_register (const gchar *name, GType parent, gsize class_size, gsize instance_size,
    void (*class_init) (gpointer), void (*instance_init) (gpointer), gpointer *parent_class);
gpointer g_fake_type_class (GType t);
#define G_DECLARE_FINAL_TYPE(Name, name, MOD, OBJ, Parent) \
  GType name##_get_type (void); typedef struct _##Name Name; typedef struct { Parent##Class parent_class; } Name##Class; \
  static inline Name *MOD##_##OBJ (gpointer p) { return (Name *) p; }
#define G_DEFINE_TYPE(Name, name, PARENT_TYPE) \
  static void name##_class_init (Name##Class *klass); static void name##_init (Name *self); \
  static gpointer name##_parent_class = NULL; \
  GType name##_get_type (void) { static GType t = 0; \
    if (!t) t = g_fake_type_register (#Name, PARENT_TYPE, sizeof (Name##Class), sizeof (Name), \
        (void (*) (gpointer)) name##_class_init, (void (*) (gpointer)) name##_init, &name##_parent_class); \
    return t; }

/* ---- GstMiniObject / GstObject -------------------------------------------- */
typedef struct _GstMiniObject GstMiniObject;
typedef void (*GstMiniObjectNotify) (gpointer user_data, GstMiniObject *where_the_object_was);
struct _GstMiniObject { GType type; volatile gint refcount; void (*free) (GstMiniObject *);
  struct { GQuark quark; gpointer data; GDestroyNotify destroy; } qdata[4]; guint n_qdata;
  struct { GstMiniObjectNotify notify; gpointer data; } weak[4]; guint n_weak; };
#define GST_MINI_OBJECT_CAST(o) ((GstMiniObject *) (o))
#define GST_MINI_OBJECT(o) ((GstMiniObject *) (o))
void gst_mini_object_weak_ref (GstMiniObject *, GstMiniObjectNotify, gpointer);
void gst_mini_object_set_qdata (GstMiniObject *, GQuark, gpointer, GDestroyNotify);
gpointer gst_mini_object_get_qdata (GstMiniObject *, GQuark);
GstMiniObject *gst_mini_object_ref (GstMiniObject *); void gst_mini_object_unref (GstMiniObject *);

typedef struct _GstObject { GObject object; guint flags; const gchar *name; } GstObject;
typedef struct _GstObjectClass { GObjectClass parent_class; } GstObjectClass;
#define GST_TYPE_OBJECT (gst_object_get_type ())
GType gst_object_get_type (void);
#define GST_OBJECT(o) ((GstObject *) (o))
#define GST_OBJECT_FLAG_SET(o, f) (((GstObject *) (o))->flags |= (f))
gpointer gst_object_ref_sink (gpointer o); gpointer gst_object_ref (gpointer o); void gst_object_unref (gpointer o);

typedef struct _GstPlugin GstPlugin; typedef struct _GstStructure GstStructure;
/* fake caps: what video/x-raw caps carry for this element */
typedef struct _GstCaps { GstMiniObject mini; gchar format[16]; gint width, height; } GstCaps;
GstCaps *gst_caps_ref (GstCaps *); void gst_caps_unref (GstCaps *);
GstCaps *gst_fake_video_caps_new (const gchar *format, gint width, gint height);

/* ---- segments, events ----------------------------------------------------- */
typedef enum { GST_FORMAT_UNDEFINED = 0, GST_FORMAT_TIME = 3 } GstFormat;
typedef struct _GstSegment { guint flags; gdouble rate, applied_rate; GstFormat format; guint64 base, offset, start, stop, time,
  position, duration; } GstSegment;
void gst_segment_init (GstSegment *, GstFormat);
guint64 gst_segment_to_running_time (const GstSegment *, GstFormat, guint64 position);
typedef enum { GST_EVENT_UNKNOWN = 0, GST_EVENT_FLUSH_START = 1, GST_EVENT_FLUSH_STOP = 2, GST_EVENT_CAPS = 3,
  GST_EVENT_SEGMENT = 4, GST_EVENT_EOS = 5, GST_EVENT_GAP = 6 } GstEventType;
typedef struct _GstEvent { GstMiniObject mini; GstEventType type; GstSegment segment; GstClockTime gap_ts, gap_duration;
  GstCaps *caps; } GstEvent;
#define GST_EVENT_TYPE(e) ((e)->type)
GstEvent *gst_event_new_segment (const GstSegment *); GstEvent *gst_event_new_gap (GstClockTime, GstClockTime);
GstEvent *gst_event_new_eos (void); GstEvent *gst_event_new_flush_start (void); GstEvent *gst_event_new_flush_stop (gboolean);
GstEvent *gst_event_new_caps (GstCaps *);
void gst_event_parse_segment (GstEvent *, const GstSegment **); void gst_event_copy_segment (GstEvent *, GstSegment *);
void gst_event_parse_gap (GstEvent *, GstClockTime *, GstClockTime *); void gst_event_parse_caps (GstEvent *, GstCaps **);
void gst_event_unref (GstEvent *);

/* ---- memory, allocator, buffer --------------------------------------------- */
typedef struct _GstAllocator GstAllocator; typedef struct _GstMemory GstMemory;
typedef struct _GstAllocationParams { guint flags; gsize align, prefix, padding; } GstAllocationParams;
typedef enum { GST_MAP_READ = 1, GST_MAP_WRITE = 2, GST_MAP_READWRITE = 3 } GstMapFlags;
struct _GstMemory { GstMiniObject mini_object; GstAllocator *allocator; GstMemory *parent; gsize maxsize, align, offset, size;
  gpointer fake_data; /* fake: system memory payload (allocator == NULL) */ gboolean fake_owned; };
#define GST_MEMORY_CAST(m) ((GstMemory *) (m))
struct _GstAllocator { GstObject object; const gchar *mem_type;
  gpointer (*mem_map) (GstMemory *, gsize, GstMapFlags); void (*mem_unmap) (GstMemory *); };
typedef struct _GstAllocatorClass { GstObjectClass object_class;
  GstMemory *(*alloc) (GstAllocator *, gsize, GstAllocationParams *); void (*free) (GstAllocator *, GstMemory *); } GstAllocatorClass;
#define GST_TYPE_ALLOCATOR (gst_allocator_get_type ())
GType gst_allocator_get_type (void);
#define GST_ALLOCATOR_CLASS(k) ((GstAllocatorClass *) (k))
#define GST_ALLOCATOR_GET_CLASS(o) ((GstAllocatorClass *) ((GTypeInstance *) (o))->g_class)
#define GST_ALLOCATOR_CAST(o) ((GstAllocator *) (o))
#define GST_ALLOCATOR(o) ((GstAllocator *) (o))
#define GST_ALLOCATOR_FLAG_CUSTOM_ALLOC 16
void gst_memory_init (GstMemory *, guint flags, GstAllocator *, GstMemory *parent, gsize maxsize, gsize align, gsize offset, gsize size);
GstMemory *gst_allocator_alloc (GstAllocator *, gsize, GstAllocationParams *);
GstMemory *gst_memory_ref (GstMemory *); void gst_memory_unref (GstMemory *);
typedef struct { GstMemory *memory; GstMapFlags flags; guint8 *data; gsize size, maxsize; } GstMapInfo;

typedef struct _GstVideoMetaFake GstVideoMetaFake;
typedef struct _GstBuffer { GstMiniObject mini_object; GstClockTime pts, dts, duration; GstMemory *mem[4]; guint n_mem;
  GstVideoMetaFake *video_meta; } GstBuffer;
#define GST_BUFFER_PTS(b) ((b)->pts)
#define GST_BUFFER_DURATION(b) ((b)->duration)
#define GST_BUFFER_PTS_IS_VALID(b) GST_CLOCK_TIME_IS_VALID ((b)->pts)
#define GST_BUFFER_DURATION_IS_VALID(b) GST_CLOCK_TIME_IS_VALID ((b)->duration)
GstBuffer *gst_buffer_new (void); GstBuffer *gst_buffer_new_allocate (GstAllocator *, gsize, GstAllocationParams *);
GstBuffer *gst_buffer_new_wrapped (gpointer data, gsize size);          /* takes ownership: g_free */
GstBuffer *gst_buffer_new_wrapped_full (guint flags, gpointer data, gsize maxsize, gsize offset, gsize size, gpointer user_data,
    GDestroyNotify notify);
void gst_buffer_append_memory (GstBuffer *, GstMemory *);
guint gst_buffer_n_memory (GstBuffer *); GstMemory *gst_buffer_peek_memory (GstBuffer *, guint idx);
gboolean gst_buffer_map (GstBuffer *, GstMapInfo *, GstMapFlags); void gst_buffer_unmap (GstBuffer *, GstMapInfo *);
GstBuffer *gst_buffer_ref (GstBuffer *); void gst_buffer_unref (GstBuffer *);

/* ---- queries (allocation only) ---------------------------------------------- */
typedef struct _GstQuery { GstCaps *caps; GstAllocator *allocator; gboolean has_video_meta; } GstQuery;
void gst_query_parse_allocation (GstQuery *, GstCaps **, gboolean *need_pool);
void gst_query_add_allocation_param (GstQuery *, GstAllocator *, const GstAllocationParams *);
void gst_query_add_allocation_meta (GstQuery *, GType api, const GstStructure *);

/* ---- element, pad ----------------------------------------------------------- */
typedef enum { GST_FLOW_OK = 0, GST_FLOW_FLUSHING = -2, GST_FLOW_EOS = -3, GST_FLOW_NOT_NEGOTIATED = -4, GST_FLOW_ERROR = -5 } GstFlowReturn;
typedef enum { GST_PAD_SRC = 1, GST_PAD_SINK = 2 } GstPadDirection;
typedef enum { GST_PAD_ALWAYS = 0 } GstPadPresence;
typedef struct { const gchar *string; } GstStaticCaps;
#define GST_STATIC_CAPS(s) { s }
typedef struct { const gchar *name_template; GstPadDirection direction; GstPadPresence presence; GstStaticCaps static_caps; } GstStaticPadTemplate;
#define GST_STATIC_PAD_TEMPLATE(n, d, p, c) { n, d, p, c }
typedef struct _GstPad GstPad;
typedef GstFlowReturn (*GstPadChainFunction) (GstPad *, GstObject *, GstBuffer *);
typedef gboolean (*GstPadEventFunction) (GstPad *, GstObject *, GstEvent *);
struct _GstPad { GstObject object; GstPadDirection direction; GstPadChainFunction chainfunc; GstPadEventFunction eventfunc;
  GstCaps *current_caps; GstObject *parent; };
GstPad *gst_pad_new_from_static_template (GstStaticPadTemplate *, const gchar *);
void gst_pad_set_chain_function (GstPad *, GstPadChainFunction); void gst_pad_set_event_function (GstPad *, GstPadEventFunction);
GstCaps *gst_pad_get_current_caps (GstPad *);
/* fake-only: what a peer pad does */
GstFlowReturn gst_fake_pad_chain (GstPad *, GstBuffer *); gboolean gst_fake_pad_send_event (GstPad *, GstEvent *);
void gst_fake_pad_set_caps (GstPad *, GstCaps *);

typedef struct _GstElement { GstObject object; GstPad *pads[8]; guint n_pads; gint fake_errors; } GstElement;
typedef struct _GstElementClass { GstObjectClass parent_class; } GstElementClass;
#define GST_TYPE_ELEMENT (gst_element_get_type ())
GType gst_element_get_type (void);
#define GST_ELEMENT(o) ((GstElement *) (o))
#define GST_ELEMENT_CAST(o) ((GstElement *) (o))
#define GST_ELEMENT_CLASS(k) ((GstElementClass *) (k))
typedef enum { GST_RANK_NONE = 0 } GstRank;
gboolean gst_element_register (GstPlugin *, const gchar *, guint, GType);
gboolean gst_element_add_pad (GstElement *, GstPad *);
GstPad *gst_element_get_static_pad (GstElement *, const gchar *);
void gst_element_class_add_static_pad_template (GstElementClass *, GstStaticPadTemplate *);
void gst_element_class_set_static_metadata (GstElementClass *, const gchar *, const gchar *, const gchar *, const gchar *);
#define GST_DEBUG_FUNCPTR(f) (f)
#define GST_DEBUG_CATEGORY_STATIC(c) static int c
#define GST_DEBUG_CATEGORY_INIT(c, n, col, d) ((c) = 0)
#define GST_DEBUG_OBJECT(o, ...) ((void) (o))
#define GST_LOG_OBJECT(o, ...) ((void) (o))
#define GST_WARNING_OBJECT(o, ...) ((void) (o))
#define GST_ELEMENT_ERROR(el, dom, code, text, debug) do { ((GstElement *) (el))->fake_errors++; \
    printf ("element error: "); printf text; printf (" -- "); printf debug; printf ("\n"); } while (0)
void gst_init (int *, char ***); const gchar *gst_version_string (void);
#endif
