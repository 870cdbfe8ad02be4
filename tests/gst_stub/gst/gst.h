/* stub: see ../README.md */
#ifndef GST_STUB_GST_H
#define GST_STUB_GST_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef int gboolean; typedef int gint; typedef unsigned int guint; typedef char gchar;
typedef unsigned char guint8; typedef uint32_t guint32; typedef uint64_t guint64; typedef int64_t gint64;
typedef size_t gsize; typedef void *gpointer; typedef const void *gconstpointer; typedef float gfloat;
typedef double gdouble; typedef unsigned long GType; typedef guint64 GstClockTime;
#define TRUE 1
#define FALSE 0
#define G_N_ELEMENTS(a) (sizeof (a) / sizeof ((a)[0]))
#define G_GSIZE_FORMAT "zu"
#define GST_CLOCK_TIME_NONE ((GstClockTime) -1)
#define GST_CLOCK_TIME_IS_VALID(t) (((GstClockTime) (t)) != GST_CLOCK_TIME_NONE)

typedef struct { int dummy; } GMutex;
void g_mutex_init (GMutex *m); void g_mutex_clear (GMutex *m); void g_mutex_lock (GMutex *m); void g_mutex_unlock (GMutex *m);
gpointer g_malloc (gsize n); gpointer g_malloc0 (gsize n); void g_free (gpointer p); gpointer g_memdup2 (gconstpointer p, gsize n);
#define g_new0(type, n) ((type *) g_malloc0 (sizeof (type) * (n)))
gint g_atomic_int_add (volatile gint *atomic, gint val);

typedef struct _GValue GValue; typedef struct _GParamSpec GParamSpec;
typedef struct _GObject { int ref; } GObject;
typedef struct _GObjectClass {
  void (*set_property) (GObject *, guint, const GValue *, GParamSpec *);
  void (*get_property) (GObject *, guint, GValue *, GParamSpec *);
  void (*finalize) (GObject *);
} GObjectClass;
#define G_OBJECT_CLASS(k) ((GObjectClass *) (k))
typedef enum { G_PARAM_READWRITE = 3, G_PARAM_STATIC_STRINGS = 0xe0 } GParamFlags;
GParamSpec *g_param_spec_int (const gchar *, const gchar *, const gchar *, gint, gint, gint, GParamFlags);
void g_object_class_install_property (GObjectClass *, guint, GParamSpec *);
gint g_value_get_int (const GValue *); void g_value_set_int (GValue *, gint);
gpointer g_object_new (GType type, const gchar *first, ...);
#define G_OBJECT_WARN_INVALID_PROPERTY_ID(o, id, p) ((void) (o), (void) (id), (void) (p))
#define G_DECLARE_FINAL_TYPE(Name, name, MOD, OBJ, Parent) \
  GType name##_get_type (void); typedef struct _##Name Name; typedef struct { Parent##Class parent_class; } Name##Class; \
  static inline Name *MOD##_##OBJ (gpointer p) { return (Name *) p; }
#define G_DEFINE_TYPE(Name, name, PARENT_TYPE) \
  static void name##_class_init (Name##Class *klass); static void name##_init (Name *self); \
  static gpointer name##_parent_class = NULL; \
  GType name##_get_type (void) { (void) name##_class_init; (void) name##_init; (void) name##_parent_class; return 0; }

typedef struct _GstObject { GObject object; guint flags; } GstObject;
#define GST_OBJECT_FLAG_SET(o, f) (((GstObject *) (o))->flags |= (f))
gpointer gst_object_ref_sink (gpointer o);
typedef struct _GstPlugin GstPlugin; typedef struct _GstCaps GstCaps; typedef struct _GstEvent GstEvent;
typedef struct _GstElement { GstObject object; } GstElement;
typedef struct _GstElementClass { GObjectClass parent_class; } GstElementClass;
#define GST_ELEMENT(o) ((GstElement *) (o))
#define GST_ELEMENT_CLASS(k) ((GstElementClass *) (k))
#define GST_TYPE_ELEMENT 0
typedef enum { GST_RANK_NONE = 0 } GstRank;
gboolean gst_element_register (GstPlugin *, const gchar *, guint, GType);
typedef enum { GST_FLOW_OK = 0, GST_FLOW_NOT_NEGOTIATED = -4, GST_FLOW_ERROR = -5 } GstFlowReturn;
typedef enum { GST_PAD_SRC = 1, GST_PAD_SINK = 2 } GstPadDirection;
typedef enum { GST_PAD_ALWAYS = 0 } GstPadPresence;
typedef struct { const gchar *string; } GstStaticCaps;
#define GST_STATIC_CAPS(s) { s }
typedef struct { const gchar *name; GstPadDirection dir; GstPadPresence presence; GstStaticCaps caps; } GstStaticPadTemplate;
#define GST_STATIC_PAD_TEMPLATE(n, d, p, c) { n, d, p, c }
typedef struct _GstPad GstPad;
typedef struct _GstBuffer { GstClockTime pts, duration; } GstBuffer;
#define GST_BUFFER_PTS(b) ((b)->pts)
#define GST_BUFFER_DURATION(b) ((b)->duration)
#define GST_BUFFER_DURATION_IS_VALID(b) GST_CLOCK_TIME_IS_VALID ((b)->duration)
typedef enum { GST_MAP_READ = 1, GST_MAP_WRITE = 2, GST_MAP_READWRITE = 3 } GstMapFlags;
typedef struct { guint8 *data; gsize size; } GstMapInfo;
gboolean gst_buffer_map (GstBuffer *, GstMapInfo *, GstMapFlags); void gst_buffer_unmap (GstBuffer *, GstMapInfo *);
void gst_buffer_unref (GstBuffer *); GstBuffer *gst_buffer_new_allocate (gpointer, gsize, gpointer);
GstBuffer *gst_buffer_new_wrapped (gpointer, gsize);
GstCaps *gst_pad_get_current_caps (GstPad *); void gst_caps_unref (GstCaps *);
typedef GstFlowReturn (*GstPadChainFunction) (GstPad *, GstObject *, GstBuffer *);
typedef gboolean (*GstPadEventFunction) (GstPad *, GstObject *, GstEvent *);
GstPad *gst_pad_new_from_static_template (GstStaticPadTemplate *, const gchar *);
void gst_pad_set_chain_function (GstPad *, GstPadChainFunction); void gst_pad_set_event_function (GstPad *, GstPadEventFunction);
gboolean gst_element_add_pad (GstElement *, GstPad *);
void gst_element_class_add_static_pad_template (GstElementClass *, GstStaticPadTemplate *);
void gst_element_class_set_static_metadata (GstElementClass *, const gchar *, const gchar *, const gchar *, const gchar *);
typedef enum { GST_EVENT_FLUSH_STOP = 1, GST_EVENT_EOS = 2, GST_EVENT_CAPS = 3 } GstEventType;
GstEventType gst_event_type_stub (GstEvent *);
#define GST_EVENT_TYPE(e) gst_event_type_stub (e)
void gst_event_unref (GstEvent *);
#define GST_DEBUG_FUNCPTR(f) (f)
#define GST_DEBUG_CATEGORY_STATIC(c) static int c
#define GST_DEBUG_CATEGORY_INIT(c, n, col, d) ((c) = 0)
#define GST_ELEMENT_ERROR(el, dom, code, text, debug) do { (void) (el); printf text; printf debug; } while (0)
void gst_init (int *, char ***); const gchar *gst_version_string (void);

/* allocator */
typedef struct _GstAllocator GstAllocator; typedef struct _GstMemory GstMemory; typedef struct _GstAllocationParams GstAllocationParams;
struct _GstMemory { GstAllocator *allocator; gsize size; };
#define GST_MEMORY_CAST(m) ((GstMemory *) (m))
struct _GstAllocator { GstObject object; const gchar *mem_type; gpointer (*mem_map) (GstMemory *, gsize, GstMapFlags); void (*mem_unmap) (GstMemory *); };
typedef struct _GstAllocatorClass { GObjectClass object_class; GstMemory *(*alloc) (GstAllocator *, gsize, GstAllocationParams *); void (*free) (GstAllocator *, GstMemory *); } GstAllocatorClass;
#define GST_ALLOCATOR_CLASS(k) ((GstAllocatorClass *) (k))
#define GST_ALLOCATOR_CAST(o) ((GstAllocator *) (o))
#define GST_TYPE_ALLOCATOR 0
#define GST_ALLOCATOR_FLAG_CUSTOM_ALLOC 16
void gst_memory_init (GstMemory *, guint flags, GstAllocator *, GstMemory *parent, gsize maxsize, gsize align, gsize offset, gsize size);
#endif
