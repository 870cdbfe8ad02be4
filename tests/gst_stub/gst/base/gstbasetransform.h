/* functional fake: see ../../README.md */
#ifndef GST_STUB_BASETRANSFORM_H
#define GST_STUB_BASETRANSFORM_H
#include <gst/gst.h>
typedef struct _GstBaseTransform { GstElement element; GstPad *sinkpad, *srcpad; gboolean have_segment; GstSegment segment;
  gboolean fake_in_place; } GstBaseTransform;
typedef struct _GstBaseTransformClass {
  GstElementClass parent_class;
  gboolean (*set_caps) (GstBaseTransform *, GstCaps *, GstCaps *);
  gboolean (*propose_allocation) (GstBaseTransform *, GstQuery *decide_query, GstQuery *query);
  gboolean (*sink_event) (GstBaseTransform *, GstEvent *);
  GstFlowReturn (*transform_ip) (GstBaseTransform *, GstBuffer *);
  gboolean (*start) (GstBaseTransform *);
  gboolean (*stop) (GstBaseTransform *);
} GstBaseTransformClass;
#define GST_TYPE_BASE_TRANSFORM (gst_base_transform_get_type ())
GType gst_base_transform_get_type (void);
#define GST_BASE_TRANSFORM(o) ((GstBaseTransform *) (o))
#define GST_BASE_TRANSFORM_CLASS(k) ((GstBaseTransformClass *) (k))
#define GST_BASE_TRANSFORM_GET_CLASS(o) ((GstBaseTransformClass *) ((GTypeInstance *) (o))->g_class)
#define GST_BASE_TRANSFORM_SINK_NAME "sink"
#define GST_BASE_TRANSFORM_SRC_NAME "src"
void gst_base_transform_set_in_place (GstBaseTransform *, gboolean);
#endif
