/* stub: see ../../README.md */
#ifndef GST_STUB_BASETRANSFORM_H
#define GST_STUB_BASETRANSFORM_H
#include <gst/gst.h>
typedef struct _GstBaseTransform { GstElement element; } GstBaseTransform;
typedef struct _GstBaseTransformClass {
  GstElementClass parent_class;
  gboolean (*set_caps) (GstBaseTransform *, GstCaps *, GstCaps *);
  GstFlowReturn (*transform_ip) (GstBaseTransform *, GstBuffer *);
  gboolean (*start) (GstBaseTransform *);
  gboolean (*stop) (GstBaseTransform *);
} GstBaseTransformClass;
#define GST_TYPE_BASE_TRANSFORM 0
#define GST_BASE_TRANSFORM(o) ((GstBaseTransform *) (o))
#define GST_BASE_TRANSFORM_CLASS(k) ((GstBaseTransformClass *) (k))
void gst_base_transform_set_in_place (GstBaseTransform *, gboolean);
#endif
