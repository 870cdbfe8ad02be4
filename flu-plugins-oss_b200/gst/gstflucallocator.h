/* gstflucallocator.h -- pinned pool frames of the B200 blend context as GstMemory. */
#ifndef __GST_FLUC_ALLOCATOR_H__
#define __GST_FLUC_ALLOCATOR_H__

#include <gst/gst.h>
#include <gst/video/video.h>

#include "fluc_ttmlblend.h"

GType gst_fluc_allocator_get_type (void);
/* frames of one geometry; ctx must outlive the allocator and every memory it handed out */
GstAllocator *gst_fluc_allocator_new (FlucTtmlBlend * ctx, const GstVideoInfo * info,
    FlucTtmlBlendFormat format);
/* the strides / offsets a buffer with memory of this allocator must carry as GstVideoMeta */
void gst_fluc_allocator_fill_video_meta (GstAllocator * allocator, GstMemory * memory,
    gsize offset[GST_VIDEO_MAX_PLANES], gint stride[GST_VIDEO_MAX_PLANES]);
/* one video buffer: memory from the allocator plus the GstVideoMeta that goes with it (what a
 * buffer pool configured with this allocator hands out) */
GstBuffer *gst_fluc_allocator_alloc_video_buffer (GstAllocator * allocator);
gboolean gst_is_fluc_memory (GstMemory * memory);

#endif
