/* stub: see README.md -- the pixman API gstttmlblur.c uses. The fixed-point macros are pixman's
 * public ones; the functions are implemented in shim.c. */
#ifndef REFSTUB_PIXMAN_H
#define REFSTUB_PIXMAN_H
#include <stdint.h>
typedef int32_t pixman_fixed_16_16_t;
typedef pixman_fixed_16_16_t pixman_fixed_t;
#define pixman_int_to_fixed(i) ((pixman_fixed_t) ((uint32_t) (i) << 16))
#define pixman_double_to_fixed(d) ((pixman_fixed_t) ((d) * 65536.0))
typedef struct _pixman_image pixman_image_t;
typedef enum { PIXMAN_a8r8g8b8 = 0x20028888 } pixman_format_code_t;
typedef enum { PIXMAN_FILTER_CONVOLUTION = 6 } pixman_filter_t;
typedef enum { PIXMAN_OP_SRC = 1 } pixman_op_t;
typedef int pixman_bool_t;
pixman_image_t *pixman_image_create_bits (pixman_format_code_t format, int width, int height,
    uint32_t *bits, int rowstride_bytes);
pixman_bool_t pixman_image_set_filter (pixman_image_t *image, pixman_filter_t filter,
    const pixman_fixed_t *filter_params, int n_filter_params);
void pixman_image_composite (pixman_op_t op, pixman_image_t *src, pixman_image_t *mask, pixman_image_t *dest,
    int16_t src_x, int16_t src_y, int16_t mask_x, int16_t mask_y, int16_t dest_x, int16_t dest_y,
    uint16_t width, uint16_t height);
pixman_bool_t pixman_image_unref (pixman_image_t *image);
#endif
