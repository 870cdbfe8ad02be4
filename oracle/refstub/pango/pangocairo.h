/* stub: see ../README.md -- the cairo calls gstttmlblur.c makes (implemented in shim.c) */
#ifndef REFSTUB_PANGOCAIRO_H
#define REFSTUB_PANGOCAIRO_H
typedef struct _cairo_surface cairo_surface_t;
typedef struct { int unused; } cairo_user_data_key_t;
typedef void (*cairo_destroy_func_t) (void *data);
typedef enum { CAIRO_FORMAT_ARGB32 = 0 } cairo_format_t;
int cairo_image_surface_get_width (cairo_surface_t *surface);
int cairo_image_surface_get_height (cairo_surface_t *surface);
int cairo_image_surface_get_stride (cairo_surface_t *surface);
unsigned char *cairo_image_surface_get_data (cairo_surface_t *surface);
cairo_surface_t *cairo_image_surface_create_for_data (unsigned char *data, cairo_format_t format,
    int width, int height, int stride);
int cairo_surface_set_user_data (cairo_surface_t *surface, const cairo_user_data_key_t *key,
    void *user_data, cairo_destroy_func_t destroy);
#endif
