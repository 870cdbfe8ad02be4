/* stub: see ../README.md -- the GLib bits gstttmlblur.c / gstttmlblur.h use */
#ifndef REFSTUB_GST_H
#define REFSTUB_GST_H
#include <alloca.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef __cplusplus
#define G_BEGIN_DECLS extern "C" {
#define G_END_DECLS }
#else
#define G_BEGIN_DECLS
#define G_END_DECLS
#endif
#define G_PI 3.1415926535897932384626433832795028841971693993751
#define g_newa(type, n) ((type *) alloca (sizeof (type) * (size_t) (n)))
#define g_new(type, n) ((type *) malloc (sizeof (type) * (size_t) (n)))
#define g_malloc0(n) calloc (1, (size_t) (n))
#define g_free free
#endif
