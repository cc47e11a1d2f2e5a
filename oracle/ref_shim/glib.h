/* oracle/ref_shim/glib.h -- TEST INFRASTRUCTURE: the three GLib calls src/xmi_spline.c makes, on libc (GLib is not in the image). */
#ifndef ORC_REF_SHIM_GLIB_H
#define ORC_REF_SHIM_GLIB_H
#include <stdlib.h>
#include <string.h>
#define g_malloc(n) malloc(n)
#define g_free(p) free(p)
#endif
