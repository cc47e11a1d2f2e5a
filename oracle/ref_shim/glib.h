/* oracle/ref_shim/glib.h -- TEST INFRASTRUCTURE: the GLib names used by the reference sources that oracle/build_ref.sh
 * compiles (src/xmi_spline.c; include/xmi_data_structs.h and xmi_output_raw2struct of src/xmi_data_structs.c), on libc
 * (GLib is not in the image).  Container types are opaque: nothing compiled here touches them. */
#ifndef ORC_REF_SHIM_GLIB_H
#define ORC_REF_SHIM_GLIB_H
#include <stdlib.h>
#include <string.h>
#define G_BEGIN_DECLS
#define G_END_DECLS
typedef int gboolean;
typedef char gchar;
typedef int gint;
typedef unsigned int guint;
typedef double gdouble;
typedef void *gpointer;
typedef unsigned long GType;
typedef struct _GPtrArray GPtrArray;
typedef struct _GArray GArray;
typedef struct _GHashTable GHashTable;
typedef struct _GError GError;
#define g_malloc(n) malloc(n)
#define g_malloc0(n) calloc(1, (n))
#define g_realloc(p, n) realloc((p), (n))
#define g_free(p) free(p)
#define g_strdup(s) ((s) ? strdup(s) : NULL)
#define g_ascii_strtod(s, e) strtod((s), (e))
#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#endif
