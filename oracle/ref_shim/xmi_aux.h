/* oracle/ref_shim/xmi_aux.h -- TEST INFRASTRUCTURE: xmi_memdup (src/xmi_aux.c:33-35 is g_memdup) for src/xmi_spline.c. */
#ifndef ORC_REF_SHIM_XMI_AUX_H
#define ORC_REF_SHIM_XMI_AUX_H
#include <stdlib.h>
#include <string.h>
static inline void *xmi_memdup(const void *mem, size_t bytes) {
	void *p = malloc(bytes);
	if (p && mem) memcpy(p, mem, bytes);
	return p;
}
#endif
