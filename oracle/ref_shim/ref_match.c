/* oracle/ref_shim/ref_match.c -- TEST INFRASTRUCTURE: runs the reference's own cache match rules, xmi_check_solid_angle_match
 * (src/xmi_solid_angle.c:420-673) and xmi_check_escape_ratios_match (src/xmi_detector.c:143-172), cut out of their files
 * by oracle/build_ref.sh into oracle/_ref/ at build time (the rest of both files is HDF5 I/O), so that host_cache.cpp's
 * xmb_check_solid_angle_match / xmb_check_escape_ratios_match can be pinned against them.  Supplied here:
 *   - xmi_normalize_vector_double (Fortran in the reference, src/xmi_aux_f.F90:1240-1251);
 *   - CS_Total_Kissel (xraylib): forwarded to the provider both sides of the comparison use (the surrogate);
 *   - GLib names: oracle/ref_shim/glib.h. */
#include <math.h>
#include "xmi_data_structs.h"
#include "xmimsim_b200.h"

void xmi_normalize_vector_double(double *array, int n) {
	double s = 0.0;
	for (int i = 0; i < n; i++) s += array[i] * array[i];
	s = sqrt(s);
	for (int i = 0; i < n; i++) array[i] /= s;
}
static double CS_Total_Kissel(int Z, double E, void *error) { (void)error; return xmb_xrl_surrogate()->CS_Total_Kissel(Z, E); }

#include "solid_angle_match.inc"
#include "escape_ratios_match.inc"

/* both functions normalise the orientation vectors of their arguments in place: pass copies */
int ref_check_solid_angle_match(void *A, void *B) { return xmi_check_solid_angle_match((xmi_input *)A, (xmi_input *)B); }
int ref_check_escape_ratios_match(void *A, void *B) { return xmi_check_escape_ratios_match((xmi_input *)A, (xmi_input *)B); }

/* the reference's defaults: xmi_get_default_escape_ratios_options (src/xmi_detector.c:643-646) and the initialiser of
 * __default_main_options (src/xmi_data_structs.c:2531-2547), both cut out by oracle/build_ref.sh */
#include "xmi_detector.h"
#include "default_escape_options.inc"
#include "default_main_options.inc"
void ref_default_escape_ratios_options(void *out) { *(xmi_escape_ratios_options *)out = xmi_get_default_escape_ratios_options(); }
void ref_default_main_options(void *out) { *(xmi_main_options *)out = __default_main_options; }
