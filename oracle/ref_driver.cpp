// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE: C entry points around the reference's own code compiled by
// oracle/build_ref.sh into oracle/_ref/libxmi_ref.so (see oracle/ref_shim/README.md).  The kernel text included below
// is /root/reference/src/xmi_kernels.cl as it lies there (one sed substitution for the OpenCL vector literal); the
// launch loop mirrors src/xmi_solid_angle_cl.c:379-401 (global work size = grid / RANGE_DIVIDER, RANGE_DIVIDER^2
// launches with offsets), the argument list src/xmi_solid_angle_cl.c:340-372.
#include "cl_shim.hpp"
// the reference concatenates these in this order before xmi_kernels.cl (src/xmi_solid_angle_cl.c:296-303); its
// "openclfeatures.h" is replaced by ref_shim/gccfeatures.h, which compilerfeatures.h picks for GCC
#include <compilerfeatures.h>
#include <sse.h>
#include <array.h>
#include <threefry.h>
#ifndef RANGE_DIVIDER
#define RANGE_DIVIDER 1
#endif
#include "xmi_kernels_cl.inc"

extern "C" {
#include "xmi_spline.h"

// solid_angles[theta][r] (r fastest), float, as the OpenCL host reads it back (src/xmi_solid_angle_cl.c:404-420)
int ref_solid_angle_calculation_cl(const float *r_vals, int nr, const float *theta_vals, int nt, float *solid_angles,
                                   int collimator_present, float detector_radius, float collimator_radius,
                                   float collimator_height, int hits_per_single) {
#pragma omp parallel for schedule(dynamic, 1)
	for (int t1 = 0; t1 < nt; t1++) {
		orc_cl_gsz[0] = (size_t)nr; orc_cl_gsz[1] = (size_t)nt;
		for (int t0 = 0; t0 < nr; t0++) {
			orc_cl_gid[0] = (size_t)t0; orc_cl_gid[1] = (size_t)t1;
			xmi_solid_angle_calculation(r_vals, theta_vals, solid_angles, collimator_present, detector_radius, collimator_radius,
			                            collimator_height, hits_per_single);
		}
	}
	return 1;
}

// natural cubic spline of the reference (src/xmi_spline.c), one evaluation
double ref_cubic_spline(double *x, double *y, size_t n, double v) {
	xmi_cubic_spline *s = xmi_cubic_spline_init(x, y, n);
	const double r = xmi_cubic_spline_eval(s, v);
	xmi_cubic_spline_free(s);
	return r;
}
}
