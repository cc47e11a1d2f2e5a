/* orc_core.c -- Philox, quadratic solver, xmi_init_input restatement.  TEST INFRASTRUCTURE ONLY. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* Philox4x32-10, as published in Random123 (philox.h): multipliers 0xD2511F53 / 0xCD9E8D57,
 * Weyl key increments 0x9E3779B9 / 0xBB67AE85, ten rounds. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for (int round = 0; round < 10; round++) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c0;
		uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* src/xmi_aux_f.F90:1872-1905 */
int orc_poly_solve_quadratic(double a, double b, double c, double *rv1, double *rv2) {
	if (a == 0.0) {
		if (b == 0.0) return 0;
		*rv1 = -1.0 * c / b;
		return 1;
	}
	double delta = b * b - 4.0 * a * c;
	if (delta < 0.0) return 0;
	if (delta == 0.0) {
		*rv1 = -b / 2.0 / a;
		*rv2 = *rv1;
		return 2;
	}
	double sq = sqrt(delta);
	double t1 = (-b + sq) / 2.0 / a, t2 = (-b - sq) / 2.0 / a;
	*rv1 = t1 < t2 ? t1 : t2;
	*rv2 = t1 < t2 ? t2 : t1;
	return 2;
}

static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(const double *a, const double *b, double *c) {
	c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static void norm3(double *a) { double n = sqrt(dot3(a, a)); a[0] /= n; a[1] /= n; a[2] /= n; }

/* src/xmi_main.F90:1741-1918 */
int orc_init_input(xmb_input *input, orc_derived *d) {
	xmb_geometry *g = input->geometry;
	memset(d, 0, sizeof(*d));
	norm3(g->n_sample_orientation);                                   /* :1759 */
	if (g->n_sample_orientation[2] < 0.0)                             /* :1760-1762 */
		for (int i = 0; i < 3; i++) g->n_sample_orientation[i] *= -1.0;
	norm3(g->n_detector_orientation);                                 /* :1763 */
	d->detector_radius = sqrt(g->area_detector / M_PI);               /* :1769 */
	d->collimator_height = g->collimator_height;
	if (g->collimator_height > 0.0 && g->collimator_diameter > 0.0) { /* :1780-1796 */
		d->collimator_present = 1;
		d->collimator_radius = g->collimator_diameter / 2.0;
		if (d->collimator_radius >= d->detector_radius) return 0;
		d->half_apex = atan((d->detector_radius - d->collimator_radius) / g->collimator_height);
		d->vertex[0] = d->detector_radius / tan(d->half_apex);
	} else d->collimator_present = 0;
	double nx[3], ny[3], nz[3], ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0};
	memcpy(nx, g->n_detector_orientation, sizeof(nx));                /* :1803 */
	if (fabs(dot3(nx, ex)) > 1.0e-6) cross3(ey, nx, ny); else cross3(ex, nx, ny);   /* :1812-1823 */
	norm3(ny);
	cross3(nx, ny, nz);
	double A[3][3], B[3][3];
	for (int i = 0; i < 3; i++) { A[i][0] = nx[i]; A[i][1] = ny[i]; A[i][2] = nz[i]; }   /* columns, :1836-1841 */
	/* src/xmi_aux_f.F90:2076-2106 */
	double det = A[0][0] * A[1][1] * A[2][2] - A[0][0] * A[1][2] * A[2][1] - A[0][1] * A[1][0] * A[2][2] +
	             A[0][1] * A[1][2] * A[2][0] + A[0][2] * A[1][0] * A[2][1] - A[0][2] * A[1][1] * A[2][0];
	double di = 1.0 / det;
	B[0][0] = +di * (A[1][1] * A[2][2] - A[1][2] * A[2][1]);
	B[1][0] = -di * (A[1][0] * A[2][2] - A[1][2] * A[2][0]);
	B[2][0] = +di * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
	B[0][1] = -di * (A[0][1] * A[2][2] - A[0][2] * A[2][1]);
	B[1][1] = +di * (A[0][0] * A[2][2] - A[0][2] * A[2][0]);
	B[2][1] = -di * (A[0][0] * A[2][1] - A[0][1] * A[2][0]);
	B[0][2] = +di * (A[0][1] * A[1][2] - A[0][2] * A[1][1]);
	B[1][2] = -di * (A[0][0] * A[1][2] - A[0][2] * A[1][0]);
	B[2][2] = +di * (A[0][0] * A[1][1] - A[0][1] * A[1][0]);
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { d->ndo_new[i * 3 + j] = A[i][j]; d->ndo_inv[i * 3 + j] = B[i][j]; }
	double p0[3] = {0.0, 0.0, g->d_sample_source};                    /* :1862-1868 */
	double dd = sqrt(pow(g->p_detector_window[0] - p0[0], 2) + pow(g->p_detector_window[1] - p0[1], 2) +
	                 pow(g->p_detector_window[2] - p0[2], 2));
	d->detector_solid_angle = 2 * M_PI * (1.0 - cos(atan(d->detector_radius / dd)));
	for (int i = 0; i < 3; i++) d->n_sample_orientation_det[i] = dot3(B[i], g->n_sample_orientation);   /* :1876-1878 */
	const xmb_composition *c = input->composition;
	int n = c->n_layers, ref = c->reference_layer - 1;
	d->n_layers = n;
	d->thickness_along_Z = (double *)calloc(n, sizeof(double));
	d->Z_coord_begin = (double *)calloc(n, sizeof(double));
	d->Z_coord_end = (double *)calloc(n, sizeof(double));
	double ez[3] = {0, 0, 1};
	for (int j = 0; j < n; j++) d->thickness_along_Z[j] = fabs(c->layers[j].thickness / dot3(g->n_sample_orientation, ez));   /* :1885-1890 */
	d->Z_coord_begin[ref] = 0.0 + g->d_sample_source;                 /* :1892-1896 */
	d->Z_coord_end[ref] = d->thickness_along_Z[ref] + g->d_sample_source;
	for (int j = ref + 1; j < n; j++) { d->Z_coord_begin[j] = d->Z_coord_end[j - 1]; d->Z_coord_end[j] = d->Z_coord_begin[j] + d->thickness_along_Z[j]; }
	for (int j = ref - 1; j >= 0; j--) { d->Z_coord_end[j] = d->Z_coord_begin[j + 1]; d->Z_coord_begin[j] = d->Z_coord_end[j] - d->thickness_along_Z[j]; }
	return 1;
}

void orc_free_derived(orc_derived *d) {
	free(d->thickness_along_Z); free(d->Z_coord_begin); free(d->Z_coord_end);
	d->thickness_along_Z = d->Z_coord_begin = d->Z_coord_end = NULL;
}
