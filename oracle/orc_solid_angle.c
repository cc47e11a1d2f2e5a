/* orc_solid_angle.c -- CPU restatement of the solid-angle grid.  TEST INFRASTRUCTURE ONLY.
 * Follows src/xmi_solid_angle_f.F90 (Fortran/OpenMP path), cross-checked with src/xmi_kernels.cl:219-452. */
#include <math.h>
#include <stdlib.h>
#include "oracle.h"
#include "orc_rng.h"

/* src/xmi_aux_f.F90:1143-1170; returns 0 for a line parallel to the plane */
static int intersection_plane_line(const double pp[3], const double pn[3], const double lp[3], const double ld[3], double out[3]) {
	double ItimesN = ld[0] * pn[0] + ld[1] * pn[1] + ld[2] * pn[2];
	if (ItimesN == 0.0) return 0;
	double d = ((pp[0] - lp[0]) * pn[0] + (pp[1] - lp[1]) * pn[1] + (pp[2] - lp[2]) * pn[2]) / ItimesN;
	for (int i = 0; i < 3; i++) out[i] = d * ld[i] + lp[i];
	return 1;
}
static double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

/* src/xmi_solid_angle_f.F90:432-710 */
static double single_solid_angle(const orc_derived *D, double r1, double theta1, long hits_per_single,
                                 uint64_t seed, uint64_t point_id, uint32_t block0, uint32_t tag, long *hits_out);
double orc_single_solid_angle(const orc_derived *D, double r1, double theta1, long hits_per_single,
                              uint64_t seed, uint64_t point_id, long *hits_out) {
	return single_solid_angle(D, r1, theta1, hits_per_single, seed, point_id, 0, ORC_TAG_SOLID_ANGLE, hits_out);
}
/* The same calculation for an interaction point beyond the grid (xmi_get_solid_angle, src/xmi_solid_angle_f.F90:783-789):
 * rays of photon g at interaction `order` come from stream (g, blocks (order << 20) + ray pair, ORC_TAG_SA_FALLBACK). */
double orc_single_solid_angle_photon(const orc_derived *D, double r1, double theta1, long hits_per_single,
                                     uint64_t seed, uint64_t g, int order, long *hits_out) {
	return single_solid_angle(D, r1, theta1, hits_per_single, seed, g, (uint32_t)order << 20, ORC_TAG_SA_FALLBACK, hits_out);
}
static double single_solid_angle(const orc_derived *D, double r1, double theta1, long hits_per_single,
                                 uint64_t seed, uint64_t point_id, uint32_t block0, uint32_t tag, long *hits_out) {
	const double detector_normal[3] = {0.0, 0.0, 1.0};
	double r, theta, full_cone_base_radius;
	int outside_collimator;
	if (hits_out) *hits_out = 0;
	if (!D->collimator_present) {                                              /* :481-487 */
		r = r1; theta = theta1; full_cone_base_radius = D->detector_radius; outside_collimator = 0;
	} else if (fabs(D->collimator_radius - D->detector_radius) < 0.000001) {    /* :488-519 cylindrical */
		if (r1 * cos(theta1) <= D->detector_radius) { r = r1; theta = theta1; full_cone_base_radius = D->detector_radius; }
		else {
			r = sqrt(r1 * r1 - 2.0 * r1 * sin(theta1) * D->collimator_height + D->collimator_height * D->collimator_height);
			theta = acos(r1 * cos(theta1) / r);
			full_cone_base_radius = D->collimator_radius;
		}
		outside_collimator = r1 * sin(theta1) > D->collimator_height;
	} else {                                                                   /* :520-558 conical */
		if (r1 * cos(theta1) <= D->detector_radius &&
		    r1 * sin(theta1) <= D->collimator_height * (r1 * cos(theta1) - D->detector_radius) / (D->collimator_radius - D->detector_radius)) {
			r = r1; theta = theta1; full_cone_base_radius = D->detector_radius;
		} else if (r1 * sin(theta1) <= D->collimator_height) {
			return 0.0;
		} else {
			r = sqrt(r1 * r1 - 2.0 * r1 * sin(theta1) * D->collimator_height + D->collimator_height * D->collimator_height);
			theta = acos(r1 * cos(theta1) / r);
			full_cone_base_radius = D->collimator_radius;
		}
		outside_collimator = r1 * sin(theta1) > D->collimator_height;
	}
	double beta = atan(full_cone_base_radius / r);                              /* :569 */
	double alpha1 = atan(full_cone_base_radius * sin(theta) / (r - full_cone_base_radius * cos(theta)));
	if (alpha1 <= 0.0) alpha1 += M_PI;                                          /* :575-576 */
	double full_cone_apex = beta > alpha1 ? beta : alpha1;                      /* :584 */
	double cos_full_cone_apex = cos(full_cone_apex);
	double full_cone_solid_angle = 2 * M_PI * (1.0 - cos_full_cone_apex);       /* :590 */
	/* rotation matrix, columns (1,0,0), (0,-sin,cos), (0,-cos,-sin)  (:593-597) */
	double st = sin(theta), ct = cos(theta);
	double det_point[3] = {0, 0, 0}, coll_point[3] = {0, 0, D->collimator_height};
	double line_point[3] = {0.0, r1 * cos(theta1), r1 * sin(theta1)};           /* :609 */
	long detector_hits = 0;
	orc_rng rng;
	orc_rng_init(&rng, seed, point_id, tag);
	rng.ctr[2] = block0;
	for (long i = 0; i < hits_per_single; i++) {                                /* :630-693 */
		double theta_rng = acos(1.0 - orc_rng_uniform(&rng) * (1.0 - cos_full_cone_apex));
		double phi_rng = orc_rng_uniform(&rng) * 2.0 * M_PI;
		double c[3] = {sin(theta_rng) * cos(phi_rng), sin(theta_rng) * sin(phi_rng), cos(theta_rng)};
		double dv[3] = {c[0], -st * c[1] - ct * c[2], ct * c[1] - st * c[2]};   /* MATMUL(rotation_matrix, c) */
		if (dv[2] >= 0.0) continue;                                             /* :664 */
		double ip[3];
		if (outside_collimator) {                                               /* :667-679 */
			if (!intersection_plane_line(coll_point, detector_normal, line_point, dv, ip)) continue;
			ip[2] = 0.0;
			if (norm3(ip) > D->collimator_radius) continue;
		}
		if (!intersection_plane_line(det_point, detector_normal, line_point, dv, ip)) continue;
		if (norm3(ip) <= D->detector_radius) detector_hits++;                   /* :691 */
	}
	if (hits_out) *hits_out = detector_hits;
	return full_cone_solid_angle * (double)detector_hits / (double)hits_per_single;   /* :703 */
}

/* src/xmi_solid_angle_f.F90:373-409 */
void orc_solid_angle_grid(const orc_derived *d, const double *r_vals, const int *r_idx, int n_r,
                          const double *theta_vals, const int *theta_idx, int n_theta, long full_n_r,
                          long hits_per_single, uint64_t seed, double *solid_angles, int32_t *hits, int n_threads) {
	if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(n_threads) collapse(2)
	for (int it = 0; it < n_theta; it++)
		for (int ir = 0; ir < n_r; ir++) {
			long h = 0;
			uint64_t id = (uint64_t)theta_idx[it] * (uint64_t)full_n_r + (uint64_t)r_idx[ir];
			solid_angles[(size_t)it * n_r + ir] = orc_single_solid_angle(d, r_vals[ir], theta_vals[it], hits_per_single, seed, id, &h);
			if (hits) hits[(size_t)it * n_r + ir] = (int32_t)h;
		}
}

/* src/xmi_solid_angle_f.F90:123-292 */
static double mu_layer(const xmb_xrl_provider *xrl, const xmb_layer *l, double E) {   /* src/xmi_aux_f.F90:1125-1141 */
	double rv = 0.0;
	for (int i = 0; i < l->n_elements; i++) rv += xrl->CS_Total_Kissel(l->Z[i], E) * l->weight[i];
	return rv;
}
static double depth(const xmb_input *in, const orc_derived *d, const xmb_xrl_provider *xrl, double energy, double R) {
	const xmb_composition *c = in->composition;
	int n = c->n_layers, m = n - 1;
	double *mu = (double *)malloc(sizeof(double) * n), my_sum = 0.0;
	for (int i = 0; i < n; i++) { mu[i] = mu_layer(xrl, &c->layers[i], energy); my_sum += mu[i] * c->layers[i].density * d->thickness_along_Z[i]; }
	double Pabs = -1.0 * expm1(-1.0 * my_sum), myln = -1.0 * log1p(-1.0 * R * Pabs);
	my_sum = 0.0;
	for (int i = 0; i < n; i++) { my_sum += mu[i] * c->layers[i].density * d->thickness_along_Z[i]; if (my_sum > myln) { m = i; break; } }
	my_sum = 0.0;
	for (int i = 0; i < m; i++) my_sum += (1.0 - (mu[i] * c->layers[i].density) / (mu[m] * c->layers[m].density)) * d->thickness_along_Z[i];
	double S = my_sum + myln / (mu[m] * c->layers[m].density) + d->Z_coord_begin[0];
	free(mu);
	return S;
}
int orc_solid_angle_axes(const xmb_input *in, const orc_derived *d, const xmb_xrl_provider *xrl, double *r_vals, double *theta_vals, long n) {
	const xmb_excitation *e = in->excitation;
	double e_lo, e_hi;
	if (e->n_continuous > 1 && e->n_discrete > 0) {
		e_lo = fmin(e->continuous[0].energy, e->discrete[0].energy);
		e_hi = fmax(e->continuous[e->n_continuous - 1].energy, e->discrete[e->n_discrete - 1].energy);
	} else if (e->n_continuous > 1) { e_lo = e->continuous[0].energy; e_hi = e->continuous[e->n_continuous - 1].energy; }
	else if (e->n_discrete > 0) { e_lo = e->discrete[0].energy; e_hi = e->discrete[e->n_discrete - 1].energy; }
	else return 0;
	double S1 = depth(in, d, xrl, e_lo, 0.00001), S2 = depth(in, d, xrl, e_hi, 0.99999);
	const double *pw = in->geometry->p_detector_window;
	double d1 = sqrt(pw[0] * pw[0] + pw[1] * pw[1] + (pw[2] - S1) * (pw[2] - S1));
	double d2 = sqrt(pw[0] * pw[0] + pw[1] * pw[1] + (pw[2] - S2) * (pw[2] - S2));
	double r_hi = fmax(d1, d2) * 1.25, r_lo = r_hi / n;                        /* :259-264 */
	double t_hi = M_PI / 2.0, t_lo = 0.00001;                                  /* :267-268 */
	for (long i = 0; i < n; i++) {
		r_vals[i] = r_lo + (r_hi - r_lo) * (double)i / (double)(n - 1);
		theta_vals[i] = t_lo + (t_hi - t_lo) * (double)i / (double)(n - 1);
	}
	return 1;
}
