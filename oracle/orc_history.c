/* orc_history.c -- CPU restatement of the photon-history loop with forced detection.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows, function by function:
 *   xmi_main_msim              src/xmi_main.F90:66-954      (driver, source loops, export)
 *   xmi_coords_dir / _point / _gaussian, xmi_photon_shift_first_layer   :957-1186
 *   xmi_simulate_photon        :1188-1685  (variance-reduction branch :1414-1518)
 *   xmi_simulate_photon_rayleigh / _compton / _fluorescence              :1986-2411
 *   xmi_update_photon_energy_compton2, _dirv, _elecv                     :4985-5182
 *   xmi_coster_kronig_check, xmi_fluorescence_line_check                 :5184-5437
 *   xmi_variance_reduction, xmi_compton_varred2, ..._compton_var_red     src/xmi_variance_reduction.F90:29-1101
 *   xmi_get_solid_angle        src/xmi_solid_angle_f.F90:712-801
 *   helpers                    src/xmi_aux_f.F90:1109-1428, :1841-1941
 * Cross sections come from the table bundle (xmb_tables_host) where the reference calls xraylib.
 * Photon g of the run draws from Philox stream g, sequentially, in the reference's draw order.
 * Not restated (documented in DESIGN.md): brute-force mode (variance reduction off), advanced
 * Compton, the on-the-fly solid-angle Monte-Carlo fallback for off-grid points (counted instead).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "orc_rng.h"
#include "xmb_lines.h"

#define ENERGY_THRESHOLD 1.0      /* src/xmi_aux_f.F90:673 */
#define ENERGY_MAX 200.0          /* :674 */
#define XMI_MEC2 (9.10938188e-31 * 2.99792458e8 * 2.99792458e8 / 1.602176487e-19 / 1000.0)   /* src/xmi_main.F90:49-54 */
#define RAYLEIGH 1
#define COMPTON 2
#define PHOTO 3
#define KEV2ANGST 12.39841930
#define AVOGNUM 0.602252
#define RE2 0.07940775

/* Random-number layout (shared with the GPU engine; the reference's MT19937 order is not reproducible across
 * thread counts anyway): Philox4x32-10, key = seed,
 * counter = (photon id lo, hi, (order << 20) | (stage << 16) | (element << 8) | block, 0x48).
 *   order 0, stage 0        : source sampling, words consumed sequentially
 *   order k, stage 1, blk 0 : {path length, detector point r, detector point phi, atom selection}
 *   order k, stage 1, blk 1 : {interaction type, s0, s1, s2}  s* = first three draws of the interaction
 *                             (Rayleigh: theta, phi; Compton: theta, phi, depolarisation; photo: shell, line, theta)
 *   order k, stage 2, elem e: forced-detection Compton energy of element e, two trials per block
 *   order k, stage 3        : further draws of the interaction (Doppler trials; photo: phi, Coster-Kronig hops)
 * Every draw has a fixed address, so streams never depend on branch history. */
static void draw_block(uint64_t seed, uint64_t g, int order, int stage, int elem, int block, double u[4]) {
	uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
	uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32),
	                   ((uint32_t)order << 20) | ((uint32_t)stage << 16) | ((uint32_t)elem << 8) | (uint32_t)block, ORC_TAG_HISTORY};
	uint32_t out[4];
	orc_philox4x32_10(ctr, key, out);
	for (int i = 0; i < 4; i++) u[i] = out[i] * (1.0 / 4294967296.0);
}
/* sequential cursor over the blocks of one (order, stage, element) sub-stream */
typedef struct { uint64_t seed, g; int order, stage, elem, block, have; double u[4]; } substream_t;
static void sub_init(substream_t *s, uint64_t seed, uint64_t g, int order, int stage, int elem) {
	s->seed = seed; s->g = g; s->order = order; s->stage = stage; s->elem = elem; s->block = 0; s->have = 0;
}
static double sub_uniform(substream_t *s) {
	if (s->have == 0) { draw_block(s->seed, s->g, s->order, s->stage, s->elem, s->block & 0xFF, s->u); s->block++; s->have = 4; }
	return s->u[4 - s->have--];
}

typedef struct {
	double coords[3], dirv[3], elecv[3];
	double energy, weight, theta, phi;
	double weight_escape;     /* escape-ratio mode only (src/xmi_main.F90:1462-1464) */
	int current_layer;        /* 0-based */
	int current_element;      /* Z */
	int current_element_index;
	int n_interactions, last_interaction;
	int gen;                  /* 0, or GEN_BIT for a cascade offspring (brute-force mode): OR-ed into the RNG order field */
	uint64_t seed, g;         /* Philox key and stream (global photon id) */
	int hist_line[32];        /* history(k,1): interaction code or negative line */
	int hist_Z[32];           /* history(k,2) */
	double mus[64];
} photon_t;

typedef struct {
	const xmb_input *in;
	const orc_derived *d;
	const xmb_tables_host *T;
	const xmb_main_options *opt;
	const xmb_solid_angle *sa;
	int cascade;              /* 1..4 */
	int escape_mode;          /* options%escape_ratios_mode */
	int n_int, nch, nL;
	double *channels;         /* [(n_int+1)][nch] */
	double *var_red;          /* Fortran (100,385,n_int) stored as [k][line][Z] -> index ((k*385+line)*100+Z) */
	uint64_t sa_not_found;
	uint64_t n_interactions_total;
} ctx_t;

/* ---- table lookups ------------------------------------------------------------------------ */
typedef struct { int pos; double f; } nodepos_t;

static nodepos_t node_find(const xmb_tables_host *T, double E) {
	int b = (int)floor((E - T->bucket_E0) * T->bucket_inv_dE);
	if (b < 0) b = 0;
	if (b > T->n_buckets - 1) b = T->n_buckets - 1;
	int i = T->bucket_start[b];
	while (i > 0 && T->node_E[i] > E) i--;
	while (i + 1 < T->n_nodes - 1 && T->node_E[i + 1] <= E) i++;
	if (i > T->n_nodes - 2) i = T->n_nodes - 2;
	nodepos_t p;
	p.pos = i;
	p.f = (E - T->node_E[i]) / (T->node_E[i + 1] - T->node_E[i]);
	return p;
}
static double lerp_at(const double *row, nodepos_t p) { return row[p.pos] + (row[p.pos + 1] - row[p.pos]) * p.f; }

/* CS_Total_Kissel(Z,E) etc. from the bundle */
static double cs_total(const ctx_t *c, int Z, nodepos_t p) { return lerp_at(c->T->cs_total + (size_t)c->T->uniqZ[Z] * c->T->n_nodes, p); }

/* xmi_mu_calc for the composition (src/xmi_aux_f.F90:1109-1141) */
static void mu_calc(const ctx_t *c, double E, double *mus) {
	nodepos_t p = node_find(c->T, E);
	for (int i = 0; i < c->nL; i++) {
		const xmb_layer *l = &c->in->composition->layers[i];
		double rv = 0.0;
		for (int j = 0; j < l->n_elements; j++) rv += cs_total(c, l->Z[j], p) * l->weight[j];
		mus[i] = rv;
	}
}

/* bilinear_interpolation on uniform axes (src/xmi_aux_f.F90:1337-1428); findpos semantics :1305-1335 */
static int findpos_uniform(double x0, double dx, int n, double x) {
	if (fabs(x - x0) < 1e-10) return 0;
	int i = (int)ceil((x - x0) / dx) - 1;
	if (i < 0) i = 0;
	if (i > n - 2) i = n - 2;
	/* guard against rounding of the division: enforce axis(i) < x <= axis(i+1) where possible */
	while (i > 0 && x <= x0 + dx * i) i--;
	while (i < n - 2 && x > x0 + dx * (i + 1)) i++;
	return i;
}
static double bilinear(const double *a, int n2, const double *ax1, int n1, const double *ax2, double x1, double x2) {
	/* a[i1][i2] row-major (i2 fastest) */
	int p1 = findpos_uniform(ax1[0], ax1[1] - ax1[0], n1, x1);
	int p2 = findpos_uniform(ax2[0], ax2[1] - ax2[0], n2, x2);
	double denom = (ax1[p1 + 1] - ax1[p1]) * (ax2[p2 + 1] - ax2[p2]);
	double c1 = (ax1[p1 + 1] - x1) * (ax2[p2 + 1] - x2) / denom;
	double c2 = (x1 - ax1[p1]) * (ax2[p2 + 1] - x2) / denom;
	double c3 = (ax1[p1 + 1] - x1) * (x2 - ax2[p2]) / denom;
	double c4 = (x1 - ax1[p1]) * (x2 - ax2[p2]) / denom;
	return c1 * a[(size_t)p1 * n2 + p2] + c2 * a[(size_t)(p1 + 1) * n2 + p2] + c3 * a[(size_t)p1 * n2 + p2 + 1] +
	       c4 * a[(size_t)(p1 + 1) * n2 + p2 + 1];
}

/* ---- small vector helpers (src/xmi_aux_f.F90:1143-1238) ------------------------------------ */
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm3(const double *a) { return sqrt(dot3(a, a)); }
static void normalize3(double *a) { double n = norm3(a); a[0] /= n; a[1] /= n; a[2] /= n; }
static void cross3(const double *a, const double *b, double *c) {
	c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static int plane_line(const double *pp, const double *pn, const double *lp, const double *ld, double *out) {
	double ItimesN = dot3(ld, pn);
	if (ItimesN == 0.0) return 0;
	double diff[3] = {pp[0] - lp[0], pp[1] - lp[1], pp[2] - lp[2]};
	double d = dot3(diff, pn) / ItimesN;
	for (int i = 0; i < 3; i++) out[i] = d * ld[i] + lp[i];
	return 1;
}
static double dist3(const double *a, const double *b) {
	return sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
}
static void matvec(const double *M, const double *v, double *out) {
	for (int i = 0; i < 3; i++) out[i] = M[i * 3] * v[0] + M[i * 3 + 1] * v[1] + M[i * 3 + 2] * v[2];
}

/* standard normal via Box-Muller on two uniforms (reference: ziggurat; only distributional parity
 * is meaningful for these draws) */
static double ran_gaussian(orc_rng *r, double sigma) {
	double u1 = orc_rng_uniform(r), u2 = orc_rng_uniform(r);
	return sigma * sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * M_PI * u2);
}

/* ---- xmi_update_photon_dirv (src/xmi_main.F90:5071-5148) ----------------------------------- */
static void update_dirv(photon_t *p, double theta_i, double phi_i) {
	double phi_new = phi_i;
	if (phi_i > 2.0 * M_PI) phi_new = phi_i - 2.0 * M_PI;
	else if (phi_i < 0.0) phi_new = phi_i + 2.0 * M_PI;
	double cph = cos(p->phi), sph = sin(p->phi), cth = cos(p->theta), sth = sin(p->theta);
	double m[9] = {cth * cph, -sph, sth * cph, cth * sph, cph, sth * sph, -sth, 0.0, cth};
	double ts = sin(theta_i);
	double v[3] = {ts * cos(phi_new), ts * sin(phi_new), cos(theta_i)};
	matvec(m, v, p->dirv);
	normalize3(p->dirv);
	p->theta = acos(p->dirv[2]);
	p->phi = atan2(p->dirv[1], p->dirv[0]);
	if (p->phi > 2.0 * M_PI) p->phi -= 2.0 * M_PI;
	else if (p->phi < 0.0) p->phi += 2.0 * M_PI;
}
/* xmi_update_photon_elecv (:5150-5182) */
static void update_elecv(photon_t *p) {
	double cosalfa = dot3(p->dirv, p->elecv);
	double c_alfa = acos(cosalfa), sinalfa = sin(c_alfa);
	double c_ae = 1.0 / sinalfa, c_be = -c_ae * cosalfa;
	for (int i = 0; i < 3; i++) p->elecv[i] = c_ae * p->elecv[i] + c_be * p->dirv[i];
	normalize3(p->elecv);
}
/* phi0 of the electric vector in the photon frame (:2055-2066, :2148-2159) */
static double elec_phi0(const photon_t *p) {
	double cph = cos(p->phi), sph = sin(p->phi), cth = cos(p->theta), sth = sin(p->theta);
	double a[3] = {cph * cth, cth * sph, -sth}, b[3] = {sph, -cph, 0.0};
	double cosphi0 = dot3(p->elecv, a), sinphi0 = dot3(p->elecv, b);
	if (fabs(cosphi0) > 1.0) cosphi0 = cosphi0 > 0 ? 1.0 : -1.0;
	double phi0 = acos(cosphi0);
	if (sinphi0 > 0.0) phi0 = -phi0;
	return phi0;
}

/* ---- Doppler-broadened Compton energy (src/xmi_main.F90:4985-5067; VR variant
 *      src/xmi_variance_reduction.F90:1010-1101) ------------------------------------------------ */
static double compton_energy(const ctx_t *c, int zi, double E0, double theta_i, substream_t *rng, int varred) {
	const double cc = 1.2399E-6, c0 = 4.85E-12, c1 = 1.456E-2;
	const xmb_tables_host *T = c->T;
	const double *icdf = T->cp_icdf + (size_t)zi * T->n_cp;
	double c_lamb0 = cc / (E0 * 1000.0);
	double sth2 = sin(theta_i / 2.0);
	double energy;
	int tries = 0;
	for (;;) {
		double r = sub_uniform(rng), r_sign = sub_uniform(rng);
		int pos = (int)(r / (T->cp_R[1] - T->cp_R[0]));          /* 0-based INT(r/dr) */
		if (varred && pos == T->n_cp - 2) continue;               /* :1058 skip the last interval */
		if (pos > T->n_cp - 2) pos = T->n_cp - 2;
		double pz = icdf[pos] + (icdf[pos + 1] - icdf[pos]) * (r - T->cp_R[pos]) / (T->cp_R[pos + 1] - T->cp_R[pos]);
		if (r_sign < 0.5) pz = -pz;
		double dlamb = c0 * sth2 * sth2 - c1 * c_lamb0 * sth2 * pz;
		double c_lamb = c_lamb0 + dlamb;
		energy = cc / c_lamb / 1000.0;
		if (energy <= E0) break;
		if (tries == (varred ? 100 : 500)) break;                  /* VR: reference aborts the run here (:1084-1089) */
		tries++;
	}
	return energy;
}


/* ---- shell-resolved ("advanced") Compton: src/xmi_aux_f.F90:1951-2073, src/xmi_main.F90:4785-4983,
 *      src/xmi_variance_reduction.F90:752-947 --------------------------------------------------------- */
static double adv_qimax(double energy, double Ii, double theta) {                  /* xmi_get_qimax */
	if (Ii != 0.0 && energy < Ii) return 0.0;
	double EminIi = energy - Ii, costheta = cos(theta);
	double Q = 137.0 * (EminIi * energy * (1.0 - costheta) / XMI_MEC2 - Ii);
	return Q / sqrt(EminIi * EminIi + energy * energy - 2.0 * EminIi * energy * costheta);
}
static double adv_q_from_energy(double e0, double e1, double theta) {              /* xmi_get_q_from_energy */
	double ct = cos(theta);
	double Q = 137.0 * (e1 - e0 + (1.0 - ct) * e0 * e1 / XMI_MEC2);
	return Q / sqrt(e1 * e1 + e0 * e0 - 2.0 * e0 * e1 * ct);
}
static double adv_energy_from_q(double e0, double Q, double theta) {               /* xmi_get_energy_from_q */
	double a = e0, b = XMI_MEC2, c = cos(theta);
	if (fabs(c - 1.0) < 1E-8) return 0.0;
	if (fabs(Q) < 1E-4) return e0 / (1.0 + e0 * (1.0 - c) / XMI_MEC2);
	double d = 1.0 + a / b - a * c / b;
	double aq = 137.0 * 137.0 * d * d - Q * Q;
	double bq = -2.0 * 137.0 * 137.0 * a * d + 2.0 * a * c * Q * Q;
	double cq = 137.0 * 137.0 * a * a - a * a * Q * Q;
	double E1 = 0.0, E2 = 0.0;
	if (orc_poly_solve_quadratic(aq, bq, cq, &E1, &E2) == 0) return 0.0;
	double Q1 = adv_q_from_energy(e0, E1, theta), Q2 = adv_q_from_energy(e0, E2, theta);
	if (Q * Q1 > 0.0) return E1;
	if (Q * Q2 > 0.0) return E2;
	if (fabs(E1 - E2) < 1E-10 || fabs(Q1 - Q2) < 1E-10) return E1;
	return 0.0;                                                                    /* the reference exits here */
}
/* cumulative probability of the subshell row r up to Qimax (:4808-4848) */
static double adv_shell_cdf(const xmb_tables_host *T, int r, double energy, double theta) {
	const double Qimax = adv_qimax(energy, T->adv_edge[r], theta);
	if (Qimax < -100.0) return 0.0;
	if (Qimax > 100.0) return 1.0;
	const double *cdf = T->adv_cdf + (size_t)r * T->n_cp;
	const double dq = 100.0 / (T->n_cp - 1.0), qa = fabs(Qimax);
	int pos = (int)(qa / dq);
	if (pos > T->n_cp - 2) pos = T->n_cp - 2;
	const double v = cdf[pos] + (cdf[pos + 1] - cdf[pos]) * (qa - dq * pos) / dq;
	return Qimax < 0.0 ? 1.0 - (0.5 + v) : 0.5 + v;
}
/* Q for the cumulative value cdf in [0, cdfs(r)] (:4880-4940) */
static double adv_sample_q(const xmb_tables_host *T, int r, double cdf) {
	const double *qinv = T->adv_qinv + (size_t)r * T->n_cp;
	const double dc = 0.5 / (T->n_cp - 1.0), cp = cdf < 0.5 ? 0.5 - cdf : cdf - 0.5;
	int pos = (int)(cp / dc);
	if (pos > T->n_cp - 2) pos = T->n_cp - 2;
	const double q = qinv[pos] + (qinv[pos + 1] - qinv[pos]) * (cp - dc * pos) / dc;
	return cdf < 0.5 ? -q : q;
}
/* xmi_update_photon_energy_compton (:4785-4983): two draws {subshell, Q} */
static double compton_energy_adv(const ctx_t *c, int zi, double E0, double theta_i, double u_shell, double u_q) {
	const xmb_tables_host *T = c->T;
	const int r0 = T->adv_off[zi], r1 = T->adv_off[zi + 1];
	double cdfs[32], cdf_sum = 0.0;
	for (int r = r0; r < r1; r++) { cdfs[r - r0] = adv_shell_cdf(T, r, E0, theta_i); cdf_sum += T->adv_config[r] * cdfs[r - r0]; }
	if (cdf_sum == 0.0) return 0.0;
	double temp_sum = 0.0;
	int i = r1 - 1;
	for (int r = r0; r < r1; r++) { temp_sum += T->adv_config[r] * cdfs[r - r0] / cdf_sum; if (u_shell <= temp_sum) { i = r; break; } }
	return adv_energy_from_q(E0, adv_sample_q(T, i, u_q * cdfs[i - r0]), theta_i);
}

/* ---- xmi_get_solid_angle (src/xmi_solid_angle_f.F90:712-801) ------------------------------- */
static long g_hits_per_single = 5000;                                        /* src/xmi_solid_angle_f.F90:43 */
void orc_set_hits_per_single(long n) { g_hits_per_single = n > 0 ? n : 5000; }
static double get_solid_angle(ctx_t *c, const photon_t *p) {
	const double *coords = p->coords;
	const xmb_solid_angle *sa = c->sa;
	const xmb_geometry *g = c->in->geometry;
	double r = dist3(g->p_detector_window, coords);
	double dirv[3] = {coords[0] - g->p_detector_window[0], coords[1] - g->p_detector_window[1], coords[2] - g->p_detector_window[2]};
	normalize3(dirv);
	double temp_theta = acos(dot3(dirv, g->n_detector_orientation));
	if (temp_theta > M_PI / 2.0) temp_theta = M_PI - temp_theta;
	double theta = (M_PI / 2.0) - temp_theta;
	if (theta < sa->grid_dims_theta_vals[0]) return 0.0;
	int nr = (int)sa->grid_dims_r_n, nt = (int)sa->grid_dims_theta_n;
	/* findpos = -1 only beyond the last axis value (src/xmi_aux_f.F90:1305-1335; below the first it returns 1 and the
	 * bilinear form extrapolates from the first cell): on-the-fly Monte Carlo with hits_per_single rays (:783-789) */
	if (r > sa->grid_dims_r_vals[nr - 1] || theta > sa->grid_dims_theta_vals[nt - 1]) {
		c->sa_not_found++;
		return orc_single_solid_angle_photon(c->d, r, theta, g_hits_per_single, p->seed, p->g, p->n_interactions, NULL);
	}
	/* Fortran array solid_angles(r, theta) -> C [theta][r]; bilinear with x1 = r, x2 = theta */
	int p1 = findpos_uniform(sa->grid_dims_r_vals[0], sa->grid_dims_r_vals[1] - sa->grid_dims_r_vals[0], nr, r);
	int p2 = findpos_uniform(sa->grid_dims_theta_vals[0], sa->grid_dims_theta_vals[1] - sa->grid_dims_theta_vals[0], nt, theta);
	const double *R = sa->grid_dims_r_vals, *Th = sa->grid_dims_theta_vals, *A = sa->solid_angles;
	double denom = (R[p1 + 1] - R[p1]) * (Th[p2 + 1] - Th[p2]);
	double c1 = (R[p1 + 1] - r) * (Th[p2 + 1] - theta) / denom, c2 = (r - R[p1]) * (Th[p2 + 1] - theta) / denom;
	double c3 = (R[p1 + 1] - r) * (theta - Th[p2]) / denom, c4 = (r - R[p1]) * (theta - Th[p2]) / denom;
	return c1 * A[(size_t)p2 * nr + p1] + c2 * A[(size_t)p2 * nr + p1 + 1] + c3 * A[(size_t)(p2 + 1) * nr + p1] +
	       c4 * A[(size_t)(p2 + 1) * nr + p1 + 1];
}

static void deposit(ctx_t *c, int Z, int slot, int n_ia, double energy, double w) {
	/* var_red_history(Z, slot, n_ia) += w ; channels(n_ia:, ch) += w   (src/xmi_variance_reduction.F90:353-369) */
	c->var_red[((size_t)(n_ia - 1) * 385 + (slot - 1)) * 100 + (Z - 1)] += w;
	int channel = -1;
	if (energy >= ENERGY_THRESHOLD) channel = (int)((energy - c->in->detector->zero) / c->in->detector->gain);
	if (channel >= 0 && channel <= c->nch - 1)
		for (int k = n_ia; k <= c->n_int; k++) c->channels[(size_t)k * c->nch + channel] += w;
}

/* ---- xmi_variance_reduction (src/xmi_variance_reduction.F90:29-726) ------------------------ */
static void variance_reduction(ctx_t *c, photon_t *p, double u_det_r, double u_det_phi) {
	const xmb_tables_host *T = c->T;
	const orc_derived *D = c->d;
	const xmb_geometry *g = c->in->geometry;
	if (p->energy <= ENERGY_THRESHOLD) return;                                   /* :79 */
	double radius = sqrt(u_det_r) * D->detector_radius;                           /* :91 */
	double theta = 2.0 * M_PI * u_det_phi;
	double detector_point[3] = {0.0, cos(theta) * radius, sin(theta) * radius};
	double rel[3] = {p->coords[0] - g->p_detector_window[0], p->coords[1] - g->p_detector_window[1], p->coords[2] - g->p_detector_window[2]};
	double lc_point[3], dirv[3], lc_dirv[3];
	matvec(D->ndo_inv, rel, lc_point);                                            /* :107 */
	matvec(D->ndo_inv, p->dirv, dirv);
	for (int i = 0; i < 3; i++) lc_dirv[i] = detector_point[i] - lc_point[i];
	if (lc_dirv[0] >= 0.0) return;                                                /* :114 heading away */
	double total_distance = dist3(detector_point, lc_point);
	normalize3(dirv);
	normalize3(lc_dirv);
	double new_dirv_coords[3];
	matvec(D->ndo_new, lc_dirv, new_dirv_coords);                                 /* :151 */
	double dotprod = dot3(dirv, lc_dirv);
	if (dotprod > 1.0) dotprod = 1.0; else if (dotprod < -1.0) dotprod = -1.0;
	theta = acos(dotprod);                                                        /* :166 scattering angle */
	double proj[3], dp = dot3(new_dirv_coords, p->dirv);
	for (int i = 0; i < 3; i++) proj[i] = new_dirv_coords[i] - dp * p->dirv[i];    /* :198 */
	normalize3(proj);
	double en = norm3(p->elecv), elecv_norm[3] = {p->elecv[0] / en, p->elecv[1] / en, p->elecv[2] / en};
	dotprod = dot3(proj, elecv_norm);
	if (dotprod > 1.0) dotprod = 1.0; else if (dotprod < -1.0) dotprod = -1.0;
	double phi = acos(dotprod);                                                   /* :212 */
	int step_max, step_dir;
	if (dot3(new_dirv_coords, g->n_sample_orientation) > 0.0) { step_max = c->nL - 1; step_dir = 1; }
	else { step_max = 0; step_dir = -1; }
	double distances[64];
	for (int i = 0; i < c->nL; i++) distances[i] = 0.0;
	double temp_coords[3] = {p->coords[0], p->coords[1], p->coords[2]};
	for (int i = p->current_layer; step_dir > 0 ? i <= step_max : i >= step_max; i += step_dir) {   /* :253-284 */
		double pp[3] = {0.0, 0.0, step_dir == 1 ? D->Z_coord_end[i] : D->Z_coord_begin[i]}, inter[3];
		if (!plane_line(pp, g->n_sample_orientation, temp_coords, new_dirv_coords, inter)) return;
		distances[i] = dist3(temp_coords, inter);
		if (distances[i] > total_distance) { distances[i] = total_distance; break; }
		memcpy(temp_coords, inter, sizeof(inter));
		total_distance -= distances[i];
	}
	const xmb_layer *layer = &c->in->composition->layers[p->current_layer];
	const xmb_layer *layers = c->in->composition->layers;
	int n_ia = p->n_interactions;
	double temp_murhod = 0.0;
	for (int j = p->current_layer; step_dir > 0 ? j <= step_max : j >= step_max; j += step_dir)
		temp_murhod += p->mus[j] * layers[j].density * distances[j];
	double Pesc_rayl = exp(-temp_murhod);                                         /* :306 */
	double detector_solid_angle = get_solid_angle(c, p);                  /* :318 */
	double Pdir_fluo = detector_solid_angle / 4.0 / M_PI;
	int line_last = c->opt->use_M_lines ? XMB_M5P5 : XMB_L3Q1;
	nodepos_t np = node_find(T, p->energy);
	double q = p->energy / KEV2ANGST * sin(theta / 2.0);
	double qx = q / T->q_max * (T->n_q - 1);
	int qi = (int)qx;
	if (qi > T->n_q - 2) qi = T->n_q - 2;
	double qf = qx - qi;
	for (int i = 0; i < layer->n_elements; i++) {                                 /* var_red: :338 */
		int Z = layer->Z[i], zi = T->uniqZ[Z];
		/* RAYLEIGH  (:342-369)  DCSP_Rayl = N_A/A F^2 r_e^2 (1 - sin^2 theta cos^2 phi) */
		double Pconv = layer->weight[i] / p->mus[p->current_layer];
		double F = T->ff[(size_t)zi * T->n_q + qi] * (1.0 - qf) + T->ff[(size_t)zi * T->n_q + qi + 1] * qf;
		double dcsp_rayl = AVOGNUM / T->atomic_weight[zi] * F * F * RE2 * (1.0 - sin(theta) * sin(theta) * cos(phi) * cos(phi));
		double Pdir = detector_solid_angle * dcsp_rayl;
		deposit(c, Z, 383 + 1, n_ia, p->energy, Pconv * Pdir * Pesc_rayl * p->weight);
		/* COMPTON  (xmi_compton_varred2, :949-1008; shell-resolved xmi_compton_varred, :752-947) */
		{
			double S = T->sf[(size_t)zi * T->n_q + qi] * (1.0 - qf) + T->sf[(size_t)zi * T->n_q + qi + 1] * qf;
			double k0k = 1.0 / (1.0 + (1.0 - cos(theta)) * p->energy / 510.998928);
			double dcsp_kn = RE2 / 2.0 * k0k * k0k * (k0k + 1.0 / k0k - 2.0 * sin(theta) * sin(theta) * cos(phi) * cos(phi));
			double Pdir_c = detector_solid_angle * AVOGNUM / T->atomic_weight[zi] * S * dcsp_kn;
			if (c->opt->use_advanced_compton) {
				const int r0 = T->adv_off[zi], r1 = T->adv_off[zi + 1];
				double cdfs[32], cdf_sum = 0.0;
				for (int r = r0; r < r1; r++) { cdfs[r - r0] = adv_shell_cdf(T, r, p->energy, theta); cdf_sum += T->adv_config[r] * cdfs[r - r0]; }
				for (int r = r0; r < r1 && cdf_sum != 0.0; r++) {
					double shell_weight = T->adv_config[r] * cdfs[r - r0] / cdf_sum;
					if (shell_weight == 0.0) continue;
					double u[4];
					draw_block(p->seed, p->g, n_ia, 2, i, (r - r0) >> 2, u);          /* one word per subshell */
					double e_c = adv_energy_from_q(p->energy, adv_sample_q(T, r, u[(r - r0) & 3] * cdfs[r - r0]), theta);
					if (e_c == 0.0) continue;
					double mus_c[64], tm = 0.0;
					mu_calc(c, e_c, mus_c);
					for (int j = p->current_layer; step_dir > 0 ? j <= step_max : j >= step_max; j += step_dir)
						tm += mus_c[j] * layers[j].density * distances[j];
					deposit(c, Z, 383 + 2, n_ia, e_c, Pconv * Pdir_c * exp(-tm) * p->weight * shell_weight);
				}
			} else {
				substream_t cs;
				sub_init(&cs, p->seed, p->g, n_ia, 2, i);
				double e_c = compton_energy(c, zi, p->energy, theta, &cs, 1);
				double mus_c[64];
				mu_calc(c, e_c, mus_c);
				double tm = 0.0;
				for (int j = p->current_layer; step_dir > 0 ? j <= step_max : j >= step_max; j += step_dir)
					tm += mus_c[j] * layers[j].density * distances[j];
				double Pesc_comp = exp(-tm);
				deposit(c, Z, 383 + 2, n_ia, e_c, Pconv * Pdir_c * Pesc_comp * p->weight);
			}
		}
		/* FLUORESCENCE  (:391-709).  Shell vacancy cross sections under the selected cascade mode at the
		 * photon energy; after a fluorescence interaction that energy is a line energy, which is a node of
		 * the table grid, so this lookup is the reference's precalc_xrf_cs (:392-433). */
		double P[9];
		for (int s = 0; s < 9; s++) P[s] = lerp_at(T->cs_vacancy + (((size_t)(c->cascade - 1) * T->nZ + zi) * 9 + s) * T->n_nodes, np);
		if (!(p->energy >= T->edge_energy[zi * 9 + 0])) P[0] = 0.0;               /* :445 */
		if (!c->opt->use_M_lines) for (int s = 4; s < 9; s++) P[s] = 0.0;
		for (int l = 1; l <= line_last; l++) {                                    /* :575 */
			double energy_fluo = T->line_energy[(size_t)zi * 384 + l];
			if (energy_fluo < ENERGY_THRESHOLD) continue;
			int shell;
			if (l >= 1 && l <= 29) shell = 0;                                     /* KP5..KL1 */
			else if (l >= XMB_L1M1 && l <= 58) shell = 1;                         /* L1P5..L1M1 */
			else if (l >= XMB_L2M1 && l <= 85) shell = 2;                         /* L2Q1..L2M1 */
			else if (l >= 86 && l <= 113) shell = 3;                              /* L3Q1..L3M1 */
			else if (l >= 118 && l <= 136) shell = 4;
			else if (l >= 140 && l <= 158) shell = 5;
			else if (l >= 161 && l <= 180) shell = 6;
			else if (l >= 182 && l <= 200) shell = 7;
			else if (l >= 201 && l <= 219) shell = 8;
			else continue;
			if (P[shell] == 0.0) continue;
			double Pc = layer->weight[i] * P[shell] * T->fluor_yield[zi * 9 + shell] * T->rad_rate[(size_t)zi * 384 + l] /
			            p->mus[p->current_layer];
			nodepos_t lp = node_find(T, energy_fluo);
			double tm = 0.0;
			for (int j = p->current_layer; step_dir > 0 ? j <= step_max : j >= step_max; j += step_dir)
				tm += lerp_at(T->mu_layer + (size_t)j * T->n_nodes, lp) * layers[j].density * distances[j];   /* precalc_mu_cs */
			double Pesc = exp(-tm);
			double tw = Pc * Pdir_fluo * Pesc * p->weight;
			if (tw == 0.0) continue;
			deposit(c, Z, l, n_ia, energy_fluo, tw);
		}
	}
}

/* ---- interactions --------------------------------------------------------------------------- */
static int do_rayleigh(ctx_t *c, photon_t *p, const double *sd) {                  /* src/xmi_main.F90:1986-2101 */
	const xmb_tables_host *T = c->T;
	int zi = T->uniqZ[p->current_element];
	double r = sd[0];
	double theta_i = bilinear(T->rayl_theta_icdf + (size_t)zi * T->n_icdf_E * T->n_icdf_R, T->n_icdf_R, T->icdf_E, T->n_icdf_E,
	                          T->icdf_R, p->energy, r);
	double tt = sin(theta_i) * sin(theta_i);
	tt = tt / (4.0 - 2.0 * tt);
	double phi_i = bilinear(T->phi_icdf, T->n_icdf_R, T->phi_T, T->n_phi_T, T->icdf_R, tt, sd[1]);
	double phi0 = elec_phi0(p);
	update_dirv(p, theta_i, phi0 + phi_i);
	update_elecv(p);
	p->hist_line[p->n_interactions] = RAYLEIGH;
	p->hist_Z[p->n_interactions] = p->current_element;
	return 1;
}

static int do_compton(ctx_t *c, photon_t *p, const double *sd) {                   /* :2103-2229 */
	const xmb_tables_host *T = c->T;
	int zi = T->uniqZ[p->current_element];
	double theta_i = bilinear(T->compt_theta_icdf + (size_t)zi * T->n_icdf_E * T->n_icdf_R, T->n_icdf_R, T->icdf_E, T->n_icdf_E,
	                          T->icdf_R, p->energy, sd[0]);
	double K0K = 1.0 + p->energy * (1.0 - cos(theta_i)) / XMI_MEC2;
	double tt = sin(theta_i) * sin(theta_i);
	tt = tt / (K0K + (1.0 / K0K) - tt) / 2.0;
	double phi_i = bilinear(T->phi_icdf, T->n_icdf_R, T->phi_T, T->n_phi_T, T->icdf_R, tt, sd[1]);
	double phi0 = elec_phi0(p);
	substream_t ds;
	sub_init(&ds, p->seed, p->g, p->n_interactions | p->gen, 3, 0);
	if (c->opt->use_advanced_compton) {
		double u_shell = sub_uniform(&ds), u_q = sub_uniform(&ds);
		p->energy = compton_energy_adv(c, zi, p->energy, theta_i, u_shell, u_q);
	} else
		p->energy = compton_energy(c, zi, p->energy, theta_i, &ds, 0);
	mu_calc(c, p->energy, p->mus);                                                /* :5059 */
	if (p->energy == 0.0) return 1;
	update_dirv(p, theta_i, phi_i + phi0);
	update_elecv(p);
	double pp = 2.0 * (pow(cos(theta_i) * cos(phi_i), 2) + pow(sin(phi_i), 2));   /* :2201-2211 */
	double rat = 1.0 / (1.0 + (1 - cos(theta_i)) * p->energy / 510.998910);
	double rk = rat - 2.0 + 1.0 / rat;
	pp = pp / (rk + pp);
	double r = sd[2];
	double w_h = (1.0 + pp) / 2.0;
	if (r > w_h) { double t[3]; cross3(p->dirv, p->elecv, t); memcpy(p->elecv, t, sizeof(t)); }
	p->hist_line[p->n_interactions] = COMPTON;
	p->hist_Z[p->n_interactions] = p->current_element;
	return 1;
}

/* xmi_coster_kronig_check (:5184-5323) */
static int coster_kronig(const ctx_t *c, int zi, int shell, substream_t *rng) {
	const double *ck = c->T->cos_kron + (size_t)zi * XMB_N_CK;
	static const int first[9] = {-1, XMB_FL12, XMB_FL23, -1, XMB_FM12, XMB_FM23, XMB_FM34, XMB_FM45, -1};
	static const int ntr[9] = {0, 2, 1, 0, 4, 3, 2, 1, 0};
	while (shell == 1 || shell == 2 || (shell >= 4 && shell <= 7)) {
		double r = sub_uniform(rng), sumz = 0.0;
		int found = -1;
		for (int t = 0; t < ntr[shell]; t++) {
			sumz += ck[first[shell] + t];
			if (r < sumz) { found = t; break; }
		}
		if (found < 0) break;              /* nothing happened */
		shell = shell + 1 + found;         /* L1: t=0 -> L2, t=1 -> L3 ; M1: t -> M(2+t) ; ... */
		/* the reference exits after landing in the last subshell of the group or after a single-option hop */
	}
	return shell;
}

static int do_photo(ctx_t *c, photon_t *p, const double *sd) {                     /* :2231-2411 */
	const xmb_tables_host *T = c->T;
	int Z = p->current_element, zi = T->uniqZ[Z];
	nodepos_t np = node_find(T, p->energy);
	double photo_total = lerp_at(T->cs_photo_total + (size_t)zi * T->n_nodes, np);
	double sumz = 0.0, r = sd[0];
	int max_shell = c->opt->use_M_lines ? 8 : 3, shell, shell_found = 0;
	for (shell = 0; shell <= max_shell; shell++) {
		sumz += lerp_at(T->cs_photo_partial + ((size_t)zi * 9 + shell) * T->n_nodes, np) / photo_total;
		if (r < sumz) { shell_found = 1; break; }
	}
	if (!shell_found) { p->energy = 0.0; return 1; }                              /* :2280-2291 */
	/* the reference draws one unused number here (xmi_variance_reduction.F90:737); not reproduced */
	if (c->escape_mode) p->weight_escape *= T->fluor_yield_corr[zi * 9 + shell];  /* xmi_variance_reduction.F90:740-744 */
	else p->weight *= T->fluor_yield_corr[zi * 9 + shell];                        /* :745 */
	substream_t xs;
	sub_init(&xs, p->seed, p->g, p->n_interactions, 3, 0);
	double u_phi = sub_uniform(&xs);
	shell = coster_kronig(c, zi, shell, &xs);                                     /* :2329 */
	/* xmi_fluorescence_line_check (:5352-5437) */
	r = sd[1];
	sumz = 0.0;
	int line = 0;
	for (int l = xmb_shell_line_first[shell]; l <= xmb_shell_line_last[shell]; l++) {
		sumz += T->rad_rate[(size_t)zi * 384 + l];
		if (r < sumz) { line = l; break; }
	}
	if (!line) { p->energy = 0.0; return 1; }                                     /* :5421-5426, then :2336-2340 */
	p->energy = T->line_energy[(size_t)zi * 384 + line];
	{
		nodepos_t lp = node_find(T, p->energy);                                   /* precalc_mu_cs, :2346-2349 */
		for (int i = 0; i < c->nL; i++) p->mus[i] = lerp_at(T->mu_layer + (size_t)i * T->n_nodes, lp);
	}
	double theta_i = acos(-2.0 * sd[2] + 1.0);                                    /* :2352-2353 */
	double phi_i = 2.0 * M_PI * u_phi;
	update_dirv(p, theta_i, phi_i);
	update_elecv(p);
	p->hist_line[p->n_interactions] = -line;
	p->hist_Z[p->n_interactions] = Z;
	return 1;
}

/* ---- xmi_simulate_photon, variance-reduction branch (src/xmi_main.F90:1188-1685) ------------ */
static void simulate_photon(ctx_t *c, photon_t *p) {
	const xmb_geometry *g = c->in->geometry;
	const orc_derived *D = c->d;
	const xmb_layer *layers = c->in->composition->layers;
	for (;;) {
		if (p->energy < ENERGY_THRESHOLD) break;                                   /* :1229 */
		int step_max, step_dir;
		if (dot3(p->dirv, g->n_sample_orientation) > 0.0) { step_max = c->nL - 1; step_dir = 1; }
		else { step_max = 0; step_dir = -1; }
		if (p->n_interactions == c->n_int) break;                                  /* :1417-1420 */
		double b0[4], b1[4];
		draw_block(p->seed, p->g, p->n_interactions + 1, 1, 0, 0, b0);
		draw_block(p->seed, p->g, p->n_interactions + 1, 1, 0, 1, b1);
		double interactionR = b0[0];                                               /* :1264 */
		double distances[64], lp[3] = {p->coords[0], p->coords[1], p->coords[2]};
		for (int i = p->current_layer; step_dir > 0 ? i <= step_max : i >= step_max; i += step_dir) {   /* :1429-1449 */
			double pp[3] = {0.0, 0.0, step_dir == 1 ? D->Z_coord_end[i] : D->Z_coord_begin[i]}, inter[3];
			if (!plane_line(pp, g->n_sample_orientation, lp, p->dirv, inter)) return;
			distances[i] = dist3(lp, inter);
			memcpy(lp, inter, sizeof(inter));
		}
		double Pabs = 0.0;
		for (int i = p->current_layer; step_dir > 0 ? i <= step_max : i >= step_max; i += step_dir)
			Pabs += p->mus[i] * layers[i].density * distances[i];
		double Pabs2 = -1.0 * expm1(-1.0 * Pabs);                                  /* :1460 */
		p->weight *= Pabs2;
		double negln = -1.0 * log1p(-1.0 * interactionR * Pabs2);                  /* :1468 */
		int my_index = p->current_layer;
		double my_sum = 0.0;
		for (int i = p->current_layer; step_dir > 0 ? i <= step_max : i >= step_max; i += step_dir) {
			my_sum += p->mus[i] * layers[i].density * distances[i];
			if (my_sum > negln) { my_index = i; break; }
		}
		double temp_sum = 0.0;
		for (int i = p->current_layer; step_dir > 0 ? i <= my_index : i >= my_index; i += step_dir)   /* :1497-1502 */
			temp_sum += (1.0 - (p->mus[i] * layers[i].density / (p->mus[my_index] * layers[my_index].density))) * distances[i];
		temp_sum = temp_sum - 1.0 * log1p(-1.0 * interactionR * Pabs2) / (p->mus[my_index] * layers[my_index].density);
		for (int i = 0; i < 3; i++) p->coords[i] += temp_sum * p->dirv[i];          /* :1507 */
		p->current_layer = my_index;
		p->n_interactions++;                                                       /* :1542 */
		c->n_interactions_total++;
		variance_reduction(c, p, b0[1], b0[2]);                                    /* :1548 */
		/* atom selection (:1558-1572) */
		interactionR = b0[3];
		{
			const xmb_layer *l = &layers[p->current_layer];
			nodepos_t np = node_find(c->T, p->energy);
			double thr = 0.0;
			for (int i = 0; i < l->n_elements; i++) {
				thr += l->weight[i] * cs_total(c, l->Z[i], np) / p->mus[p->current_layer];
				if (interactionR < thr || i == l->n_elements - 1) {   /* last element absorbs rounding (reference keeps the stale element) */
					p->current_element = l->Z[i];
					p->current_element_index = i;
					break;
				}
			}
			/* interaction type (:1580-1652) */
			interactionR = b1[0];
			int zi = c->T->uniqZ[p->current_element];
			double pr = lerp_at(c->T->p_rayl + (size_t)zi * c->T->n_nodes, np);
			double prc = lerp_at(c->T->p_rayl_compt + (size_t)zi * c->T->n_nodes, np);
			if (interactionR < pr) { p->last_interaction = RAYLEIGH; do_rayleigh(c, p, b1 + 1); }
			else if (interactionR < prc) { p->last_interaction = COMPTON; do_compton(c, p, b1 + 1); }
			else { p->last_interaction = PHOTO; do_photo(c, p, b1 + 1); }
		}
	}
}

/* ---- source sampling and driver (src/xmi_main.F90:280-867) ---------------------------------- */
typedef struct { int is_cont; int idx; uint64_t first; uint64_t count; } segment_t;

static void start_photon(ctx_t *c, photon_t *p, orc_rng *rng, const segment_t *seg, uint64_t j /* 0-based within segment */, int *skip) {
	const xmb_input *in = c->in;
	const xmb_geometry *g = in->geometry;
	const xmb_excitation *exc = in->excitation;
	const xmb_tables_host *T = c->T;
	memset(p, 0, sizeof(*p));
	*skip = 0;
	double weight, hor_ver_ratio = 0.0, sx, sy, sxp, syp;
	if (seg->is_cont) {                                                            /* :319-438 */
		const xmb_energy_continuous *a = &exc->continuous[seg->idx], *b = &exc->continuous[seg->idx + 1];
		double x1 = a->energy, x2 = b->energy, y1 = a->vertical_intensity + a->horizontal_intensity,
		       y2 = b->vertical_intensity + b->horizontal_intensity;
		double total_intensity = (y1 + y2) * (x2 - x1) / 2.0;
		/* xmi_ran_trap (src/xmi_aux_f.F90:1841-1941) */
		double m = (y2 - y1) / (x2 - x1);
		double denom = (x2 - x1) * (y1 - x1 * m) + m * (x2 * x2 - x1 * x1) / 2.0;
		double qa = m / 2.0, qb = y1 - x1 * m, qc = -x1 * y1 + m * x1 * x1 / 2.0 - denom * orc_rng_uniform(rng);
		double rv1 = 0, rv2 = 0;
		orc_poly_solve_quadratic(qa, qb, qc, &rv1, &rv2);
		if (x1 <= rv1 && rv1 <= x2) p->energy = rv1; else p->energy = rv2;
		double hi = a->horizontal_intensity + (b->horizontal_intensity - a->horizontal_intensity) * (p->energy - x1) / (x2 - x1);
		double ti = y1 + (y2 - y1) * (p->energy - x1) / (x2 - x1);
		hor_ver_ratio = hi / ti;                                                  /* :359-364 */
		double exc_corr = exp(-lerp_at(T->exc_murhod, node_find(T, p->energy)));  /* :366-372 */
		weight = total_intensity * exc_corr / in->general->n_photons_interval;    /* :378 */
		mu_calc(c, p->energy, p->mus);
		sx = a->sigma_x; sy = a->sigma_y; sxp = a->sigma_xp; syp = a->sigma_yp;
	} else {                                                                       /* :579-724 */
		const xmb_energy_discrete *e = &exc->discrete[seg->idx];
		double total_intensity = e->vertical_intensity + e->horizontal_intensity;
		hor_ver_ratio = e->horizontal_intensity * (double)in->general->n_photons_line / total_intensity;   /* :582, on the global index */
		double exc_corr = exp(-lerp_at(T->exc_murhod, node_find(T, e->energy)));
		weight = total_intensity * exc_corr / in->general->n_photons_line;        /* :601 */
		if (e->distribution_type == XMB_DISCRETE_GAUSSIAN) p->energy = ran_gaussian(rng, e->scale_parameter) + e->energy;
		else if (e->distribution_type == XMB_DISCRETE_LORENTZIAN) {
			double u = orc_rng_uniform(rng);
			p->energy = e->scale_parameter * tan(M_PI * u) + e->energy;           /* gsl_ran_cauchy */
		} else p->energy = e->energy;
		if (p->energy <= ENERGY_THRESHOLD || p->energy > ENERGY_MAX) { *skip = 1; return; }   /* :639-641 */
		mu_calc(c, p->energy, p->mus);
		sx = e->sigma_x; sy = e->sigma_y; sxp = e->sigma_xp; syp = e->sigma_yp;
	}
	/* xmi_coords_dir (:957-1136) */
	double x1, y1;
	if (fabs(sx * sy) < 1.0E-20) {
		double x1_max = atan(g->slit_size_x / g->d_source_slit / 2.0), y1_max = atan(g->slit_size_y / g->d_source_slit / 2.0);
		x1 = x1_max * (-1.0 + 2.0 * orc_rng_uniform(rng));                        /* ran_flat(-1, 1) */
		y1 = y1_max * (-1.0 + 2.0 * orc_rng_uniform(rng));
		p->coords[0] = p->coords[1] = p->coords[2] = 0.0;
	} else {
		x1 = ran_gaussian(rng, sxp);
		y1 = ran_gaussian(rng, syp);
		p->coords[0] = ran_gaussian(rng, sx) - g->d_source_slit * sin(x1);
		p->coords[1] = ran_gaussian(rng, sy) - g->d_source_slit * sin(y1);
		p->coords[2] = 0.0;
	}
	p->dirv[0] = tan(x1); p->dirv[1] = tan(y1); p->dirv[2] = 1.0;
	normalize3(p->dirv);
	p->theta = acos(p->dirv[2]);
	p->phi = atan2(p->dirv[1], p->dirv[0]);
	/* polarisation (:396-414 continuous: random; :682 discrete: by index) */
	int horizontal;
	if (seg->is_cont) horizontal = orc_rng_uniform(rng) <= hor_ver_ratio;
	else horizontal = (double)(j + 1) <= hor_ver_ratio;
	p->weight = weight;
	if (horizontal) { p->elecv[0] = 0.0; p->elecv[1] = 1.0; p->elecv[2] = 0.0; }
	else { p->elecv[0] = 1.0; p->elecv[1] = 0.0; p->elecv[2] = 0.0; }
	double cosalfa = dot3(p->elecv, p->dirv);
	double c_alfa = acos(cosalfa), c_ae = 1.0 / sin(c_alfa), c_be = -c_ae * cosalfa;
	for (int i = 0; i < 3; i++) p->elecv[i] = c_ae * p->elecv[i] + c_be * p->dirv[i];
	/* xmi_photon_shift_first_layer (:1140-1186) */
	const orc_derived *D = c->d;
	p->current_layer = -1;
	if (p->coords[2] >= D->Z_coord_begin[0]) {
		for (int i = 0; i < c->nL; i++) if (p->coords[2] < D->Z_coord_end[i]) { p->current_layer = i; break; }
		if (p->current_layer < 0) { *skip = 1; return; }
	} else {
		double pp[3] = {0.0, 0.0, D->Z_coord_begin[0]}, inter[3];
		if (!plane_line(pp, g->n_sample_orientation, p->coords, p->dirv, inter)) { *skip = 1; return; }
		memcpy(p->coords, inter, sizeof(inter));
		p->current_layer = 0;
	}
}

static int build_segments(const xmb_input *in, segment_t **out) {
	const xmb_excitation *exc = in->excitation;
	int n = 0, cap = exc->n_continuous + exc->n_discrete + 1;
	segment_t *s = (segment_t *)calloc(cap, sizeof(segment_t));
	uint64_t first = 0;
	for (int i = 0; i + 1 < exc->n_continuous; i++) {
		double y1 = exc->continuous[i].vertical_intensity + exc->continuous[i].horizontal_intensity;
		double y2 = exc->continuous[i + 1].vertical_intensity + exc->continuous[i + 1].horizontal_intensity;
		double total = (y1 + y2) * (exc->continuous[i + 1].energy - exc->continuous[i].energy) / 2.0;
		if (total == 0.0) continue;                                               /* :331-333 */
		s[n].is_cont = 1; s[n].idx = i; s[n].first = first; s[n].count = (uint64_t)in->general->n_photons_interval;
		first += s[n].count; n++;
	}
	for (int i = 0; i < exc->n_discrete; i++) {
		s[n].is_cont = 0; s[n].idx = i; s[n].first = first; s[n].count = (uint64_t)in->general->n_photons_line;
		first += s[n].count; n++;
	}
	*out = s;
	return n;
}

uint64_t orc_total_histories(const xmb_input *in) {
	segment_t *s;
	int n = build_segments(in, &s);
	uint64_t t = n ? s[n - 1].first + s[n - 1].count : 0;
	free(s);
	return t;
}

/* Simulates global photon ids [g_begin, g_end).  Outputs are RAW sums (no live_time), layouts:
 *   channels[(n_int+1)][nch] cumulative over interaction order (as the reference's channels(0:n_int, :))
 *   var_red[n_int][385][100]  (k, slot, Z-1)  -- Fortran var_red_history(Z, slot, k)
 * Returns the number of histories run. */
static uint64_t main_msim_sharded(const xmb_input *in, const orc_derived *d, const xmb_tables_host *T, const xmb_main_options *opt,
                                  const xmb_solid_angle *sa, uint64_t seed, uint64_t g_begin, uint64_t g_end, int shard_rank, int shard_n,
                                  int n_threads, double *channels, double *var_red, uint64_t *counters);

uint64_t orc_main_msim_range(const xmb_input *in, const orc_derived *d, const xmb_tables_host *T, const xmb_main_options *opt,
                             const xmb_solid_angle *sa, uint64_t seed, uint64_t g_begin, uint64_t g_end, int n_threads,
                             double *channels, double *var_red, uint64_t *counters /* [2]: sa_not_found, interactions */) {
	return main_msim_sharded(in, d, T, opt, sa, seed, g_begin, g_end, 0, 1, n_threads, channels, var_red, counters);
}

/* The block-cyclic shard of a rank (blocks of 1024 photon ids, block b to rank b % n_ranks: every rank simulates the
 * same share of every source line, as the reference's MPI split does, src/xmi_main.F90:314,574). */
uint64_t orc_main_msim_shard(const xmb_input *in, const orc_derived *d, const xmb_tables_host *T, const xmb_main_options *opt,
                             const xmb_solid_angle *sa, uint64_t seed, int rank, int n_ranks, int n_threads,
                             double *channels, double *var_red, uint64_t *counters) {
	return main_msim_sharded(in, d, T, opt, sa, seed, 0, orc_total_histories(in), rank, n_ranks, n_threads, channels, var_red, counters);
}

static uint64_t main_msim_sharded(const xmb_input *in, const orc_derived *d, const xmb_tables_host *T, const xmb_main_options *opt,
                                  const xmb_solid_angle *sa, uint64_t seed, uint64_t g_begin, uint64_t g_end, int shard_rank, int shard_n,
                                  int n_threads, double *channels, double *var_red, uint64_t *counters) {
	segment_t *segs;
	int nseg = build_segments(in, &segs);
	int n_int = in->general->n_interactions_trajectory, nch = in->detector->nchannels, nL = in->composition->n_layers;
	int cascade = (opt->use_cascade_auger ? 1 : 0) + (opt->use_cascade_radiative ? 2 : 0) + 1;   /* none 1, auger 2, rad 3, full 4 (:141-153) */
	size_t nchn = (size_t)(n_int + 1) * nch, nvr = (size_t)n_int * 385 * 100;
	memset(channels, 0, sizeof(double) * nchn);
	memset(var_red, 0, sizeof(double) * nvr);
	if (n_threads < 1) n_threads = 1;
	uint64_t c0 = 0, c1 = 0, n_run = 0;
#pragma omp parallel num_threads(n_threads) reduction(+ : n_run)
	{
		ctx_t c;
		memset(&c, 0, sizeof(c));
		c.in = in; c.d = d; c.T = T; c.opt = opt; c.sa = sa; c.cascade = cascade; c.n_int = n_int; c.nch = nch; c.nL = nL;
		c.channels = (double *)calloc(nchn, sizeof(double));
		c.var_red = (double *)calloc(nvr, sizeof(double));
#pragma omp for schedule(dynamic, 256)
		for (uint64_t gidx = g_begin; gidx < g_end; gidx++) {
			if ((int)((gidx >> 10) % (uint64_t)shard_n) != shard_rank) continue;
			n_run++;
			int s = 0;
			while (s + 1 < nseg && gidx >= segs[s + 1].first) s++;
			photon_t p;
			orc_rng rng;
			orc_rng_init(&rng, seed, gidx, ORC_TAG_HISTORY);   /* order 0, stage 0: ctr[2] = block */
			int skip;
			start_photon(&c, &p, &rng, &segs[s], gidx - segs[s].first, &skip);
			p.seed = seed; p.g = gidx;
			if (!skip) simulate_photon(&c, &p);
		}
#pragma omp critical
		{
			for (size_t i = 0; i < nchn; i++) channels[i] += c.channels[i];
			for (size_t i = 0; i < nvr; i++) var_red[i] += c.var_red[i];
			c0 += c.sa_not_found; c1 += c.n_interactions_total;
		}
		free(c.channels); free(c.var_red);
	}
	if (counters) { counters[0] = c0; counters[1] = c1; }
	free(segs);
	return n_run;
}

/* ---- escape-peak ratios (src/xmi_main.F90:5473-5801; driver src/xmi_detector.c:91-141) ---------------------
 * `in` is the escape-mode input (composition = crystal, pencil beam, one interaction; its discrete lines are
 * the input energies).  Streams: g = energy index * n_photons + j; order 0 = source draws (slit x, slit y,
 * polarisation angle), order 1 = the forced interaction, order 2 stage 1 block 0 word 0 = analogue free path.
 * fluo[(i*109 + line-1)*nZ + zi], compt[c*nE + i] (the Fortran layouts), both divided by photons_interacted. */
int orc_escape_ratios(const xmb_input *in, const orc_derived *d, const xmb_tables_host *T, uint64_t seed, long n_photons,
                      int n_out, double out_min, double out_delta, int n_threads, double *fluo, double *compt) {
	const xmb_geometry *g = in->geometry;
	const xmb_layer *layers = in->composition->layers;
	const int nE = in->excitation->n_discrete, nL = in->composition->n_layers, nZ = T->nZ;
	xmb_main_options opt;
	memset(&opt, 0, sizeof(opt));
	opt.escape_ratios_mode = 1;                                                    /* :5561-5569 */
	memset(fluo, 0, sizeof(double) * (size_t)nE * 109 * nZ);
	memset(compt, 0, sizeof(double) * (size_t)nE * n_out);
	if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(n_threads)
	for (int i = 0; i < nE; i++) {
		ctx_t c;
		memset(&c, 0, sizeof(c));
		c.in = in; c.d = d; c.T = T; c.opt = &opt; c.cascade = 1; c.n_int = 1; c.nL = nL; c.escape_mode = 1;
		const double E0 = in->excitation->discrete[i].energy;
		double initial_mus[64];
		mu_calc(&c, E0, initial_mus);                                              /* :5608-5610 */
		double photons_interacted = 0.0;
		double *fl = fluo + (size_t)i * 109 * nZ;
		for (long j = 0; j < n_photons; j++) {
			photon_t p;
			memset(&p, 0, sizeof(p));
			p.seed = seed; p.g = (uint64_t)i * (uint64_t)n_photons + (uint64_t)j;
			orc_rng rng;
			orc_rng_init(&rng, seed, p.g, ORC_TAG_HISTORY);
			p.energy = E0;
			memcpy(p.mus, initial_mus, sizeof(double) * nL);
			/* xmi_coords_dir, point source (:957-1136) */
			double x1_max = atan(g->slit_size_x / g->d_source_slit / 2.0), y1_max = atan(g->slit_size_y / g->d_source_slit / 2.0);
			double x1 = x1_max * (-1.0 + 2.0 * orc_rng_uniform(&rng));
			double y1 = y1_max * (-1.0 + 2.0 * orc_rng_uniform(&rng));
			p.dirv[0] = tan(x1); p.dirv[1] = tan(y1); p.dirv[2] = 1.0;
			normalize3(p.dirv);
			p.theta = acos(p.dirv[2]);
			p.phi = atan2(p.dirv[1], p.dirv[0]);
			p.weight = 1.0;                                                        /* :5644-5645 */
			p.weight_escape = p.weight;
			double theta_elecv = orc_rng_uniform(&rng) * M_PI * 2.0;               /* :5646-5649 */
			p.elecv[0] = cos(theta_elecv); p.elecv[1] = sin(theta_elecv); p.elecv[2] = 0.0;
			double cosalfa = dot3(p.elecv, p.dirv);
			double c_ae = 1.0 / sin(acos(cosalfa)), c_be = -c_ae * cosalfa;
			for (int k = 0; k < 3; k++) p.elecv[k] = c_ae * p.elecv[k] + c_be * p.dirv[k];
			{   /* xmi_photon_shift_first_layer (:1140-1186) */
				double pp[3] = {0.0, 0.0, d->Z_coord_begin[0]}, inter[3];
				if (!plane_line(pp, g->n_sample_orientation, p.coords, p.dirv, inter)) continue;
				memcpy(p.coords, inter, sizeof(inter));
				p.current_layer = 0;
			}
			/* first pass of xmi_simulate_photon: forced interaction (:1417-1518), no forced detection */
			double b0[4], b1[4];
			draw_block(seed, p.g, 1, 1, 0, 0, b0);
			draw_block(seed, p.g, 1, 1, 0, 1, b1);
			{
				double distances[64], lp[3] = {p.coords[0], p.coords[1], p.coords[2]}, Pabs = 0.0;
				int ok = 1;
				for (int k = 0; k < nL; k++) {
					double pp[3] = {0.0, 0.0, d->Z_coord_end[k]}, inter[3];
					if (!plane_line(pp, g->n_sample_orientation, lp, p.dirv, inter)) { ok = 0; break; }
					distances[k] = dist3(lp, inter);
					memcpy(lp, inter, sizeof(inter));
					Pabs += p.mus[k] * layers[k].density * distances[k];
				}
				if (!ok) continue;
				double Pabs2 = -1.0 * expm1(-1.0 * Pabs);
				p.weight *= Pabs2;
				p.weight_escape = p.weight;                                        /* :1462-1464 */
				double l1p = log1p(-1.0 * b0[0] * Pabs2), negln = -1.0 * l1p;
				int my_index = 0;
				double my_sum = 0.0;
				for (int k = 0; k < nL; k++) {
					my_sum += p.mus[k] * layers[k].density * distances[k];
					if (my_sum > negln) { my_index = k; break; }
				}
				double temp_sum = 0.0;
				for (int k = 0; k <= my_index; k++)
					temp_sum += (1.0 - (p.mus[k] * layers[k].density / (p.mus[my_index] * layers[my_index].density))) * distances[k];
				temp_sum = temp_sum - 1.0 * l1p / (p.mus[my_index] * layers[my_index].density);
				for (int k = 0; k < 3; k++) p.coords[k] += temp_sum * p.dirv[k];
				p.current_layer = my_index;
				p.n_interactions = 1;
			}
			photons_interacted += p.weight;                                        /* :5685-5688 */
			{   /* atom and interaction type (:1558-1652) */
				const xmb_layer *l = &layers[p.current_layer];
				nodepos_t np = node_find(T, p.energy);
				double thr = 0.0;
				for (int k = 0; k < l->n_elements; k++) {
					thr += l->weight[k] * cs_total(&c, l->Z[k], np) / p.mus[p.current_layer];
					if (b0[3] < thr || k == l->n_elements - 1) { p.current_element = l->Z[k]; p.current_element_index = k; break; }
				}
				int zi = T->uniqZ[p.current_element];
				double pr = lerp_at(T->p_rayl + (size_t)zi * T->n_nodes, np), prc = lerp_at(T->p_rayl_compt + (size_t)zi * T->n_nodes, np);
				if (b1[0] < pr) { p.last_interaction = RAYLEIGH; do_rayleigh(&c, &p, b1 + 1); }
				else if (b1[0] < prc) { p.last_interaction = COMPTON; do_compton(&c, &p, b1 + 1); }
				else { p.last_interaction = PHOTO; p.hist_line[1] = 0; do_photo(&c, &p, b1 + 1); }
			}
			/* second pass: energy cut (:1229-1231), then the analogue step of escape mode (:1274-1413) */
			if (p.energy < ENERGY_THRESHOLD) continue;
			int inside = 0;
			{
				int step_max, step_dir;
				if (dot3(p.dirv, g->n_sample_orientation) > 0.0) { step_max = nL - 1; step_dir = 1; } else { step_max = 0; step_dir = -1; }
				double u[4];
				draw_block(seed, p.g, 2, 1, 0, 0, u);
				double interactionR = u[0], blbs = 1.0, max_random_layer = 0.0;
				double cur[3] = {p.coords[0], p.coords[1], p.coords[2]};
				for (int k = p.current_layer; step_dir > 0 ? k <= step_max : k >= step_max; k += step_dir) {
					double pp[3] = {0.0, 0.0, step_dir == 1 ? d->Z_coord_end[k] : d->Z_coord_begin[k]}, inter[3];
					if (!plane_line(pp, g->n_sample_orientation, cur, p.dirv, inter)) { inside = 1; break; }
					double dist = dist3(cur, inter);
					double temp_prod = -1.0 * dist * layers[k].density * p.mus[k];
					double tempexp = exp(temp_prod);
					max_random_layer = max_random_layer - blbs * expm1(temp_prod);
					if (interactionR <= max_random_layer) { inside = 1; break; }
					memcpy(cur, inter, sizeof(inter));
					blbs = blbs * tempexp;
				}
			}
			if (inside) continue;                                                  /* :5692-5693 */
			if (p.last_interaction == COMPTON) {                                   /* :5701-5713 */
				int ci = (int)((p.energy - out_min) / out_delta) + 1;
				if (ci >= 1 && ci <= n_out) {
#pragma omp atomic
					compt[(size_t)(ci - 1) * nE + i] += p.weight;
				}
			} else if (p.last_interaction == PHOTO) {                              /* :5714-5732 */
				int line = -p.hist_line[1];
				if (line >= 1 && line <= 109) fl[(size_t)(line - 1) * nZ + T->uniqZ[p.current_element]] += p.weight_escape;
			}
		}
		for (size_t k = 0; k < (size_t)109 * nZ; k++) fl[k] /= photons_interacted;  /* :5757-5760 */
		for (int k = 0; k < n_out; k++) compt[(size_t)k * nE + i] /= photons_interacted;
	}
	return 1;
}

/* =====================================================================================================
 * Brute-force mode (options%use_variance_reduction = 0): analogue random walk, photons are scored only
 * when they actually reach the detector; Auger / radiative cascades spawn one offspring photon.
 * src/xmi_main.F90:1229-1416 (analogue step), :1920-1984, :2231-2411, :2413-4783; src/xmi_aux_f.F90:1622-1833.
 *
 * Random-number addresses (counter word 2 = (gen<<31)|(order<<20)|(stage<<16)|(elem<<8)|block, gen = 1 for an
 * offspring photon's own walk):
 *   order k, stage 1, blk 0 : {free path, -, -, atom}          blk 1 : {interaction type, s0, s1, s2}
 *   order k, stage 3        : Compton: Doppler trials; photo: phi, yield check, Coster-Kronig hops
 *   order k, stage 4, elem 0: Auger cascade: transition, then the parent's re-emission (yield, CK hops, line,
 *                             theta, phi, polarisation angle);  elem 1: the same for the offspring
 *   order k, stage 5        : radiative cascade: yield, CK hops, line, theta, phi, polarisation angle
 * ===================================================================================================== */
#define GEN_BIT 0x800   /* OR-ed into the order field: bit 31 of the counter word */

enum { DET_NONE = 0, DET_HIT = 1, DET_COLLIMATOR = 2, DET_BAD = 3 };

/* xmi_check_detector_intersection (src/xmi_aux_f.F90:1622-1833) for the segment begin -> end */
static int check_detector_intersection(const ctx_t *c, const double *begin, const double *end) {
	const orc_derived *D = c->d;
	const xmb_geometry *g = c->in->geometry;
	double b[3], e[3], t[3];
	for (int i = 0; i < 3; i++) t[i] = begin[i] - g->p_detector_window[i];
	matvec(D->ndo_inv, t, b);
	for (int i = 0; i < 3; i++) t[i] = end[i] - g->p_detector_window[i];
	matvec(D->ndo_inv, t, e);
	const double pp[3] = {0.0, 0.0, 0.0}, pn[3] = {1.0, 0.0, 0.0};
	double dirv[3] = {e[0] - b[0], e[1] - b[1], e[2] - b[2]}, inter[3];
	if (!D->collimator_present) {
		if (b[0] * e[0] > 0) return DET_NONE;
		/* (the reference assigns the scalar norm to the direction here, :1662 -- SURVEY App. A quirk; the segment
		 * direction is used instead) */
		if (!plane_line(pp, pn, e, dirv, inter)) return DET_NONE;
		inter[0] = 0.0;
		if (norm3(inter) <= D->detector_radius) return dot3(dirv, pn) >= 0.0 ? DET_BAD : DET_HIT;
		return DET_NONE;
	}
	/* conical collimator: quadric intersection of the line with the cone (:1690-1830) */
	if (dirv[0] == 0.0) return DET_NONE;
	const double t_begin = (b[0] - e[0]) / dirv[0], t_end = 0.0;
	double delta[3] = {e[0] - D->vertex[0], e[1] - D->vertex[1], e[2] - D->vertex[2]};
	const double cos2theta = cos(D->half_apex) * cos(D->half_apex);
	const double M[3] = {1.0 - cos2theta, -cos2theta, -cos2theta};
	double dM[3] = {dirv[0] * M[0], dirv[1] * M[1], dirv[2] * M[2]}, deltaM[3] = {delta[0] * M[0], delta[1] * M[1], delta[2] * M[2]};
	const double c2 = dot3(dM, dirv), c1 = dot3(dM, delta), c0 = dot3(deltaM, delta);
	const double disc = c1 * c1 - c0 * c2;
	if (disc < 0.0) return DET_NONE;
	const double t1 = (-c1 + sqrt(disc)) / c2, t2 = (-c1 - sqrt(disc)) / c2;
	const double X1x = e[0] + t1 * dirv[0], X2x = e[0] + t2 * dirv[0];
	const int v1 = -(X1x - D->vertex[0]) >= 0.0, v2 = -(X2x - D->vertex[0]) >= 0.0;
	const double tmax = fmax(t_begin, t_end), tmin = fmin(t_begin, t_end);
	const int in1 = t1 <= tmax && t1 >= tmin && X1x <= g->collimator_height;
	const int in2 = t2 <= tmax && t2 >= tmin && X2x <= g->collimator_height;
	if (!v1 && !v2) return DET_NONE;
	if (v1 && v2) return (in1 || in2) ? DET_COLLIMATOR : DET_NONE;
	if (v1 ? in1 : in2) return DET_COLLIMATOR;
	if (!plane_line(pp, pn, e, dirv, inter)) return DET_NONE;
	inter[0] = 0.0;
	if (norm3(inter) <= D->detector_radius && dist3(b, inter) <= dist3(b, e)) return dot3(dirv, pn) >= 0.0 ? DET_BAD : DET_HIT;
	return DET_NONE;
}

/* xmi_check_photon_detector_hit (src/xmi_main.F90:1920-1984): a photon that left the sample */
static int check_photon_detector_hit(const ctx_t *c, const photon_t *p) {
	const orc_derived *D = c->d;
	const xmb_geometry *g = c->in->geometry;
	if (dot3(p->dirv, g->n_detector_orientation) >= 0.0) return 0;
	double dd[3], cd[3], t[3], inter[3];
	matvec(D->ndo_inv, p->dirv, dd);
	for (int i = 0; i < 3; i++) t[i] = p->coords[i] - g->p_detector_window[i];
	matvec(D->ndo_inv, t, cd);
	double pp[3] = {0.0, 0.0, 0.0};
	const double pn[3] = {1.0, 0.0, 0.0};
	if (!plane_line(pp, pn, cd, dd, inter)) return 0;
	if (norm3(inter) > D->detector_radius) return 0;
	if (!D->collimator_present) return 1;
	pp[0] = g->collimator_height;
	if (!plane_line(pp, pn, cd, dd, inter)) return 0;
	inter[0] = 0.0;
	if (norm3(inter) > D->collimator_radius) return 0;
	return 1;
}

static int line_from_shell(const xmb_tables_host *T, int zi, int shell, double r) {     /* xmi_fluorescence_line_check (:5352-5437) */
	if (shell < 0 || shell > 8) return 0;
	double sumz = 0.0;
	for (int l = xmb_shell_line_first[shell]; l <= xmb_shell_line_last[shell]; l++) {
		sumz += T->rad_rate[(size_t)zi * 384 + l];
		if (r < sumz) return l;
	}
	return 0;
}

/* isotropic re-emission of a cascade photon (:4455-4481, :4733-4767): direction set directly from (theta, phi) */
static void cascade_emit(ctx_t *c, photon_t *q, int line, substream_t *xs) {
	q->energy = c->T->line_energy[(size_t)c->T->uniqZ[q->current_element] * 384 + line];
	mu_calc(c, q->energy, q->mus);
	q->theta = acos(2.0 * sub_uniform(xs) - 1.0);
	q->phi = 2.0 * M_PI * sub_uniform(xs);
	q->dirv[0] = sin(q->theta) * cos(q->phi); q->dirv[1] = sin(q->theta) * sin(q->phi); q->dirv[2] = cos(q->theta);
	q->hist_line[q->n_interactions] = -line;
	q->hist_Z[q->n_interactions] = q->current_element;
	double r = 2.0 * M_PI * sub_uniform(xs);
	q->elecv[0] = cos(r); q->elecv[1] = sin(r); q->elecv[2] = 0.0;
	q->last_interaction = PHOTO;
	double cosalfa = dot3(q->elecv, q->dirv);
	double c_ae = 1.0 / sin(acos(cosalfa)), c_be = -c_ae * cosalfa;
	for (int i = 0; i < 3; i++) q->elecv[i] = c_ae * q->elecv[i] + c_be * q->dirv[i];
}

/* one vacancy of a cascade: yield check (:5325-5350), Coster-Kronig, line; returns the line or 0 */
static int cascade_vacancy(ctx_t *c, int zi, int shell, substream_t *xs, double *energy) {
	const xmb_tables_host *T = c->T;
	if (shell > 8) return 0;                                                       /* N..Q vacancies: no tabulated yield */
	if (shell >= 4 && !c->opt->use_M_lines) return 0;
	if (sub_uniform(xs) > T->fluor_yield_corr[zi * 9 + shell]) return 0;
	shell = coster_kronig(c, zi, shell, xs);
	int line = line_from_shell(T, zi, shell, sub_uniform(xs));
	if (!line) return 0;
	*energy = T->line_energy[(size_t)zi * 384 + line];
	if (*energy <= ENERGY_THRESHOLD) return 0;
	return line;
}

typedef struct { int use_auger, use_rad; } cascade_opt_t;

/* xmi_simulate_photon_cascade_auger (:2413-4594): p has energy 0 on entry (the yield check failed) */
static void cascade_auger(ctx_t *c, photon_t *p, int shell, cascade_opt_t *co, photon_t *off, int *have_off) {
	const xmb_tables_host *T = c->T;
	if (shell < 0 || shell > 3) return;
	const int zi = T->uniqZ[p->current_element];
	const double *a = T->auger_rate + (size_t)zi * XMB_N_AUGER;
	const int first = shell == 0 ? 0 : 240 + 135 * (shell - 1), n = shell == 0 ? 240 : 135;
	substream_t xs;
	sub_init(&xs, p->seed, p->g, p->n_interactions | p->gen, 4, 0);
	const double r = sub_uniform(&xs);
	double sumz = 0.0;
	int found = -1;
	for (int k = 0; k < n; k++) { sumz += a[first + k]; if (r < sumz) { found = k; break; } }
	if (found < 0) return;
	const int new1 = shell == 0 ? 1 + found / 30 : 4 + found / 27, new2 = shell == 0 ? 1 + found % 30 : 4 + found % 27;
	/* the offspring starts as a copy of the parent (:4421-4440) */
	photon_t o = *p;
	double e1 = 0.0, e2 = 0.0;
	int l1 = cascade_vacancy(c, zi, new1, &xs, &e1);
	if (l1) {
		co->use_auger = co->use_rad = 0;                                           /* :4452-4453 */
		cascade_emit(c, p, l1, &xs);
	}
	substream_t ys;
	sub_init(&ys, p->seed, p->g, p->n_interactions | p->gen, 4, 1);
	int l2 = cascade_vacancy(c, zi, new2, &ys, &e2);
	if (l2) {
		cascade_emit(c, &o, l2, &ys);
		*off = o;
		*have_off = 1;
	}
}

/* xmi_simulate_photon_cascade_radiative (:4596-4783): vacancy left behind by the emitted line */
static void cascade_radiative(ctx_t *c, photon_t *p, int shell, int line, cascade_opt_t *co, photon_t *off, int *have_off) {
	const xmb_tables_host *T = c->T;
	int shell_new = -1;
	if (shell == 0) { if (line >= 1 && line <= 8) shell_new = line; }             /* KL1..KM5 -> L1..M5 (lines 1..8) */
	else if (shell >= 1 && shell <= 3 && c->opt->use_M_lines) {
		const int base = shell == 1 ? XMB_L1M1 : shell == 2 ? XMB_L2M1 : XMB_L3M1;
		if (line >= base && line <= base + 4) shell_new = 4 + (line - base);
	} else return;
	if (shell_new < 0) return;
	if (shell_new >= 4 && !c->opt->use_M_lines) return;
	const int zi = T->uniqZ[p->current_element];
	substream_t xs;
	sub_init(&xs, p->seed, p->g, p->n_interactions | p->gen, 5, 0);
	double e = 0.0;
	int l = cascade_vacancy(c, zi, shell_new, &xs, &e);
	if (!l) return;
	photon_t o = *p;
	co->use_auger = co->use_rad = 0;                                               /* :4727-4728 */
	cascade_emit(c, &o, l, &xs);
	*off = o;
	*have_off = 1;
}

static void do_photo_brute(ctx_t *c, photon_t *p, const double *sd, cascade_opt_t *co, photon_t *off, int *have_off) {
	const xmb_tables_host *T = c->T;
	int Z = p->current_element, zi = T->uniqZ[Z];
	nodepos_t np = node_find(T, p->energy);
	double photo_total = lerp_at(T->cs_photo_total + (size_t)zi * T->n_nodes, np);
	double sumz = 0.0;
	int max_shell = c->opt->use_M_lines ? 8 : 3, shell, shell_found = 0;
	for (shell = 0; shell <= max_shell; shell++) {
		sumz += lerp_at(T->cs_photo_partial + ((size_t)zi * 9 + shell) * T->n_nodes, np) / photo_total;
		if (sd[0] < sumz) { shell_found = 1; break; }
	}
	if (!shell_found) { p->energy = 0.0; return; }
	substream_t xs;
	sub_init(&xs, p->seed, p->g, p->n_interactions | p->gen, 3, 0);
	double u_phi = sub_uniform(&xs);
	if (sub_uniform(&xs) > T->fluor_yield_corr[zi * 9 + shell]) {                  /* :2297-2319, :5335-5339 */
		p->energy = 0.0;
		if (co->use_auger) cascade_auger(c, p, shell, co, off, have_off);
		return;
	}
	shell = coster_kronig(c, zi, shell, &xs);
	int line = line_from_shell(T, zi, shell, sd[1]);
	if (!line) { p->energy = 0.0; return; }
	p->energy = T->line_energy[(size_t)zi * 384 + line];
	{
		nodepos_t lp = node_find(T, p->energy);
		for (int i = 0; i < c->nL; i++) p->mus[i] = lerp_at(T->mu_layer + (size_t)i * T->n_nodes, lp);
	}
	update_dirv(p, acos(-2.0 * sd[2] + 1.0), 2.0 * M_PI * u_phi);
	update_elecv(p);
	p->hist_line[p->n_interactions] = -line;
	p->hist_Z[p->n_interactions] = Z;
	if (co->use_rad) cascade_radiative(c, p, shell, line, co, off, have_off);
}

/* xmi_simulate_photon, analogue branch.  Returns 1 if the photon reached the detector. */
static int simulate_photon_brute(ctx_t *c, photon_t *p, cascade_opt_t *co, photon_t *off, int *have_off) {
	const xmb_geometry *g = c->in->geometry;
	const orc_derived *D = c->d;
	const xmb_layer *layers = c->in->composition->layers;
	for (;;) {
		if (p->energy < ENERGY_THRESHOLD) return 0;                                /* :1229 */
		int step_max, step_dir;
		if (dot3(p->dirv, g->n_sample_orientation) > 0.0) { step_max = c->nL - 1; step_dir = 1; } else { step_max = 0; step_dir = -1; }
		double b0[4], b1[4];
		const int order = (p->n_interactions + 1) | p->gen;
		draw_block(p->seed, p->g, order, 1, 0, 0, b0);
		draw_block(p->seed, p->g, order, 1, 0, 1, b1);
		const double interactionR = b0[0];
		double blbs = 1.0, max_random_layer = 0.0, min_random_layer;
		int inside = 0;
		for (int i = p->current_layer; step_dir > 0 ? i <= step_max : i >= step_max; i += step_dir) {   /* :1278-1413 */
			double pp[3] = {0.0, 0.0, step_dir == 1 ? D->Z_coord_end[i] : D->Z_coord_begin[i]}, inter[3];
			if (!plane_line(pp, g->n_sample_orientation, p->coords, p->dirv, inter)) return 0;
			double dist = dist3(p->coords, inter);
			double temp_prod = -1.0 * dist * layers[i].density * p->mus[i];
			double tempexp = exp(temp_prod);
			min_random_layer = max_random_layer;
			max_random_layer = max_random_layer - blbs * expm1(temp_prod);
			if (interactionR <= max_random_layer) {
				dist = -1.0 * log1p(-1.0 * (interactionR - min_random_layer) / blbs) / p->mus[i] / layers[i].density;
				double old[3] = {p->coords[0], p->coords[1], p->coords[2]};
				for (int k = 0; k < 3; k++) p->coords[k] += dist * p->dirv[k];       /* xmi_move_photon_with_dist */
				int rv = check_detector_intersection(c, old, p->coords);
				if (rv == DET_COLLIMATOR || rv == DET_BAD) return 0;
				if (rv == DET_HIT) return 1;
				p->current_layer = i;
				inside = 1;
				break;
			}
			int rv = check_detector_intersection(c, p->coords, inter);
			if (rv == DET_COLLIMATOR || rv == DET_BAD) return 0;
			if (rv == DET_HIT) return 1;
			memcpy(p->coords, inter, sizeof(inter));
			blbs = blbs * tempexp;
		}
		if (!inside) return check_photon_detector_hit(c, p);                       /* :1525-1533 */
		if (p->n_interactions == c->n_int) return 0;                               /* :1536-1539 */
		p->n_interactions++;
		c->n_interactions_total++;
		{
			const xmb_layer *l = &layers[p->current_layer];
			nodepos_t np = node_find(c->T, p->energy);
			double thr = 0.0;
			for (int i = 0; i < l->n_elements; i++) {
				thr += l->weight[i] * cs_total(c, l->Z[i], np) / p->mus[p->current_layer];
				if (b0[3] < thr || i == l->n_elements - 1) { p->current_element = l->Z[i]; p->current_element_index = i; break; }
			}
			int zi = c->T->uniqZ[p->current_element];
			double pr = lerp_at(c->T->p_rayl + (size_t)zi * c->T->n_nodes, np);
			double prc = lerp_at(c->T->p_rayl_compt + (size_t)zi * c->T->n_nodes, np);
			if (b1[0] < pr) { p->last_interaction = RAYLEIGH; do_rayleigh(c, p, b1 + 1); }
			else if (b1[0] < prc) { p->last_interaction = COMPTON; do_compton(c, p, b1 + 1); }
			else { p->last_interaction = PHOTO; do_photo_brute(c, p, b1 + 1, co, off, have_off); }
		}
	}
}

/* Brute-force counterpart of orc_main_msim_range.  channels[(n_int+1)][nch] cumulative from the row of the
 * photon's interaction count upwards (:470-485); brute[n_int][385][100] = Fortran brute_history(Z, slot, k) (:497-523).
 * RAW sums (no live_time).  counters: [0] detector hits, [1] interactions, [2] offspring photons simulated. */
uint64_t orc_main_msim_brute_range(const xmb_input *in, const orc_derived *d, const xmb_tables_host *T, const xmb_main_options *opt,
                                   uint64_t seed, uint64_t g_begin, uint64_t g_end, int n_threads, double *channels, double *brute,
                                   uint64_t *counters) {
	segment_t *segs;
	int nseg = build_segments(in, &segs);
	int n_int = in->general->n_interactions_trajectory, nch = in->detector->nchannels, nL = in->composition->n_layers;
	size_t nchn = (size_t)(n_int + 1) * nch, nbr = (size_t)n_int * 385 * 100;
	memset(channels, 0, sizeof(double) * nchn);
	memset(brute, 0, sizeof(double) * nbr);
	if (n_threads < 1) n_threads = 1;
	uint64_t c0 = 0, c1 = 0, c2 = 0;
#pragma omp parallel num_threads(n_threads)
	{
		ctx_t c;
		memset(&c, 0, sizeof(c));
		c.in = in; c.d = d; c.T = T; c.opt = opt; c.cascade = 1; c.n_int = n_int; c.nch = nch; c.nL = nL;
		double *ch = (double *)calloc(nchn, sizeof(double)), *br = (double *)calloc(nbr, sizeof(double));
		uint64_t hits = 0, n_off = 0;
#pragma omp for schedule(dynamic, 256)
		for (uint64_t gidx = g_begin; gidx < g_end; gidx++) {
			int s = 0;
			while (s + 1 < nseg && gidx >= segs[s + 1].first) s++;
			photon_t p, off;
			orc_rng rng;
			orc_rng_init(&rng, seed, gidx, ORC_TAG_HISTORY);
			int skip, have_off = 0;
			start_photon(&c, &p, &rng, &segs[s], gidx - segs[s].first, &skip);
			p.seed = seed; p.g = gidx;
			if (skip) continue;
			cascade_opt_t co = {opt->use_cascade_auger, opt->use_cascade_radiative};
			photon_t *cur = &p;
			for (int gen = 0; gen < 2; gen++) {
				int hit;
				if (gen == 0) hit = simulate_photon_brute(&c, cur, &co, &off, &have_off);
				else {
					cascade_opt_t none = {0, 0};
					int dummy = 0;
					photon_t unused;
					off.gen = GEN_BIT;
					n_off++;
					cur = &off;
					hit = simulate_photon_brute(&c, cur, &none, &unused, &dummy);
				}
				if (hit) {                                                         /* :443-523 */
					hits++;
					int channel = cur->energy >= ENERGY_THRESHOLD ? (int)((cur->energy - in->detector->zero) / in->detector->gain) : -1;
					if (channel >= 0 && channel < nch)
						for (int k = cur->n_interactions; k <= n_int; k++) ch[(size_t)k * nch + channel] += cur->weight;
					if (cur->n_interactions > 0) {
						int k = cur->n_interactions, hl = cur->hist_line[k], Zel = cur->hist_Z[k];
						int slot = hl < 0 ? -hl : hl == RAYLEIGH ? 384 : hl == COMPTON ? 385 : 0;
						if (slot) br[((size_t)(k - 1) * 385 + (slot - 1)) * 100 + (Zel - 1)] += cur->weight;
					}
				}
				if (!have_off) break;
			}
		}
#pragma omp critical
		{
			for (size_t i = 0; i < nchn; i++) channels[i] += ch[i];
			for (size_t i = 0; i < nbr; i++) brute[i] += br[i];
			c0 += hits; c1 += c.n_interactions_total; c2 += n_off;
		}
		free(ch); free(br);
	}
	if (counters) { counters[0] = c0; counters[1] = c1; counters[2] = c2; }
	free(segs);
	return g_end - g_begin;
}
