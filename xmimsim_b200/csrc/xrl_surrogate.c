/* xrl_surrogate.c -- analytic stand-in for xraylib (third-party, absent offline).
 *
 * NOT physics-grade data.  It implements the subset of the xraylib API the reference calls on the
 * hot path (call sites listed in SURVEY.md 8c) with smooth, self-consistent closed forms so that
 * the engine, the oracle and the benchmarks have representative inputs (edges, jumps, line sets,
 * Coster-Kronig, cascades, form factors, Compton profiles) at every Z in 1..94.  Real data plugs
 * in by filling xmb_xrl_provider with libxrl's functions.
 *
 * Models:
 *   binding energies : log-log interpolation in Z through a table of anchor elements, subshells
 *                      gated by Aufbau occupancy
 *   photo-ionisation : single power law above each edge, partitioned by edge-jump ratios
 *   Rayleigh/Compton : two-component form factor F(q), S(q)=Z(1-(...)^-2); CS by Gauss-Legendre
 *                      quadrature of the DCS so that CS and DCS are mutually consistent
 *   yields           : Wentzel-type Z^4/(Z^4+a); fixed Coster-Kronig; dipole-allowed line sets
 *   cascades         : generic restatement of xraylib's P*_kissel recurrences (vacancy transfer by
 *                      Coster-Kronig, radiative and non-radiative decay of deeper shells)
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include "xmimsim_b200.h"
#include "xmb_lines.h"

#define NSH 31
#define AVOGNUM 0.602252       /* xraylib's Avogadro constant in mol^-1 barn^-1 cm^2 units */
#define RE2 0.07940775         /* classical electron radius squared, barn */
#define MEC2 510.998928        /* keV */
#define KEV2ANGST 12.39841930

static const double aw_tab[95] = {0,
 1.008,4.0026,6.94,9.0122,10.81,12.011,14.007,15.999,18.998,20.180,
 22.990,24.305,26.982,28.085,30.974,32.06,35.45,39.948,39.098,40.078,
 44.956,47.867,50.942,51.996,54.938,55.845,58.933,58.693,63.546,65.38,
 69.723,72.630,74.922,78.971,79.904,83.798,85.468,87.62,88.906,91.224,
 92.906,95.95,98.0,101.07,102.91,106.42,107.87,112.41,114.82,118.71,
 121.76,127.60,126.90,131.29,132.91,137.33,138.91,140.12,140.91,144.24,
 145.0,150.36,151.96,157.25,158.93,162.50,164.93,167.26,168.93,173.05,
 174.97,178.49,180.95,183.84,186.21,190.23,192.22,195.08,196.97,200.59,
 204.38,207.2,208.98,209.0,210.0,222.0,223.0,226.0,227.0,232.04,
 231.04,238.03,237.0,244.0};

/* first Z at which each subshell is occupied (Aufbau, j-split) */
static const int z_first[NSH] = {
 1,                       /* K  */
 3, 5, 5,                 /* L1 L2 L3 */
 11, 13, 13, 21, 21,      /* M1..M5 */
 19, 31, 31, 39, 39, 58, 58,  /* N1..N7 */
 37, 49, 49, 71, 71, 91, 91,  /* O1..O7 */
 55, 81, 81, 89, 89,      /* P1..P5 */
 87, 200, 200 };          /* Q1..Q3 */

/* anchor binding energies (keV) for K..M5; 0 = not used as anchor */
#define NANCH 12
static const int anch_Z[NANCH] = {8, 14, 20, 26, 29, 42, 47, 56, 74, 79, 82, 92};
static const double anch_E[NANCH][9] = {
 /* O  */ {0.5320, 0.0237, 0.0071, 0.0071, 0, 0, 0, 0, 0},
 /* Si */ {1.8389, 0.1487, 0.0995, 0.0989, 0.0114, 0.0051, 0.0051, 0, 0},
 /* Ca */ {4.0381, 0.4378, 0.3500, 0.3464, 0.0437, 0.0254, 0.0254, 0, 0},
 /* Fe */ {7.1120, 0.8461, 0.7211, 0.7081, 0.0929, 0.0540, 0.0540, 0.0036, 0.0036},
 /* Cu */ {8.9789, 1.0961, 0.9510, 0.9311, 0.1198, 0.0736, 0.0736, 0.0016, 0.0016},
 /* Mo */ {19.9995, 2.8655, 2.6251, 2.5202, 0.5046, 0.4097, 0.3923, 0.2303, 0.2270},
 /* Ag */ {25.5140, 3.8058, 3.5237, 3.3511, 0.7175, 0.6024, 0.5714, 0.3728, 0.3667},
 /* Ba */ {37.4406, 5.9888, 5.6236, 5.2470, 1.2928, 1.1367, 1.0622, 0.7961, 0.7807},
 /* W  */ {69.5250, 12.0998, 11.5440, 10.2068, 2.8196, 2.5749, 2.2810, 1.8716, 1.8092},
 /* Au */ {80.7249, 14.3528, 13.7336, 11.9187, 3.4249, 3.1478, 2.7430, 2.2911, 2.2057},
 /* Pb */ {88.0045, 15.8608, 15.2000, 13.0352, 3.8507, 3.5542, 3.0664, 2.5856, 2.4840},
 /* U  */ {115.6061, 21.7574, 20.9476, 17.1663, 5.5480, 5.1822, 4.3034, 3.7276, 3.5517}};

static double edge_cache[95][NSH];
static pthread_once_t edge_once = PTHREAD_ONCE_INIT;   /* the table generator calls the provider from OpenMP threads */

static double interp_anchor(int Z, int s) {
	/* log-log interpolation/extrapolation through anchors that have shell s */
	int idx[NANCH], n = 0, i;
	for (i = 0; i < NANCH; i++) if (anch_E[i][s] > 0.0) idx[n++] = i;
	if (n == 0) return 0.0;
	if (n == 1) return anch_E[idx[0]][s];
	int a = 0;
	while (a < n - 2 && Z > anch_Z[idx[a + 1]]) a++;
	double x0 = log((double)anch_Z[idx[a]]), x1 = log((double)anch_Z[idx[a + 1]]);
	double y0 = log(anch_E[idx[a]][s]), y1 = log(anch_E[idx[a + 1]][s]);
	double x = log((double)Z);
	return exp(y0 + (y1 - y0) * (x - x0) / (x1 - x0));
}

static void build_edges(void) {
	static const double nfrac[7] = {1.0, 0.855, 0.721, 0.487, 0.462, 0.160, 0.154};
	static const double ofrac[7] = {1.0, 0.728, 0.571, 0.141, 0.123, 0.030, 0.028};
	static const double pfrac[5] = {1.0, 0.60, 0.45, 0.20, 0.18};
	int Z, s;
	for (Z = 1; Z <= 94; Z++) {
		double *e = edge_cache[Z];
		memset(e, 0, sizeof(double) * NSH);
		for (s = 0; s < 9; s++) if (Z >= z_first[s]) e[s] = interp_anchor(Z, s);
		/* keep strict ordering K > L1 > L2 >= L3 > M1 ... even in extrapolated regions */
		for (s = 1; s < 9; s++) if (e[s] > 0.0 && e[s - 1] > 0.0 && e[s] > 0.98 * e[s - 1]) e[s] = 0.98 * e[s - 1];
		double m1 = e[4] > 0.0 ? e[4] : (e[1] > 0 ? 0.12 * e[1] : 0.0);
		double n1 = m1 * 0.232 * (Z / 82.0);
		for (s = 0; s < 7; s++) if (Z >= z_first[9 + s]) e[9 + s] = n1 * nfrac[s];
		double o1 = n1 * 0.165 * (Z / 82.0);
		for (s = 0; s < 7; s++) if (Z >= z_first[16 + s]) e[16 + s] = o1 * ofrac[s];
		double p1 = o1 * 0.05;
		for (s = 0; s < 5; s++) if (Z >= z_first[23 + s]) e[23 + s] = p1 * pfrac[s];
		if (Z >= z_first[28]) e[28] = p1 * 0.1;
	}
}

static double s_AtomicWeight(int Z) { return (Z >= 1 && Z <= 94) ? aw_tab[Z] : 0.0; }

static double s_EdgeEnergy(int Z, int shell) {
	pthread_once(&edge_once, build_edges);
	if (Z < 1 || Z > 94 || shell < 0 || shell >= NSH) return 0.0;
	return edge_cache[Z][shell];
}

/* ---- radiative transitions -------------------------------------------------------------- */
/* relative emission weights upper(K..M5) -> lower shell; dipole-allowed set */
typedef struct { signed char up, lo; float w; } trans_t;
static const trans_t trans_tab[] = {
 {0, 2, 0.290f}, {0, 3, 0.570f}, {0, 5, 0.040f}, {0, 6, 0.078f}, {0, 10, 0.007f}, {0, 11, 0.013f}, {0, 17, 0.001f}, {0, 18, 0.001f},
 {1, 5, 0.330f}, {1, 6, 0.450f}, {1, 10, 0.080f}, {1, 11, 0.110f}, {1, 17, 0.015f}, {1, 18, 0.015f},
 {2, 4, 0.030f}, {2, 7, 0.790f}, {2, 9, 0.010f}, {2, 12, 0.150f}, {2, 19, 0.020f},
 {3, 4, 0.040f}, {3, 7, 0.080f}, {3, 8, 0.700f}, {3, 9, 0.010f}, {3, 12, 0.015f}, {3, 13, 0.140f}, {3, 19, 0.005f}, {3, 20, 0.010f},
 {4, 10, 0.45f}, {4, 11, 0.50f}, {4, 17, 0.02f}, {4, 18, 0.03f},
 {5, 9, 0.25f}, {5, 12, 0.70f}, {5, 19, 0.05f},
 {6, 9, 0.20f}, {6, 12, 0.08f}, {6, 13, 0.67f}, {6, 20, 0.05f},
 {7, 10, 0.08f}, {7, 11, 0.02f}, {7, 14, 0.88f}, {7, 17, 0.02f},
 {8, 11, 0.08f}, {8, 14, 0.04f}, {8, 15, 0.86f}, {8, 18, 0.02f}};
#define NTRANS ((int)(sizeof(trans_tab) / sizeof(trans_tab[0])))

/* dense = 1: besides the dipole-allowed set every transition of xraylib's line enumeration whose lower subshell exists
 * for the element carries a (small) weight, as in xraylib's tables, where the forbidden satellites have rates of 1e-4 .. 1e-2:
 * the number of ACTIVE forced-detection lines then matches real data (srm1155: ~320 instead of ~150), which is what the
 * line phase of the history kernel scales with. */
static double raw_rate_mode(int Z, int up, int lo, int dense) {
	int i;
	if (up < 0 || lo <= up || lo >= NSH || Z < z_first[up] || Z < z_first[lo]) return 0.0;
	for (i = 0; i < NTRANS; i++) if (trans_tab[i].up == up && trans_tab[i].lo == lo) return trans_tab[i].w;
	if (!dense || lo - up > 11) return 0.0;   /* satellites to the next subshells only: ~320 active lines for srm1155, as with xraylib data */
	return 2.0e-3 / (1.0 + 0.25 * (lo - up));
}
static double rad_rate_mode(int Z, int line, int dense) {
	int l = -line, j;
	if (Z < 1 || Z > 94 || l < 1 || l > XMB_N_LINES) return 0.0;
	int up = xmb_line_upper[l], lo = xmb_line_lower[l];
	if (up < 0 || up > 8 || lo < 0) return 0.0;
	double r = raw_rate_mode(Z, up, lo, dense);
	if (r == 0.0) return 0.0;
	double tot = 0.0;
	for (j = up + 1; j < NSH; j++) tot += raw_rate_mode(Z, up, j, dense);
	return tot > 0.0 ? r / tot : 0.0;
}
static double s_RadRate(int Z, int line) { return rad_rate_mode(Z, line, 0); }
static double s_RadRate_dense(int Z, int line) { return rad_rate_mode(Z, line, 1); }

static double s_LineEnergy(int Z, int line) {
	int l = -line;
	if (Z < 1 || Z > 94 || l < 1 || l > XMB_N_LINES) return 0.0;
	int up = xmb_line_upper[l], lo = xmb_line_lower[l];
	if (up < 0 || lo < 0) return 0.0;
	double eu = s_EdgeEnergy(Z, up), el = s_EdgeEnergy(Z, lo);
	if (eu <= 0.0 || el <= 0.0 || eu <= el) return 0.0;
	return eu - el;
}

static double s_FluorYield(int Z, int shell) {
	if (Z < 1 || Z > 94 || shell < 0 || shell > 8 || Z < z_first[shell]) return 0.0;
	double z4 = (double)Z * Z * Z * Z;
	switch (shell) {
	case 0: return z4 / (z4 + 1.12e6);
	case 1: return 0.40 * z4 / (z4 + 1.0e8);
	case 2: return 1.05 * z4 / (z4 + 1.0e8);
	case 3: return z4 / (z4 + 1.0e8);
	default: return (0.6 + 0.1 * (shell - 4)) * z4 / (z4 + 1.3e9);
	}
}

static const double ck_tab[XMB_N_CK] = {0.12, 0.50, 0.12, 0.10, 0.10, 0.10, 0.10, 0.10, 0.10, 0.10, 0.05, 0.05, 0.05};
static const signed char ck_from[XMB_N_CK] = {1, 1, 2, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7};
static const signed char ck_to[XMB_N_CK]   = {2, 3, 3, 5, 6, 7, 8, 6, 7, 8, 7, 8, 8};

static double s_CosKron(int Z, int trans) {
	if (Z < 1 || Z > 94 || trans < 0 || trans >= XMB_N_CK) return 0.0;
	if (Z < z_first[(int)ck_from[trans]] || Z < z_first[(int)ck_to[trans]]) return 0.0;
	return ck_tab[trans];
}

static double s_JumpRatio(int Z, int shell) {
	double z = Z;
	switch (shell) {
	case 0: { double r = 17.54 - 0.6608 * z + 0.01427 * z * z - 1.1e-4 * z * z * z; if (Z > 60) r = 5.5 - 0.02 * (z - 60); return r < 4.5 ? 4.5 : r; }
	case 1: return 1.17;
	case 2: return 1.39;
	case 3: { double r = 20.03 - 0.7732 * z + 0.01159 * z * z - 5.835e-5 * z * z * z; if (Z < 30) r = 5.7; return r < 2.2 ? 2.2 : r; }
	case 4: return 1.10;
	case 5: return 1.10;
	case 6: return 1.20;
	case 7: return 1.50;
	case 8: return 2.00;
	default: return 1.0;
	}
}

static double s_JumpFactor(int Z, int shell) {
	if (Z < 1 || Z > 94 || shell < 0 || shell > 8 || Z < z_first[shell]) return 0.0;
	return s_JumpRatio(Z, shell);
}

/* ---- photo-ionisation ------------------------------------------------------------------- */
#define PHOTO_EXP 2.85
static double photo_envelope(int Z, double E) {
	/* total photo CS the atom would have with every shell open (cm2/g) */
	double z = Z;
	return 22.0 * z * z * z * z / aw_tab[Z] * pow(E, -PHOTO_EXP) * pow(10.0, PHOTO_EXP - 3.0);
}

/* share[s] of the envelope carried by subshell s (K..M5) and by the outer remainder (index 9) */
static void photo_shares(int Z, double share[10]) {
	double rem = 1.0;
	int s;
	for (s = 0; s < 9; s++) {
		if (Z >= z_first[s]) {
			double r = s_JumpRatio(Z, s);
			share[s] = rem * (1.0 - 1.0 / r);
			rem = rem / r;
		} else share[s] = 0.0;
	}
	share[9] = rem;
}

static double s_CS_Photo_Partial(int Z, int shell, double E) {
	if (Z < 1 || Z > 94 || shell < 0 || shell > 8 || E <= 0.0) return 0.0;
	double edge = s_EdgeEnergy(Z, shell);
	if (edge <= 0.0 || E < edge) return 0.0;
	double share[10];
	photo_shares(Z, share);
	return share[shell] * photo_envelope(Z, E);
}

static double s_CS_Photo_Total(int Z, double E) {
	if (Z < 1 || Z > 94 || E <= 0.0) return 0.0;
	double share[10], f = 0.0;
	int s;
	photo_shares(Z, share);
	for (s = 0; s < 9; s++) { double edge = s_EdgeEnergy(Z, s); if (edge > 0.0 && E >= edge) f += share[s]; }
	f += share[9];
	return f * photo_envelope(Z, E);
}

/* ---- scattering ------------------------------------------------------------------------- */
static double s_FF(int Z, double q) {
	if (Z < 1 || Z > 94) return 0.0;
	double z = Z;
	if (Z <= 2) { double x2 = (z - 0.3 * (Z - 1)) / 3.32; double t = 1.0 + (q / x2) * (q / x2); return z / (t * t); }
	double x1 = 0.22 * cbrt(z), x2 = (z - 0.3) / 3.32;
	double t1 = 1.0 + (q / x1) * (q / x1), t2 = 1.0 + (q / x2) * (q / x2);
	return (z - 2.0) / (t1 * t1) + 2.0 / (t2 * t2);
}

static double s_SF(int Z, double q) {
	if (Z < 1 || Z > 94) return 0.0;
	double z = Z;
	double xs = 0.18 * cbrt(z);
	double t = 1.0 + (q / xs) * (q / xs);
	return z * (1.0 - 1.0 / (t * t));
}

static double dcs_thoms(double ct) { return RE2 * 0.5 * (1.0 + ct * ct); }
static double dcs_kn(double E, double ct) {
	double k = 1.0 / (1.0 + E / MEC2 * (1.0 - ct));
	return RE2 * 0.5 * k * k * (k + 1.0 / k - (1.0 - ct * ct));
}

/* 96-point Gauss-Legendre nodes would be long to type: use composite 3-point GL on 64 panels in theta */
static double integrate_dcs(int Z, double E, int compton) {
	static const double gx[3] = {-0.7745966692414834, 0.0, 0.7745966692414834};
	static const double gw[3] = {0.5555555555555556, 0.8888888888888888, 0.5555555555555556};
	const int NP = 96;
	double sum = 0.0;
	int p, g;
	/* panels graded towards theta = 0 where F^2 is sharply peaked */
	for (p = 0; p < NP; p++) {
		double a = M_PI * pow((double)p / NP, 2.0), b = M_PI * pow((double)(p + 1) / NP, 2.0);
		for (g = 0; g < 3; g++) {
			double th = 0.5 * (a + b) + 0.5 * (b - a) * gx[g];
			double ct = cos(th), st = sin(th);
			double q = E / KEV2ANGST * sin(th * 0.5);
			double f;
			if (compton) f = dcs_kn(E, ct) * s_SF(Z, q);
			else { double F = s_FF(Z, q); f = dcs_thoms(ct) * F * F; }
			sum += gw[g] * 0.5 * (b - a) * f * st;
		}
	}
	return 2.0 * M_PI * sum * AVOGNUM / aw_tab[Z];
}

/* CS_Rayl/CS_Compt are smooth in E: cache on a log grid per element and interpolate log-log */
#define NCS 192
static double cs_cache[95][2][NCS];
static unsigned char cs_have[95];
static pthread_mutex_t cs_lock = PTHREAD_MUTEX_INITIALIZER;
static const double cs_e0 = 0.05, cs_e1 = 250.0;

static void build_cs(int Z) {
	int i;
	for (i = 0; i < NCS; i++) {
		double E = cs_e0 * pow(cs_e1 / cs_e0, (double)i / (NCS - 1));
		cs_cache[Z][0][i] = log(integrate_dcs(Z, E, 0));
		cs_cache[Z][1][i] = log(integrate_dcs(Z, E, 1));
	}
	__atomic_store_n(&cs_have[Z], 1, __ATOMIC_RELEASE);
}

static double cs_lookup(int Z, double E, int which) {
	if (Z < 1 || Z > 94 || E <= 0.0) return 0.0;
	if (!__atomic_load_n(&cs_have[Z], __ATOMIC_ACQUIRE)) {
		pthread_mutex_lock(&cs_lock);
		if (!__atomic_load_n(&cs_have[Z], __ATOMIC_ACQUIRE)) build_cs(Z);
		pthread_mutex_unlock(&cs_lock);
	}
	if (E < cs_e0) E = cs_e0;
	if (E > cs_e1) E = cs_e1;
	double x = log(E / cs_e0) / log(cs_e1 / cs_e0) * (NCS - 1);
	int i = (int)x;
	if (i >= NCS - 1) i = NCS - 2;
	double f = x - i;
	return exp(cs_cache[Z][which][i] * (1.0 - f) + cs_cache[Z][which][i + 1] * f);
}

static double s_CS_Rayl(int Z, double E) { return cs_lookup(Z, E, 0); }
static double s_CS_Compt(int Z, double E) { return cs_lookup(Z, E, 1); }
static double s_CS_Total(int Z, double E) { return s_CS_Photo_Total(Z, E) + s_CS_Rayl(Z, E) + s_CS_Compt(Z, E); }

/* ---- Compton profile -------------------------------------------------------------------- */
static double s_ComptonProfile(int Z, double pz) {
	if (Z < 1 || Z > 94) return 0.0;
	double z = Z, J = 0.0;
	double nK = Z >= 2 ? 2.0 : 1.0, nL = Z > 2 ? (Z >= 10 ? 8.0 : z - 2.0) : 0.0, nO = Z > 10 ? z - 10.0 : 0.0;
	double pK = Z >= 2 ? z - 0.3 : 1.0, pL = Z > 2 ? (z - 2.0 > 1.0 ? (z - 2.0) * 0.5 : 0.6) : 1.0, pO = 0.9 + 0.01 * z;
	double c = 8.0 / (3.0 * M_PI);
	double t;
	t = 1.0 + (pz / pK) * (pz / pK); J += nK * c / pK / (t * t * t);
	if (nL > 0.0) { t = 1.0 + (pz / pL) * (pz / pL); J += nL * c / pL / (t * t * t); }
	if (nO > 0.0) { t = 1.0 + (pz / pO) * (pz / pO); J += nO * c / pO / (t * t * t); }
	return J;
}

/* shell-resolved profiles: electrons fill the subshells in the order of z_first; every subshell gets the hydrogenic
 * 1s shape with the momentum scale of its binding energy (normalised per electron) */
static const int sh_cap[NSH] = {2, 2, 2, 4, 2, 2, 4, 4, 6, 2, 2, 4, 4, 6, 6, 8, 2, 2, 4, 4, 6, 6, 8, 2, 2, 4, 4, 6, 2, 2, 4};
static double s_ElectronConfig(int Z, int shell) {
	if (Z < 1 || Z > 94 || shell < 0 || shell >= NSH) return 0.0;
	/* fill in order of first occupation (ties: shell number) */
	int left = Z, s, zf;
	for (zf = 1; zf <= 94 && left > 0; zf++)
		for (s = 0; s < NSH && left > 0; s++) {
			if (z_first[s] != zf) continue;
			int n = left < sh_cap[s] ? left : sh_cap[s];
			if (s == shell) return (double)n;
			left -= n;
		}
	return 0.0;
}
static double s_ComptonProfile_Partial(int Z, int shell, double pz) {
	if (s_ElectronConfig(Z, shell) <= 0.0) return 0.0;
	double I = s_EdgeEnergy(Z, shell);
	if (I <= 0.0) I = 0.005;                       /* outer shells without an anchored edge: ~5 eV */
	double p = sqrt(I / 0.0136057);                /* a.u.: p = sqrt(I / Ry) */
	double t = 1.0 + (pz / p) * (pz / p);
	return 8.0 / (3.0 * M_PI) / p / (t * t * t);
}

/* ---- cascade vacancy cross sections ----------------------------------------------------- */
/* vacancies created in `lo` per non-radiative decay of a vacancy in `up` (surrogate constants) */
static double auger_transfer(int Z, int up, int lo) {
	static const double kl[3] = {0.45, 0.55, 0.85};             /* K -> L1 L2 L3   (KLL, KLX) */
	static const double km[5] = {0.03, 0.03, 0.05, 0.02, 0.02}; /* K -> M1..M5 */
	static const double lm[5] = {0.25, 0.25, 0.45, 0.45, 0.55}; /* L  -> M1..M5 (LMM) */
	if (Z < z_first[up] || Z < z_first[lo]) return 0.0;
	if (up == 0 && lo >= 1 && lo <= 3) return kl[lo - 1];
	if (up == 0 && lo >= 4 && lo <= 8) return km[lo - 4];
	if (up >= 1 && up <= 3 && lo >= 4 && lo <= 8) return lm[lo - 4];
	return 0.0;
}

static double auger_yield(int Z, int shell) {
	/* xraylib: AugerYield = 1 - FluorYield - sum of Coster-Kronig from this shell */
	double a = 1.0 - s_FluorYield(Z, shell);
	int t;
	for (t = 0; t < XMB_N_CK; t++) if (ck_from[t] == shell) a -= s_CosKron(Z, t);
	return a < 0.0 ? 0.0 : a;
}

/* AugerRate: ordered pair distribution whose expected vacancy counts reproduce auger_transfer(), so that the explicit
 * cascade of the brute-force mode and the analytic cascade of the forced-detection mode (VacancyCS) agree:
 * p(X,Y) = t(X) t(Y) / (2 S), S = sum t  =>  sum_Y p(lo,Y) + sum_X p(X,lo) = t(lo), total S/2 <= 1. */
static double s_AugerRate(int Z, int shell, int n1, int n2) {
	int lo0, lo1, s;
	double S = 0.0;
	if (shell == 0) { lo0 = 1; lo1 = 8; } else if (shell >= 1 && shell <= 3) { lo0 = 4; lo1 = 8; } else return 0.0;
	if (n1 < lo0 || n1 > lo1 || n2 < lo0 || n2 > lo1) return 0.0;   /* outer partners (N..Q) carry no weight here */
	for (s = lo0; s <= lo1; s++) S += auger_transfer(Z, shell, s);
	if (S <= 0.0) return 0.0;
	return auger_transfer(Z, shell, n1) * auger_transfer(Z, shell, n2) / (2.0 * S);
}

static int line_index(int up, int lo) {
	int l;
	for (l = 1; l <= XMB_M5P5; l++) if (xmb_line_upper[l] == up && xmb_line_lower[l] == lo) return l;
	return 0;
}

static double vacancy_cs_mode(int Z, int shell, double E, int cascade, const double *P, int dense) {
	/* P[u] for u < shell already evaluated under the same cascade mode */
	if (shell < 0 || shell > 8) return 0.0;
	double rv = s_CS_Photo_Partial(Z, shell, E);
	if (shell == 0) return rv;
	int t, u;
	for (t = 0; t < XMB_N_CK; t++)
		if (ck_to[t] == shell && P[(int)ck_from[t]] > 0.0) rv += s_CosKron(Z, t) * P[(int)ck_from[t]];
	int first_same = shell <= 3 ? 1 : 4;   /* deeper principal shells only */
	for (u = 0; u < first_same; u++) {
		if (P[u] <= 0.0) continue;
		if (cascade == 3 || cascade == 4) {
			int l = line_index(u, shell);
			if (l) rv += s_FluorYield(Z, u) * rad_rate_mode(Z, -l, dense) * P[u];
		}
		if (cascade == 2 || cascade == 4) rv += auger_yield(Z, u) * auger_transfer(Z, u, shell) * P[u];
	}
	return rv;
}
static double s_VacancyCS(int Z, int shell, double E, int cascade, const double *P) { return vacancy_cs_mode(Z, shell, E, cascade, P, 0); }
static double s_VacancyCS_dense(int Z, int shell, double E, int cascade, const double *P) { return vacancy_cs_mode(Z, shell, E, cascade, P, 1); }

static const xmb_xrl_provider surrogate = {
	"xrl-surrogate-1 (analytic stand-in, NOT xraylib data)",
	s_AtomicWeight, s_EdgeEnergy, s_LineEnergy, s_FluorYield, s_RadRate, s_CosKron, s_JumpFactor,
	s_CS_Total, s_CS_Photo_Total, s_CS_Photo_Partial, s_CS_Rayl, s_CS_Compt, s_FF, s_SF,
	s_ComptonProfile, s_VacancyCS, s_AugerRate, s_ElectronConfig, s_ComptonProfile_Partial};

static const xmb_xrl_provider surrogate_dense = {
	"xrl-surrogate-1-dense (analytic stand-in with xraylib's line density, NOT xraylib data)",
	s_AtomicWeight, s_EdgeEnergy, s_LineEnergy, s_FluorYield, s_RadRate_dense, s_CosKron, s_JumpFactor,
	s_CS_Total, s_CS_Photo_Total, s_CS_Photo_Partial, s_CS_Rayl, s_CS_Compt, s_FF, s_SF,
	s_ComptonProfile, s_VacancyCS_dense, s_AugerRate, s_ElectronConfig, s_ComptonProfile_Partial};

/* The stand-in with as many active fluorescence lines per element as xraylib has (workload realism for benchmarks). */
const xmb_xrl_provider *xmb_xrl_surrogate_dense(void) {
	pthread_once(&edge_once, build_edges);
	return &surrogate_dense;
}

const xmb_xrl_provider *xmb_xrl_surrogate(void) {
	pthread_once(&edge_once, build_edges);
	return &surrogate;
}
