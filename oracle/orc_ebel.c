/* orc_ebel.c -- CPU restatement of xmi_tube_ebel (src/xmi_ebel.F90:114-521) and the natural cubic spline
 * it uses (src/xmi_spline.c:38-140).  TEST INFRASTRUCTURE ONLY (see oracle.h). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "xmb_lines.h"

static const double orc_omegaL[100] = {                                                    /* :27-38 */
 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
 2.17E-4, 3.04E-4, 4.15E-4, 5.53E-4, 7.24E-4, 9.30E-4, 0.00118, 0.00147, 0.00181, 0.00221,
 0.00268, 0.00321, 0.00381, 0.00450, 0.00527, 0.00614, 0.00711, 0.00819, 0.00939, 0.0107,
 0.122, 0.0138, 0.0155, 0.0174, 0.0195, 0.0218, 0.0242, 0.0263, 0.0285, 0.0309,
 0.0335, 0.0363, 0.0393, 0.0425, 0.0459, 0.0495, 0.0534, 0.0575, 0.0618, 0.0655,
 0.0714, 0.0765, 0.0820, 0.0877, 0.0938, 0.100, 0.107, 0.114, 0.121, 0.129,
 0.137, 0.145, 0.153, 0.163, 0.172, 0.182, 0.192, 0.202, 0.212, 0.223,
 0.234, 0.245, 0.257, 0.269, 0.281, 0.293, 0.305, 0.318, 0.331, 0.343,
 0.356, 0.369, 0.382, 0.395, 0.409, 0.422, 0.435, 0.448, 0.461, 0.474,
 0.486, 0.499, 0.511, 0.524, 0.536, 0.548, 0.560, 0.572, 0.583, 0.595};

/* natural cubic spline evaluated at v (src/xmi_spline.c) */
double orc_cubic_spline(const double *x, const double *y, size_t np, double v) {
	size_t n = np - 1;
	double *h = calloc(np, sizeof(double)), *al = calloc(np, sizeof(double)), *l = calloc(np, sizeof(double)),
	       *mu = calloc(np, sizeof(double)), *z = calloc(np, sizeof(double)), *c = calloc(np, sizeof(double));
	for (size_t i = 0; i < n; i++) h[i] = x[i + 1] - x[i];
	for (size_t i = 1; i < n; i++) al[i] = 3 * (y[i + 1] - y[i]) / h[i] - 3 * (y[i] - y[i - 1]) / h[i - 1];
	l[0] = 1.0;
	for (size_t i = 1; i < n; i++) {
		l[i] = 2 * (x[i + 1] - x[i - 1]) - h[i - 1] * mu[i - 1];
		mu[i] = h[i] / l[i];
		z[i] = (al[i] - h[i - 1] * z[i - 1]) / l[i];
	}
	c[n] = 0.0;
	for (size_t jj = n; jj-- > 0;) c[jj] = z[jj] - mu[jj] * c[jj + 1];
	long j;
	for (j = 0; j < (long)n; j++) if (x[j] > v) { if (j == 0) j++; break; }
	j--;
	double b = (y[j + 1] - y[j]) / h[j] - h[j] * (c[j + 1] + 2 * c[j]) / 3, d = (c[j + 1] - c[j]) / 3 / h[j], dx = v - x[j];
	double rv = y[j] + b * dx + c[j] * dx * dx + d * dx * dx * dx;
	free(h); free(al); free(l); free(mu); free(z); free(c);
	return rv;
}

/* Returns ncont; *ndisc_out lines.  cont_E/cont_I sized >= (V-1)/dE + 3, disc_E/disc_I sized >= 113.  Intensities are
 * the horizontal (= vertical) intensities. */
int orc_tube_ebel(const xmb_xrl_provider *xrl, const xmb_layer *anode, const xmb_layer *window, const xmb_layer *filter,
                  double V, double current, double angle_e, double angle_x, double dE, double solid_angle, int transmission,
                  size_t n_eff, const double *eff_E, const double *eff, double *cont_E, double *cont_I, int *ndisc_out,
                  double *disc_E, double *disc_I) {
	const double DEG2RAD = 0.01745329, const1 = 1.35E+09, const2_K = 5.0E+13, zk = 2.0, zl = 8.0, bk = 0.35, bl = 0.25;
	const int Z = anode->Z[0];
	double sinalphae = sin(DEG2RAD * angle_e), sinalphax = sin(DEG2RAD * angle_x), sinfactor = sinalphae / sinalphax;
	int ncont = (int)floor((V - 1.0) / dE) + 1;
	if (V / dE != nearbyint(V / dE)) ncont++;                                              /* :205-209 */
	for (int i = 0; i < ncont - 1; i++) cont_E[i] = 1.0 + i * dE;
	cont_E[ncont - 1] = V;
	double x = 1.109 - 0.00435 * Z + 0.00175 * V, m = 0.1382 - 0.9211 / sqrt((double)Z), logz = log((double)Z);
	double eta = (0.1904 - 0.2236 * logz + 0.1292 * (logz * logz) - 0.0149 * (logz * logz * logz)) * pow(V, m);
	double p3 = 0.787E-05 * sqrt(Z * 0.0135) * pow(V, 1.5) + 0.735E-06 * (V * V);
	double rhozmax = xrl->AtomicWeight(Z) * p3 / Z;
	for (int i = 0; i < ncont; i++) {                                                      /* :233-259 */
		double u0 = V / (cont_E[i] + dE / 2.0), logu0 = log(u0);
		double p1 = logu0 * (0.49269 - 1.09870 * eta + 0.78557 * eta * eta), p2 = 0.70256 - 1.09865 * eta + 1.00460 * eta * eta + logu0;
		double rhoz = rhozmax * (p1 / p2), tau = xrl->CS_Total_Kissel(Z, cont_E[i] + dE / 2.0), rhelp = tau * 2.0 * rhoz * sinfactor;
		cont_I[i] = 0.0;
		if (rhelp > 0.0) {
			if (!transmission) cont_I[i] = const1 * Z * pow(u0 - 1.0, x) * (1.0 - exp(-1.0 * rhelp)) / rhelp;
			else cont_I[i] = const1 * Z * pow(u0 - 1.0, x) * (exp(-tau * (anode->density * anode->thickness - 2.0 * rhoz) / sinalphax) -
			                                                   exp(-tau * anode->density * anode->thickness / sinalphax)) / rhelp;
		}
	}
	int nd = 0;
	for (int l = 1; l <= XMB_L3Q1; l++) {                                                  /* :262-391 */
		if (!(xrl->RadRate(Z, -l) > 0.0 && xrl->LineEnergy(Z, -l) > 1.0)) continue;
		int shell = l <= 29 ? 0 : l <= 58 ? 1 : l <= 85 ? 2 : 3;
		double edge = xrl->EdgeEnergy(Z, shell);
		if (edge == 0.0 || edge > V) continue;
		double E = xrl->LineEnergy(Z, -l), u0 = V / edge, logu0 = log(u0);
		double oneovers = (sqrt(u0) * logu0 + 2.0 * (1.0 - sqrt(u0)));
		oneovers = oneovers / (u0 * logu0 + 1.0 - u0);
		oneovers = 1.0 + (16.05 * sqrt(0.0135 * Z / edge) * oneovers);
		if (shell == 0) oneovers = (zk * bk / Z) * (u0 * logu0 + 1.0 - u0) * oneovers;
		else oneovers = (zl * bl / Z) * (u0 * logu0 + 1.0 - u0) * oneovers;
		double r = 1.0 - (0.0081517 * Z) + (3.613e-05 * Z * Z) + (0.009583 * Z * exp(-1.0 * u0)) + (V * 0.001141);
		double p1 = logu0 * (0.49269 - 1.09870 * eta + 0.78557 * eta * eta), p2 = 0.70256 - 1.09865 * eta + 1.00460 * eta * eta + logu0;
		double rhoz = rhozmax * (p1 / p2), tau = xrl->CS_Total_Kissel(Z, E), rhelp = tau * 2.0 * rhoz * sinfactor;
		if (rhelp > 0.0) {
			if (!transmission) rhelp = (1.0 - exp(-1.0 * rhelp)) / rhelp;
			else rhelp = (exp(-tau * (anode->density * anode->thickness - 2.0 * rhoz) / sinalphax) - exp(-tau * anode->density * anode->thickness / sinalphax)) / rhelp;
		}
		double fcorr = Z >= 80 ? 1.0 : -0.4814 + 0.03781 * Z - 2.413E-4 * (Z * Z), I;
		if (shell == 0) I = rhelp * const2_K * oneovers * r * xrl->RadRate(Z, -l) * xrl->FluorYield(Z, 0);
		else if (shell == 1) I = rhelp * fcorr * 0.71E13 * oneovers * r * xrl->RadRate(Z, -l) * orc_omegaL[Z - 1];
		else if (shell == 2) I = rhelp * fcorr * 2.70E13 * oneovers * r * xrl->RadRate(Z, -l) * orc_omegaL[Z - 1];
		else I = rhelp * 4.94E13 * oneovers * r * xrl->RadRate(Z, -l) * orc_omegaL[Z - 1];
		disc_E[nd] = E; disc_I[nd] = I; nd++;
	}
	if (window) {                                                                          /* :393-405 */
		for (int i = 0; i < nd; i++) disc_I[i] *= exp(-1.0 * window->density * window->thickness * xrl->CS_Total_Kissel(window->Z[0], disc_E[i]));
		for (int i = 0; i < ncont; i++) cont_I[i] *= exp(-1.0 * window->density * window->thickness * xrl->CS_Total_Kissel(window->Z[0], cont_E[i] + dE / 2.0));
	}
	if (filter) {                                                                          /* :407-419 */
		for (int i = 0; i < nd; i++) disc_I[i] *= exp(-1.0 * filter->density * filter->thickness * xrl->CS_Total_Kissel(filter->Z[0], disc_E[i]));
		for (int i = 0; i < ncont; i++) cont_I[i] *= exp(-1.0 * filter->density * filter->thickness * xrl->CS_Total_Kissel(filter->Z[0], cont_E[i]));
	}
	for (int i = 0; i < ncont; i++) cont_I[i] *= solid_angle * current / 2.0;
	for (int i = 0; i < nd; i++) disc_I[i] *= solid_angle * current / 2.0;
	if (n_eff > 0) {
		for (int i = 0; i < ncont; i++) cont_I[i] *= orc_cubic_spline(eff_E, eff, n_eff, cont_E[i]);
		for (int i = 0; i < nd; i++) disc_I[i] *= orc_cubic_spline(eff_E, eff, n_eff, disc_E[i]);
	}
	*ndisc_out = nd;
	return ncont;
}
