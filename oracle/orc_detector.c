/* orc_detector.c -- CPU restatement of the detector response.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Follows src/xmi_detector_f.F90: xmi_detector_convolute_spectrum :339-580, xmi_detector_escape :582-809,
 * xmi_detector_sum_peaks :56-217, xmi_detector_poisson :862-905, xmi_detector_convolute_history :291-337.
 * Random draws (pile-up, Poisson) use Philox streams instead of MT19937: only distributional parity is
 * meaningful there. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "orc_rng.h"
#include "xmb_lines.h"

static double mu_layer(const xmb_xrl_provider *xrl, const xmb_layer *l, double E) {   /* src/xmi_aux_f.F90:1125-1141 */
	double rv = 0.0;
	for (int i = 0; i < l->n_elements; i++) rv += xrl->CS_Total_Kissel(l->Z[i], E) * l->weight[i];
	return rv;
}

/* per-energy detector efficiency: absorbers exp(-mu rho t), crystal 1-exp(-mu rho t)  (:437-454, :315-329) */
double orc_detector_correction(const xmb_input *in, const xmb_xrl_provider *xrl, double E) {
	double det_corr = 1.0;
	for (int j = 0; j < in->absorbers->n_det_layers; j++) {
		const xmb_layer *l = &in->absorbers->det_layers[j];
		det_corr = det_corr * exp(-1.0 * l->density * l->thickness * mu_layer(xrl, l, E));
	}
	for (int j = 0; j < in->detector->n_crystal_layers; j++) {
		const xmb_layer *l = &in->detector->crystal_layers[j];
		det_corr = -1.0 * det_corr * expm1(-1.0 * l->density * l->thickness * mu_layer(xrl, l, E));
	}
	return det_corr;
}

/* findpos (src/xmi_aux_f.F90:1305-1335), 0-based return, -1 when not found */
static int findpos(const double *a, int n, double x) {
	if (fabs(x - a[0]) < 1e-10) return 0;
	for (int i = 1; i < n; i++) if (x <= a[i]) return i - 1;
	return -1;
}

/* xmi_detector_escape (:582-809); ch[0..nch-1] modified in place */
static void detector_escape(double *ch, int nch, const xmb_input *in, const xmb_xrl_provider *xrl, const xmb_escape_ratios *er) {
	const double gain = in->detector->gain, zero = in->detector->zero;
	const double channel_1e = 0.5 * gain + zero;
	const double compton_out_diff = er->compton_escape_output_energies[1] - er->compton_escape_output_energies[0];
	const int nfe = er->n_fluo_input_energies, nel = er->n_elements, nci = er->n_compton_input_energies, nco = er->n_compton_output_energies;
	/* line ranges per shell: K KL1..KP5, L1 L1M1..L1P5, L2 L2M1..L2Q1, L3 L3M1..L3P3 (109) */
	const int first[4] = {1, XMB_L1M1, XMB_L2M1, 86}, last[4] = {29, 58, 85, 109};
	for (int i = 0; i < nch; i++) {
		double channel_e = ((double)(float)i + 0.5) * gain + zero;     /* REAL(i) is single precision in the reference */
		double sum_ratio = 0.0;
		for (int j = 0; j < nel; j++) {
			for (int sh = 0; sh < 4; sh++) {
				if (!(channel_e > xrl->EdgeEnergy(er->Z[j], sh))) continue;
				for (int k = first[sh]; k <= last[sh]; k++) {
					double line_e = xrl->LineEnergy(er->Z[j], -k);
					if ((channel_e - line_e) >= channel_1e && channel_e >= er->fluo_escape_input_energies[0] &&
					    channel_e < er->fluo_escape_input_energies[nfe - 1]) {
						int pos = findpos(er->fluo_escape_input_energies, nfe, channel_e);
						if (pos < 0) continue;
						/* fluo_escape_ratios(element, |line|, energy), element fastest */
						double a1 = er->fluo_escape_input_energies[pos], b1 = er->fluo_escape_input_energies[pos + 1];
						double a2 = er->fluo_escape_ratios[((size_t)pos * 109 + (k - 1)) * nel + j];
						double b2 = er->fluo_escape_ratios[((size_t)(pos + 1) * 109 + (k - 1)) * nel + j];
						double ratio = a2 + ((b2 - a2) * (channel_e - a1) / (b1 - a1));
						sum_ratio += ratio;
						int escape_i = (int)((channel_e - line_e - zero) / gain);
						/* K lines: escape_i < UBOUND (:644-645); L lines: <= UBOUND (:683-684) */
						int ok = sh == 0 ? (escape_i >= 0 && escape_i < nch - 1) : (escape_i >= 0 && escape_i <= nch - 1);
						if (ok) ch[escape_i] += ratio * ch[i];
					}
				}
			}
		}
		for (int j = 0; j < i; j++) {                                  /* Compton escape (:777-805) */
			double channel_c = ((double)(float)j + 0.5) * gain + zero;
			double compton_diff = channel_e - channel_c;
			if (channel_e >= er->compton_escape_input_energies[0] && channel_e < er->compton_escape_input_energies[nci - 1] &&
			    compton_diff >= er->compton_escape_output_energies[0] && compton_diff < er->compton_escape_output_energies[nco - 1]) {
				int p1 = findpos(er->compton_escape_input_energies, nci, channel_e);
				int p2 = findpos(er->compton_escape_output_energies, nco, compton_diff);
				if (p1 < 0 || p2 < 0) continue;
				const double *x1 = er->compton_escape_input_energies, *x2 = er->compton_escape_output_energies;
				double denom = (x1[p1 + 1] - x1[p1]) * (x2[p2 + 1] - x2[p2]);
				double c1 = (x1[p1 + 1] - channel_e) * (x2[p2 + 1] - compton_diff) / denom, c2 = (channel_e - x1[p1]) * (x2[p2 + 1] - compton_diff) / denom;
				double c3 = (x1[p1 + 1] - channel_e) * (compton_diff - x2[p2]) / denom, c4 = (channel_e - x1[p1]) * (compton_diff - x2[p2]) / denom;
				/* compton_escape_ratios(input, output), input fastest */
				const double *A = er->compton_escape_ratios;
				double v = c1 * A[(size_t)p2 * nci + p1] + c2 * A[(size_t)p2 * nci + p1 + 1] + c3 * A[(size_t)(p2 + 1) * nci + p1] +
				           c4 * A[(size_t)(p2 + 1) * nci + p1 + 1];
				double ratio = v * gain / compton_out_diff;
				ch[j] += ratio * ch[i];
				sum_ratio += ratio;
			}
		}
		ch[i] = ch[i] * (1.0 - sum_ratio);
	}
}

/* xmi_detector_sum_peaks (:56-217): sequential pulse-train Monte Carlo */
static void detector_sum_peaks(double *ch, int nch, const xmb_input *in, uint64_t seed, int order) {
	double Nt = 0.0;
	for (int i = 0; i < nch; i++) Nt += ch[i];
	long Nt_long = (long)Nt;
	double lambda = Nt / in->detector->live_time, mu = 1.0 / lambda;
	double *cdf = (double *)malloc(sizeof(double) * nch), *nw = (double *)calloc(nch, sizeof(double));
	double run = 0.0;
	for (int i = 0; i < nch; i++) { run += ch[i] > 0 ? ch[i] : 0.0; cdf[i] = run; }
	orc_rng rng;
	orc_rng_init(&rng, seed, (uint64_t)order, ORC_TAG_DETECTOR);
	long pulses[100], npulses = 0, npulses_all = 0;
	if (Nt_long > 0 && run > 0.0)
	for (;;) {
		npulses++; npulses_all++;
		if (npulses > 100) break;                                   /* reference aborts: pulsetrain maximum */
		double u = orc_rng_uniform(&rng) * run;                       /* ran_discrete: channel ~ counts */
		int lo = 0, hi = nch - 1;
		while (lo < hi) { int mid = (lo + hi) / 2; if (cdf[mid] > u) hi = mid; else lo = mid + 1; }
		pulses[npulses - 1] = lo + 1;                                /* +1: channels count from 1 here (:160) */
		double deltaT = -mu * log(1.0 - orc_rng_uniform(&rng));       /* ran_exponential(mu) */
		if (deltaT > in->detector->pulse_width) {
			if (npulses == 1) nw[pulses[0] - 1] += 1.0;
			else {
				double energies_sum = 0.0;
				for (long k = 0; k < npulses; k++) energies_sum += (pulses[k] * in->detector->gain) + in->detector->zero;
				long pulses_sum = (long)((energies_sum - in->detector->zero) / in->detector->gain);
				if (pulses_sum > 0 && pulses_sum <= nch) nw[pulses_sum - 1] += 1.0;
			}
			if (npulses_all >= Nt_long) break;
			npulses = 0;
		}
	}
	memcpy(ch, nw, sizeof(double) * nch);
	free(cdf); free(nw);
}

/* Poisson deviate: inversion for small means, PTRS (Hoermann 1993) otherwise */
static double ran_poisson(orc_rng *r, double lam) {
	if (lam < 10.0) {
		double L = exp(-lam), p = 1.0; long k = 0;
		do { k++; p *= orc_rng_uniform(r); } while (p > L);
		return (double)(k - 1);
	}
	double slam = sqrt(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b, inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
	for (;;) {
		double U = orc_rng_uniform(r) - 0.5, V = orc_rng_uniform(r), us = 0.5 - fabs(U);
		double k = floor((2.0 * a / us + b) * U + lam + 0.43);
		if (us >= 0.07 && V <= vr) return k;
		if (k < 0 || (us < 0.013 && V > us)) continue;
		if (log(V) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * log(lam) - lgamma(k + 1.0)) return k;
	}
}

/* Gaussian + tail/shelf response (:502-558); temp[nch] -> conv[nch] */
void orc_detector_gaussian(const xmb_input *in, const double *temp, double *conv) {
	const xmb_detector *d = in->detector;
	const int nch = d->nchannels, nlim = nch - 1;
	const double a = d->noise * d->noise, b = (2.3548) * (2.3548) * 3.85 * d->fano / 1000.0;
	const double c = sqrt(2.0) / (2.0 * sqrt(2.0 * log(2.0)));
	const double M_SQRTPI = 1.77245385090551602729816748334;
	double *R = (double *)calloc(nlim + 101, sizeof(double));
	memset(conv, 0, sizeof(double) * nch);
	for (int I0 = 0; I0 <= nlim; I0++) {
		double E0 = d->zero + d->gain * I0;
		if (E0 < 1.0) continue;
		double FWHM = sqrt(a + b * E0), B0 = c * FWHM, A0 = 1.0 / (B0 * M_SQRTPI);
		double A3 = 2.73E-3 * exp(-0.21 * E0) + 1.E-4;
		double A4 = 0.000188 * exp(-0.00296 * pow(E0, 0.763)) + 1.355E-5 * exp(0.968 * pow(E0, 0.498));
		double ALFA = 1.179 * exp(8.6E-4 * pow(E0, 1.877)) - 7.793 * exp(-3.81 * pow(E0, -0.0716));
		double my_sum = 0.0;
		for (int I = 0; I <= I0 + 100; I++) {
			if (I >= nch) break;
			double E = d->zero + d->gain * I, X = (E - E0) / B0, G = exp(-X * X), F = erfc(X);
			if (E0 > 50.0) R[I] = A0 * G;
			else if (d->detector_type == XMB_DETECTOR_SI_SDD) R[I] = A0 * G + 1.0 * (0.63 * A3 + 15.0 * A4 * exp(ALFA * (E - E0))) * F;
			else R[I] = A0 * G + 1.0 * (2.7 * A3 + 15.0 * A4 * exp(ALFA * (E - E0))) * F;   /* SiLi and Ge */
			my_sum += R[I];
		}
		for (int I = 0; I <= I0 + 100; I++) {
			if (I >= nch) break;
			conv[I] += R[I] * temp[I0] / my_sum;
		}
	}
	free(R);
}

/* xmi_detector_convolute_spectrum (:339-580).  noconv is modified in place (absorption correction, escape
 * and pile-up act on the caller's array, exactly as the reference's pointer remap :412-413 does). */
void orc_detector_convolute_spectrum(const xmb_input *in, const xmb_xrl_provider *xrl, double *noconv, double *conv,
                                     const xmb_main_options *opt, const xmb_escape_ratios *er, int n_interactions, uint64_t seed) {
	const int nch = in->detector->nchannels;
	for (int i = 0; i < nch; i++) noconv[i] *= orc_detector_correction(in, xrl, i * in->detector->gain + in->detector->zero);
	if (opt->use_escape_peaks == 1 && er) detector_escape(noconv, nch, in, xrl, er);
	if (opt->use_sum_peaks == 1) detector_sum_peaks(noconv, nch, in, seed, n_interactions);
	orc_detector_gaussian(in, noconv, conv);
	if (opt->use_poisson == 1) {                                                 /* :862-905 */
		orc_rng rng;
		orc_rng_init(&rng, seed, (uint64_t)(1000 + n_interactions), ORC_TAG_DETECTOR);
		for (int i = 0; i < nch; i++) if (conv[i] <= 4294967295.0 && conv[i] > 1.0) conv[i] = ran_poisson(&rng, conv[i]);
	}
}

/* xmi_detector_convolute_history (:291-337); history is the exported C array [100][385][n_int] */
void orc_detector_convolute_history(const xmb_input *in, const xmb_xrl_provider *xrl, double *history) {
	const int n_int = in->general->n_interactions_trajectory;
	for (int k = 0; k < 100; k++)
		for (int j = 0; j < 383; j++)
			for (int i = 0; i < n_int; i++) {
				double counts = history[((size_t)k * 385 + j) * n_int + i];
				if (counts > 0.0) {
					double line_energy = xrl->LineEnergy(k + 1, -(j + 1));
					if (line_energy > 0.0) history[((size_t)k * 385 + j) * n_int + i] = counts * orc_detector_correction(in, xrl, line_energy);
				}
			}
}
