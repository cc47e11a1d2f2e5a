// host_tables.cpp -- builds the physics-table bundle from a cross-section provider.
//
// Replaces (reference): xmi_db / xmi_db_Z_specific / xmi_db_Z_independent
//                         src/xmi_data.c:149-487, src/xmi_data_f.F90:880-1551   (generator)
//                       xmi_init_from_hdf5 / xmi_update_input_from_hdf5
//                         src/xmi_data_f.F90:98-848                             (loader)
//                       LineEnergies + precalc_mu_cs prologue of xmi_main_msim
//                         src/xmi_main.F90:203-237
// and pre-tabulates, on one common energy-node grid, what the reference asks xraylib for inside
// the photon loop.  Everything here is host-side, one-time, untimed set-up.
#include <cmath>
#include <algorithm>
#include <set>
#include "engine.h"
#include "xmb_lines.h"

static const double KEV2ANGST = 12.39841930;
static const double MEC2 = 510.998928;

namespace {

// reference grid constants (src/xmi_data.c:150-152, src/xmi_data_f.F90:904-907)
const int N_ICDF_E_FULL = 400, N_ICDF_R = 2000, N_IP_E_FULL = 20000, N_PHI_T = 200, N_CP = 10001;
const double LOWE = 0.1, MAXE = 200.0, MAXPZ = 100.0;

// inverse-CDF by the reference's cumulative walk (src/xmi_data_f.F90:1022-1037): the k-th output is
// the abscissa at which the running sum first reaches rs[k]; at most one output per step.
void icdf_walk(const std::vector<double> &mass, const std::vector<double> &absc, const double *rs, int n_r,
               double *out, double first, double last) {
	double run = 0.0;
	size_t l = 0;
	int m = 0;
	const size_t n = mass.size();
	for (int k = 0; k < n_r; k++) out[k] = last;
	for (;;) {
		run += mass[l];
		if (run >= rs[m]) {
			out[m] = absc[l];
			if (m == n_r - 1) break;
			m++;
		}
		if (l == n - 1) break;
		l++;
	}
	out[n_r - 1] = last;
	out[0] = first;
}

}  // namespace

// skip_icdf: leave the theta and Compton-profile inverse CDFs zero-filled (the GPU generator fills them, tables_gpu.cu)
int xmb_build_tables(const xmb_xrl_provider *xrl, xmb_inputFPtr inputF, int quality, xmb_hdf5FPtr *out, bool skip_icdf) {
	XmbInputF *in = xmb_as_input(inputF);
	if (!in || !xrl || !out) { xmb_set_error("xmb_init_from_provider: bad arguments"); return 0; }
	const xmb_composition &comp = *in->in.composition;
	const xmb_excitation &exc = *in->in.excitation;
	XmbHdf5F *h = new XmbHdf5F();
	h->xrl = xrl;
	h->quality = quality;

	// ---- unique elements (ascending) ----------------------------------------------------------
	std::set<int> zs;
	for (int i = 0; i < comp.n_layers; i++)
		for (int j = 0; j < comp.layers[i].n_elements; j++) zs.insert(comp.layers[i].Z[j]);
	h->Z.assign(zs.begin(), zs.end());
	const int nZ = (int)h->Z.size();
	h->uniqZ.assign(95, -1);
	for (int i = 0; i < nZ; i++) h->uniqZ[h->Z[i]] = i;
	h->atomic_weight.resize(nZ);
	for (int i = 0; i < nZ; i++) h->atomic_weight[i] = xrl->AtomicWeight(h->Z[i]);

	// ---- energy window --------------------------------------------------------------------------
	double emax = 0.0;
	bool broad = false;
	for (int i = 0; i < exc.n_discrete; i++) {
		emax = std::max(emax, exc.discrete[i].energy);
		if (exc.discrete[i].distribution_type != XMB_DISCRETE_MONOCHROMATIC) broad = true;
	}
	for (int i = 0; i < exc.n_continuous; i++) emax = std::max(emax, exc.continuous[i].energy);
	if (broad) emax = MAXE;       // Gaussian/Lorentzian lines are accepted up to energy_max = 200 keV
	if (emax <= 0.0 || emax > MAXE) { xmb_set_error("source energies must lie in (0, 200] keV"); delete h; return 0; }
	h->e_max = emax;

	// ---- node grid: reference's interaction-prob grid (20000 pts on 0.1..200) + edge doublets -----
	const double step = (MAXE - LOWE) / (N_IP_E_FULL - 1.0);
	int n_uniform = std::min(N_IP_E_FULL, (int)std::ceil((emax - LOWE) / step) + 2);
	std::vector<double> nodes;
	nodes.reserve(n_uniform + 18 * nZ);
	for (int j = 0; j < n_uniform; j++) nodes.push_back(LOWE + (MAXE - LOWE) * j / (N_IP_E_FULL - 1.0));
	const double top = nodes.back();
	for (int i = 0; i < nZ; i++)
		for (int s = 0; s < 9; s++) {
			double e = xrl->EdgeEnergy(h->Z[i], s);
			if (e < LOWE || e <= 0.0) continue;          // src/xmi_data_f.F90:1085-1099
			if (e + 0.00001 >= top) continue;
			nodes.push_back(e + 0.00001);
			nodes.push_back(e - 0.00001);
		}
	// fluorescence-line energies and the nominal energies of the discrete source lines become nodes: lookups there are exact
	// (broadened lines too: the reference evaluates their excitation-absorber correction at the nominal energy,
	// src/xmi_main.F90:586-592, and a table lookup there must return the provider's value)
	for (int i = 0; i < nZ; i++)
		for (int l = 1; l <= XMB_M5P5; l++) {
			double e = xrl->LineEnergy(h->Z[i], -l);
			if (e >= LOWE && e < top) nodes.push_back(e);
		}
	for (int i = 0; i < exc.n_discrete; i++)
		if (exc.discrete[i].energy >= LOWE && exc.discrete[i].energy < top)
			nodes.push_back(exc.discrete[i].energy);
	std::sort(nodes.begin(), nodes.end());
	nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
	h->node_E = nodes;
	const int nN = (int)nodes.size();
	// bucket index: bucket b covers [E0 + b*dE, E0 + (b+1)*dE)
	const int nB = n_uniform - 1;
	h->bucket_start.assign(nB + 1, 0);
	{
		int k = 0;
		for (int b = 0; b <= nB; b++) {
			double lo = LOWE + step * b;
			while (k < nN && nodes[k] < lo - 1e-12) k++;
			h->bucket_start[b] = std::min(k, nN - 1);
		}
	}

	// ---- energy-dependent cross sections on the nodes ---------------------------------------------
	h->cs_total.resize((size_t)nZ * nN);
	h->cs_photo_total.resize((size_t)nZ * nN);
	h->p_rayl.resize((size_t)nZ * nN);
	h->p_rayl_compt.resize((size_t)nZ * nN);
	h->cs_photo_partial.resize((size_t)nZ * 9 * nN);
	h->cs_vacancy.resize((size_t)4 * nZ * 9 * nN);
	for (int i = 0; i < nZ; i++) (void)xrl->CS_Rayl(h->Z[i], 1.0);   // warm provider caches serially
#pragma omp parallel for schedule(dynamic, 64) collapse(2)
	for (int i = 0; i < nZ; i++)
		for (int k = 0; k < nN; k++) {
			const int Z = h->Z[i];
			const double E = nodes[k];
			const double tot = xrl->CS_Total_Kissel(Z, E);
			const double ray = xrl->CS_Rayl(Z, E), com = xrl->CS_Compt(Z, E);
			h->cs_total[(size_t)i * nN + k] = tot;
			h->cs_photo_total[(size_t)i * nN + k] = xrl->CS_Photo_Total(Z, E);
			h->p_rayl[(size_t)i * nN + k] = ray / tot;                       // src/xmi_data_f.F90:1109-1114
			h->p_rayl_compt[(size_t)i * nN + k] = com / tot + ray / tot;
			for (int s = 0; s < 9; s++) h->cs_photo_partial[((size_t)i * 9 + s) * nN + k] = xrl->CS_Photo_Partial(Z, s, E);
			for (int mode = 1; mode <= 4; mode++) {
				double P[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
				for (int s = 0; s < 9; s++) {
					P[s] = xrl->VacancyCS(Z, s, E, mode, P);
					h->cs_vacancy[(((size_t)(mode - 1) * nZ + i) * 9 + s) * nN + k] = P[s];
				}
			}
		}

	// ---- scattering-angle inverse CDFs ------------------------------------------------------------
	const long n_theta = quality >= 1 ? 100000 : 20000;
	const long n_phi = quality >= 1 ? 100000 : 20000;
	const long n_pz = quality >= 1 ? 10000000 : 400000;
	int n_icdf_E = std::min(N_ICDF_E_FULL, (int)std::ceil((emax - LOWE) / ((MAXE - LOWE) / (N_ICDF_E_FULL - 1.0))) + 2);
	h->icdf_E.resize(n_icdf_E);
	for (int j = 0; j < n_icdf_E; j++) h->icdf_E[j] = LOWE + (MAXE - LOWE) * j / (N_ICDF_E_FULL - 1.0);
	h->icdf_R.resize(N_ICDF_R);
	for (int j = 0; j < N_ICDF_R; j++) h->icdf_R[j] = (double)j / (N_ICDF_R - 1.0);
	h->rayl_theta_icdf.resize((size_t)nZ * n_icdf_E * N_ICDF_R);
	h->compt_theta_icdf.resize((size_t)nZ * n_icdf_E * N_ICDF_R);
	std::vector<double> thetas(n_theta), sinth(n_theta), costh(n_theta), sinhalf(n_theta);
	for (long k = 0; k < n_theta; k++) {
		thetas[k] = M_PI * k / (n_theta - 1.0);
		sinth[k] = std::sin(thetas[k]);
		costh[k] = std::cos(thetas[k]);
		sinhalf[k] = std::sin(thetas[k] * 0.5);
	}
#pragma omp parallel if (!skip_icdf)
	if (!skip_icdf) {
		std::vector<double> fr(n_theta), fc(n_theta), mass(n_theta - 1);
#pragma omp for schedule(dynamic, 1) collapse(2)
		for (int i = 0; i < nZ; i++)
			for (int j = 0; j < n_icdf_E; j++) {
				const int Z = h->Z[i];
				const double E = h->icdf_E[j];
				for (long k = 0; k < n_theta; k++) {
					const double q = E / KEV2ANGST * sinhalf[k];
					const double F = xrl->FF_Rayl(Z, q);
					const double kk = 1.0 / (1.0 + E / MEC2 * (1.0 - costh[k]));
					// DCS_Rayl ~ (1+cos^2) F^2 ; DCS_Compt ~ k^2 (k + 1/k - sin^2) S ; constants cancel
					fr[k] = (1.0 + costh[k] * costh[k]) * F * F * sinth[k];
					fc[k] = kk * kk * (kk + 1.0 / kk - sinth[k] * sinth[k]) * xrl->SF_Compt(Z, q) * sinth[k];
				}
				double sum = 0.0;
				for (long k = 0; k < n_theta - 1; k++) { mass[k] = (fr[k] + fr[k + 1]) * (thetas[k + 1] - thetas[k]) / 2.0; sum += mass[k]; }
				for (long k = 0; k < n_theta - 1; k++) mass[k] /= sum;
				icdf_walk(mass, thetas, h->icdf_R.data(), N_ICDF_R, &h->rayl_theta_icdf[((size_t)i * n_icdf_E + j) * N_ICDF_R], 0.0, M_PI);
				sum = 0.0;
				for (long k = 0; k < n_theta - 1; k++) { mass[k] = (fc[k] + fc[k + 1]) * (thetas[k + 1] - thetas[k]) / 2.0; sum += mass[k]; }
				for (long k = 0; k < n_theta - 1; k++) mass[k] /= sum;
				icdf_walk(mass, thetas, h->icdf_R.data(), N_ICDF_R, &h->compt_theta_icdf[((size_t)i * n_icdf_E + j) * N_ICDF_R], 0.0, M_PI);
			}
	}
	// phi inverse CDF: CDF(phi) = (phi - a sin 2phi)/2pi, a in [0, 0.5]  (src/xmi_data_f.F90:1508-1545)
	h->phi_T.resize(N_PHI_T);
	for (int i = 0; i < N_PHI_T; i++) h->phi_T[i] = 0.5 * i / (N_PHI_T - 1.0);
	h->phi_icdf.assign((size_t)N_PHI_T * N_ICDF_R, 0.0);
#pragma omp parallel for schedule(dynamic, 4)
	for (int i = 0; i < N_PHI_T; i++) {
		double *row = &h->phi_icdf[(size_t)i * N_ICDF_R];
		int k = 0;
		for (long j = 0; j < n_phi; j++) {
			const double phi = 2.0 * M_PI * j / (n_phi - 1.0);
			const double cdf = (phi - h->phi_T[i] * std::sin(2.0 * phi)) / 2.0 / M_PI;
			if (cdf >= h->icdf_R[k]) {
				row[k] = phi;
				if (k == N_ICDF_R - 1) break;
				k++;
			}
		}
		row[0] = 0.0;
		row[N_ICDF_R - 1] = 2.0 * M_PI;
	}
	// Compton-profile inverse CDF (src/xmi_data_f.F90:1162-1186)
	h->cp_R.resize(N_CP);
	for (int i = 0; i < N_CP; i++) h->cp_R[i] = (double)i / (N_CP - 1.0);
	h->cp_icdf.resize((size_t)nZ * N_CP);
#pragma omp parallel if (!skip_icdf)
	if (!skip_icdf) {
		std::vector<double> mass(n_pz - 1), pzs(n_pz);
		for (long k = 0; k < n_pz; k++) pzs[k] = MAXPZ * k / (n_pz - 1.0);
#pragma omp for schedule(dynamic, 1)
		for (int i = 0; i < nZ; i++) {
			const int Z = h->Z[i];
			double prev = xrl->ComptonProfile(Z, pzs[0]), sum = 0.0;
			for (long k = 0; k < n_pz - 1; k++) {
				const double next = xrl->ComptonProfile(Z, pzs[k + 1]);
				mass[k] = (prev + next) * (pzs[k + 1] - pzs[k]) / 2.0 / Z;
				sum += mass[k];
				prev = next;
			}
			for (long k = 0; k < n_pz - 1; k++) mass[k] /= sum;
			icdf_walk(mass, pzs, h->cp_R.data(), N_CP, &h->cp_icdf[(size_t)i * N_CP], 0.0, MAXPZ);
		}
	}

	// ---- form factor / scattering function grid -----------------------------------------------------
	const int n_q = 8192;
	const double q_max = emax / KEV2ANGST * 1.0001;
	h->ff.resize((size_t)nZ * n_q);
	h->sf.resize((size_t)nZ * n_q);
	for (int i = 0; i < nZ; i++)
		for (int k = 0; k < n_q; k++) {
			const double q = q_max * k / (n_q - 1.0);
			h->ff[(size_t)i * n_q + k] = xrl->FF_Rayl(h->Z[i], q);
			h->sf[(size_t)i * n_q + k] = xrl->SF_Compt(h->Z[i], q);
		}

	// ---- atomic constants ---------------------------------------------------------------------------
	h->fluor_yield.assign((size_t)nZ * 9, 0.0);
	h->fluor_yield_corr.assign((size_t)nZ * 9, 0.0);
	h->cos_kron.assign((size_t)nZ * XMB_N_CK, 0.0);
	h->rad_rate.assign((size_t)nZ * 384, 0.0);
	h->line_energy.assign((size_t)nZ * 384, 0.0);
	h->edge_energy.assign((size_t)nZ * 9, 0.0);
	static const int ck_from[XMB_N_CK] = {1, 1, 2, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7};
	static const int ck_to[XMB_N_CK] = {2, 3, 3, 5, 6, 7, 8, 6, 7, 8, 7, 8, 8};
	for (int i = 0; i < nZ; i++) {
		const int Z = h->Z[i];
		for (int s = 0; s < 9; s++) {
			h->fluor_yield[i * 9 + s] = xrl->FluorYield(Z, s);
			h->edge_energy[i * 9 + s] = xrl->EdgeEnergy(Z, s);
		}
		for (int t = 0; t < XMB_N_CK; t++) h->cos_kron[i * XMB_N_CK + t] = xrl->CosKronTransProb(Z, t);
		for (int l = 1; l <= XMB_N_LINES; l++) {
			h->rad_rate[i * 384 + l] = xrl->RadRate(Z, -l);
			h->line_energy[i * 384 + l] = xrl->LineEnergy(Z, -l);
		}
		// corrected yields in terms of the primary vacancy distribution: the closed forms of
		// src/xmi_data_f.F90:1244-1304 are the expansion of corr(s) = w_s + sum_t f_st corr(t)
		for (int s = 8; s >= 0; s--) {
			double c = h->fluor_yield[i * 9 + s];
			for (int t = 0; t < XMB_N_CK; t++)
				if (ck_from[t] == s) c += h->cos_kron[i * XMB_N_CK + t] * h->fluor_yield_corr[i * 9 + ck_to[t]];
			h->fluor_yield_corr[i * 9 + s] = c;
		}
	}

	// Auger transition rates in the reference's enumeration order (src/xmi_main.F90:2441-2470, :2482-4418)
	h->auger_rate.assign((size_t)nZ * XMB_N_AUGER, 0.0);
	if (xrl->AugerRate)
		for (int i = 0; i < nZ; i++) {
			double *a = &h->auger_rate[(size_t)i * XMB_N_AUGER];
			for (int x = 0; x < 8; x++)
				for (int y = 0; y < 30; y++) a[x * 30 + y] = xrl->AugerRate(h->Z[i], 0, 1 + x, 1 + y);
			for (int s = 1; s <= 3; s++)
				for (int x = 0; x < 5; x++)
					for (int y = 0; y < 27; y++) a[240 + 135 * (s - 1) + x * 27 + y] = xrl->AugerRate(h->Z[i], s, 4 + x, 4 + y);
		}

	// ---- per-layer attenuation on the nodes (xmi_mu_calc, src/xmi_aux_f.F90:1109-1141) -----------------
	h->mu_layer.assign((size_t)comp.n_layers * nN, 0.0);
	for (int k = 0; k < comp.n_layers; k++)
		for (int n = 0; n < nN; n++) {
			double mu = 0.0;
			for (int e = 0; e < comp.layers[k].n_elements; e++)
				mu += h->cs_total[(size_t)h->uniqZ[comp.layers[k].Z[e]] * nN + n] * comp.layers[k].weight[e];
			h->mu_layer[(size_t)k * nN + n] = mu;
		}
	// excitation-path absorbers (src/xmi_main.F90:366-372, :586-592): sum of mu*rho*t on the nodes
	h->exc_murhod.assign(nN, 0.0);
	{
		const xmb_absorbers &ab = *in->in.absorbers;
		if (ab.n_exc_layers > 0) {
#pragma omp parallel for schedule(static)
			for (int n = 0; n < nN; n++) {
				double s = 0.0;
				for (int k = 0; k < ab.n_exc_layers; k++)
					s += ab.exc_layers[k].density * ab.exc_layers[k].thickness * xmb_host_mu_layer(xrl, &ab.exc_layers[k], nodes[n]);
				h->exc_murhod[n] = s;
			}
		}
	}

	// ---- publish the view ------------------------------------------------------------------------------
	xmb_tables_host &v = h->view;
	v.nZ = nZ; v.Z = h->Z.data(); v.uniqZ = h->uniqZ.data(); v.atomic_weight = h->atomic_weight.data();
	v.n_nodes = nN; v.node_E = h->node_E.data(); v.bucket_E0 = LOWE; v.bucket_inv_dE = 1.0 / step;
	v.n_buckets = nB; v.bucket_start = h->bucket_start.data();
	v.cs_total = h->cs_total.data(); v.cs_photo_total = h->cs_photo_total.data();
	v.p_rayl = h->p_rayl.data(); v.p_rayl_compt = h->p_rayl_compt.data();
	v.cs_photo_partial = h->cs_photo_partial.data(); v.cs_vacancy = h->cs_vacancy.data();
	v.n_icdf_E = n_icdf_E; v.n_icdf_R = N_ICDF_R; v.icdf_E = h->icdf_E.data(); v.icdf_R = h->icdf_R.data();
	v.rayl_theta_icdf = h->rayl_theta_icdf.data(); v.compt_theta_icdf = h->compt_theta_icdf.data();
	v.n_phi_T = N_PHI_T; v.phi_T = h->phi_T.data(); v.phi_icdf = h->phi_icdf.data();
	v.n_cp = N_CP; v.cp_R = h->cp_R.data(); v.cp_icdf = h->cp_icdf.data();
	v.n_q = n_q; v.q_max = q_max; v.ff = h->ff.data(); v.sf = h->sf.data();
	v.fluor_yield = h->fluor_yield.data(); v.fluor_yield_corr = h->fluor_yield_corr.data();
	v.cos_kron = h->cos_kron.data(); v.rad_rate = h->rad_rate.data(); v.line_energy = h->line_energy.data();
	v.auger_rate = h->auger_rate.data();
	v.n_adv_rows = 0;
	v.edge_energy = h->edge_energy.data();
	v.n_layers = comp.n_layers; v.mu_layer = h->mu_layer.data(); v.exc_murhod = h->exc_murhod.data();
	*out = h;
	return 1;
}

extern "C" int xmb_init_from_provider(const xmb_xrl_provider *xrl, xmb_inputFPtr inputF, int quality, xmb_hdf5FPtr *out) {
	return xmb_build_tables(xrl, inputF, quality, out, false);
}

extern "C" const xmb_tables_host *xmb_get_tables(xmb_hdf5FPtr p) {
	XmbHdf5F *h = xmb_as_hdf5(p);
	return h ? &h->view : nullptr;
}

extern "C" void xmb_free_hdf5_F(xmb_hdf5FPtr *p) {
	if (!p || !*p) return;
	XmbHdf5F *h = xmb_as_hdf5(*p);
	if (!h) return;
	xmb_free_all_device_tables(h);
	h->magic = 0;
	delete h;
	*p = nullptr;
}

// ---- solid-angle grid bounds ---------------------------------------------------------------------
// xmi_solid_angle_inputs_f (src/xmi_solid_angle_f.F90:62-301)
static double penetration_depth(const XmbInputF *in, const xmb_xrl_provider *xrl, double energy, double R) {
	const xmb_composition &c = *in->in.composition;
	const int n = c.n_layers;
	std::vector<double> mu(n);
	double my_sum = 0.0;
	for (int i = 0; i < n; i++) {
		mu[i] = xmb_host_mu_layer(xrl, &c.layers[i], energy);
		my_sum += mu[i] * c.layers[i].density * in->thickness_along_Z[i];
	}
	const double Pabs = -1.0 * expm1(-1.0 * my_sum);
	const double myln = -1.0 * log1p(-1.0 * R * Pabs);
	my_sum = 0.0;
	int m = n - 1;   // reference leaves m unset when the loop never exits early; last layer is the safe reading
	for (int i = 0; i < n; i++) {
		my_sum += mu[i] * c.layers[i].density * in->thickness_along_Z[i];
		if (my_sum > myln) { m = i; break; }
	}
	my_sum = 0.0;
	for (int i = 0; i < m; i++)
		my_sum += (1.0 - (mu[i] * c.layers[i].density) / (mu[m] * c.layers[m].density)) * in->thickness_along_Z[i];
	return my_sum + myln / (mu[m] * c.layers[m].density) + in->Z_coord_begin[0];
}

extern "C" int xmb_solid_angle_inputs(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, xmb_solid_angle **out) {
	XmbInputF *in = xmb_as_input(inputF);
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	if (!in || !h || !in->inited || !out) { xmb_set_error("xmb_solid_angle_inputs: bad arguments"); return 0; }
	const xmb_excitation &exc = *in->in.excitation;
	const long NR = 1024, NT = 1024;   // grid_dims_r_n, grid_dims_theta_n (src/xmi_solid_angle_f.F90:41)
	double e_lo, e_hi;
	if (exc.n_continuous > 1 && exc.n_discrete > 0) {
		e_lo = std::min(exc.continuous[0].energy, exc.discrete[0].energy);
		e_hi = std::max(exc.continuous[exc.n_continuous - 1].energy, exc.discrete[exc.n_discrete - 1].energy);
	} else if (exc.n_continuous > 1) {
		e_lo = exc.continuous[0].energy;
		e_hi = exc.continuous[exc.n_continuous - 1].energy;
	} else if (exc.n_discrete > 0) {
		e_lo = exc.discrete[0].energy;
		e_hi = exc.discrete[exc.n_discrete - 1].energy;
	} else { xmb_set_error("no excitation energies"); return 0; }
	const double S1 = penetration_depth(in, h->xrl, e_lo, 0.00001);
	const double S2 = penetration_depth(in, h->xrl, e_hi, 0.99999);
	const double *pw = in->in.geometry->p_detector_window;
	auto dist = [&](double z) { return std::sqrt(pw[0] * pw[0] + pw[1] * pw[1] + (pw[2] - z) * (pw[2] - z)); };
	const double r_hi = std::max(dist(S1), dist(S2)) * 1.25;
	const double r_lo = r_hi / NR;
	const double t_hi = M_PI / 2.0, t_lo = 0.00001;
	xmb_solid_angle *sa = (xmb_solid_angle *)calloc(1, sizeof(xmb_solid_angle));
	sa->solid_angles = (double *)calloc((size_t)NR * NT, sizeof(double));
	sa->grid_dims_r_vals = (double *)malloc(sizeof(double) * NR);
	sa->grid_dims_theta_vals = (double *)malloc(sizeof(double) * NT);
	sa->grid_dims_r_n = NR;
	sa->grid_dims_theta_n = NT;
	for (long i = 0; i < NR; i++) sa->grid_dims_r_vals[i] = r_lo + (r_hi - r_lo) * (double)i / (double)(NR - 1);
	for (long i = 0; i < NT; i++) sa->grid_dims_theta_vals[i] = t_lo + (t_hi - t_lo) * (double)i / (double)(NT - 1);
	*out = sa;
	return 1;
}

extern "C" void xmb_free_solid_angle(xmb_solid_angle *sa) {
	if (!sa) return;
	free(sa->solid_angles);
	free(sa->grid_dims_r_vals);
	free(sa->grid_dims_theta_vals);
	free(sa->xmi_input_string);
	free(sa);
}

// overridden by the strong definition in history.cu once device layouts exist
__attribute__((weak)) void xmb_free_device_tables(XmbDeviceTables *) {}


// ---- shell-resolved Compton profiles (src/xmi_data_f.F90:1120-1235), built on demand ------------------------------
extern "C" int xmb_tables_enable_advanced_compton(xmb_hdf5FPtr p) {
	XmbHdf5F *h = xmb_as_hdf5(p);
	if (!h) { xmb_set_error("xmb_tables_enable_advanced_compton: bad handle"); return 0; }
	if (h->view.n_adv_rows > 0) return 1;
	const xmb_xrl_provider *xrl = h->xrl;
	if (!xrl->ElectronConfig_Biggs || !xrl->ComptonProfile_Partial) {
		xmb_set_error("use_advanced_compton: the cross-section provider has no ElectronConfig_Biggs / ComptonProfile_Partial");
		return 0;
	}
	const int nZ = (int)h->Z.size();
	const long n_pz = h->quality >= 1 ? 10000000 : 400000;
	h->adv_off.assign(nZ + 1, 0);
	h->adv_shell.clear(); h->adv_config.clear(); h->adv_edge.clear();
	for (int i = 0; i < nZ; i++) {
		h->adv_off[i] = (int)h->adv_shell.size();
		for (int s = 0; s <= 30; s++) {                    // K always, then the occupied subshells (:1126-1145)
			const double n = xrl->ElectronConfig_Biggs(h->Z[i], s);
			if (s > 0 && !(n > 0)) continue;
			h->adv_shell.push_back(s);
			h->adv_config.push_back(n);
			h->adv_edge.push_back(xrl->EdgeEnergy(h->Z[i], s));
		}
	}
	h->adv_off[nZ] = (int)h->adv_shell.size();
	const int rows = (int)h->adv_shell.size();
	h->adv_cdf.assign((size_t)rows * N_CP, 0.0);
	h->adv_qinv.assign((size_t)rows * N_CP, 0.0);
	std::vector<int> row_Z(rows);
	for (int i = 0; i < nZ; i++) for (int r = h->adv_off[i]; r < h->adv_off[i + 1]; r++) row_Z[r] = h->Z[i];
#pragma omp parallel
	{
		std::vector<double> big(n_pz);
#pragma omp for schedule(dynamic, 1)
		for (int r = 0; r < rows; r++) {
			const int Z = row_Z[r], s = h->adv_shell[r];
			double *cdf = &h->adv_cdf[(size_t)r * N_CP], *qinv = &h->adv_qinv[(size_t)r * N_CP];
			// cumulative trapezoid on the coarse grid Qs, normalised to 0.5 at Q = 100 (:1189-1197)
			const double dq = MAXPZ / (N_CP - 1.0);
			double prev = xrl->ComptonProfile_Partial(Z, s, 0.0);
			cdf[0] = 0.0;
			for (int j = 1; j < N_CP; j++) {
				const double next = xrl->ComptonProfile_Partial(Z, s, MAXPZ * j / (N_CP - 1.0));
				cdf[j] = cdf[j - 1] + dq * (next + prev) * 0.5;
				prev = next;
			}
			const double last = cdf[N_CP - 1];
			if (last > 0.0) for (int j = 0; j < N_CP; j++) cdf[j] = cdf[j] * 0.5 / last;
			// inverse on the fine grid (:1199-1232)
			const double dqb = MAXPZ / (n_pz - 1.0);
			prev = xrl->ComptonProfile_Partial(Z, s, 0.0);
			big[0] = 0.0;
			for (long j = 1; j < n_pz; j++) {
				const double next = xrl->ComptonProfile_Partial(Z, s, MAXPZ * j / (n_pz - 1.0));
				big[j] = big[j - 1] + dqb * (next + prev) * 0.5;
				prev = next;
			}
			const double lastb = big[n_pz - 1];
			if (lastb > 0.0) for (long j = 0; j < n_pz; j++) big[j] = big[j] * 0.5 / lastb;
			long pos = 0;
			for (int j = 0; j < N_CP; j++) {
				const double c = 0.5 * j / (N_CP - 1.0);
				while (pos < n_pz - 2 && big[pos + 1] <= c) pos++;      // findpos: big[pos] <= c < big[pos+1]
				const double d = big[pos + 1] - big[pos];
				const double q0 = MAXPZ * pos / (n_pz - 1.0), q1 = MAXPZ * (pos + 1) / (n_pz - 1.0);
				qinv[j] = d > 0.0 ? q0 + (q1 - q0) * (c - big[pos]) / d : q0;
			}
			qinv[0] = 0.0;
		}
	}
	xmb_tables_host &v = h->view;
	v.adv_off = h->adv_off.data(); v.adv_shell = h->adv_shell.data(); v.adv_config = h->adv_config.data();
	v.adv_edge = h->adv_edge.data(); v.adv_cdf = h->adv_cdf.data(); v.adv_qinv = h->adv_qinv.data();
	v.n_adv_rows = rows;
	xmb_free_all_device_tables(h);   // device layouts are rebuilt with the new tables
	return 1;
}
