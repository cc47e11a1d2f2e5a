// host_ebel.cpp -- X-ray tube spectrum after Ebel (X-Ray Spectrom. 28 (1999) 255; 32 (2003) 46), the source
// generator that feeds the history engine with continuous intervals + characteristic lines.
// Replaces xmi_tube_ebel (include/xmi_ebel.h; src/xmi_ebel.F90:114-521).  Host only.
#include <cmath>
#include <vector>
#include "engine.h"
#include "xmb_lines.h"

namespace {

const double TUBE_MINIMUM_ENERGY = 1.0;   // src/xmi_ebel.F90:18
const double DEG2RAD = 0.01745329;        // :26 (the reference's truncated constant)

// L-shell fluorescence yields used by the model (:27-38; entry 31 is 0.122 in the reference as well)
const double omegaL[100] = {
    0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 2.17E-4, 3.04E-4, 4.15E-4, 5.53E-4, 7.24E-4, 9.30E-4, 0.00118, 0.00147, 0.00181, 0.00221,
    0.00268, 0.00321, 0.00381, 0.00450, 0.00527, 0.00614, 0.00711, 0.00819, 0.00939, 0.0107, 0.122, 0.0138, 0.0155, 0.0174, 0.0195, 0.0218, 0.0242,
    0.0263, 0.0285, 0.0309, 0.0335, 0.0363, 0.0393, 0.0425, 0.0459, 0.0495, 0.0534, 0.0575, 0.0618, 0.0655, 0.0714, 0.0765, 0.0820, 0.0877, 0.0938,
    0.100, 0.107, 0.114, 0.121, 0.129, 0.137, 0.145, 0.153, 0.163, 0.172, 0.182, 0.192, 0.202, 0.212, 0.223, 0.234, 0.245, 0.257, 0.269, 0.281, 0.293,
    0.305, 0.318, 0.331, 0.343, 0.356, 0.369, 0.382, 0.395, 0.409, 0.422, 0.435, 0.448, 0.461, 0.474, 0.486, 0.499, 0.511, 0.524, 0.536, 0.548, 0.560,
    0.572, 0.583, 0.595};

// natural cubic spline through (x, y) (src/xmi_spline.c:38-140): second derivatives by the tridiagonal sweep
struct Spline {
	std::vector<double> x, a, b, c, d;
	Spline(const double *xs, const double *ys, size_t n) : x(xs, xs + n), a(ys, ys + n), b(n - 1), c(n, 0.0), d(n - 1) {
		const size_t m = n - 1;
		std::vector<double> h(m), alpha(m, 0.0), l(n, 1.0), mu(n, 0.0), z(n, 0.0);
		for (size_t i = 0; i < m; i++) h[i] = x[i + 1] - x[i];
		for (size_t i = 1; i < m; i++) alpha[i] = 3 * (a[i + 1] - a[i]) / h[i] - 3 * (a[i] - a[i - 1]) / h[i - 1];
		for (size_t i = 1; i < m; i++) {
			l[i] = 2 * (x[i + 1] - x[i - 1]) - h[i - 1] * mu[i - 1];
			mu[i] = h[i] / l[i];
			z[i] = (alpha[i] - h[i - 1] * z[i - 1]) / l[i];
		}
		for (size_t j = m; j-- > 0;) {
			c[j] = z[j] - mu[j] * c[j + 1];
			b[j] = (a[j + 1] - a[j]) / h[j] - h[j] * (c[j + 1] + 2 * c[j]) / 3;
			d[j] = (c[j + 1] - c[j]) / 3 / h[j];
		}
	}
	double operator()(double v) const {
		size_t j = 0;
		while (j + 1 < b.size() && x[j + 1] <= v) j++;          // last knot interval also extrapolates, as the reference's loop does
		const double dx = v - x[j];
		return a[j] + b[j] * dx + c[j] * dx * dx + d[j] * dx * dx * dx;
	}
};

// thick-target (or transmission) absorption term shared by continuum and lines (:247-259, :351-361)
double absorption_term(double tau, double rhoz, double sinfactor, double sinalphax, bool transmission, double rho_t, bool &valid) {
	const double rhelp = tau * 2.0 * rhoz * sinfactor;
	valid = rhelp > 0.0;
	if (!valid) return 0.0;
	if (!transmission) return (1.0 - std::exp(-rhelp)) / rhelp;
	return (std::exp(-tau * (rho_t - 2.0 * rhoz) / sinalphax) - std::exp(-tau * rho_t / sinalphax)) / rhelp;
}

}  // namespace

extern "C" int xmb_tube_ebel(const xmb_xrl_provider *xrl, const xmb_layer *anode, const xmb_layer *window, const xmb_layer *filter,
                             double voltage, double current, double angle_electron, double angle_xray, double delta_energy,
                             double solid_angle, int transmission, size_t n_eff, const double *eff_E, const double *eff,
                             xmb_excitation **out) {
	if (!xrl) xrl = xmb_xrl_surrogate();
	if (!anode || !out || anode->n_elements < 1 || !(delta_energy > 0.0) || !(voltage > TUBE_MINIMUM_ENERGY)) {
		xmb_set_error("xmb_tube_ebel: bad arguments");
		return 0;
	}
	const int Z = anode->Z[0];
	const double sinalphae = std::sin(DEG2RAD * angle_electron), sinalphax = std::sin(DEG2RAD * angle_xray);
	const double sinfactor = sinalphae / sinalphax, rho_t = anode->density * anode->thickness;
	// ---- continuum grid (:203-217): 1 keV, 1 + dE, ... and the tube voltage as last point ------------------------
	size_t ncont = (size_t)std::floor((voltage - TUBE_MINIMUM_ENERGY) / delta_energy) + 1;
	if (voltage / delta_energy != std::nearbyint(voltage / delta_energy)) ncont++;
	std::vector<xmb_energy_continuous> cont(ncont);
	for (auto &c : cont) c = xmb_energy_continuous{};
	for (size_t i = 0; i + 1 < ncont; i++) cont[i].energy = TUBE_MINIMUM_ENERGY + i * delta_energy;
	cont[ncont - 1].energy = voltage;
	// ---- model constants (:219-231) ------------------------------------------------------------------------------------
	const double const1 = 1.35E+09, const2_K = 5.0E+13, zk = 2.0, zl = 8.0, bk = 0.35, bl = 0.25;
	const double x = 1.109 - 0.00435 * Z + 0.00175 * voltage;
	const double m = 0.1382 - 0.9211 / std::sqrt((double)Z);
	const double logz = std::log((double)Z);
	const double eta = (0.1904 - 0.2236 * logz + 0.1292 * logz * logz - 0.0149 * logz * logz * logz) * std::pow(voltage, m);
	const double p3 = 0.787E-05 * std::sqrt(Z * 0.0135) * std::pow(voltage, 1.5) + 0.735E-06 * voltage * voltage;
	const double rhozmax = xrl->AtomicWeight(Z) * p3 / Z;
	const double pa = 0.49269 - 1.09870 * eta + 0.78557 * eta * eta, pb = 0.70256 - 1.09865 * eta + 1.00460 * eta * eta;
	auto depth = [&](double logu0) { return rhozmax * (logu0 * pa / (pb + logu0)); };
	// (the reference calls xraylib's CS_Total here, :243-244; the provider carries the Kissel total only)
	for (size_t i = 0; i < ncont; i++) {
		const double E = cont[i].energy + delta_energy / 2.0;
		const double u0 = voltage / E;
		bool ok;
		const double term = absorption_term(xrl->CS_Total_Kissel(Z, E), depth(std::log(u0)), sinfactor, sinalphax, transmission != 0, rho_t, ok);
		cont[i].horizontal_intensity = ok ? const1 * Z * std::pow(u0 - 1.0, x) * term : 0.0;
	}
	// ---- characteristic lines (:262-391) --------------------------------------------------------------------------------------
	std::vector<xmb_energy_discrete> disc;
	for (int l = 1; l <= XMB_L3Q1; l++) {
		if (!(xrl->RadRate(Z, -l) > 0.0 && xrl->LineEnergy(Z, -l) > TUBE_MINIMUM_ENERGY)) continue;
		int shell;
		if (l <= xmb_shell_line_last[0]) shell = 0;
		else if (l >= xmb_shell_line_first[1] && l <= xmb_shell_line_last[1]) shell = 1;
		else if (l >= xmb_shell_line_first[2] && l <= xmb_shell_line_last[2]) shell = 2;
		else if (l >= xmb_shell_line_first[3] && l <= XMB_L3Q1) shell = 3;
		else continue;
		const double edge = xrl->EdgeEnergy(Z, shell);
		if (edge == 0.0 || edge > voltage) continue;
		const double E = xrl->LineEnergy(Z, -l);
		const double u0 = voltage / edge, logu0 = std::log(u0);
		double oneovers = (std::sqrt(u0) * logu0 + 2.0 * (1.0 - std::sqrt(u0))) / (u0 * logu0 + 1.0 - u0);
		oneovers = 1.0 + 16.05 * std::sqrt(0.0135 * Z / edge) * oneovers;
		// (:330-334: `disc_lines > L1L2_LINE` on negative macros = the K lines)
		oneovers = ((l < xmb_shell_line_first[1] ? zk * bk : zl * bl) / Z) * (u0 * logu0 + 1.0 - u0) * oneovers;
		const double r = 1.0 - 0.0081517 * Z + 3.613e-05 * Z * Z + 0.009583 * Z * std::exp(-u0) + voltage * 0.001141;
		bool ok;
		double term = absorption_term(xrl->CS_Total_Kissel(Z, E), depth(logu0), sinfactor, sinalphax, transmission != 0, rho_t, ok);
		if (!ok) term = xrl->CS_Total_Kissel(Z, E) * 2.0 * depth(logu0) * sinfactor;   // the reference keeps rhelp itself here (WHERE leaves it)
		const double fcorr = Z >= 80 ? 1.0 : -0.4814 + 0.03781 * Z - 2.413E-4 * Z * Z;
		xmb_energy_discrete d{};
		d.energy = E;
		d.distribution_type = XMB_DISCRETE_MONOCHROMATIC;
		const double rr = xrl->RadRate(Z, -l);
		if (shell == 0) d.horizontal_intensity = term * const2_K * oneovers * r * rr * xrl->FluorYield(Z, 0);
		else if (shell == 1) d.horizontal_intensity = term * fcorr * 0.71E13 * oneovers * r * rr * omegaL[Z - 1];
		else if (shell == 2) d.horizontal_intensity = term * fcorr * 2.70E13 * oneovers * r * rr * omegaL[Z - 1];
		else d.horizontal_intensity = term * 4.94E13 * oneovers * r * rr * omegaL[Z - 1];
		disc.push_back(d);
	}
	// ---- window, filter (first element's cross section only, as the reference: :393-420), solid angle, current -----------
	auto attenuate = [&](const xmb_layer *lay, bool half_shift) {
		if (!lay || lay->n_elements < 1) return;
		for (auto &d : disc) d.horizontal_intensity *= std::exp(-lay->density * lay->thickness * xrl->CS_Total_Kissel(lay->Z[0], d.energy));
		for (auto &c : cont)
			c.horizontal_intensity *= std::exp(-lay->density * lay->thickness * xrl->CS_Total_Kissel(lay->Z[0], c.energy + (half_shift ? delta_energy / 2.0 : 0.0)));
	};
	attenuate(window, true);     // :400-404 evaluates the window at E + dE/2 ...
	attenuate(filter, false);    // ... and the filter at E (:414-418)
	for (auto &c : cont) c.horizontal_intensity *= solid_angle * current / 2.0;
	for (auto &d : disc) d.horizontal_intensity *= solid_angle * current / 2.0;
	if (n_eff > 1 && eff_E && eff) {   // transmission-efficiency curve (:443-458)
		const Spline s(eff_E, eff, n_eff);
		for (auto &c : cont) c.horizontal_intensity *= s(c.energy);
		for (auto &d : disc) d.horizontal_intensity *= s(d.energy);
	}
	for (auto &c : cont) c.vertical_intensity = c.horizontal_intensity;
	for (auto &d : disc) d.vertical_intensity = d.horizontal_intensity;
	xmb_excitation *e = (xmb_excitation *)calloc(1, sizeof(xmb_excitation));
	e->n_continuous = (int)cont.size();
	e->continuous = (xmb_energy_continuous *)malloc(sizeof(xmb_energy_continuous) * cont.size());
	memcpy(e->continuous, cont.data(), sizeof(xmb_energy_continuous) * cont.size());
	e->n_discrete = (int)disc.size();
	if (!disc.empty()) {
		e->discrete = (xmb_energy_discrete *)malloc(sizeof(xmb_energy_discrete) * disc.size());
		memcpy(e->discrete, disc.data(), sizeof(xmb_energy_discrete) * disc.size());
	}
	*out = e;
	return 1;
}

extern "C" void xmb_free_excitation(xmb_excitation **e) {
	if (!e || !*e) return;
	free((*e)->discrete); free((*e)->continuous); free(*e);
	*e = nullptr;
}
