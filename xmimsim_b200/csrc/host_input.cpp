// host_input.cpp -- input tree handling and one-time geometry precompute.
//
// Replaces (reference): xmi_input_C2F  src/xmi_aux_f.F90:782-1025
//                       xmi_init_input src/xmi_main.F90:1741-1918
//                       xmi_main_options_new defaults src/xmi_data_structs.c:2531-2565
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include "engine.h"

static thread_local char g_err[512] = "";

void xmb_set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char *xmb_last_error(void) { return g_err; }
extern "C" const char *xmb_version(void) { return "xmimsim-b200 0.1 (sm_100a)"; }

XmbInputF *xmb_as_input(xmb_inputFPtr p) {
	XmbInputF *h = static_cast<XmbInputF *>(p);
	if (!h || h->magic != XMB_MAGIC_INPUT) { xmb_set_error("not an xmb input handle"); return nullptr; }
	return h;
}
XmbHdf5F *xmb_as_hdf5(xmb_hdf5FPtr p) {
	XmbHdf5F *h = static_cast<XmbHdf5F *>(p);
	if (!h || h->magic != XMB_MAGIC_HDF5) { xmb_set_error("not an xmb table handle"); return nullptr; }
	return h;
}

static char *dup_str(const char *s) { return s ? strdup(s) : nullptr; }

static void copy_layer(const xmb_layer &a, xmb_layer &b, bool normalise) {
	b.n_elements = a.n_elements;
	b.density = a.density;
	b.thickness = a.thickness;
	b.Z = (int *)malloc(sizeof(int) * std::max(1, a.n_elements));
	b.weight = (double *)malloc(sizeof(double) * std::max(1, a.n_elements));
	double sum = 0.0;
	for (int i = 0; i < a.n_elements; i++) { b.Z[i] = a.Z[i]; b.weight[i] = a.weight[i]; sum += a.weight[i]; }
	// reference normalises every layer's weights to sum 1 at C->F copy (src/xmi_aux_f.F90:857-858)
	if (normalise && sum > 0.0) for (int i = 0; i < a.n_elements; i++) b.weight[i] /= sum;
}

static xmb_layer *copy_layers(const xmb_layer *a, int n) {
	if (n <= 0 || !a) return nullptr;
	xmb_layer *b = (xmb_layer *)calloc(n, sizeof(xmb_layer));
	for (int i = 0; i < n; i++) copy_layer(a[i], b[i], true);
	return b;
}

static void free_layers(xmb_layer *l, int n) {
	if (!l) return;
	for (int i = 0; i < n; i++) { free(l[i].Z); free(l[i].weight); }
	free(l);
}

extern "C" int xmb_input_C2F(const xmb_input *input, xmb_inputFPtr *out) {
	if (!input || !out || !input->general || !input->composition || !input->geometry ||
	    !input->excitation || !input->absorbers || !input->detector) {
		xmb_set_error("xmb_input_C2F: incomplete input tree");
		return 0;
	}
	const xmb_composition *c = input->composition;
	if (c->n_layers < 1 || c->reference_layer < 1 || c->reference_layer > c->n_layers) {
		xmb_set_error("xmb_input_C2F: invalid composition");
		return 0;
	}
	for (int i = 0; i < c->n_layers; i++)
		for (int j = 0; j < c->layers[i].n_elements; j++)
			if (c->layers[i].Z[j] < 1 || c->layers[i].Z[j] > 94) { xmb_set_error("xmb_input_C2F: Z out of 1..94"); return 0; }
	XmbInputF *h = new XmbInputF();
	xmb_input &d = h->in;
	d.general = (xmb_general *)calloc(1, sizeof(xmb_general));
	*d.general = *input->general;
	d.general->outputfile = dup_str(input->general->outputfile);
	d.general->comments = dup_str(input->general->comments);
	d.composition = (xmb_composition *)calloc(1, sizeof(xmb_composition));
	d.composition->n_layers = c->n_layers;
	d.composition->reference_layer = c->reference_layer;
	d.composition->layers = copy_layers(c->layers, c->n_layers);
	d.geometry = (xmb_geometry *)calloc(1, sizeof(xmb_geometry));
	*d.geometry = *input->geometry;
	d.excitation = (xmb_excitation *)calloc(1, sizeof(xmb_excitation));
	const xmb_excitation *e = input->excitation;
	d.excitation->n_discrete = e->n_discrete;
	d.excitation->n_continuous = e->n_continuous;
	if (e->n_discrete > 0) {
		d.excitation->discrete = (xmb_energy_discrete *)malloc(sizeof(xmb_energy_discrete) * e->n_discrete);
		memcpy(d.excitation->discrete, e->discrete, sizeof(xmb_energy_discrete) * e->n_discrete);
	}
	if (e->n_continuous > 0) {
		d.excitation->continuous = (xmb_energy_continuous *)malloc(sizeof(xmb_energy_continuous) * e->n_continuous);
		memcpy(d.excitation->continuous, e->continuous, sizeof(xmb_energy_continuous) * e->n_continuous);
	}
	d.absorbers = (xmb_absorbers *)calloc(1, sizeof(xmb_absorbers));
	d.absorbers->n_exc_layers = input->absorbers->n_exc_layers;
	d.absorbers->exc_layers = copy_layers(input->absorbers->exc_layers, input->absorbers->n_exc_layers);
	d.absorbers->n_det_layers = input->absorbers->n_det_layers;
	d.absorbers->det_layers = copy_layers(input->absorbers->det_layers, input->absorbers->n_det_layers);
	d.detector = (xmb_detector *)calloc(1, sizeof(xmb_detector));
	*d.detector = *input->detector;
	d.detector->crystal_layers = copy_layers(input->detector->crystal_layers, input->detector->n_crystal_layers);
	*out = h;
	return 1;
}

extern "C" const xmb_input *xmb_input_F2C(xmb_inputFPtr p) {
	XmbInputF *h = xmb_as_input(p);
	return h ? &h->in : nullptr;
}

extern "C" void xmb_free_input_F(xmb_inputFPtr *p) {
	if (!p || !*p) return;
	XmbInputF *h = xmb_as_input(*p);
	if (!h) return;
	xmb_input &d = h->in;
	free(d.general->outputfile); free(d.general->comments); free(d.general);
	free_layers(d.composition->layers, d.composition->n_layers); free(d.composition);
	free(d.geometry);
	free(d.excitation->discrete); free(d.excitation->continuous); free(d.excitation);
	free_layers(d.absorbers->exc_layers, d.absorbers->n_exc_layers);
	free_layers(d.absorbers->det_layers, d.absorbers->n_det_layers); free(d.absorbers);
	free_layers(d.detector->crystal_layers, d.detector->n_crystal_layers); free(d.detector);
	h->magic = 0;
	delete h;
	*p = nullptr;
}

static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const double *a, const double *b, double *c) {
	c[0] = a[1] * b[2] - a[2] * b[1];
	c[1] = a[2] * b[0] - a[0] * b[2];
	c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void normalize3(double *a) { double n = std::sqrt(dot3(a, a)); a[0] /= n; a[1] /= n; a[2] /= n; }

extern "C" int xmb_init_input(xmb_inputFPtr *p) {
	XmbInputF *h = p ? xmb_as_input(*p) : nullptr;
	if (!h) return 0;
	xmb_geometry &g = *h->in.geometry;
	xmb_derived &d = h->der;
	// geometry normalisations: sample normal flipped to +z (src/xmi_main.F90:1759-1763)
	normalize3(g.n_sample_orientation);
	if (g.n_sample_orientation[2] < 0.0) for (int i = 0; i < 3; i++) g.n_sample_orientation[i] = -g.n_sample_orientation[i];
	normalize3(g.n_detector_orientation);
	memcpy(d.n_sample_orientation, g.n_sample_orientation, sizeof(double) * 3);
	memcpy(d.n_detector_orientation, g.n_detector_orientation, sizeof(double) * 3);
	d.detector_radius = std::sqrt(g.area_detector / M_PI);
	d.collimator_height = g.collimator_height;
	if (g.collimator_height > 0.0 && g.collimator_diameter > 0.0) {
		d.collimator_present = 1;
		d.collimator_radius = g.collimator_diameter / 2.0;
		if (d.collimator_radius >= d.detector_radius) {
			xmb_set_error("Non conical collimator found");   // reference exits here (:1785-1789)
			return 0;
		}
		d.half_apex = std::atan((d.detector_radius - d.collimator_radius) / g.collimator_height);
		d.vertex[0] = d.detector_radius / std::tan(d.half_apex);
		d.vertex[1] = d.vertex[2] = 0.0;
	} else {
		d.collimator_present = 0;
		d.collimator_radius = 0.0;
		d.half_apex = 0.0;
		d.vertex[0] = d.vertex[1] = d.vertex[2] = 0.0;
	}
	// detector frame: x' = detector normal, y' = (y or x) cross x', z' = x' cross y' (:1803-1827)
	double nx[3], ny[3], nz[3];
	memcpy(nx, g.n_detector_orientation, sizeof(nx));
	const double ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0};
	if (std::fabs(dot3(nx, ex)) > 1.0e-6) cross3(ey, nx, ny); else cross3(ex, nx, ny);
	normalize3(ny);
	cross3(nx, ny, nz);
	// matrix with columns nx, ny, nz (:1836-1841), row-major storage
	double *A = d.ndo_new;
	for (int i = 0; i < 3; i++) { A[i * 3 + 0] = nx[i]; A[i * 3 + 1] = ny[i]; A[i * 3 + 2] = nz[i]; }
	// direct 3x3 inverse (src/xmi_aux_f.F90:2076-2106)
	double det = A[0] * A[4] * A[8] - A[0] * A[5] * A[7] - A[1] * A[3] * A[8] + A[1] * A[5] * A[6] +
	             A[2] * A[3] * A[7] - A[2] * A[4] * A[6];
	double di = 1.0 / det;
	double *B = d.ndo_inv;
	B[0] = +di * (A[4] * A[8] - A[5] * A[7]);
	B[3] = -di * (A[3] * A[8] - A[5] * A[6]);
	B[6] = +di * (A[3] * A[7] - A[4] * A[6]);
	B[1] = -di * (A[1] * A[8] - A[2] * A[7]);
	B[4] = +di * (A[0] * A[8] - A[2] * A[6]);
	B[7] = -di * (A[0] * A[7] - A[1] * A[6]);
	B[2] = +di * (A[1] * A[5] - A[2] * A[4]);
	B[5] = -di * (A[0] * A[5] - A[2] * A[3]);
	B[8] = +di * (A[0] * A[4] - A[1] * A[3]);
	// nominal detector solid angle seen from the reference-layer surface (:1862-1868)
	double dx = g.p_detector_window[0], dy = g.p_detector_window[1], dz = g.p_detector_window[2] - g.d_sample_source;
	double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
	d.detector_solid_angle = 2.0 * M_PI * (1.0 - std::cos(std::atan(d.detector_radius / dist)));
	for (int i = 0; i < 3; i++) d.n_sample_orientation_det[i] = dot3(&B[i * 3], g.n_sample_orientation);
	// layer coordinates along the beam axis (:1885-1907)
	const xmb_composition &c = *h->in.composition;
	int n = c.n_layers, ref = c.reference_layer - 1;
	h->thickness_along_Z.assign(n, 0.0);
	h->Z_coord_begin.assign(n, 0.0);
	h->Z_coord_end.assign(n, 0.0);
	for (int j = 0; j < n; j++) h->thickness_along_Z[j] = std::fabs(c.layers[j].thickness / g.n_sample_orientation[2]);
	h->Z_coord_begin[ref] = 0.0 + g.d_sample_source;
	h->Z_coord_end[ref] = h->thickness_along_Z[ref] + g.d_sample_source;
	for (int j = ref + 1; j < n; j++) {
		h->Z_coord_begin[j] = h->Z_coord_end[j - 1];
		h->Z_coord_end[j] = h->Z_coord_begin[j] + h->thickness_along_Z[j];
	}
	for (int j = ref - 1; j >= 0; j--) {
		h->Z_coord_end[j] = h->Z_coord_begin[j + 1];
		h->Z_coord_begin[j] = h->Z_coord_end[j] - h->thickness_along_Z[j];
	}
	d.n_layers = n;
	d.thickness_along_Z = h->thickness_along_Z.data();
	d.Z_coord_begin = h->Z_coord_begin.data();
	d.Z_coord_end = h->Z_coord_end.data();
	h->inited = true;
	return 1;
}

extern "C" const xmb_derived *xmb_get_derived(xmb_inputFPtr p) {
	XmbInputF *h = xmb_as_input(p);
	return (h && h->inited) ? &h->der : nullptr;
}

extern "C" void xmb_main_options_defaults(xmb_main_options *o) {
	if (!o) return;
	memset(o, 0, sizeof(*o));
	o->use_M_lines = 1;
	o->use_cascade_auger = 1;
	o->use_cascade_radiative = 1;
	o->use_variance_reduction = 1;
	o->use_sum_peaks = 0;
	o->use_escape_peaks = 1;
	o->escape_ratios_mode = 0;
	o->verbose = 0;
	o->use_poisson = 0;
	o->use_gpu = 1;
	o->omp_num_threads = 1;
	o->extra_verbose = 0;
	o->custom_detector_response = nullptr;
	o->use_advanced_compton = 0;
	o->use_default_seeds = 0;
}

double xmb_host_mu_layer(const xmb_xrl_provider *xrl, const xmb_layer *layer, double E) {
	// xmi_mu_calc for one layer (src/xmi_aux_f.F90:1125-1141)
	double rv = 0.0;
	for (int i = 0; i < layer->n_elements; i++) rv += xrl->CS_Total_Kissel(layer->Z[i], E) * layer->weight[i];
	return rv;
}

// ---- escape-ratio mode input (src/xmi_detector.c:91-141, src/xmi_main.F90:1687-1738) -------------------------
extern "C" xmb_escape_ratios_options xmb_get_default_escape_ratios_options(void) {
	xmb_escape_ratios_options rv = {1990, 1999, 500000, 1.0, 0.1, 0.1, 0.1};
	return rv;
}

extern "C" int xmb_escape_ratios_input(const xmb_input *input, const xmb_escape_ratios_options *ero, xmb_inputFPtr *out) {
	if (!input || !ero || !out || !input->detector || !input->general || !input->geometry || !input->absorbers) {
		xmb_set_error("xmb_escape_ratios_input: bad arguments");
		return 0;
	}
	if (input->detector->n_crystal_layers < 1) { xmb_set_error("xmb_escape_ratios_input: detector has no crystal"); return 0; }
	if (ero->n_input_energies < 1 || ero->n_compton_output_energies < 1 || ero->n_photons < 1 ||
	    ero->n_photons >= (1L << 23) || ero->n_input_energies > 65535) {
		xmb_set_error("xmb_escape_ratios_input: options out of range");
		return 0;
	}
	// shallow tree with the reference's overrides; xmb_input_C2F takes the deep copy
	xmb_general gen = *input->general;
	gen.n_interactions_trajectory = 1;
	gen.n_photons_line = ero->n_photons;
	xmb_composition comp;
	comp.n_layers = input->detector->n_crystal_layers;
	comp.layers = input->detector->crystal_layers;
	comp.reference_layer = 1;
	xmb_geometry geo = *input->geometry;
	geo.d_sample_source = 1.0;
	geo.d_source_slit = 1.0;
	geo.slit_size_x = 0.0001;
	geo.slit_size_y = 0.0001;
	geo.n_sample_orientation[0] = 0.0; geo.n_sample_orientation[1] = 0.0; geo.n_sample_orientation[2] = 1.0;
	std::vector<xmb_energy_discrete> lines((size_t)ero->n_input_energies);
	for (long i = 0; i < ero->n_input_energies; i++) {
		xmb_energy_discrete &e = lines[i];
		memset(&e, 0, sizeof(e));
		e.energy = ero->input_energy_min + i * ero->input_energy_delta;   // src/xmi_main.F90:5534-5537
		e.horizontal_intensity = 0.5; e.vertical_intensity = 0.5;
		e.distribution_type = XMB_DISCRETE_MONOCHROMATIC;
	}
	xmb_excitation exc;
	exc.n_discrete = (int)lines.size(); exc.discrete = lines.data();
	exc.n_continuous = 0; exc.continuous = nullptr;
	xmb_absorbers abs = *input->absorbers;
	abs.n_exc_layers = 0; abs.exc_layers = nullptr;
	xmb_input esc;
	esc.general = &gen; esc.composition = &comp; esc.geometry = &geo; esc.excitation = &exc; esc.absorbers = &abs;
	esc.detector = input->detector;
	if (!xmb_input_C2F(&esc, out)) return 0;
	if (!xmb_init_input(out)) { xmb_free_input_F(out); return 0; }
	return 1;
}
