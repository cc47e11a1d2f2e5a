// host_cache.cpp -- the solid-angle and escape-ratio caches of the reference (~/.local/share/XMI-MSIM/
// xmimsim-solid-angles.h5 / xmimsim-escape-ratios.h5; src/xmi_solid_angle.c:193-790, src/xmi_detector.c:40-435).
// libhdf5 does not exist in this build environment, so the files are a side-car container with the reference's
// logical schema: a "kind" tag, then one entry per cached result (the reference: one HDF5 group) holding the XML string
// of the input that produced it (dataset xmi_input_string) and the same datasets with the same dimensions.  The match
// rules that decide whether an entry can be re-used are the reference's.
//
//   file   := "XMBCACHE" u32 version(2) u32 kind        kind 1 = XMI_HDF5_SOLID_ANGLES, 2 = XMI_HDF5_ESCAPE_RATIOS
//   entry  := u64 payload_bytes, u32 xml_len, xml, u32 tag_len, tag, datasets
// `tag` names the cross-section provider the entry was computed with (xmb_cache_set_provider; the reference has one
// provider, xraylib, and needs no such field): the axes of a solid-angle grid and every escape ratio depend on it, so an
// entry made with the analytic stand-in is never handed to a run on xraylib data, and the other way round.
//   kind 1 := i64 n_theta, i64 n_r, f64 solid_angles[n_theta][n_r], f64 grid_dims_r_vals[n_r], f64 grid_dims_theta_vals[n_theta]
//   kind 2 := i32 n_elements, n_fluo_input_energies, n_compton_input_energies, n_compton_output_energies, i32 Z[],
//             f64 fluo_escape_ratios[n_fluo_in][109][n_elements], f64 fluo_escape_input_energies[], f64 compton_escape_ratios
//             [n_out][n_in], f64 compton_escape_input_energies[], f64 compton_escape_output_energies[]
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>
#include "engine.h"

namespace {
const double COMPARE_THRESHOLD = 1E-10;   // XMI_COMPARE_THRESHOLD, include/xmi_data_structs.h:517
const char MAGIC[8] = {'X', 'M', 'B', 'C', 'A', 'C', 'H', 'E'};

struct Reader {
	FILE *f;
	bool ok = true;
	template <typename T> T get() { T v{}; if (fread(&v, sizeof(T), 1, f) != 1) ok = false; return v; }
	template <typename T> void arr(T *p, size_t n) { if (n && fread(p, sizeof(T), n, f) != n) ok = false; }
};

std::string g_cache_tag;   // provider name stamped into new entries and required of matches

// reads the tag of an entry; false when it differs from the current one (the caller skips the entry)
bool read_tag_matches(Reader &r) {
	const unsigned n = r.get<unsigned>();
	if (!r.ok || n > 4096) { r.ok = false; return false; }
	std::string tag(n, '\0');
	r.arr(&tag[0], n);
	return r.ok && tag == g_cache_tag;
}
void write_tag(FILE *f) {
	const unsigned n = (unsigned)g_cache_tag.size();
	fwrite(&n, 4, 1, f); fwrite(g_cache_tag.data(), 1, n, f);
}

FILE *open_cache(const char *file, unsigned kind, bool create, const char *mode) {
	FILE *f = fopen(file, mode);
	if (!f && create) {
		f = fopen(file, "wb+");
		if (!f) { xmb_set_error("Cannot create cache file %s", file); return nullptr; }
		const unsigned version = 2;
		fwrite(MAGIC, 1, 8, f); fwrite(&version, 4, 1, f); fwrite(&kind, 4, 1, f);
		fflush(f);
		return f;
	}
	if (!f) { xmb_set_error("Cannot open file %s for reading", file); return nullptr; }
	char m[8];
	unsigned version = 0, k = 0;
	if (fread(m, 1, 8, f) != 8 || memcmp(m, MAGIC, 8) != 0 || fread(&version, 4, 1, f) != 1 || fread(&k, 4, 1, f) != 1 || version != 2) {
		xmb_set_error("%s is not a cache file of this library (or of an earlier format without the provider tag)", file); fclose(f); return nullptr;
	}
	if (k != kind) { xmb_set_error("%s has kind %u, expected %u", file, k, kind); fclose(f); return nullptr; }   // the reference's kind attribute check
	return f;
}

// depth at which the cumulative interaction probability reaches R for photons of `energy` entering the layer stack
// along Z (the S1 / S2 extremes of src/xmi_solid_angle.c:487-640)
double extreme_depth(const xmb_input *A, const xmb_xrl_provider *xrl, double energy, double R, const std::vector<double> &tz, double z0) {
	const int n = A->composition->n_layers;
	std::vector<double> mu(n);
	double sum = 0.0;
	for (int i = 0; i < n; i++) {
		mu[i] = xmb_host_mu_layer(xrl, &A->composition->layers[i], energy);
		sum += mu[i] * A->composition->layers[i].density * tz[i];
	}
	const double Pabs = -1.0 * expm1(-1.0 * sum);
	const double myln = -1.0 * log1p(-1.0 * R * Pabs);
	sum = 0.0;
	int m = 0;
	for (m = 0; m < n; m++) { sum += mu[m] * A->composition->layers[m].density * tz[m]; if (sum > myln) break; }
	if (m == n) m = n - 1;
	sum = 0.0;
	for (int i = 0; i < m; i++) sum += (1.0 - (mu[i] * A->composition->layers[i].density) / (mu[m] * A->composition->layers[m].density)) * tz[i];
	return sum + myln / (mu[m] * A->composition->layers[m].density) + z0;
}

void extremes(const xmb_input *A, const xmb_xrl_provider *xrl, double *S1, double *S2) {
	double ns[3] = {A->geometry->n_sample_orientation[0], A->geometry->n_sample_orientation[1], A->geometry->n_sample_orientation[2]};
	const double nn = std::sqrt(ns[0] * ns[0] + ns[1] * ns[1] + ns[2] * ns[2]);
	const int n = A->composition->n_layers, ref = A->composition->reference_layer - 1;
	std::vector<double> tz(n), zb(n, 0.0), ze(n, 0.0);
	for (int i = 0; i < n; i++) tz[i] = std::fabs(A->composition->layers[i].thickness / (ns[2] / nn));
	zb[ref] = 0.0; ze[ref] = tz[ref];
	for (int i = ref + 1; i < n; i++) { zb[i] = ze[i - 1]; ze[i] = zb[i] + tz[i]; }
	for (int i = ref - 1; i >= 0; i--) { ze[i] = zb[i + 1]; zb[i] = ze[i] - tz[i]; }   // (the reference's loop test `i == 0` only walks one layer, :455)
	const xmb_excitation *e = A->excitation;
	double lo, hi;
	if (e->n_continuous > 1 && e->n_discrete > 0) {
		lo = std::min(e->continuous[0].energy, e->discrete[0].energy);
		hi = std::max(e->continuous[e->n_continuous - 1].energy, e->discrete[e->n_discrete - 1].energy);
	} else if (e->n_continuous > 1) { lo = e->continuous[0].energy; hi = e->continuous[e->n_continuous - 1].energy; }
	else { lo = e->discrete[0].energy; hi = e->discrete[e->n_discrete - 1].energy; }
	*S1 = extreme_depth(A, xrl, lo, 0.00001, tz, zb[0]);
	*S2 = extreme_depth(A, xrl, hi, 0.99999, tz, zb[0]);
}

}  // namespace

// xmi_check_solid_angle_match (src/xmi_solid_angle.c:420-670): A = cached (old), B = new input.  The cached grid can be
// re-used when it covers the depth range the new sample can be excited in and the detector geometry is the same.
extern "C" void xmb_cache_set_provider(const xmb_xrl_provider *xrl) { g_cache_tag = xrl && xrl->name ? xrl->name : ""; }

extern "C" int xmb_check_solid_angle_match(const xmb_input *A, const xmb_input *B, const xmb_xrl_provider *xrl) {
	if (!xrl) xrl = xmb_xrl_surrogate();
	double S1a, S2a, S1b, S2b;
	extremes(A, xrl, &S1a, &S2a);
	extremes(B, xrl, &S1b, &S2b);
	if (S2a - S2b < -0.0001 || S1a - S1b > 0.0001) return 0;
	const xmb_geometry *a = A->geometry, *b = B->geometry;
	if (std::fabs(a->p_detector_window[0] - b->p_detector_window[0]) > COMPARE_THRESHOLD) return 0;
	if (std::fabs(a->p_detector_window[1] - b->p_detector_window[1]) > COMPARE_THRESHOLD) return 0;
	if (std::fabs((a->p_detector_window[2] - a->d_sample_source) - (b->p_detector_window[2] - b->d_sample_source)) > COMPARE_THRESHOLD) return 0;
	double na[3], nb[3], la = 0.0, lb = 0.0;
	for (int i = 0; i < 3; i++) { la += a->n_detector_orientation[i] * a->n_detector_orientation[i]; lb += b->n_detector_orientation[i] * b->n_detector_orientation[i]; }
	for (int i = 0; i < 3; i++) { na[i] = a->n_detector_orientation[i] / std::sqrt(la); nb[i] = b->n_detector_orientation[i] / std::sqrt(lb); }
	for (int i = 0; i < 3; i++) if (std::fabs(na[i] - nb[i]) > COMPARE_THRESHOLD) return 0;
	if (std::fabs(a->area_detector - b->area_detector) / a->area_detector > COMPARE_THRESHOLD) return 0;
	if (std::fabs(a->collimator_height - b->collimator_height) > COMPARE_THRESHOLD) return 0;
	if (std::fabs(a->collimator_diameter - b->collimator_diameter) > COMPARE_THRESHOLD) return 0;
	return 1;
}

// xmi_check_escape_ratios_match (src/xmi_detector.c:143-172): same crystal
extern "C" int xmb_check_escape_ratios_match(const xmb_input *A, const xmb_input *B) {
	const xmb_detector *a = A->detector, *b = B->detector;
	if (a->n_crystal_layers != b->n_crystal_layers) return 0;
	for (int i = 0; i < a->n_crystal_layers; i++) {
		const xmb_layer &x = a->crystal_layers[i], &y = b->crystal_layers[i];
		if (std::fabs(x.thickness - y.thickness) / x.thickness > COMPARE_THRESHOLD) return 0;
		if (std::fabs(x.density - y.density) / x.density > COMPARE_THRESHOLD) return 0;
		if (x.n_elements != y.n_elements) return 0;
		for (int j = 0; j < x.n_elements; j++) {
			if (x.Z[j] != y.Z[j]) return 0;
			if (std::fabs(x.weight[j] - y.weight[j]) > COMPARE_THRESHOLD) return 0;
		}
	}
	return 1;
}

// xmi_find_solid_angle_match (src/xmi_solid_angle.c:672-790): 1 = file read (*rv NULL when nothing matches), 0 = error.
// A missing file counts as an empty cache.
extern "C" int xmb_find_solid_angle_match(const char *file, const xmb_input *A, const xmb_xrl_provider *xrl, xmb_solid_angle **rv,
                                          const xmb_main_options *options) {
	if (!file || !A || !rv) { xmb_set_error("xmb_find_solid_angle_match: bad arguments"); return 0; }
	*rv = nullptr;
	FILE *probe = fopen(file, "rb");
	if (!probe) return 1;
	fclose(probe);
	FILE *f = open_cache(file, 1, false, "rb");
	if (!f) return 0;
	Reader r{f};
	for (;;) {
		const unsigned long long bytes = r.get<unsigned long long>();
		if (!r.ok) break;                                   // end of file
		const long start = ftell(f);
		const unsigned xml_len = r.get<unsigned>();
		std::string xml(xml_len, '\0');
		r.arr(&xml[0], xml_len);
		xmb_input *cached = nullptr;
		if (!r.ok || !xmb_input_read_from_xml_string(xml.c_str(), &cached)) { fclose(f); xmb_set_error("%s: corrupt entry", file); return 0; }
		const bool same_provider = read_tag_matches(r);
		if (!r.ok) { xmb_input_free(&cached); fclose(f); xmb_set_error("%s: corrupt entry", file); return 0; }
		const int match = same_provider && xmb_check_solid_angle_match(cached, A, xrl);
		xmb_input_free(&cached);
		if (options && options->extra_verbose) printf(match ? "Match in solid angle grid\n" : "No match in solid angle grid\n");
		if (!match) { fseek(f, start + (long)bytes, SEEK_SET); continue; }
		xmb_solid_angle *sa = (xmb_solid_angle *)calloc(1, sizeof(xmb_solid_angle));
		sa->grid_dims_theta_n = (long)r.get<long long>();
		sa->grid_dims_r_n = (long)r.get<long long>();
		const size_t n = (size_t)sa->grid_dims_theta_n * sa->grid_dims_r_n;
		sa->solid_angles = (double *)malloc(sizeof(double) * n);
		sa->grid_dims_r_vals = (double *)malloc(sizeof(double) * sa->grid_dims_r_n);
		sa->grid_dims_theta_vals = (double *)malloc(sizeof(double) * sa->grid_dims_theta_n);
		r.arr(sa->solid_angles, n); r.arr(sa->grid_dims_r_vals, sa->grid_dims_r_n); r.arr(sa->grid_dims_theta_vals, sa->grid_dims_theta_n);
		sa->xmi_input_string = strdup(xml.c_str());
		fclose(f);
		if (!r.ok) { xmb_free_solid_angle(sa); xmb_set_error("%s: truncated entry", file); return 0; }
		*rv = sa;
		return 1;
	}
	fclose(f);
	return 1;
}

// xmi_update_solid_angle_hdf5_file (src/xmi_solid_angle.c:193-300): appends one entry; creates the file when missing
extern "C" int xmb_update_solid_angle_cache_file(const char *file, const xmb_solid_angle *sa) {
	if (!file || !sa || !sa->xmi_input_string || !sa->solid_angles) { xmb_set_error("xmb_update_solid_angle_cache_file: bad arguments (the grid needs its xmi_input_string)"); return 0; }
	FILE *f = open_cache(file, 1, true, "rb+");
	if (!f) return 0;
	fseek(f, 0, SEEK_END);
	const unsigned xml_len = (unsigned)strlen(sa->xmi_input_string);
	const long long nt = sa->grid_dims_theta_n, nr = sa->grid_dims_r_n;
	const unsigned long long bytes = 4 + xml_len + 4 + g_cache_tag.size() + 16 + sizeof(double) * ((size_t)nt * nr + nt + nr);
	fwrite(&bytes, 8, 1, f); fwrite(&xml_len, 4, 1, f); fwrite(sa->xmi_input_string, 1, xml_len, f);
	write_tag(f);
	fwrite(&nt, 8, 1, f); fwrite(&nr, 8, 1, f);
	fwrite(sa->solid_angles, sizeof(double), (size_t)nt * nr, f);
	fwrite(sa->grid_dims_r_vals, sizeof(double), nr, f);
	const size_t w = fwrite(sa->grid_dims_theta_vals, sizeof(double), nt, f);
	fclose(f);
	if (w != (size_t)nt) { xmb_set_error("short write to %s", file); return 0; }
	return 1;
}

// xmi_find_escape_ratios_match (src/xmi_detector.c:300-435)
extern "C" int xmb_find_escape_ratios_match(const char *file, const xmb_input *A, xmb_escape_ratios **rv, const xmb_main_options *options) {
	if (!file || !A || !rv) { xmb_set_error("xmb_find_escape_ratios_match: bad arguments"); return 0; }
	*rv = nullptr;
	FILE *probe = fopen(file, "rb");
	if (!probe) return 1;
	fclose(probe);
	FILE *f = open_cache(file, 2, false, "rb");
	if (!f) return 0;
	Reader r{f};
	for (;;) {
		const unsigned long long bytes = r.get<unsigned long long>();
		if (!r.ok) break;
		const long start = ftell(f);
		const unsigned xml_len = r.get<unsigned>();
		std::string xml(xml_len, '\0');
		r.arr(&xml[0], xml_len);
		xmb_input *cached = nullptr;
		if (!r.ok || !xmb_input_read_from_xml_string(xml.c_str(), &cached)) { fclose(f); xmb_set_error("%s: corrupt entry", file); return 0; }
		const bool same_provider = read_tag_matches(r);
		if (!r.ok) { xmb_input_free(&cached); fclose(f); xmb_set_error("%s: corrupt entry", file); return 0; }
		const int match = same_provider && xmb_check_escape_ratios_match(cached, A);
		xmb_input_free(&cached);
		if (options && options->extra_verbose) printf(match ? "Match in escape ratios\n" : "No match in escape ratios\n");
		if (!match) { fseek(f, start + (long)bytes, SEEK_SET); continue; }
		xmb_escape_ratios *e = (xmb_escape_ratios *)calloc(1, sizeof(xmb_escape_ratios));
		e->n_elements = r.get<int>(); e->n_fluo_input_energies = r.get<int>(); e->n_compton_input_energies = r.get<int>(); e->n_compton_output_energies = r.get<int>();
		const size_t nf = (size_t)e->n_fluo_input_energies * 109 * e->n_elements, nc = (size_t)e->n_compton_input_energies * e->n_compton_output_energies;
		e->Z = (int *)malloc(sizeof(int) * e->n_elements);
		e->fluo_escape_ratios = (double *)malloc(sizeof(double) * nf);
		e->fluo_escape_input_energies = (double *)malloc(sizeof(double) * e->n_fluo_input_energies);
		e->compton_escape_ratios = (double *)malloc(sizeof(double) * nc);
		e->compton_escape_output_energies = (double *)malloc(sizeof(double) * e->n_compton_output_energies);
		r.arr(e->Z, e->n_elements); r.arr(e->fluo_escape_ratios, nf); r.arr(e->fluo_escape_input_energies, e->n_fluo_input_energies);
		r.arr(e->compton_escape_ratios, nc);
		std::vector<double> cin(e->n_compton_input_energies);
		r.arr(cin.data(), cin.size());
		e->compton_escape_input_energies = e->fluo_escape_input_energies;     // one array in the reference too (src/xmi_main.F90:5525)
		r.arr(e->compton_escape_output_energies, e->n_compton_output_energies);
		e->xmi_input_string = strdup(xml.c_str());
		fclose(f);
		if (!r.ok) { xmb_set_error("%s: truncated entry", file); return 0; }
		*rv = e;
		return 1;
	}
	fclose(f);
	return 1;
}

// xmi_update_escape_ratios_hdf5_file (src/xmi_detector.c:174-298)
extern "C" int xmb_update_escape_ratios_cache_file(const char *file, const xmb_escape_ratios *e) {
	if (!file || !e || !e->xmi_input_string) { xmb_set_error("xmb_update_escape_ratios_cache_file: bad arguments (the ratios need their xmi_input_string)"); return 0; }
	FILE *f = open_cache(file, 2, true, "rb+");
	if (!f) return 0;
	fseek(f, 0, SEEK_END);
	const unsigned xml_len = (unsigned)strlen(e->xmi_input_string);
	const size_t nf = (size_t)e->n_fluo_input_energies * 109 * e->n_elements, nc = (size_t)e->n_compton_input_energies * e->n_compton_output_energies;
	const unsigned long long bytes = 4 + xml_len + 4 + g_cache_tag.size() + 16 + 4ULL * e->n_elements +
	                                 sizeof(double) * (nf + e->n_fluo_input_energies + nc + e->n_compton_input_energies + e->n_compton_output_energies);
	fwrite(&bytes, 8, 1, f); fwrite(&xml_len, 4, 1, f); fwrite(e->xmi_input_string, 1, xml_len, f);
	write_tag(f);
	fwrite(&e->n_elements, 4, 1, f); fwrite(&e->n_fluo_input_energies, 4, 1, f); fwrite(&e->n_compton_input_energies, 4, 1, f); fwrite(&e->n_compton_output_energies, 4, 1, f);
	fwrite(e->Z, 4, e->n_elements, f);
	fwrite(e->fluo_escape_ratios, sizeof(double), nf, f);
	fwrite(e->fluo_escape_input_energies, sizeof(double), e->n_fluo_input_energies, f);
	fwrite(e->compton_escape_ratios, sizeof(double), nc, f);
	fwrite(e->compton_escape_input_energies, sizeof(double), e->n_compton_input_energies, f);
	const size_t w = fwrite(e->compton_escape_output_energies, sizeof(double), e->n_compton_output_energies, f);
	fclose(f);
	if (w != (size_t)e->n_compton_output_energies) { xmb_set_error("short write to %s", file); return 0; }
	return 1;
}
