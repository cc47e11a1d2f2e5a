/* plugin_harness.c -- calls the reference-named entry points exactly as the reference's host does, from C, without
 * Python or ctypes in between:
 *   xmi_solid_angle_calculation_cl     resolved with dlsym from the plugin file, as g_module_symbol does
 *                                      (src/xmi_solid_angle.c:121-160, called at bin/xmimsim.c:320 through xmi_solid_angle_calculation)
 *   xmi_main_msim                      include/xmi_main.h:29, called at bin/xmimsim.c:361
 *   xmi_detector_convolute_all_custom  resolved with dlsym, called at bin/xmimsim.c:501-526
 * The handle it passes is an OPAQUE object of its own (standing for the reference's Fortran xmi_inputFPtr): the shim must
 * turn it into an input through the host's xmi_input_F2C, which this executable exports (-rdynamic) the way libxmimsim
 * does.  Every result is compared bit for bit with the xmb_* entry points called on the same input.
 * usage: plugin_harness <libxmimsim_b200.so> <libxmimsim-b200-interpose.so> <file.xmsi> [photons per line] */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "xmimsim_b200.h"

struct fake_fortran_input { char pad[24]; long tag; xmb_input *c_tree; };   /* nothing like an xmb handle */
static int f2c_calls = 0;

/* the host's converter (src/xmi_aux_f.F90:766-776): the shim finds it with dlsym(RTLD_DEFAULT, ...) */
void xmi_input_F2C(void *inputFPtr, xmb_input **out) {
	struct fake_fortran_input *f = (struct fake_fortran_input *)inputFPtr;
	f2c_calls++;
	*out = f->tag == 0x46303346L ? f->c_tree : NULL;
}

#define SYM(lib, type, name) type name = (type)dlsym(lib, #name); if (!name) { fprintf(stderr, "missing symbol %s: %s\n", #name, dlerror()); return 2; }
#define CHECK(cond, what) do { if (!(cond)) { fprintf(stderr, "FAIL: %s\n", what); return 1; } printf("ok: %s\n", what); } while (0)

typedef int (*sa_cl_t)(void *, xmb_solid_angle **, char *, xmb_main_options *);
typedef int (*main_msim_t)(void *, void *, int, double **, xmb_main_options *, double **, double **, xmb_solid_angle *);
typedef void (*conv_all_t)(void *, double **, double **, double *, double *, xmb_main_options *, xmb_escape_ratios *, int, int);

int main(int argc, char **argv) {
	if (argc < 4) { fprintf(stderr, "usage: %s libxmimsim_b200.so libxmimsim-b200-interpose.so file.xmsi [photons per line]\n", argv[0]); return 2; }
	void *lib = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
	if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
	void *inter = dlopen(argv[2], RTLD_NOW | RTLD_LOCAL);
	if (!inter) { fprintf(stderr, "%s\n", dlerror()); return 2; }
	SYM(lib, sa_cl_t, xmi_solid_angle_calculation_cl)
	SYM(lib, conv_all_t, xmi_detector_convolute_all_custom)
	SYM(inter, main_msim_t, xmi_main_msim)

	xmb_input *input = NULL;
	if (!xmb_input_read_from_xml_file(argv[3], &input)) { fprintf(stderr, "%s\n", xmb_last_error()); return 2; }
	if (argc > 4) input->general->n_photons_line = atol(argv[4]);
	const int n_int = input->general->n_interactions_trajectory, nch = input->detector->nchannels;
	struct fake_fortran_input F;
	memset(&F, 0x5A, sizeof(F));
	F.tag = 0x46303346L; F.c_tree = input;
	xmb_plugin_set_provider(xmb_xrl_surrogate());          /* the test box has no xraylib: say so explicitly */
	xmb_main_options opt;
	xmb_main_options_defaults(&opt);
	opt.use_escape_peaks = 0;                              /* no escape ratios in this harness */

	/* ---- the same steps through the library's own entry points --------------------------------------------- */
	xmb_inputFPtr inputF = NULL;
	xmb_hdf5FPtr tables = NULL;
	if (!xmb_input_C2F(input, &inputF) || !xmb_init_input(&inputF) || !xmb_init_from_provider(xmb_xrl_surrogate(), inputF, 1, &tables)) {
		fprintf(stderr, "%s\n", xmb_last_error()); return 2;
	}
	xmb_solid_angle *sa_ref = NULL;
	if (!xmb_solid_angle_calculation(inputF, tables, &sa_ref, NULL, &opt, xmb_get_hits_per_single(), 0)) { fprintf(stderr, "%s\n", xmb_last_error()); return 2; }
	double *ch_ref = NULL, *br_ref = NULL, *vr_ref = NULL;
	if (!xmb_main_msim(inputF, tables, 1, &ch_ref, &opt, &br_ref, &vr_ref, sa_ref)) { fprintf(stderr, "%s\n", xmb_last_error()); return 2; }

	/* ---- the reference-named symbols with the opaque handle --------------------------------------------------- */
	xmb_solid_angle *sa = NULL;
	char *xml = strdup("<harness/>");
	CHECK(xmi_solid_angle_calculation_cl(&F, &sa, xml, &opt) == 1 && sa, "xmi_solid_angle_calculation_cl returned a grid");
	CHECK(f2c_calls >= 1, "the shim converted the opaque handle with the host's xmi_input_F2C");
	CHECK(sa->xmi_input_string == xml, "the grid keeps the caller's input string (src/xmi_solid_angle_cl.c:433)");
	CHECK(sa->grid_dims_r_n == sa_ref->grid_dims_r_n && sa->grid_dims_theta_n == sa_ref->grid_dims_theta_n, "grid dimensions");
	CHECK(!memcmp(sa->grid_dims_r_vals, sa_ref->grid_dims_r_vals, sizeof(double) * sa->grid_dims_r_n) &&
	      !memcmp(sa->grid_dims_theta_vals, sa_ref->grid_dims_theta_vals, sizeof(double) * sa->grid_dims_theta_n), "grid axes bit-identical");
	CHECK(!memcmp(sa->solid_angles, sa_ref->solid_angles, sizeof(double) * sa->grid_dims_r_n * sa->grid_dims_theta_n), "solid angles bit-identical");

	double *ch = NULL, *br = NULL, *vr = NULL;
	CHECK(xmi_main_msim(&F, NULL /* a Fortran hdf5 handle the shim cannot read: tables are rebuilt */, 1, &ch, &opt, &br, &vr, sa) == 1,
	      "xmi_main_msim returned 1");
	CHECK(ch && vr && !memcmp(ch, ch_ref, sizeof(double) * (n_int + 1) * nch), "channels bit-identical to xmb_main_msim");
	CHECK(!memcmp(vr, vr_ref, sizeof(double) * 100 * 385 * n_int), "var_red_history bit-identical");
	double tot = 0.0;
	for (int i = 0; i < nch; i++) tot += ch[(size_t)n_int * nch + i];
	CHECK(tot > 0.0, "the spectrum is not empty");

	double **rows = (double **)calloc(n_int + 1, sizeof(double *)), **conv = (double **)calloc(n_int + 1, sizeof(double *));
	double **rows_ref = (double **)calloc(n_int + 1, sizeof(double *)), **conv_ref = (double **)calloc(n_int + 1, sizeof(double *));
	for (int i = 0; i <= n_int; i++) { rows[i] = ch + (size_t)i * nch; rows_ref[i] = ch_ref + (size_t)i * nch; }
	xmi_detector_convolute_all_custom(&F, rows, conv, br, vr, &opt, NULL, n_int, 0);
	xmb_detector_convolute_all(inputF, tables, rows_ref, conv_ref, br_ref, vr_ref, &opt, NULL, n_int, 0);
	int same = 1;
	for (int i = 1; i <= n_int; i++) same = same && conv[i] && conv_ref[i] && !memcmp(conv[i], conv_ref[i], sizeof(double) * nch);
	CHECK(same, "xmi_detector_convolute_all_custom bit-identical to xmb_detector_convolute_all");
	CHECK(!memcmp(ch, ch_ref, sizeof(double) * (n_int + 1) * nch), "rows corrected in place identically (src/xmi_detector_f.F90:412-413)");
	printf("harness OK (%d xmi_input_F2C calls)\n", f2c_calls);
	return 0;
}
