// plugin_shim.cpp -- the reference's plugin/ABI symbol names, forwarding to the xmb_* engine.
//
// Exported under the exact names and signatures the reference resolves with g_module_symbol / links:
//   xmi_solid_angle_calculation_cl     typedef XmiSolidAngleCalculation, src/xmi_solid_angle.c:51; symbol looked up at :139
//   xmi_detector_convolute_all_custom  typedef XmiDetectorConvoluteAll, include/xmi_main.h:37; looked up in bin/xmimsim.c:513
//   xmi_main_msim                      include/xmi_main.h:29, src/xmi_main.F90:66-77.  Compiled with -DXMB_EXPORT_XMI_MAIN_MSIM into a second file,
//                                      libxmimsim-b200-interpose.so (Makefile): in the plugin file itself the name would shadow, or be shadowed by,
//                                      libxmimsim's own; LD_PRELOAD the interpose file to route the host's xmi_main_msim calls to the GPU
//
// The reference passes an opaque xmi_inputFPtr (a Fortran derived type, not readable from C).  The shim accepts
// either (a) one of this library's own handles (magic-tagged), or (b) a reference handle, which it converts
// with the reference's own xmi_input_F2C (src/xmi_aux_f.F90:766-776; resolved lazily from the host process, as
// custom-detector-response/detector-response2.c:39 does) and re-derives xmi_init_input's fields itself.
// Cross sections come from the provider registered with xmb_plugin_set_provider; without one the shim binds xraylib at run
// time (xmb_xrl_from_library: the host process of the reference has libxrl loaded already).  The analytic stand-in is used
// only when XMB_ALLOW_SURROGATE=1 is set, with a warning: a grid computed from stand-in attenuation bounds would otherwise
// end up in the host's permanent HDF5 cache.
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include "engine.h"

#ifndef XMB_INTERPOSE_ONLY   // the interpose file holds xmi_main_msim alone and takes everything else from the library
static const xmb_xrl_provider *g_provider = nullptr;
static long g_hits_per_single = 0;   // 0: not set through xmb_set_hits_per_single

extern "C" void xmb_set_hits_per_single(long n) { g_hits_per_single = n > 0 ? n : 0; }
extern "C" long xmb_get_hits_per_single(void) {
	if (g_hits_per_single > 0) return g_hits_per_single;
	// the reference reads its global `hits_per_single` (src/xmi_solid_angle_cl.c:54); honour it when the host exports it
	long *hps = (long *)dlsym(RTLD_DEFAULT, "hits_per_single");
	return hps && *hps > 0 ? *hps : 5000;
}

extern "C" void xmb_plugin_set_provider(const xmb_xrl_provider *p) { g_provider = p; }

// registered provider, else xraylib, else (only on request) the stand-in; NULL + xmb_last_error otherwise
extern "C" const xmb_xrl_provider *xmb_plugin_provider(void) {
	if (g_provider) return g_provider;
	static const xmb_xrl_provider *cached = nullptr;
	if (cached) return cached;
	const xmb_xrl_provider *p = xmb_xrl_from_library(nullptr);
	if (!p) {
		const char *allow = getenv("XMB_ALLOW_SURROGATE");
		if (!allow || allow[0] != '1') {
			xmb_set_error("no cross-section provider: xraylib could not be bound and none was registered with xmb_plugin_set_provider "
			              "(XMB_ALLOW_SURROGATE=1 accepts the analytic stand-in, NOT physics-grade)");
			return nullptr;
		}
		p = xmb_xrl_surrogate();
		fprintf(stderr, "xmimsim-b200: WARNING: cross sections from the analytic stand-in (%s), NOT physics-grade\n", p->name);
	}
	return cached = p;
}

typedef void (*xmi_input_F2C_t)(void *, xmb_input **);

// Resolve the caller's handle to one of ours.  *owned is set when a temporary handle was created.
extern "C" int xmb_plugin_resolve_input(void *inputFPtr, xmb_inputFPtr *out, int *owned) {
	*owned = 0;
	XmbInputF *h = static_cast<XmbInputF *>(inputFPtr);
	// our own handles start with the magic word; reading 8 bytes of a foreign object is safe (it is at least a struct)
	if (h && h->magic == XMB_MAGIC_INPUT) { *out = h; return 1; }
	xmi_input_F2C_t f2c = (xmi_input_F2C_t)dlsym(RTLD_DEFAULT, "xmi_input_F2C");
	if (!f2c) { xmb_set_error("handle is not an xmb handle and xmi_input_F2C is not available in this process"); return 0; }
	xmb_input *c_tree = nullptr;
	f2c(inputFPtr, &c_tree);
	if (!c_tree) { xmb_set_error("xmi_input_F2C returned NULL"); return 0; }
	xmb_inputFPtr mine = nullptr;
	if (!xmb_input_C2F(c_tree, &mine) || !xmb_init_input(&mine)) return 0;
	*out = mine;
	*owned = 1;
	return 1;
}

extern "C" int xmi_solid_angle_calculation_cl(void *inputFPtr, xmb_solid_angle **solid_angle, char *input_string, xmb_main_options *options) {
	xmb_inputFPtr in = nullptr;
	int owned = 0;
	if (!xmb_plugin_resolve_input(inputFPtr, &in, &owned)) { fprintf(stderr, "xmimsim-b200: %s\n", xmb_last_error()); return 0; }
	xmb_hdf5FPtr tables = nullptr;
	int rv = 0;
	const xmb_xrl_provider *xrl = xmb_plugin_provider();
	if (xrl && xmb_init_from_provider(xrl, in, 1, &tables)) {
		rv = xmb_solid_angle_calculation(in, tables, solid_angle, input_string, options, xmb_get_hits_per_single(), 0);
		xmb_free_hdf5_F(&tables);
	}
	if (!rv) fprintf(stderr, "xmimsim-b200: %s\n", xmb_last_error());   // 0 = caller falls through to its next backend
	if (owned) xmb_free_input_F(&in);
	return rv;
}

extern "C" void xmi_detector_convolute_all_custom(void *inputFPtr, double **channels_noconv, double **channels_conv, double *brute_history,
                                                  double *var_red_history, xmb_main_options *options, xmb_escape_ratios *escape_ratios,
                                                  int n_interactions_all, int zero_interaction) {
	xmb_inputFPtr in = nullptr;
	int owned = 0;
	if (!xmb_plugin_resolve_input(inputFPtr, &in, &owned)) { fprintf(stderr, "xmimsim-b200: %s\n", xmb_last_error()); return; }
	if (options && options->verbose) printf("xmimsim-b200 detector response (CUDA sm_100a)\n");
	// provider for the absorber / crystal attenuation.  The hook has no error return (include/xmi_main.h:37): without a
	// provider it stops the run, as the reference does on its own fatal errors
	const xmb_xrl_provider *xrl = xmb_plugin_provider();
	if (!xrl) { fprintf(stderr, "xmimsim-b200: %s\n", xmb_last_error()); exit(1); }
	XmbHdf5F shell;   // carries only the provider pointer
	shell.xrl = xrl;
	xmb_hdf5FPtr tables = &shell;
	xmb_detector_convolute_all(in, tables, channels_noconv, channels_conv, brute_history, var_red_history, options, escape_ratios,
	                           n_interactions_all, zero_interaction);
	if (owned) xmb_free_input_F(&in);
}

#endif   // XMB_INTERPOSE_ONLY

#ifdef XMB_EXPORT_XMI_MAIN_MSIM
extern "C" int xmi_main_msim(void *inputFPtr, void *hdf5FPtr, int n_mpi_hosts, double **channels, xmb_main_options *options,
                             double **brute_history, double **var_red_history, xmb_solid_angle *solid_angles) {
	xmb_inputFPtr in = nullptr;
	int owned = 0;
	if (!xmb_plugin_resolve_input(inputFPtr, &in, &owned)) { fprintf(stderr, "xmimsim-b200: %s\n", xmb_last_error()); return 0; }
	XmbHdf5F *h = static_cast<XmbHdf5F *>(hdf5FPtr);
	xmb_hdf5FPtr tables = (h && h->magic == XMB_MAGIC_HDF5) ? hdf5FPtr : nullptr;
	bool own_tables = false;
	if (!tables) {
		const xmb_xrl_provider *xrl = xmb_plugin_provider();
		if (!xrl || !xmb_init_from_provider(xrl, in, 1, &tables)) { fprintf(stderr, "xmimsim-b200: %s\n", xmb_last_error()); if (owned) xmb_free_input_F(&in); return 0; }
		own_tables = true;
	}
	const int rv = xmb_main_msim(in, tables, n_mpi_hosts, channels, options, brute_history, var_red_history, solid_angles);
	if (own_tables) xmb_free_hdf5_F(&tables);
	if (owned) xmb_free_input_F(&in);
	return rv;
}
#endif
