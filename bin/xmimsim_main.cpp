// xmimsim-b200 -- command-line driver with the reference's options and progress lines (bin/xmimsim.c:78-700):
//   xmimsim-b200 [options] inputfile.xmsi
// reads the XMSI file, computes the solid-angle grid, runs the histories on the GPU, computes the escape ratios of
// the detector crystal when escape peaks are on, applies the detector response and writes the XMSO file named in
// the input (plus optional SPE / CSV spectra).  The solid-angle grid and the escape ratios are recomputed on the GPU
// (54 ms and 0.3 s on a B200) unless cache files are named (--with-solid-angles-data, --with-escape-ratios-data: the
// reference's HDF5 caches as a side-car container with the same schema and match rules, host_cache.cpp).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>
#include "xmimsim_b200.h"

static void usage(FILE *f) {
	fputs("Usage:\n  xmimsim-b200 [OPTION...] inputfile\n\n"
	      "xmimsim-b200: Monte-Carlo simulation of X-ray fluorescence spectra on an NVIDIA B200\n\n"
	      "  --enable-M-lines / --disable-M-lines                      M lines (default: enabled)\n"
	      "  --enable-auger-cascade / --disable-auger-cascade          non-radiative cascade (default: enabled)\n"
	      "  --enable-radiative-cascade / --disable-radiative-cascade  radiative cascade (default: enabled)\n"
	      "  --enable-variance-reduction / --disable-variance-reduction forced detection (default: enabled)\n"
	      "  --enable-pile-up / --disable-pile-up                      pulse pile-up (default: disabled)\n"
	      "  --enable-escape-peaks / --disable-escape-peaks            escape peaks (default: enabled)\n"
	      "  --enable-poisson / --disable-poisson                      Poisson noise (default: disabled)\n"
	      "  --enable-advanced-compton / --disable-advanced-compton    shell-resolved Compton (default: disabled)\n"
	      "  --spe-file=F --spe-file-unconvoluted=F                    write F_<order>.spe\n"
	      "  --csv-file=F --csv-file-unconvoluted=F                    write CSV spectra\n"
	      "  --custom-detector-response=LIB                            use xmi_detector_convolute_all_custom from LIB (bin/xmimsim.c:505-522)\n"
	      "  --with-solid-angles-data=F --with-escape-ratios-data=F    cache files (queried first, updated after a calculation)\n"
	      "  --set-seed=N                                              Philox key (default: library seed)\n"
	      "  --with-xraylib[=LIB]                                      libxrl to take the cross sections from (default: libxrl.so.11 / .7 / libxrl.so)\n"
	      "  --surrogate-cross-sections                                run with the built-in analytic stand-in instead of xraylib: NOT physics-grade,\n"
	      "                                                            for tests and benchmarks on machines without xraylib\n"
	      "  --gpus=N                                                  shard the histories over N GPUs of this machine (one NCCL all-reduce; 0 = all)\n"
	      "  --table-quality=0|1                                       inverse-CDF integration resolution (default 1 = reference)\n"
	      "  -v, --verbose    -V, --very-verbose    --version\n", f);
}

int main(int argc, char **argv) {
	xmb_main_options opt;
	xmb_main_options_defaults(&opt);
	std::string spe_conv, spe_noconv, csv_conv, csv_noconv, infile, sa_cache, er_cache, custom_response, xraylib_path;
	bool use_surrogate = false;
	int n_gpus = 1;
	unsigned long long seed = 0;
	int quality = 1;
	struct Flag { const char *name; int *target; };
	const Flag flags[] = {{"M-lines", &opt.use_M_lines}, {"auger-cascade", &opt.use_cascade_auger}, {"radiative-cascade", &opt.use_cascade_radiative},
	                      {"variance-reduction", &opt.use_variance_reduction}, {"pile-up", &opt.use_sum_peaks}, {"escape-peaks", &opt.use_escape_peaks},
	                      {"poisson", &opt.use_poisson}, {"advanced-compton", &opt.use_advanced_compton}, {"gpu", &opt.use_gpu},
	                      {"default-seeds", &opt.use_default_seeds}};
	for (int i = 1; i < argc; i++) {
		const std::string a = argv[i];
		bool done = false;
		for (const Flag &f : flags) {
			if (a == std::string("--enable-") + f.name) { *f.target = 1; done = true; }
			else if (a == std::string("--disable-") + f.name) { *f.target = 0; done = true; }
		}
		if (done) continue;
		auto val = [&](const char *key, std::string &out) {
			const std::string k = std::string(key) + "=";
			if (a.compare(0, k.size(), k) == 0) { out = a.substr(k.size()); return true; }
			if (a == key && i + 1 < argc) { out = argv[++i]; return true; }
			return false;
		};
		std::string tmp;
		if (val("--spe-file-unconvoluted", spe_noconv) || val("--spe-file", spe_conv) || val("--csv-file-unconvoluted", csv_noconv) || val("--csv-file", csv_conv)) continue;
		if (val("--custom-detector-response", custom_response)) continue;
		if (val("--with-solid-angles-data", sa_cache) || val("--with-escape-ratios-data", er_cache)) continue;
		if (a == "--with-xraylib") continue;
		if (a.compare(0, 15, "--with-xraylib=") == 0) { xraylib_path = a.substr(15); continue; }
		if (a == "--surrogate-cross-sections") { use_surrogate = true; continue; }
		if (val("--gpus", tmp)) { n_gpus = atoi(tmp.c_str()); continue; }
		if (val("--set-seed", tmp)) { seed = strtoull(tmp.c_str(), nullptr, 0); continue; }
		if (val("--table-quality", tmp)) { quality = atoi(tmp.c_str()); continue; }
		if (val("--set-threads", tmp)) { opt.omp_num_threads = atoi(tmp.c_str()); continue; }
		if (a == "-v" || a == "--verbose") { opt.verbose = 1; continue; }
		if (a == "-V" || a == "--very-verbose") { opt.verbose = 1; opt.extra_verbose = 1; continue; }
		if (a == "--version") { printf("%s\n", xmb_version()); return 0; }
		if (a == "-h" || a == "--help") { usage(stdout); return 0; }
		if (!a.empty() && a[0] == '-') { fprintf(stderr, "Unknown option %s\n", a.c_str()); usage(stderr); return 1; }
		infile = a;
	}
	if (infile.empty()) { usage(stderr); return 1; }
	if (xmb_cuda_device_count() < 1) { fprintf(stderr, "No CUDA device found: xmimsim-b200 has no CPU fallback\n"); return 1; }

	// cross sections: xraylib, as the reference links it (configure.ac:115-116).  The analytic stand-in only on request.
	const xmb_xrl_provider *xrl = nullptr;
	if (use_surrogate) {
		xrl = xmb_xrl_surrogate();
		fprintf(stderr, "WARNING: cross sections from the built-in analytic stand-in (%s): the spectra are NOT physics-grade\n", xrl->name);
	} else {
		xrl = xmb_xrl_from_library(xraylib_path.empty() ? nullptr : xraylib_path.c_str());
		if (!xrl) {
			fprintf(stderr, "%s\nxraylib is required (--with-xraylib=LIB names the library); --surrogate-cross-sections runs with the "
			                "analytic stand-in instead (not physics-grade)\n", xmb_last_error());
			return 1;
		}
	}
	xmb_cache_set_provider(xrl);    // cache entries carry the provider's name: stand-in grids never serve an xraylib run
	xmb_plugin_set_provider(xrl);   // a --custom-detector-response plugin that is this library computes with the same cross sections
	xmb_input *input = nullptr;
	if (!xmb_input_read_from_xml_file(infile.c_str(), &input)) { fprintf(stderr, "Could not read %s: %s\n", infile.c_str(), xmb_last_error()); return 1; }
	if (opt.verbose) printf("Inputfile %s successfully parsed\n", infile.c_str());
	xmb_inputFPtr inputF = nullptr;
	xmb_hdf5FPtr tables = nullptr;
	if (!xmb_input_C2F(input, &inputF) || !xmb_init_input(&inputF)) { fprintf(stderr, "%s\n", xmb_last_error()); return 1; }
	if (opt.verbose) printf("Building cross-section tables (%s)\n", xrl->name);
	if (!xmb_init_from_provider(xrl, inputF, quality, &tables)) { fprintf(stderr, "Could not build the tables: %s\n", xmb_last_error()); return 1; }

	const int n_int = input->general->n_interactions_trajectory, nch = input->detector->nchannels;
	xmb_solid_angle *sa = nullptr;
	if (opt.use_variance_reduction) {
		// bin/xmimsim.c:300-333: query the cache, calculate on a miss, update the cache
		if (!sa_cache.empty()) {
			if (opt.verbose) printf("Querying %s for solid angle grid\n", sa_cache.c_str());
			if (!xmb_find_solid_angle_match(sa_cache.c_str(), input, xrl, &sa, &opt)) { fprintf(stderr, "%s\n", xmb_last_error()); return 1; }
		}
		if (!sa) {
			if (opt.verbose) printf("Precalculating solid angle grid\n");
			char *xml = nullptr;
			if (!xmb_input_write_to_xml_string(input, &xml)) { fprintf(stderr, "Could not write input to XML string: %s\n", xmb_last_error()); return 1; }
			if (!xmb_solid_angle_calculation(inputF, tables, &sa, xml, &opt, 5000, seed)) { fprintf(stderr, "Solid angle calculation failed: %s\n", xmb_last_error()); return 1; }
			if (opt.verbose) printf("Solid angle calculation finished\n");
			if (!sa_cache.empty()) {
				if (!xmb_update_solid_angle_cache_file(sa_cache.c_str(), sa)) { fprintf(stderr, "%s\n", xmb_last_error()); return 1; }
				if (opt.verbose) printf("%s was successfully updated with new solid angle grid\n", sa_cache.c_str());
			}
		} else if (opt.verbose) printf("Solid angle grid already present in %s\n", sa_cache.c_str());
	} else if (opt.verbose) printf("Operating in brute-force mode: solid angle grid is redundant\n");
	double *channels = nullptr, *brute = nullptr, *var_red = nullptr;
	if (n_gpus == 1) {
		xmb_msim_ex ex{};
		ex.n_ranks = 1; ex.device = -1; ex.seed = seed;
		uint64_t *limbs = nullptr;
		size_t n_slots = 0;
		if (opt.verbose) { printf("Simulating interactions\n"); fflush(stdout); }
		if (!xmb_main_msim_raw(inputF, tables, &opt, sa, &ex, &limbs, &n_slots) ||
		    !xmb_main_msim_finish(inputF, tables, &opt, limbs, n_slots, &channels, &brute, &var_red)) { fprintf(stderr, "Error in xmi_main_msim: %s\n", xmb_last_error()); return 1; }
		free(limbs);
		if (opt.verbose) printf("Interactions simulation finished\n");
	} else {
		// the reference's MPI run (bin/xmimsim.c:396-413), here over the GPUs of this machine
		xmb_msim_ex ex{};
		ex.seed = seed;
		if (opt.verbose) { printf("Simulating interactions on %d GPUs\n", n_gpus > 0 ? n_gpus : xmb_cuda_device_count()); fflush(stdout); }
		if (!xmb_main_msim_all_devices(inputF, tables, n_gpus, nullptr, &channels, &opt, &brute, &var_red, sa, &ex)) { fprintf(stderr, "Error in xmi_main_msim: %s\n", xmb_last_error()); return 1; }
		if (opt.verbose) printf("Interactions simulation finished (%llu histories, slowest GPU %.1f ms)\n", (unsigned long long)ex.n_histories, ex.kernel_ms);
	}
	double zero_sum = 0.0;
	for (int j = 0; j < nch; j++) zero_sum += channels[j];
	const int first = zero_sum > 0.0 ? 0 : 1;                                     // bin/xmimsim.c:436, 499

	xmb_escape_ratios *er = nullptr;
	if (opt.use_escape_peaks) {
		// bin/xmimsim.c:462-495
		if (!er_cache.empty()) {
			if (opt.verbose) printf("Querying %s for escape peak ratios\n", er_cache.c_str());
			if (!xmb_find_escape_ratios_match(er_cache.c_str(), input, &er, &opt)) { fprintf(stderr, "%s\n", xmb_last_error()); return 1; }
		}
		if (!er) {
			if (opt.verbose) printf("Precalculating escape peak ratios\n");
			char *xml = nullptr;
			if (!xmb_input_write_to_xml_string(input, &xml)) { fprintf(stderr, "Could not write input to XML string: %s\n", xmb_last_error()); return 1; }
			const int ok = xmb_escape_ratios_calculation(input, &er, xml, xrl, &opt, xmb_get_default_escape_ratios_options(), seed);
			free(xml);
			if (!ok) { fprintf(stderr, "Escape ratios calculation failed: %s\n", xmb_last_error()); return 1; }
			if (!er_cache.empty()) {
				if (!xmb_update_escape_ratios_cache_file(er_cache.c_str(), er)) { fprintf(stderr, "%s\n", xmb_last_error()); return 1; }
				if (opt.verbose) printf("%s was successfully updated with new escape peak ratios\n", er_cache.c_str());
			}
		} else if (opt.verbose) printf("Escape peak ratios already present in %s\n", er_cache.c_str());
	}
	// The response corrects its input rows in place (detector absorbers, crystal efficiency, escape peaks:
	// src/xmi_detector_f.F90:412-468) and the reference writes THOSE rows as the "unconvoluted" spectra (channelsdef after
	// xmi_detector_convolute_all, bin/xmimsim.c:498-642): so does this driver.
	std::vector<double *> rows(n_int + 1), conv(n_int + 1, nullptr);
	for (int i = 0; i <= n_int; i++) rows[i] = channels + (size_t)i * nch;
	std::vector<double *> &raw_rows = rows;
	if (!custom_response.empty()) {
		// the reference's plugin hook (bin/xmimsim.c:505-522): same symbol, same signature
		void *mod = dlopen(custom_response.c_str(), RTLD_NOW | RTLD_LOCAL);
		if (!mod) { fprintf(stderr, "Could not open %s: %s\n", custom_response.c_str(), dlerror()); return 1; }
		typedef void (*ConvoluteAll)(void *, double **, double **, double *, double *, xmb_main_options *, xmb_escape_ratios *, int, int);
		ConvoluteAll fn = (ConvoluteAll)dlsym(mod, "xmi_detector_convolute_all_custom");
		if (!fn) { fprintf(stderr, "Could not get symbol xmi_detector_convolute_all_custom from %s: %s\n", custom_response.c_str(), dlerror()); return 1; }
		if (opt.verbose) printf("xmi_detector_convolute_all_custom loaded from %s\n", custom_response.c_str());
		fn(inputF, rows.data(), conv.data(), brute, var_red, &opt, er, n_int, first == 0 ? 1 : 0);
	} else
		xmb_detector_convolute_all(inputF, tables, rows.data(), conv.data(), brute, var_red, &opt, er, n_int, first == 0 ? 1 : 0);
	for (int i = first; i <= n_int; i++) if (!conv[i]) { fprintf(stderr, "Detector response failed: %s\n", xmb_last_error()); return 1; }
	if (!conv[0]) conv[0] = (double *)calloc(nch, sizeof(double));

	for (int i = first; i <= n_int; i++) {
		char name[4096];
		if (!spe_noconv.empty()) {
			snprintf(name, sizeof(name), "%s_%d.spe", spe_noconv.c_str(), i);
			if (!xmb_write_spe_file(name, input, raw_rows[i])) { fprintf(stderr, "%s\n", xmb_last_error()); return 1; }
			if (opt.verbose) printf("Writing to SPE file %s\n", name);
		}
		if (!spe_conv.empty()) {
			snprintf(name, sizeof(name), "%s_%d.spe", spe_conv.c_str(), i);
			if (!xmb_write_spe_file(name, input, conv[i])) { fprintf(stderr, "%s\n", xmb_last_error()); return 1; }
			if (opt.verbose) printf("Writing to SPE file %s\n", name);
		}
	}
	if (!csv_noconv.empty()) {
		if (!xmb_write_csv_file(csv_noconv.c_str(), input, raw_rows.data(), first)) { fprintf(stderr, "%s\n", xmb_last_error()); return 1; }
		if (opt.verbose) printf("Writing to CSV file %s\n", csv_noconv.c_str());
	}
	if (!csv_conv.empty()) {
		if (!xmb_write_csv_file(csv_conv.c_str(), input, conv.data(), first)) { fprintf(stderr, "%s\n", xmb_last_error()); return 1; }
		if (opt.verbose) printf("Writing to CSV file %s\n", csv_conv.c_str());
	}
	if (!xmb_output_write_to_xml_file(input, infile.c_str(), input->general->outputfile, channels, conv.data(), brute,
	                                  opt.use_variance_reduction ? var_red : nullptr, first == 0 ? 1 : 0, xrl)) {
		fprintf(stderr, "Could not write to %s: %s\n", input->general->outputfile, xmb_last_error());
		return 1;
	}
	if (opt.verbose) printf("Output written to XMSO file %s\n", input->general->outputfile);
	for (int i = 0; i <= n_int; i++) free(conv[i]);
	free(channels); free(brute); free(var_red);
	if (er) xmb_free_escape_ratios(&er);
	if (sa) xmb_free_solid_angle(sa);
	xmb_free_hdf5_F(&tables);
	xmb_free_input_F(&inputF);
	xmb_input_free(&input);
	return 0;
}
