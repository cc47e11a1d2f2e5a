// engine.h -- internal structures shared by the host-side C++ and the CUDA translation units.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "xmimsim_b200.h"

#define XMB_MAGIC_INPUT 0x584D42494E505554ULL  // "XMBINPUT"
#define XMB_MAGIC_HDF5  0x584D425441424C45ULL  // "XMBTABLE"
#define XMB_DEFAULT_SEED 0x584D494D53494DULL   // "XMIMSIM"

struct XmbInputF {
	uint64_t magic = XMB_MAGIC_INPUT;
	xmb_input in{};            // deep copy, owned
	bool inited = false;
	xmb_derived der{};
	std::vector<double> thickness_along_Z, Z_coord_begin, Z_coord_end;
};

struct XmbDeviceTables;        // device.cu

struct XmbHdf5F {
	uint64_t magic = XMB_MAGIC_HDF5;
	const xmb_xrl_provider *xrl = nullptr;
	xmb_tables_host view{};
	// storage behind the view
	std::vector<int> Z, uniqZ, bucket_start;
	std::vector<double> atomic_weight, node_E, cs_total, cs_photo_total, p_rayl, p_rayl_compt,
	    cs_photo_partial, cs_vacancy, icdf_E, icdf_R, rayl_theta_icdf, compt_theta_icdf, phi_T,
	    phi_icdf, cp_R, cp_icdf, ff, sf, fluor_yield, fluor_yield_corr, cos_kron, rad_rate,
	    line_energy, edge_energy, mu_layer, exc_murhod, auger_rate;
	std::vector<int> adv_off, adv_shell;
	std::vector<double> adv_config, adv_edge, adv_cdf, adv_qinv;
	int quality = 0;
	double e_max = 0.0;
	// lazily built device-side layouts (history.cu): one handle per CUDA device the simulation has run on (a single
	// process may drive every GPU of the box, multi_gpu.cu); `dev` is the handle of the last run (accessors read it)
	XmbDeviceTables *dev = nullptr;
	std::vector<XmbDeviceTables *> devs;
};

void xmb_set_error(const char *fmt, ...);
XmbInputF *xmb_as_input(xmb_inputFPtr p);
XmbHdf5F *xmb_as_hdf5(xmb_hdf5FPtr p);

// host-side table evaluation (used for the solid-angle bounds and the exciter absorbers)
double xmb_host_mu_layer(const xmb_xrl_provider *xrl, const xmb_layer *layer, double E);
void xmb_free_device_tables(XmbDeviceTables *dev);
void xmb_free_all_device_tables(XmbHdf5F *h);   // every per-device handle of h
int xmb_build_tables(const xmb_xrl_provider *xrl, xmb_inputFPtr inputF, int quality, xmb_hdf5FPtr *out, bool skip_icdf);
