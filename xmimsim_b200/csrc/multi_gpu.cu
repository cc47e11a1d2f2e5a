// multi_gpu.cu -- photon histories sharded over the GPUs of one box, the histograms summed by ONE NCCL all-reduce.
//
// Replaces the reference's MPI split and reduction: every host simulates n_photons / n_mpi_hosts photons of every
// source line (src/xmi_main.F90:314,574) and rank 0 collects three MPI_Reduce(MPI_SUM) of the double arrays
// (bin/xmimsim.c:396-413).  Here a rank (= one GPU) runs the history kernel on its block-cyclic shard of the global
// photon ids, converts its exact 128-bit accumulators to 48-bit limbs (xmb_limbs_kernel) and the limbs of all ranks are
// added by ncclAllReduce(ncclUint64, ncclSum) ON THE SAME STREAM, without a host synchronisation in between; the
// epilogue (xmb_main_msim_finish: 128-bit integer sums -> cumulative rows -> double) then runs on identical integers on
// every rank.  Integer sums commute, every random draw has a fixed address: the result is bit-identical at 1/2/4/8 GPUs.
//
// Two ways to form the communicator:
//   * one process per GPU (torchrun / mpirun, what the reference's MPI build does): xmb_comm_unique_id on rank 0, the
//     128-byte id travels through the launcher's own channel (MPI_Bcast in the reference host, torch.distributed in
//     bench.py), xmb_comm_init_rank everywhere;
//   * one process, every visible GPU (bin/xmimsim-b200 --gpus N): xmb_main_msim_all_devices, ncclCommInitAll.
// NCCL is bound at run time (dlopen libnccl.so.2: the system library, or the one a host such as PyTorch has already
// loaded), so the library itself carries no link-time dependency; without NCCL the multi-GPU calls fail loudly.
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include <mutex>
#include <nccl.h>
#include "cuda_util.cuh"
#include "device_tables.h"

namespace {
struct NcclApi {
	void *so = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	ncclResult_t (*GetVersion)(int *) = nullptr;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
	const char *names[] = {getenv("XMB_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};   // RTLD_LOCAL: the symbols are taken with dlsym
	for (const char *n : names) {
		if (!n || !*n) continue;
		g_nccl.so = dlopen(n, RTLD_NOW | RTLD_LOCAL);
		if (g_nccl.so) break;
	}
	if (!g_nccl.so) return;
	bool ok = true;
	auto sym = [&](const char *name) { void *p = dlsym(g_nccl.so, name); if (!p) ok = false; return p; };
	g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
	g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
	g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
	g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
	g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
	g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
	g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
	g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
	g_nccl.GetVersion = (decltype(g_nccl.GetVersion))sym("ncclGetVersion");
	if (!ok) { dlclose(g_nccl.so); g_nccl = NcclApi(); }
}
bool nccl_ready() {
	std::call_once(g_nccl_once, load_nccl);
	if (!g_nccl.so) xmb_set_error("NCCL not available: dlopen(libnccl.so.2) failed (%s); set XMB_NCCL_LIBRARY", dlerror() ? dlerror() : "symbols missing");
	return g_nccl.so != nullptr;
}
#define XMB_NCCL_OK(call)                                                                           \
	do {                                                                                            \
		ncclResult_t r_ = (call);                                                                   \
		if (r_ != ncclSuccess) { xmb_set_error("%s: %s", #call, g_nccl.GetErrorString(r_)); return 0; } \
	} while (0)
}   // namespace

struct xmb_comm {
	ncclComm_t comm = nullptr;
	int rank = 0, n_ranks = 1, device = 0;
};

static_assert(sizeof(ncclUniqueId) == XMB_COMM_ID_BYTES, "XMB_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");

extern "C" int xmb_comm_unique_id(char *id) {
	if (!id) { xmb_set_error("xmb_comm_unique_id: NULL"); return 0; }
	if (!nccl_ready()) return 0;
	ncclUniqueId u;
	XMB_NCCL_OK(g_nccl.GetUniqueId(&u));
	memcpy(id, &u, sizeof(u));
	return 1;
}

extern "C" int xmb_comm_init_rank(const char *id, int rank, int n_ranks, int device, xmb_comm **out) {
	if (!id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) { xmb_set_error("xmb_comm_init_rank: bad arguments"); return 0; }
	if (!nccl_ready()) return 0;
	if (device >= 0) XMB_CUDA_OK(cudaSetDevice(device));
	int dev = 0;
	XMB_CUDA_OK(cudaGetDevice(&dev));
	ncclUniqueId u;
	memcpy(&u, id, sizeof(u));
	xmb_comm *c = new xmb_comm();
	c->rank = rank; c->n_ranks = n_ranks; c->device = dev;
	const ncclResult_t r = g_nccl.CommInitRank(&c->comm, n_ranks, u, rank);
	if (r != ncclSuccess) { xmb_set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(r)); delete c; return 0; }
	*out = c;
	return 1;
}

extern "C" void xmb_comm_free(xmb_comm **c) {
	if (!c || !*c) return;
	if ((*c)->comm && g_nccl.so) { cudaSetDevice((*c)->device); g_nccl.CommDestroy((*c)->comm); }
	delete *c;
	*c = nullptr;
}

extern "C" int xmb_comm_rank(const xmb_comm *c) { return c ? c->rank : -1; }
extern "C" int xmb_comm_size(const xmb_comm *c) { return c ? c->n_ranks : 0; }
extern "C" int xmb_nccl_version(void) {
	int v = 0;
	if (!nccl_ready() || g_nccl.GetVersion(&v) != ncclSuccess) return 0;
	return v;
}

// This rank's shard: history kernel -> limbs -> all-reduce, all enqueued on the device's default stream.  The reduced
// limbs stay in HBM (xmb_msim_device_limbs); ex receives this rank's counters once the stream has drained.
extern "C" int xmb_main_msim_multi_raw(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, const xmb_main_options *options,
                                       const xmb_solid_angle *sa, xmb_comm *comm, xmb_msim_ex *ex) {
	XmbInputF *in = xmb_as_input(inputF);
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	if (!in || !h || !options || !comm || !ex) { xmb_set_error("xmb_main_msim_multi_raw: bad arguments"); return 0; }
	if (!nccl_ready()) return 0;
	ex->rank = comm->rank; ex->n_ranks = comm->n_ranks; ex->device = comm->device;
	XmbDeviceTables *D = nullptr;
	if (!xmb_msim_launch(in, h, options, sa, ex, &D)) return 0;
	XMB_NCCL_OK(g_nccl.AllReduce(D->limbs, D->limbs, 2 * D->run_slots, ncclUint64, ncclSum, comm->comm, (cudaStream_t)0));
	D->run_launches++;
	if (!xmb_msim_collect(D, options, ex)) return 0;
	XMB_CUDA_OK(cudaStreamSynchronize((cudaStream_t)0));
	return 1;
}

// xmi_main_msim over the communicator: every rank returns the full (summed) result, as if it had run alone.
extern "C" int xmb_main_msim_multi(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, xmb_comm *comm, double **channels,
                                   const xmb_main_options *options, double **brute_history, double **var_red_history,
                                   const xmb_solid_angle *solid_angles) {
	xmb_msim_ex ex{};
	ex.keep_on_device = 0;
	if (!xmb_main_msim_multi_raw(inputF, hdf5F, options, solid_angles, comm, &ex)) return 0;
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	XmbDeviceTables *D = h->dev;
	std::vector<uint64_t> limbs(2 * D->run_slots);
	XMB_CUDA_OK(cudaMemcpy(limbs.data(), D->limbs, sizeof(uint64_t) * limbs.size(), cudaMemcpyDeviceToHost));
	return xmb_main_msim_finish(inputF, hdf5F, options, limbs.data(), D->run_slots, channels, brute_history, var_red_history);
}

// One process, n_devices GPUs (device ordinals 0 .. n_devices-1, or devices[]): tables and grid are replicated, every
// device gets its shard's kernel before any of them is waited for, the all-reduces run as one NCCL group.
extern "C" int xmb_main_msim_all_devices(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, int n_devices, const int *devices,
                                         double **channels, const xmb_main_options *options, double **brute_history,
                                         double **var_red_history, const xmb_solid_angle *solid_angles, xmb_msim_ex *ex_out) {
	XmbInputF *in = xmb_as_input(inputF);
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	if (!in || !h || !options) { xmb_set_error("xmb_main_msim_all_devices: bad arguments"); return 0; }
	const int visible = xmb_cuda_device_count();
	if (n_devices <= 0) n_devices = visible;
	if (n_devices < 1 || n_devices > visible) { xmb_set_error("%d GPUs requested, %d visible", n_devices, visible); return 0; }
	std::vector<int> devs(n_devices);
	for (int i = 0; i < n_devices; i++) devs[i] = devices ? devices[i] : i;
	int dev0 = 0;
	cudaGetDevice(&dev0);
	std::vector<ncclComm_t> comms(n_devices, nullptr);
	if (n_devices > 1) {
		if (!nccl_ready()) return 0;
		XMB_NCCL_OK(g_nccl.CommInitAll(comms.data(), n_devices, devs.data()));
	}
	std::vector<xmb_msim_ex> ex(n_devices);
	std::vector<XmbDeviceTables *> D(n_devices, nullptr);
	int ok = 1;
	for (int i = 0; i < n_devices && ok; i++) {
		ex[i] = xmb_msim_ex{};
		ex[i].rank = i; ex[i].n_ranks = n_devices; ex[i].device = devs[i]; ex[i].seed = ex_out ? ex_out->seed : 0;
		ok = xmb_msim_launch(in, h, options, solid_angles, &ex[i], &D[i]);
	}
	if (ok && n_devices > 1) {
		g_nccl.GroupStart();
		for (int i = 0; i < n_devices; i++) {
			cudaSetDevice(devs[i]);
			const ncclResult_t r = g_nccl.AllReduce(D[i]->limbs, D[i]->limbs, 2 * D[i]->run_slots, ncclUint64, ncclSum, comms[i], (cudaStream_t)0);
			if (r != ncclSuccess) { xmb_set_error("ncclAllReduce: %s", g_nccl.GetErrorString(r)); ok = 0; }
		}
		const ncclResult_t r = g_nccl.GroupEnd();
		if (r != ncclSuccess) { xmb_set_error("ncclGroupEnd: %s", g_nccl.GetErrorString(r)); ok = 0; }
	}
	xmb_msim_ex total{};
	for (int i = 0; i < n_devices; i++) {
		if (!D[i]) continue;
		if (!xmb_msim_collect(D[i], options, &ex[i])) ok = 0;
		cudaSetDevice(devs[i]);
		if (cudaStreamSynchronize((cudaStream_t)0) != cudaSuccess) { xmb_set_error("device %d: %s", devs[i], cudaGetErrorString(cudaGetLastError())); ok = 0; }
		total.n_histories += ex[i].n_histories; total.n_interactions += ex[i].n_interactions;
		total.n_launches += ex[i].n_launches + (n_devices > 1 ? 1 : 0);
		if (ex[i].kernel_ms > total.kernel_ms) total.kernel_ms = ex[i].kernel_ms;
	}
	int rv = 0;
	if (ok) {
		// every device holds the same reduced limbs; read device 0's
		cudaSetDevice(devs[0]);
		std::vector<uint64_t> limbs(2 * D[0]->run_slots);
		if (cudaMemcpy(limbs.data(), D[0]->limbs, sizeof(uint64_t) * limbs.size(), cudaMemcpyDeviceToHost) != cudaSuccess) {
			xmb_set_error("reading the reduced histograms: %s", cudaGetErrorString(cudaGetLastError()));
		} else {
			h->dev = D[0];
			rv = xmb_main_msim_finish(inputF, hdf5F, options, limbs.data(), D[0]->run_slots, channels, brute_history, var_red_history);
		}
	}
	for (int i = 0; i < n_devices; i++) if (comms[i]) { cudaSetDevice(devs[i]); g_nccl.CommDestroy(comms[i]); }
	cudaSetDevice(dev0);
	if (ex_out) { total.rank = 0; total.n_ranks = n_devices; total.seed = ex_out->seed; total.device = devs[0]; *ex_out = total; }
	return rv;
}
