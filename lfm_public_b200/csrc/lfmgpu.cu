// liblfmgpu.so -- the C ABI of include/lfmgpu.h over the sm_100a kernels.
//
// Host-side bookkeeping only: uploads the flattened rank (lfmgpu_desc) into SoA device arrays, orders
// the kernels on a compute stream and a high-priority halo stream exactly as Mesh::solve orders the
// ISolver virtuals (reference: src/mesh_solver.cpp:474-691), and moves halo buffers either between
// handles of the same process or over NCCL (dlopen'ed, so the library loads on a box without NCCL).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "lfmgpu.h"
#include "stage_pipe.cuh"
#include "tile_kernels.cuh"

#ifndef LFM_AUSM_MINB64
#define LFM_AUSM_MINB64 2
#endif

using namespace lfm;

namespace {

thread_local std::string g_err;
int fail(const char* fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	g_err = buf;
	return 1;
}
#define CU(call) \
	do { \
		cudaError_t e_ = (call); \
		if (e_ != cudaSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)
#define TRY(call) \
	do { \
		int r_ = (call); \
		if (r_) return r_; \
	} while (0)

// ---- NCCL through dlopen ---------------------------------------------------------------------------
struct NcclApi {
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
int load_nccl() {
	if (g_nccl.lib) return 0;
	const char* names[] = {"libnccl.so.2", "libnccl.so"};
	void* lib = nullptr;
	for (const char* n : names) {
		lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (lib) break;
	}
	if (!lib) return fail("cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name) \
	*(void**)(&g_nccl.field) = dlsym(lib, name); \
	if (!g_nccl.field) return fail("libnccl lacks %s", name);
	SYM(GetUniqueId, "ncclGetUniqueId")
	SYM(CommInitRank, "ncclCommInitRank")
	SYM(CommDestroy, "ncclCommDestroy")
	SYM(Send, "ncclSend")
	SYM(Recv, "ncclRecv")
	SYM(AllReduce, "ncclAllReduce")
	SYM(GroupStart, "ncclGroupStart")
	SYM(GroupEnd, "ncclGroupEnd")
	SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
	g_nccl.lib = lib;
	return 0;
}
#define NC(call) \
	do { \
		ncclResult_t r_ = (call); \
		if (r_ != ncclSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
	} while (0)

struct TimedLaunch {
	const char* name;
	cudaEvent_t a, b;
};

}  // namespace

struct lfmgpu_ctx {
	int device = 0, prec = 8, D = 3, NQ = 5, F = 0;
	int n_cells = 0, n_faces = 0, n_bc = 0, n_mpi = 0, n_tot = 0, n_sub = 1;
	size_t ncs = 0, nfs = 0, ngs = 0;
	int sub_cell_start[LFMGPU_MAX_SUBMESH + 1] = {0};
	int sub_face_start[LFMGPU_MAX_SUBMESH + 1] = {0};
	lfmgpu_consts c{};
	std::vector<void*> allocs;
	DevMesh<double> md{};
	DevMesh<float> mf{};
	void* q[2] = {nullptr, nullptr};     // [ncs][QW] Q records (Rec<D>): conservatives + the derived values 1/rho, Rpsi, c|H
	bool drv_valid[2] = {false, false};  // the derived values in the records of q[b] match its conservatives
	bool drv_dirty_next = false;         // a submesh of the running stage was advanced without writing drv
	int drv_scheme = -1;                 // scheme the third derived value was computed for
	int cur = 0;
	unsigned updated_mask = 0;         // submeshes already advanced in the running stage
	bool dq_zero = true;               // prepare_for_timestep seen, no stage yet
	int rk_pending = 0;
	bool vis_on_cur = true;            // the conservatives calc_VIS last saw are q[cur] (false after a flip: q[1-cur])
	bool stage_done = false;           // at least one stage completed (q[1-cur] holds the pre-stage state)
	double* d_res = nullptr;           // [NQ] sum of RES^2
	double* d_res_partial = nullptr;
	void* d_partial = nullptr;         // block partials for cfl/dt
	void* d_scalar = nullptr;
	void* d_force_contrib = nullptr;
	int* d_force_used = nullptr;
	void* d_force_out = nullptr;
	cudaStream_t s_main = nullptr, s_comm = nullptr;
	// pipelined host I/O (lfmgpu_pipe_*): staging buffers and copy streams, created on first use
	cudaStream_t s_in = nullptr, s_out = nullptr;
	void *stage_in = nullptr, *stage_out = nullptr;
	cudaEvent_t ev_in_ready = nullptr, ev_in_free = nullptr, ev_out_ready = nullptr, ev_out_free = nullptr;
	cudaEvent_t ev_ready = nullptr;
	cudaEvent_t ev_user[8] = {nullptr};
	cudaEvent_t ev_packed[2] = {nullptr, nullptr}, ev_arrived[2] = {nullptr, nullptr};
	bool pending[2] = {false, false};
	// halo
	int n_nbr = 0, rank = 0, n_ranks = 1;
	std::vector<int> nbr_rank, send_start, recv_start;
	int* d_send_cell = nullptr;
	void* send_buf[2] = {nullptr, nullptr};
	void* recv_buf[2] = {nullptr, nullptr};
	size_t last_send_count[2] = {0, 0};
	int transport = 0;                 // 0 none, 1 local, 2 nccl
	std::vector<lfmgpu_ctx*> peers;
	ncclComm_t comm = nullptr;
	// tiles (fused path)
	TilePlan tiles;
	int use_tiles = 1;
	int tile_cells = 128;              // cells per tile requested (halved until the plan fits the budget)
	int stage_cfg = 0;
	int fixed_strides = 1;             // compile-time shared-memory strides when the plan fits them
	int les_opt = 0;                   // option "laminar" == 0: lfmgpu_step / _step_multi use calc_VIS_Smagorinsky
	int les = 0;                       // closure of the latest calc_VIS call (what the stage kernels and the halo pack see)
	int minmod_opt = 0;                // option "minmod": lfmgpu_step / _step_multi call calc_gradients_M2AUSM before calc_VIS (solver 2)
	bool ausm_set = false;             // calc_gradients_M2AUSM ran in the running stage
	bool ausm_zero = false;            // g_rho / g_p / g_U hold what prepare_for_RKstep leaves (zeros)
	std::vector<int> smag_owner;       // cell whose face loop leaves each cell's final tauMC (Smagorinsky constant)
	int tile_smem_budget = 75 * 1024;  // bytes of shared memory one tile CTA may use
	// persistent TMA-fed stage kernel (stage_pipe.cuh)
	bool pipe_ok = false;              // tensor maps and ring geometry are in place
	int pipe_enable = 3;               // LFMGPU_PIPE: bit 0 stage kernel, bit 1 gradient kernel; 0: the tile kernels of round 1 serve everything
	int pipe_slots_cap = 6;            // LFMGPU_PIPE_SLOTS
	int pipe_pf_dist = 2;              // L2 prefetch distance of the persistent kernels in tiles per CTA (LFMGPU_PIPE_PF, 0: off)
	int pipe_grad_slots_cap = 12;      // LFMGPU_PIPE_GSLOTS
	int grad_groups = 3;               // consumer groups of the gradient kernel, 3 or 4 (LFMGPU_PIPE_GGROUPS)
	int pipe_spare_sms = 8;            // SMs left to the halo stream's kernels on a rank with neighbours (LFMGPU_PIPE_SPARE)
	int n_sms = 0;
	PipeGeom pipe{};
	size_t pipe_smem = 0;
	bool grad_pipe_ok = false;
	GradGeom gpipe{};
	size_t gpipe_smem = 0;
	CUtensorMap map_q[2], map_v;       // record arrays as 2D tensors [ncs][QW | VW], box = one tile of own cells, swizzled
	CUtensorMap gmap_q[2], gmap_v;     // the same tensors with a box of one record: the maps of the halo gathers (tile::gather4)
	int smem_pad_kb = 0;               // extra shared memory requested per stage CTA (occupancy experiments)
	// CUDA graph of two consecutive time steps of a rank without neighbours (lfmgpu_step): ten RK stages return the double
	// buffer to its starting parity, so the same executable graph serves every following pair of steps with the same arguments
	int use_graph = 1;                 // LFMGPU_GRAPH=0: every kernel is launched from the host
	cudaGraphExec_t graph_exec = nullptr;
	uint64_t graph_kernels = 0;        // kernel launches inside the graph (added to `launches` at every replay)
	struct GraphKey {
		int scheme, want_res, minmod, les, cur, dq_zero_after;
		double dt;
		bool operator==(const GraphKey& o) const { return scheme == o.scheme && want_res == o.want_res && minmod == o.minmod && les == o.les && cur == o.cur && dt == o.dt; }
	} graph_key{};
	// introspection
	uint64_t launches = 0;
	bool timing = false;
	std::vector<TimedLaunch> timed;

	template <class R> DevMesh<R>& mesh();
};
template <> DevMesh<double>& lfmgpu_ctx::mesh<double>() { return md; }
template <> DevMesh<float>& lfmgpu_ctx::mesh<float>() { return mf; }

namespace {

int tile_plan_build(lfmgpu_ctx* h, const lfmgpu_desc* ds);
template <class R, int D> int tile_grad(lfmgpu_ctx* h, int sub);
template <class R, int D> int ensure_drv(lfmgpu_ctx* h, int scheme);
template <class R, int D> int tile_stage(lfmgpu_ctx* h, int sub, int scheme, R dt, R Ak, R Bk, int first, int res);

struct LaunchScope {
	lfmgpu_ctx* h;
	cudaStream_t s;
	TimedLaunch t{};
	LaunchScope(lfmgpu_ctx* h_, const char* name, cudaStream_t s_) : h(h_), s(s_) {
		h->launches++;
		if (h->timing) {
			t.name = name;
			cudaEventCreate(&t.a);
			cudaEventCreate(&t.b);
			cudaEventRecord(t.a, s);
		}
	}
	~LaunchScope() {
		if (h->timing) {
			cudaEventRecord(t.b, s);
			h->timed.push_back(t);
		}
	}
};
#define LAUNCH(h, name, stream, ...) \
	do { \
		LaunchScope ls_(h, name, stream); \
		__VA_ARGS__; \
	} while (0)
#define CHECK_LAUNCH() CU(cudaGetLastError())

inline int blocks_for(int n) { return (n + kBlock - 1) / kBlock; }

int dev_alloc(lfmgpu_ctx* h, void** p, size_t bytes, bool zero = true) {
	if (bytes == 0) bytes = 16;
	CU(cudaMalloc(p, bytes));
	h->allocs.push_back(*p);
	if (zero) CU(cudaMemset(*p, 0, bytes));
	return 0;
}

template <class T> int upload(lfmgpu_ctx* h, T** dst, const T* src, size_t n) {
	TRY(dev_alloc(h, (void**)dst, n * sizeof(T), n == 0));
	if (n) CU(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
	return 0;
}

// AoS [n][k] host -> SoA [k][stride] device
template <class R> int upload_soa(lfmgpu_ctx* h, R** dst, const R* src, size_t n, int k, size_t stride) {
	std::vector<R> tmp((size_t)k * stride, R(0));
	for (size_t i = 0; i < n; i++)
		for (int j = 0; j < k; j++) tmp[(size_t)j * stride + i] = src[i * k + j];
	return upload<R>(h, dst, tmp.data(), tmp.size());
}

inline size_t pad32(size_t n) { return (n + 31) / 32 * 32; }

template <class R> int build(lfmgpu_ctx* h, const lfmgpu_desc* ds) {
	const int D = h->D, NQ = h->NQ, F = h->F;
	DevMesh<R>& m = h->mesh<R>();
	m.n_cells = h->n_cells;
	m.n_faces = h->n_faces;
	m.n_bc = h->n_bc;
	m.n_mpi = h->n_mpi;
	m.n_tot = h->n_tot;
	m.F = F;
	m.ncs = h->ncs;
	m.nfs = h->nfs;
	m.ngs = h->ngs;
	int* p = nullptr;
	TRY(upload<int>(h, &p, ds->face_owner, (size_t)h->n_faces));
	m.face_owner = p;
	TRY(upload<int>(h, &p, ds->face_neigh, (size_t)h->n_faces));
	m.face_neigh = p;
	R* r = nullptr;
	TRY(upload_soa<R>(h, &r, (const R*)ds->face_S, (size_t)h->n_faces, D, h->nfs));
	m.S = r;
	TRY(upload_soa<R>(h, &r, (const R*)ds->face_d, (size_t)h->n_faces, D, h->nfs));
	m.d = r;
	TRY(upload<R>(h, &r, (const R*)ds->face_w, (size_t)h->n_faces));
	m.w = r;
	TRY(upload<R>(h, &r, (const R*)ds->vol_inv, (size_t)h->n_cells));
	m.vol_inv = r;
	TRY(upload<R>(h, &r, (const R*)ds->sponge_sigma, (size_t)h->n_cells));
	m.sigma = r;
	// cell -> faces: reference slot order and ascending-face order, both transposed to [F][n_cells]
	{
		std::vector<int> slot((size_t)F * h->n_cells, 0), csr((size_t)F * h->n_cells, 0);
		std::vector<int> tmp((size_t)F);
		for (int c = 0; c < h->n_cells; c++) {
			int n = 0;
			for (int s = 0; s < F; s++) {
				const int e = ds->cell_slot_face[(size_t)c * F + s];
				slot[(size_t)s * h->n_cells + c] = e;
				if (e) tmp[(size_t)n++] = e;
			}
			std::sort(tmp.begin(), tmp.begin() + n, [](int a, int b) { return std::abs(a) < std::abs(b); });
			for (int s = 0; s < n; s++) csr[(size_t)s * h->n_cells + c] = tmp[(size_t)s];
		}
		TRY(upload<int>(h, &p, slot.data(), slot.size()));
		m.slot = p;
		TRY(upload<int>(h, &p, csr.data(), csr.size()));
		m.csr = p;
	}
	const int QW = Rec<3>::QW, VW = D == 3 ? Rec<3>::VW : Rec<2>::VW;
	for (int b = 0; b < 2; b++) TRY(dev_alloc(h, &h->q[b], (size_t)QW * h->ncs * sizeof(R)));
	{
		std::vector<R> tmp((size_t)QW * h->ncs, R(0));
		const R* q0 = (const R*)ds->q0;
		for (int c = 0; c < h->n_cells; c++)
			for (int i = 0; i < NQ; i++) tmp[(size_t)c * QW + i] = q0[(size_t)c * NQ + i];
		CU(cudaMemcpy(h->q[0], tmp.data(), tmp.size() * sizeof(R), cudaMemcpyHostToDevice));
		CU(cudaMemcpy(h->q[1], tmp.data(), tmp.size() * sizeof(R), cudaMemcpyHostToDevice));
	}
	TRY(dev_alloc(h, (void**)&m.dq, (size_t)NQ * h->n_cells * sizeof(R)));
	TRY(dev_alloc(h, (void**)&m.RES, (size_t)NQ * h->n_cells * sizeof(R)));
	TRY(dev_alloc(h, (void**)&m.vis, (size_t)VW * h->ncs * sizeof(R)));
	m.tauMC = nullptr;   // allocated when the Smagorinsky closure is first used
	m.g_rho = m.g_p = m.g_U = nullptr;   // allocated when solver 2 (M2-AUSM) is first used
	m.les = 0;
	{
		// Smagorinsky constant -2 (Cs Delta)^2 per cell (cfd_v0.cpp:1601-1602), evaluated here with the C library's pow so that
		// it equals the reference's value bit for bit.  The reference recomputes a cell's tauMC at every face visit with the
		// constant of the VISITING cell: the final value comes from the cell's own loop if it owns a face, else from the owner
		// of its last incoming face.
		std::vector<R> sc((size_t)h->n_cells, R(0));
		const R* vol_inv = (const R*)ds->vol_inv;
		for (int c = 0; c < h->n_cells; c++) {
			int owner = -1, last_inc = 0;
			for (int s = 0; s < F; s++) {
				const int e = ds->cell_slot_face[(size_t)c * F + s];
				if (e > 0) owner = c;
				if (e < 0 && -e > last_inc) last_inc = -e;
			}
			if (owner < 0 && last_inc > 0) owner = ds->face_owner[last_inc - 1];
			if (owner >= 0) sc[(size_t)c] = -(R)2.0 * pow((R)0.16 * pow(1.0 / vol_inv[owner], 1. / 3.), 2);
		}
		TRY(upload<R>(h, &r, sc.data(), sc.size()));
		m.smag_c = r;
	}
	m.flux = nullptr;   // allocated on first use of the materialised path
	TRY(dev_alloc(h, (void**)&m.pAVG, (size_t)h->n_cells * sizeof(R)));
	TRY(dev_alloc(h, (void**)&m.pRMS, (size_t)h->n_cells * sizeof(R)));
	TRY(upload<int>(h, &p, ds->bc_cell, (size_t)h->n_bc));
	m.bc_cell = p;
	TRY(upload<int>(h, &p, ds->bc_kind, (size_t)h->n_bc));
	m.bc_kind = p;
	TRY(upload<int>(h, &p, ds->bc_face, (size_t)h->n_bc));
	m.bc_face = p;
	TRY(upload<int>(h, &p, ds->bc_patch, (size_t)h->n_bc));
	m.bc_patch = p;
	// constants, narrowed exactly as the CPU restatement narrows them
	const lfmgpu_consts& c = ds->c;
	Consts<R>& k = m.k;
	k.gamma = (R)c.gamma;
	k.gm1 = (R)c.gamma_m1;
	k.Rgas_inv = (R)c.Rgas_inv;
	k.mu = (R)c.mu;
	k.Cp = (R)c.Cp;
	k.Pr_inv = (R)c.Pr_inv;
	{
		volatile R t = k.Cp * k.mu;
		k.kappa = t * k.Pr_inv;
	}
	k.c_tau = (R)(2.0 / 3.0 * (double)k.mu);
	k.c_diag = (R)((double)k.mu * 2.0 / 3.0);
	k.rhoInf = (R)c.rhoInf;
	for (int i = 0; i < 3; i++) k.UInf[i] = (R)c.UInf[i];
	k.EInf = (R)c.EInf;
	k.pInf = (R)c.pInf;
	k.TInf = (R)c.TInf;
	for (int i = 0; i < 3; i++) {
		volatile R t = k.rhoInf * k.UInf[i];
		k.rhoUInf[i] = t;
	}
	{
		volatile R t = k.rhoInf * k.EInf;
		k.rhoEInf = t;
	}
	// scratch
	TRY(dev_alloc(h, &h->d_partial, (size_t)(blocks_for(h->n_cells) + 1) * sizeof(R)));
	TRY(dev_alloc(h, &h->d_scalar, 64));
	TRY(dev_alloc(h, (void**)&h->d_res, 8 * sizeof(double)));
	TRY(dev_alloc(h, (void**)&h->d_res_partial, (size_t)NQ * (blocks_for(h->n_cells) + 1) * sizeof(double)));
	TRY(dev_alloc(h, &h->d_force_contrib, (size_t)h->n_bc * (2 * D + D * D) * sizeof(R)));
	TRY(dev_alloc(h, (void**)&h->d_force_used, (size_t)h->n_bc * sizeof(int)));
	TRY(dev_alloc(h, &h->d_force_out, 16 * sizeof(R)));
	// halo
	if (h->n_nbr) {
		const size_t ns = (size_t)h->send_start[(size_t)h->n_nbr], nr = (size_t)h->recv_start[(size_t)h->n_nbr];
		TRY(upload<int>(h, &h->d_send_cell, ds->send_cell, ns));
		const size_t per = (size_t)(NQ + 2 * D * D + 2 * D) * sizeof(R);
		for (int s = 0; s < 2; s++) {
			TRY(dev_alloc(h, &h->send_buf[s], ns * per));
			TRY(dev_alloc(h, &h->recv_buf[s], nr * per));
		}
	}
	return 0;
}

int halo_mode(const lfmgpu_ctx* h, int step) { return h->c.comm_type == LFMGPU_COMM_SPLIT ? (step == 0 ? 1 : 2) : 3; }
int halo_spc(const lfmgpu_ctx* h, int step) {
	const int mode = halo_mode(h, step), D = h->D;
	return ((mode & 1) ? D + 2 : 0) + ((mode & 2) ? 2 * D * D + 2 * D : 0);
}

// buffer that holds the latest conservatives of cell range of submesh 0 (for packing)
void* q_for_pack(lfmgpu_ctx* h) { return (h->updated_mask & 1u) ? h->q[1 - h->cur] : h->q[h->cur]; }

template <class R, int D> int t_set_bc(lfmgpu_ctx* h) {
	if (!h->n_bc) return 0;
	LAUNCH(h, "k_set_bc", h->s_main, (k_set_bc<R, D><<<blocks_for(h->n_bc), kBlock, 0, h->s_main>>>(h->mesh<R>(), (R*)h->q[h->cur], h->drv_scheme < 0 ? 1 : h->drv_scheme)));
	CHECK_LAUNCH();
	return 0;
}

void sub_range(const lfmgpu_ctx* h, int sub, int& c0, int& c1, int& f0, int& f1) {
	if (sub < 0) {
		c0 = 0;
		c1 = h->n_cells;
		f0 = 0;
		f1 = h->n_faces;
	} else {
		c0 = h->sub_cell_start[sub];
		c1 = h->sub_cell_start[sub + 1];
		f0 = h->sub_face_start[sub];
		f1 = h->sub_face_start[sub + 1];
	}
}

// The tile kernels copy the derived values (1/rho, Rpsi, c|H) of every staged cell from drv[cur]: make them match
// q[cur].  In steady state the previous stage kernel, set_bc and the halo unpack have written them already; after
// an upload, a stage served by the unfused kernels or a change of scheme they are rebuilt here (ghosts included).
template <class R, int D> int ensure_drv(lfmgpu_ctx* h, int scheme) {
	const int want = scheme >= 0 ? scheme : (h->drv_scheme >= 0 ? h->drv_scheme : 1);
	if (h->drv_valid[h->cur] && want == h->drv_scheme) return 0;
	LAUNCH(h, "k_derive", h->s_main, (k_derive<R, D><<<blocks_for(h->n_tot), kBlock, 0, h->s_main>>>(h->mesh<R>(), (R*)h->q[h->cur], 0, h->n_tot, want)));
	CHECK_LAUNCH();
	h->drv_valid[h->cur] = true;
	h->drv_scheme = want;
	return 0;
}

template <class R, int D> int t_vis(lfmgpu_ctx* h, int sub, int les) {
	int c0, c1, f0, f1;
	sub_range(h, sub, c0, c1, f0, f1);
	h->vis_on_cur = true;
	DevMesh<R>& m = h->mesh<R>();
	if (les && !(h->use_tiles && h->tiles.ready)) return fail("the Smagorinsky closure is served by the tile kernels only (use_tiles=1 and a tileable mesh)");
	if (les && !m.tauMC) TRY(dev_alloc(h, (void**)&m.tauMC, (size_t)D * D * h->ncs * sizeof(R)));
	h->les = les;
	m.les = les;
	if (c1 <= c0) return 0;
	if (h->use_tiles && h->tiles.ready) {
		TRY((ensure_drv<R, D>(h, -1)));
		return tile_grad<R, D>(h, sub);
	}
	LAUNCH(h, "k_grad_cell", h->s_main, (k_grad_cell<R, D><<<blocks_for(c1 - c0), kBlock, 0, h->s_main>>>(h->mesh<R>(), (const R*)h->q[h->cur], c0, c1)));
	CHECK_LAUNCH();
	return 0;
}

// solver 2: the three minmod-gradient arrays exist only once M2-AUSM is used
template <class R, int D> int ausm_arrays(lfmgpu_ctx* h) {
	DevMesh<R>& m = h->mesh<R>();
	if (m.g_rho) return 0;
	TRY(dev_alloc(h, (void**)&m.g_rho, (size_t)D * h->ncs * sizeof(R)));
	TRY(dev_alloc(h, (void**)&m.g_p, (size_t)D * h->ncs * sizeof(R)));
	TRY(dev_alloc(h, (void**)&m.g_U, (size_t)D * D * h->ncs * sizeof(R)));
	h->ausm_zero = true;
	return 0;
}

// calc_gradients_M2AUSM (cfd_v0.cpp:1384-1495) of one submesh (or all: sub < 0)
template <class R, int D> int t_gradients_ausm(lfmgpu_ctx* h, int sub) {
	int c0, c1, f0, f1;
	sub_range(h, sub, c0, c1, f0, f1);
	TRY((ausm_arrays<R, D>(h)));
	h->ausm_set = true;
	h->ausm_zero = false;
	if (c1 <= c0) return 0;
	LAUNCH(h, "k_grad_ausm", h->s_main, (k_grad_ausm<R, D><<<blocks_for(c1 - c0), kBlock, 0, h->s_main>>>(h->mesh<R>(), (const R*)h->q[h->cur], c0, c1)));
	CHECK_LAUNCH();
	return 0;
}

template <class R, int D> int t_rk_stage(lfmgpu_ctx* h, int sub, int scheme, int rk, double dt, int want_res) {
	int c0, c1, f0, f1;
	sub_range(h, sub, c0, c1, f0, f1);
	DevMesh<R>& m = h->mesh<R>();
	if (scheme == LFMGPU_SCHEME_M2AUSM) {
		// without calc_gradients_M2AUSM in this stage the reference reads the zeros prepare_for_RKstep left (cfd_v0.cpp:1362-1371)
		TRY((ausm_arrays<R, D>(h)));
		if (!h->ausm_set && !h->ausm_zero) {
			CU(cudaMemsetAsync(m.g_rho, 0, (size_t)D * h->ncs * sizeof(R), h->s_main));
			CU(cudaMemsetAsync(m.g_p, 0, (size_t)D * h->ncs * sizeof(R), h->s_main));
			CU(cudaMemsetAsync(m.g_U, 0, (size_t)D * D * h->ncs * sizeof(R), h->s_main));
			h->ausm_zero = true;
		}
	}
	const R* q = (const R*)h->q[h->cur];
	R* qn = (R*)h->q[1 - h->cur];
	const int res = (want_res && rk == 0) ? 1 : 0;
	const int first = h->dq_zero ? 1 : 0;
	const R Ak = (R)h->c.Ak[rk], Bk = (R)h->c.Bk[rk];
	if (c1 > c0) {
		// solver 2 with the Smagorinsky closure would need both sets of extra rows staged: not served by either path
		if (scheme == LFMGPU_SCHEME_M2AUSM && h->les) return fail("M2-AUSM is served with the laminar closure only");
		if (h->use_tiles && h->tiles.ready) {
			TRY((ensure_drv<R, D>(h, scheme)));
			TRY((tile_stage<R, D>(h, sub, scheme, (R)dt, Ak, Bk, first, res)));
		} else {
			if (h->les) return fail("the Smagorinsky closure is served by the tile kernels only");
			h->drv_dirty_next = true;
			if (!m.flux) TRY(dev_alloc(h, (void**)&m.flux, (size_t)h->NQ * h->nfs * sizeof(R)));
			if (f1 > f0) {
				if (scheme == LFMGPU_SCHEME_M1)
					LAUNCH(h, "k_flux_face", h->s_main, (k_flux_face<R, D, 0><<<blocks_for(f1 - f0), kBlock, 0, h->s_main>>>(m, q, f0, f1)));
				else if (scheme == LFMGPU_SCHEME_M2AUSM)
					LAUNCH(h, "k_flux_face", h->s_main, (k_flux_face<R, D, 2><<<blocks_for(f1 - f0), kBlock, 0, h->s_main>>>(m, q, f0, f1)));
				else
					LAUNCH(h, "k_flux_face", h->s_main, (k_flux_face<R, D, 1><<<blocks_for(f1 - f0), kBlock, 0, h->s_main>>>(m, q, f0, f1)));
				CHECK_LAUNCH();
			}
			LAUNCH(h, "k_update_cell", h->s_main, (k_update_cell<R, D><<<blocks_for(c1 - c0), kBlock, 0, h->s_main>>>(m, q, qn, c0, c1, (R)dt, Ak, Bk, first, res)));
			CHECK_LAUNCH();
		}
		if (res) {
			const int nb = blocks_for(c1 - c0);
			LAUNCH(h, "k_res_partial", h->s_main, (k_res_partial<R><<<nb, kBlock, 0, h->s_main>>>(m.RES, h->n_cells, h->NQ, c0, c1, h->d_res_partial)));
			LAUNCH(h, "k_res_final", h->s_main, (k_res_final<<<h->NQ, kBlock, 0, h->s_main>>>(h->d_res_partial, nb, h->d_res)));
			CHECK_LAUNCH();
		}
	}
	// bookkeeping: flip the conservatives once every submesh has been advanced
	const unsigned all = (1u << h->n_sub) - 1u;
	h->updated_mask |= (sub < 0) ? all : (1u << sub);
	if (h->updated_mask == all) {
		h->cur = 1 - h->cur;
		h->drv_valid[h->cur] = !h->drv_dirty_next;
		h->drv_dirty_next = false;
		h->updated_mask = 0;
		h->dq_zero = false;
		h->stage_done = true;
		h->vis_on_cur = false;
		h->ausm_set = false;
	}
	return 0;
}

template <class R, int D> int t_pack(lfmgpu_ctx* h, int step, cudaStream_t s) {
	const int ns = h->send_start[(size_t)h->n_nbr];
	const int mode = halo_mode(h, step);
	h->last_send_count[step] = (size_t)ns * halo_spc(h, step);
	if (!ns) return 0;
	LAUNCH(h, "k_pack", s, (k_pack<R, D><<<blocks_for(ns), kBlock, 0, s>>>(h->mesh<R>(), (const R*)q_for_pack(h), h->d_send_cell, ns, mode, (R*)h->send_buf[step])));
	CHECK_LAUNCH();
	return 0;
}

template <class R, int D> int t_unpack(lfmgpu_ctx* h, int step) {
	const int nr = h->recv_start[(size_t)h->n_nbr];
	if (!nr) return 0;
	LAUNCH(h, "k_unpack", h->s_main, (k_unpack<R, D><<<blocks_for(nr), kBlock, 0, h->s_main>>>(h->mesh<R>(), (R*)h->q[h->cur], h->drv_scheme < 0 ? 1 : h->drv_scheme, nr, halo_mode(h, step), (const R*)h->recv_buf[step])));
	CHECK_LAUNCH();
	return 0;
}

template <class R, int D> int t_cfl_dt(lfmgpu_ctx* h, double arg, int what, double* out) {
	const int nb = blocks_for(h->n_cells);
	LAUNCH(h, "k_cfl_dt", h->s_main, (k_cfl_dt<R, D><<<nb, kBlock, 0, h->s_main>>>(h->mesh<R>(), (const R*)h->q[h->cur], (R)arg, what, (R*)h->d_partial)));
	LAUNCH(h, "k_reduce_minmax", h->s_main, (k_reduce_minmax<R><<<1, kBlock, 0, h->s_main>>>((const R*)h->d_partial, nb, what, (R*)h->d_scalar)));
	CHECK_LAUNCH();
	R v;
	CU(cudaMemcpyAsync(&v, h->d_scalar, sizeof(R), cudaMemcpyDeviceToHost, h->s_main));
	CU(cudaStreamSynchronize(h->s_main));
	*out = (double)v;
	return 0;
}

template <class R, int D> int t_average(lfmgpu_ctx* h, int time_step) {
	if (time_step <= 0) return 0;
	LAUNCH(h, "k_average", h->s_main, (k_average<R, D><<<blocks_for(h->n_cells), kBlock, 0, h->s_main>>>(h->mesh<R>(), (const R*)h->q[h->cur], time_step)));
	CHECK_LAUNCH();
	return 0;
}

template <class R, int D> int t_forces(lfmgpu_ctx* h, int patch, double* Fpre, double* Fvis) {
	for (int i = 0; i < D; i++) Fpre[i] = Fvis[i] = 0.0;
	if (!h->n_bc) return 0;
	LAUNCH(h, "k_forces_face", h->s_main, (k_forces_face<R, D><<<blocks_for(h->n_bc), kBlock, 0, h->s_main>>>(h->mesh<R>(), (const R*)h->q[h->cur], (const R*)h->q[h->stage_done ? 1 - h->cur : h->cur], patch, (R*)h->d_force_contrib, h->d_force_used)));
	LAUNCH(h, "k_forces_sum", h->s_main, (k_forces_sum<R, D><<<1, 32, 0, h->s_main>>>(h->n_bc, (const R*)h->d_force_contrib, h->d_force_used, (R*)h->d_force_out)));
	CHECK_LAUNCH();
	R out[6];
	CU(cudaMemcpyAsync(out, h->d_force_out, 2 * D * sizeof(R), cudaMemcpyDeviceToHost, h->s_main));
	CU(cudaStreamSynchronize(h->s_main));
	for (int i = 0; i < D; i++) {
		Fpre[i] = (double)out[i];
		Fvis[i] = (double)out[D + i];
	}
	return 0;
}

// ---- fused tile path: plan (host) + launchers --------------------------------------------------------
template <class R> TileView<R> tile_view(lfmgpu_ctx* h, int smax, int fmax) {
	TilePlan& p = h->tiles;
	TileView<R> v;
	v.tiles = p.d_tiles;
	v.halo_cell = p.d_halo_cell;
	v.f_idx = p.d_f_idx;
	v.f_gface = p.d_f_gface;
	v.fS = (const R*)p.d_fS;
	v.fK = (const R*)p.d_fK;
	v.fw = (const R*)p.d_fw;
	v.fdm = (const R*)p.d_fdm;
	v.fdi = (const R*)p.d_fdi;
	v.fSmag = (const R*)p.d_fSmag;
	v.T = p.T;
	v.csr_local = p.d_csr_local;
	v.smax = smax;
	v.fmax = fmax;
	return v;
}

void tile_range(const lfmgpu_ctx* h, int sub, int& t0, int& t1, int& smax, int& fmax) {
	const TilePlan& p = h->tiles;
	if (sub < 0) {
		t0 = 0;
		t1 = p.n_tiles;
		smax = fmax = 0;
		for (int s = 0; s < h->n_sub; s++) {
			smax = std::max(smax, p.sub_smax[s]);
			fmax = std::max(fmax, p.sub_fmax[s]);
		}
	} else {
		t0 = p.sub_tile_start[sub];
		t1 = p.sub_tile_start[sub + 1];
		smax = p.sub_smax[sub];
		fmax = p.sub_fmax[sub];
	}
}

constexpr int kGradThreads = 128;

// every submesh of the plan fits the compile-time shared-memory strides
bool all_fixed(const lfmgpu_ctx* h) {
	if (!h->fixed_strides) return false;
	for (int s = 0; s < h->n_sub; s++)
		if (h->tiles.sub_smax[s] > kFixedSmax || h->tiles.sub_fmax[s] > kFixedFmax) return false;
	return true;
}


template <class R, int D> size_t stage_smem(int smax, int fmax) { return ((size_t)StagedLayout<D>::NS * smax + (size_t)(D + 2) * fmax) * sizeof(R); }
template <class R, int D> size_t grad_smem(int smax, int fmax) { return ((size_t)(D + 2) * smax + (size_t)(D + 1) * fmax) * sizeof(R) + (size_t)fmax * sizeof(uint32_t); }

template <class R, int D> int grad_pipe(lfmgpu_ctx* h, int t0, int t1);
template <class R, int D> int tile_grad(lfmgpu_ctx* h, int sub) {
	if (h->grad_pipe_ok && !h->les) {   // the persistent TMA-fed kernel: laminar closure
		int t0, t1, smax, fmax;
		tile_range(h, sub, t0, t1, smax, fmax);
		if (t1 <= t0) return 0;
		return grad_pipe<R, D>(h, t0, t1);
	}
	// every submesh at once (a rank without neighbours): one launch when all of them fit the compile-time strides (same
	// shared-memory size anyway), else one launch each so that each gets its own shared-memory size
	if (sub < 0 && !all_fixed(h)) {
		for (int s = 0; s < h->n_sub; s++) TRY((tile_grad<R, D>(h, s)));
		return 0;
	}
	int t0, t1, smax, fmax;
	tile_range(h, sub, t0, t1, smax, fmax);
	if (t1 <= t0) return 0;
	const bool fixed = h->fixed_strides && smax <= kFixedSmax && fmax <= kFixedFmax;
	if (fixed) {
		smax = kFixedSmax;
		fmax = kFixedFmax;
	}
	const size_t smem = grad_smem<R, D>(smax, fmax);
	auto kern = h->les ? (fixed ? k_tile_grad<R, D, kGradThreads, kFixedSmax, kFixedFmax, 1> : k_tile_grad<R, D, kGradThreads, 0, 0, 1>)
	                   : (fixed ? k_tile_grad<R, D, kGradThreads, kFixedSmax, kFixedFmax, 0> : k_tile_grad<R, D, kGradThreads, 0, 0, 0>);
	if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	LAUNCH(h, "tile_grad", h->s_main, (kern<<<t1 - t0, kGradThreads, smem, h->s_main>>>(h->mesh<R>(), tile_view<R>(h, smax, fmax), (const R*)h->q[h->cur], t0)));
	CHECK_LAUNCH();
	return 0;
}

// ---- persistent TMA-fed stage kernel: tensor maps, ring geometry, launch -----------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// records [rows][width] of `prec`-byte values as a 2D tensor whose box is `box_rows` whole records, written to shared memory
// with the swizzle whose span is the record size (32, 64 or 128 bytes)
int make_record_map(CUtensorMap* out, void* base, int prec, int width, size_t rows, int box_rows) {
	static EncodeTiledFn encode = nullptr;
	if (!encode) {
		void* fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
		if (!fn || qres != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled is not available in this driver");
		encode = (EncodeTiledFn)fn;
	}
	const int bytes = width * prec;
	const CUtensorMapSwizzle sw = bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
	if (bytes != 128 && bytes != 64 && bytes != 32) return fail("record of %d bytes has no TMA swizzle mode", bytes);
	const cuuint64_t gdim[2] = {(cuuint64_t)width, (cuuint64_t)rows};
	const cuuint64_t gstride[1] = {(cuuint64_t)bytes};
	const cuuint32_t box[2] = {(cuuint32_t)width, (cuuint32_t)box_rows};
	const cuuint32_t estr[2] = {1, 1};
	const CUresult r = encode(out, prec == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
	                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for %d-byte records, box of %d rows", (int)r, bytes, box_rows);
	return 0;
}

inline uint32_t round_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

// dynamic shared memory of the persistent kernels of this rank's precision and dimension, on the handle's device
template <class R, int D> int pipe_attrs(lfmgpu_ctx* h) {
	if (h->pipe_ok) {
		CU(cudaFuncSetAttribute(k_stage_pipe<R, D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->pipe_smem));
		CU(cudaFuncSetAttribute(k_stage_pipe<R, D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->pipe_smem));
	}
	if (h->grad_pipe_ok) {
		CU(cudaFuncSetAttribute(k_grad_pipe<R, D, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gpipe_smem));
		CU(cudaFuncSetAttribute(k_grad_pipe<R, D, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gpipe_smem));
	}
	return 0;
}

// Called once the tile plan exists: decides whether the persistent kernel can serve this rank (every submesh cut with the
// same tile size <= 256 -- the height of the TMA box -- halos of at most 256 cells, at least two ring slots in shared memory)
int pipe_setup(lfmgpu_ctx* h) {
	h->pipe_ok = false;
	const TilePlan& p = h->tiles;
	if (!h->pipe_enable || !p.ready) return 0;
	int hmax = 0, fmax = 0;
	for (int s = 0; s < h->n_sub; s++) {
		if (p.sub_tile_start[s + 1] > p.sub_tile_start[s] && p.sub_tc[s] != h->tile_cells) return 0;
		hmax = std::max(hmax, p.sub_hmax[s]);
		fmax = std::max(fmax, p.sub_fmax[s]);
	}
	const int TC = h->tile_cells;
	if (TC > 256 || TC % 2 || hmax > kPipeMaxHalo) return 0;
	hmax = (std::max(hmax, 4) + 3) / 4 * 4;   // halo cells are gathered four at a time
	const int D = h->D, QW = Rec<3>::QW, VW = D == 3 ? Rec<3>::VW : Rec<2>::VW;
	const uint32_t QB = (uint32_t)(QW * h->prec), VB = (uint32_t)(VW * h->prec);
	int dev_smem = 0;
	CU(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
	CU(cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, h->device));
	PipeGeom g{};
	g.box_cells = TC;
	g.pf_dist = h->pipe_pf_dist;
	g.dbg = getenv("LFMGPU_PIPE_DBG") ? atoi(getenv("LFMGPU_PIPE_DBG")) : 0;
	g.coop = getenv("LFMGPU_PIPE_COOP") ? atoi(getenv("LFMGPU_PIPE_COOP")) : 1;
	g.wstore = getenv("LFMGPU_PIPE_WSTORE") ? atoi(getenv("LFMGPU_PIPE_WSTORE")) : 1;
	const int direct = getenv("LFMGPU_PIPE_DIRECT") ? atoi(getenv("LFMGPU_PIPE_DIRECT")) : 1;   // bit 0: stage kernel, bit 1: gradient kernel
	g.direct = direct & 1;
	g.pf_face = getenv("LFMGPU_PIPE_PFF") ? atoi(getenv("LFMGPU_PIPE_PFF")) : 0;
	g.pf_cell = getenv("LFMGPU_PIPE_PFC") ? atoi(getenv("LFMGPU_PIPE_PFC")) : 1;
	g.hmax = hmax;
	g.smax = TC + hmax;
	g.fmax = (fmax + 3) / 4 * 4;
	g.q_bytes = round_up((uint32_t)g.smax * QB, 1024);
	g.slot_bytes = g.q_bytes + round_up((uint32_t)g.smax * VB, 1024);
	for (int ns = std::min(6, h->pipe_slots_cap); ns >= 2; ns--) {
		g.n_slots = ns;
		g.off_bar = 0;
		g.off_meta = 256;   // slot barriers (16 bytes per slot, at most 16 slots) in front, then the producers' id-ring barriers and the id rings
		g.off_ids = g.off_meta + 2 * kProducerWarps * 8;
		g.off_fl = round_up(g.off_ids + (uint32_t)(2 * kProducerWarps) * ((hmax + 3) / 4 * 4) * 4u, 128);
		g.off_slot = round_up(g.off_fl + 2u * (uint32_t)h->NQ * g.fmax * (uint32_t)h->prec, 1024);
		const size_t total = (size_t)g.off_slot + (size_t)ns * g.slot_bytes + 1024;   // + slack to align the base to 1024
		if (total <= (size_t)dev_smem) {
			h->pipe = g;
			h->pipe_smem = total;
			h->pipe_ok = true;
			break;
		}
	}
	// the gradient kernel's ring: Q records only + the tile's face slice; its four groups need at least four slots
	h->grad_pipe_ok = false;
	if (TC <= kGradGroupThreads) {
		GradGeom gg{};
		gg.box_cells = TC;
		gg.pf_dist = getenv("LFMGPU_PIPE_GPF") ? atoi(getenv("LFMGPU_PIPE_GPF")) : 0;   // (the gradient kernel's fills are short: prefetching them measured slower)
		gg.dbg = g.dbg;
		gg.direct = (direct >> 1) & 1;
		gg.wstore = getenv("LFMGPU_PIPE_GWSTORE") ? atoi(getenv("LFMGPU_PIPE_GWSTORE")) : 1;
		gg.hmax = hmax;
		gg.smax = TC + hmax;
		gg.fmax = g.fmax;
		gg.q_bytes = round_up(std::max((uint32_t)gg.smax * QB, (uint32_t)TC * VB), 128);
		gg.slot_bytes = round_up(gg.q_bytes + (uint32_t)gg.fmax * (uint32_t)((D + 1) * h->prec + 4), 1024);   // only the Q region needs the swizzle's alignment
		const int NG = h->grad_groups;
		for (int ns = std::max(NG, std::min(12, h->pipe_grad_slots_cap)); ns >= NG; ns--) {
			if ((2 * ns) % NG) continue;   // see stage_pipe.cuh: the depth must keep a slot's use k - 2 inside the waiting group
			gg.n_slots = ns;
			gg.off_bar = 0;
			gg.off_meta = 256;   // (a ring of 9 slots needs 144 bytes of slot barriers: 128 was one slot short and let them overlap the id-ring barriers)
			gg.off_ids = gg.off_meta + 2 * kProducerWarps * 8;
			gg.off_slot = round_up(gg.off_ids + (uint32_t)(2 * kProducerWarps) * ((hmax + 3) / 4 * 4) * 4u, 1024);
			const size_t total = (size_t)gg.off_slot + (size_t)ns * gg.slot_bytes + 1024;
			if (total <= (size_t)dev_smem) {
				h->gpipe = gg;
				h->gpipe_smem = total;
				h->grad_pipe_ok = h->pipe_ok;
				break;
			}
		}
	}
	if (getenv("LFMGPU_PLAN_STATS"))
		fprintf(stderr, "[lfmgpu pipe] stage kernel: %s, %d slots of %u bytes, %zu bytes of shared memory; gradient kernel: %s, %d slots of %u bytes, %zu bytes; halo <= %d cells, faces <= %d\n",
		        h->pipe_ok ? "on" : "off", h->pipe.n_slots, h->pipe.slot_bytes, h->pipe_smem, h->grad_pipe_ok ? "on" : "off", h->gpipe.n_slots, h->gpipe.slot_bytes, h->gpipe_smem, hmax, g.fmax);
	if (!h->pipe_ok) return 0;
	for (int b = 0; b < 2; b++) TRY(make_record_map(&h->map_q[b], h->q[b], h->prec, QW, h->ncs, TC));
	for (int b = 0; b < 2; b++) TRY(make_record_map(&h->gmap_q[b], h->q[b], h->prec, QW, h->ncs, 1));
	const void* vis = h->prec == 8 ? (const void*)h->md.vis : (const void*)h->mf.vis;
	TRY(make_record_map(&h->map_v, (void*)vis, h->prec, VW, h->ncs, TC));
	TRY(make_record_map(&h->gmap_v, (void*)vis, h->prec, VW, h->ncs, 1));
	TRY(h->prec == 8 ? (D == 3 ? pipe_attrs<double, 3>(h) : pipe_attrs<double, 2>(h)) : (D == 3 ? pipe_attrs<float, 3>(h) : pipe_attrs<float, 2>(h)));
	return 0;
}

template <class R, int D> int grad_pipe(lfmgpu_ctx* h, int t0, int t1) {
	auto kern = h->grad_groups == 4 ? k_grad_pipe<R, D, 4> : k_grad_pipe<R, D, 3>;
	// (the kernel's shared-memory limit was raised in pipe_setup, not here: a launch may be under stream capture)
	const int sms = std::max(1, h->n_sms - (h->n_nbr ? h->pipe_spare_sms : 0));
	const int grid = std::min(t1 - t0, sms);
	LAUNCH(h, "tile_grad", h->s_main,
	       (kern<<<grid, grad_threads_total(h->grad_groups), h->gpipe_smem, h->s_main>>>(h->mesh<R>(), tile_view<R>(h, h->gpipe.smax, h->gpipe.fmax), h->map_q[h->cur], h->gmap_q[h->cur], h->gpipe, (const R*)h->q[h->cur], t0, t1 - t0)));
	CHECK_LAUNCH();
	return 0;
}

template <class R, int D, int SCHEME> int stage_pipe(lfmgpu_ctx* h, int t0, int t1, R dt, R Ak, R Bk, int first, int res) {
	auto kern = k_stage_pipe<R, D, SCHEME>;

	// a rank with neighbours keeps a few SMs free: the halo stream's pack / NCCL / unpack kernels must be able to start while
	// the interior submesh's persistent CTAs hold every register of the SMs they run on
	const int sms = std::max(1, h->n_sms - (h->n_nbr ? h->pipe_spare_sms : 0));
	const int grid = std::min(t1 - t0, sms);
	LAUNCH(h, "tile_stage", h->s_main,
	       (kern<<<grid, kPipeThreads, h->pipe_smem, h->s_main>>>(h->mesh<R>(), tile_view<R>(h, h->pipe.smax, h->pipe.fmax), h->map_q[h->cur], h->map_v, h->gmap_q[h->cur], h->gmap_v, h->pipe, (const R*)h->q[h->cur],
	                                                              (R*)h->q[1 - h->cur], t0, t1 - t0, dt, Ak, Bk, first, res)));
	CHECK_LAUNCH();
	return 0;
}

// resident CTAs the M2-AUSM stage kernel is compiled for (register cap 65536 / (256 * n))
template <class R> constexpr int kAusmMinBlocks = sizeof(R) == 8 ? LFM_AUSM_MINB64 : 2;

template <class R, int D, int SCHEME> int tile_stage_s(lfmgpu_ctx* h, int sub, R dt, R Ak, R Bk, int first, int res) {
	if constexpr (SCHEME != 2) {
		if (h->pipe_ok && !h->les) {   // the persistent TMA-fed kernel: laminar M1 / M2
			int t0, t1, smax, fmax;
			tile_range(h, sub, t0, t1, smax, fmax);
			if (t1 <= t0) return 0;
			return stage_pipe<R, D, SCHEME>(h, t0, t1, dt, Ak, Bk, first, res);
		}
	}
	if (sub < 0 && (!all_fixed(h) || SCHEME == 2)) {
		for (int s = 0; s < h->n_sub; s++) TRY((tile_stage_s<R, D, SCHEME>(h, s, dt, Ak, Bk, first, res)));
		return 0;
	}
	int t0, t1, smax, fmax;
	tile_range(h, sub, t0, t1, smax, fmax);
	if (t1 <= t0) return 0;
	bool fixed = h->fixed_strides && smax <= kFixedSmax && fmax <= kFixedFmax;
	if (fixed) {
		smax = kFixedSmax;
		fmax = kFixedFmax;
	}
	if (SCHEME == 2) {
		// solver 2: the 2D + D^2 gradient rows are staged too; with the launch's own strides (not the compile-time ones) two
		// CTAs still fit an SM
		fixed = false;
		tile_range(h, sub, t0, t1, smax, fmax);
	}
	size_t smem = stage_smem<R, D>(smax, fmax) + (h->les ? (size_t)D * D * smax * sizeof(R) : 0) + (SCHEME == 2 ? (size_t)(2 * D + D * D) * smax * sizeof(R) : 0);
	smem += (size_t)h->smem_pad_kb * 1024;   // experiment knob LFMGPU_SMEM_PAD: fewer resident CTAs
	TileView<R> tview = tile_view<R>(h, smax, fmax);
	if constexpr (SCHEME == 2) {   // M2-AUSM: one configuration (256 threads), laminar closure
		const int nt_ = 256;
		int dev_smem = 0;
		CU(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
		if (smem > (size_t)dev_smem) return fail("M2-AUSM: a tile needs %zu bytes of shared memory (limit %d): lower LFMGPU_TILE_CELLS or set use_tiles=0", smem, dev_smem);
		auto kern = k_tile_stage<R, D, SCHEME, 256, kAusmMinBlocks<R>, 0, 0, 2>;
		if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		LAUNCH(h, "tile_stage", h->s_main,
		       (kern<<<t1 - t0, nt_, smem, h->s_main>>>(h->mesh<R>(), tview, (const R*)h->q[h->cur], (R*)h->q[1 - h->cur], t0, dt, Ak, Bk,
		                                              first, res)));
		CHECK_LAUNCH();
		return 0;
	} else {
	if (h->les) {   // Smagorinsky closure: tauMC is staged too (one configuration: 256 threads, 2 CTAs/SM)
		const int nt_ = 256;
		int dev_smem = 0;
		CU(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
		if (smem > (size_t)dev_smem) return fail("Smagorinsky closure: a tile needs %zu bytes of shared memory (limit %d): lower LFMGPU_TILE_CELLS", smem, dev_smem);
#define LFM_STAGE_LES(SM_, FM_) \
		{ \
			auto kern = k_tile_stage<R, D, SCHEME, 256, 2, SM_, FM_, 1>; \
			if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
			LAUNCH(h, "tile_stage", h->s_main, \
			       (kern<<<t1 - t0, nt_, smem, h->s_main>>>(h->mesh<R>(), tview, (const R*)h->q[h->cur], (R*)h->q[1 - h->cur], t0, dt, Ak, \
			                                              Bk, first, res))); \
		}
		if (fixed)
			LFM_STAGE_LES(kFixedSmax, kFixedFmax)
		else
			LFM_STAGE_LES(0, 0)
#undef LFM_STAGE_LES
		CHECK_LAUNCH();
		return 0;
	}
#define LFM_STAGE_LAUNCH(...) \
	{ \
		auto kern = k_tile_stage<R, D, SCHEME, __VA_ARGS__>; \
		if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
		LAUNCH(h, "tile_stage", h->s_main, \
		       (kern<<<t1 - t0, nt_, smem, h->s_main>>>(h->mesh<R>(), tview, (const R*)h->q[h->cur], (R*)h->q[1 - h->cur], t0, dt, Ak, Bk, \
		                                              first, res))); \
	}
#define LFM_STAGE_CFG(NT_, MB_) \
	{ \
		const int nt_ = NT_; \
		if (fixed) \
			LFM_STAGE_LAUNCH(NT_, MB_, kFixedSmax, kFixedFmax) \
		else \
			LFM_STAGE_LAUNCH(NT_, MB_, 0, 0) \
	}
	switch (h->stage_cfg) {
		case 1: LFM_STAGE_CFG(128, 4) break;
		case 6: LFM_STAGE_CFG(256, 2) break;
		case 12: LFM_STAGE_CFG(512, 1) break;
		default: LFM_STAGE_CFG(256, 3) break;
	}
#undef LFM_STAGE_CFG
#undef LFM_STAGE_LAUNCH
	CHECK_LAUNCH();
	return 0;
	}   // SCHEME != 2
}
template <class R, int D> int tile_stage(lfmgpu_ctx* h, int sub, int scheme, R dt, R Ak, R Bk, int first, int res) {
	if (scheme == LFMGPU_SCHEME_M2AUSM) return tile_stage_s<R, D, 2>(h, sub, dt, Ak, Bk, first, res);
	return scheme == LFMGPU_SCHEME_M1 ? tile_stage_s<R, D, 0>(h, sub, dt, Ak, Bk, first, res) : tile_stage_s<R, D, 1>(h, sub, dt, Ak, Bk, first, res);
}

// per-mesh-face constants (temporary) -> the plan's tile-ordered face tables
template <class R, int D> int tile_geo(lfmgpu_ctx* h) {
	TilePlan& p = h->tiles;
	const size_t T = p.T;
	R *gK = nullptr, *gdm = nullptr, *gdi = nullptr, *gSm = nullptr;
	CU(cudaMalloc((void**)&gK, std::max<size_t>(16, (size_t)D * h->nfs * sizeof(R))));
	CU(cudaMalloc((void**)&gdm, std::max<size_t>(16, h->nfs * sizeof(R))));
	CU(cudaMalloc((void**)&gdi, std::max<size_t>(16, h->nfs * sizeof(R))));
	CU(cudaMalloc((void**)&gSm, std::max<size_t>(16, h->nfs * sizeof(R))));
	int rc = 0;
	do {
		if ((rc = dev_alloc(h, &p.d_fS, (size_t)D * T * sizeof(R)))) break;
		if ((rc = dev_alloc(h, &p.d_fK, (size_t)D * T * sizeof(R)))) break;
		if ((rc = dev_alloc(h, &p.d_fw, T * sizeof(R)))) break;
		if ((rc = dev_alloc(h, &p.d_fdm, T * sizeof(R)))) break;
		if ((rc = dev_alloc(h, &p.d_fdi, T * sizeof(R)))) break;
		if ((rc = dev_alloc(h, &p.d_fSmag, T * sizeof(R)))) break;
		if (h->n_faces && p.n_table) {
			k_face_geo<R, D><<<blocks_for(h->n_faces), kBlock>>>(h->mesh<R>(), gK, gdm, gdi, gSm);
			const unsigned nb = (unsigned)((p.n_table + kBlock - 1) / kBlock);
			k_tile_face_tables<R, D><<<nb, kBlock>>>(h->mesh<R>(), gK, gdm, gdi, gSm, p.d_f_gface, p.n_table, T, (R*)p.d_fS, (R*)p.d_fK, (R*)p.d_fw, (R*)p.d_fdm, (R*)p.d_fdi, (R*)p.d_fSmag);
			if (cudaDeviceSynchronize() != cudaSuccess) rc = fail("tile face tables: %s", cudaGetErrorString(cudaGetLastError()));
		}
	} while (0);
	cudaFree(gK);
	cudaFree(gdm);
	cudaFree(gdi);
	cudaFree(gSm);
	return rc;
}

// Cuts every submesh into runs of about `tile_cells` consecutive cells and derives, per tile, the halo cells, the
// tile-ordered face list (own faces grouped by rank among the owner's faces, then the incoming faces) and the
// tile-local indices the kernels use.  Returns 0 with plan.ready == false when the mesh does not fit the
// shared-memory budget even with the smallest tile (the unfused kernels serve it then).
// The host half of the plan: everything the tile kernels index with, built from the flattened rank alone (no device), so
// that it can be checked on a machine without a GPU (lfmgpu_plan_check, tests/test_tile_plan.py).
struct HostPlan {
	bool ok = false;                   // false: not tileable within the budget (the unfused kernels serve the rank)
	std::vector<TileDesc> tiles;
	std::vector<int> halo_cell, f_gface;
	std::vector<uint32_t> f_idx;
	std::vector<int16_t> csr_local;
	int TCs[LFMGPU_MAX_SUBMESH] = {0};
};

int tile_plan_host(lfmgpu_ctx* h, const lfmgpu_desc* ds, int dev_smem, HostPlan& hp) {
	TilePlan& p = h->tiles;
	hp.ok = false;
	const int nc = h->n_cells, nf = h->n_faces, F = h->F, D = h->D;
	if (nc == 0) return 0;
	const size_t budget = std::min<size_t>((size_t)dev_smem, (size_t)h->tile_smem_budget);
	const size_t es = (size_t)h->prec;
	const int NS = (D == 3 ? StagedLayout<3>::NS : StagedLayout<2>::NS), NQ = h->NQ;

	// sorted signed cell->face lists (same order the device csr uses)
	std::vector<int> csr((size_t)nc * F, 0);
	for (int c = 0; c < nc; c++) {
		int n = 0;
		int* row = &csr[(size_t)c * F];
		for (int s = 0; s < F; s++) {
			const int e = ds->cell_slot_face[(size_t)c * F + s];
			if (e) row[n++] = e;
		}
		std::sort(row, row + n, [](int a, int b) { return std::abs(a) < std::abs(b); });
	}
	std::vector<int> cfs((size_t)nc + 1, 0);   // first own face of each cell
	for (int f = 0; f < nf; f++) cfs[(size_t)ds->face_owner[f] + 1]++;
	for (int c = 0; c < nc; c++) cfs[(size_t)c + 1] += cfs[(size_t)c];

	std::vector<TileDesc>& tiles = hp.tiles;
	std::vector<int>&halo_cell = hp.halo_cell, &f_gface = hp.f_gface;
	std::vector<uint32_t>& f_idx = hp.f_idx;
	std::vector<int16_t>& csr_local = hp.csr_local;
	std::vector<int> stamp((size_t)h->n_tot, -1), local_of((size_t)h->n_tot, 0);
	std::vector<int> own_pos, inc_rank((size_t)nc, 0);
	std::vector<std::pair<int, int>> inc_sorted;   // (mesh face, tile-local index) of the tile's incoming faces
	struct Inc {
		int rank, ln, f;
	};
	std::vector<Inc> inc;
	int* TCs = hp.TCs;
	for (int s = 0; s < LFMGPU_MAX_SUBMESH; s++) TCs[s] = h->tile_cells;
	int probe_id = -1;
	for (;;) {
		int failed_sub = -1;
		tiles.clear();
		halo_cell.clear();
		f_gface.clear();
		f_idx.clear();
		csr_local.assign((size_t)F * nc, 0);
		std::fill(stamp.begin(), stamp.end(), -1);
		bool ok = true;
		long long tot_own = 0, tot_inc = 0, tot_halo = 0;
		for (int s = 0; s < h->n_sub && ok; s++) {
			const int TC = TCs[s];
			// shared-memory caps of one tile: faces ~ (F/2 own + 25% incoming) per cell, the rest of the budget for staged cells
			int f_cap = std::max(F, (int)(0.625 * F * TC));
			int s_cap = (int)(((long long)(budget / es) - (long long)NQ * (f_cap + 4)) / NS) - 4;
			if (h->fixed_strides && TC <= 160 && s_cap >= kFixedSmax / 2) {   // keep ordinary tiles inside the compile-time strides
				f_cap = std::min(f_cap, kFixedFmax);
				s_cap = std::min(s_cap, kFixedSmax);
			}
			if (s_cap < 2 * F) {
				ok = false;
				failed_sub = s;
				break;
			}
			p.sub_tile_start[s] = (int)tiles.size();
			p.sub_tc[s] = TC;
			int smax = 0, fmax = 0, hmax = 0;
			for (int c0 = h->sub_cell_start[s], c1 = 0; c0 < h->sub_cell_start[s + 1]; c0 = c1) {
				const int cmax = std::min(c0 + TC, h->sub_cell_start[s + 1]);
				// Choose the cut: grow the run cell by cell, tracking halo cells and incoming faces incrementally, and
				// close the tile where halo/tile is smallest inside the window [0.55 TC, TC] among the prefixes that fit
				// the shared-memory caps (ties: the longer run); if no prefix in the window fits, the longest that does.
				// For block- or space-filling-curve-numbered meshes this snaps tiles to the compact runs of the numbering.
				{
					const int probe = --probe_id;   // negative stamps: never collide with tile ids
					int nh = 0, ninc = 0;
					double best = 1e300;
					int best_in_window = -1, longest_fit = -1;
					const int cmin = std::min(cmax, c0 + std::max(1, (TC * 11) / 20));
					for (int c = c0; c < cmax; c++) {
						if (stamp[(size_t)c] == probe) nh--;
						const int* row = &csr[(size_t)c * F];
						for (int k = 0; k < F && row[k]; k++) {
							const int e = row[k], f = std::abs(e) - 1;
							const int other = e > 0 ? ds->face_neigh[f] : ds->face_owner[f];
							if (other >= c0 && other <= c) continue;
							if (e < 0) ninc++;
							if (stamp[(size_t)other] != probe) {
								stamp[(size_t)other] = probe;
								nh++;
							}
						}
						const int nt = c + 1 - c0;
						const bool fits = TC + nh <= s_cap && (cfs[(size_t)c + 1] - cfs[(size_t)c0]) + ninc <= f_cap;   // halo cells are staged from slot TC on
						if (!fits) continue;   // (not monotone: a later prefix may fit again once halo cells join the tile)
						longest_fit = c + 1;
						if (c + 1 >= cmin) {
							const double ratio = (double)nh / (double)nt;
							if (ratio <= best) {
								best = ratio;
								best_in_window = c + 1;
							}
						}
					}
					c1 = best_in_window > 0 ? best_in_window : longest_fit;
					if (c1 <= c0) {
						ok = false;   // not even one cell fits behind a halo base of TC slots: retry with shorter tiles
						failed_sub = s;
						break;
					}
				}
				const int tid = (int)tiles.size();
				TileDesc td{};
				td.c0 = c0;
				td.nt = c1 - c0;
				const int fo0 = cfs[(size_t)c0];
				td.nfo = cfs[(size_t)c1] - fo0;
				while (halo_cell.size() % 4) halo_cell.push_back(0);   // a tile's halo list starts on a 16-byte boundary (16-byte async copies)
				td.halo_off = (int)halo_cell.size();
				while (f_gface.size() % 4) {   // a tile's slice of the face tables starts on a 16-byte boundary (bulk copies)
					f_gface.push_back(0);
					f_idx.push_back(0);
				}
				td.f_off = (int)f_gface.size();
				// halo cells and incoming faces
				inc.clear();
				for (int c = c0; c < c1; c++) {
					const int* row = &csr[(size_t)c * F];
					int r = 0;
					for (int k = 0; k < F && row[k]; k++) {
						const int e = row[k], f = std::abs(e) - 1;
						const int other = e > 0 ? ds->face_neigh[f] : ds->face_owner[f];
						if (other >= c0 && other < c1) continue;
						if (stamp[(size_t)other] != tid) {
							stamp[(size_t)other] = tid;
							halo_cell.push_back(other);
						}
						if (e < 0) inc.push_back({r++, c - c0, f});
					}
				}
				std::sort(halo_cell.begin() + td.halo_off, halo_cell.end());
				td.nh = (int)halo_cell.size() - td.halo_off;
				td.ninc = (int)inc.size();
				td.hb = TC;
				for (int i = 0; i < td.nh; i++) local_of[(size_t)halo_cell[(size_t)td.halo_off + i]] = td.hb + i;
				if (td.hb + td.nh >= 32768 || td.nfo + td.ninc >= 32767 || f_gface.size() + (size_t)td.nfo + td.ninc >= (size_t)0x7fffffff) {
					ok = false;
					failed_sub = s;
					break;
				}
				auto staged = [&](int x) { return (x >= c0 && x < c1) ? x - c0 : local_of[(size_t)x]; };
				// own faces: grouped by rank among the owner's faces, owners ascending inside a group
				own_pos.assign((size_t)td.nfo, 0);
				for (int r = 0, placed = 0; placed < td.nfo; r++)
					for (int c = c0; c < c1; c++) {
						const int f = cfs[(size_t)c] + r;
						if (f >= cfs[(size_t)c + 1]) continue;
						own_pos[(size_t)(f - fo0)] = placed++;
						const int n = ds->face_neigh[f];
						const bool ghost = n >= nc && n < nc + h->n_bc;
						f_gface.push_back(f);
						f_idx.push_back((uint32_t)(c - c0) | ((uint32_t)staged(n) << 16) | (ghost ? 0x80000000u : 0u));
					}
				// incoming faces: grouped by rank among the tile-side cell's incoming faces
				std::sort(inc.begin(), inc.end(), [](const Inc& a, const Inc& b) { return a.rank != b.rank ? a.rank < b.rank : (a.ln != b.ln ? a.ln < b.ln : a.f < b.f); });
				inc_sorted.clear();
				for (int k = 0; k < td.ninc; k++) {
					const int f = inc[(size_t)k].f;
					f_gface.push_back(f);
					f_idx.push_back((uint32_t)staged(ds->face_owner[f]) | ((uint32_t)inc[(size_t)k].ln << 16));
					inc_sorted.push_back({f, td.nfo + k});
				}
				std::sort(inc_sorted.begin(), inc_sorted.end());
				for (int c = c0; c < c1; c++) {
					const int* row = &csr[(size_t)c * F];
					for (int k = 0; k < F && row[k]; k++) {
						const int e = row[k], f = std::abs(e) - 1;
						int lf;
						if (f >= fo0 && f < fo0 + td.nfo) {
							lf = own_pos[(size_t)(f - fo0)];
						} else {
							auto it = std::lower_bound(inc_sorted.begin(), inc_sorted.end(), std::make_pair(f, -1));
							lf = it->second;
						}
						csr_local[(size_t)k * nc + c] = (int16_t)(e > 0 ? lf + 1 : -(lf + 1));
					}
				}
				smax = std::max(smax, td.hb + td.nh);
				hmax = std::max(hmax, td.nh);
				fmax = std::max(fmax, td.nfo + td.ninc);
				tot_own += td.nfo;
				tot_inc += td.ninc;
				tot_halo += td.nh;
				tiles.push_back(td);
			}
			// keep shared-memory rows 16-byte aligned
			smax = (smax + 3) / 4 * 4;
			fmax = (fmax + 3) / 4 * 4;
			p.sub_smax[s] = smax;
			p.sub_fmax[s] = fmax;
			p.sub_hmax[s] = hmax;
			if (ok && ((size_t)NS * smax + (size_t)NQ * fmax) * es > budget) {
				ok = false;
				failed_sub = s;
			}
		}
		p.sub_tile_start[h->n_sub] = (int)tiles.size();
		if (ok) {
			p.halo_face_ratio = tot_own ? (double)tot_inc / (double)tot_own : 0.0;
			p.halo_cell_ratio = (double)tot_halo / (double)nc;
			break;
		}
		if (failed_sub < 0 || TCs[failed_sub] <= 16) return 0;   // not tileable within the budget: unfused kernels
		TCs[failed_sub] /= 2;
	}
	hp.ok = true;
	return 0;
}

int tile_plan_build(lfmgpu_ctx* h, const lfmgpu_desc* ds) {
	TilePlan& p = h->tiles;
	p.ready = false;
	const int D = h->D;
	if (h->n_cells == 0) return 0;
	int dev_smem = 0;
	CU(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
	HostPlan hp;
	TRY(tile_plan_host(h, ds, dev_smem, hp));
	if (!hp.ok) return 0;
	std::vector<TileDesc>& tiles = hp.tiles;
	std::vector<int>&halo_cell = hp.halo_cell, &f_gface = hp.f_gface;
	std::vector<uint32_t>& f_idx = hp.f_idx;
	std::vector<int16_t>& csr_local = hp.csr_local;
	const int* TCs = hp.TCs;
	const size_t es = (size_t)h->prec;
	const int NS = (D == 3 ? StagedLayout<3>::NS : StagedLayout<2>::NS), NQ = h->NQ;
	if (getenv("LFMGPU_PLAN_STATS")) {
		for (int sb = 0; sb < h->n_sub; sb++) {
			std::vector<int> a, b, c;
			for (int t = p.sub_tile_start[sb]; t < p.sub_tile_start[sb + 1]; t++) {
				a.push_back(tiles[(size_t)t].hb + tiles[(size_t)t].nh);
				b.push_back(tiles[(size_t)t].nfo + tiles[(size_t)t].ninc);
				c.push_back(tiles[(size_t)t].nt);
			}
			if (a.empty()) continue;
			std::sort(a.begin(), a.end());
			std::sort(b.begin(), b.end());
			std::sort(c.begin(), c.end());
			auto q = [](const std::vector<int>& v, double f) { return v[(size_t)(f * (v.size() - 1))]; };
			fprintf(stderr, "[lfmgpu plan] sub %d: TC %d, %zu tiles; cells/tile min %d med %d max %d; staged med %d p90 %d p99 %d max %d; faces med %d p90 %d p99 %d max %d\n", sb,
			        TCs[sb], a.size(), c.front(), q(c, 0.5), c.back(), q(a, 0.5), q(a, 0.9), q(a, 0.99), a.back(), q(b, 0.5), q(b, 0.9), q(b, 0.99), b.back());
		}
	}
	p.tile_cells = TCs[h->n_sub - 1];
	p.n_tiles = (int)tiles.size();
	int smax = 0, fmax = 0;
	for (int s = 0; s < h->n_sub; s++) {
		smax = std::max(smax, p.sub_smax[s]);
		fmax = std::max(fmax, p.sub_fmax[s]);
	}
	p.smem_bytes = ((size_t)NS * smax + (size_t)NQ * fmax) * es;
	p.n_table = f_gface.size();
	p.T = pad32(p.n_table + 1);
	TRY(upload<TileDesc>(h, &p.d_tiles, tiles.data(), tiles.size()));
	halo_cell.resize(halo_cell.size() + 4, 0);   // the last tile's 16-byte id copies may read up to three entries past its list
	TRY(upload<int>(h, &p.d_halo_cell, halo_cell.data(), halo_cell.size()));
	f_gface.resize(p.T, 0);
	f_idx.resize(p.T, 0);
	TRY(upload<int>(h, &p.d_f_gface, f_gface.data(), f_gface.size()));
	TRY(upload<uint32_t>(h, &p.d_f_idx, f_idx.data(), f_idx.size()));
	TRY(upload<int16_t>(h, &p.d_csr_local, csr_local.data(), csr_local.size()));
	TRY(h->prec == 8 ? (D == 3 ? tile_geo<double, 3>(h) : tile_geo<double, 2>(h)) : (D == 3 ? tile_geo<float, 3>(h) : tile_geo<float, 2>(h)));
	p.ready = true;
	return 0;
}

#define DISPATCH(h, fn, ...) \
	((h)->prec == 8 ? ((h)->D == 3 ? fn<double, 3>(__VA_ARGS__) : fn<double, 2>(__VA_ARGS__)) \
	                : ((h)->D == 3 ? fn<float, 3>(__VA_ARGS__) : fn<float, 2>(__VA_ARGS__)))

int halo_start_impl(lfmgpu_ctx* h, int step) {
	if (h->n_nbr == 0 || h->transport == 0) return 0;
	// the pack reads what the compute stream has produced up to here
	CU(cudaEventRecord(h->ev_ready, h->s_main));
	CU(cudaStreamWaitEvent(h->s_comm, h->ev_ready, 0));
	if (h->transport == 1) {
		// consumers of my previous send buffer must have copied it
		for (int i = 0; i < h->n_nbr; i++) {
			lfmgpu_ctx* p = h->peers[(size_t)h->nbr_rank[(size_t)i]];
			CU(cudaStreamWaitEvent(h->s_comm, p->ev_arrived[step], 0));
		}
	}
	TRY(DISPATCH(h, t_pack, h, step, h->s_comm));
	CU(cudaEventRecord(h->ev_packed[step], h->s_comm));
	if (h->transport == 2) {
		const int spc = halo_spc(h, step);
		const ncclDataType_t ty = h->prec == 8 ? ncclDouble : ncclFloat;
		NC(g_nccl.GroupStart());
		for (int i = 0; i < h->n_nbr; i++) {
			const size_t so = (size_t)h->send_start[(size_t)i] * spc, sn = (size_t)(h->send_start[(size_t)i + 1] - h->send_start[(size_t)i]) * spc;
			const size_t ro = (size_t)h->recv_start[(size_t)i] * spc, rn = (size_t)(h->recv_start[(size_t)i + 1] - h->recv_start[(size_t)i]) * spc;
			NC(g_nccl.Send((const char*)h->send_buf[step] + so * h->prec, sn, ty, h->nbr_rank[(size_t)i], h->comm, h->s_comm));
			NC(g_nccl.Recv((char*)h->recv_buf[step] + ro * h->prec, rn, ty, h->nbr_rank[(size_t)i], h->comm, h->s_comm));
		}
		NC(g_nccl.GroupEnd());
		h->launches++;
		CU(cudaEventRecord(h->ev_arrived[step], h->s_comm));
	}
	h->pending[step] = true;
	return 0;
}

int halo_wait_impl(lfmgpu_ctx* h, int step) {
	if (h->n_nbr == 0 || h->transport == 0 || !h->pending[step]) return 0;
	if (h->transport == 1) {
		const int spc = halo_spc(h, step);
		for (int i = 0; i < h->n_nbr; i++) {
			lfmgpu_ctx* p = h->peers[(size_t)h->nbr_rank[(size_t)i]];
			int j = -1;
			for (int k = 0; k < p->n_nbr; k++)
				if (p->nbr_rank[(size_t)k] == h->rank) j = k;
			if (j < 0) return fail("rank %d: neighbour %d does not list me", h->rank, h->nbr_rank[(size_t)i]);
			if (!p->pending[step] && p->last_send_count[step] == 0) return fail("rank %d: neighbour %d has not started halo step %d", h->rank, p->rank, step);
			const size_t rn = (size_t)(h->recv_start[(size_t)i + 1] - h->recv_start[(size_t)i]) * spc;
			const size_t sn = (size_t)(p->send_start[(size_t)j + 1] - p->send_start[(size_t)j]) * spc;
			if (rn != sn) return fail("halo size mismatch between ranks %d and %d (%zu vs %zu)", h->rank, p->rank, rn, sn);
			CU(cudaStreamWaitEvent(h->s_comm, p->ev_packed[step], 0));
			const char* src = (const char*)p->send_buf[step] + (size_t)p->send_start[(size_t)j] * spc * h->prec;
			char* dst = (char*)h->recv_buf[step] + (size_t)h->recv_start[(size_t)i] * spc * h->prec;
			if (p->device == h->device)
				CU(cudaMemcpyAsync(dst, src, rn * h->prec, cudaMemcpyDeviceToDevice, h->s_comm));
			else
				CU(cudaMemcpyPeerAsync(dst, h->device, src, p->device, rn * h->prec, h->s_comm));
		}
		CU(cudaEventRecord(h->ev_arrived[step], h->s_comm));
	}
	CU(cudaStreamWaitEvent(h->s_main, h->ev_arrived[step], 0));
	TRY(DISPATCH(h, t_unpack, h, step));
	h->pending[step] = false;
	return 0;
}

int use(lfmgpu_ctx* h) {
	if (!h) return fail("null handle");
	CU(cudaSetDevice(h->device));
	return 0;
}

// one RK stage of every listed rank, in the call order of Mesh::solve (mesh_solver.cpp:500-679)
int stage_all(lfmgpu_ctx** hs, int n, int scheme, int rk, double dt, int want_res) {
	for (int r = 0; r < n; r++) {
		lfmgpu_ctx* h = hs[r];
		TRY(use(h));
		h->rk_pending = rk;
		TRY(halo_wait_impl(h, 0));
		TRY(DISPATCH(h, t_set_bc, h));
		const bool grads = scheme == LFMGPU_SCHEME_M2AUSM && h->minmod_opt;   // mesh_solver.cpp:537-548
		if (grads) TRY(DISPATCH(h, t_gradients_ausm, h, h->n_nbr ? 0 : -1));
		TRY(DISPATCH(h, t_vis, h, h->n_nbr ? 0 : -1, h->les_opt));
		TRY(halo_start_impl(h, 1));
	}
	for (int r = 0; r < n; r++) {
		lfmgpu_ctx* h = hs[r];
		TRY(use(h));
		if (h->n_nbr) {
			if (scheme == LFMGPU_SCHEME_M2AUSM && h->minmod_opt)   // mesh_solver.cpp:582-594
				for (int s = 1; s < h->n_sub; s++) TRY(DISPATCH(h, t_gradients_ausm, h, s));
			for (int s = 1; s < h->n_sub; s++) TRY(DISPATCH(h, t_vis, h, s, h->les_opt));
		}
	}
	for (int r = 0; r < n; r++) {
		lfmgpu_ctx* h = hs[r];
		TRY(use(h));
		TRY(halo_wait_impl(h, 1));
		TRY(DISPATCH(h, t_rk_stage, h, h->n_nbr ? 0 : -1, scheme, rk, dt, want_res));
		TRY(halo_start_impl(h, 0));
	}
	for (int r = 0; r < n; r++) {
		lfmgpu_ctx* h = hs[r];
		TRY(use(h));
		if (h->n_nbr)
			for (int s = 1; s < h->n_sub; s++) TRY(DISPATCH(h, t_rk_stage, h, s, scheme, rk, dt, want_res));
	}
	return 0;
}

int steps_all(lfmgpu_ctx** hs, int n, int scheme, double dt, int n_steps, int first, int want_res);

// lfmgpu_step on a rank without neighbours: pairs of time steps are replayed from a CUDA graph (a small mesh is launch-bound:
// a 2D case of half a million cells spends 150 us per stage in five launches).  The graph is captured from the very code
// path that launches eagerly; the host-side bookkeeping that capture advances (buffer parity, flags) is the state after the
// two steps, and ten stages leave the parity where it was, so replaying needs no bookkeeping at all.
int steps_graphed(lfmgpu_ctx* h, int scheme, double dt, int n_steps, int want_res) {
	lfmgpu_ctx* hs[1] = {h};
	const bool eligible = h->use_graph && h->n_nbr == 0 && !h->timing && h->c.rk_order % 2 == 1 && n_steps >= 4;
	if (!eligible) return steps_all(hs, 1, scheme, dt, n_steps, 0, want_res);
	// the first step runs eagerly: one-off work (derived values after an upload, lazy allocations) stays out of the graph
	TRY(steps_all(hs, 1, scheme, dt, 1, 0, want_res));
	n_steps--;
	const lfmgpu_ctx::GraphKey key{scheme, want_res, h->minmod_opt, h->les_opt, h->cur, 0, dt};
	if (!h->graph_exec || !(h->graph_key == key)) {
		if (h->graph_exec) {
			cudaGraphExecDestroy(h->graph_exec);
			h->graph_exec = nullptr;
		}
		const uint64_t l0 = h->launches;
		cudaGraph_t graph = nullptr;
		CU(cudaStreamBeginCapture(h->s_main, cudaStreamCaptureModeThreadLocal));
		const int rc = steps_all(hs, 1, scheme, dt, 2, 0, want_res);
		const cudaError_t ce = cudaStreamEndCapture(h->s_main, &graph);
		if (rc || ce != cudaSuccess || !graph) {
			if (graph) cudaGraphDestroy(graph);
			return rc ? rc : fail("CUDA graph capture of two time steps failed: %s", cudaGetErrorString(ce));
		}
		const cudaError_t ci = cudaGraphInstantiate(&h->graph_exec, graph, 0);
		cudaGraphDestroy(graph);
		if (ci != cudaSuccess) return fail("cudaGraphInstantiate: %s", cudaGetErrorString(ci));
		h->graph_kernels = h->launches - l0;
		h->launches = l0;   // nothing has run yet: the launches are counted when the graph is
		h->graph_key = key;
	}
	for (; n_steps >= 2; n_steps -= 2) {
		CU(cudaGraphLaunch(h->graph_exec, h->s_main));
		h->launches += h->graph_kernels;
	}
	if (n_steps) TRY(steps_all(hs, 1, scheme, dt, n_steps, 0, want_res));
	return 0;
}

int steps_all(lfmgpu_ctx** hs, int n, int scheme, double dt, int n_steps, int first, int want_res) {
	if (first) {
		// pre-loop warm-up (mesh_solver.cpp:409-428)
		for (int r = 0; r < n; r++) {
			TRY(use(hs[r]));
			TRY(halo_start_impl(hs[r], 0));
		}
		for (int r = 0; r < n; r++) {
			TRY(use(hs[r]));
			TRY(DISPATCH(hs[r], t_set_bc, hs[r]));
			TRY(halo_wait_impl(hs[r], 0));
		}
		for (int r = 0; r < n; r++) {
			TRY(use(hs[r]));
			TRY(halo_start_impl(hs[r], 1));
		}
		for (int r = 0; r < n; r++) {
			TRY(use(hs[r]));
			TRY(halo_wait_impl(hs[r], 1));
		}
	}
	for (int s = 0; s < n_steps; s++) {
		for (int r = 0; r < n; r++) hs[r]->dq_zero = true;
		const int order = hs[0]->c.rk_order;
		for (int rk = 0; rk < order; rk++) TRY(stage_all(hs, n, scheme, rk, dt, want_res));
		for (int r = 0; r < n; r++) {
			TRY(use(hs[r]));
			TRY(halo_wait_impl(hs[r], 0));
		}
	}
	return 0;
}

// sizes and index ranges of the rank, as the descriptor gives them (no device work)
int ctx_from_desc(lfmgpu_ctx* h, const lfmgpu_desc* ds) {
	h->prec = ds->precision;
	h->D = ds->dim;
	h->NQ = ds->dim + 2;
	h->F = ds->max_slots;
	h->n_cells = ds->n_cells;
	h->n_faces = ds->n_faces;
	h->n_bc = ds->n_bc_ghosts;
	h->n_mpi = ds->n_mpi_ghosts;
	h->n_tot = ds->n_cells + ds->n_bc_ghosts + ds->n_mpi_ghosts;
	h->n_sub = ds->n_sub;
	h->ncs = pad32((size_t)h->n_tot);
	h->nfs = pad32((size_t)h->n_faces);
	h->ngs = pad32((size_t)std::max(1, h->n_mpi));
	memcpy(h->sub_cell_start, ds->sub_cell_start, sizeof h->sub_cell_start);
	memcpy(h->sub_face_start, ds->sub_face_start, sizeof h->sub_face_start);
	h->c = ds->c;
	h->n_nbr = ds->n_nbr;
	h->nbr_rank.assign(ds->nbr_rank, ds->nbr_rank + ds->n_nbr);
	h->send_start.assign(1, 0);
	h->recv_start.assign(1, 0);
	if (ds->n_nbr) {
		h->send_start.assign(ds->send_start, ds->send_start + ds->n_nbr + 1);
		h->recv_start.assign(ds->recv_start, ds->recv_start + ds->n_nbr + 1);
		if (h->recv_start[(size_t)ds->n_nbr] != ds->n_mpi_ghosts) return fail("lfmgpu_create: recv_start does not cover n_mpi_ghosts");
	}
	return 0;
}

}  // namespace

// ======================================================================================================
// C ABI
// ======================================================================================================
extern "C" {

const char* lfmgpu_last_error(void) { return g_err.c_str(); }

int lfmgpu_device_count(int* n) {
	CU(cudaGetDeviceCount(n));
	return 0;
}

int lfmgpu_create(const lfmgpu_desc* ds, int device, lfmgpu_t* out) {
	if (!ds || !out) return fail("lfmgpu_create: null argument");
	if (ds->precision != 4 && ds->precision != 8) return fail("lfmgpu_create: precision must be 4 or 8");
	if (ds->dim != 2 && ds->dim != 3) return fail("lfmgpu_create: dim must be 2 or 3");
	if (ds->n_sub < 1 || ds->n_sub > LFMGPU_MAX_SUBMESH) return fail("lfmgpu_create: bad submesh count");
	CU(cudaSetDevice(device));
	lfmgpu_ctx* h = new lfmgpu_ctx();
	h->device = device;
	if (ctx_from_desc(h, ds)) {
		delete h;
		return 1;
	}
	int lo, hi;
	cudaDeviceGetStreamPriorityRange(&lo, &hi);
	int rc = 0;
	if (cudaStreamCreateWithPriority(&h->s_main, cudaStreamNonBlocking, lo) != cudaSuccess ||
	    cudaStreamCreateWithPriority(&h->s_comm, cudaStreamNonBlocking, hi) != cudaSuccess)
		rc = fail("stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
	if (!rc) {
		cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming);
		for (int s = 0; s < 2; s++) {
			cudaEventCreateWithFlags(&h->ev_packed[s], cudaEventDisableTiming);
			cudaEventCreateWithFlags(&h->ev_arrived[s], cudaEventDisableTiming);
		}
		rc = h->prec == 8 ? build<double>(h, ds) : build<float>(h, ds);
	}
	h->stage_cfg = 6;   // tile kernels (fallback of the persistent kernel): 256 threads x 2 CTAs/SM (128 registers, no spills)
	if (const char* e = getenv("LFMGPU_TILE_CELLS")) h->tile_cells = std::max(16, atoi(e));
	if (const char* e = getenv("LFMGPU_STAGE_CFG")) h->stage_cfg = atoi(e);
	if (const char* e = getenv("LFMGPU_USE_TILES")) h->use_tiles = atoi(e);
	if (const char* e = getenv("LFMGPU_FIXED_STRIDES")) h->fixed_strides = atoi(e);
	if (const char* e = getenv("LFMGPU_TILE_SMEM")) h->tile_smem_budget = std::max(16, atoi(e)) * 1024;
	if (const char* e = getenv("LFMGPU_SMEM_PAD")) h->smem_pad_kb = std::max(0, atoi(e));
	if (const char* e = getenv("LFMGPU_GRAPH")) h->use_graph = atoi(e);
	if (const char* e = getenv("LFMGPU_PIPE")) h->pipe_enable = atoi(e);   // bit 0: stage kernel, bit 1: gradient kernel
	if (const char* e = getenv("LFMGPU_PIPE_SLOTS")) h->pipe_slots_cap = std::max(2, atoi(e));
	if (const char* e = getenv("LFMGPU_PIPE_SPARE")) h->pipe_spare_sms = std::max(0, atoi(e));
	if (const char* e = getenv("LFMGPU_PIPE_PF")) h->pipe_pf_dist = std::max(0, atoi(e));
	if (const char* e = getenv("LFMGPU_PIPE_GSLOTS")) h->pipe_grad_slots_cap = atoi(e);
	if (const char* e = getenv("LFMGPU_PIPE_GGROUPS")) h->grad_groups = atoi(e) == 4 ? 4 : 3;
	if (!rc) rc = tile_plan_build(h, ds);
	if (!rc) rc = pipe_setup(h);
	if (!(h->pipe_enable & 2)) h->grad_pipe_ok = false;
	if (!(h->pipe_enable & 1)) h->pipe_ok = false;
	if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = fail("upload failed: %s", cudaGetErrorString(cudaGetLastError()));
	if (rc) {
		lfmgpu_destroy(h);
		return rc;
	}
	*out = h;
	return 0;
}

// ---- host-only check of the tile plan (no device) ----------------------------------------------------
// Builds the host half of the plan exactly as lfmgpu_create does and verifies what the tile kernels rely on: the tiles
// partition the cells without straddling a submesh; a tile's halo is sorted, disjoint from the tile and entirely used; its
// face table holds every face owned by a tile cell (owner / neighbour staged indices and the physical-ghost flag right)
// followed by every incoming face; the local gather lists name, slot by slot in ascending mesh-face order, the table
// entry of the same mesh face with the same sign; the shared-memory strides cover every tile and fit the budget.
// stats[10]: tileable, tiles, max staged cells, max faces, incoming/own faces, halo cells per cell, stage-kernel shared
// memory in bytes, cells per tile requested after halving, mean halo cells per tile, mean runs of consecutive cell ids per
// tile's halo list (what a run-wise copy would issue per staged row instead of one copy per halo cell).
int lfmgpu_plan_check(const lfmgpu_desc* ds, int tile_cells, int smem_limit_bytes, double* stats) {
	if (!ds || !stats) return fail("lfmgpu_plan_check: null argument");
	if (ds->precision != 4 && ds->precision != 8) return fail("lfmgpu_plan_check: precision must be 4 or 8");
	if (ds->dim != 2 && ds->dim != 3) return fail("lfmgpu_plan_check: dim must be 2 or 3");
	if (ds->n_sub < 1 || ds->n_sub > LFMGPU_MAX_SUBMESH) return fail("lfmgpu_plan_check: bad submesh count");
	for (int i = 0; i < 10; i++) stats[i] = 0.0;
	lfmgpu_ctx ctx;
	lfmgpu_ctx* h = &ctx;
	TRY(ctx_from_desc(h, ds));
	if (tile_cells > 0) h->tile_cells = std::max(16, tile_cells);
	HostPlan hp;
	TRY(tile_plan_host(h, ds, smem_limit_bytes > 0 ? smem_limit_bytes : 227 * 1024, hp));
	if (!hp.ok) return 0;
	const TilePlan& p = h->tiles;
	const int nc = h->n_cells, F = h->F, D = h->D;
	const int NS = (D == 3 ? StagedLayout<3>::NS : StagedLayout<2>::NS);
	auto bad = [](const char* what, int tile) { return fail("tile plan check: %s (tile %d)", what, tile); };
	int next = 0, smax_all = 0, fmax_all = 0;
	long long own = 0, incoming = 0, halo = 0, halo_runs = 0;
	std::vector<int> row((size_t)F), used;
	for (int s = 0; s < h->n_sub; s++) {
		if (p.sub_tile_start[s] > p.sub_tile_start[s + 1]) return bad("submesh tile ranges out of order", p.sub_tile_start[s]);
		if (p.sub_smax[s] % 4 || p.sub_fmax[s] % 4) return bad("strides not multiples of 4", p.sub_tile_start[s]);
		if (((size_t)NS * p.sub_smax[s] + (size_t)h->NQ * p.sub_fmax[s]) * (size_t)h->prec > (size_t)h->tile_smem_budget) return bad("shared memory over the budget", p.sub_tile_start[s]);
		for (int t = p.sub_tile_start[s]; t < p.sub_tile_start[s + 1]; t++) {
			const TileDesc& td = hp.tiles[(size_t)t];
			const int c0 = td.c0, c1 = td.c0 + td.nt;
			if (c0 != next || td.nt < 1) return bad("tiles do not tile the cells", t);
			next = c1;
			if (c0 < h->sub_cell_start[s] || c1 > h->sub_cell_start[s + 1]) return bad("tile straddles a submesh", t);
			if (td.hb != p.sub_tc[s] || td.nt > td.hb) return bad("halo base is not the submesh's tile size", t);
			if (td.hb + td.nh > p.sub_smax[s] || td.nfo + td.ninc > p.sub_fmax[s]) return bad("tile larger than the strides of its launch", t);
			const int* hc = hp.halo_cell.data() + td.halo_off;
			for (int i = 0; i < td.nh; i++) {
				if (i && hc[i] <= hc[i - 1]) return bad("halo not strictly ascending", t);
				if (hc[i] >= c0 && hc[i] < c1) return bad("halo cell inside the tile", t);
				if (hc[i] < 0 || hc[i] >= h->n_tot) return bad("halo cell out of range", t);
				if (!i || hc[i] != hc[i - 1] + 1) halo_runs++;
			}
			used.assign((size_t)td.nh, 0);
			auto cell_of = [&](int staged) { return staged < td.nt ? c0 + staged : hc[staged - td.hb]; };
			int n_own = 0;
			for (int c = c0; c < c1; c++)
				for (int k = 0; k < F; k++) n_own += ds->cell_slot_face[(size_t)c * F + k] > 0 ? 1 : 0;
			if (n_own != td.nfo) return bad("own-face count differs from the faces the tile's cells own", t);
			for (int j = 0; j < td.nfo + td.ninc; j++) {
				const int f = hp.f_gface[(size_t)td.f_off + j];
				const uint32_t idx = hp.f_idx[(size_t)td.f_off + j];
				const int lo = (int)(idx & 0xffffu), ln = (int)((idx >> 16) & 0x7fffu);
				const bool ghost = (idx >> 31) != 0;
				auto staged_ok = [&](int x) { return (x >= 0 && x < td.nt) || (x >= td.hb && x < td.hb + td.nh); };
				if (f < 0 || f >= h->n_faces || !staged_ok(lo) || !staged_ok(ln)) return bad("face table entry out of range", t);
				if (cell_of(lo) != ds->face_owner[f] || cell_of(ln) != ds->face_neigh[f]) return bad("staged owner/neighbour of a table face is not the mesh's", t);
				const int n = ds->face_neigh[f];
				if (ghost != (n >= nc && n < nc + h->n_bc)) return bad("physical-ghost flag wrong", t);
				if (j < td.nfo) {
					if (lo >= td.nt) return bad("own face whose owner is outside the tile", t);
				} else {
					if (lo < td.nt || ln >= td.nt || ghost) return bad("incoming face not from a halo cell into the tile", t);
				}
				if (lo >= td.nt) used[(size_t)(lo - td.hb)] = 1;
				if (ln >= td.nt) used[(size_t)(ln - td.hb)] = 1;
			}
			for (int i = 0; i < td.nh; i++)
				if (!used[(size_t)i]) return bad("halo cell no face refers to", t);
			int n_inc = 0;
			for (int c = c0; c < c1; c++) {
				int n = 0;
				for (int k = 0; k < F; k++) {
					const int e = ds->cell_slot_face[(size_t)c * F + k];
					if (e) row[(size_t)n++] = e;
				}
				std::sort(row.begin(), row.begin() + n, [](int a, int b) { return std::abs(a) < std::abs(b); });
				for (int k = 0; k < F; k++) {
					const int l = hp.csr_local[(size_t)k * nc + c];
					if (k >= n) {
						if (l) return bad("gather list longer than the cell's face list", t);
						continue;
					}
					const int e = row[(size_t)k], lf = std::abs(l) - 1;
					if (!l || (l > 0) != (e > 0) || lf >= td.nfo + td.ninc) return bad("gather entry missing, of the wrong sign or out of range", t);
					if (hp.f_gface[(size_t)td.f_off + lf] != std::abs(e) - 1) return bad("gather entry names another mesh face", t);
					if (e < 0) {
						const int o = ds->face_owner[std::abs(e) - 1];
						if (o < c0 || o >= c1) {
							n_inc++;
							if (lf < td.nfo) return bad("incoming face filed among the own faces", t);
						} else if (lf >= td.nfo) {
							return bad("own face filed among the incoming faces", t);
						}
					}
				}
			}
			if (n_inc != td.ninc) return bad("incoming-face count differs from the gather lists", t);
			smax_all = std::max(smax_all, td.hb + td.nh);
			fmax_all = std::max(fmax_all, td.nfo + td.ninc);
			own += td.nfo;
			incoming += td.ninc;
			halo += td.nh;
		}
	}
	if (next != nc || p.sub_tile_start[h->n_sub] != (int)hp.tiles.size()) return bad("tiles do not cover the cells", (int)hp.tiles.size());
	int smax = 0, fmax = 0;
	for (int s = 0; s < h->n_sub; s++) {
		smax = std::max(smax, p.sub_smax[s]);
		fmax = std::max(fmax, p.sub_fmax[s]);
	}
	stats[0] = 1.0;
	stats[1] = (double)hp.tiles.size();
	stats[2] = smax_all;
	stats[3] = fmax_all;
	stats[4] = own ? (double)incoming / (double)own : 0.0;
	stats[5] = (double)halo / (double)nc;
	stats[6] = (double)(((size_t)NS * smax + (size_t)h->NQ * fmax) * (size_t)h->prec);
	stats[7] = hp.TCs[h->n_sub - 1];
	stats[8] = (double)halo / (double)hp.tiles.size();
	stats[9] = (double)halo_runs / (double)hp.tiles.size();
	return 0;
}

int lfmgpu_destroy(lfmgpu_t h) {
	if (!h) return 0;
	cudaSetDevice(h->device);
	cudaDeviceSynchronize();
	if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
	if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
	for (auto& t : h->timed) {
		cudaEventDestroy(t.a);
		cudaEventDestroy(t.b);
	}
	for (void* p : h->allocs) cudaFree(p);
	if (h->ev_ready) cudaEventDestroy(h->ev_ready);
	for (int s = 0; s < 8; s++)
		if (h->ev_user[s]) cudaEventDestroy(h->ev_user[s]);
	for (int s = 0; s < 2; s++) {
		if (h->ev_packed[s]) cudaEventDestroy(h->ev_packed[s]);
		if (h->ev_arrived[s]) cudaEventDestroy(h->ev_arrived[s]);
	}
	if (h->s_in) cudaStreamDestroy(h->s_in);
	if (h->s_out) cudaStreamDestroy(h->s_out);
	for (cudaEvent_t e : {h->ev_in_ready, h->ev_in_free, h->ev_out_ready, h->ev_out_free})
		if (e) cudaEventDestroy(e);
	if (h->s_main) cudaStreamDestroy(h->s_main);
	if (h->s_comm) cudaStreamDestroy(h->s_comm);
	delete h;
	return 0;
}

int lfmgpu_sync(lfmgpu_t h) {
	TRY(use(h));
	if (h->s_in) CU(cudaStreamSynchronize(h->s_in));
	if (h->s_out) CU(cudaStreamSynchronize(h->s_out));
	CU(cudaStreamSynchronize(h->s_comm));
	CU(cudaStreamSynchronize(h->s_main));
	return 0;
}

int lfmgpu_set_option(lfmgpu_t h, const char* name, int value) {
	TRY(use(h));
	if (!strcmp(name, "laminar")) {   // 0: the time loops of lfmgpu_step / _step_multi call calc_VIS_Smagorinsky
		h->les_opt = value ? 0 : 1;
		return 0;
	}
	if (!strcmp(name, "minmod")) {   // fvSchemes lfm/minmodExists: the time loops call calc_gradients_M2AUSM (solver 2 only)
		h->minmod_opt = value ? 1 : 0;
		return 0;
	}
	if (!strcmp(name, "use_tiles")) {
		h->use_tiles = value;
		h->drv_valid[0] = h->drv_valid[1] = false;
		return 0;
	}
	return fail("unknown option %s", name);
}

int lfmgpu_prepare_timestep(lfmgpu_t h) {
	TRY(use(h));
	h->dq_zero = true;
	return 0;
}
int lfmgpu_prepare_rkstep(lfmgpu_t h, int rk_step) {
	TRY(use(h));
	if (rk_step < 0 || rk_step >= LFMGPU_MAX_RK) return fail("rk_step out of range");
	h->rk_pending = rk_step;
	return 0;
}
int lfmgpu_set_bc(lfmgpu_t h) {
	TRY(use(h));
	return DISPATCH(h, t_set_bc, h);
}
int lfmgpu_gradients(lfmgpu_t h, int submesh) {
	// calc_gradients feeds only one_rk_step_M2AUSM: for the M1/M2 schemes served here its
	// outputs are never read, so the call is accepted and does no work.
	TRY(use(h));
	(void)submesh;
	return 0;
}
int lfmgpu_gradients_m2ausm(lfmgpu_t h, int submesh) {
	TRY(use(h));
	if (submesh >= h->n_sub) return fail("submesh out of range");
	return DISPATCH(h, t_gradients_ausm, h, submesh);
}
int lfmgpu_vis(lfmgpu_t h, int submesh) {
	TRY(use(h));
	if (submesh >= h->n_sub) return fail("submesh out of range");
	return DISPATCH(h, t_vis, h, submesh, 0);
}
int lfmgpu_vis_smagorinsky(lfmgpu_t h, int submesh) {
	TRY(use(h));
	if (submesh >= h->n_sub) return fail("submesh out of range");
	return DISPATCH(h, t_vis, h, submesh, 1);
}
int lfmgpu_rk_stage(lfmgpu_t h, int submesh, int scheme, int rk_step, double dt, int want_res) {
	TRY(use(h));
	if (submesh >= h->n_sub) return fail("submesh out of range");
	if (scheme != LFMGPU_SCHEME_M1 && scheme != LFMGPU_SCHEME_M2 && scheme != LFMGPU_SCHEME_M2AUSM) return fail("scheme %d is not served by the GPU path", scheme);
	if (rk_step < 0 || rk_step >= LFMGPU_MAX_RK) return fail("rk_step out of range");
	return DISPATCH(h, t_rk_stage, h, submesh, scheme, rk_step, dt, want_res);
}
int lfmgpu_halo_start(lfmgpu_t h, int comm_step) {
	TRY(use(h));
	if (comm_step != 0 && comm_step != 1) return fail("comm_step must be 0 or 1");
	return halo_start_impl(h, comm_step);
}
int lfmgpu_halo_wait(lfmgpu_t h, int comm_step) {
	TRY(use(h));
	if (comm_step != 0 && comm_step != 1) return fail("comm_step must be 0 or 1");
	return halo_wait_impl(h, comm_step);
}
int lfmgpu_cfl(lfmgpu_t h, double dt, double* cfl_max) {
	TRY(use(h));
	return DISPATCH(h, t_cfl_dt, h, dt, 0, cfl_max);
}
int lfmgpu_dt(lfmgpu_t h, double cfl_max, double* dt_min) {
	TRY(use(h));
	return DISPATCH(h, t_cfl_dt, h, cfl_max, 1, dt_min);
}
int lfmgpu_average(lfmgpu_t h, int time_step) {
	TRY(use(h));
	return DISPATCH(h, t_average, h, time_step);
}
int lfmgpu_forces(lfmgpu_t h, int patch, double* Fpre, double* Fvis) {
	TRY(use(h));
	return DISPATCH(h, t_forces, h, patch, Fpre, Fvis);
}
int lfmgpu_residual(lfmgpu_t h, double* res) {
	TRY(use(h));
	CU(cudaMemcpyAsync(res, h->d_res, (size_t)h->NQ * sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
	CU(cudaMemsetAsync(h->d_res, 0, 8 * sizeof(double), h->s_main));
	CU(cudaStreamSynchronize(h->s_main));
	return 0;
}

int lfmgpu_warmup(lfmgpu_t h) {
	TRY(use(h));
	if (h->transport == 1 && h->n_ranks > 1) return fail("lfmgpu_warmup: in-process ranks must be driven together (lfmgpu_step_multi)");
	lfmgpu_ctx* hs[1] = {h};
	return steps_all(hs, 1, 0, 0.0, 0, 1, 0);
}

int lfmgpu_step(lfmgpu_t h, int scheme, double dt, int n_steps, int minmod, int want_res) {
	TRY(use(h));
	h->minmod_opt = minmod ? 1 : 0;
	if (scheme != LFMGPU_SCHEME_M1 && scheme != LFMGPU_SCHEME_M2 && scheme != LFMGPU_SCHEME_M2AUSM) return fail("scheme %d is not served by the GPU path", scheme);
	if (h->transport == 1 && h->n_ranks > 1) return fail("lfmgpu_step: in-process ranks must be driven together (lfmgpu_step_multi)");
	if (h->n_nbr && h->transport == 0) return fail("lfmgpu_step: rank has neighbours but no halo transport was initialised");
	return steps_graphed(h, scheme, dt, n_steps, want_res);
}

int lfmgpu_step_multi(const lfmgpu_t* hs, int n_ranks, int scheme, double dt, int n_steps, int first, int want_res) {
	if (!hs || n_ranks < 1) return fail("lfmgpu_step_multi: bad arguments");
	if (scheme != LFMGPU_SCHEME_M1 && scheme != LFMGPU_SCHEME_M2 && scheme != LFMGPU_SCHEME_M2AUSM) return fail("scheme %d is not served by the GPU path", scheme);
	std::vector<lfmgpu_ctx*> v(hs, hs + n_ranks);
	for (auto* h : v)
		if (h->n_nbr && h->transport == 0) return fail("lfmgpu_step_multi: rank has neighbours but no halo transport");
	return steps_all(v.data(), n_ranks, scheme, dt, n_steps, first, want_res);
}

// ---- data movement ---------------------------------------------------------------------------------
}  // extern "C" (templates below)
namespace {
// dense device scratch of the data-movement entry points
struct DevTmp {
	void* p = nullptr;
	~DevTmp() {
		if (p) cudaFree(p);
	}
};
template <class R, int D> int t_download(lfmgpu_ctx* h, int field, void* dst, size_t dst_bytes) {
	const int NQ = h->NQ, nc = h->n_cells;
	constexpr int QW = Rec<D>::QW, VW = Rec<D>::VW;
	DevMesh<R>& m = h->mesh<R>();
	// record fields: (records, width, first value, values); SoA fields: (base, stride)
	const R* rec = nullptr;
	int W = 0, comp0 = 0, comps = 1;
	const R* soa = nullptr;
	size_t stride = 0, n = (size_t)nc, off = 0;
	bool taumc_laminar = false;
	switch (field) {
		case LFMGPU_FIELD_Q: rec = (const R*)h->q[h->cur]; W = QW; comps = NQ; break;
		case LFMGPU_FIELD_QGHOST: rec = (const R*)h->q[h->cur]; W = QW; comps = NQ; off = (size_t)nc; n = (size_t)(h->n_bc + h->n_mpi); break;
		case LFMGPU_FIELD_DUDX: rec = m.vis; W = VW; comp0 = Rec<D>::DUDX; comps = D * D; break;
		case LFMGPU_FIELD_DTDX: rec = m.vis; W = VW; comp0 = Rec<D>::DTDX; comps = D; break;
		case LFMGPU_FIELD_SIGMAU: rec = m.vis; W = VW; comp0 = Rec<D>::SIGMAU; comps = D; break;
		case LFMGPU_FIELD_DQ: soa = m.dq; stride = (size_t)nc; comps = NQ; break;
		case LFMGPU_FIELD_RES: soa = m.RES; stride = (size_t)nc; comps = NQ; break;
		case LFMGPU_FIELD_PAVG: soa = m.pAVG; stride = (size_t)nc; break;
		case LFMGPU_FIELD_PRMS: soa = m.pRMS; stride = (size_t)nc; break;
		case LFMGPU_FIELD_TAUMC:
			comps = D * D;
			if (h->les) {
				soa = m.tauMC;
				stride = h->ncs;
			} else {
				taumc_laminar = true;   // a function of the stored dudx (calc_VIS does not keep it): rebuilt on the device
			}
			break;
		default: return fail("unknown field %d", field);
	}
	const size_t bytes = n * comps * sizeof(R);
	if (dst_bytes < bytes) return fail("lfmgpu_download: destination too small (%zu < %zu)", dst_bytes, bytes);
	if (n == 0) return 0;
	if (soa) {
		std::vector<R> tmp((size_t)comps * n);
		for (int i = 0; i < comps; i++) CU(cudaMemcpy(tmp.data() + (size_t)i * n, soa + (size_t)i * stride, n * sizeof(R), cudaMemcpyDeviceToHost));
		R* o = (R*)dst;
		for (size_t c = 0; c < n; c++)
			for (int i = 0; i < comps; i++) o[c * comps + i] = tmp[(size_t)i * n + c];
		return 0;
	}
	DevTmp t;
	CU(cudaMalloc(&t.p, bytes));
	const unsigned nb = (unsigned)((n + kBlock - 1) / kBlock);
	if (taumc_laminar) {
		k_taumc_laminar<R, D><<<nb, kBlock, 0, h->s_main>>>(m, n, (R*)t.p);
	} else {
		k_rec_gather<R><<<nb, kBlock, 0, h->s_main>>>(rec, W, comp0, comps, off, n, (R*)t.p);
		if (field == LFMGPU_FIELD_QGHOST && h->n_bc && h->stage_done) {
			// physical ghosts were last written by set_boundary_conditions of the latest stage, i.e. into the other buffer
			k_rec_gather<R><<<(unsigned)((h->n_bc + kBlock - 1) / kBlock), kBlock, 0, h->s_main>>>((const R*)h->q[1 - h->cur], W, comp0, comps, off, (size_t)h->n_bc, (R*)t.p);
		}
	}
	CHECK_LAUNCH();
	CU(cudaMemcpyAsync(dst, t.p, bytes, cudaMemcpyDeviceToHost, h->s_main));
	CU(cudaStreamSynchronize(h->s_main));
	return 0;
}

// conservatives of the real cells from a dense device array (AoS [n][NQ] or component-major [NQ][n]) into the records of q[cur]
template <class R, int D> int t_scatter_q(lfmgpu_ctx* h, const void* dev_src, int aos, cudaStream_t s) {
	const size_t n = (size_t)h->n_cells;
	if (!n) return 0;
	k_rec_scatter<R><<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, s>>>((const R*)dev_src, aos, n, h->NQ, (R*)h->q[h->cur], Rec<D>::QW);
	CHECK_LAUNCH();
	return 0;
}
template <class R, int D> int t_gather_q_soa(lfmgpu_ctx* h, void* dev_dst, cudaStream_t s) {
	const size_t n = (size_t)h->n_cells;
	if (!n) return 0;
	k_rec_to_soa<R><<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, s>>>((const R*)h->q[h->cur], Rec<D>::QW, n, h->NQ, (R*)dev_dst);
	CHECK_LAUNCH();
	return 0;
}
// a new state was uploaded into q[cur]: nothing of the previous run describes it any more
void mark_uploaded(lfmgpu_ctx* h) {
	h->stage_done = false;
	h->vis_on_cur = true;
	h->drv_valid[h->cur] = false;
}
int pipe_init(lfmgpu_ctx* h) {
	if (h->s_in) return 0;
	const size_t bytes = (size_t)h->NQ * h->n_cells * (size_t)h->prec;
	CU(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
	TRY(dev_alloc(h, &h->stage_in, bytes, false));
	TRY(dev_alloc(h, &h->stage_out, bytes, false));
	cudaEvent_t* evs[4] = {&h->ev_in_ready, &h->ev_in_free, &h->ev_out_ready, &h->ev_out_free};
	for (auto* e : evs) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
	CU(cudaEventRecord(h->ev_in_free, h->s_main));
	CU(cudaEventRecord(h->ev_in_ready, h->s_in));
	CU(cudaEventRecord(h->ev_out_free, h->s_out));
	CU(cudaEventRecord(h->ev_out_ready, h->s_main));
	return 0;
}
}  // namespace
extern "C" {

int lfmgpu_download(lfmgpu_t h, int field, void* dst, size_t dst_bytes) {
	TRY(use(h));
	TRY(lfmgpu_sync(h));
	return DISPATCH(h, t_download, h, field, dst, dst_bytes);
}

int lfmgpu_upload_q(lfmgpu_t h, const void* q, size_t bytes) {
	TRY(use(h));
	TRY(lfmgpu_sync(h));
	const size_t es = (size_t)h->prec, n = (size_t)h->n_cells;
	if (bytes != n * h->NQ * es) return fail("lfmgpu_upload_q: expected %zu bytes", n * h->NQ * es);
	DevTmp t;
	CU(cudaMalloc(&t.p, std::max<size_t>(bytes, 16)));
	CU(cudaMemcpyAsync(t.p, q, bytes, cudaMemcpyHostToDevice, h->s_main));
	TRY(DISPATCH(h, t_scatter_q, h, t.p, 1, h->s_main));
	CU(cudaStreamSynchronize(h->s_main));
	mark_uploaded(h);
	return 0;
}

// Component-major variants for the end-to-end path: host arrays are [NQ][n_cells], pinned or not; asynchronous on the
// compute stream (through the device staging buffers of the pipelined path: the records are filled by a kernel).
int lfmgpu_upload_q_soa_async(lfmgpu_t h, const void* q, size_t bytes) {
	TRY(use(h));
	TRY(pipe_init(h));
	const size_t es = (size_t)h->prec, n = (size_t)h->n_cells;
	if (bytes != n * h->NQ * es) return fail("lfmgpu_upload_q_soa_async: expected %zu bytes", n * h->NQ * es);
	CU(cudaStreamWaitEvent(h->s_main, h->ev_in_ready, 0));   // a pipelined upload still filling the staging buffer
	CU(cudaMemcpyAsync(h->stage_in, q, bytes, cudaMemcpyHostToDevice, h->s_main));
	TRY(DISPATCH(h, t_scatter_q, h, h->stage_in, 0, h->s_main));
	CU(cudaEventRecord(h->ev_in_free, h->s_main));
	mark_uploaded(h);
	return 0;
}
int lfmgpu_download_q_soa_async(lfmgpu_t h, void* q, size_t bytes) {
	TRY(use(h));
	TRY(pipe_init(h));
	const size_t es = (size_t)h->prec, n = (size_t)h->n_cells;
	if (bytes != n * h->NQ * es) return fail("lfmgpu_download_q_soa_async: expected %zu bytes", n * h->NQ * es);
	CU(cudaStreamWaitEvent(h->s_main, h->ev_out_free, 0));
	TRY(DISPATCH(h, t_gather_q_soa, h, h->stage_out, h->s_main));
	CU(cudaMemcpyAsync(q, h->stage_out, bytes, cudaMemcpyDeviceToHost, h->s_main));
	CU(cudaEventRecord(h->ev_out_ready, h->s_main));
	return 0;
}

// Pipelined host I/O for back-to-back batches: the upload of the next batch and the download of the previous result run
// on their own streams, through device staging buffers, while the compute stream advances the current batch.
//   pipe_in_start(host)  : H2D host -> staging (copy-in stream), once the previous commit has drained the staging buffer
//   pipe_in_commit()     : compute stream waits for that upload, then staging -> q (device copy)
//   pipe_out_start()     : compute stream copies q -> out staging, once the previous fetch has drained it
//   pipe_out_fetch(host) : D2H out staging -> host (copy-out stream)
int lfmgpu_pipe_in_start(lfmgpu_t h, const void* q, size_t bytes) {
	TRY(use(h));
	TRY(pipe_init(h));
	const size_t es = (size_t)h->prec, n = (size_t)h->n_cells;
	if (bytes != n * h->NQ * es) return fail("lfmgpu_pipe_in_start: expected %zu bytes", n * h->NQ * es);
	CU(cudaStreamWaitEvent(h->s_in, h->ev_in_free, 0));
	CU(cudaMemcpyAsync(h->stage_in, q, bytes, cudaMemcpyHostToDevice, h->s_in));
	CU(cudaEventRecord(h->ev_in_ready, h->s_in));
	return 0;
}
int lfmgpu_pipe_in_commit(lfmgpu_t h) {
	TRY(use(h));
	TRY(pipe_init(h));
	CU(cudaStreamWaitEvent(h->s_main, h->ev_in_ready, 0));
	TRY(DISPATCH(h, t_scatter_q, h, h->stage_in, 0, h->s_main));
	CU(cudaEventRecord(h->ev_in_free, h->s_main));
	mark_uploaded(h);
	return 0;
}
int lfmgpu_pipe_out_start(lfmgpu_t h) {
	TRY(use(h));
	TRY(pipe_init(h));
	CU(cudaStreamWaitEvent(h->s_main, h->ev_out_free, 0));
	TRY(DISPATCH(h, t_gather_q_soa, h, h->stage_out, h->s_main));
	CU(cudaEventRecord(h->ev_out_ready, h->s_main));
	return 0;
}
int lfmgpu_pipe_out_fetch(lfmgpu_t h, void* q, size_t bytes) {
	TRY(use(h));
	TRY(pipe_init(h));
	const size_t es = (size_t)h->prec, n = (size_t)h->n_cells;
	if (bytes != n * h->NQ * es) return fail("lfmgpu_pipe_out_fetch: expected %zu bytes", n * h->NQ * es);
	CU(cudaStreamWaitEvent(h->s_out, h->ev_out_ready, 0));
	CU(cudaMemcpyAsync(q, h->stage_out, bytes, cudaMemcpyDeviceToHost, h->s_out));
	CU(cudaEventRecord(h->ev_out_free, h->s_out));
	return 0;
}

int lfmgpu_host_alloc(void** p, size_t bytes) {
	CU(cudaMallocHost(p, bytes));
	return 0;
}
int lfmgpu_host_free(void* p) {
	CU(cudaFreeHost(p));
	return 0;
}

// ---- halo transport --------------------------------------------------------------------------------
int lfmgpu_nccl_unique_id(void* id128) {
	TRY(load_nccl());
	ncclUniqueId id;
	NC(g_nccl.GetUniqueId(&id));
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	memcpy(id128, &id, 128);
	return 0;
}
int lfmgpu_comm_init_nccl(lfmgpu_t h, const void* id128, int rank, int n_ranks) {
	TRY(use(h));
	TRY(load_nccl());
	ncclUniqueId id;
	memcpy(&id, id128, 128);
	NC(g_nccl.CommInitRank(&h->comm, n_ranks, id, rank));
	h->rank = rank;
	h->n_ranks = n_ranks;
	h->transport = 2;
	return 0;
}
int lfmgpu_comm_init_local(lfmgpu_t h, int rank, int n_ranks, const lfmgpu_t* peers) {
	TRY(use(h));
	if (rank < 0 || rank >= n_ranks) return fail("bad rank");
	h->rank = rank;
	h->n_ranks = n_ranks;
	h->peers.assign(peers, peers + n_ranks);
	for (int i = 0; i < h->n_nbr; i++) {
		const int r = h->nbr_rank[(size_t)i];
		if (r < 0 || r >= n_ranks || !h->peers[(size_t)r]) return fail("neighbour rank %d has no handle", r);
		lfmgpu_ctx* p = h->peers[(size_t)r];
		if (p->device != h->device) {
			int can = 0;
			CU(cudaDeviceCanAccessPeer(&can, h->device, p->device));
			if (can) {
				cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
				cudaGetLastError();
			}
		}
	}
	h->transport = 1;
	return 0;
}
int lfmgpu_allreduce(lfmgpu_t h, double* values, int n, int op) {
	// op: 0 sum, 1 min, 2 max -- the per-step scalar reductions of Mesh::solve (mesh_solver.cpp:715, 763-779)
	TRY(use(h));
	if (h->transport != 2) return 0;   // single rank, or in-process ranks reduced by the caller
	if (n > 8) return fail("lfmgpu_allreduce: at most 8 values");
	CU(cudaMemcpyAsync(h->d_scalar, values, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->s_main));
	NC(g_nccl.AllReduce(h->d_scalar, h->d_scalar, (size_t)n, ncclDouble, op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax), h->comm, h->s_main));
	CU(cudaMemcpyAsync(values, h->d_scalar, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
	CU(cudaStreamSynchronize(h->s_main));
	return 0;
}
int lfmgpu_halo_send_count(lfmgpu_t h, int comm_step, size_t* n_scalars) {
	if (!h || comm_step < 0 || comm_step > 1) return fail("bad argument");
	*n_scalars = h->last_send_count[comm_step];
	return 0;
}
int lfmgpu_download_send_buffer(lfmgpu_t h, int comm_step, void* dst, size_t dst_bytes) {
	TRY(use(h));
	if (comm_step < 0 || comm_step > 1) return fail("bad argument");
	const size_t bytes = h->last_send_count[comm_step] * (size_t)h->prec;
	if (dst_bytes < bytes) return fail("destination too small");
	CU(cudaStreamSynchronize(h->s_comm));
	if (bytes) CU(cudaMemcpy(dst, h->send_buf[comm_step], bytes, cudaMemcpyDeviceToHost));
	return 0;
}

// Host-staged halo (more ranks than GPUs, or a caller that keeps its own transport such as the reference's MPI_env):
// pack on the device and hand the packed buffer to the host / take a received buffer from the host and unpack it.
int lfmgpu_halo_pack_to_host(lfmgpu_t h, int comm_step, void* dst, size_t dst_bytes) {
	TRY(use(h));
	if (comm_step < 0 || comm_step > 1) return fail("comm_step must be 0 or 1");
	if (h->n_nbr == 0) return 0;
	CU(cudaEventRecord(h->ev_ready, h->s_main));
	CU(cudaStreamWaitEvent(h->s_comm, h->ev_ready, 0));
	TRY(DISPATCH(h, t_pack, h, comm_step, h->s_comm));
	const size_t bytes = h->last_send_count[comm_step] * (size_t)h->prec;
	if (dst_bytes < bytes) return fail("lfmgpu_halo_pack_to_host: destination too small (%zu < %zu)", dst_bytes, bytes);
	if (bytes) CU(cudaMemcpyAsync(dst, h->send_buf[comm_step], bytes, cudaMemcpyDeviceToHost, h->s_comm));
	CU(cudaStreamSynchronize(h->s_comm));
	return 0;
}
int lfmgpu_halo_unpack_from_host(lfmgpu_t h, int comm_step, const void* src, size_t bytes) {
	TRY(use(h));
	if (comm_step < 0 || comm_step > 1) return fail("comm_step must be 0 or 1");
	if (h->n_nbr == 0) return 0;
	const size_t want = (size_t)h->recv_start[(size_t)h->n_nbr] * halo_spc(h, comm_step) * (size_t)h->prec;
	if (bytes < want) return fail("lfmgpu_halo_unpack_from_host: source too small (%zu < %zu)", bytes, want);
	if (want) CU(cudaMemcpyAsync(h->recv_buf[comm_step], src, want, cudaMemcpyHostToDevice, h->s_main));
	TRY(DISPATCH(h, t_unpack, h, comm_step));
	CU(cudaStreamSynchronize(h->s_main));   // the caller may reuse its buffer as soon as this returns
	return 0;
}

// ---- introspection ---------------------------------------------------------------------------------
int lfmgpu_launch_count(lfmgpu_t h, uint64_t* n) {
	if (!h) return fail("null handle");
	*n = h->launches;
	return 0;
}
int lfmgpu_enable_kernel_timing(lfmgpu_t h, int on) {
	TRY(use(h));
	TRY(lfmgpu_sync(h));
	for (auto& t : h->timed) {
		cudaEventDestroy(t.a);
		cudaEventDestroy(t.b);
	}
	h->timed.clear();
	h->timing = on != 0;
	return 0;
}
int lfmgpu_kernel_time(lfmgpu_t h, const char* prefix, double* total_ms, uint64_t* launches) {
	TRY(use(h));
	TRY(lfmgpu_sync(h));
	double tot = 0;
	uint64_t n = 0;
	const size_t pl = strlen(prefix);
	for (auto& t : h->timed) {
		if (strncmp(t.name, prefix, pl)) continue;
		float ms = 0;
		CU(cudaEventElapsedTime(&ms, t.a, t.b));
		tot += ms;
		n++;
	}
	*total_ms = tot;
	*launches = n;
	return 0;
}
int lfmgpu_event_record(lfmgpu_t h, int slot) {
	TRY(use(h));
	if (slot < 0 || slot >= 8) return fail("event slot out of range");
	if (!h->ev_user[slot]) CU(cudaEventCreate(&h->ev_user[slot]));
	// the compute stream waits for the halo stream first so that the event closes everything enqueued so far
	CU(cudaEventRecord(h->ev_ready, h->s_comm));
	CU(cudaStreamWaitEvent(h->s_main, h->ev_ready, 0));
	if (h->s_in) {   // pipelined host I/O in flight: the event closes the copy streams too
		CU(cudaEventRecord(h->ev_ready, h->s_in));
		CU(cudaStreamWaitEvent(h->s_main, h->ev_ready, 0));
		CU(cudaEventRecord(h->ev_ready, h->s_out));
		CU(cudaStreamWaitEvent(h->s_main, h->ev_ready, 0));
	}
	CU(cudaEventRecord(h->ev_user[slot], h->s_main));
	return 0;
}
int lfmgpu_event_elapsed_ms(lfmgpu_t h, int slot_a, int slot_b, double* ms) {
	TRY(use(h));
	if (slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8 || !h->ev_user[slot_a] || !h->ev_user[slot_b]) return fail("event slot not recorded");
	CU(cudaEventSynchronize(h->ev_user[slot_b]));
	float f = 0;
	CU(cudaEventElapsedTime(&f, h->ev_user[slot_a], h->ev_user[slot_b]));
	*ms = (double)f;
	return 0;
}
int lfmgpu_tile_info(lfmgpu_t h, int* n_tiles, int* tile_cells, size_t* smem_bytes, double* halo_face_ratio) {
	if (!h) return fail("null handle");
	*n_tiles = h->tiles.ready ? h->tiles.n_tiles : 0;
	*tile_cells = h->tiles.tile_cells;
	*smem_bytes = h->tiles.smem_bytes;
	*halo_face_ratio = h->tiles.halo_face_ratio;
	return 0;
}

}  // extern "C"
