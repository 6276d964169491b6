// CFDv0_solver_gpu -- the reference-side binding of the B200 hot path.
//
// A drop-in subclass of the reference's `CFDv0_solver<PRECISION, DIM_CNT, FACE_CNT>` (api/cfdv0_solver.h:114-260).
// Everything the reference does before the time loop (allocate, allocate_ghost_cells, assign_pointers, initialize,
// init_params, reorder_faces; src/cfd_v0.cpp:28-1005) runs UNCHANGED in the base class.  After reorder_faces() the
// pointer-linked AoS state of all submesh solvers of the rank is flattened into one `lfmgpu_desc`
// (include/lfmgpu.h) and uploaded; from then on every per-iteration virtual that `Mesh::solve`
// (src/mesh_solver.cpp:474-853) calls is forwarded to the C ABI of liblfmgpu.so.  Output hooks download the
// fields back into the base class's cells and call the base implementation.
//
// Selecting it: the only place the reference names the concrete solver is the switch in
// `Mesh::initializeSolver` (src/mesh_solver.cpp:56-75): `new CFDv0_solver<...>` -> `new CFDv0_solver_gpu<...>`
// (INTEGRATION.md shows the patch; the drop-in test build does the same substitution with `-include
// gpu_seam_select.h`, leaving the reference sources untouched).
//
// Halo transport (replaces the bodies of mpi_communication / mpi_wait, src/cfd_v0.cpp:3547-3603):
//   * "nccl": one rank per GPU; packed buffers go device-to-device with ncclSend/ncclRecv on a side stream
//     (lfmgpu_halo_start / lfmgpu_halo_wait); the ncclUniqueId is distributed with MPI_Allgather;
//   * "host": more ranks than GPUs (or LFM_GPU_HALO=host): the device packs, the packed buffer lands in the
//     reference's own m_SendBuf*List and travels through the reference's own MPI_env calls, unchanged; the
//     received buffer is uploaded and unpacked on the device.  This keeps every haloCommType working.
//
// Error convention of the reference: print on the failing rank and MPI_Abort(MPI_COMM_WORLD, code); code 700 here.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "api/cfdv0_solver.h"
#include "lfmgpu.h"

namespace lfmgpu_seam {

// F-erased view of one submesh solver (the submeshes of a rank may have different FACE_CNT)
template <class P, unsigned D> struct SubView {
	virtual ~SubView() {}
	virtual int subIndex() const = 0;
	virtual int nCells() const = 0;
	virtual int faceCnt() const = 0;
	virtual int nValid(int c) const = 0;
	virtual size_t cellStride() const = 0;
	virtual t_solution_vars<P, D>* vars(int c) = 0;
	virtual const t_solution_vars<P, D>* neigh(int c, int s) const = 0;
	virtual const P* S(int c, int s) const = 0;
	virtual const P* d(int c, int s) const = 0;
	virtual P w(int c, int s) const = 0;
	virtual P sigma(int c) const = 0;
	virtual int cellIndex(int c) const = 0;
};

// one per process (one MPI rank == one process == one GPU context)
template <class P, unsigned D> struct RankState {
	std::vector<SubView<P, D>*> subs;   // by submesh index
	lfmgpu_t h = nullptr;
	bool built = false;
	int n_cells = 0, n_bc = 0, n_mpi = 0;
	std::vector<int> sub_start;
	bool use_nccl = false;
	int res_pending = 0;
	static RankState& get() {
		static RankState s;
		return s;
	}
};

inline void die(const char* what) {
	fprintf(stderr, "lfmgpu: %s: %s\n", what, lfmgpu_last_error());
	MPI_Abort(MPI_COMM_WORLD, 700);
	abort();
}
#define LFMGPU_OK(call) \
	do { \
		if ((call) != 0) lfmgpu_seam::die(#call); \
	} while (0)

}  // namespace lfmgpu_seam

template <typename PRECISION, unsigned DIM_CNT, unsigned FACE_CNT>
class CFDv0_solver_gpu : public CFDv0_solver<PRECISION, DIM_CNT, FACE_CNT>, public lfmgpu_seam::SubView<PRECISION, DIM_CNT> {
	using Base = CFDv0_solver<PRECISION, DIM_CNT, FACE_CNT>;
	using P = PRECISION;
	using Vars = t_solution_vars<PRECISION, DIM_CNT>;
	using State = lfmgpu_seam::RankState<PRECISION, DIM_CNT>;
	static constexpr int D = (int)DIM_CNT;
	static constexpr int NQ = (int)DIM_CNT + 2;

public:
	// ---- SubView -----------------------------------------------------------------------------------------
	int subIndex() const override { return this->m_nSubmeshIndex; }
	int nCells() const override { return (int)this->cells_cfd.size(); }
	int faceCnt() const override { return (int)FACE_CNT; }
	int nValid(int c) const override { return this->cells_cfd[(size_t)c].m_nFaceCount; }
	size_t cellStride() const override { return sizeof(CFDv0_cell<P, DIM_CNT, FACE_CNT>); }
	Vars* vars(int c) override { return &this->cells_cfd[(size_t)c].vars; }
	const Vars* neigh(int c, int s) const override { return this->cells_cfd[(size_t)c].neighs[s]; }
	const P* S(int c, int s) const override { return this->cells_cfd[(size_t)c].S[s]; }
	const P* d(int c, int s) const override { return this->cells_cfd[(size_t)c].d[s]; }
	P w(int c, int s) const override { return this->cells_cfd[(size_t)c].weight_linear[s]; }
	P sigma(int c) const override { return this->cells_cfd[(size_t)c].sponge_sigma; }
	int cellIndex(int c) const override { return this->cells_cfd[(size_t)c].m_nCellIndex; }

	// ---- setup seam: the last init virtual Mesh::initializeSolver calls (mesh_solver.cpp:119-120) --------
	void reorder_faces() override {
		Base::reorder_faces();
		State& st = State::get();
		if ((int)st.subs.size() <= this->m_nSubmeshIndex) st.subs.resize((size_t)this->m_nSubmeshIndex + 1, nullptr);
		st.subs[(size_t)this->m_nSubmeshIndex] = this;
	}
	void deallocate() override {
		State& st = State::get();
		if (st.h) {
			lfmgpu_destroy(st.h);
			st.h = nullptr;
			st.built = false;
			st.subs.clear();
		}
		Base::deallocate();
	}

	// ---- per-iteration virtuals, in Mesh::solve call order ----------------------------------------------
	void prepare_for_timestep() override {
		if (!isBnd()) return;          // one device context per rank: the boundary solver drives rank-wide calls
		LFMGPU_OK(lfmgpu_prepare_timestep(handle()));
	}
	void prepare_for_RKstep(int rk_step, int) override {
		if (!isBnd()) return;
		LFMGPU_OK(lfmgpu_prepare_rkstep(handle(), rk_step));
	}
	void set_boundary_conditions() override { LFMGPU_OK(lfmgpu_set_bc(handle())); }
	void calc_gradients(MPI_env&) override { LFMGPU_OK(lfmgpu_gradients(handle(), this->m_nSubmeshIndex)); }
	void calc_gradients_M2AUSM(MPI_env&) override { LFMGPU_OK(lfmgpu_gradients_m2ausm(handle(), this->m_nSubmeshIndex)); }
	void calc_VIS(MPI_env&) override { LFMGPU_OK(lfmgpu_vis(handle(), this->m_nSubmeshIndex)); }
	void calc_VIS_Smagorinsky(MPI_env&) override { LFMGPU_OK(lfmgpu_vis_smagorinsky(handle(), this->m_nSubmeshIndex)); }
	void one_rk_step_M1(int rk_step, P dt, MPI_env&, P* RES) override { stage(LFMGPU_SCHEME_M1, rk_step, dt, RES); }
	void one_rk_step_M2(int rk_step, P dt, MPI_env&, P* RES) override { stage(LFMGPU_SCHEME_M2, rk_step, dt, RES); }
	void one_rk_step_M2AUSM(int rk_step, P dt, MPI_env&, P* RES) override { stage(LFMGPU_SCHEME_M2AUSM, rk_step, dt, RES); }

	void mpi_communication(MPI_env& mpi_env, int comm_step) override {
		State& st = State::get();
		lfmgpu_t h = handle(&mpi_env);
		if (mpi_env.is_serial() || this->ghost_mpi.empty()) return;
		if (st.use_nccl) {
			LFMGPU_OK(lfmgpu_halo_start(h, comm_step));
			return;
		}
		// host-staged: device pack -> the reference's own send buffer -> the reference's own MPI_env calls
		std::vector<P>*sb, *rb;
		std::vector<int>*so, *sc, *ro, *rc, *wo;
		buffers(mpi_env, comm_step, sb, rb, so, sc, ro, rc, wo);
		const int win = mpi_env.m_nCommType == MPI_EXCHANGE_SPLIT ? comm_step : 0;
		if (mpi_env.get_halo_comm_type() == MPI_ONESIDED_NONB) mpi_env.waitWindow(win);
		LFMGPU_OK(lfmgpu_halo_pack_to_host(h, comm_step, sb->data(), sb->size() * sizeof(P)));
		if (mpi_env.get_halo_comm_type() == MPI_ONESIDED_NONB || mpi_env.get_halo_comm_type() == MPI_ONESIDED_BLCK) {
			mpi_env.postWindow(win);
			mpi_env.startWindow(win);
		}
		if (mpi_env.get_halo_comm_type() == MPI_TWOSIDED_PERS) mpi_env.startAllPersistantSendRecv(win);
		for (size_t n = 0; n < mpi_env.mpi_neighbors.size(); n++) {
			switch (mpi_env.get_halo_comm_type()) {
				case MPI_TWOSIDED_NONB:
				case MPI_TWOSIDED_BLCK:
					mpi_env.template new_isendrecv<P>(n, (*sc)[n], &(*sb)[(size_t)(*so)[n]], (*rc)[n], &(*rb)[(size_t)(*ro)[n]], win);
					break;
				case MPI_ONESIDED_NONB:
				case MPI_ONESIDED_BLCK:
					mpi_env.new_irget(n, (*rc)[n], &(*rb)[(size_t)(*ro)[n]], (*wo)[n], win);
					break;
				case MPI_TWOSIDED_PERS:
					break;
				default:
					fprintf(stderr, "lfmgpu: haloCommType %d is not served by the packed GPU halo path\n", (int)mpi_env.get_halo_comm_type());
					MPI_Abort(MPI_COMM_WORLD, 909);
			}
		}
		if (mpi_env.get_halo_comm_type() == MPI_TWOSIDED_BLCK) mpi_env.waitAll(win);
		if (mpi_env.get_halo_comm_type() == MPI_ONESIDED_BLCK) {
			mpi_env.completeWindow(win);
			mpi_env.waitWindow(win);
		}
	}
	void mpi_wait(MPI_env& mpi_env, int comm_step) override {
		State& st = State::get();
		lfmgpu_t h = handle(&mpi_env);
		if (mpi_env.is_serial() || this->ghost_mpi.empty()) return;
		if (st.use_nccl) {
			LFMGPU_OK(lfmgpu_halo_wait(h, comm_step));
			return;
		}
		const int win = mpi_env.m_nCommType == MPI_EXCHANGE_SPLIT ? comm_step : 0;
		mpi_env.halo_comm_wait(win);
		std::vector<P>*sb, *rb;
		std::vector<int>*so, *sc, *ro, *rc, *wo;
		buffers(mpi_env, comm_step, sb, rb, so, sc, ro, rc, wo);
		LFMGPU_OK(lfmgpu_halo_unpack_from_host(h, comm_step, rb->data(), rb->size() * sizeof(P)));
	}

	P compute_cfl(P dt) override {
		if (!isBnd()) return P(0);     // Mesh::solve takes the max over the solvers: the boundary solver answers for the rank
		double v = 0;
		LFMGPU_OK(lfmgpu_cfl(handle(), (double)dt, &v));
		return (P)v;
	}
	P compute_dt(P cflMax) override {
		if (!isBnd()) return P(100.0);
		double v = 0;
		LFMGPU_OK(lfmgpu_dt(handle(), (double)cflMax, &v));
		return (P)v;
	}
	void postProcAverage(int time_step) override {
		if (!isBnd()) return;
		LFMGPU_OK(lfmgpu_average(handle(), time_step));
	}
	void postProcForces(MPI_env& mpi_env, P time, MPI_Datatype mpi_precision) override {
		// cfd_v0.cpp:3169-3250: per wall patch local sums on the device, then the reference's two Allreduces
		this->m_dtimeVec.push_back(time);
		for (size_t itype = 0; itype < this->m_nWallBCList.size(); itype++) {
			double fp[3] = {0, 0, 0}, fv[3] = {0, 0, 0};
			LFMGPU_OK(lfmgpu_forces(handle(), this->m_nWallBCList[itype], fp, fv));
			std::array<P, DIM_CNT> Fpre_local, Fvis_local, Fpre_global, Fvis_global;
			for (int i = 0; i < D; i++) {
				Fpre_local[(size_t)i] = (P)fp[i];
				Fvis_local[(size_t)i] = (P)fv[i];
			}
			MPI_Allreduce(&Fpre_local, &Fpre_global, DIM_CNT, mpi_precision, MPI_SUM, MPI_COMM_WORLD);
			MPI_Allreduce(&Fvis_local, &Fvis_global, DIM_CNT, mpi_precision, MPI_SUM, MPI_COMM_WORLD);
			if (mpi_env.is_master()) {
				this->m_dFpre[itype].push_back(Fpre_global);
				this->m_dFvis[itype].push_back(Fvis_global);
			}
		}
	}

	// ---- output hooks: bring the fields back, then the reference's own writers ---------------------------
	void updateSolutionPrimitives() override {
		pull(LFMGPU_FIELD_Q);
		Base::updateSolutionPrimitives();
	}
	void updateSolutionResidual() override {
		pull(LFMGPU_FIELD_RES);
		Base::updateSolutionResidual();
	}
	void updateAverageField(const int nTimeStep) override {
		pull(LFMGPU_FIELD_PAVG);
		pull(LFMGPU_FIELD_PRMS);
		Base::updateAverageField(nTimeStep);
	}

private:
	bool isBnd() const { return this->m_nSubmeshIndex == 0; }
	void stage(int scheme, int rk_step, P dt, P* RES) {
		lfmgpu_t h = handle();
		const int want_res = (RES != nullptr && rk_step == 0) ? 1 : 0;
		LFMGPU_OK(lfmgpu_rk_stage(h, this->m_nSubmeshIndex, scheme, rk_step, (double)dt, want_res));
		if (want_res) {
			// cfd_v0.cpp:2825-2831: RES[i] += sum over this solver's cells of RES_c[i]^2
			double r[8] = {0};
			LFMGPU_OK(lfmgpu_residual(h, r));
			for (int i = 0; i < NQ; i++) RES[i] += (P)r[i];
		}
	}

	void buffers(MPI_env& mpi_env, int comm_step, std::vector<P>*& sb, std::vector<P>*& rb, std::vector<int>*& so, std::vector<int>*& sc, std::vector<int>*& ro, std::vector<int>*& rc,
	             std::vector<int>*& wo) {
		if (mpi_env.m_nCommType == MPI_EXCHANGE_SPLIT && comm_step == 0) {
			sb = &this->m_SendBufSolvarsList, rb = &this->m_RecvBufSolvarsList;
			so = &this->m_nSendBufSolvarOffsetList, sc = &this->m_nSendBufSolvarCountList;
			ro = &this->m_nRecvBufSolvarOffsetList, rc = &this->m_nRecvBufSolvarCountList, wo = &this->m_nRecvSolvarsWindowOffsetList;
		} else if (mpi_env.m_nCommType == MPI_EXCHANGE_SPLIT) {
			sb = &this->m_SendBufViscousList, rb = &this->m_RecvBufViscousList;
			so = &this->m_nSendBufViscousOffsetList, sc = &this->m_nSendBufViscousCountList;
			ro = &this->m_nRecvBufViscousOffsetList, rc = &this->m_nRecvBufViscousCountList, wo = &this->m_nRecvViscousWindowOffsetList;
		} else {   // PACKED; FULL_BND is served as PACKED (the whole-AoS message has no device meaning)
			sb = &this->m_SendBufList, rb = &this->m_RecvBufList;
			so = &this->m_nSendBufOffsetList, sc = &this->m_nSendBufCountList;
			ro = &this->m_nRecvBufOffsetList, rc = &this->m_nRecvBufCountList, wo = &this->m_nRecvWindowOffsetList;
		}
	}

	// device field -> the base class's cells (this solver's submesh only)
	void pull(int field) {
		State& st = State::get();
		lfmgpu_t h = handle();
		const size_t n = (size_t)st.n_cells;
		const int c0 = st.sub_start[(size_t)this->m_nSubmeshIndex];
		const int comps = (field == LFMGPU_FIELD_PAVG || field == LFMGPU_FIELD_PRMS) ? 1 : NQ;
		std::vector<P> tmp(n * (size_t)comps);
		LFMGPU_OK(lfmgpu_download(h, field, tmp.data(), tmp.size() * sizeof(P)));
		if ((field == LFMGPU_FIELD_PAVG || field == LFMGPU_FIELD_PRMS) && this->m_dpAVG.size() != this->cells_cfd.size()) {
			this->m_dpAVG.assign(this->cells_cfd.size(), P(0));
			this->m_dpRMS.assign(this->cells_cfd.size(), P(0));
		}
		for (size_t c = 0; c < this->cells_cfd.size(); c++) {
			const P* v = &tmp[((size_t)c0 + c) * (size_t)comps];
			switch (field) {
				case LFMGPU_FIELD_Q:
					for (int i = 0; i < NQ; i++) this->cells_cfd[c].vars.q_old[i] = v[i];
					break;
				case LFMGPU_FIELD_RES:
					for (int i = 0; i < NQ; i++) this->cells_cfd[c].vars.RES[i] = v[i];
					break;
				case LFMGPU_FIELD_PAVG: this->m_dpAVG[c] = v[0]; break;
				case LFMGPU_FIELD_PRMS: this->m_dpRMS[c] = v[0]; break;
			}
		}
	}

	// ---- flattening: pointer-linked AoS of every submesh solver of the rank -> lfmgpu_desc ----------------
	lfmgpu_t handle(MPI_env* mpi_env = nullptr) {
		State& st = State::get();
		if (!st.built) build(st, mpi_env);
		return st.h;
	}

	struct Range {
		const char *b, *e;
		size_t stride;
		int kind, id, base;   // kind 0: submesh cells, 1: boundary ghosts of patch id, 2: MPI ghosts of neighbour id
	};

	void build(State& st, MPI_env* mpi_env) {
		using namespace lfmgpu_seam;
		for (auto* s : st.subs)
			if (!s) {
				fprintf(stderr, "lfmgpu: a submesh solver is not a CFDv0_solver_gpu\n");
				MPI_Abort(MPI_COMM_WORLD, 702);
			}
		// the boundary solver owns all ghosts and all halo tables (cfd_v0.cpp:345-535)
		auto* bnd = dynamic_cast<CFDv0_solver_gpu*>(st.subs[0]);
		if (!bnd) {   // Mesh::solve calls prepare_for_timestep on the boundary solver first, so this is the boundary solver's type
			fprintf(stderr, "lfmgpu: the device state must be built from the boundary-submesh solver\n");
			MPI_Abort(MPI_COMM_WORLD, 702);
		}
		CFDv0_solver_gpu& B = *bnd;
		const int n_sub = (int)st.subs.size();
		st.sub_start.assign((size_t)n_sub + 1, 0);
		for (int s = 0; s < n_sub; s++) st.sub_start[(size_t)s + 1] = st.sub_start[(size_t)s] + st.subs[(size_t)s]->nCells();
		const int nc = st.sub_start[(size_t)n_sub];
		int Fmax = 0;
		for (auto* s : st.subs) Fmax = std::max(Fmax, s->faceCnt());

		// ghost index space
		const int n_patch = (int)B.ghost_bnd.size();
		std::vector<int> bc_off((size_t)n_patch + 1, 0);
		for (int p = 0; p < n_patch; p++) bc_off[(size_t)p + 1] = bc_off[(size_t)p] + (int)B.ghost_bnd[(size_t)p].size();
		const int n_bc = bc_off[(size_t)n_patch];
		const int n_nbr = (int)B.ghost_mpi.size();
		std::vector<int> recv_start((size_t)n_nbr + 1, 0), send_start((size_t)n_nbr + 1, 0);
		std::vector<std::vector<int>> inv((size_t)n_nbr);
		for (int n = 0; n < n_nbr; n++) {
			recv_start[(size_t)n + 1] = recv_start[(size_t)n] + (int)B.neigh_cells_to_recv[(size_t)n].size();
			send_start[(size_t)n + 1] = send_start[(size_t)n] + (int)B.local_cells_to_send[(size_t)n].size();
			inv[(size_t)n].assign(B.ghost_mpi[(size_t)n].size(), -1);
			for (size_t i = 0; i < B.neigh_cells_to_recv[(size_t)n].size(); i++) inv[(size_t)n][(size_t)B.neigh_cells_to_recv[(size_t)n][i]] = (int)i;
		}
		const int n_mpi = recv_start[(size_t)n_nbr];

		std::vector<Range> ranges;
		for (int s = 0; s < n_sub; s++) {
			SubView<P, DIM_CNT>* v = st.subs[(size_t)s];
			if (!v->nCells()) continue;
			const char* b = (const char*)v->vars(0);
			ranges.push_back({b, b + v->cellStride() * (size_t)v->nCells(), v->cellStride(), 0, s, st.sub_start[(size_t)s]});
		}
		const size_t gs = sizeof(CFDv0_cell<P, DIM_CNT, FACE_CNT>);
		for (int p = 0; p < n_patch; p++)
			if (!B.ghost_bnd[(size_t)p].empty()) {
				const char* b = (const char*)&B.ghost_bnd[(size_t)p][0].vars;
				ranges.push_back({b, b + gs * B.ghost_bnd[(size_t)p].size(), gs, 1, p, nc + bc_off[(size_t)p]});
			}
		for (int n = 0; n < n_nbr; n++)
			if (!B.ghost_mpi[(size_t)n].empty()) {
				const char* b = (const char*)&B.ghost_mpi[(size_t)n][0].vars;
				ranges.push_back({b, b + gs * B.ghost_mpi[(size_t)n].size(), gs, 2, n, nc + n_bc + recv_start[(size_t)n]});
			}
		auto resolve = [&](const Vars* ptr) -> int {
			const char* p = (const char*)ptr;
			for (const Range& r : ranges)
				if (p >= r.b && p < r.e) {
					const int j = (int)((size_t)(p - r.b) / r.stride);
					if (r.kind != 2) return r.base + j;
					const int i = inv[(size_t)r.id][(size_t)j];
					if (i < 0) {
						fprintf(stderr, "lfmgpu: a face points at an MPI ghost that is never received\n");
						MPI_Abort(MPI_COMM_WORLD, 703);
					}
					return r.base + i;
				}
			fprintf(stderr, "lfmgpu: dangling neighbour pointer\n");
			MPI_Abort(MPI_COMM_WORLD, 704);
			return -1;
		};

		// faces: valid (cell, slot) pairs in the order the reference's stage loops visit them
		std::vector<int32_t> cfs((size_t)nc + 1, 0);
		for (int s = 0; s < n_sub; s++)
			for (int c = 0; c < st.subs[(size_t)s]->nCells(); c++) cfs[(size_t)st.sub_start[(size_t)s] + c + 1] = st.subs[(size_t)s]->nValid(c);
		for (int c = 0; c < nc; c++) cfs[(size_t)c + 1] += cfs[(size_t)c];
		const int nf = cfs[(size_t)nc];
		std::vector<int32_t> face_owner((size_t)nf), face_neigh((size_t)nf), slot_face((size_t)nc * Fmax, 0), cell_gid((size_t)nc);
		std::vector<P> face_S((size_t)nf * D), face_d((size_t)nf * D), face_w((size_t)nf), vol_inv((size_t)nc), sigma((size_t)nc), q0((size_t)nc * NQ);
		lfmgpu_desc ds;
		memset(&ds, 0, sizeof ds);
		for (int s = 0; s < n_sub; s++) {
			SubView<P, DIM_CNT>* v = st.subs[(size_t)s];
			ds.sub_cell_start[s] = st.sub_start[(size_t)s];
			ds.sub_face_start[s] = cfs[(size_t)st.sub_start[(size_t)s]];
			ds.sub_face_cnt[s] = v->faceCnt();
			for (int c = 0; c < v->nCells(); c++) {
				const int g = st.sub_start[(size_t)s] + c;
				const Vars* cv = v->vars(c);
				for (int i = 0; i < NQ; i++) q0[(size_t)g * NQ + i] = cv->q_old[i];
				vol_inv[(size_t)g] = cv->vol_inv;
				sigma[(size_t)g] = v->sigma(c);
				cell_gid[(size_t)g] = v->cellIndex(c);
				for (int k = 0; k < v->nValid(c); k++) {
					const int f = cfs[(size_t)g] + k;
					face_owner[(size_t)f] = g;
					face_neigh[(size_t)f] = resolve(v->neigh(c, k));
					for (int i = 0; i < D; i++) {
						face_S[(size_t)f * D + i] = v->S(c, k)[i];
						face_d[(size_t)f * D + i] = v->d(c, k)[i];
					}
					face_w[(size_t)f] = v->w(c, k);
					slot_face[(size_t)g * Fmax + k] = f + 1;
				}
			}
		}
		ds.sub_cell_start[n_sub] = nc;
		ds.sub_face_start[n_sub] = nf;
		// the slots behind m_nFaceCount are faces owned by the neighbour: find the neighbour's slot that points back
		std::vector<char> used((size_t)nf, 0);
		for (int s = 0; s < n_sub; s++) {
			SubView<P, DIM_CNT>* v = st.subs[(size_t)s];
			for (int c = 0; c < v->nCells(); c++) {
				const int g = st.sub_start[(size_t)s] + c;
				for (int k = v->nValid(c); k < v->faceCnt(); k++) {
					const Vars* np = v->neigh(c, k);
					if (!np) continue;
					const int o = resolve(np);
					int found = -1;
					for (int f = cfs[(size_t)o]; f < cfs[(size_t)o + 1] && found < 0; f++)
						if (face_neigh[(size_t)f] == g && !used[(size_t)f]) {
							bool opposite = true;   // several faces between the same two cells (thin cyclic meshes): match S = -S
							for (int i = 0; i < D; i++) opposite = opposite && face_S[(size_t)f * D + i] == -v->S(c, k)[i];
							if (opposite) found = f;
						}
					for (int f = cfs[(size_t)o]; f < cfs[(size_t)o + 1] && found < 0; f++)
						if (face_neigh[(size_t)f] == g && !used[(size_t)f]) found = f;
					if (found < 0) {
						fprintf(stderr, "lfmgpu: cannot pair an invalid slot with its owner's face\n");
						MPI_Abort(MPI_COMM_WORLD, 705);
					}
					used[(size_t)found] = 1;
					slot_face[(size_t)g * Fmax + k] = -(found + 1);
				}
			}
		}
		// physical boundary ghosts (cfd_v0.cpp:382-431, roles :973-1005)
		std::vector<int32_t> bc_cell((size_t)n_bc), bc_kind((size_t)n_bc, LFMGPU_BC_NONE), bc_patch((size_t)n_bc), bc_face((size_t)n_bc, 0);
		// farfield patches: the reference's ghost state depends on m_dAoA, which it never assigns (cfd_v0.cpp:1141): not served
		if (!B.m_nFarfieldBCList.empty()) {
			fprintf(stderr, "lfmgpu: boundary patch of type farfield: not served by the GPU solver (the reference reads the unset m_dAoA there)\n");
			MPI_Abort(MPI_COMM_WORLD, 706);
		}
		auto role = [&](int p) {
			for (int x : B.m_nWallBCList)
				if (x == p) return (int)LFMGPU_BC_WALL;
			for (int x : B.m_nInletBCList)
				if (x == p) return (int)LFMGPU_BC_INLET;
			for (int x : B.m_nOutletBCList)
				if (x == p) return (int)LFMGPU_BC_OUTLET;
			return (int)LFMGPU_BC_NONE;
		};
		for (int p = 0; p < n_patch; p++)
			for (size_t k = 0; k < B.ghost_bnd[(size_t)p].size(); k++) {
				const int g = bc_off[(size_t)p] + (int)k;
				const int c = B.boundaries[(size_t)p][k][0];   // boundary-submesh cell == traversal index (submesh 0 comes first)
				bc_cell[(size_t)g] = c;
				bc_kind[(size_t)g] = role(p);
				bc_patch[(size_t)g] = p;
				for (int f = cfs[(size_t)c]; f < cfs[(size_t)c + 1]; f++)
					if (face_neigh[(size_t)f] == nc + g) bc_face[(size_t)g] = f;
			}
		// halo tables
		std::vector<int32_t> nbr_rank((size_t)n_nbr, -1), send_cell((size_t)send_start[(size_t)n_nbr]);
		for (size_t r = 0; r < B.rank2local.size(); r++)
			if (B.rank2local[r] >= 0) nbr_rank[(size_t)B.rank2local[r]] = (int)r;
		for (int n = 0; n < n_nbr; n++)
			for (size_t i = 0; i < B.local_cells_to_send[(size_t)n].size(); i++) send_cell[(size_t)send_start[(size_t)n] + i] = B.local_cells_to_send[(size_t)n][i];

		ds.precision = (int)sizeof(P);
		ds.dim = D;
		ds.max_slots = Fmax;
		ds.n_sub = n_sub;
		ds.n_cells = nc;
		ds.n_faces = nf;
		ds.n_bc_ghosts = n_bc;
		ds.n_mpi_ghosts = n_mpi;
		ds.face_owner = face_owner.data();
		ds.face_neigh = face_neigh.data();
		ds.face_S = face_S.data();
		ds.face_d = face_d.data();
		ds.face_w = face_w.data();
		ds.vol_inv = vol_inv.data();
		ds.sponge_sigma = sigma.data();
		ds.q0 = q0.data();
		ds.cell_gid = cell_gid.data();
		ds.cell_slot_face = slot_face.data();
		ds.bc_cell = bc_cell.data();
		ds.bc_kind = bc_kind.data();
		ds.bc_patch = bc_patch.data();
		ds.bc_face = bc_face.data();
		ds.n_nbr = n_nbr;
		ds.nbr_rank = nbr_rank.data();
		ds.send_start = send_start.data();
		ds.send_cell = send_cell.data();
		ds.recv_start = recv_start.data();
		lfmgpu_consts& k = ds.c;
		k.gamma = B.m_dGamma;
		k.gamma_m1 = B.m_dGammaMinusOne;
		k.Rgas_inv = B.m_dRgas_inv;
		k.mu = B.m_dMu0;
		k.Cp = B.m_dCp;
		k.Pr_inv = B.m_dPr_inv;
		k.rhoInf = B.m_drhoInf;
		for (int i = 0; i < D; i++) k.UInf[i] = B.m_dUInf[i];
		k.EInf = B.m_dEInf;
		k.pInf = B.m_dpInf;
		k.TInf = B.m_dTInf;
		k.rk_order = (int)B.m_dAk.size();
		for (size_t i = 0; i < B.m_dAk.size() && i < LFMGPU_MAX_RK; i++) {
			k.Ak[i] = B.m_dAk[i];
			k.Bk[i] = B.m_dBk[i];
		}
		k.comm_type = this->m_pInput ? this->m_pInput->m_nCommType : LFMGPU_COMM_SPLIT;

		int n_dev = 0, rank = 0, size = 1;
		LFMGPU_OK(lfmgpu_device_count(&n_dev));
		if (n_dev < 1) {
			fprintf(stderr, "lfmgpu: no CUDA device (the GPU solver has no CPU fallback)\n");
			MPI_Abort(MPI_COMM_WORLD, 706);
		}
		MPI_Comm_rank(MPI_COMM_WORLD, &rank);
		MPI_Comm_size(MPI_COMM_WORLD, &size);
		LFMGPU_OK(lfmgpu_create(&ds, rank % n_dev, &st.h));
		st.n_cells = nc;
		st.n_bc = n_bc;
		st.n_mpi = n_mpi;
		st.built = true;
		// halo transport
		const char* mode = getenv("LFM_GPU_HALO");
		st.use_nccl = size > 1 && (mode ? !strcmp(mode, "nccl") : size <= n_dev);
		if (st.use_nccl) {
			std::vector<char> ids((size_t)size * 128, 0), mine(128, 0);
			if (rank == 0) LFMGPU_OK(lfmgpu_nccl_unique_id(mine.data()));
			MPI_Allgather(mine.data(), 128, MPI_CHAR, ids.data(), 128, MPI_CHAR, MPI_COMM_WORLD);
			LFMGPU_OK(lfmgpu_comm_init_nccl(st.h, ids.data(), rank, size));
		}
		(void)mpi_env;
	}
};
