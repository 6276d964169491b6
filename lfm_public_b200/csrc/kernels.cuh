// Unfused kernels of the per-iteration solve on SoA state (the fallback of the tile kernels, tile_kernels.cuh) and the
// small per-step kernels (boundary ghosts, halo pack/unpack, reductions, post-processing).
//
// Determinism: every accumulated cell quantity (dudx, dTdx, dq, RES) is produced by ONE thread that
// walks the cell's faces in ascending face id.  Faces are numbered in the order the reference's stage
// loops visit them (owner ascending, slot ascending), so ascending face id IS the reference's
// summation order: first the faces owned by earlier cells, then the cell's own valid
// faces in slot order.  No atomics anywhere.
//
// Kernels (reference function -> kernel):
//   set_boundary_conditions  cfd_v0.cpp:1010  -> k_set_bc
//   calc_VIS                 cfd_v0.cpp:1744  -> k_grad_cell      (cell-centric Green-Gauss gather)
//   one_rk_step_M1/_M2       cfd_v0.cpp:2530/1897 -> k_flux_face (face-ordered) + k_update_cell (gather+sponge+RK)
//   pack / unpack            cfd_v0.cpp:3303.. / 3608.. -> k_pack / k_unpack
//   compute_cfl / compute_dt cfd_v0.cpp:2887 / 2913 -> k_cfl_dt
//   postProcAverage          cfd_v0.cpp:3083  -> k_average
//   postProcForces           cfd_v0.cpp:3169  -> k_forces_face + k_sum_sequential
#pragma once
#include "device_math.cuh"

namespace lfm {

// Record layout of the per-cell state the stage kernels stage in shared memory (one contiguous, sector-aligned record
// per cell, so that a halo cell is one or two long copies instead of one element per staged row):
//   Q record, QW = 8 values : q[0 .. NQ) conservatives | 1/rho | R*psi | c (M1) or H (M2)      (2D: one value of padding)
//   V record, VW = 16 / 8   : dudx[D*D] | dTdx[D] | sigmaU[D] | 3D only: tr = ((0 - d00) - d11) - d22 (the diagonal sum the
//                             flux loops form from dudx before scaling it into tauMC; stored so that a face reads it once)
// Record sizes in bytes (32, 64 or 128) are the TMA swizzle spans, see stage_pipe.cuh.
template <int D> struct Rec {
	static constexpr int NQ = D + 2;
	static constexpr int QW = 8;
	static constexpr int VW = D == 3 ? 16 : 8;
	static constexpr int RHO_INV = NQ, RPSI = NQ + 1, AUX = NQ + 2;
	static constexpr int DUDX = 0, DTDX = D * D, SIGMAU = D * D + D, TR = D * D + 2 * D;   // TR exists in 3D only
};
// tr of a finished dudx (3D): the expression of the flux loops' `diagSum -= dudx[nD][nD]` (cfd_v0.cpp:2716-2720)
template <class R, int D> __host__ __device__ __forceinline__ R dudx_trace_neg(const R* dudx_rowmajor) {
	R t = R(0);
#pragma unroll
	for (int nD = 0; nD < D; nD++) t -= dudx_rowmajor[nD * D + nD];
	return t;
}

// Device view of one rank (all pointers are device pointers).  Per-cell state lives in records (Rec<D>): value k of cell
// x at base[x * QW + k] / vis[x * VW + k]; the other per-cell arrays are SoA (base[k * n_cells + c]), per-face arrays
// base[k * nfs + f].
template <class R> struct DevMesh {
	int n_cells, n_faces, n_bc, n_mpi, n_tot, F;
	size_t ncs, nfs, ngs;              // strides: cells(+ghosts), faces, mpi ghosts
	const int *face_owner, *face_neigh;
	const R *S, *d, *w;                // [D][nfs], [D][nfs], [nfs]
	const R *vol_inv, *sigma;          // [n_cells]
	const int* csr;                    // [F][n_cells] signed (face+1), ascending |face|, 0-padded
	const int* slot;                   // [F][n_cells] signed (face+1) in the reference's slot order
	R* dq;                             // [NQ][n_cells]
	R* RES;                            // [NQ][n_cells]
	R* vis;                            // [ncs][VW]   V records: dudx, dTdx, sigmaU = U.tau of calc_VIS (real cells: k_*grad*; MPI ghosts: as received)
	R* tauMC;                          // [D*D][ncs]  only with the Smagorinsky closure (laminar: rebuilt from dudx where needed)
	const R* smag_c;                   // [n_cells]   -2 (Cs Delta)^2 of the cell whose face loop leaves the cell's final tauMC
	int les;                           // calc_VIS_Smagorinsky instead of calc_VIS
	R *g_rho, *g_p, *g_U;              // [D|D|D*D][ncs] minmod gradients of solver 2 (calc_gradients_M2AUSM); ghost slots stay 0
	R* flux;                           // [NQ][nfs]   materialised face fluxes (unfused path)
	R *pAVG, *pRMS;
	const int *bc_cell, *bc_kind, *bc_face, *bc_patch;
	Consts<R> k;
};

constexpr int kBlock = 256;

// ---------------------------------------------------------------------------------------------------
// set_boundary_conditions (cfd_v0.cpp:1010-1133): ghost conservatives of wall / inlet / outlet faces
// ---------------------------------------------------------------------------------------------------
// The three derived values that travel with the conservatives of a cell (tile kernels copy them instead of
// recomputing them per staged cell): 1/rho, R*psi (cfdv0_solver.h:252-261) and, per scheme, the speed of sound
// sqrt(gamma R psi) (M1, cfd_v0.cpp:2600-2603) or the total enthalpy rhoE/rho + R psi (M2, cfd_v0.cpp:1975-1978).
template <class R, int D> __device__ __forceinline__ void store_derived(const Consts<R>& k, const R* cq, int scheme, R* __restrict__ q, size_t x) {
	CellState<R, D> s;
#pragma unroll
	for (int i = 0; i < D + 2; i++) s.q[i] = cq[i];
	if (scheme == 0)
		derive_state<R, D, 0>(k, s);
	else
		derive_state<R, D, 1>(k, s);
	R* rec = q + x * Rec<D>::QW;
	rec[Rec<D>::RHO_INV] = s.rho_inv;
	rec[Rec<D>::RPSI] = s.Rpsi;
	rec[Rec<D>::AUX] = s.aux;
}

// derived values of cells [c0, c1) (after an upload, or when the scheme changes)
template <class R, int D> __global__ void __launch_bounds__(kBlock) k_derive(DevMesh<R> m, R* __restrict__ q, int c0, int c1, int scheme) {
	const int x = c0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (x >= c1) return;
	R cq[D + 2];
#pragma unroll
	for (int i = 0; i < D + 2; i++) cq[i] = q[(size_t)x * Rec<D>::QW + i];
	store_derived<R, D>(m.k, cq, scheme, q, (size_t)x);
}

template <class R, int D> __global__ void __launch_bounds__(kBlock) k_set_bc(DevMesh<R> m, R* __restrict__ q, int scheme) {
	const int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= m.n_bc) return;
	const int b = m.bc_cell[g];
	const int kind = m.bc_kind[g];
	const size_t gi = (size_t)m.n_cells + g;
	R cq[D + 2];
#pragma unroll
	for (int i = 0; i < D + 2; i++) cq[i] = q[(size_t)b * Rec<D>::QW + i];
	R gq[D + 2];
	if (kind == 1) {          // wall
		gq[0] = cq[0];
#pragma unroll
		for (int i = 0; i < D; i++) gq[i + 1] = -cq[i + 1];
		gq[D + 1] = cq[D + 1];
	} else if (kind == 2 || kind == 3) {
		const R rhoInt = cq[0];
		R UvecInt[D];
		R Umag_sqrtInt = R(0.0);
#pragma unroll
		for (int i = 0; i < D; i++) {
			UvecInt[i] = cq[i + 1] / rhoInt;
			Umag_sqrtInt += UvecInt[i] * UvecInt[i];
		}
		const R pInt = (cq[D + 1] - R(0.5) * Umag_sqrtInt * rhoInt) * m.k.gm1;
		R rhoExt, pExt, UvecExt[D];
		if (kind == 2) {      // inlet: T = Tinf, p from inside, U = Uinf
			const R TInt = pInt * m.k.Rgas_inv / rhoInt;
			rhoExt = rhoInt * TInt / m.k.TInf;
			pExt = pInt;
#pragma unroll
			for (int i = 0; i < D; i++) UvecExt[i] = m.k.UInf[i];
		} else {              // outlet: p = pinf, U from inside
			pExt = m.k.pInf;
			rhoExt = rhoInt * pExt / pInt;
#pragma unroll
			for (int i = 0; i < D; i++) UvecExt[i] = UvecInt[i];
		}
		R Umag_sqrtExt = R(0.0);
#pragma unroll
		for (int i = 0; i < D; i++) Umag_sqrtExt += UvecExt[i] * UvecExt[i];
		const R EExt = pExt / (rhoExt * m.k.gm1) + R(0.5) * Umag_sqrtExt;
		gq[0] = rhoExt;
#pragma unroll
		for (int i = 0; i < D; i++) gq[i + 1] = rhoExt * UvecExt[i];
		gq[D + 1] = rhoExt * EExt;
	} else {
		return;               // unknown role: the reference leaves the ghost untouched
	}
#pragma unroll
	for (int i = 0; i < D + 2; i++) q[gi * Rec<D>::QW + i] = gq[i];
	store_derived<R, D>(m.k, gq, scheme, q, gi);
}

// writes the V record of cell x (dudx, dTdx, sigmaU and, in 3D, the diagonal sum tr)
template <class R, int D> __device__ __forceinline__ void store_vis_record(R* __restrict__ vis, size_t x, const R (*dudx)[D], const R* dTdx, const R* sigmaU) {
	R* rec = vis + x * Rec<D>::VW;
#pragma unroll
	for (int i = 0; i < D; i++) {
#pragma unroll
		for (int j = 0; j < D; j++) rec[Rec<D>::DUDX + i * D + j] = dudx[i][j];
		rec[Rec<D>::DTDX + i] = dTdx[i];
		rec[Rec<D>::SIGMAU + i] = sigmaU[i];
	}
	if (D == 3) rec[Rec<D>::TR] = dudx_trace_neg<R, D>(&dudx[0][0]);
}

// ---------------------------------------------------------------------------------------------------
// calc_VIS (cfd_v0.cpp:1744-1860) as a cell-centric gather: dudx, dTdx of cells [c0, c1)
// ---------------------------------------------------------------------------------------------------
template <class R, int D> __global__ void __launch_bounds__(kBlock) k_grad_cell(DevMesh<R> m, const R* __restrict__ q, int c0, int c1) {
	const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= c1) return;
	R cq[D + 2], cU[D], c_rho_inv, c_Rpsi, c_T;
#pragma unroll
	for (int i = 0; i < D + 2; i++) cq[i] = q[(size_t)c * Rec<D>::QW + i];
	primitives<R, D>(m.k, cq, c_rho_inv, cU, c_Rpsi, c_T);
	const R vinv = m.vol_inv[c];
	R dudx[D][D], dTdx[D];
#pragma unroll
	for (int i = 0; i < D; i++) {
		dTdx[i] = R(0);
#pragma unroll
		for (int j = 0; j < D; j++) dudx[i][j] = R(0);
	}
	for (int s = 0; s < m.F; s++) {
		const int e = m.csr[(size_t)s * m.n_cells + c];
		if (e == 0) break;
		const bool own = e > 0;
		const int f = (own ? e : -e) - 1;
		const int o = own ? m.face_neigh[f] : m.face_owner[f];
		R oq[D + 2], oU[D], o_rho_inv, o_Rpsi, o_T;
#pragma unroll
		for (int i = 0; i < D + 2; i++) oq[i] = q[(size_t)o * Rec<D>::QW + i];
		primitives<R, D>(m.k, oq, o_rho_inv, oU, o_Rpsi, o_T);
		const R w = m.w[f];
		R face_U[D], face_T, sov[D];
		if (own) {
			grad_face_values<R, D>(m.k, w, cU, c_Rpsi, oU, o_Rpsi, face_U, face_T);
#pragma unroll
			for (int i = 0; i < D; i++) sov[i] = m.S[i * m.nfs + f] * vinv;
		} else {
			grad_face_values<R, D>(m.k, w, oU, o_Rpsi, cU, c_Rpsi, face_U, face_T);
#pragma unroll
			for (int i = 0; i < D; i++) sov[i] = -m.S[i * m.nfs + f] * vinv;
		}
#pragma unroll
		for (int i = 0; i < D; i++) {
#pragma unroll
			for (int j = 0; j < D; j++) dudx[i][j] += face_U[i] * sov[j];
			dTdx[i] += face_T * sov[i];
		}
	}
	R tauMC[D][D], sigmaU[D];
	vis_cell_terms<R, D>(m.k, cq, dudx, tauMC, sigmaU);
	store_vis_record<R, D>(m.vis, (size_t)c, dudx, dTdx, sigmaU);
}

// ---------------------------------------------------------------------------------------------------
// calc_gradients_M2AUSM (cfd_v0.cpp:1384-1495) as a cell-centric gather: Green-Gauss gradients of rho, p and U of cells
// [c0, c1) (the three one_rk_step_M2AUSM reads; the reference also fills rhoU/rhoE/Rpsi/c gradients that nothing reads).
// ---------------------------------------------------------------------------------------------------
template <class R, int D> __device__ __forceinline__ R pressure_of(const Consts<R>& k, const R* q) {
	const R r = q[0];
	const R E = q[D + 1] / r;
	R vmag = R(0.0);
#pragma unroll
	for (int i = 0; i < D; i++) vmag += (q[i + 1] / r) * (q[i + 1] / r);
	return r * k.gm1 * (E - R(0.5) * vmag);
}
template <class R, int D> __global__ void __launch_bounds__(kBlock) k_grad_ausm(DevMesh<R> m, const R* __restrict__ q, int c0, int c1) {
	const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= c1) return;
	R cq[D + 2];
#pragma unroll
	for (int i = 0; i < D + 2; i++) cq[i] = q[(size_t)c * Rec<D>::QW + i];
	const R cp = pressure_of<R, D>(m.k, cq);
	const R vinv = m.vol_inv[c];
	R g_rho[D], g_p[D], g_U[D][D];
#pragma unroll
	for (int i = 0; i < D; i++) {
		g_rho[i] = g_p[i] = R(0);
#pragma unroll
		for (int j = 0; j < D; j++) g_U[i][j] = R(0);
	}
	for (int s = 0; s < m.F; s++) {
		const int e = m.csr[(size_t)s * m.n_cells + c];
		if (e == 0) break;
		const bool own = e > 0;
		const int f = (own ? e : -e) - 1;
		const int o = own ? m.face_neigh[f] : m.face_owner[f];
		R oq[D + 2];
#pragma unroll
		for (int i = 0; i < D + 2; i++) oq[i] = q[(size_t)o * Rec<D>::QW + i];
		const R op = pressure_of<R, D>(m.k, oq);
		const R w = m.w[f];
		// face values seen from the face's owner: INTERP_LINEAR(w, owner, neighbour)
		const R* a = own ? cq : oq;
		const R* b = own ? oq : cq;
		R UU[D], sov[D];
#pragma unroll
		for (int i = 0; i < D; i++) {
			UU[i] = interp<R>(w, a[i + 1] / a[0], b[i + 1] / b[0]);
			sov[i] = own ? m.S[i * m.nfs + f] * vinv : -m.S[i * m.nfs + f] * vinv;
		}
		const R rho = interp<R>(w, a[0], b[0]);
		const R p = own ? interp<R>(w, cp, op) : interp<R>(w, op, cp);
#pragma unroll
		for (int i = 0; i < D; i++) {
			g_p[i] += p * sov[i];
#pragma unroll
			for (int j = 0; j < D; j++) g_U[j][i] += UU[j] * sov[i];
			g_rho[i] += rho * sov[i];
		}
	}
#pragma unroll
	for (int i = 0; i < D; i++) {
		m.g_rho[(size_t)i * m.ncs + c] = g_rho[i];
		m.g_p[(size_t)i * m.ncs + c] = g_p[i];
#pragma unroll
		for (int j = 0; j < D; j++) m.g_U[(size_t)(i * D + j) * m.ncs + c] = g_U[i][j];
	}
}

// Loads everything face_flux needs about cell x (real cell, physical ghost or MPI ghost).
template <class R, int D, int SCHEME> __device__ __forceinline__ void load_state(const DevMesh<R>& m, const R* __restrict__ q, int x, CellState<R, D>& s) {
#pragma unroll
	for (int i = 0; i < D + 2; i++) s.q[i] = q[(size_t)x * Rec<D>::QW + i];
	const bool bc_ghost = x >= m.n_cells && x < m.n_cells + m.n_bc;
	if (bc_ghost) {
#pragma unroll
		for (int i = 0; i < D; i++) {
			s.dTdx[i] = s.sigmaU[i] = R(0);
#pragma unroll
			for (int j = 0; j < D; j++) s.dudx[i][j] = R(0);
		}
	} else {
		const R* rec = m.vis + (size_t)x * Rec<D>::VW;
#pragma unroll
		for (int i = 0; i < D; i++) {
#pragma unroll
			for (int j = 0; j < D; j++) s.dudx[i][j] = rec[Rec<D>::DUDX + i * D + j];
			s.dTdx[i] = rec[Rec<D>::DTDX + i];
			s.sigmaU[i] = rec[Rec<D>::SIGMAU + i];
		}
	}
	derive_state<R, D, SCHEME>(m.k, s);
}

// ---------------------------------------------------------------------------------------------------
// face-ordered flux kernel: rhs of faces [f0, f1) -> m.flux   (one_rk_step_M1/_M2 face body)
// ---------------------------------------------------------------------------------------------------
template <class R, int D, int SCHEME> __global__ void __launch_bounds__(kBlock) k_flux_face(DevMesh<R> m, const R* __restrict__ q, int f0, int f1) {
	const int f = f0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= f1) return;
	const int o = m.face_owner[f], n = m.face_neigh[f];
	CellState<R, D> c, a;
	load_state<R, D, SCHEME>(m, q, o, c);
	load_state<R, D, SCHEME>(m, q, n, a);
	R S[D], dv[D], rhs[D + 2];
#pragma unroll
	for (int i = 0; i < D; i++) {
		S[i] = m.S[i * m.nfs + f];
		dv[i] = m.d[i * m.nfs + f];
	}
	FaceGeo<R, D> g;
	make_geo<R, D>(S, dv, m.w[f], g);
	const bool ghost = n >= m.n_cells && n < m.n_cells + m.n_bc;
	R extra = R(0);
	if (SCHEME == 2) {
		R cUg[D][D], nUg[D][D], crg[D], nrg[D], cpg[D], npg[D];
#pragma unroll
		for (int i = 0; i < D; i++) {
			crg[i] = m.g_rho[(size_t)i * m.ncs + o];
			nrg[i] = m.g_rho[(size_t)i * m.ncs + n];
			cpg[i] = m.g_p[(size_t)i * m.ncs + o];
			npg[i] = m.g_p[(size_t)i * m.ncs + n];
#pragma unroll
			for (int j = 0; j < D; j++) {
				cUg[i][j] = m.g_U[(size_t)(i * D + j) * m.ncs + o];
				nUg[i][j] = m.g_U[(size_t)(i * D + j) * m.ncs + n];
			}
		}
		extra = ausm_pressure_term<R, D>(m.k, c.q, a.q, c.Rpsi, a.Rpsi, c.dudx, a.dudx, cUg, nUg, crg, nrg, cpg, npg, S, dv, g.w);
	}
	face_flux<R, D, SCHEME>(m.k, RegSide<R, D>{c}, RegSide<R, D>{a}, g, ghost, dv, rhs, extra);
#pragma unroll
	for (int i = 0; i < D + 2; i++) m.flux[i * m.nfs + f] = rhs[i];
}

// ---------------------------------------------------------------------------------------------------
// fused RK-stage cell kernel: dq = A_k dq + sum_faces(+-dt rhs / V) + sponge ; q_new = q + B_k dq
// (prepare_for_RKstep's dq *= A_k, the scatter of cfd_v0.cpp:2792-2804 as a gather, sponge 2810-2814,
//  update 2819-2822).  first: dq starts from zero (prepare_for_timestep).  res: also gather RES.
// ---------------------------------------------------------------------------------------------------
template <class R, int D> __global__ void __launch_bounds__(kBlock) k_update_cell(DevMesh<R> m, const R* __restrict__ q, R* __restrict__ qn, int c0, int c1, R dt, R Ak, R Bk, int first, int res) {
	const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= c1) return;
	constexpr int NQ = D + 2;
	R dq[NQ], RES[NQ];
#pragma unroll
	for (int i = 0; i < NQ; i++) {
		dq[i] = first ? R(0) : m.dq[(size_t)i * m.n_cells + c] * Ak;
		RES[i] = R(0);
	}
	const R vinv = m.vol_inv[c];
	for (int s = 0; s < m.F; s++) {
		const int e = m.csr[(size_t)s * m.n_cells + c];
		if (e == 0) break;
		const int f = (e > 0 ? e : -e) - 1;
		if (e > 0) {
#pragma unroll
			for (int i = 0; i < NQ; i++) {
				const R r = m.flux[i * m.nfs + f];
				if (res) RES[i] += r;
				dq[i] += dt * r * vinv;
			}
		} else {
#pragma unroll
			for (int i = 0; i < NQ; i++) {
				const R r = m.flux[i * m.nfs + f];
				if (res) RES[i] -= r;
				dq[i] -= dt * r * vinv;
			}
		}
	}
	R cq[NQ];
#pragma unroll
	for (int i = 0; i < NQ; i++) cq[i] = q[(size_t)c * Rec<D>::QW + i];
	const R sg = m.sigma[c];
	dq[0] += dt * sg * (m.k.rhoInf - cq[0]);
#pragma unroll
	for (int i = 0; i < D; i++) dq[i + 1] += dt * sg * (m.k.rhoUInf[i] - cq[i + 1]);
	dq[D + 1] += dt * sg * (m.k.rhoEInf - cq[D + 1]);
#pragma unroll
	for (int i = 0; i < NQ; i++) {
		m.dq[(size_t)i * m.n_cells + c] = dq[i];
		qn[(size_t)c * Rec<D>::QW + i] = cq[i] + Bk * dq[i];
		if (res) m.RES[(size_t)i * m.n_cells + c] = RES[i];
	}
}

// copies q of cells [c0,c1) between the two buffers (used when only part of the submeshes is advanced)
template <class R> __global__ void __launch_bounds__(kBlock) k_copy_q(const R* __restrict__ src, R* __restrict__ dst, int QW, int c0, int c1) {
	const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= c1) return;
	for (int i = 0; i < QW; i++) dst[(size_t)c * QW + i] = src[(size_t)c * QW + i];
}

// ---------------------------------------------------------------------------------------------------
// record <-> dense array converters of the data-movement entry points (lfmgpu_download / _upload_q / _pipe_*)
// ---------------------------------------------------------------------------------------------------
// out[i * comps + k] = rec[(x0 + i) * W + comp0 + k]      (AoS, the layout lfmgpu_download hands to the host)
template <class R> __global__ void __launch_bounds__(kBlock) k_rec_gather(const R* __restrict__ rec, int W, int comp0, int comps, size_t x0, size_t n, R* __restrict__ out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	for (int k = 0; k < comps; k++) out[i * comps + k] = rec[(x0 + i) * W + comp0 + k];
}
// rec[i * W + k] = in[i * comps + k]   (aos != 0)   or   in[k * n + i]   (component-major host arrays of the end-to-end path)
template <class R> __global__ void __launch_bounds__(kBlock) k_rec_scatter(const R* __restrict__ in, int aos, size_t n, int comps, R* __restrict__ rec, int W) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	for (int k = 0; k < comps; k++) rec[i * W + k] = aos ? in[i * comps + k] : in[(size_t)k * n + i];
}
// out[k * n + i] = rec[i * W + k]
template <class R> __global__ void __launch_bounds__(kBlock) k_rec_to_soa(const R* __restrict__ rec, int W, size_t n, int comps, R* __restrict__ out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	for (int k = 0; k < comps; k++) out[(size_t)k * n + i] = rec[i * W + k];
}
// laminar tauMC of cells [0, n) rebuilt from the stored dudx with the expression of calc_VIS (cfd_v0.cpp:1838-1850): AoS [n][D][D]
template <class R, int D> __global__ void __launch_bounds__(kBlock) k_taumc_laminar(DevMesh<R> m, size_t n, R* __restrict__ out) {
	const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n) return;
	R dudx[D][D], tauMC[D][D];
	const R* rec = m.vis + c * Rec<D>::VW;
#pragma unroll
	for (int i = 0; i < D; i++)
#pragma unroll
		for (int j = 0; j < D; j++) dudx[i][j] = rec[Rec<D>::DUDX + i * D + j];
	tauMC_from<R, D>(m.k, dudx, tauMC);
#pragma unroll
	for (int i = 0; i < D; i++)
#pragma unroll
		for (int j = 0; j < D; j++) out[(c * D + i) * D + j] = tauMC[i][j];
}

// ---------------------------------------------------------------------------------------------------
// halo pack (cfd_v0.cpp:3303-3331 / 3398-3408 / 3475-3499) and unpack (3608-3692)
// mode bit 0: q payload, bit 1: viscous payload.  Wire order per cell: q | dudx | dTdx | tauMC | sigmaU
// ---------------------------------------------------------------------------------------------------
template <class R, int D> __global__ void __launch_bounds__(kBlock) k_pack(DevMesh<R> m, const R* __restrict__ q, const int* __restrict__ send_cell, int n_send, int mode, R* __restrict__ buf) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_send) return;
	constexpr int NQ = D + 2, NV = 2 * D * D + 2 * D;
	const int spc = ((mode & 1) ? NQ : 0) + ((mode & 2) ? NV : 0);
	const int c = send_cell[i];
	R* o = buf + (size_t)i * spc;
	R cq[NQ];
#pragma unroll
	for (int k = 0; k < NQ; k++) cq[k] = q[(size_t)c * Rec<D>::QW + k];
	if (mode & 1) {
#pragma unroll
		for (int k = 0; k < NQ; k++) *o++ = cq[k];
	}
	if (mode & 2) {
		R dudx[D][D], dTdx[D], tauMC[D][D], sigmaU[D];
		const R* rec = m.vis + (size_t)c * Rec<D>::VW;
#pragma unroll
		for (int a = 0; a < D; a++) {
#pragma unroll
			for (int b = 0; b < D; b++) dudx[a][b] = rec[Rec<D>::DUDX + a * D + b];
			dTdx[a] = rec[Rec<D>::DTDX + a];
		}
		if (m.les) {
#pragma unroll
			for (int a = 0; a < D; a++)
#pragma unroll
				for (int b = 0; b < D; b++) tauMC[a][b] = m.tauMC[(size_t)(a * D + b) * m.ncs + c];
		} else {
			tauMC_from<R, D>(m.k, dudx, tauMC);
		}
#pragma unroll
		for (int a = 0; a < D; a++) sigmaU[a] = rec[Rec<D>::SIGMAU + a];
#pragma unroll
		for (int a = 0; a < D; a++)
#pragma unroll
			for (int b = 0; b < D; b++) *o++ = dudx[a][b];
#pragma unroll
		for (int a = 0; a < D; a++) *o++ = dTdx[a];
#pragma unroll
		for (int a = 0; a < D; a++)
#pragma unroll
			for (int b = 0; b < D; b++) *o++ = tauMC[a][b];
#pragma unroll
		for (int a = 0; a < D; a++) *o++ = sigmaU[a];
	}
}

template <class R, int D> __global__ void __launch_bounds__(kBlock) k_unpack(DevMesh<R> m, R* __restrict__ q, int scheme, int n_recv, int mode, const R* __restrict__ buf) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_recv) return;
	constexpr int NQ = D + 2, NV = 2 * D * D + 2 * D;
	const int spc = ((mode & 1) ? NQ : 0) + ((mode & 2) ? NV : 0);
	const size_t g = (size_t)m.n_cells + m.n_bc + i;
	const R* o = buf + (size_t)i * spc;
	if (mode & 1) {
		R gq[NQ];
#pragma unroll
		for (int k = 0; k < NQ; k++) {
			gq[k] = *o++;
			q[g * Rec<D>::QW + k] = gq[k];
		}
		store_derived<R, D>(m.k, gq, scheme, q, g);
	}
	if (mode & 2) {
		R* rec = m.vis + g * Rec<D>::VW;
		R du[D * D];
#pragma unroll
		for (int a = 0; a < D * D; a++) rec[Rec<D>::DUDX + a] = du[a] = *o++;
		if (D == 3) rec[Rec<D>::TR] = dudx_trace_neg<R, D>(du);
#pragma unroll
		for (int a = 0; a < D; a++) rec[Rec<D>::DTDX + a] = *o++;
		if (m.les) {   // with the Smagorinsky closure tauMC is not a function of dudx alone: keep what the neighbour computed
#pragma unroll
			for (int a = 0; a < D * D; a++) m.tauMC[(size_t)a * m.ncs + g] = o[a];
		}
		o += D * D;   // laminar: tauMC is a function of the dudx just stored (same expression on both ranks): not kept
#pragma unroll
		for (int a = 0; a < D; a++) rec[Rec<D>::SIGMAU + a] = *o++;
	}
}

// ---------------------------------------------------------------------------------------------------
// compute_cfl (cfd_v0.cpp:2887-2909) / compute_dt (2913-2942): block max / min -> partial[blockIdx]
// what = 0: cfl (max), 1: dt (min)
// ---------------------------------------------------------------------------------------------------
template <class R, int D> __global__ void __launch_bounds__(kBlock) k_cfl_dt(DevMesh<R> m, const R* __restrict__ q, R arg, int what, R* __restrict__ partial) {
	__shared__ R sh[kBlock];
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	R v = what == 0 ? R(0) : R(100.0);
	if (c < m.n_cells) {
		R cq[D + 1];
#pragma unroll
		for (int i = 0; i < D + 1; i++) cq[i] = q[(size_t)c * Rec<D>::QW + i];
		R acc = R(0);
		for (int s = 0; s < m.F; s++) {
			const int e = m.slot[(size_t)s * m.n_cells + c];
			if (e == 0) continue;
			const int f = (e > 0 ? e : -e) - 1;
			R S[D];
#pragma unroll
			for (int i = 0; i < D; i++) S[i] = e > 0 ? m.S[i * m.nfs + f] : -m.S[i * m.nfs + f];
			R flux = R(0);
			if (what == 0) {
#pragma unroll
				for (int i = 0; i < D; i++) flux += cq[i + 1] * S[i];
			} else {
				R sMag = R(0);
#pragma unroll
				for (int i = 0; i < D; i++) sMag += S[i] * S[i];
				sMag = LFM_SQRT(sMag);
#pragma unroll
				for (int i = 0; i < D; i++) flux += cq[i + 1] * S[i] / sMag;
			}
			acc += flux < R(0) ? -flux : flux;
		}
		if (what == 0) {
			const R rho_inv = R(1.0) / cq[0];
			v = acc;
			v *= rho_inv * m.vol_inv[c] * arg;
		} else {
			const R vol = R(1.0) / m.vol_inv[c];
			v = arg * cq[0] * R(2.0) * vol / acc;
		}
	}
	sh[threadIdx.x] = v;
	__syncthreads();
	for (int s = kBlock / 2; s > 0; s >>= 1) {
		if (threadIdx.x < s) {
			const R o = sh[threadIdx.x + s];
			if (what == 0 ? (sh[threadIdx.x] < o) : (o < sh[threadIdx.x])) sh[threadIdx.x] = o;
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

template <class R> __global__ void k_reduce_minmax(const R* __restrict__ partial, int n, int what, R* __restrict__ out) {
	__shared__ R sh[kBlock];
	R v = what == 0 ? R(0) : R(100.0);
	for (int i = threadIdx.x; i < n; i += kBlock) {
		const R o = partial[i];
		if (what == 0 ? (v < o) : (o < v)) v = o;
	}
	sh[threadIdx.x] = v;
	__syncthreads();
	for (int s = kBlock / 2; s > 0; s >>= 1) {
		if (threadIdx.x < s) {
			const R o = sh[threadIdx.x + s];
			if (what == 0 ? (sh[threadIdx.x] < o) : (o < sh[threadIdx.x])) sh[threadIdx.x] = o;
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) out[0] = sh[0];
}

// ---------------------------------------------------------------------------------------------------
// postProcAverage (cfd_v0.cpp:3083-3103)
// ---------------------------------------------------------------------------------------------------
template <class R, int D> __global__ void __launch_bounds__(kBlock) k_average(DevMesh<R> m, const R* __restrict__ q, int time_step) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= m.n_cells) return;
	const R* rec = q + (size_t)c * Rec<D>::QW;
	const R r = rec[0];
	R velMag = R(0);
#pragma unroll
	for (int nD = 1; nD <= D; nD++) {
		const R vel = rec[nD] / r;
		velMag += vel * vel;
	}
	const R E = rec[D + 1] / r;
	const R p = (r * m.k.gm1) * (E - R(0.5) * velMag);
	const R avg = m.pAVG[c] + p;
	m.pAVG[c] = avg;
	const R dev = p - avg / R(time_step);
	m.pRMS[c] += dev * dev;
}

// ---------------------------------------------------------------------------------------------------
// stage-0 residual norm (cfd_v0.cpp:2825-2831): sum over cells [c0,c1) of RES^2, fixed-shape tree
// ---------------------------------------------------------------------------------------------------
template <class R> __global__ void __launch_bounds__(kBlock) k_res_partial(const R* __restrict__ RES, int n_cells, int NQ, int c0, int c1, double* __restrict__ partial) {
	__shared__ double sh[kBlock];
	for (int i = 0; i < NQ; i++) {
		const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
		double v = 0.0;
		if (c < c1) {
			const R r = RES[(size_t)i * n_cells + c];
			v = (double)(r * r);
		}
		sh[threadIdx.x] = v;
		__syncthreads();
		for (int s = kBlock / 2; s > 0; s >>= 1) {
			if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
			__syncthreads();
		}
		if (threadIdx.x == 0) partial[(size_t)i * gridDim.x + blockIdx.x] = sh[0];
		__syncthreads();
	}
}
// out[i] += sum(partial[i][:]) with a fixed-shape tree (one block per component)
__global__ void __launch_bounds__(kBlock) k_res_final(const double* __restrict__ partial, int n, double* __restrict__ out) {
	__shared__ double sh[kBlock];
	double v = 0.0;
	for (int i = threadIdx.x; i < n; i += kBlock) v += partial[(size_t)blockIdx.x * n + i];
	sh[threadIdx.x] = v;
	__syncthreads();
	for (int s = kBlock / 2; s > 0; s >>= 1) {
		if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[blockIdx.x] += sh[0];
}

// ---------------------------------------------------------------------------------------------------
// postProcForces (cfd_v0.cpp:3173-3240): per wall face contributions, then a sequential sum in the
// reference's face order (bit-exact with the CPU loop)
// ---------------------------------------------------------------------------------------------------
template <class R, int D> __global__ void __launch_bounds__(kBlock) k_forces_face(DevMesh<R> m, const R* __restrict__ q, const R* __restrict__ qghost, int patch, R* __restrict__ contrib /*[n_bc][2*D]*/, int* __restrict__ used) {
	const int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= m.n_bc) return;
	if (m.bc_patch[g] != patch) {
		used[g] = 0;
		return;
	}
	used[g] = 1;
	const int t = m.bc_cell[g], f = m.bc_face[g];
	const size_t n = (size_t)m.n_cells + g;
	R S[D], dv[D], cq[D + 2], nq[D + 2];
#pragma unroll
	for (int i = 0; i < D; i++) {
		S[i] = m.S[i * m.nfs + f];
		dv[i] = m.d[i * m.nfs + f];
	}
#pragma unroll
	for (int i = 0; i < D + 2; i++) {
		cq[i] = q[(size_t)t * Rec<D>::QW + i];
		nq[i] = qghost[n * Rec<D>::QW + i];
	}
	const R r = cq[0], rE = cq[D + 1];
	R Umag = R(0);
#pragma unroll
	for (int i = 0; i < D; i++) {
		const R u = cq[i + 1] / r;
		Umag += u * u;
	}
	Umag = LFM_SQRT(Umag);
	const R E = rE / r;
	const R p = (r * m.k.gm1) * (E - R(0.5) * Umag * Umag);
	R* o = contrib + (size_t)g * 2 * D;
#pragma unroll
	for (int nD = 0; nD < D; nD++) o[nD] = S[nD] * (p - m.k.pInf);
	R d_mag = dv[0] * dv[0];
#pragma unroll
	for (int i = 1; i < D; i++) d_mag += dv[i] * dv[i];
	d_mag = LFM_SQRT(d_mag);
	const R dmag_inv = R(1.0) / d_mag;
	R d_norm[D], dudx[D][D], tau[D][D];
#pragma unroll
	for (int i = 0; i < D; i++) d_norm[i] = dv[i] * dmag_inv;
#pragma unroll
	for (int i = 0; i < D; i++) {
		const R cell_U = cq[i + 1] / cq[0];
		const R adjc_U = nq[i + 1] / nq[0];
#pragma unroll
		for (int j = 0; j < D; j++) dudx[i][j] = (adjc_U - cell_U) * d_norm[j] * dmag_inv;
	}
	stress<R, D>(m.k, dudx, tau);
	// Fvis[a] -= tau[a][b]*S[b] for b ascending: keep the products separate so the serial sum can replay the order
#pragma unroll
	for (int a = 0; a < D; a++) {
		// the reference subtracts term by term from the running total; store the D products per row
		o[D + a] = R(0);
	}
	R* prod = contrib + (size_t)m.n_bc * 2 * D + (size_t)g * D * D;
#pragma unroll
	for (int a = 0; a < D; a++)
#pragma unroll
		for (int b = 0; b < D; b++) prod[a * D + b] = tau[a][b] * S[b];
}

template <class R, int D> __global__ void k_forces_sum(int n_bc, const R* __restrict__ contrib, const int* __restrict__ used, R* __restrict__ out /*[2*D]*/) {
	const int a = threadIdx.x;
	if (a >= D) return;
	R fp = R(0), fv = R(0);
	const R* prod = contrib + (size_t)n_bc * 2 * D;
	for (int g = 0; g < n_bc; g++) {
		if (!used[g]) continue;
		fp += contrib[(size_t)g * 2 * D + a];
		for (int b = 0; b < D; b++) fv -= prod[(size_t)g * D * D + a * D + b];
	}
	out[a] = fp;
	out[D + a] = fv;
}

}  // namespace lfm
