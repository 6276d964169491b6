// Fused shared-memory tile kernels ("v3"): the production path of the per-iteration solve.
//
// A tile is a run of consecutive cells in traversal order (never straddling a submesh).  One CTA owns one
// tile per launch and
//   A. copies the state of the tile's cells AND of every cell across one of their faces (the halo) into shared
//      memory with asynchronous copies (cp.async / LDGSTS): no registers are tied up, nothing is computed, every
//      byte the tile needs from HBM is requested in the first few hundred cycles of the CTA.  The per-cell
//      divisions / square roots the flux loops need (1/rho, R*psi, c or H) are NOT recomputed here: they travel
//      with the conservatives as three "derived" values per cell, written by whoever writes the conservatives
//      (the stage kernel's epilogue, set_boundary_conditions, the halo unpack; k_derive after an upload);
//   B. evaluates every face that touches a tile cell exactly once per tile, reading both sides from shared
//      memory and the face constants from the plan's TILE-ORDERED face tables (fully coalesced, no indirection;
//      the first face's constants are requested before the wait of phase A, the next face's before the current
//      face is evaluated).  Faces owned by earlier cells outside the tile ("incoming") are the only redundant
//      flux evaluations, a surface-to-volume effect of the cell numbering;
//   C. gathers per cell in ascending face id (the reference's summation order, kernels.cuh) from shared memory,
//      adds the sponge term, applies the low-storage RK update and writes q_new, the derived values of q_new and
//      dq (and RES) coalesced.
// Nothing but the final cell state goes back to HBM: face fluxes never leave the SM, prepare_for_RKstep's
// `dq *= A_k` and zeroing passes are folded in.  The conservatives are double-buffered (q -> qn) because other
// tiles still read the pre-stage state of this tile's cells (the reference's in-place sweep is a Jacobi update,
// SURVEY.md 3.2).
//
// Inside a tile the faces are ordered for conflict-free shared-memory access: the tile's own faces grouped by
// their rank among the owner's faces (so consecutive threads have consecutive owners and, on a structured
// numbering, consecutive neighbours), then the incoming faces grouped the same way by their tile-side cell.
//
//   calc_VIS (cfd_v0.cpp:1744)                 -> k_tile_grad   (Green-Gauss gather from staged primitives)
//   one_rk_step_M1/_M2 (cfd_v0.cpp:2530/1897)  -> k_tile_stage  (flux + gather + sponge + RK update)
#pragma once
#include <cstddef>
#include <cstdint>

#include "kernels.cuh"
#include "lfmgpu.h"

namespace lfm {

struct TileDesc {
	int c0, nt;          // first cell, number of cells
	int halo_off, nh;    // halo cells: halo_cell[halo_off .. +nh)
	int f_off, nfo;      // tile faces: tables[f_off .. f_off+nfo+ninc), own faces first
	int ninc, hb;        // hb: staged index of the first halo cell (own cells are staged at [0, nt), halo cells at [hb, hb + nh);
	                     // hb = the submesh's tile size, so that a fixed-height TMA box of own cells never lands on a halo slot)
};

template <class R> struct TileView {
	const TileDesc* tiles;
	const int* halo_cell;            // global cell id of each halo entry (ascending inside a tile)
	const uint32_t* f_idx;           // [T] staged index of the owner | neighbour << 16 | physical-ghost flag << 31
	const int* f_gface;              // [T] mesh face id (only read for physical-boundary faces: owner->ghost vector)
	const R *fS, *fK;                // [D][T] area vector, non-orthogonal split K = S - delta
	const R *fw, *fdm, *fdi, *fSmag; // [T] weight_linear, |delta|, 1/|d|, |S|
	size_t T;                        // stride of the face tables
	const int16_t* csr_local;        // [F][n_cells] +-(tile-local face index + 1), ascending mesh face id, 0-padded
	int smax, fmax;                  // shared-memory strides of this launch (staged cells, faces)
};

// per-face constants (make_geo) computed once on the device with the same expressions the flux loops use
template <class R, int D> __global__ void __launch_bounds__(kBlock) k_face_geo(DevMesh<R> m, R* gK, R* g_delta_mag, R* g_dmag_inv, R* g_Smag) {
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= m.n_faces) return;
	R S[D], dv[D];
#pragma unroll
	for (int i = 0; i < D; i++) {
		S[i] = m.S[i * m.nfs + f];
		dv[i] = m.d[i * m.nfs + f];
	}
	FaceGeo<R, D> g;
	make_geo<R, D>(S, dv, m.w[f], g);
#pragma unroll
	for (int i = 0; i < D; i++) gK[i * m.nfs + f] = g.K[i];
	g_delta_mag[f] = g.delta_mag;
	g_dmag_inv[f] = g.dmag_inv;
	g_Smag[f] = g.S_mag;
}

// mesh-face constants -> tile-ordered tables (entry j describes mesh face f_gface[j])
template <class R, int D>
__global__ void __launch_bounds__(kBlock) k_tile_face_tables(DevMesh<R> m, const R* gK, const R* g_delta_mag, const R* g_dmag_inv, const R* g_Smag, const int* __restrict__ f_gface, size_t n,
                                                             size_t T, R* fS, R* fK, R* fw, R* fdm, R* fdi, R* fSmag) {
	const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const int f = f_gface[j];
#pragma unroll
	for (int i = 0; i < D; i++) {
		fS[i * T + j] = m.S[i * m.nfs + f];
		fK[i * T + j] = gK[i * m.nfs + f];
	}
	fw[j] = m.w[f];
	fdm[j] = g_delta_mag[f];
	fdi[j] = g_dmag_inv[f];
	fSmag[j] = g_Smag[f];
}

template <int D> struct StagedLayout {
	static constexpr int NQ = D + 2;
	static constexpr int RHO_INV = NQ, RPSI = NQ + 1, AUX = NQ + 2, DUDX = NQ + 3, DTDX = DUDX + D * D, SIGMAU = DTDX + D;
	static constexpr int NS = SIGMAU + D;
	static constexpr int TAUMC = NS;             // staged only with the Smagorinsky closure
	static constexpr int NS_LES = NS + D * D;
	// solver 2 (M2-AUSM): the minmod gradients of calc_gradients_M2AUSM ride along (instead of tauMC: laminar closure only)
	static constexpr int GRHO = NS, GP = NS + D, GU = NS + 2 * D;
	static constexpr int NS_AUSM = NS + 2 * D + D * D;
	__host__ __device__ static constexpr int rows(int mode) { return mode == 1 ? NS_LES : (mode == 2 ? NS_AUSM : NS); }
};

// a side of a face staged in shared memory (SoA rows of stride smax); values are loaded where they are used
template <class R, int D> struct SmemSide {
	static constexpr bool kHasTrace = false;
	const R* st;
	int smax, i;
	using L = StagedLayout<D>;
	__device__ __forceinline__ R q(int k) const { return st[k * smax + i]; }
	__device__ __forceinline__ R rho_inv() const { return st[L::RHO_INV * smax + i]; }
	__device__ __forceinline__ R Rpsi() const { return st[L::RPSI * smax + i]; }
	__device__ __forceinline__ R aux() const { return st[L::AUX * smax + i]; }
	__device__ __forceinline__ R dudx(int a, int b) const { return st[(L::DUDX + a * D + b) * smax + i]; }
	__device__ __forceinline__ R dTdx(int a) const { return st[(L::DTDX + a) * smax + i]; }
	__device__ __forceinline__ R sigmaU(int a) const { return st[(L::SIGMAU + a) * smax + i]; }
	__device__ __forceinline__ R tauMC(int a, int b) const { return st[(L::TAUMC + a * D + b) * smax + i]; }
	__device__ __forceinline__ R g_rho(int a) const { return st[(L::GRHO + a) * smax + i]; }
	__device__ __forceinline__ R g_p(int a) const { return st[(L::GP + a) * smax + i]; }
	__device__ __forceinline__ R g_U(int a, int b) const { return st[(L::GU + a * D + b) * smax + i]; }
};

// The AUSM+up pressure term of one face of solver 2 with both sides staged in shared memory (ausm_pressure_term of
// device_math.cuh fed from the staged rows; dv = owner -> neighbour vector of the face).
template <class R, int D> __device__ __forceinline__ R ausm_extra_staged(const Consts<R>& k, const SmemSide<R, D>& c, const SmemSide<R, D>& n, const R* S, const R* dv, R weight) {
	R cq[D + 2], nq[D + 2], cd[D][D], nd[D][D], cUg[D][D], nUg[D][D], crg[D], nrg[D], cpg[D], npg[D];
#pragma unroll
	for (int i = 0; i < D + 2; i++) {
		cq[i] = c.q(i);
		nq[i] = n.q(i);
	}
#pragma unroll
	for (int i = 0; i < D; i++) {
		crg[i] = c.g_rho(i);
		nrg[i] = n.g_rho(i);
		cpg[i] = c.g_p(i);
		npg[i] = n.g_p(i);
#pragma unroll
		for (int j = 0; j < D; j++) {
			cd[i][j] = c.dudx(i, j);
			nd[i][j] = n.dudx(i, j);
			cUg[i][j] = c.g_U(i, j);
			nUg[i][j] = n.g_U(i, j);
		}
	}
	return ausm_pressure_term<R, D>(k, cq, nq, c.Rpsi(), n.Rpsi(), cd, nd, cUg, nUg, crg, nrg, cpg, npg, S, dv, weight);
}

constexpr int kMaxSlots = 6;   // FACE_CNT of the reference is at most 6 (hexahedra)

// asynchronous global -> shared copy of one element (LDGSTS); completion: cp_async_wait_all()
template <class T> __device__ __forceinline__ void cp_async_elem(T* dst_smem, const T* src) {
	const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
	if (sizeof(T) == 8)
		asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
	else
		asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
	asm volatile("cp.async.commit_group;" ::: "memory");
	asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// calc_VIS on one tile: dudx, dTdx, sigmaU of the tile's cells
//   A. (rhoU, 1/rho, Rpsi) of tile + halo cells and (S, w, owner/neighbour staged index) of the tile's faces go
//      to shared memory with asynchronous copies;
//   C. per cell, ordered Green-Gauss gather from shared memory, then the tau/sigmaU block of calc_VIS.
// ---------------------------------------------------------------------------------------------------
// SMAX / FMAX > 0: compile-time shared-memory strides (row offsets become immediates of the LDS/STS instructions
// instead of one address computation per access); 0: the strides of the launch (tv.smax / tv.fmax).
constexpr int kFixedSmax = 304, kFixedFmax = 480;   // (304 staged cells: six ring slots of the persistent gradient kernel still fit)

template <class R, int D, int NT, int SMAX = 0, int FMAX = 0, int LES = 0>
__global__ void __launch_bounds__(NT) k_tile_grad(DevMesh<R> m, TileView<R> tv, const R* __restrict__ q, int tile0) {
	constexpr int QW = Rec<D>::QW;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int smax = SMAX ? SMAX : tv.smax, fmax = FMAX ? FMAX : tv.fmax;
	R* pr = reinterpret_cast<R*>(smem_raw);      // [D+2][smax]: rhoU, 1/rho, Rpsi
	R* fg = pr + (size_t)(D + 2) * smax;         // [D+1][fmax]: S, w
	uint32_t* fi = reinterpret_cast<uint32_t*>(fg + (size_t)(D + 1) * fmax);   // [fmax]: owner | neighbour << 16 (staged indices)
	const TileDesc td = tv.tiles[tile0 + blockIdx.x];
	const int ns = td.nt + td.nh;
	const int nf = td.nfo + td.ninc;
	for (int j = threadIdx.x; j < ns; j += NT) {
		const int x = j < td.nt ? td.c0 + j : tv.halo_cell[td.halo_off + j - td.nt];
		const int i = j < td.nt ? j : td.hb + j - td.nt;   // staged index
#pragma unroll
		for (int k = 0; k < D; k++) cp_async_elem(pr + k * smax + i, q + (size_t)x * QW + k + 1);
		cp_async_elem(pr + D * smax + i, q + (size_t)x * QW + Rec<D>::RHO_INV);
		cp_async_elem(pr + (D + 1) * smax + i, q + (size_t)x * QW + Rec<D>::RPSI);
	}
	for (int lf = threadIdx.x; lf < nf; lf += NT) {
		const size_t j = (size_t)td.f_off + lf;
#pragma unroll
		for (int k = 0; k < D; k++) cp_async_elem(fg + k * fmax + lf, tv.fS + k * tv.T + j);
		cp_async_elem(fg + D * fmax + lf, tv.fw + j);
		cp_async_elem(fi + lf, tv.f_idx + j);
	}
	// what the gather needs from global memory, requested before the wait
	const int lc0 = threadIdx.x;
	int e0[kMaxSlots];
	R vinv0 = R(0), cq0[D + 1];
	if (lc0 < td.nt) {
		const int c = td.c0 + lc0;
#pragma unroll
		for (int s = 0; s < kMaxSlots; s++) e0[s] = s < m.F ? (int)tv.csr_local[(size_t)s * m.n_cells + c] : 0;
		vinv0 = m.vol_inv[c];
#pragma unroll
		for (int k = 0; k < D + 1; k++) cq0[k] = q[(size_t)c * QW + k];
	}
	cp_async_wait_all();
	__syncthreads();
	for (int lc = threadIdx.x; lc < td.nt; lc += NT) {
		const int c = td.c0 + lc;
		int e[kMaxSlots];
		R vinv, cq[D + 1];
		if (lc == lc0) {
#pragma unroll
			for (int s = 0; s < kMaxSlots; s++) e[s] = e0[s];
			vinv = vinv0;
#pragma unroll
			for (int k = 0; k < D + 1; k++) cq[k] = cq0[k];
		} else {
#pragma unroll
			for (int s = 0; s < kMaxSlots; s++) e[s] = s < m.F ? (int)tv.csr_local[(size_t)s * m.n_cells + c] : 0;
			vinv = m.vol_inv[c];
#pragma unroll
			for (int k = 0; k < D + 1; k++) cq[k] = q[(size_t)c * QW + k];
		}
		R cU[D];
		{
			const R cri = pr[D * smax + lc];
#pragma unroll
			for (int k = 0; k < D; k++) cU[k] = pr[k * smax + lc] * cri;
		}
		const R c_Rpsi = pr[(D + 1) * smax + lc];
		R dudx[D][D], dTdx[D];
#pragma unroll
		for (int i = 0; i < D; i++) {
			dTdx[i] = R(0);
#pragma unroll
			for (int j = 0; j < D; j++) dudx[i][j] = R(0);
		}
#pragma unroll
		for (int s = 0; s < kMaxSlots; s++) {
			if (e[s] == 0) break;
			const bool own = e[s] > 0;
			const int lf = (own ? e[s] : -e[s]) - 1;
			const uint32_t idx = fi[lf];
			const int lo = own ? (int)((idx >> 16) & 0x7fffu) : (int)(idx & 0xffffu);   // the other side
			R oU[D];
			{
				const R ori = pr[D * smax + lo];
#pragma unroll
				for (int k = 0; k < D; k++) oU[k] = pr[k * smax + lo] * ori;
			}
			const R o_Rpsi = pr[(D + 1) * smax + lo];
			const R w = fg[D * fmax + lf];
			R face_U[D], face_T, sov[D];
			if (own) {
				grad_face_values<R, D>(m.k, w, cU, c_Rpsi, oU, o_Rpsi, face_U, face_T);
#pragma unroll
				for (int i = 0; i < D; i++) sov[i] = fg[i * fmax + lf] * vinv;
			} else {
				grad_face_values<R, D>(m.k, w, oU, o_Rpsi, cU, c_Rpsi, face_U, face_T);
#pragma unroll
				for (int i = 0; i < D; i++) sov[i] = -fg[i * fmax + lf] * vinv;
			}
#pragma unroll
			for (int i = 0; i < D; i++) {
#pragma unroll
				for (int j = 0; j < D; j++) dudx[i][j] += face_U[i] * sov[j];
				dTdx[i] += face_T * sov[i];
			}
		}
		// the tau / sigmaU block of calc_VIS (U = rhoU / rho: a true division, cfd_v0.cpp:1806-1857)
		R Ud[D], tau[D][D], sigmaU[D];
#pragma unroll
		for (int i = 0; i < D; i++) Ud[i] = cq[i + 1] / cq[0];
		stress<R, D>(m.k, dudx, tau);
#pragma unroll
		for (int i = 0; i < D; i++) sigmaU[i] = dotD<R, D>(Ud, tau[i]);
		store_vis_record<R, D>(m.vis, (size_t)c, dudx, dTdx, sigmaU);
		if (LES) {
			R tauMC[D][D];
			tauMC_smagorinsky<R, D>(m.k, m.smag_c[c], dudx, tauMC);
#pragma unroll
			for (int i = 0; i < D; i++)
#pragma unroll
				for (int j = 0; j < D; j++) m.tauMC[(size_t)(i * D + j) * m.ncs + c] = tauMC[i][j];
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// one_rk_step_M1/_M2 on one tile (flux, gather, sponge, RK update)
// ---------------------------------------------------------------------------------------------------
// everything phase B needs about one face, fetched from the tile-ordered tables ahead of its use
template <class R, int D> struct FaceIn {
	uint32_t idx;
	FaceGeo<R, D> g;
};
template <class R, int D, int SCHEME> __device__ __forceinline__ void fetch_face(const TileView<R>& tv, size_t j, FaceIn<R, D>& in) {
	in.idx = tv.f_idx[j];
#pragma unroll
	for (int i = 0; i < D; i++) {
		in.g.S[i] = tv.fS[i * tv.T + j];
		in.g.K[i] = tv.fK[i * tv.T + j];
	}
	in.g.w = tv.fw[j];
	in.g.delta_mag = tv.fdm[j];
	in.g.dmag_inv = tv.fdi[j];
	in.g.S_mag = SCHEME == 0 ? tv.fSmag[j] : R(0);
}

// Phase C of the stage kernel for components [I0, I1) of tile cell lc: dq = A_k dq + sum_faces(+-dt rhs / V) + sponge,
// q_new = q + B_k dq (prepare_for_RKstep's scaling, the scatter of cfd_v0.cpp:2792-2804 as an ordered gather, sponge
// 2810-2814, update 2819-2822).  `x -= y` is evaluated as `x += (-y)`: bit-identical, and branch-free on the sign.
// The global loads are issued before the CTA barrier that closes phase B (sync != 0) so they travel during the wait.
template <class R, int D, int I0, int I1>
__device__ __forceinline__ void gather_update(const DevMesh<R>& m, const TileView<R>& tv, const TileDesc& td, R* st, const R* fl, int smax, int fmax, R* __restrict__ qn, int lc, bool active,
                                              bool sync, R dt, R Ak, R Bk, int first, int res) {
	constexpr int NQ = D + 2;
	const int c = td.c0 + (active ? lc : 0);
	int e[kMaxSlots];
	R dq[NQ], vinv = R(0), sg = R(0);
#pragma unroll
	for (int s = 0; s < kMaxSlots; s++) e[s] = 0;
#pragma unroll
	for (int i = 0; i < NQ; i++) dq[i] = R(0);
	if (active) {
#pragma unroll
		for (int s = 0; s < kMaxSlots; s++)
			if (s < m.F) e[s] = (int)tv.csr_local[(size_t)s * m.n_cells + c];
		if (!first) {
#pragma unroll
			for (int i = I0; i < I1; i++) dq[i] = m.dq[(size_t)i * m.n_cells + c];
		}
		vinv = m.vol_inv[c];
		sg = m.sigma[c];
	}
	if (sync) __syncthreads();
	if (!active) return;
	R RES[NQ];
#pragma unroll
	for (int i = I0; i < I1; i++) {
		dq[i] *= Ak;
		RES[i] = R(0);
	}
#pragma unroll
	for (int s = 0; s < kMaxSlots; s++) {
		if (e[s] == 0) break;
		const bool own = e[s] > 0;
		const int lfc = (own ? e[s] : -e[s]) - 1;
#pragma unroll
		for (int i = I0; i < I1; i++) {
			const R v = fl[i * fmax + lfc];
			const R rr = own ? v : -v;
			if (res) RES[i] += rr;
			dq[i] += dt * rr * vinv;
		}
	}
#pragma unroll
	for (int i = I0; i < I1; i++) {
		const R cqi = st[i * smax + lc];
		const R target = i == 0 ? m.k.rhoInf : (i == NQ - 1 ? m.k.rhoEInf : m.k.rhoUInf[i > 0 && i < NQ - 1 ? i - 1 : 0]);
		dq[i] += dt * sg * (target - cqi);
		m.dq[(size_t)i * m.n_cells + c] = dq[i];
		const R qi = cqi + Bk * dq[i];
		qn[(size_t)c * Rec<D>::QW + i] = qi;
		st[i * smax + lc] = qi;      // only this cell's own threads touch these slots after phase B
		if (res) m.RES[(size_t)i * m.n_cells + c] = RES[i];
	}
}

// EXTRA: rows staged on top of StagedLayout::NS -- 0 none, 1 tauMC (Smagorinsky closure), 2 the minmod gradients of solver 2 (M2-AUSM)
template <class R, int D, int SCHEME, int NT, int MINB, int SMAX = 0, int FMAX = 0, int EXTRA = 0>
__global__ void __launch_bounds__(NT, MINB)
    k_tile_stage(DevMesh<R> m, TileView<R> tv, const R* __restrict__ q, R* __restrict__ qn, int tile0, R dt, R Ak, R Bk, int first, int res) {
	using L = StagedLayout<D>;
	constexpr int NQ = D + 2, QW = Rec<D>::QW, VW = Rec<D>::VW;
	static_assert(NT % 64 == 0, "phase C splits the CTA into two halves of whole warps (a barrier sits inside each half's code path)");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int smax = SMAX ? SMAX : tv.smax, fmax = FMAX ? FMAX : tv.fmax;
	R* st = reinterpret_cast<R*>(smem_raw);      // [NS (+D*D)][smax]
	R* fl = st + (size_t)L::rows(EXTRA) * smax;   // [NQ][fmax]
	const TileDesc td = tv.tiles[tile0 + blockIdx.x];
	const int ns = td.nt + td.nh;
	const int nf = td.nfo + td.ninc;

	// ---- A: request the cell states (asynchronous copies, no registers) ----------------------------------
	for (int j = threadIdx.x; j < ns; j += NT) {
		const int x = j < td.nt ? td.c0 + j : tv.halo_cell[td.halo_off + j - td.nt];
		const int i = j < td.nt ? j : td.hb + j - td.nt;   // staged index
		// the staged rows [0, NS) are the values of the cell's Q record (q | 1/rho, Rpsi, aux) followed by its V record
		// (dudx | dTdx | sigmaU): StagedLayout and Rec order them alike
#pragma unroll
		for (int k = 0; k < NQ + 3; k++) cp_async_elem(st + k * smax + i, q + (size_t)x * QW + (k < NQ ? k : Rec<D>::RHO_INV + k - NQ));
#pragma unroll
		for (int k = 0; k < D * D + 2 * D; k++) cp_async_elem(st + (L::DUDX + k) * smax + i, m.vis + (size_t)x * VW + k);
		if (EXTRA == 1) {
#pragma unroll
			for (int k = 0; k < D * D; k++) cp_async_elem(st + (L::TAUMC + k) * smax + i, m.tauMC + (size_t)k * m.ncs + x);
		}
		if (EXTRA == 2) {   // solver 2: minmod gradients (ghost slots hold zeros, as in the reference)
#pragma unroll
			for (int k = 0; k < D; k++) {
				cp_async_elem(st + (L::GRHO + k) * smax + i, m.g_rho + (size_t)k * m.ncs + x);
				cp_async_elem(st + (L::GP + k) * smax + i, m.g_p + (size_t)k * m.ncs + x);
			}
#pragma unroll
			for (int k = 0; k < D * D; k++) cp_async_elem(st + (L::GU + k) * smax + i, m.g_U + (size_t)k * m.ncs + x);
		}
	}
	// the first face's constants travel while the copies land
	FaceIn<R, D> cur;
	if ((int)threadIdx.x < nf) fetch_face<R, D, SCHEME>(tv, (size_t)td.f_off + threadIdx.x, cur);
	cp_async_wait_all();
	__syncthreads();

	// ---- B: every face of the tile once ----------------------------------------------------------------
	for (int lf = threadIdx.x; lf < nf; lf += NT) {
		FaceIn<R, D> nxt;
		if (lf + NT < nf) fetch_face<R, D, SCHEME>(tv, (size_t)td.f_off + lf + NT, nxt);
		const int lo = (int)(cur.idx & 0xffffu), ln = (int)((cur.idx >> 16) & 0x7fffu);
		const bool ghost = (cur.idx >> 31) != 0;
		R dv[D];
#pragma unroll
		for (int i = 0; i < D; i++) dv[i] = R(0);
		if (ghost || EXTRA == 2) {   // solver 2 reconstructs along d on every face
			const int f = tv.f_gface[(size_t)td.f_off + lf];
#pragma unroll
			for (int i = 0; i < D; i++) dv[i] = m.d[i * m.nfs + f];
		}
		R rhs[NQ];
		R extra = R(0);
		if (EXTRA == 2) extra = ausm_extra_staged<R, D>(m.k, SmemSide<R, D>{st, smax, lo}, SmemSide<R, D>{st, smax, ln}, cur.g.S, dv, cur.g.w);
		face_flux<R, D, SCHEME, SmemSide<R, D>, SmemSide<R, D>, (EXTRA == 1 ? 1 : 0)>(m.k, SmemSide<R, D>{st, smax, lo}, SmemSide<R, D>{st, smax, ln}, cur.g, ghost, dv, rhs, extra);
#pragma unroll
		for (int i = 0; i < NQ; i++) fl[i * fmax + lf] = rhs[i];
		if (lf + NT < nf) cur = nxt;
	}

	// ---- C: ordered gather, sponge, RK update ----------------------------------------------------------
	// When the tile leaves half the CTA idle the components of a cell are split between the two halves of the CTA
	// (warp-uniform, so neither half pays for the other's components): the ordered sums are per component, so the
	// split does not change any result.
	const bool split = 2 * td.nt <= NT;
	const int rounds = split ? 1 : (td.nt + NT - 1) / NT;   // the same trip count for every thread of the CTA
	for (int r = 0; r < rounds; r++) {
		constexpr int H = (NQ + 1) / 2;
		const int part = split ? (threadIdx.x >= NT / 2 ? 1 : 0) : 2;           // 0: [0,H)  1: [H,NQ)  2: all
		const int lc = split ? (int)threadIdx.x - (part ? NT / 2 : 0) : (int)threadIdx.x + r * NT;
		const bool active = lc < td.nt;
		if (part == 0)
			gather_update<R, D, 0, H>(m, tv, td, st, fl, smax, fmax, qn, lc, active, r == 0, dt, Ak, Bk, first, res);
		else if (part == 1)
			gather_update<R, D, H, NQ>(m, tv, td, st, fl, smax, fmax, qn, lc, active, r == 0, dt, Ak, Bk, first, res);
		else
			gather_update<R, D, 0, NQ>(m, tv, td, st, fl, smax, fmax, qn, lc, active, r == 0, dt, Ak, Bk, first, res);
	}
	// derived values of the NEW state travel with it (1/rho, R*psi, c or H): the next stage copies them
	__syncthreads();
	for (int lc = threadIdx.x; lc < td.nt; lc += NT) {
		CellState<R, D> s;
#pragma unroll
		for (int i = 0; i < NQ; i++) s.q[i] = st[i * smax + lc];
		derive_state<R, D, SCHEME>(m.k, s);
		R* rec = qn + (size_t)(td.c0 + lc) * QW;
		rec[Rec<D>::RHO_INV] = s.rho_inv;
		rec[Rec<D>::RPSI] = s.Rpsi;
		rec[Rec<D>::AUX] = s.aux;
	}
}

// host-side description of the plan (device arrays are owned by the handle's allocation list)
struct TilePlan {
	bool ready = false;
	int n_tiles = 0, tile_cells = 0;
	int sub_tile_start[LFMGPU_MAX_SUBMESH + 1] = {0};
	int sub_smax[LFMGPU_MAX_SUBMESH] = {0}, sub_fmax[LFMGPU_MAX_SUBMESH] = {0};
	int sub_tc[LFMGPU_MAX_SUBMESH] = {0}, sub_hmax[LFMGPU_MAX_SUBMESH] = {0};   // tile size (= halo base hb) and largest halo of each submesh
	size_t smem_bytes = 0;            // largest k_tile_stage request
	double halo_face_ratio = 0.0;     // incoming faces / own faces (redundant flux evaluations)
	double halo_cell_ratio = 0.0;     // halo cells / tile cells
	size_t T = 0, n_table = 0;        // stride and entries of the tile-ordered face tables
	TileDesc* d_tiles = nullptr;
	int* d_halo_cell = nullptr;
	uint32_t* d_f_idx = nullptr;
	int* d_f_gface = nullptr;
	int16_t* d_csr_local = nullptr;
	void *d_fS = nullptr, *d_fK = nullptr, *d_fw = nullptr, *d_fdm = nullptr, *d_fdi = nullptr, *d_fSmag = nullptr;
};

}  // namespace lfm
