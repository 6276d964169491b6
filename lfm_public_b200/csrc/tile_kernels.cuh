// Fused shared-memory tile kernels ("v2"): the production path of the per-iteration solve.
//
// A tile is a run of consecutive cells in traversal order (never straddling a submesh).  One CTA owns one
// tile per launch and
//   A. stages the state of the tile's cells AND of every cell across one of their faces (the halo) in shared
//      memory, evaluating the per-cell divisions / square roots once per staged cell,
//   B. evaluates every face that touches a tile cell exactly once per tile: the tile's own faces (a contiguous
//      face range, because faces are numbered by owner) plus the "incoming" faces owned by earlier cells outside
//      the tile -- the only redundant flux evaluations, a surface-to-volume effect of the cell numbering,
//   C. gathers per cell in ascending face id (the reference's summation order, kernels.cuh) from shared memory,
//      adds the sponge term and applies the low-storage RK update, writing q_new / dq (and RES) coalesced.
// Nothing but the final cell state goes back to HBM: face fluxes never leave the SM, prepare_for_RKstep's
// `dq *= A_k` and zeroing passes are folded in.  The conservatives are double-buffered (q -> qn) because other
// tiles still read the pre-stage state of this tile's cells (the reference's in-place sweep is a Jacobi update,
// SURVEY.md 3.2).
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
	int fo0, nfo;        // own faces: [fo0, fo0+nfo)
	int inc_off, ninc;   // incoming faces: inc_face[inc_off .. +ninc)
};

template <class R> struct TileView {
	const TileDesc* tiles;
	const int* halo_cell;            // global cell id of each halo entry (ascending inside a tile)
	const int* inc_face;             // global face id of each incoming face (ascending inside a tile)
	const uint16_t* inc_lowner;      // staged index of the incoming face's owner
	const uint16_t* face_lneigh;     // [n_faces] staged index of the neighbour inside the owner's tile; bit 15: physical ghost
	const int16_t* csr_local;        // [F][n_cells] +-(tile-local face index + 1), ascending face id, 0-padded
	const R *gK, *g_delta_mag, *g_dmag_inv, *g_Smag;   // per-face constants: [D][nfs], [nfs], [nfs], [nfs]
	int smax, fmax;                  // shared-memory strides of this launch (staged cells, faces)
	int prefetch_distance;           // tiles ahead whose rows are pulled into L2 (0: off)
	int n_launch_tiles;              // tiles of this launch
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

template <int D> struct StagedLayout {
	static constexpr int NQ = D + 2;
	static constexpr int RHO_INV = NQ, RPSI = NQ + 1, AUX = NQ + 2, DUDX = NQ + 3, DTDX = DUDX + D * D, SIGMAU = DTDX + D;
	static constexpr int NS = SIGMAU + D;
};

// a side of a face staged in shared memory (SoA rows of stride smax); values are loaded where they are used
template <class R, int D> struct SmemSide {
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
};

constexpr int kMaxSlots = 6;   // FACE_CNT of the reference is at most 6 (hexahedra)

// ---------------------------------------------------------------------------------------------------
// calc_VIS on one tile: dudx, dTdx, sigmaU of the tile's cells
//   A. primitives (U, Rpsi) of tile + halo cells and (w, S, owner/neighbour staged index) of the tile's faces
//      go to shared memory with independent, coalesced loads;
//   C. per cell, ordered Green-Gauss gather from shared memory, then the tau/sigmaU block of calc_VIS.
// ---------------------------------------------------------------------------------------------------
template <class R, int D, int NT> __global__ void __launch_bounds__(NT) k_tile_grad(DevMesh<R> m, TileView<R> tv, const R* __restrict__ q, int tile0) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int smax = tv.smax, fmax = tv.fmax;
	R* pr = reinterpret_cast<R*>(smem_raw);      // [D+1][smax]: U, Rpsi
	R* fg = pr + (size_t)(D + 1) * smax;         // [D+1][fmax]: S, w
	uint32_t* fi = reinterpret_cast<uint32_t*>(fg + (size_t)(D + 1) * fmax);   // [fmax]: owner | neighbour << 16 (staged indices)
	const TileDesc td = tv.tiles[tile0 + blockIdx.x];
	const int ns = td.nt + td.nh;
	const int nf = td.nfo + td.ninc;
	for (int i = threadIdx.x; i < ns; i += NT) {
		const int x = i < td.nt ? td.c0 + i : tv.halo_cell[td.halo_off + i - td.nt];
		R cq[D + 2], U[D], rho_inv, Rpsi, T;
#pragma unroll
		for (int k = 0; k < D + 2; k++) cq[k] = q[k * m.ncs + x];
		primitives<R, D>(m.k, cq, rho_inv, U, Rpsi, T);
#pragma unroll
		for (int k = 0; k < D; k++) pr[k * smax + i] = U[k];
		pr[D * smax + i] = Rpsi;
	}
	for (int lf = threadIdx.x; lf < nf; lf += NT) {
		int f;
		uint32_t lo, ln;
		if (lf < td.nfo) {
			f = td.fo0 + lf;
			lo = (uint32_t)(m.face_owner[f] - td.c0);
			ln = tv.face_lneigh[f] & 0x7fffu;
		} else {
			const int k = td.inc_off + lf - td.nfo;
			f = tv.inc_face[k];
			lo = tv.inc_lowner[k];
			ln = (uint32_t)(m.face_neigh[f] - td.c0);
		}
#pragma unroll
		for (int k = 0; k < D; k++) fg[k * fmax + lf] = m.S[k * m.nfs + f];
		fg[D * fmax + lf] = m.w[f];
		fi[lf] = lo | (ln << 16);
	}
	__syncthreads();
	for (int lc = threadIdx.x; lc < td.nt; lc += NT) {
		const int c = td.c0 + lc;
		int e[kMaxSlots];
#pragma unroll
		for (int s = 0; s < kMaxSlots; s++) e[s] = s < m.F ? (int)tv.csr_local[(size_t)s * m.n_cells + c] : 0;
		const R vinv = m.vol_inv[c];
		R cq[D + 2];
#pragma unroll
		for (int k = 0; k < D + 2; k++) cq[k] = q[k * m.ncs + c];
		R cU[D];
#pragma unroll
		for (int k = 0; k < D; k++) cU[k] = pr[k * smax + lc];
		const R c_Rpsi = pr[D * smax + lc];
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
			const int lo = own ? (int)(idx >> 16) : (int)(idx & 0xffffu);   // the other side
			R oU[D];
#pragma unroll
			for (int k = 0; k < D; k++) oU[k] = pr[k * smax + lo];
			const R o_Rpsi = pr[D * smax + lo];
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
		R tauMC[D][D], sigmaU[D];
		vis_cell_terms<R, D>(m.k, cq, dudx, tauMC, sigmaU);
#pragma unroll
		for (int i = 0; i < D; i++) {
#pragma unroll
			for (int j = 0; j < D; j++) m.dudx[(size_t)(i * D + j) * m.ncs + c] = dudx[i][j];
			m.dTdx[(size_t)i * m.ncs + c] = dTdx[i];
			m.sigmaU[(size_t)i * m.ncs + c] = sigmaU[i];
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// one_rk_step_M1/_M2 on one tile (flux, gather, sponge, RK update)
// ---------------------------------------------------------------------------------------------------
template <class R, int D> __device__ __forceinline__ void staged_store(R* st, int smax, int i, const CellState<R, D>& s) {
	using L = StagedLayout<D>;
#pragma unroll
	for (int k = 0; k < D + 2; k++) st[k * smax + i] = s.q[k];
	st[L::RHO_INV * smax + i] = s.rho_inv;
	st[L::RPSI * smax + i] = s.Rpsi;
	st[L::AUX * smax + i] = s.aux;
#pragma unroll
	for (int a = 0; a < D; a++) {
#pragma unroll
		for (int b = 0; b < D; b++) st[(L::DUDX + a * D + b) * smax + i] = s.dudx[a][b];
		st[(L::DTDX + a) * smax + i] = s.dTdx[a];
		st[(L::SIGMAU + a) * smax + i] = s.sigmaU[a];
	}
}

// L2 prefetch of `nrows` rows (row r starts at base + r*stride elements) of `n` elements starting at element `first`:
// one prefetch instruction per 128-byte line, spread over the CTA.  No registers are tied up and nothing waits.
template <class T> __device__ __forceinline__ void l2_prefetch_rows(const T* base, size_t stride, int nrows, int first, int n, int tid, int nthreads) {
	if (n <= 0) return;
	const int lines = (n * (int)sizeof(T) + 127) / 128 + 1;
	for (int i = tid; i < nrows * lines; i += nthreads) {
		const int r = i / lines, l = i - r * lines;
		const char* p = reinterpret_cast<const char*>(base + (size_t)r * stride + first) + (size_t)l * 128;
		asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
	}
}

// everything phase B needs about one face, fetched from global memory ahead of its use
template <class R, int D> struct FaceIn {
	int lo, ln;
	bool ghost;
	FaceGeo<R, D> g;
	R dv[D];
};
template <class R, int D, int SCHEME> __device__ __forceinline__ void fetch_face(const DevMesh<R>& m, const TileView<R>& tv, const TileDesc& td, int lf, FaceIn<R, D>& in) {
	int f;
	in.ghost = false;
	if (lf < td.nfo) {
		f = td.fo0 + lf;
		in.lo = m.face_owner[f] - td.c0;
		const unsigned v = tv.face_lneigh[f];
		in.ln = (int)(v & 0x7fffu);
		in.ghost = (v >> 15) != 0;
	} else {
		const int k = td.inc_off + lf - td.nfo;
		f = tv.inc_face[k];
		in.lo = tv.inc_lowner[k];
		in.ln = m.face_neigh[f] - td.c0;
	}
#pragma unroll
	for (int i = 0; i < D; i++) {
		in.g.S[i] = m.S[i * m.nfs + f];
		in.g.K[i] = tv.gK[i * m.nfs + f];
		in.dv[i] = R(0);
	}
	in.g.w = m.w[f];
	in.g.delta_mag = tv.g_delta_mag[f];
	in.g.dmag_inv = tv.g_dmag_inv[f];
	in.g.S_mag = SCHEME == 0 ? tv.g_Smag[f] : R(0);
	if (in.ghost) {
#pragma unroll
		for (int i = 0; i < D; i++) in.dv[i] = m.d[i * m.nfs + f];
	}
}

template <class R, int D, int SCHEME, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_tile_stage(DevMesh<R> m, TileView<R> tv, const R* __restrict__ q, R* __restrict__ qn, int tile0, R dt, R Ak, R Bk, int first, int res) {
	using L = StagedLayout<D>;
	constexpr int NQ = D + 2;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int smax = tv.smax, fmax = tv.fmax;
	R* st = reinterpret_cast<R*>(smem_raw);      // [NS][smax]
	R* fl = st + (size_t)L::NS * smax;           // [NQ][fmax]
	const TileDesc td = tv.tiles[tile0 + blockIdx.x];
	const int ns = td.nt + td.nh;
	const int nf = td.nfo + td.ninc;

	// Pull the streamed rows of a tile that will be scheduled a little later into L2 (its cells' state and its own
	// faces' constants), so that its staging phase meets L2 latency instead of HBM latency.
	if (tv.prefetch_distance > 0 && (int)blockIdx.x + tv.prefetch_distance < tv.n_launch_tiles) {
		const TileDesc pd = tv.tiles[tile0 + blockIdx.x + tv.prefetch_distance];
		const int t = threadIdx.x;
		l2_prefetch_rows<R>(q, m.ncs, NQ, pd.c0, pd.nt, t, NT);
		l2_prefetch_rows<R>(m.dudx, m.ncs, D * D, pd.c0, pd.nt, t, NT);
		l2_prefetch_rows<R>(m.dTdx, m.ncs, D, pd.c0, pd.nt, t, NT);
		l2_prefetch_rows<R>(m.sigmaU, m.ncs, D, pd.c0, pd.nt, t, NT);
		if (!first) l2_prefetch_rows<R>(m.dq, (size_t)m.n_cells, NQ, pd.c0, pd.nt, t, NT);
		l2_prefetch_rows<R>(m.vol_inv, 0, 1, pd.c0, pd.nt, t, NT);
		l2_prefetch_rows<R>(m.sigma, 0, 1, pd.c0, pd.nt, t, NT);
		l2_prefetch_rows<R>(m.S, m.nfs, D, pd.fo0, pd.nfo, t, NT);
		l2_prefetch_rows<R>(tv.gK, m.nfs, D, pd.fo0, pd.nfo, t, NT);
		l2_prefetch_rows<R>(m.w, 0, 1, pd.fo0, pd.nfo, t, NT);
		l2_prefetch_rows<R>(tv.g_delta_mag, 0, 1, pd.fo0, pd.nfo, t, NT);
		l2_prefetch_rows<R>(tv.g_dmag_inv, 0, 1, pd.fo0, pd.nfo, t, NT);
	}

	// ---- A: stage cell states ------------------------------------------------------------------------
	for (int i = threadIdx.x; i < ns; i += NT) {
		const int x = i < td.nt ? td.c0 + i : tv.halo_cell[td.halo_off + i - td.nt];
		CellState<R, D> s;
		load_state<R, D, SCHEME>(m, q, x, s);
		staged_store<R, D>(st, smax, i, s);
	}
	__syncthreads();

	// ---- B: every face of the tile once ----------------------------------------------------------------
	for (int lf = threadIdx.x; lf < nf; lf += NT) {
		FaceIn<R, D> cur;
		fetch_face<R, D, SCHEME>(m, tv, td, lf, cur);
		R rhs[NQ];
		face_flux<R, D, SCHEME>(m.k, SmemSide<R, D>{st, smax, cur.lo}, SmemSide<R, D>{st, smax, cur.ln}, cur.g, cur.ghost, cur.dv, rhs);
#pragma unroll
		for (int i = 0; i < NQ; i++) fl[i * fmax + lf] = rhs[i];
	}

	// ---- C: ordered gather, sponge, RK update ----------------------------------------------------------
	// Two threads share a cell when the tile leaves half the CTA idle: the ordered sums are per component, so
	// splitting the components between threads does not change any result.
	const int parts = (2 * td.nt <= NT) ? 2 : 1;
	const int items = td.nt * parts;
	bool synced = false;
	for (int it = threadIdx.x; it < items || !synced; it += NT) {
		const bool active = it < items;
		const int lc = active ? (parts == 2 ? (it >> 1) : it) : 0;
		const int part = parts == 2 ? (it & 1) : 0;
		const int i0 = parts == 2 ? (part == 0 ? 0 : (NQ + 1) / 2) : 0;
		const int i1 = parts == 2 ? (part == 0 ? (NQ + 1) / 2 : NQ) : NQ;
		const int c = td.c0 + lc;
		int e[kMaxSlots];
		R dq[NQ], RES[NQ], vinv = R(0), sg = R(0);
		if (active) {
#pragma unroll
			for (int s = 0; s < kMaxSlots; s++) e[s] = s < m.F ? (int)tv.csr_local[(size_t)s * m.n_cells + c] : 0;
#pragma unroll
			for (int i = 0; i < NQ; i++) {
				dq[i] = (first || i < i0 || i >= i1) ? R(0) : m.dq[(size_t)i * m.n_cells + c] * Ak;
				RES[i] = R(0);
			}
			vinv = m.vol_inv[c];
			sg = m.sigma[c];
		}
		if (!synced) {          // the loads above are in flight while the CTA waits for phase B to finish
			__syncthreads();
			synced = true;
		}
		if (!active) continue;
#pragma unroll
		for (int s = 0; s < kMaxSlots; s++) {
			if (e[s] == 0) break;
			const int lfc = (e[s] > 0 ? e[s] : -e[s]) - 1;
			if (e[s] > 0) {
#pragma unroll
				for (int i = 0; i < NQ; i++)
					if (i >= i0 && i < i1) {
						const R r = fl[i * fmax + lfc];
						if (res) RES[i] += r;
						dq[i] += dt * r * vinv;
					}
			} else {
#pragma unroll
				for (int i = 0; i < NQ; i++)
					if (i >= i0 && i < i1) {
						const R r = fl[i * fmax + lfc];
						if (res) RES[i] -= r;
						dq[i] -= dt * r * vinv;
					}
			}
		}
#pragma unroll
		for (int i = 0; i < NQ; i++)
			if (i >= i0 && i < i1) {
				const R cqi = st[i * smax + lc];
				const R target = i == 0 ? m.k.rhoInf : (i == NQ - 1 ? m.k.rhoEInf : m.k.rhoUInf[i > 0 && i < NQ - 1 ? i - 1 : 0]);
				dq[i] += dt * sg * (target - cqi);
				m.dq[(size_t)i * m.n_cells + c] = dq[i];
				qn[i * m.ncs + c] = cqi + Bk * dq[i];
				if (res) m.RES[(size_t)i * m.n_cells + c] = RES[i];
			}
	}
}

// host-side description of the plan (device arrays are owned by the handle's allocation list)
struct TilePlan {
	bool ready = false;
	int n_tiles = 0, tile_cells = 0, threads = 128;
	int sub_tile_start[LFMGPU_MAX_SUBMESH + 1] = {0};
	int sub_smax[LFMGPU_MAX_SUBMESH] = {0}, sub_fmax[LFMGPU_MAX_SUBMESH] = {0};
	size_t smem_bytes = 0;            // largest k_tile_stage request
	double halo_face_ratio = 0.0;     // incoming faces / own faces (redundant flux evaluations)
	double halo_cell_ratio = 0.0;     // halo cells / tile cells
	TileDesc* d_tiles = nullptr;
	int* d_halo_cell = nullptr;
	int* d_inc_face = nullptr;
	uint16_t* d_inc_lowner = nullptr;
	uint16_t* d_face_lneigh = nullptr;
	int16_t* d_csr_local = nullptr;
	void *d_gK = nullptr, *d_delta_mag = nullptr, *d_dmag_inv = nullptr, *d_Smag = nullptr;
};

}  // namespace lfm
