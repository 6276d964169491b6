// Per-cell and per-face arithmetic of the hot path, written once and used by every kernel variant.
//
// Each function restates one block of the reference's stage loops with the SAME expression trees
// (operand order, reciprocal-vs-division, re-evaluated interpolation weights), so that the fp64
// instantiation, compiled with --fmad=false, reproduces the CPU reference bit for bit:
//   compute_Rpsi                 api/cfdv0_solver.h:252-261
//   tau / sigmaU / tauMC block   src/cfd_v0.cpp:1806-1857 (calc_VIS)
//   M1 central-upwind flux       src/cfd_v0.cpp:2586-2675 (one_rk_step_M1)
//   M2 Pirozzoli flux            src/cfd_v0.cpp:1963-2023 (one_rk_step_M2)
//   viscous face terms           src/cfd_v0.cpp:2678-2790 (== 2026-2137)
// The float instantiation computes in pure fp32 (the reference's float build promotes a few
// sub-expressions to double through literals; the north-star tolerance for fp32 is 1e-5).
#pragma once
#include <cuda_runtime.h>

namespace lfm {

template <class R> __device__ __forceinline__ R rsqrt_exact(R x);
template <> __device__ __forceinline__ double rsqrt_exact<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float rsqrt_exact<float>(float x) { return sqrtf(x); }
#define LFM_SQRT(x) rsqrt_exact<R>(x)

// INTERP_LINEAR(weight, a, b)  (api/cfdv0_solver.h:14)
template <class R> __device__ __forceinline__ R interp(R w, R a, R b) { return w * a + (R(1.0) - w) * b; }

template <class R, int D> __device__ __forceinline__ R dotD(const R* l, const R* r) {
	R s = l[0] * r[0];
#pragma unroll
	for (int i = 1; i < D; i++) s += l[i] * r[i];
	return s;
}

// Gas / scheme constants in the working precision (lfmgpu_consts narrowed once at create time).
template <class R> struct Consts {
	R gamma, gm1, Rgas_inv, mu, Cp, Pr_inv, kappa;   // kappa = Cp*mu*Pr_inv
	R c_tau;                                         // 2.0/3.0*mu      (tau diagonal scale)
	R c_diag;                                        // mu*2.0/3.0      (tauMC diagonal scale)
	R rhoInf, UInf[3], EInf, pInf, TInf;
	R rhoUInf[3], rhoEInf;                           // rhoInf*UInf[i], rhoInf*EInf (sponge targets)
};

// What one side of a face needs (the record the tile kernels stage in shared memory).
//   aux: M1 -> sqrt(gamma*Rpsi) (speed of sound), M2 -> rhoE/rho + Rpsi (total enthalpy): the per-cell
//   square root / true division of the flux loops, evaluated once per staged cell.
//   sigmaU: U.tau of calc_VIS (uses rhoU/rho, a true division, unlike the flux loops' rhoU*rho_inv).
// tauMC is a cheap function of dudx and is rebuilt per face side instead of being stored.
template <class R, int D> struct CellState {
	R q[D + 2];
	R rho_inv, Rpsi, aux;
	R dudx[D][D], dTdx[D], sigmaU[D];
};

// Per-face geometry that does not change in time, derived once from (S, d) with the expressions of the
// viscous block (cfd_v0.cpp:2681-2690, 2747-2760): S_mag, 1/|d|, the non-orthogonal split K = S - delta and |delta|.
template <class R, int D> struct FaceGeo {
	R S[D], w, K[D], delta_mag, dmag_inv, S_mag;
};

template <class R, int D> __device__ __forceinline__ void primitives(const Consts<R>& k, const R* q, R& rho_inv, R* U, R& Rpsi, R& T) {
	rho_inv = R(1.0) / q[0];
#pragma unroll
	for (int i = 0; i < D; i++) U[i] = q[i + 1] * rho_inv;
	R rhoU_sqr = q[1] * q[1];
#pragma unroll
	for (int i = 1; i < D; i++) rhoU_sqr += q[i + 1] * q[i + 1];
	Rpsi = k.gm1 * (q[D + 1] - R(0.5) * rhoU_sqr * rho_inv) * rho_inv;
	T = Rpsi * k.Rgas_inv;
}

// stress tensor of calc_VIS / one_rk_step (same loop order as the reference: nD2 = nD1 % D)
template <class R, int D> __device__ __forceinline__ void stress(const Consts<R>& k, const R (*dudx)[D], R (*tau)[D]) {
#pragma unroll
	for (int nD = 0; nD < D; nD++) {
		tau[nD][nD] = R(2.0) * dudx[nD][nD];
#pragma unroll
		for (int nD1 = nD + 1; nD1 < D + nD; nD1++) {
			const int nD2 = nD1 % D;
			tau[nD][nD2] = k.mu * (dudx[nD][nD2] + dudx[nD2][nD]);
			tau[nD][nD] -= dudx[nD2][nD2];
		}
		tau[nD][nD] *= k.c_tau;
	}
}

template <class R, int D> __device__ __forceinline__ void tauMC_from(const Consts<R>& k, const R (*dudx)[D], R (*tauMC)[D]) {
	R diagSum = R(0);
#pragma unroll
	for (int nD = 0; nD < D; nD++) diagSum -= dudx[nD][nD];
	diagSum *= k.c_diag;
#pragma unroll
	for (int i = 0; i < D; i++) {
#pragma unroll
		for (int j = 0; j < D; j++) tauMC[i][j] = k.mu * dudx[j][i];
		tauMC[i][i] += diagSum;
	}
}

// tauMC of calc_VIS_Smagorinsky (cfd_v0.cpp:1660-1690): the laminar tauMC plus the eddy-viscosity contribution,
// smag = -2 (Cs Delta)^2 of the visiting cell (cfd_v0.cpp:1601-1602; computed on the host with the C library's pow)
template <class R, int D> __device__ __forceinline__ void tauMC_smagorinsky(const Consts<R>& k, R smag, const R (*dudx)[D], R (*tauMC)[D]) {
	tauMC_from<R, D>(k, dudx, tauMC);
	const R HALF = R(0.5);
	R Strain_Mag = R(0), divu = R(0);
#pragma unroll
	for (int i = 0; i < D; i++) {
#pragma unroll
		for (int j = 0; j < D; j++) {
			const R S_ij = HALF * (dudx[i][j] + dudx[j][i]);
			Strain_Mag += S_ij * S_ij;
		}
		divu += dudx[i][i];
	}
	Strain_Mag = LFM_SQRT(R(2.0) * Strain_Mag);
#pragma unroll
	for (int i = 0; i < D; i++) {
#pragma unroll
		for (int j = 0; j < D; j++) {
			const R S_ij = HALF * (dudx[i][j] + dudx[j][i]);
			tauMC[i][j] -= smag * Strain_Mag * S_ij;
		}
		tauMC[i][i] += smag * Strain_Mag * divu / R(3.0);
	}
}

// the per-cell block of calc_VIS (cfd_v0.cpp:1806-1857): sigmaU and tauMC from the finished dudx
template <class R, int D> __device__ __forceinline__ void vis_cell_terms(const Consts<R>& k, const R* q, const R (*dudx)[D], R (*tauMC)[D], R* sigmaU) {
	R Ud[D], tau[D][D];
#pragma unroll
	for (int i = 0; i < D; i++) Ud[i] = q[i + 1] / q[0];
	stress<R, D>(k, dudx, tau);
#pragma unroll
	for (int i = 0; i < D; i++) sigmaU[i] = dotD<R, D>(Ud, tau[i]);
	tauMC_from<R, D>(k, dudx, tauMC);
}

// Green-Gauss face values of calc_VIS (cfd_v0.cpp:1771-1790), owner perspective
template <class R, int D> __device__ __forceinline__ void grad_face_values(const Consts<R>& k, R w, const R* Uo, R Rpsi_o, const R* Un, R Rpsi_n, R* face_U, R& face_T) {
#pragma unroll
	for (int i = 0; i < D; i++) face_U[i] = interp<R>(w, Uo[i], Un[i]);
	face_T = interp<R>(w, Rpsi_o, Rpsi_n) * k.Rgas_inv;
}

template <class R, int D> __device__ __forceinline__ void make_geo(const R* S, const R* dv, R w, FaceGeo<R, D>& g) {
	const R ONE = R(1.0), ZERO = R(0.0);
	R S_mag = ZERO;
#pragma unroll
	for (int i = 0; i < D; i++) S_mag += S[i] * S[i];
	S_mag = LFM_SQRT(S_mag);
	R d_mag = dv[0] * dv[0];
#pragma unroll
	for (int i = 1; i < D; i++) d_mag += dv[i] * dv[i];
	d_mag = LFM_SQRT(d_mag);
	g.dmag_inv = ONE / d_mag;
	const R Sd = dotD<R, D>(S, dv);
	R delta_mag = ZERO;
#pragma unroll
	for (int i = 0; i < D; i++) {
		const R delta = dv[i] * S_mag * S_mag / Sd;
		delta_mag += delta * delta;
		g.K[i] = S[i] - delta;
		g.S[i] = S[i];
	}
	g.delta_mag = LFM_SQRT(delta_mag);
	g.S_mag = S_mag;
	g.w = w;
}

// Accessors: face_flux reads a side's state through one of these, so the same expression tree serves values held
// in registers (unfused kernels) and values staged in shared memory (tile kernels, loaded where they are used to
// keep the register footprint small).
template <class R, int D> struct RegSide {
	static constexpr bool kHasTrace = false;   // the side carries -(tr dudx) precomputed (trace_neg()) instead of forming it per face
	const CellState<R, D>& s;
	__device__ __forceinline__ R q(int k) const { return s.q[k]; }
	__device__ __forceinline__ R rho_inv() const { return s.rho_inv; }
	__device__ __forceinline__ R Rpsi() const { return s.Rpsi; }
	__device__ __forceinline__ R aux() const { return s.aux; }
	__device__ __forceinline__ R dudx(int i, int j) const { return s.dudx[i][j]; }
	__device__ __forceinline__ R dTdx(int i) const { return s.dTdx[i]; }
	__device__ __forceinline__ R sigmaU(int i) const { return s.sigmaU[i]; }
	__device__ __forceinline__ R tauMC(int, int) const { return R(0); }   // the unfused kernels serve the laminar closure only
};

// one-sided viscous terms of a physical-boundary face (cfd_v0.cpp:2701-2707, 2732-2745, 2774-2782); rare, kept out of line
template <class R, int D> __device__ __forceinline__ void ghost_face_viscous(const Consts<R>& k, const R* cU, const R* nU, R cT, R nT, const R* S, const R* K, R weight, R delta_mag, R dmag_inv, const R* dv, R* rhs) {
	R d_norm[D], dudx[D][D], dTdx[D], tauMC[D][D], sigmaU[D];
#pragma unroll
	for (int i = 0; i < D; i++) d_norm[i] = dv[i] * dmag_inv;
#pragma unroll
	for (int i = 0; i < D; i++) {
#pragma unroll
		for (int j = 0; j < D; j++) dudx[i][j] = (nU[i] - cU[i]) * d_norm[j] * dmag_inv;
		dTdx[i] = (nT - cT) * d_norm[i] * dmag_inv;
	}
	R tau[D][D], U_f[D];
	stress<R, D>(k, dudx, tau);
	tauMC_from<R, D>(k, dudx, tauMC);
#pragma unroll
	for (int i = 0; i < D; i++) U_f[i] = interp<R>(weight, cU[i], nU[i]);
#pragma unroll
	for (int i = 0; i < D; i++) sigmaU[i] = dotD<R, D>(U_f, tau[i]);
#pragma unroll
	for (int i = 0; i < D; i++) {
		const R divTauMC = dotD<R, D>(tauMC[i], S);
		const R lapU = k.mu * (delta_mag * (nU[i] - cU[i]) * dmag_inv + dotD<R, D>(K, dudx[i]));
		rhs[i + 1] += divTauMC + lapU;
	}
	const R lapT = k.kappa * (delta_mag * (nT - cT) * dmag_inv + dotD<R, D>(K, dTdx));
	const R divSigmaU = dotD<R, D>(sigmaU, S);
	rhs[D + 1] += divSigmaU + lapT;
}

// One face of one_rk_step_M1 / _M2: rhs[D+2] seen from the owner (c = owner, n = neighbour).
// ghost: the neighbour is a physical-boundary ghost (is_ghost) -> one-sided gradients; dv (the owner->ghost
// vector) is only read in that case.
// LES != 0: tauMC of both sides is read from the sides (c.tauMC(i, j), stored by the Smagorinsky gradient pass) instead of
// being rebuilt from dudx.
// SCHEME 2 (M2-AUSM): the M2 flux with `pavg_extra` (ausm_pressure_term) added to the face pressure (cfd_v0.cpp:2341).
template <class R, int D, int SCHEME, class SideC, class SideN, int LES = 0>
__device__ __forceinline__ void face_flux(const Consts<R>& k, const SideC& c, const SideN& n, const FaceGeo<R, D>& g, bool ghost, const R* dv, R* rhs, R pavg_extra = R(0)) {
	const R ONE = R(1.0), HALF = R(0.5), ZERO = R(0.0);
	const R weight = g.w;
	const R* S = g.S;
	R cU[D], nU[D];
	{
		const R cri = c.rho_inv(), nri = n.rho_inv();
#pragma unroll
		for (int i = 0; i < D; i++) {
			cU[i] = c.q(i + 1) * cri;
			nU[i] = n.q(i + 1) * nri;
		}
	}
	if (SCHEME == 0) {
		const R S_mag = g.S_mag;
		const R omw = ONE - weight;
		const R rhoPos = interp<R>(weight, c.q(0), n.q(0));
		const R rhoNeg = interp<R>(omw, n.q(0), c.q(0));
		const R rhoPos_inv = ONE / rhoPos;
		const R rhoNeg_inv = ONE / rhoNeg;
		R rhoUPos[D], rhoUNeg[D];
#pragma unroll
		for (int i = 0; i < D; i++) {
			rhoUPos[i] = interp<R>(weight, c.q(i + 1), n.q(i + 1));
			rhoUNeg[i] = interp<R>(omw, n.q(i + 1), c.q(i + 1));
		}
		R cell_e = R(2) * c.q(D + 1) * c.rho_inv();
		R adjc_e = R(2) * n.q(D + 1) * n.rho_inv();
#pragma unroll
		for (int nD = 0; nD < D; nD++) {
			cell_e -= cU[nD] * cU[nD];
			adjc_e -= nU[nD] * nU[nD];
		}
		cell_e *= HALF;
		adjc_e *= HALF;
		const R ePos = interp<R>(weight, cell_e, adjc_e);
		const R eNeg = interp<R>(omw, adjc_e, cell_e);
		const R RpsiPos = interp<R>(weight, c.Rpsi(), n.Rpsi());
		const R RpsiNeg = interp<R>(omw, n.Rpsi(), c.Rpsi());
		const R pPos = rhoPos * RpsiPos;
		const R pNeg = rhoNeg * RpsiNeg;
		const R cPos = interp<R>(weight, c.aux(), n.aux());
		const R cNeg = interp<R>(omw, n.aux(), c.aux());
		R phiPos = ZERO, phiNeg = ZERO;
		R uPos[D], uNeg[D];
#pragma unroll
		for (int i = 0; i < D; i++) {
			uPos[i] = rhoUPos[i] * rhoPos_inv;
			uNeg[i] = rhoUNeg[i] * rhoNeg_inv;
			phiPos += uPos[i] * S[i];
			phiNeg += -uNeg[i] * S[i];
		}
		R psiPos = phiPos + cPos * S_mag;
		{
			const R b = -phiNeg + cNeg * S_mag;
			if (psiPos < b) psiPos = b;
			if (psiPos < ZERO) psiPos = ZERO;
		}
		R psiNeg = phiPos - cPos * S_mag;
		{
			const R b = -phiNeg - cNeg * S_mag;
			if (b < psiNeg) psiNeg = b;
			if (ZERO < psiNeg) psiNeg = ZERO;
		}
		const R a0 = ONE / (psiPos - psiNeg);
		const R a1 = psiPos * psiNeg;
		const R aPos = psiPos * phiPos;
		const R aNeg = psiNeg * phiNeg;
		R rhoEPos = ePos, rhoENeg = eNeg;
#pragma unroll
		for (int nD = 0; nD < D; nD++) {
			rhoEPos += HALF * uPos[nD] * uPos[nD];
			rhoENeg += HALF * uNeg[nD] * uNeg[nD];
		}
		rhoEPos *= rhoPos;
		rhoENeg *= rhoNeg;
		rhs[0] = -(aPos * rhoPos + aNeg * rhoNeg + (rhoNeg - rhoPos) * a1) * a0;
#pragma unroll
		for (int i = 0; i < D; i++) {
			const R phiUp = (aPos * rhoUPos[i] + aNeg * rhoUNeg[i] + (rhoUNeg[i] - rhoUPos[i]) * a1) * a0 + (pPos * psiPos - pNeg * psiNeg) * a0 * S[i];
			rhs[i + 1] = -phiUp;
		}
		rhs[D + 1] = -(aPos * rhoEPos + aNeg * rhoENeg + (rhoENeg - rhoEPos) * a1 + (aPos * pPos + aNeg * pNeg)) * a0;
	} else {
		const R rhoavg = HALF * (c.q(0) + n.q(0));
		const R rhoavg_inv = ONE / rhoavg;
		const R Rpsiavg = HALF * (c.Rpsi() + n.Rpsi());
		R pavg = rhoavg * Rpsiavg;
		if (SCHEME == 2) pavg += pavg_extra;
		const R Havg = HALF * (c.aux() + n.aux());
		R rhoUavg[D];
		R phiavg = ZERO;
#pragma unroll
		for (int i = 0; i < D; i++) {
			rhoUavg[i] = HALF * (c.q(i + 1) + n.q(i + 1));
			const R uavg = rhoUavg[i] * rhoavg_inv;
			phiavg += uavg * S[i];
		}
		rhs[0] = -rhoavg * phiavg;
#pragma unroll
		for (int i = 0; i < D; i++) rhs[i + 1] = -(rhoUavg[i] * phiavg + pavg * S[i]);
		rhs[D + 1] = -(rhoavg * Havg * phiavg);
	}

	// ---- viscosity ----
	const R cT = c.Rpsi() * k.Rgas_inv, nT = n.Rpsi() * k.Rgas_inv;
	if (ghost) {
		ghost_face_viscous<R, D>(k, cU, nU, cT, nT, S, g.K, weight, g.delta_mag, g.dmag_inv, dv, rhs);
		return;
	}
	const R dmag_inv = g.dmag_inv;
	// diagonal part of tauMC of each side: (-tr dudx) * (mu*2/3)
	R cdiag = ZERO, ndiag = ZERO;
	if constexpr (SideC::kHasTrace && SideN::kHasTrace) {
		cdiag = c.trace_neg();   // ((0 - d00) - d11) - d22, formed once per cell by whoever wrote the V record
		ndiag = n.trace_neg();
	} else {
#pragma unroll
		for (int nD = 0; nD < D; nD++) {
			cdiag -= c.dudx(nD, nD);
			ndiag -= n.dudx(nD, nD);
		}
	}
	cdiag *= k.c_diag;
	ndiag *= k.c_diag;
	// divTauMC[i] = sum_j interp(w, tauMC_c[i][j], tauMC_n[i][j]) * S[j],  tauMC[i][j] = mu*dudx[j][i] (+ diag on i == j)
	// Kdudx[i]    = sum_j K[j] * interp(w, dudx_c[i][j], dudx_n[i][j])
	// One pass over the elements (a, b) of dudx in row-major order feeds both sums -- element (a, b) is term j = a of
	// divTauMC[b] and term j = b of Kdudx[a] -- so every element is read once and each sum still runs over ascending j.
	R divTauMC[D], Kdudx[D];
#pragma unroll
	for (int a = 0; a < D; a++) {
#pragma unroll
		for (int b = 0; b < D; b++) {
			const R cab = c.dudx(a, b), nab = n.dudx(a, b);
			R ct, nt;
			if (LES) {
				ct = c.tauMC(b, a);
				nt = n.tauMC(b, a);
			} else {
				ct = k.mu * cab;
				nt = k.mu * nab;
				if (a == b) {
					ct += cdiag;
					nt += ndiag;
				}
			}
			const R t = interp<R>(weight, ct, nt) * S[a];
			const R d = g.K[b] * interp<R>(weight, cab, nab);
			if (a == 0)
				divTauMC[b] = t;
			else
				divTauMC[b] += t;
			if (b == 0)
				Kdudx[a] = d;
			else
				Kdudx[a] += d;
		}
	}
#pragma unroll
	for (int i = 0; i < D; i++) {
		const R lapU = k.mu * (g.delta_mag * (nU[i] - cU[i]) * dmag_inv + Kdudx[i]);
		rhs[i + 1] += divTauMC[i] + lapU;
	}
	R KdTdx = ZERO, divSigmaU = ZERO;
#pragma unroll
	for (int j = 0; j < D; j++) {
		const R d = g.K[j] * interp<R>(weight, c.dTdx(j), n.dTdx(j));
		const R t = interp<R>(weight, c.sigmaU(j), n.sigmaU(j)) * S[j];
		if (j == 0) {
			KdTdx = d;
			divSigmaU = t;
		} else {
			KdTdx += d;
			divSigmaU += t;
		}
	}
	const R lapT = k.kappa * (g.delta_mag * (nT - cT) * dmag_inv + KdTdx);
	rhs[D + 1] += divSigmaU + lapT;
}

// ---- solver 2: AUSM+up pressure dissipation behind a shock sensor (cfd_v0.cpp:2270-2341) -------------------------
// calc_r / interp_minmod (api/cfdv0_solver.h:311-370), p5Pos / p5Neg (cfd_v0.cpp:1863-1894)
template <class R> __device__ __forceinline__ R lfm_abs(R x) { return x < R(0) ? -x : x; }
template <class R> __device__ __forceinline__ R sign_of(R a) { return a < R(0.0) ? R(-1.0) : R(1.0); }
template <class R, int D> __device__ __forceinline__ R calc_r(R phiP, R phiN, const R* phiGrad, const R* d) {
	const R gradf = phiN - phiP + R(1.0e-30);
	R gradcf = R(0);
#pragma unroll
	for (int i = 0; i < D; i++) gradcf += d[i] * phiGrad[i];
	if (lfm_abs(gradcf) >= R(1000.0) * lfm_abs(gradf)) return R(2.0) * R(1000.0) * sign_of(gradcf) * sign_of(gradf) - R(1.0);
	return R(2.0) * (gradcf / gradf) - R(1.0);
}
template <class R, int D> __device__ __forceinline__ R interp_minmod(R cell_phi, R adjc_phi, const R* grad_phi, const R* d, R weight_linear, R flux) {
	const R ONE = R(1.0), ZERO = R(0.0), HALF = R(0.5);
	const R r = calc_r<R, D>(cell_phi, adjc_phi, grad_phi, d);
	const R rm = r < ONE ? r : ONE;
	const R limiter = rm < ZERO ? ZERO : rm;
	const R weight = limiter * weight_linear + (ONE - limiter) * (ONE + flux) * HALF;
	return weight * cell_phi + (ONE - weight) * adjc_phi;
}
template <class R> __device__ __forceinline__ R p5Pos(R M, R alpha) {
	const R M2Pos = R(0.25) * (M + R(1.0)) * (M + R(1.0));
	const R M2Neg = R(-0.25) * (M - R(1.0)) * (M - R(1.0));
	const R M1Pos = R(0.5) * (M + lfm_abs(M));
	if (lfm_abs(M) < R(1)) return M2Pos * ((R(2.0) - M) - R(16.0) * alpha * M * M2Neg);
	return M1Pos / M;
}
template <class R> __device__ __forceinline__ R p5Neg(R M, R alpha) {
	const R M2Pos = R(0.25) * (M + R(1.0)) * (M + R(1.0));
	const R M2Neg = R(-0.25) * (M - R(1.0)) * (M - R(1.0));
	const R M1Neg = R(0.5) * (M - lfm_abs(M));
	if (lfm_abs(M) < R(1)) return M2Neg * ((R(-2.0) - M) + R(16.0) * alpha * M * M2Pos);
	return M1Neg / M;
}
// theta_avg * (pu - HALF * phalf): what one_rk_step_M2AUSM adds to the mid-point pressure of a face.
// cq/nq: conservatives, cd/nd: dudx, cUg/nUg: U_grad ([i][j] = dU_i/dx_j), crg/nrg: rho_grad, cpg/npg: p_grad of the two sides.
template <class R, int D>
__device__ __forceinline__ R ausm_pressure_term(const Consts<R>& k, const R* cq, const R* nq, R cell_Rpsi, R adjc_Rpsi, const R (*cd)[D], const R (*nd)[D], const R (*cUg)[D], const R (*nUg)[D],
                                                const R* crg, const R* nrg, const R* cpg, const R* npg, const R* S, const R* dv, R weight) {
	const R ONE = R(1.0), M_ONE = R(-1.0), HALF = R(0.5);
	const R cP = LFM_SQRT(k.gamma * cell_Rpsi);
	const R cN = LFM_SQRT(k.gamma * adjc_Rpsi);
	const R cavg = HALF * (cP + cN);
	R cell_divu = R(0), neigh_divu = R(0), cell_curlu = R(0), neigh_curlu = R(0), cell_vmag = R(0), neigh_vmag = R(0);
	R unPos = R(0), unNeg = R(0), norm = R(0);
#pragma unroll
	for (int nD = 0; nD < D; nD++) {
		cell_divu += cd[nD][nD];
		neigh_divu += nd[nD][nD];
#pragma unroll
		for (int nD2 = nD + 1; nD2 < D; nD2++) {
			cell_curlu += (cd[nD][nD2] - cd[nD2][nD]) * (cd[nD][nD2] - cd[nD2][nD]);
			neigh_curlu += (nd[nD][nD2] - nd[nD2][nD]) * (nd[nD][nD2] - nd[nD2][nD]);
		}
		const R cu = cq[nD + 1] / cq[0];
		cell_vmag += cu * cu;
		const R nu = nq[nD + 1] / nq[0];
		neigh_vmag += nu * nu;
		const R uP = interp_minmod<R, D>(cu, nu, cUg[nD], dv, weight, ONE);
		const R uN = interp_minmod<R, D>(cu, nu, nUg[nD], dv, weight, M_ONE);
		unPos += uP * S[nD];
		unNeg += uN * S[nD];
		norm += S[nD] * S[nD];
	}
	R th = -(cell_divu / LFM_SQRT(cell_divu * cell_divu + cell_curlu + R(4e-2)));
	const R cell_theta = th > R(0) ? th : R(0);
	th = -(neigh_divu / LFM_SQRT(neigh_divu * neigh_divu + neigh_curlu + R(4e-2)));
	const R neigh_theta = th > R(0) ? th : R(0);
	const R theta_avg = HALF * (cell_theta + neigh_theta);
	R r = cq[0];
	R E = cq[D + 1] / r;
	const R cell_p = r * k.gm1 * (E - HALF * cell_vmag);
	r = nq[0];
	E = nq[D + 1] / r;
	const R neigh_p = r * k.gm1 * (E - HALF * neigh_vmag);
	const R rhoP = interp_minmod<R, D>(cq[0], nq[0], crg, dv, weight, ONE);
	const R rhoN = interp_minmod<R, D>(cq[0], nq[0], nrg, dv, weight, M_ONE);
	const R pP = interp_minmod<R, D>(cell_p, neigh_p, cpg, dv, weight, ONE);
	const R pN = interp_minmod<R, D>(cell_p, neigh_p, npg, dv, weight, M_ONE);
	const R Msq = (unPos * unPos + unNeg * unNeg) / (R(2.) * cavg * cavg * norm);
	const R MPos = unPos / (LFM_SQRT(norm) * cavg);
	const R MNeg = unNeg / (LFM_SQRT(norm) * cavg);
	const R Minf = R(0.2);
	const R mx = Msq > Minf * Minf ? Msq : Minf * Minf;
	const R M0 = LFM_SQRT(R(1.0) < mx ? R(1.0) : mx);
	const R fa = M0 * (R(2.0) - M0);
	const R alpha = R(3.0) * (R(5.0) * fa * fa - R(4.0)) / R(16.0);
	const R phalf = pN * (p5Pos<R>(MNeg, alpha) - p5Neg<R>(MNeg, alpha)) - pP * (p5Pos<R>(MPos, alpha) - p5Neg<R>(MPos, alpha));
	const R pu = R(-0.75) * p5Pos<R>(MPos, alpha) * p5Neg<R>(MNeg, alpha) * (rhoP + rhoN) * cavg * fa * (unNeg - unPos);
	return theta_avg * (pu - HALF * phalf);
}

// Fills the derived members of a CellState whose q is already loaded (dudx, dTdx, sigmaU come from memory).
template <class R, int D, int SCHEME> __device__ __forceinline__ void derive_state(const Consts<R>& k, CellState<R, D>& s) {
	R U[D], T;
	primitives<R, D>(k, s.q, s.rho_inv, U, s.Rpsi, T);
	if (SCHEME == 0)
		s.aux = LFM_SQRT(k.gamma * s.Rpsi);
	else
		s.aux = s.q[D + 1] / s.q[0] + s.Rpsi;
}

}  // namespace lfm
