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

// What one side of a face needs.  `U` is rhoU*rho_inv (flux loops), `Ud` is rhoU/rho (calc_VIS block);
// the reference uses both forms and they differ in the last bit.
template <class R, int D> struct CellState {
	R q[D + 2];
	R rho_inv, Rpsi, T;
	R U[D];
	R dudx[D][D], dTdx[D], tauMC[D][D], sigmaU[D];
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

// One face of one_rk_step_M1 / _M2: rhs[D+2] seen from the owner (c = owner, n = neighbour).
// ghost: the neighbour is a physical-boundary ghost (is_ghost) -> one-sided gradients.
template <class R, int D, int SCHEME> __device__ __forceinline__ void face_flux(const Consts<R>& k, const CellState<R, D>& c, const CellState<R, D>& n, const R* S, const R* dv, R weight, bool ghost, R* rhs) {
	const R ONE = R(1.0), HALF = R(0.5), ZERO = R(0.0);
	R S_mag;
	if (SCHEME == 0) {
		const R omw = ONE - weight;
		const R rhoPos = interp<R>(weight, c.q[0], n.q[0]);
		const R rhoNeg = interp<R>(omw, n.q[0], c.q[0]);
		const R rhoPos_inv = ONE / rhoPos;
		const R rhoNeg_inv = ONE / rhoNeg;
		R rhoUPos[D], rhoUNeg[D];
#pragma unroll
		for (int i = 0; i < D; i++) {
			rhoUPos[i] = interp<R>(weight, c.q[i + 1], n.q[i + 1]);
			rhoUNeg[i] = interp<R>(omw, n.q[i + 1], c.q[i + 1]);
		}
		R cell_e = R(2) * c.q[D + 1] * c.rho_inv;
		R adjc_e = R(2) * n.q[D + 1] * n.rho_inv;
#pragma unroll
		for (int nD = 0; nD < D; nD++) {
			cell_e -= c.U[nD] * c.U[nD];
			adjc_e -= n.U[nD] * n.U[nD];
		}
		cell_e *= HALF;
		adjc_e *= HALF;
		const R ePos = interp<R>(weight, cell_e, adjc_e);
		const R eNeg = interp<R>(omw, adjc_e, cell_e);
		const R RpsiPos = interp<R>(weight, c.Rpsi, n.Rpsi);
		const R RpsiNeg = interp<R>(omw, n.Rpsi, c.Rpsi);
		const R pPos = rhoPos * RpsiPos;
		const R pNeg = rhoNeg * RpsiNeg;
		const R cP = LFM_SQRT(k.gamma * c.Rpsi);
		const R cN = LFM_SQRT(k.gamma * n.Rpsi);
		const R cPos = interp<R>(weight, cP, cN);
		const R cNeg = interp<R>(omw, cN, cP);
		R phiPos = ZERO, phiNeg = ZERO;
		R uPos[D], uNeg[D];
#pragma unroll
		for (int i = 0; i < D; i++) {
			uPos[i] = rhoUPos[i] * rhoPos_inv;
			uNeg[i] = rhoUNeg[i] * rhoNeg_inv;
			phiPos += uPos[i] * S[i];
			phiNeg += -uNeg[i] * S[i];
		}
		S_mag = ZERO;
#pragma unroll
		for (int i = 0; i < D; i++) S_mag += S[i] * S[i];
		S_mag = LFM_SQRT(S_mag);
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
		const R rhoavg = HALF * (c.q[0] + n.q[0]);
		const R rhoavg_inv = ONE / rhoavg;
		R rhoUavg[D];
#pragma unroll
		for (int i = 0; i < D; i++) rhoUavg[i] = HALF * (c.q[i + 1] + n.q[i + 1]);
		const R Rpsiavg = HALF * (c.Rpsi + n.Rpsi);
		const R pavg = rhoavg * Rpsiavg;
		const R cell_H = c.q[D + 1] / c.q[0] + c.Rpsi;
		const R adjc_H = n.q[D + 1] / n.q[0] + n.Rpsi;
		const R Havg = HALF * (cell_H + adjc_H);
		R phiavg = ZERO;
#pragma unroll
		for (int i = 0; i < D; i++) {
			const R uavg = rhoUavg[i] * rhoavg_inv;
			phiavg += uavg * S[i];
		}
		S_mag = ZERO;
#pragma unroll
		for (int i = 0; i < D; i++) S_mag += S[i] * S[i];
		S_mag = LFM_SQRT(S_mag);
		rhs[0] = -rhoavg * phiavg;
#pragma unroll
		for (int i = 0; i < D; i++) rhs[i + 1] = -(rhoUavg[i] * phiavg + pavg * S[i]);
		rhs[D + 1] = -(rhoavg * Havg * phiavg);
	}

	// ---- viscosity ----
	R d_mag = dv[0] * dv[0];
#pragma unroll
	for (int i = 1; i < D; i++) d_mag += dv[i] * dv[i];
	d_mag = LFM_SQRT(d_mag);
	const R dmag_inv = ONE / d_mag;
	R d_norm[D];
#pragma unroll
	for (int i = 0; i < D; i++) d_norm[i] = dv[i] * dmag_inv;
	R dudx[D][D], dTdx[D], tauMC[D][D], sigmaU[D];
	if (!ghost) {
#pragma unroll
		for (int i = 0; i < D; i++) {
#pragma unroll
			for (int j = 0; j < D; j++) {
				dudx[i][j] = interp<R>(weight, c.dudx[i][j], n.dudx[i][j]);
				tauMC[i][j] = interp<R>(weight, c.tauMC[i][j], n.tauMC[i][j]);
			}
			dTdx[i] = interp<R>(weight, c.dTdx[i], n.dTdx[i]);
			sigmaU[i] = interp<R>(weight, c.sigmaU[i], n.sigmaU[i]);
		}
	} else {
#pragma unroll
		for (int i = 0; i < D; i++) {
#pragma unroll
			for (int j = 0; j < D; j++) dudx[i][j] = (n.U[i] - c.U[i]) * d_norm[j] * dmag_inv;
			dTdx[i] = (n.T - c.T) * d_norm[i] * dmag_inv;
		}
		R tau[D][D], U_f[D];
		stress<R, D>(k, dudx, tau);
		tauMC_from<R, D>(k, dudx, tauMC);
#pragma unroll
		for (int i = 0; i < D; i++) U_f[i] = interp<R>(weight, c.U[i], n.U[i]);
#pragma unroll
		for (int i = 0; i < D; i++) sigmaU[i] = dotD<R, D>(U_f, tau[i]);
	}
	R divTauMC[D];
#pragma unroll
	for (int i = 0; i < D; i++) divTauMC[i] = dotD<R, D>(tauMC[i], S);
	const R Sd = dotD<R, D>(S, dv);
	R delta_mag = ZERO, K[D];
#pragma unroll
	for (int i = 0; i < D; i++) {
		const R delta = dv[i] * S_mag * S_mag / Sd;
		delta_mag += delta * delta;
		K[i] = S[i] - delta;
	}
	delta_mag = LFM_SQRT(delta_mag);
#pragma unroll
	for (int i = 0; i < D; i++) {
		const R lapU = k.mu * (delta_mag * (n.U[i] - c.U[i]) * dmag_inv + dotD<R, D>(K, dudx[i]));
		rhs[i + 1] += divTauMC[i] + lapU;
	}
	const R lapT = k.kappa * (delta_mag * (n.T - c.T) * dmag_inv + dotD<R, D>(K, dTdx));
	const R divSigmaU = dotD<R, D>(sigmaU, S);
	rhs[D + 1] += divSigmaU + lapT;
}

}  // namespace lfm
