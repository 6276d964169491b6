/* Body of the CPU restatement, included twice by lfm_oracle.c with REAL = double and REAL = float.
 * TEST INFRASTRUCTURE ONLY.  Every function follows the reference loop it cites literally: same loop
 * order, same in-place updates, same expression trees (including the double literals that promote
 * sub-expressions in the float build), so that with gcc (no FMA contraction) the double instantiation
 * reproduces oracle/_ref bit for bit on the same flattened mesh.
 *
 * Required macros: REAL, NAME(x), SQRT, FABS.
 */

#define INTERP_LINEAR(weight, a, b) ((weight) * (a) + (1.0 - (weight)) * (b))   /* api/cfdv0_solver.h:14 */

typedef struct NAME(ctx) {
	int D, NQ, F, n_cells, n_faces, n_bc, n_mpi, n_tot, n_sub, n_nbr;
	int sub_cell_start[LFMGPU_MAX_SUBMESH + 1];
	int *face_neigh, *cell_face_start, *cell_slot_face;
	int *bc_cell, *bc_kind, *bc_patch, *bc_face;
	int *send_start, *send_cell, *recv_start;
	REAL *S, *d, *w;                /* per face, stride 3 */
	REAL *vol_inv, *sigma;          /* vol_inv has n_tot entries (ghosts: 0) */
	REAL *q, *dq, *RES;             /* stride 5 */
	REAL *dudx, *tauMC;             /* stride 9 ([i*3+j]) */
	REAL *dTdx, *sigmaU;            /* stride 3 */
	REAL *rho_grad, *p_grad;        /* stride 3; solver 2 (M2-AUSM) only */
	REAL *U_grad;                   /* stride 9 ([i*3+j] = dU_i/dx_j) */
	unsigned char* is_ghost;
	REAL *pAVG, *pRMS;
	REAL gamma, gm1, Rgas_inv, mu, Cp, Pr_inv, rhoInf, UInf[3], EInf, pInf, TInf;
	REAL Ak[LFMGPU_MAX_RK], Bk[LFMGPU_MAX_RK];
	int comm_type;
	int les;                        /* 0: calc_VIS, 1: calc_VIS_Smagorinsky */
	int minmod;                     /* fvSchemes lfm/minmodExists: Mesh::solve calls calc_gradients[_M2AUSM] */
} NAME(ctx);

static void* NAME(xcalloc)(size_t n, size_t sz) {
	void* p = calloc(n ? n : 1, sz);
	if (!p) {
		fprintf(stderr, "lfm_oracle: out of memory\n");
		abort();
	}
	return p;
}

static NAME(ctx)* NAME(create)(const lfmgpu_desc* ds) {
	NAME(ctx)* c = (NAME(ctx)*)NAME(xcalloc)(1, sizeof *c);
	const int D = ds->dim;
	c->D = D;
	c->NQ = D + 2;
	c->F = ds->max_slots;
	c->n_cells = ds->n_cells;
	c->n_faces = ds->n_faces;
	c->n_bc = ds->n_bc_ghosts;
	c->n_mpi = ds->n_mpi_ghosts;
	c->n_tot = ds->n_cells + ds->n_bc_ghosts + ds->n_mpi_ghosts;
	c->n_sub = ds->n_sub;
	c->n_nbr = ds->n_nbr;
	memcpy(c->sub_cell_start, ds->sub_cell_start, sizeof c->sub_cell_start);
	c->face_neigh = (int*)NAME(xcalloc)((size_t)c->n_faces, sizeof(int));
	memcpy(c->face_neigh, ds->face_neigh, (size_t)c->n_faces * sizeof(int));
	c->cell_face_start = (int*)NAME(xcalloc)((size_t)c->n_cells + 1, sizeof(int));
	for (int f = 0; f < c->n_faces; f++) c->cell_face_start[ds->face_owner[f] + 1]++;
	for (int t = 0; t < c->n_cells; t++) c->cell_face_start[t + 1] += c->cell_face_start[t];
	c->cell_slot_face = (int*)NAME(xcalloc)((size_t)c->n_cells * c->F, sizeof(int));
	memcpy(c->cell_slot_face, ds->cell_slot_face, (size_t)c->n_cells * c->F * sizeof(int));
#define COPY_INT(dst, src, n) \
	c->dst = (int*)NAME(xcalloc)((size_t)(n), sizeof(int)); \
	if ((n) > 0) memcpy(c->dst, ds->src, (size_t)(n) * sizeof(int));
	COPY_INT(bc_cell, bc_cell, c->n_bc)
	COPY_INT(bc_kind, bc_kind, c->n_bc)
	COPY_INT(bc_patch, bc_patch, c->n_bc)
	COPY_INT(bc_face, bc_face, c->n_bc)
	COPY_INT(send_start, send_start, c->n_nbr + 1)
	COPY_INT(recv_start, recv_start, c->n_nbr + 1)
	COPY_INT(send_cell, send_cell, ds->n_nbr ? ds->send_start[ds->n_nbr] : 0)
#undef COPY_INT
	c->S = (REAL*)NAME(xcalloc)((size_t)c->n_faces * 3, sizeof(REAL));
	c->d = (REAL*)NAME(xcalloc)((size_t)c->n_faces * 3, sizeof(REAL));
	c->w = (REAL*)NAME(xcalloc)((size_t)c->n_faces, sizeof(REAL));
	for (int f = 0; f < c->n_faces; f++) {
		for (int k = 0; k < D; k++) {
			c->S[f * 3 + k] = ((const REAL*)ds->face_S)[f * D + k];
			c->d[f * 3 + k] = ((const REAL*)ds->face_d)[f * D + k];
		}
		c->w[f] = ((const REAL*)ds->face_w)[f];
	}
	c->vol_inv = (REAL*)NAME(xcalloc)((size_t)c->n_tot, sizeof(REAL));
	c->sigma = (REAL*)NAME(xcalloc)((size_t)c->n_cells, sizeof(REAL));
	c->q = (REAL*)NAME(xcalloc)((size_t)c->n_tot * 5, sizeof(REAL));
	c->dq = (REAL*)NAME(xcalloc)((size_t)c->n_tot * 5, sizeof(REAL));
	c->RES = (REAL*)NAME(xcalloc)((size_t)c->n_tot * 5, sizeof(REAL));
	c->dudx = (REAL*)NAME(xcalloc)((size_t)c->n_tot * 9, sizeof(REAL));
	c->tauMC = (REAL*)NAME(xcalloc)((size_t)c->n_tot * 9, sizeof(REAL));
	c->dTdx = (REAL*)NAME(xcalloc)((size_t)c->n_tot * 3, sizeof(REAL));
	c->sigmaU = (REAL*)NAME(xcalloc)((size_t)c->n_tot * 3, sizeof(REAL));
	c->rho_grad = (REAL*)NAME(xcalloc)((size_t)c->n_tot * 3, sizeof(REAL));
	c->p_grad = (REAL*)NAME(xcalloc)((size_t)c->n_tot * 3, sizeof(REAL));
	c->U_grad = (REAL*)NAME(xcalloc)((size_t)c->n_tot * 9, sizeof(REAL));
	c->is_ghost = (unsigned char*)NAME(xcalloc)((size_t)c->n_tot, 1);
	c->pAVG = (REAL*)NAME(xcalloc)((size_t)c->n_cells, sizeof(REAL));
	c->pRMS = (REAL*)NAME(xcalloc)((size_t)c->n_cells, sizeof(REAL));
	for (int t = 0; t < c->n_cells; t++) {
		c->vol_inv[t] = ((const REAL*)ds->vol_inv)[t];
		c->sigma[t] = ((const REAL*)ds->sponge_sigma)[t];
		for (int i = 0; i < c->NQ; i++) c->q[t * 5 + i] = ((const REAL*)ds->q0)[t * c->NQ + i];
	}
	const lfmgpu_consts* k = &ds->c;
	c->gamma = (REAL)k->gamma;
	c->gm1 = (REAL)k->gamma_m1;
	c->Rgas_inv = (REAL)k->Rgas_inv;
	c->mu = (REAL)k->mu;
	c->Cp = (REAL)k->Cp;
	c->Pr_inv = (REAL)k->Pr_inv;
	c->rhoInf = (REAL)k->rhoInf;
	for (int i = 0; i < 3; i++) c->UInf[i] = (REAL)k->UInf[i];
	c->EInf = (REAL)k->EInf;
	c->pInf = (REAL)k->pInf;
	c->TInf = (REAL)k->TInf;
	for (int i = 0; i < LFMGPU_MAX_RK; i++) {
		c->Ak[i] = (REAL)k->Ak[i];
		c->Bk[i] = (REAL)k->Bk[i];
	}
	c->comm_type = k->comm_type;
	return c;
}

static void NAME(destroy)(NAME(ctx)* c) {
	free(c->face_neigh); free(c->cell_face_start); free(c->cell_slot_face);
	free(c->bc_cell); free(c->bc_kind); free(c->bc_patch); free(c->bc_face);
	free(c->send_start); free(c->send_cell); free(c->recv_start);
	free(c->S); free(c->d); free(c->w); free(c->vol_inv); free(c->sigma);
	free(c->q); free(c->dq); free(c->RES); free(c->dudx); free(c->tauMC); free(c->dTdx); free(c->sigmaU);
	free(c->is_ghost); free(c->pAVG); free(c->pRMS);
	free(c);
}

/* api/cfdv0_solver.h:252-261 */
static REAL NAME(compute_Rpsi)(const NAME(ctx)* c, const REAL* q_sol) {
	const REAL ONE = 1.0;
	REAL rho_inv = ONE / q_sol[0];
	REAL rhoU_sqr = q_sol[1] * q_sol[1];
	for (int idim = 1; idim < c->D; idim++) rhoU_sqr += q_sol[idim + 1] * q_sol[idim + 1];
	return c->gm1 * (q_sol[c->D + 1] - 0.5 * rhoU_sqr * rho_inv) * rho_inv;
}
static REAL NAME(dot)(int D, const REAL* l, const REAL* r) {   /* cfdv0_solver.h:263-270 */
	REAL s = l[0] * r[0];
	for (int i = 1; i < D; i++) s += l[i] * r[i];
	return s;
}
static REAL NAME(vmag)(int D, const REAL* v) {                 /* cfdv0_solver.h:272-279 */
	REAL s = v[0] * v[0];
	for (int i = 1; i < D; i++) s += v[i] * v[i];
	return SQRT(s);
}
static REAL NAME(max3)(REAL a, REAL b, REAL c) {               /* std::max(initializer_list) */
	REAL m = a;
	if (m < b) m = b;
	if (m < c) m = c;
	return m;
}
static REAL NAME(min3)(REAL a, REAL b, REAL c) {
	REAL m = a;
	if (b < m) m = b;
	if (c < m) m = c;
	return m;
}

/* src/cfd_v0.cpp:1326-1333 */
static void NAME(prepare_timestep)(NAME(ctx)* c) {
	for (int t = 0; t < c->n_cells; t++)
		for (int i = 0; i < c->NQ; i++) c->dq[t * 5 + i] = c->RES[t * 5 + i] = 0.0;
}

/* src/cfd_v0.cpp:1339-1377 */
static void NAME(prepare_rkstep)(NAME(ctx)* c, int rk) {
	for (int t = 0; t < c->n_cells; t++) {
		for (int i = 0; i < c->NQ; i++) c->dq[t * 5 + i] *= c->Ak[rk];
		for (int i = 0; i < 9; i++) c->dudx[t * 9 + i] = c->tauMC[t * 9 + i] = 0.;
		for (int i = 0; i < 3; i++) c->dTdx[t * 3 + i] = c->rho_grad[t * 3 + i] = c->p_grad[t * 3 + i] = 0.;
		for (int i = 0; i < 9; i++) c->U_grad[t * 9 + i] = 0.;
	}
	for (int g = 0; g < c->n_bc; g++) c->is_ghost[c->n_cells + g] = 1;
}

/* src/cfd_v0.cpp:1010-1133 (wall / inlet / outlet; the far-field branch is unreachable) */
static void NAME(set_bc)(NAME(ctx)* c) {
	const int D = c->D;
	for (int g = 0; g < c->n_bc; g++) {
		const int gi = c->n_cells + g;
		const int b = c->bc_cell[g];
		REAL* gq = &c->q[gi * 5];
		const REAL* cq = &c->q[b * 5];
		if (c->bc_kind[g] == LFMGPU_BC_WALL) {
			gq[0] = cq[0];
			for (int i = 0; i < D; i++) gq[i + 1] = -cq[i + 1];
			gq[D + 1] = cq[D + 1];
			for (int i = 0; i < D; i++) {
				for (int j = 0; j < D; j++) {
					c->dudx[gi * 9 + i * 3 + j] = c->dudx[b * 9 + i * 3 + j];
					c->tauMC[gi * 9 + i * 3 + j] = c->tauMC[b * 9 + i * 3 + j];
				}
				c->dTdx[gi * 3 + i] = -c->dTdx[b * 3 + i];
			}
		} else if (c->bc_kind[g] == LFMGPU_BC_INLET) {
			REAL rhoInt = cq[0];
			REAL UvecInt[3];
			REAL Umag_sqrtInt = 0.0;
			for (int i = 0; i < D; i++) {
				UvecInt[i] = cq[i + 1] / rhoInt;
				Umag_sqrtInt += UvecInt[i] * UvecInt[i];
			}
			REAL pInt = (cq[D + 1] - 0.5 * Umag_sqrtInt * rhoInt) * c->gm1;
			REAL TInt = pInt * c->Rgas_inv / rhoInt;
			REAL TExt = c->TInf;
			REAL rhoExt = rhoInt * TInt / TExt;
			REAL UvecExt[3];
			REAL Umag_sqrtExt = 0.0;
			for (int i = 0; i < D; i++) {
				UvecExt[i] = c->UInf[i];
				Umag_sqrtExt += UvecExt[i] * UvecExt[i];
			}
			REAL pExt = pInt;
			REAL EExt = pExt / (rhoExt * c->gm1) + 0.5 * Umag_sqrtExt;
			gq[0] = rhoExt;
			for (int i = 0; i < D; i++) gq[i + 1] = rhoExt * UvecExt[i];
			gq[D + 1] = rhoExt * EExt;
			for (int i = 0; i < D; i++) {
				for (int j = 0; j < D; j++) {
					c->dudx[gi * 9 + i * 3 + j] = -c->dudx[b * 9 + i * 3 + j];
					c->tauMC[gi * 9 + i * 3 + j] = -c->tauMC[b * 9 + i * 3 + j];
				}
				c->dTdx[gi * 3 + i] = -c->dTdx[b * 3 + i];
			}
		} else if (c->bc_kind[g] == LFMGPU_BC_OUTLET) {
			REAL rhoInt = cq[0];
			REAL UvecInt[3];
			REAL Umag_sqrtInt = 0.0;
			for (int i = 0; i < D; i++) {
				UvecInt[i] = cq[i + 1] / rhoInt;
				Umag_sqrtInt += UvecInt[i] * UvecInt[i];
			}
			REAL pInt = (cq[D + 1] - 0.5 * Umag_sqrtInt * rhoInt) * c->gm1;
			REAL pExt = c->pInf;
			REAL rhoExt = rhoInt * pExt / pInt;
			REAL UvecExt[3];
			REAL Umag_sqrtExt = 0.0;
			for (int i = 0; i < D; i++) {
				UvecExt[i] = UvecInt[i];
				Umag_sqrtExt += UvecExt[i] * UvecExt[i];
			}
			REAL EExt = pExt / (rhoExt * c->gm1) + 0.5 * Umag_sqrtExt;
			gq[0] = rhoExt;
			for (int i = 0; i < D; i++) gq[i + 1] = rhoExt * UvecExt[i];
			gq[D + 1] = rhoExt * EExt;
			for (int i = 0; i < D; i++) {
				for (int j = 0; j < D; j++) {
					c->dudx[gi * 9 + i * 3 + j] = -c->dudx[b * 9 + i * 3 + j];
					c->tauMC[gi * 9 + i * 3 + j] = -c->tauMC[b * 9 + i * 3 + j];
				}
				c->dTdx[gi * 3 + i] = -c->dTdx[b * 3 + i];
			}
		}
	}
}

/* the tau / sigmaU / tauMC block of calc_VIS (src/cfd_v0.cpp:1806-1857) for one cell record; with les != 0 the block of
 * calc_VIS_Smagorinsky (src/cfd_v0.cpp:1649-1690): the same plus the eddy-viscosity contribution to tauMC, with the
 * smag_constant of the cell whose face loop is running (:1601-1602) */
static void NAME(vis_cell_terms)(NAME(ctx)* c, int x, REAL smag_constant) {
	const int D = c->D;
	REAL mu = c->mu, diagSum;
	REAL U[3], tau[3][3];
	const REAL* q = &c->q[x * 5];
	REAL* dudx = &c->dudx[x * 9];
	for (int i = 0; i < D; i++) U[i] = q[i + 1] / q[0];
	for (int nD = 0; nD < D; nD++) {
		tau[nD][nD] = 2.0 * dudx[nD * 3 + nD];
		for (int nD1 = nD + 1; nD1 < D + nD; nD1++) {
			const int nD2 = nD1 % D;
			tau[nD][nD2] = mu * (dudx[nD * 3 + nD2] + dudx[nD2 * 3 + nD]);
			tau[nD][nD] -= dudx[nD2 * 3 + nD2];
		}
		tau[nD][nD] *= 2.0 / 3.0 * mu;
	}
	diagSum = 0;
	for (int nD = 0; nD < D; nD++) diagSum -= dudx[nD * 3 + nD];
	diagSum *= mu * 2.0 / 3.0;
	for (int i = 0; i < D; i++) {
		c->sigmaU[x * 3 + i] = NAME(dot)(D, U, tau[i]);
		for (int j = 0; j < D; j++) c->tauMC[x * 9 + i * 3 + j] = mu * dudx[j * 3 + i];
		c->tauMC[x * 9 + i * 3 + i] += diagSum;
	}
	if (c->les) {
		const REAL HALF = 0.5;
		REAL Strain_Mag = 0.0, S_ij = 0.0, divu = 0.0;
		for (int i = 0; i < D; i++) {
			for (int j = 0; j < D; j++) {
				S_ij = HALF * (dudx[i * 3 + j] + dudx[j * 3 + i]);
				Strain_Mag += pow(S_ij, 2);
			}
			divu += dudx[i * 3 + i];
		}
		Strain_Mag = sqrt(2.0 * Strain_Mag);
		for (int i = 0; i < D; i++) {
			for (int j = 0; j < D; j++) {
				S_ij = HALF * (dudx[i * 3 + j] + dudx[j * 3 + i]);
				c->tauMC[x * 9 + i * 3 + j] -= smag_constant * Strain_Mag * S_ij;
			}
			c->tauMC[x * 9 + i * 3 + i] += smag_constant * Strain_Mag * divu / (REAL)3.0;
		}
	}
}

/* src/cfd_v0.cpp:1744-1860 calc_VIS over one submesh */
static void NAME(vis)(NAME(ctx)* c, int sub) {
	const int D = c->D;
	const REAL ONE = 1.0;
	REAL cell_Rpsi, adjc_Rpsi, cell_rho_inv, adjc_rho_inv, face_Rpsi, face_T;
	REAL face_U[3], cell_sov[3], adjc_sov[3];
	for (int t = c->sub_cell_start[sub]; t < c->sub_cell_start[sub + 1]; t++) {
		const REAL* vq = &c->q[t * 5];
		cell_Rpsi = NAME(compute_Rpsi)(c, vq);
		/* Smagorinsky, non dynamic: -2*(Cs*Delta)^2 (src/cfd_v0.cpp:1601-1602) */
		const REAL smag_constant = c->les ? -(REAL)2.0 * pow((REAL)0.16 * pow(1.0 / c->vol_inv[t], 1. / 3.), 2) : (REAL)0.0;
		for (int f = c->cell_face_start[t]; f < c->cell_face_start[t + 1]; f++) {
			const int n = c->face_neigh[f];
			const REAL* nq = &c->q[n * 5];
			const REAL* S = &c->S[f * 3];
			const REAL w = c->w[f];
			adjc_Rpsi = NAME(compute_Rpsi)(c, nq);
			cell_rho_inv = ONE / vq[0];
			adjc_rho_inv = ONE / nq[0];
			for (int i = 0; i < D; i++) face_U[i] = INTERP_LINEAR(w, vq[i + 1] * cell_rho_inv, nq[i + 1] * adjc_rho_inv);
			face_Rpsi = INTERP_LINEAR(w, cell_Rpsi, adjc_Rpsi);
			face_T = face_Rpsi * c->Rgas_inv;
			for (int i = 0; i < D; i++) {
				cell_sov[i] = S[i] * c->vol_inv[t];
				adjc_sov[i] = -S[i] * c->vol_inv[n];
			}
			for (int i = 0; i < D; i++)
				for (int j = 0; j < D; j++) {
					c->dudx[t * 9 + i * 3 + j] += face_U[i] * cell_sov[j];
					c->dudx[n * 9 + i * 3 + j] += face_U[i] * adjc_sov[j];
				}
			for (int i = 0; i < D; i++) {
				c->dTdx[t * 3 + i] += face_T * cell_sov[i];
				c->dTdx[n * 3 + i] += face_T * adjc_sov[i];
			}
			NAME(vis_cell_terms)(c, t, smag_constant);
			NAME(vis_cell_terms)(c, n, smag_constant);
		}
	}
}

/* src/cfd_v0.cpp:1384-1495 calc_gradients_M2AUSM over one submesh: Green-Gauss gradients of rho, p and U (the ones
 * one_rk_step_M2AUSM reads; rhoU_grad, rhoE_grad, Rpsi_grad and c_grad are computed by the reference and never used).
 * Ghost cells have vol_inv == 0 (never set by the reference), so nothing accumulates in them. */
static void NAME(gradients_ausm)(NAME(ctx)* c, int sub) {
	const int D = c->D;
	const REAL HALF = 0.5;
	REAL UU[3], cell_sov[3], adjc_sov[3];
	for (int t = c->sub_cell_start[sub]; t < c->sub_cell_start[sub + 1]; t++) {
		const REAL* vq = &c->q[t * 5];
		for (int f = c->cell_face_start[t]; f < c->cell_face_start[t + 1]; f++) {
			const int n = c->face_neigh[f];
			const REAL* nq = &c->q[n * 5];
			const REAL* S = &c->S[f * 3];
			const REAL w = c->w[f];
			for (int i = 0; i < D; i++) UU[i] = INTERP_LINEAR(w, vq[i + 1] / vq[0], nq[i + 1] / nq[0]);
			const REAL rho = INTERP_LINEAR(w, vq[0], nq[0]);
			for (int i = 0; i < D; i++) {
				cell_sov[i] = S[i] * c->vol_inv[t];
				adjc_sov[i] = -S[i] * c->vol_inv[n];
			}
			REAL r = vq[0];
			REAL E = vq[D + 1] / r;
			REAL vmag = 0.0;
			for (int i = 0; i < D; i++) vmag += (vq[i + 1] / r) * (vq[i + 1] / r);
			const REAL cell_p = r * c->gm1 * (E - HALF * vmag);
			r = nq[0];
			E = nq[D + 1] / r;
			vmag = 0.0;
			for (int i = 0; i < D; i++) vmag += (nq[i + 1] / r) * (nq[i + 1] / r);
			const REAL neigh_p = r * c->gm1 * (E - HALF * vmag);
			const REAL p = INTERP_LINEAR(w, cell_p, neigh_p);
			for (int i = 0; i < D; i++) {
				c->p_grad[t * 3 + i] += p * cell_sov[i];
				c->p_grad[n * 3 + i] += p * adjc_sov[i];
				for (int j = 0; j < D; j++) {
					c->U_grad[t * 9 + j * 3 + i] += UU[j] * cell_sov[i];
					c->U_grad[n * 9 + j * 3 + i] += UU[j] * adjc_sov[i];
				}
				c->rho_grad[t * 3 + i] += rho * cell_sov[i];
				c->rho_grad[n * 3 + i] += rho * adjc_sov[i];
			}
		}
	}
}

/* api/cfdv0_solver.h:311-370 calc_r / interp_minmod */
static REAL NAME(sign)(REAL a) { return a < 0.0 ? -1.0 : 1.0; }
static REAL NAME(calc_r)(int D, REAL phiP, REAL phiN, const REAL* phiGrad, const REAL* d) {
	const REAL ONE = 1.0, ZERO = 0.0;
	REAL gradf = phiN - phiP + 1.0e-30;
	REAL gradcf = ZERO;
	for (int i = 0; i < D; i++) gradcf += d[i] * phiGrad[i];
	if (FABS(gradcf) >= 1000.0 * FABS(gradf)) return 2.0 * 1000.0 * NAME(sign)(gradcf) * NAME(sign)(gradf) - ONE;
	return 2.0 * (gradcf / gradf) - ONE;
}
static REAL NAME(interp_minmod)(int D, REAL cell_phi, REAL adjc_phi, const REAL* grad_phi, const REAL* d, REAL weight_linear, REAL flux) {
	const REAL ONE = 1.0, ZERO = 0.0, HALF = 0.5;
	const REAL r = NAME(calc_r)(D, cell_phi, adjc_phi, grad_phi, d);
	const REAL rm = r < ONE ? r : ONE;               /* min(r, ONE) */
	const REAL limiter = rm < ZERO ? ZERO : rm;      /* max(.., ZERO) */
	const REAL weight = limiter * weight_linear + (ONE - limiter) * (ONE + flux) * HALF;
	return weight * cell_phi + (ONE - weight) * adjc_phi;
}
/* src/cfd_v0.cpp:1863-1894 */
static REAL NAME(p5Pos)(REAL M, REAL alpha) {
	REAL M2Pos = 0.25 * (M + 1.0) * (M + 1.0);
	REAL M2Neg = -0.25 * (M - 1.0) * (M - 1.0);
	REAL M1Pos = 0.5 * (M + FABS(M));
	REAL p5;
	if (FABS(M) < 1)
		p5 = M2Pos * ((2.0 - M) - 16.0 * alpha * M * M2Neg);
	else
		p5 = M1Pos / M;
	return p5;
}
static REAL NAME(p5Neg)(REAL M, REAL alpha) {
	REAL M2Pos = 0.25 * (M + 1.0) * (M + 1.0);
	REAL M2Neg = -0.25 * (M - 1.0) * (M - 1.0);
	REAL M1Neg = 0.5 * (M - FABS(M));
	REAL p5;
	if (FABS(M) < 1)
		p5 = M2Neg * ((-2.0 - M) + 16.0 * alpha * M * M2Pos);
	else
		p5 = M1Neg / M;
	return p5;
}

/* src/cfd_v0.cpp:2530-2832 one_rk_step_M1 and :1897-2179 one_rk_step_M2 over one submesh.
 * RES_out (D+2 accumulators) may be NULL. */
static void NAME(rk_stage)(NAME(ctx)* c, int sub, int scheme, int rk_step, REAL dt, REAL* RES_out) {
	const int D = c->D;
	const REAL ONE = 1.0, HALF = 0.5, ZERO = 0.0;
	REAL cell_rho_inv, cell_Rpsi, cell_T, adjc_rho_inv, adjc_Rpsi, adjc_T;
	REAL rhoPos, rhoNeg, rhoPos_inv, rhoNeg_inv, cell_e, adjc_e, ePos, eNeg, RpsiPos, RpsiNeg;
	REAL pPos, pNeg, cP, cN, cPos, cNeg, phiPos, phiNeg, S_mag, psiPos, psiNeg, a0, a1, aPos, aNeg, rhoEPos, rhoENeg, diagSum;
	REAL mu, d_mag, dmag_inv, Cp, k, delta_mag, laplacianT, divSigmaU, weight, oneMinusWeight;
	REAL cell_Uvec[3], adjc_Uvec[3], delta[3], K[3], rhoUPos[3], rhoUNeg[3], uPos[3], uNeg[3], phiUp[3], d_norm[3], dTdx[3];
	REAL laplacianU[3], sigmaU[3], U_f[3], rhs[5], dudx[3][3], tau[3][3], divTauMC[3], tauMC[3][3];
	REAL rhoavg, rhoavg_inv, rhoUavg[3], uavg[3], eavg, Rpsiavg, pavg, cavg, phiavg, rhoEavg, cell_H, adjc_H, Havg;
	const int bRkCheck = (RES_out && rk_step == 0) ? 1 : 0;
	(void)eavg; (void)cavg; (void)rhoEavg;

	for (int t = c->sub_cell_start[sub]; t < c->sub_cell_start[sub + 1]; t++) {
		REAL* cq = &c->q[t * 5];
		cell_rho_inv = ONE / cq[0];
		for (int i = 0; i < D; i++) cell_Uvec[i] = cq[i + 1] * cell_rho_inv;
		cell_Rpsi = NAME(compute_Rpsi)(c, cq);
		cell_T = cell_Rpsi * c->Rgas_inv;

		for (int f = c->cell_face_start[t]; f < c->cell_face_start[t + 1]; f++) {
			const int n = c->face_neigh[f];
			const REAL* nq = &c->q[n * 5];
			const REAL* S = &c->S[f * 3];
			const REAL* dv = &c->d[f * 3];
			weight = c->w[f];
			oneMinusWeight = ONE - weight;
			adjc_rho_inv = ONE / nq[0];
			for (int i = 0; i < D; i++) adjc_Uvec[i] = nq[i + 1] * adjc_rho_inv;
			adjc_Rpsi = NAME(compute_Rpsi)(c, nq);
			adjc_T = adjc_Rpsi * c->Rgas_inv;

			if (scheme == LFMGPU_SCHEME_M1) {
				/* ---- convection, M1 (cfd_v0.cpp:2586-2675) ---- */
				rhoPos = INTERP_LINEAR(weight, cq[0], nq[0]);
				rhoNeg = INTERP_LINEAR(oneMinusWeight, nq[0], cq[0]);
				rhoPos_inv = ONE / rhoPos;
				rhoNeg_inv = ONE / rhoNeg;
				for (int i = 0; i < D; i++) {
					rhoUPos[i] = INTERP_LINEAR(weight, cq[i + 1], nq[i + 1]);
					rhoUNeg[i] = INTERP_LINEAR(oneMinusWeight, nq[i + 1], cq[i + 1]);
				}
				cell_e = 2 * cq[D + 1] * cell_rho_inv;
				adjc_e = 2 * nq[D + 1] * adjc_rho_inv;
				for (int nD = 0; nD < D; nD++) {
					cell_e -= cell_Uvec[nD] * cell_Uvec[nD];
					adjc_e -= adjc_Uvec[nD] * adjc_Uvec[nD];
				}
				cell_e *= 0.5;
				adjc_e *= 0.5;
				ePos = INTERP_LINEAR(weight, cell_e, adjc_e);
				eNeg = INTERP_LINEAR(oneMinusWeight, adjc_e, cell_e);
				RpsiPos = INTERP_LINEAR(weight, cell_Rpsi, adjc_Rpsi);
				RpsiNeg = INTERP_LINEAR(oneMinusWeight, adjc_Rpsi, cell_Rpsi);
				pPos = rhoPos * RpsiPos;
				pNeg = rhoNeg * RpsiNeg;
				cP = SQRT(c->gamma * cell_Rpsi);
				cN = SQRT(c->gamma * adjc_Rpsi);
				cPos = INTERP_LINEAR(weight, cP, cN);
				cNeg = INTERP_LINEAR(oneMinusWeight, cN, cP);
				phiPos = 0.0;
				phiNeg = 0.0;
				for (int i = 0; i < D; i++) {
					uPos[i] = rhoUPos[i] * rhoPos_inv;
					uNeg[i] = rhoUNeg[i] * rhoNeg_inv;
					phiPos += uPos[i] * S[i];
					phiNeg += -uNeg[i] * S[i];
				}
				S_mag = 0.;
				for (int i = 0; i < D; i++) S_mag += pow(S[i], 2);
				S_mag = SQRT(S_mag);
				psiPos = NAME(max3)(phiPos + cPos * S_mag, -phiNeg + cNeg * S_mag, ZERO);
				psiNeg = NAME(min3)(phiPos - cPos * S_mag, -phiNeg - cNeg * S_mag, ZERO);
				a0 = ONE / (psiPos - psiNeg);
				a1 = psiPos * psiNeg;
				aPos = psiPos * phiPos;
				aNeg = psiNeg * phiNeg;
				rhoEPos = ePos;
				rhoENeg = eNeg;
				for (int nD = 0; nD < D; nD++) {
					rhoEPos += 0.5 * uPos[nD] * uPos[nD];
					rhoENeg += 0.5 * uNeg[nD] * uNeg[nD];
				}
				rhoEPos *= rhoPos;
				rhoENeg *= rhoNeg;
				rhs[0] = -(aPos * rhoPos + aNeg * rhoNeg + (rhoNeg - rhoPos) * a1) * a0;
				for (int i = 0; i < D; i++) {
					phiUp[i] = (aPos * rhoUPos[i] + aNeg * rhoUNeg[i] + (rhoUNeg[i] - rhoUPos[i]) * a1) * a0 + (pPos * psiPos - pNeg * psiNeg) * a0 * S[i];
					rhs[i + 1] = -phiUp[i];
				}
				rhs[D + 1] = -(aPos * rhoEPos + aNeg * rhoENeg + (rhoENeg - rhoEPos) * a1 + (aPos * pPos + aNeg * pNeg)) * a0;
			} else {
				/* ---- convection, M2 (cfd_v0.cpp:1963-2023) ---- */
				rhoavg = HALF * (cq[0] + nq[0]);
				rhoavg_inv = ONE / rhoavg;
				for (int i = 0; i < D; i++) rhoUavg[i] = HALF * (cq[i + 1] + nq[i + 1]);
				cell_e = cq[D + 1] * cell_rho_inv;
				adjc_e = nq[D + 1] * adjc_rho_inv;
				for (int nD = 0; nD < D; nD++) {
					cell_e -= 0.5 * cell_Uvec[nD] * cell_Uvec[nD];
					adjc_e -= 0.5 * adjc_Uvec[nD] * adjc_Uvec[nD];
				}
				eavg = HALF * (cell_e + adjc_e);
				Rpsiavg = HALF * (cell_Rpsi + adjc_Rpsi);
				pavg = rhoavg * Rpsiavg;
				cell_H = cq[D + 1] / cq[0] + cell_Rpsi;
				adjc_H = nq[D + 1] / nq[0] + adjc_Rpsi;
				Havg = HALF * (cell_H + adjc_H);
				if (scheme == 2) {
					/* ---- AUSM+up pressure dissipation behind a shock sensor (cfd_v0.cpp:2270-2341) ---- */
					const REAL M_ONE = -1.0;
					REAL cell_divu = 0., neigh_divu = 0., cell_curlu = 0., neigh_curlu = 0., cell_vmag = 0., neigh_vmag = 0.;
					REAL unPos = 0., unNeg = 0., norm = 0., cell_u[3], neigh_u[3];
					const REAL* cd = &c->dudx[t * 9];
					const REAL* nd = &c->dudx[n * 9];
					cP = SQRT(c->gamma * cell_Rpsi);
					cN = SQRT(c->gamma * adjc_Rpsi);
					cavg = HALF * (cP + cN);
					for (int nD = 0; nD < D; nD++) {
						cell_divu += cd[nD * 3 + nD];
						neigh_divu += nd[nD * 3 + nD];
						for (int nD2 = nD + 1; nD2 < D; nD2++) {
							cell_curlu += (cd[nD * 3 + nD2] - cd[nD2 * 3 + nD]) * (cd[nD * 3 + nD2] - cd[nD2 * 3 + nD]);
							neigh_curlu += (nd[nD * 3 + nD2] - nd[nD2 * 3 + nD]) * (nd[nD * 3 + nD2] - nd[nD2 * 3 + nD]);
						}
						cell_u[nD] = cq[nD + 1] / cq[0];
						cell_vmag += cell_u[nD] * cell_u[nD];
						neigh_u[nD] = nq[nD + 1] / nq[0];
						neigh_vmag += neigh_u[nD] * neigh_u[nD];
						const REAL uP = NAME(interp_minmod)(D, cell_u[nD], neigh_u[nD], &c->U_grad[t * 9 + nD * 3], dv, weight, ONE);
						const REAL uN = NAME(interp_minmod)(D, cell_u[nD], neigh_u[nD], &c->U_grad[n * 9 + nD * 3], dv, weight, M_ONE);
						unPos += uP * S[nD];
						unNeg += uN * S[nD];
						norm += S[nD] * S[nD];
					}
					REAL th = -(cell_divu / sqrt(cell_divu * cell_divu + cell_curlu + 4e-2));
					const REAL cell_theta = th > 0. ? th : 0.;
					th = -(neigh_divu / sqrt(neigh_divu * neigh_divu + neigh_curlu + 4e-2));
					const REAL neigh_theta = th > 0. ? th : 0.;
					const REAL theta_avg = HALF * (cell_theta + neigh_theta);
					REAL r = cq[0];
					REAL E = cq[D + 1] / r;
					const REAL cell_p = r * c->gm1 * (E - HALF * cell_vmag);
					r = nq[0];
					E = nq[D + 1] / r;
					const REAL neigh_p = r * c->gm1 * (E - HALF * neigh_vmag);
					const REAL rhoP = NAME(interp_minmod)(D, cq[0], nq[0], &c->rho_grad[t * 3], dv, weight, ONE);
					const REAL rhoN = NAME(interp_minmod)(D, cq[0], nq[0], &c->rho_grad[n * 3], dv, weight, M_ONE);
					const REAL pP = NAME(interp_minmod)(D, cell_p, neigh_p, &c->p_grad[t * 3], dv, weight, ONE);
					const REAL pN = NAME(interp_minmod)(D, cell_p, neigh_p, &c->p_grad[n * 3], dv, weight, M_ONE);
					const REAL Msq = (unPos * unPos + unNeg * unNeg) / (2. * cavg * cavg * norm);
					const REAL MPos = unPos / (sqrt(norm) * cavg);
					const REAL MNeg = unNeg / (sqrt(norm) * cavg);
					const REAL Minf = 0.2;
					REAL mx = Msq > Minf * Minf ? Msq : Minf * Minf;           /* max(Msq, Minf*Minf) */
					const REAL M0 = sqrt((REAL)1.0 < mx ? (REAL)1.0 : mx);     /* sqrt(min(1.0, ..)) */
					const REAL fa = M0 * (2.0 - M0);
					const REAL alpha = 3.0 * (5.0 * fa * fa - 4.0) / 16.0;
					const REAL phalf = pN * (NAME(p5Pos)(MNeg, alpha) - NAME(p5Neg)(MNeg, alpha)) - pP * (NAME(p5Pos)(MPos, alpha) - NAME(p5Neg)(MPos, alpha));
					const REAL pu = -0.75 * NAME(p5Pos)(MPos, alpha) * NAME(p5Neg)(MNeg, alpha) * (rhoP + rhoN) * cavg * fa * (unNeg - unPos);
					pavg += theta_avg * (pu - HALF * phalf);
				}
				phiavg = 0.0;
				for (int i = 0; i < D; i++) {
					uavg[i] = rhoUavg[i] * rhoavg_inv;
					phiavg += uavg[i] * S[i];
				}
				S_mag = 0.;
				for (int i = 0; i < D; i++) S_mag += pow(S[i], 2);
				S_mag = SQRT(S_mag);
				rhs[0] = -rhoavg * phiavg;
				for (int i = 0; i < D; i++) rhs[i + 1] = -(rhoUavg[i] * phiavg + pavg * S[i]);
				rhs[D + 1] = -(rhoavg * Havg * phiavg);
			}

			/* ---- viscosity (cfd_v0.cpp:2678-2790 == 2026-2137) ---- */
			mu = c->mu;
			d_mag = NAME(vmag)(D, dv);
			dmag_inv = ONE / d_mag;
			for (int i = 0; i < D; i++) d_norm[i] = dv[i] * dmag_inv;
			for (int i = 0; i < D; i++) {
				for (int j = 0; j < D; j++) dudx[i][j] = INTERP_LINEAR(weight, c->dudx[t * 9 + i * 3 + j], c->dudx[n * 9 + i * 3 + j]);
				dTdx[i] = INTERP_LINEAR(weight, c->dTdx[t * 3 + i], c->dTdx[n * 3 + i]);
			}
			if (c->is_ghost[n])
				for (int i = 0; i < D; i++) {
					for (int j = 0; j < D; j++) dudx[i][j] = (adjc_Uvec[i] - cell_Uvec[i]) * d_norm[j] * dmag_inv;
					dTdx[i] = (adjc_T - cell_T) * d_norm[i] * dmag_inv;
				}
			for (int nD = 0; nD < D; nD++) {
				tau[nD][nD] = 2.0 * dudx[nD][nD];
				for (int nD1 = nD + 1; nD1 < D + nD; nD1++) {
					const int nD2 = nD1 % D;
					tau[nD][nD2] = mu * (dudx[nD][nD2] + dudx[nD2][nD]);
					tau[nD][nD] -= dudx[nD2][nD2];
				}
				tau[nD][nD] *= 2.0 / 3.0 * mu;
			}
			Cp = c->Cp;
			k = Cp * mu * c->Pr_inv;
			for (int i = 0; i < D; i++)
				for (int j = 0; j < D; j++) tauMC[i][j] = INTERP_LINEAR(weight, c->tauMC[t * 9 + i * 3 + j], c->tauMC[n * 9 + i * 3 + j]);
			if (c->is_ghost[n]) {
				diagSum = 0;
				for (int nD = 0; nD < D; nD++) diagSum -= dudx[nD][nD];
				diagSum *= mu * 2.0 / 3.0;
				for (int i = 0; i < D; i++) {
					for (int j = 0; j < D; j++) tauMC[i][j] = mu * dudx[j][i];
					tauMC[i][i] += diagSum;
				}
			}
			for (int i = 0; i < D; i++) divTauMC[i] = NAME(dot)(D, tauMC[i], S);
			delta_mag = 0.0;
			for (int i = 0; i < D; i++) {
				delta[i] = dv[i] * S_mag * S_mag / NAME(dot)(D, S, dv);
				delta_mag += delta[i] * delta[i];
				K[i] = S[i] - delta[i];
			}
			delta_mag = SQRT(delta_mag);
			for (int i = 0; i < D; i++) laplacianU[i] = mu * (delta_mag * (adjc_Uvec[i] - cell_Uvec[i]) * dmag_inv + NAME(dot)(D, K, dudx[i]));
			laplacianT = k * (delta_mag * (adjc_T - cell_T) * dmag_inv + NAME(dot)(D, K, dTdx));
			for (int i = 0; i < D; i++) sigmaU[i] = INTERP_LINEAR(weight, c->sigmaU[t * 3 + i], c->sigmaU[n * 3 + i]);
			if (c->is_ghost[n]) {
				for (int i = 0; i < D; i++) U_f[i] = INTERP_LINEAR(weight, cell_Uvec[i], adjc_Uvec[i]);
				for (int i = 0; i < D; i++) sigmaU[i] = NAME(dot)(D, U_f, tau[i]);
			}
			divSigmaU = NAME(dot)(D, sigmaU, S);
			for (int i = 0; i < D; i++) rhs[i + 1] += divTauMC[i] + laplacianU[i];
			rhs[D + 1] += divSigmaU + laplacianT;

			if (bRkCheck)
				for (int i = 0; i < D + 2; i++) {
					c->RES[t * 5 + i] += rhs[i];
					c->RES[n * 5 + i] -= rhs[i];
				}
			for (int i = 0; i < D + 2; i++) {
				c->dq[t * 5 + i] += dt * rhs[i] * c->vol_inv[t];
				c->dq[n * 5 + i] -= dt * rhs[i] * c->vol_inv[n];
			}
		}
		/* sponge (cfd_v0.cpp:2810-2814) and RK update (2819-2822) */
		c->dq[t * 5 + 0] += dt * c->sigma[t] * (c->rhoInf - cq[0]);
		for (int i = 0; i < D; i++) c->dq[t * 5 + i + 1] += dt * c->sigma[t] * (c->rhoInf * c->UInf[i] - cq[i + 1]);
		c->dq[t * 5 + D + 1] += dt * c->sigma[t] * (c->rhoInf * c->EInf - cq[D + 1]);
		for (int i = 0; i < D + 2; i++) cq[i] += c->Bk[rk_step] * c->dq[t * 5 + i];
	}
	if (bRkCheck)
		for (int t = c->sub_cell_start[sub]; t < c->sub_cell_start[sub + 1]; t++)
			for (int i = 0; i < D + 2; i++) RES_out[i] += c->RES[t * 5 + i] * c->RES[t * 5 + i];
}

/* halo payload sizes (cfd_v0.cpp:686-709) */
static int NAME(scalars_per_cell)(const NAME(ctx)* c, int comm_step) {
	const int D = c->D;
	if (c->comm_type == LFMGPU_COMM_SPLIT) return comm_step == 0 ? D + 2 : 2 * D * D + 2 * D;
	return D + 2 + 2 * D * D + 2 * D;
}

/* pack: cfd_v0.cpp:3303-3331 (packed), 3398-3408 (solvars), 3475-3499 (viscous); all neighbours concatenated */
static void NAME(pack)(const NAME(ctx)* c, int comm_step, REAL* buf) {
	const int D = c->D;
	const int split = c->comm_type == LFMGPU_COMM_SPLIT;
	size_t o = 0;
	for (int n = 0; n < c->n_nbr; n++)
		for (int i = c->send_start[n]; i < c->send_start[n + 1]; i++) {
			const int g = c->send_cell[i];
			if (!split || comm_step == 0)
				for (int k = 0; k < D + 2; k++) buf[o++] = c->q[g * 5 + k];
			if (!split || comm_step == 1) {
				for (int a = 0; a < D; a++)
					for (int b = 0; b < D; b++) buf[o++] = c->dudx[g * 9 + a * 3 + b];
				for (int a = 0; a < D; a++) buf[o++] = c->dTdx[g * 3 + a];
				for (int a = 0; a < D; a++)
					for (int b = 0; b < D; b++) buf[o++] = c->tauMC[g * 9 + a * 3 + b];
				for (int a = 0; a < D; a++) buf[o++] = c->sigmaU[g * 3 + a];
			}
		}
}

/* unpack: cfd_v0.cpp:3608-3692 */
static void NAME(unpack)(NAME(ctx)* c, int comm_step, const REAL* buf) {
	const int D = c->D;
	const int split = c->comm_type == LFMGPU_COMM_SPLIT;
	size_t o = 0;
	for (int n = 0; n < c->n_nbr; n++)
		for (int i = c->recv_start[n]; i < c->recv_start[n + 1]; i++) {
			const int g = c->n_cells + c->n_bc + i;
			if (!split || comm_step == 0)
				for (int k = 0; k < D + 2; k++) c->q[g * 5 + k] = buf[o++];
			if (!split || comm_step == 1) {
				for (int a = 0; a < D; a++)
					for (int b = 0; b < D; b++) c->dudx[g * 9 + a * 3 + b] = buf[o++];
				for (int a = 0; a < D; a++) c->dTdx[g * 3 + a] = buf[o++];
				for (int a = 0; a < D; a++)
					for (int b = 0; b < D; b++) c->tauMC[g * 9 + a * 3 + b] = buf[o++];
				for (int a = 0; a < D; a++) c->sigmaU[g * 3 + a] = buf[o++];
			}
		}
}

/* S of slot k of cell t after reorder_faces, outward from t */
static void NAME(slot_S)(const NAME(ctx)* c, int t, int k, REAL* out, int* present) {
	const int e = c->cell_slot_face[t * c->F + k];
	*present = e != 0;
	if (!e) return;
	const int f = (e > 0 ? e : -e) - 1;
	for (int i = 0; i < c->D; i++) out[i] = e > 0 ? c->S[f * 3 + i] : -c->S[f * 3 + i];
}

/* src/cfd_v0.cpp:2887-2909 */
static REAL NAME(cfl)(const NAME(ctx)* c, REAL dt) {
	const REAL ONE = 1.0;
	REAL cfl, cfl_max = 0.;
	for (int t = 0; t < c->n_cells; t++) {
		cfl = 0;
		REAL rho_inv = ONE / c->q[t * 5];
		for (int k = 0; k < c->F; k++) {
			REAL S[3];
			int present;
			NAME(slot_S)(c, t, k, S, &present);
			if (!present) continue;
			REAL flux = 0.;
			for (int i = 0; i < c->D; i++) flux += c->q[t * 5 + i + 1] * S[i];
			cfl += FABS(flux);
		}
		cfl *= rho_inv * c->vol_inv[t] * dt;
		if (cfl_max < cfl) cfl_max = cfl;
	}
	return cfl_max;
}

/* src/cfd_v0.cpp:2913-2942 */
static REAL NAME(dtmin)(const NAME(ctx)* c, REAL cflMax) {
	REAL sumFlux, dt, dt_min = 100.0;
	for (int t = 0; t < c->n_cells; t++) {
		sumFlux = 0;
		REAL rho = c->q[t * 5];
		for (int k = 0; k < c->F; k++) {
			REAL S[3];
			int present;
			NAME(slot_S)(c, t, k, S, &present);
			if (!present) continue;
			REAL sMag = 0.;
			for (int i = 0; i < c->D; i++) sMag += S[i] * S[i];
			sMag = SQRT(sMag);
			REAL flux = 0.;
			for (int i = 0; i < c->D; i++) flux += c->q[t * 5 + i + 1] * S[i] / sMag;
			sumFlux += FABS(flux);
		}
		REAL vol = 1.0 / c->vol_inv[t];
		dt = cflMax * rho * 2.0 * vol / sumFlux;
		if (dt_min > dt) dt_min = dt;
	}
	return dt_min;
}

/* src/cfd_v0.cpp:3083-3103 */
static void NAME(average)(NAME(ctx)* c, int time_step) {
	if (time_step <= 0) return;
	for (int t = 0; t < c->n_cells; t++) {
		REAL r = c->q[t * 5];
		REAL velMag = 0;
		for (int nD = 1; nD <= c->D; nD++) {
			const REAL vel = c->q[t * 5 + nD] / r;
			velMag += (vel * vel);
		}
		REAL E = c->q[t * 5 + c->D + 1] / r;
		REAL p = (r * c->gm1) * (E - 0.5 * velMag);
		c->pAVG[t] += p;
		c->pRMS[t] += pow(p - c->pAVG[t] / time_step, 2);
	}
}

/* src/cfd_v0.cpp:3173-3240 local sums of one wall patch */
static void NAME(forces)(const NAME(ctx)* c, int patch, REAL* Fpre, REAL* Fvis) {
	const int D = c->D;
	const REAL ONE = 1.0;
	for (int i = 0; i < D; i++) Fpre[i] = Fvis[i] = 0.0;
	for (int g = 0; g < c->n_bc; g++) {
		if (c->bc_patch[g] != patch) continue;
		const int t = c->bc_cell[g], f = c->bc_face[g], n = c->n_cells + g;
		const REAL* S = &c->S[f * 3];
		const REAL* q = &c->q[t * 5];
		const REAL* nq = &c->q[n * 5];
		REAL r = q[0], rE = q[D + 1];
		REAL Uvec[3], Umag = 0.0, Smag = 0.0;
		for (int i = 0; i < D; i++) {
			Uvec[i] = q[i + 1] / r;
			Umag += Uvec[i] * Uvec[i];
			Smag += S[i] * S[i];
		}
		Umag = SQRT(Umag);
		Smag = SQRT(Smag);
		(void)Smag;
		REAL E = rE / r;
		REAL p = (r * c->gm1) * (E - 0.5 * Umag * Umag);
		for (int nD = 0; nD < D; nD++) Fpre[nD] += S[nD] * (p - c->pInf);
		REAL dudx[3][3], d_norm[3], tau[3][3];
		REAL d_mag = NAME(vmag)(D, &c->d[f * 3]);
		REAL dmag_inv = ONE / d_mag;
		for (int i = 0; i < D; i++) d_norm[i] = c->d[f * 3 + i] * dmag_inv;
		for (int i = 0; i < D; i++) {
			REAL cell_U = q[i + 1] / q[0];
			REAL adjc_U = nq[i + 1] / nq[0];
			for (int j = 0; j < D; j++) dudx[i][j] = (adjc_U - cell_U) * d_norm[j] * dmag_inv;
		}
		REAL mu = c->mu;
		for (int nD = 0; nD < D; nD++) {
			tau[nD][nD] = 2.0 * dudx[nD][nD];
			for (int nD1 = nD + 1; nD1 < D + nD; nD1++) {
				const int nD2 = nD1 % D;
				tau[nD][nD2] = mu * (dudx[nD][nD2] + dudx[nD2][nD]);
				tau[nD][nD] -= dudx[nD2][nD2];
			}
			tau[nD][nD] *= 2.0 / 3.0 * mu;
		}
		for (int a = 0; a < D; a++)
			for (int b = 0; b < D; b++) Fvis[a] -= tau[a][b] * S[b];
	}
}

#undef INTERP_LINEAR
