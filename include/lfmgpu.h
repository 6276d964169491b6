/* lfmgpu.h -- C ABI of the B200 (sm_100a) implementation of libFastMesh's per-iteration solve.
 *
 * This is the drop-in boundary for the hot path of TRC-HPC/LFM_Public: the per-stage virtuals of
 * `ISolver` (reference: api/cfdv0_solver.h:18-105) as called by `Mesh::solve`
 * (reference: src/mesh_solver.cpp:474-853).  The reference has no C ABI of its own; a
 * `CFDv0_solver_gpu<P,D,F> : CFDv0_solver<P,D,F>` selected in `Mesh::initializeSolver`
 * (src/mesh_solver.cpp:56-75) forwards each virtual to the entry point named below
 * (INTEGRATION.md shows that class).  Plain pointers and sizes only; every function returns 0 on
 * success and a non-zero code otherwise (text via lfmgpu_last_error()) -- the reference's own
 * convention is print + MPI_Abort(code), which the C++ wrapper restores.
 *
 * Threading: one host thread per handle (the reference is single-threaded per rank); calls are
 * asynchronous with respect to the device unless stated, ordered by the handle's streams exactly
 * in the call order of Mesh::solve.
 */
#ifndef LFMGPU_H
#define LFMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFMGPU_MAX_SUBMESH 8
#define LFMGPU_MAX_RK 8

/* scheme selector == fastmesh_solver_t (reference: api/fastmesh.h:36-39) */
enum { LFMGPU_SCHEME_M1 = 0, LFMGPU_SCHEME_M2 = 1, LFMGPU_SCHEME_M2AUSM = 2 };
/* physical boundary roles (reference: src/cfd_v0.cpp:973-1005 init_boundary_conditions) */
enum { LFMGPU_BC_NONE = 0, LFMGPU_BC_WALL = 1, LFMGPU_BC_INLET = 2, LFMGPU_BC_OUTLET = 3 };
/* payload modes == t_mpi_comm_type (reference: api/mpi_env.h:34-39); FULL_BND is served as PACKED */
enum { LFMGPU_COMM_FULL_BND = 0, LFMGPU_COMM_PACKED = 1, LFMGPU_COMM_SPLIT = 2 };
/* downloadable fields */
enum {
	LFMGPU_FIELD_Q = 0,      /* [n_cells][D+2]  conservatives, traversal order            */
	LFMGPU_FIELD_DQ = 1,     /* [n_cells][D+2]  low-storage RK accumulator (delta_q)      */
	LFMGPU_FIELD_DUDX = 2,   /* [n_cells][D*D]  row-major                                  */
	LFMGPU_FIELD_DTDX = 3,   /* [n_cells][D]                                               */
	LFMGPU_FIELD_RES = 4,    /* [n_cells][D+2]  stage-0 residual                           */
	LFMGPU_FIELD_PAVG = 5,   /* [n_cells]       running pressure sum                       */
	LFMGPU_FIELD_PRMS = 6,   /* [n_cells]       running squared deviation                  */
	LFMGPU_FIELD_QGHOST = 7, /* [n_bc_ghosts+n_mpi_ghosts][D+2]                            */
	LFMGPU_FIELD_TAUMC = 8,  /* [n_cells][D*D]  recomputed from dudx (what calc_VIS leaves)*/
	LFMGPU_FIELD_SIGMAU = 9  /* [n_cells][D]                                               */
};

/* Gas / scheme constants (reference: src/cfd_v0.cpp:880-902, 86-128).  Stored as double; for a
 * float build the host computes them in float and widens, so the values are exact. */
typedef struct lfmgpu_consts {
	double gamma, gamma_m1, Rgas_inv, mu, Cp, Pr_inv;
	double rhoInf, UInf[3], EInf, pInf, TInf;
	double Ak[LFMGPU_MAX_RK], Bk[LFMGPU_MAX_RK];
	int32_t rk_order;
	int32_t comm_type;
} lfmgpu_consts;

/* Flattened mesh + state of ONE rank, exactly what CFDv0_solver holds after reorder_faces()
 * (reference: src/cfd_v0.cpp:278-339), pointer-free.  Index space:
 *   [0, n_cells)                          real cells in reference traversal order: submesh 0
 *                                         (boundary) ascending, then submesh 1.. (interior)
 *   [n_cells, n_cells+n_bc_ghosts)        physical-boundary ghosts, by patch then by creation order
 *                                         (reference: ghost_bnd[patch][k], src/cfd_v0.cpp:382-431)
 *   [.., +n_mpi_ghosts)                   MPI ghosts, by neighbour then by receive order
 *                                         (reference: ghost_mpi[nbr][neigh_cells_to_recv[nbr][i]])
 * A "face" is one VALID (cell, slot) pair of the reference (slot < m_nFaceCount), numbered in the order
 * the reference's stage loops visit them (cell ascending, slot ascending); owner = that cell.
 * Real-typed arrays hold `precision`-byte floats. */
typedef struct lfmgpu_desc {
	int32_t precision;                       /* 4 | 8 == sizeof(PRECISION)                       */
	int32_t dim;                             /* DIM_CNT: 2 | 3                                   */
	int32_t max_slots;                       /* max FACE_CNT over the submeshes                   */
	int32_t n_sub;                           /* number of submeshes (1: boundary only)           */
	int32_t sub_cell_start[LFMGPU_MAX_SUBMESH + 1];
	int32_t sub_face_start[LFMGPU_MAX_SUBMESH + 1];
	int32_t sub_face_cnt[LFMGPU_MAX_SUBMESH];/* FACE_CNT of each submesh's solver                 */
	int32_t n_cells, n_faces, n_bc_ghosts, n_mpi_ghosts;
	/* faces [n_faces] */
	const int32_t* face_owner;
	const int32_t* face_neigh;
	const void* face_S;                      /* [n_faces][dim] owner-outward area vector          */
	const void* face_d;                      /* [n_faces][dim] owner centre -> neighbour centre   */
	const void* face_w;                      /* [n_faces] weight_linear                           */
	/* cells [n_cells] */
	const void* vol_inv;
	const void* sponge_sigma;
	const void* q0;                          /* [n_cells][dim+2] initial conservatives            */
	const int32_t* cell_gid;                 /* polyMesh cell index (rank-local) for output       */
	const int32_t* cell_slot_face;           /* [n_cells][max_slots]: +(face+1) owner side, -(face+1)
	                                            neighbour side, 0 = unused slot; slot order after
	                                            reorder_faces (valid first, invalid reversed)     */
	/* physical boundary ghosts [n_bc_ghosts] */
	const int32_t* bc_cell;                  /* inside cell (boundaries[p][k][0])                 */
	const int32_t* bc_kind;                  /* LFMGPU_BC_*                                       */
	const int32_t* bc_patch;                 /* polyMesh patch index                              */
	const int32_t* bc_face;                  /* face whose neighbour is this ghost                */
	/* halo */
	int32_t n_nbr;
	const int32_t* nbr_rank;                 /* [n_nbr]                                           */
	const int32_t* send_start;               /* [n_nbr+1] into send_cell                          */
	const int32_t* send_cell;                /* local_cells_to_send, boundary-submesh cell ids    */
	const int32_t* recv_start;               /* [n_nbr+1]; ghost = n_cells+n_bc_ghosts+recv_start[n]+i */
	lfmgpu_consts c;
} lfmgpu_desc;

typedef struct lfmgpu_ctx* lfmgpu_t;

const char* lfmgpu_last_error(void);
int lfmgpu_device_count(int* n);

/* ---- life cycle (CFDv0_solver::allocate .. reorder_faces happen on the host; this uploads) -------- */
int lfmgpu_create(const lfmgpu_desc* desc, int device, lfmgpu_t* out);
int lfmgpu_destroy(lfmgpu_t h);                              /* ISolver::deallocate                    */
int lfmgpu_sync(lfmgpu_t h);                                 /* wait for all streams of the handle     */

/* ---- per-iteration virtuals of ISolver, in Mesh::solve call order ---------------------------------- */
/* submesh: 0 = boundary submesh, 1.. = interior submeshes, -1 = all submeshes in one launch            */
int lfmgpu_prepare_timestep(lfmgpu_t h);                     /* cfd_v0.cpp:1326  dq = RES = 0           */
int lfmgpu_prepare_rkstep(lfmgpu_t h, int rk_step);          /* cfd_v0.cpp:1339  dq *= A_k (folded)     */
int lfmgpu_set_bc(lfmgpu_t h);                               /* cfd_v0.cpp:1010  set_boundary_conditions*/
int lfmgpu_gradients(lfmgpu_t h, int submesh);               /* cfd_v0.cpp:1501  calc_gradients (minmod; outputs unread by M1/M2: no work) */
int lfmgpu_gradients_m2ausm(lfmgpu_t h, int submesh);        /* cfd_v0.cpp:1384  calc_gradients_M2AUSM (rho, p, U gradients of solver 2) */
int lfmgpu_vis(lfmgpu_t h, int submesh);                     /* cfd_v0.cpp:1744  calc_VIS (laminar)     */
int lfmgpu_vis_smagorinsky(lfmgpu_t h, int submesh);         /* cfd_v0.cpp:1574  calc_VIS_Smagorinsky   */
int lfmgpu_rk_stage(lfmgpu_t h, int submesh, int scheme, int rk_step, double dt, int want_res);
                                                             /* cfd_v0.cpp:2530 / 1897 / 2186 one_rk_step_M1/_M2/_M2AUSM */
int lfmgpu_halo_start(lfmgpu_t h, int comm_step);            /* cfd_v0.cpp:3547  mpi_communication      */
int lfmgpu_halo_wait(lfmgpu_t h, int comm_step);             /* cfd_v0.cpp:3576  mpi_wait (+unpack)     */
int lfmgpu_cfl(lfmgpu_t h, double dt, double* cfl_max);      /* cfd_v0.cpp:2887  compute_cfl (blocking) */
int lfmgpu_dt(lfmgpu_t h, double cfl_max, double* dt_min);   /* cfd_v0.cpp:2913  compute_dt  (blocking) */
int lfmgpu_average(lfmgpu_t h, int time_step);               /* cfd_v0.cpp:3083  postProcAverage        */
int lfmgpu_forces(lfmgpu_t h, int patch, double* Fpre, double* Fvis); /* cfd_v0.cpp:3169 (local sums, blocking) */
int lfmgpu_residual(lfmgpu_t h, double* res);                /* D+2 sums of RES^2 (cfd_v0.cpp:2825-2831), blocking; zeroes them */

/* One or more whole time steps in the reference order (Mesh::solve, mesh_solver.cpp:487-691):
 * prepare_for_timestep; per stage: prepare_for_RKstep, mpi_wait(0), set_bc, [gradients], vis(bnd),
 * halo(1) overlapped with vis(int), rk_stage(bnd), halo(0) overlapped with rk_stage(int); final wait. */
int lfmgpu_step(lfmgpu_t h, int scheme, double dt, int n_steps, int minmod, int want_res);
/* The pre-loop warm-up of Mesh::solve (mesh_solver.cpp:409-428): halo(0), set_bc, wait(0), halo(1), wait(1). */
int lfmgpu_warmup(lfmgpu_t h);
/* Same as lfmgpu_step for ranks that are all handles of this process (in-process transport), advanced in
 * lockstep; first != 0 runs the warm-up before the first step. */
int lfmgpu_step_multi(const lfmgpu_t* hs, int n_ranks, int scheme, double dt, int n_steps, int first, int want_res);
/* The scalar reductions Mesh::solve does with MPI_Allreduce / MPI_Reduce (mesh_solver.cpp:715, 763-779,
 * cfd_v0.cpp:3242-3243): op 0 = sum, 1 = min, 2 = max over the NCCL ranks, in place, n <= 8; blocking. */
int lfmgpu_allreduce(lfmgpu_t h, double* values, int n, int op);
/* Options: "use_tiles" 1 (default) = fused shared-memory tile kernels (all three schemes), 0 = face kernel + gather kernels (tuning);
 * "laminar" 1 (default) / 0 = lfmgpu_step and lfmgpu_step_multi call calc_VIS / calc_VIS_Smagorinsky
 * (CInputReader::m_bLaminar, reference: src/inputReader.cpp:60, src/mesh_solver.cpp:556-560);
 * "minmod" 0 (default) / 1 = with LFMGPU_SCHEME_M2AUSM those loops call calc_gradients_M2AUSM before calc_VIS
 * (fvSchemes lfm/minmodExists, mesh_solver.cpp:537-548, 582-594; lfmgpu_step's minmod argument sets it too). */
int lfmgpu_set_option(lfmgpu_t h, const char* name, int value);

/* Host-only check of the tile plan lfmgpu_create would build for this rank (no device needed): partition, halo lists, face
 * tables, local gather lists and shared-memory strides are verified against the descriptor; tile_cells / smem_limit_bytes
 * <= 0 take the defaults (128 cells, 227 KB).  stats[10]: tileable, tiles, max staged cells, max faces per tile,
 * incoming/own faces, halo cells per cell, stage-kernel shared memory (bytes), cells per tile after halving, mean halo
 * cells per tile, mean runs of consecutive ids per halo list. */
int lfmgpu_plan_check(const lfmgpu_desc* desc, int tile_cells, int smem_limit_bytes, double* stats);

/* ---- data movement --------------------------------------------------------------------------------- */
int lfmgpu_download(lfmgpu_t h, int field, void* dst, size_t dst_bytes);   /* blocking                 */
int lfmgpu_upload_q(lfmgpu_t h, const void* q, size_t bytes);              /* [n_cells][D+2], blocking */
/* component-major host arrays [D+2][n_cells] (what a time-directory writer wants), asynchronous on the compute stream */
int lfmgpu_upload_q_soa_async(lfmgpu_t h, const void* q, size_t bytes);
int lfmgpu_download_q_soa_async(lfmgpu_t h, void* q, size_t bytes);
/* Pipelined host I/O for back-to-back batches (each batch: upload q, advance, download q): the upload of the next batch
 * and the download of the previous result run on their own copy streams through device staging buffers while the compute
 * stream advances the current batch.  Order per batch: pipe_in_start(host_in) [may be issued one batch ahead] ->
 * pipe_in_commit -> lfmgpu_step -> pipe_out_start -> pipe_out_fetch(host_out).  Host arrays are [D+2][n_cells] as in
 * lfmgpu_upload_q_soa_async and must stay valid (pinned for true overlap) until lfmgpu_sync / a later event. */
int lfmgpu_pipe_in_start(lfmgpu_t h, const void* q, size_t bytes);
int lfmgpu_pipe_in_commit(lfmgpu_t h);
int lfmgpu_pipe_out_start(lfmgpu_t h);
int lfmgpu_pipe_out_fetch(lfmgpu_t h, void* q, size_t bytes);
/* pinned host staging for the end-to-end path */
int lfmgpu_host_alloc(void** p, size_t bytes);
int lfmgpu_host_free(void* p);

/* ---- halo transport -------------------------------------------------------------------------------- */
/* NCCL (one rank per GPU).  unique_id is an opaque 128-byte ncclUniqueId from rank 0. */
int lfmgpu_nccl_unique_id(void* id128);
int lfmgpu_comm_init_nccl(lfmgpu_t h, const void* id128, int rank, int n_ranks);
/* In-process transport: all ranks are handles of this process (tests, 1-GPU multi-rank runs);
 * peers[r] is the handle of rank r. */
int lfmgpu_comm_init_local(lfmgpu_t h, int rank, int n_ranks, const lfmgpu_t* peers);
/* Packed send buffer of the last lfmgpu_halo_start(comm_step) (the reference's m_SendBuf*List contents,
 * cfd_v0.cpp:3303-3331 / 3398-3408 / 3475-3499), neighbours concatenated in nbr order. */
int lfmgpu_halo_send_count(lfmgpu_t h, int comm_step, size_t* n_scalars);
int lfmgpu_download_send_buffer(lfmgpu_t h, int comm_step, void* dst, size_t dst_bytes);

/* Host-staged halo for callers that keep their own transport (the reference's MPI_env, or more ranks than GPUs):
 * pack on the device into `dst` (same layout as lfmgpu_download_send_buffer, blocking) / upload a received buffer
 * (neighbours concatenated in nbr order, the reference's m_RecvBuf*List layout) and unpack it into the MPI ghosts. */
int lfmgpu_halo_pack_to_host(lfmgpu_t h, int comm_step, void* dst, size_t dst_bytes);
int lfmgpu_halo_unpack_from_host(lfmgpu_t h, int comm_step, const void* src, size_t bytes);

/* ---- introspection --------------------------------------------------------------------------------- */
int lfmgpu_launch_count(lfmgpu_t h, uint64_t* n);            /* kernels launched by this handle so far  */
/* average device time (ms) per launch and launch count of the kernels whose name starts with `prefix`, measured
 * with CUDA events on the launching stream when timing is enabled */
int lfmgpu_enable_kernel_timing(lfmgpu_t h, int on);
int lfmgpu_kernel_time(lfmgpu_t h, const char* prefix, double* total_ms, uint64_t* launches);
/* CUDA events on the handle's compute stream (after joining the halo stream): record into slot 0..7, read the
 * device time between two recorded slots (blocks until slot_b has happened) */
int lfmgpu_event_record(lfmgpu_t h, int slot);
int lfmgpu_event_elapsed_ms(lfmgpu_t h, int slot_a, int slot_b, double* ms);
/* shape of the fused-tile plan built at create time (n_tiles == 0: mesh served by the unfused kernels) */
int lfmgpu_tile_info(lfmgpu_t h, int* n_tiles, int* tile_cells, size_t* smem_bytes, double* halo_face_ratio);

#ifdef __cplusplus
}
#endif
#endif /* LFMGPU_H */
