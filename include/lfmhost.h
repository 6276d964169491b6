/* lfmhost.h -- C ABI of the host-side case setup (no CUDA dependency): reads an LFM/OpenFOAM case or takes
 * mesh arrays, performs the reference's pre-loop setup (submesh split, geometry, ghosts, halo lists, initial
 * state; see lfm_public_b200/host/flatten.h for the reference functions restated) and hands out the
 * `lfmgpu_desc` that lfmgpu_create() uploads.
 *
 * Multi-rank setup needs one neighbour exchange (the reference does it with MPI_Isend/Irecv at init,
 * src/mesh_reader.cpp:490-614 and src/cfd_v0.cpp:646-683): lfmhost_export() -> transport of your choice
 * (torch.distributed, NCCL, in-process) -> lfmhost_import() on the neighbour, then lfmhost_finish().
 */
#ifndef LFMHOST_H
#define LFMHOST_H

#include <stddef.h>
#include <stdint.h>

#include "lfmgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lfmhost_case lfmhost_case;

/* plain-C mirror of CInputReader's members (reference: api/inputReader.h:34-88) + Foam::Time controls */
typedef struct lfmhost_opts {
	int32_t comm_type, halo_comm_type, double_precision;
	int32_t have_average, have_forces, have_residual, save_forces_step, print_info_freq;
	double t_start_average, cfl_max;
	int32_t solver, dimension, rk_order, minmod;
	double p_inf, T_inf, U_inf[3], Ls, mach, K;
	double Cp, mol_weight, mu0, Pr;
	int32_t laminar;
	double start_time, end_time, delta_t;
	int32_t write_interval, adjust_time_step, time_precision, write_precision;
} lfmhost_opts;

typedef struct lfmhost_mesh_in {
	int32_t n_points;
	const double* points;               /* [n_points][3]                                   */
	int32_t n_faces;
	const int32_t* faces;               /* [n_faces][4], -1 padded triangles                */
	const int32_t* owner;               /* [n_faces]                                        */
	int32_t n_internal;
	const int32_t* neighbour;           /* [n_internal]                                     */
	int32_t n_cells;
	int32_t n_patches;
	const char* const* patch_name;
	const char* const* patch_type;
	const int32_t* patch_nfaces;
	const int32_t* patch_start;
	const char* const* patch_nbr_name;  /* cyclic neighbourPatch ("" otherwise)             */
	const int32_t* patch_my_proc;       /* processor patches, -1 otherwise                  */
	const int32_t* patch_nbr_proc;
	const int32_t* face_proc_addressing;/* [n_faces] or NULL                                */
	const int32_t* cell_submesh;        /* [n_cells] or NULL                                */
	const char* const* patch_refer_name;/* processorCyclic referPatch ("" otherwise), or NULL */
} lfmhost_mesh_in;

const char* lfmhost_last_error(void);
void lfmhost_default_opts(lfmhost_opts* o);

/* rank < 0: serial case (mesh in <case>/constant/polyMesh); rank >= 0: <case>/processor<rank>/... with the
 * dictionaries of the global case (reference: info/lfm_solve.cpp:117-118, runTimeManagerOF.cpp:5-13) */
int lfmhost_open_case(const char* case_dir, int rank, int n_ranks, lfmhost_case** out);
/* U: [n_cells][3]; alpha may be NULL when opts->Ls <= 0 */
int lfmhost_case_from_arrays(const lfmhost_mesh_in* mesh, const lfmhost_opts* opts, const double* p, const double* T,
                             const double* U, const double* alpha, int rank, int n_ranks, lfmhost_case** out);
int lfmhost_get_opts(const lfmhost_case* c, lfmhost_opts* out);
int lfmhost_nbr_count(const lfmhost_case* c);
int lfmhost_nbr_rank(const lfmhost_case* c, int i);
int lfmhost_export(lfmhost_case* c, int i, const void** data, size_t* bytes);   /* valid until the next export(i) */
int lfmhost_import(lfmhost_case* c, int i, const void* data, size_t bytes);
int lfmhost_finish(lfmhost_case* c);
const lfmgpu_desc* lfmhost_desc(const lfmhost_case* c);
/* polyMesh geometry as computed for this case: any pointer may be NULL */
int lfmhost_geometry(const lfmhost_case* c, double* face_areas, double* face_centres, double* cell_centres, double* cell_volumes);
/* Hamiltonian-path cell numbering of the case's mesh (reference: hpathRenumber/hpathRenumber.C, the `hpath` method of
 * renumberMesh): order[new] = old, boundary submesh first.  stats[4]: boundary cells, boundary walk ok, interior walk ok,
 * fraction of the interior cells on the path itself. */
int lfmhost_hpath_order(const lfmhost_case* c, int32_t* order, double* stats);
int lfmhost_mesh_sizes(const lfmhost_case* c, int32_t* n_points, int32_t* n_faces, int32_t* n_internal, int32_t* n_cells);
/* Writes `values` ([n_cells] traversal order, nComp components interleaved) as an OpenFOAM vol field in polyMesh cell order */
int lfmhost_write_field(const lfmhost_case* c, const char* path, const char* name, const double* values, int n_comp, int precision);
int lfmhost_close(lfmhost_case* c);

#ifdef __cplusplus
}
#endif
#endif
