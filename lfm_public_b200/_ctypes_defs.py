"""ctypes mirrors of include/lfmgpu.h and include/lfmhost.h (kept field-for-field in sync with the headers)."""
import ctypes as C

MAX_SUBMESH = 8
MAX_RK = 8

SCHEME_M1, SCHEME_M2, SCHEME_M2AUSM = 0, 1, 2
BC_NONE, BC_WALL, BC_INLET, BC_OUTLET = 0, 1, 2, 3
COMM_FULL_BND, COMM_PACKED, COMM_SPLIT = 0, 1, 2
(FIELD_Q, FIELD_DQ, FIELD_DUDX, FIELD_DTDX, FIELD_RES, FIELD_PAVG, FIELD_PRMS, FIELD_QGHOST, FIELD_TAUMC,
 FIELD_SIGMAU) = range(10)


class Consts(C.Structure):
    _fields_ = [
        ("gamma", C.c_double), ("gamma_m1", C.c_double), ("Rgas_inv", C.c_double), ("mu", C.c_double),
        ("Cp", C.c_double), ("Pr_inv", C.c_double),
        ("rhoInf", C.c_double), ("UInf", C.c_double * 3), ("EInf", C.c_double), ("pInf", C.c_double), ("TInf", C.c_double),
        ("Ak", C.c_double * MAX_RK), ("Bk", C.c_double * MAX_RK),
        ("rk_order", C.c_int32), ("comm_type", C.c_int32),
    ]


class Desc(C.Structure):
    _fields_ = [
        ("precision", C.c_int32), ("dim", C.c_int32), ("max_slots", C.c_int32), ("n_sub", C.c_int32),
        ("sub_cell_start", C.c_int32 * (MAX_SUBMESH + 1)),
        ("sub_face_start", C.c_int32 * (MAX_SUBMESH + 1)),
        ("sub_face_cnt", C.c_int32 * MAX_SUBMESH),
        ("n_cells", C.c_int32), ("n_faces", C.c_int32), ("n_bc_ghosts", C.c_int32), ("n_mpi_ghosts", C.c_int32),
        ("face_owner", C.POINTER(C.c_int32)), ("face_neigh", C.POINTER(C.c_int32)),
        ("face_S", C.c_void_p), ("face_d", C.c_void_p), ("face_w", C.c_void_p),
        ("vol_inv", C.c_void_p), ("sponge_sigma", C.c_void_p), ("q0", C.c_void_p),
        ("cell_gid", C.POINTER(C.c_int32)), ("cell_slot_face", C.POINTER(C.c_int32)),
        ("bc_cell", C.POINTER(C.c_int32)), ("bc_kind", C.POINTER(C.c_int32)), ("bc_patch", C.POINTER(C.c_int32)),
        ("bc_face", C.POINTER(C.c_int32)),
        ("n_nbr", C.c_int32),
        ("nbr_rank", C.POINTER(C.c_int32)), ("send_start", C.POINTER(C.c_int32)), ("send_cell", C.POINTER(C.c_int32)),
        ("recv_start", C.POINTER(C.c_int32)),
        ("c", Consts),
    ]


class HostOpts(C.Structure):
    _fields_ = [
        ("comm_type", C.c_int32), ("halo_comm_type", C.c_int32), ("double_precision", C.c_int32),
        ("have_average", C.c_int32), ("have_forces", C.c_int32), ("have_residual", C.c_int32),
        ("save_forces_step", C.c_int32), ("print_info_freq", C.c_int32),
        ("t_start_average", C.c_double), ("cfl_max", C.c_double),
        ("solver", C.c_int32), ("dimension", C.c_int32), ("rk_order", C.c_int32), ("minmod", C.c_int32),
        ("p_inf", C.c_double), ("T_inf", C.c_double), ("U_inf", C.c_double * 3), ("Ls", C.c_double), ("mach", C.c_double),
        ("K", C.c_double),
        ("Cp", C.c_double), ("mol_weight", C.c_double), ("mu0", C.c_double), ("Pr", C.c_double),
        ("laminar", C.c_int32),
        ("start_time", C.c_double), ("end_time", C.c_double), ("delta_t", C.c_double),
        ("write_interval", C.c_int32), ("adjust_time_step", C.c_int32), ("time_precision", C.c_int32),
        ("write_precision", C.c_int32),
    ]


class MeshIn(C.Structure):
    _fields_ = [
        ("n_points", C.c_int32), ("points", C.POINTER(C.c_double)),
        ("n_faces", C.c_int32), ("faces", C.POINTER(C.c_int32)), ("owner", C.POINTER(C.c_int32)),
        ("n_internal", C.c_int32), ("neighbour", C.POINTER(C.c_int32)),
        ("n_cells", C.c_int32), ("n_patches", C.c_int32),
        ("patch_name", C.POINTER(C.c_char_p)), ("patch_type", C.POINTER(C.c_char_p)),
        ("patch_nfaces", C.POINTER(C.c_int32)), ("patch_start", C.POINTER(C.c_int32)),
        ("patch_nbr_name", C.POINTER(C.c_char_p)),
        ("patch_my_proc", C.POINTER(C.c_int32)), ("patch_nbr_proc", C.POINTER(C.c_int32)),
        ("face_proc_addressing", C.POINTER(C.c_int32)), ("cell_submesh", C.POINTER(C.c_int32)),
        ("patch_refer_name", C.POINTER(C.c_char_p)),
    ]
