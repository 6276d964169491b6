"""Python view of liblfmgpu.so (include/lfmgpu.h): the host-side mirror of the reference's per-iteration
`ISolver` interface (reference: api/cfdv0_solver.h:18-105) as `Mesh::solve` calls it
(src/mesh_solver.cpp:474-853).  Method names follow the reference's virtuals.

There is no CPU fallback: if the CUDA library is missing or no device is present the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._ctypes_defs import (FIELD_DQ, FIELD_DTDX, FIELD_DUDX, FIELD_PAVG, FIELD_PRMS, FIELD_Q, FIELD_QGHOST, FIELD_RES,
                           FIELD_SIGMAU, FIELD_TAUMC)

_LIB = None
# LFMGPU_LIB: alternative build of the same library (tuning experiments: e.g. an --fmad=true build)
LIB_PATH = os.environ.get("LFMGPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblfmgpu.so")

# every symbol include/lfmgpu.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "lfmgpu_last_error", "lfmgpu_device_count", "lfmgpu_create", "lfmgpu_destroy", "lfmgpu_sync",
    "lfmgpu_prepare_timestep", "lfmgpu_prepare_rkstep", "lfmgpu_set_bc", "lfmgpu_gradients", "lfmgpu_gradients_m2ausm", "lfmgpu_vis", "lfmgpu_vis_smagorinsky",
    "lfmgpu_rk_stage", "lfmgpu_halo_start", "lfmgpu_halo_wait", "lfmgpu_cfl", "lfmgpu_dt", "lfmgpu_average",
    "lfmgpu_forces", "lfmgpu_residual", "lfmgpu_step", "lfmgpu_warmup", "lfmgpu_step_multi", "lfmgpu_allreduce",
    "lfmgpu_set_option", "lfmgpu_plan_check", "lfmgpu_download", "lfmgpu_upload_q", "lfmgpu_upload_q_soa_async", "lfmgpu_download_q_soa_async",
    "lfmgpu_pipe_in_start", "lfmgpu_pipe_in_commit", "lfmgpu_pipe_out_start", "lfmgpu_pipe_out_fetch",
    "lfmgpu_host_alloc", "lfmgpu_host_free", "lfmgpu_nccl_unique_id", "lfmgpu_comm_init_nccl", "lfmgpu_comm_init_local",
    "lfmgpu_halo_send_count", "lfmgpu_download_send_buffer", "lfmgpu_halo_pack_to_host", "lfmgpu_halo_unpack_from_host", "lfmgpu_launch_count", "lfmgpu_enable_kernel_timing",
    "lfmgpu_kernel_time", "lfmgpu_tile_info", "lfmgpu_event_record", "lfmgpu_event_elapsed_ms",
]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make gpu` (or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.lfmgpu_last_error.restype = C.c_char_p
        vp, i, d, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
        sig = {
            "lfmgpu_device_count": [C.POINTER(i)], "lfmgpu_create": [vp, i, C.POINTER(vp)], "lfmgpu_destroy": [vp],
            "lfmgpu_sync": [vp], "lfmgpu_prepare_timestep": [vp], "lfmgpu_prepare_rkstep": [vp, i], "lfmgpu_set_bc": [vp],
            "lfmgpu_gradients": [vp, i], "lfmgpu_gradients_m2ausm": [vp, i], "lfmgpu_vis": [vp, i], "lfmgpu_vis_smagorinsky": [vp, i], "lfmgpu_rk_stage": [vp, i, i, i, d, i],
            "lfmgpu_halo_start": [vp, i], "lfmgpu_halo_wait": [vp, i], "lfmgpu_cfl": [vp, d, C.POINTER(d)],
            "lfmgpu_dt": [vp, d, C.POINTER(d)], "lfmgpu_average": [vp, i], "lfmgpu_forces": [vp, i, vp, vp],
            "lfmgpu_residual": [vp, vp], "lfmgpu_step": [vp, i, d, i, i, i], "lfmgpu_warmup": [vp],
            "lfmgpu_step_multi": [vp, i, i, d, i, i, i], "lfmgpu_allreduce": [vp, vp, i, i],
            "lfmgpu_set_option": [vp, C.c_char_p, i], "lfmgpu_plan_check": [vp, i, i, vp], "lfmgpu_download": [vp, i, vp, sz], "lfmgpu_upload_q": [vp, vp, sz],
            "lfmgpu_upload_q_soa_async": [vp, vp, sz], "lfmgpu_download_q_soa_async": [vp, vp, sz],
            "lfmgpu_pipe_in_start": [vp, vp, sz], "lfmgpu_pipe_in_commit": [vp], "lfmgpu_pipe_out_start": [vp],
            "lfmgpu_pipe_out_fetch": [vp, vp, sz],
            "lfmgpu_host_alloc": [C.POINTER(vp), sz], "lfmgpu_host_free": [vp], "lfmgpu_nccl_unique_id": [vp],
            "lfmgpu_comm_init_nccl": [vp, vp, i, i], "lfmgpu_comm_init_local": [vp, i, i, vp],
            "lfmgpu_halo_pack_to_host": [vp, i, vp, sz], "lfmgpu_halo_unpack_from_host": [vp, i, vp, sz],
            "lfmgpu_halo_send_count": [vp, i, C.POINTER(sz)], "lfmgpu_download_send_buffer": [vp, i, vp, sz],
            "lfmgpu_launch_count": [vp, C.POINTER(C.c_uint64)], "lfmgpu_enable_kernel_timing": [vp, i],
            "lfmgpu_kernel_time": [vp, C.c_char_p, C.POINTER(d), C.POINTER(C.c_uint64)],
            "lfmgpu_event_record": [vp, i], "lfmgpu_event_elapsed_ms": [vp, i, i, C.POINTER(d)],
            "lfmgpu_tile_info": [vp, C.POINTER(i), C.POINTER(i), C.POINTER(sz), C.POINTER(d)],
        }
        for name, args in sig.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = C.c_int
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise RuntimeError("lfmgpu: " + lib().lfmgpu_last_error().decode())


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().lfmgpu_device_count(C.byref(n))
    return n.value if rc == 0 else 0


class PinnedArray:
    """numpy array over cudaMallocHost memory (lfmgpu_host_alloc)."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _check(lib().lfmgpu_host_alloc(C.byref(p), max(self.nbytes, 16)))
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().lfmgpu_host_free(C.c_void_p(self.ptr))
            self.ptr = None


def plan_check(case, tile_cells=0, smem_limit_bytes=0):
    """Host-only verification of the tile plan of a finished host_api.Case (no GPU needed); returns the plan statistics."""
    st = np.zeros(10)
    _check(lib().lfmgpu_plan_check(C.cast(case.desc_ptr, C.c_void_p), int(tile_cells), int(smem_limit_bytes), st.ctypes.data))
    return dict(tileable=bool(st[0]), n_tiles=int(st[1]), max_staged=int(st[2]), max_faces=int(st[3]), halo_face_ratio=float(st[4]),
                halo_cell_ratio=float(st[5]), smem_bytes=int(st[6]), tile_cells=int(st[7]), halo_per_tile=float(st[8]), halo_runs_per_tile=float(st[9]))


class GpuSolver:
    """One rank's device-resident solver: the GPU counterpart of CFDv0_solver<P,D,F> for the boundary + interior
    submeshes of the rank (the reference keeps one ISolver per submesh; submesh is an argument here)."""

    def __init__(self, case, device=0):
        self.case = case                      # keeps the descriptor storage alive
        d = case.desc
        self.D, self.NQ, self.n_cells, self.n_sub = d.dim, d.dim + 2, d.n_cells, d.n_sub
        self.n_ghost = d.n_bc_ghosts + d.n_mpi_ghosts
        self.real = np.float64 if d.precision == 8 else np.float32
        self.device = device
        h = C.c_void_p()
        _check(lib().lfmgpu_create(C.cast(case.desc_ptr, C.c_void_p), device, C.byref(h)))
        self.h = h
        try:
            laminar = bool(case.opts.laminar)
        except Exception:
            laminar = True
        if not laminar:                       # turbulenceProperties simulationType != laminar: calc_VIS_Smagorinsky in the time loop
            self.set_option("laminar", 0)
        try:
            self.minmod = int(case.opts.minmod)
        except Exception:
            self.minmod = 0
        if self.minmod:                       # fvSchemes lfm/minmodExists: calc_gradients_M2AUSM in the time loop of solver 2
            self.set_option("minmod", 1)

    # ---- ISolver virtuals -------------------------------------------------------------------------------
    def prepare_for_timestep(self):
        _check(lib().lfmgpu_prepare_timestep(self.h))

    def prepare_for_RKstep(self, rk):
        _check(lib().lfmgpu_prepare_rkstep(self.h, rk))

    def set_boundary_conditions(self):
        _check(lib().lfmgpu_set_bc(self.h))

    def calc_gradients(self, submesh=-1):
        _check(lib().lfmgpu_gradients(self.h, submesh))

    def calc_gradients_M2AUSM(self, submesh=-1):
        _check(lib().lfmgpu_gradients_m2ausm(self.h, submesh))

    def calc_VIS(self, submesh=-1):
        _check(lib().lfmgpu_vis(self.h, submesh))

    def calc_VIS_Smagorinsky(self, submesh=-1):
        _check(lib().lfmgpu_vis_smagorinsky(self.h, submesh))

    def one_rk_step(self, submesh, scheme, rk, dt, want_res=False):
        _check(lib().lfmgpu_rk_stage(self.h, submesh, scheme, rk, float(dt), int(want_res)))

    def mpi_communication(self, step):
        _check(lib().lfmgpu_halo_start(self.h, step))

    def mpi_wait(self, step):
        _check(lib().lfmgpu_halo_wait(self.h, step))

    def compute_cfl(self, dt):
        v = C.c_double()
        _check(lib().lfmgpu_cfl(self.h, float(dt), C.byref(v)))
        return v.value

    def compute_dt(self, cfl_max):
        v = C.c_double()
        _check(lib().lfmgpu_dt(self.h, float(cfl_max), C.byref(v)))
        return v.value

    def postProcAverage(self, time_step):
        _check(lib().lfmgpu_average(self.h, time_step))

    def postProcForces(self, patch):
        a = np.zeros(3)
        b = np.zeros(3)
        _check(lib().lfmgpu_forces(self.h, patch, a.ctypes.data, b.ctypes.data))
        return a[:self.D], b[:self.D]

    def residual(self):
        r = np.zeros(8)
        _check(lib().lfmgpu_residual(self.h, r.ctypes.data))
        return r[:self.NQ]

    # ---- whole steps --------------------------------------------------------------------------------------
    def warmup(self):
        _check(lib().lfmgpu_warmup(self.h))

    def step(self, scheme, dt, n_steps=1, want_res=False):
        _check(lib().lfmgpu_step(self.h, scheme, float(dt), n_steps, self.minmod, int(want_res)))

    def sync(self):
        _check(lib().lfmgpu_sync(self.h))

    def set_option(self, name, value):
        _check(lib().lfmgpu_set_option(self.h, name.encode(), int(value)))

    # ---- data ---------------------------------------------------------------------------------------------
    def download(self, field):
        D, NQ, n = self.D, self.NQ, self.n_cells
        shape = {FIELD_Q: (n, NQ), FIELD_DQ: (n, NQ), FIELD_RES: (n, NQ), FIELD_DUDX: (n, D, D), FIELD_TAUMC: (n, D, D),
                 FIELD_DTDX: (n, D), FIELD_SIGMAU: (n, D), FIELD_PAVG: (n,), FIELD_PRMS: (n,),
                 FIELD_QGHOST: (self.n_ghost, NQ)}[field]
        out = np.zeros(shape, dtype=self.real)
        _check(lib().lfmgpu_download(self.h, field, out.ctypes.data, out.nbytes))
        return out

    def upload_q(self, q):
        q = np.ascontiguousarray(q, dtype=self.real)
        _check(lib().lfmgpu_upload_q(self.h, q.ctypes.data, q.nbytes))

    def upload_q_soa_async(self, ptr, nbytes):
        _check(lib().lfmgpu_upload_q_soa_async(self.h, ptr, nbytes))

    def download_q_soa_async(self, ptr, nbytes):
        _check(lib().lfmgpu_download_q_soa_async(self.h, ptr, nbytes))

    def pipe_in_start(self, ptr, nbytes):
        _check(lib().lfmgpu_pipe_in_start(self.h, ptr, nbytes))

    def pipe_in_commit(self):
        _check(lib().lfmgpu_pipe_in_commit(self.h))

    def pipe_out_start(self):
        _check(lib().lfmgpu_pipe_out_start(self.h))

    def pipe_out_fetch(self, ptr, nbytes):
        _check(lib().lfmgpu_pipe_out_fetch(self.h, ptr, nbytes))

    def send_buffer(self, step):
        n = C.c_size_t()
        _check(lib().lfmgpu_halo_send_count(self.h, step, C.byref(n)))
        out = np.zeros(n.value, dtype=self.real)
        _check(lib().lfmgpu_download_send_buffer(self.h, step, out.ctypes.data, out.nbytes))
        return out

    # ---- transport ----------------------------------------------------------------------------------------
    def comm_init_nccl(self, unique_id: bytes, rank, n_ranks):
        assert len(unique_id) == 128
        _check(lib().lfmgpu_comm_init_nccl(self.h, unique_id, rank, n_ranks))

    def allreduce(self, values, op):
        v = np.ascontiguousarray(values, dtype=np.float64)
        _check(lib().lfmgpu_allreduce(self.h, v.ctypes.data, len(v), {"sum": 0, "min": 1, "max": 2}[op]))
        return v

    # ---- introspection ------------------------------------------------------------------------------------
    @property
    def launch_count(self):
        n = C.c_uint64()
        _check(lib().lfmgpu_launch_count(self.h, C.byref(n)))
        return n.value

    def enable_kernel_timing(self, on=True):
        _check(lib().lfmgpu_enable_kernel_timing(self.h, int(on)))

    def kernel_time(self, prefix):
        ms = C.c_double()
        n = C.c_uint64()
        _check(lib().lfmgpu_kernel_time(self.h, prefix.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def event_record(self, slot):
        _check(lib().lfmgpu_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_double()
        _check(lib().lfmgpu_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def tile_info(self):
        nt, tc, sm, hr = C.c_int(), C.c_int(), C.c_size_t(), C.c_double()
        _check(lib().lfmgpu_tile_info(self.h, C.byref(nt), C.byref(tc), C.byref(sm), C.byref(hr)))
        return dict(n_tiles=nt.value, tile_cells=tc.value, smem_bytes=sm.value, halo_face_ratio=hr.value)

    def close(self):
        if self.h:
            lib().lfmgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().lfmgpu_nccl_unique_id(buf))
    return buf.raw


def init_local(solvers):
    """In-process transport: solvers[r] is rank r (tests, several ranks on one GPU, or one process driving several GPUs)."""
    n = len(solvers)
    arr = (C.c_void_p * n)(*[s.h for s in solvers])
    for r, s in enumerate(solvers):
        _check(lib().lfmgpu_comm_init_local(s.h, r, n, arr))


def step_multi(solvers, scheme, dt, n_steps=1, first=False, want_res=False):
    n = len(solvers)
    arr = (C.c_void_p * n)(*[s.h for s in solvers])
    _check(lib().lfmgpu_step_multi(arr, n, scheme, float(dt), n_steps, int(first), int(want_res)))
