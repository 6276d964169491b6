"""ctypes binding of oracle/liblfm_oracle.so (the CPU restatement).  Lives under tests/ because only the
tests, smoke() and bench.py's cpu_baseline leg may touch oracle/."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liblfm_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
        L = C.CDLL(path)
        L.lfmo_create.restype = C.c_void_p
        L.lfmo_create.argtypes = [C.c_void_p]
        L.lfmo_destroy.argtypes = [C.c_void_p]
        for f in ("lfmo_prepare_timestep", "lfmo_set_bc"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.lfmo_prepare_rkstep.argtypes = [C.c_void_p, C.c_int]
        L.lfmo_vis.argtypes = [C.c_void_p, C.c_int]
        L.lfmo_set_les.argtypes = [C.c_void_p, C.c_int]
        L.lfmo_set_minmod.argtypes = [C.c_void_p, C.c_int]
        L.lfmo_gradients_ausm.argtypes = [C.c_void_p, C.c_int]
        L.lfmo_rk_stage.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.lfmo_halo_count.restype = C.c_size_t
        L.lfmo_halo_count.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.lfmo_pack.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.lfmo_unpack.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.lfmo_cfl.restype = C.c_double
        L.lfmo_cfl.argtypes = [C.c_void_p, C.c_double]
        L.lfmo_dt.restype = C.c_double
        L.lfmo_dt.argtypes = [C.c_void_p, C.c_double]
        L.lfmo_average.argtypes = [C.c_void_p, C.c_int]
        L.lfmo_forces.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.lfmo_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.lfmo_upload_q.argtypes = [C.c_void_p, C.c_void_p]
        L.lfmo_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


FIELD_SHAPES = {0: "NQ", 1: "NQ", 2: "DD", 3: "D", 4: "NQ", 5: "1", 6: "1", 7: "GNQ", 8: "DD", 9: "D"}


class Oracle:
    """CPU restatement of one rank (see oracle/lfm_oracle.c)."""

    def __init__(self, case):
        self.case = case                       # keeps the descriptor arrays alive
        d = case.desc
        self.D, self.NQ, self.n_cells = d.dim, d.dim + 2, d.n_cells
        self.n_ghost = d.n_bc_ghosts + d.n_mpi_ghosts
        self.real = np.float64 if d.precision == 8 else np.float32
        self.h = lib().lfmo_create(C.cast(case.desc_ptr, C.c_void_p))
        if not case.opts.laminar:          # turbulenceProperties simulationType != laminar: calc_VIS_Smagorinsky
            lib().lfmo_set_les(self.h, 1)
        if case.opts.minmod:
            lib().lfmo_set_minmod(self.h, 1)

    def download(self, field):
        kind = FIELD_SHAPES[field]
        D, NQ, n = self.D, self.NQ, self.n_cells
        shape = {"NQ": (n, NQ), "DD": (n, D, D), "D": (n, D), "1": (n,), "GNQ": (self.n_ghost, NQ)}[kind]
        out = np.zeros(shape, dtype=self.real)
        lib().lfmo_download(self.h, field, out.ctypes.data)
        return out

    def pack(self, step):
        n = lib().lfmo_halo_count(self.h, step, 1)
        buf = np.zeros(n, dtype=self.real)
        lib().lfmo_pack(self.h, step, buf.ctypes.data)
        return buf

    def __getattr__(self, name):
        f = getattr(lib(), "lfmo_" + name)
        return lambda *a: f(self.h, *a)

    def close(self):
        if self.h:
            lib().lfmo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run(oracles, scheme, dt, n_steps, first=True, want_res=False):
    """Mesh::solve on all ranks in lockstep; returns residual sums [n_steps, NQ] when want_res."""
    n = len(oracles)
    arr = (C.c_void_p * n)(*[o.h for o in oracles])
    res = np.zeros((n_steps, oracles[0].NQ)) if want_res else None
    rc = lib().lfmo_run(arr, n, scheme, dt, n_steps, 1 if first else 0, res.ctypes.data if want_res else None)
    assert rc == 0
    return res


def scalars_per_cell(D, comm_type, step):
    if comm_type == 2:
        return D + 2 if step == 0 else 2 * D * D + 2 * D
    return D + 2 + 2 * D * D + 2 * D


def lockstep_run(oracles, cases, scheme, dt, n_steps, record=None):
    """Python restatement of the Mesh::solve call order (reference: src/mesh_solver.cpp:409-428, 474-691) driving the
    per-function oracle API, with the neighbour exchange done in numpy.  If `record` is a dict, every message is
    appended to record[(src_rank, dst_rank)] in send order -- the same order the reference posts its MPI_Isend."""
    n = len(oracles)
    arrs = [c.arrays() for c in cases]
    D = oracles[0].D
    comm_type = cases[0].desc.c.comm_type
    recv = [{0: None, 1: None} for _ in range(n)]
    grads = scheme == 2 and bool(cases[0].opts.minmod)     # calc_gradients_M2AUSM (only solver 2 reads the minmod gradients)

    def exchange(step):
        spc = scalars_per_cell(D, comm_type, step)
        packs = [o.pack(step) for o in oracles]
        for r in range(n):
            nr = int(arrs[r]["recv_start"][-1]) if len(arrs[r]["nbr_rank"]) else 0
            recv[r][step] = np.zeros(nr * spc, dtype=oracles[r].real)
        for r in range(n):
            a = arrs[r]
            for i, nb in enumerate(a["nbr_rank"]):
                seg = packs[r][a["send_start"][i] * spc:a["send_start"][i + 1] * spc]
                if record is not None:
                    record.setdefault((r, int(nb)), []).append(seg.copy())
                b = arrs[nb]
                j = list(b["nbr_rank"]).index(r)
                recv[nb][step][b["recv_start"][j] * spc:b["recv_start"][j + 1] * spc] = seg

    def unpack(r, step):
        oracles[r].unpack(step, recv[r][step].ctypes.data)

    exchange(0)
    for r in range(n):
        oracles[r].set_bc()
        unpack(r, 0)
    exchange(1)
    for r in range(n):
        unpack(r, 1)
        oracles[r].vis(0)
    rk_order = cases[0].desc.c.rk_order
    for _ in range(n_steps):
        for o in oracles:
            o.prepare_timestep()
        for rk in range(rk_order):
            for r, o in enumerate(oracles):
                o.prepare_rkstep(rk)
                unpack(r, 0)
                o.set_bc()
                if grads:
                    o.gradients_ausm(0)
                o.vis(0)
            exchange(1)
            for r, o in enumerate(oracles):
                for s in range(1, cases[r].desc.n_sub):
                    if grads:
                        o.gradients_ausm(s)
                for s in range(1, cases[r].desc.n_sub):
                    o.vis(s)
                unpack(r, 1)
                o.rk_stage(0, scheme, rk, dt, None)
            exchange(0)
            for r, o in enumerate(oracles):
                for s in range(1, cases[r].desc.n_sub):
                    o.rk_stage(s, scheme, rk, dt, None)
        for r in range(n):
            unpack(r, 0)


def drive_rank(orc, exchange, scheme, dt, n_steps):
    """The same call order for ONE rank of a multi-process job: exchange(step) packs, moves the halo messages over the
    job's transport and unpacks them (ghost cells are only read by the boundary submesh's next call, so unpacking
    right after the exchange is equivalent to the reference's later mpi_wait)."""
    d = orc.case.desc
    grads = scheme == 2 and bool(orc.case.opts.minmod)
    exchange(0)
    orc.set_bc()
    exchange(1)
    orc.vis(0)
    for _ in range(n_steps):
        orc.prepare_timestep()
        for rk in range(d.c.rk_order):
            orc.prepare_rkstep(rk)
            orc.set_bc()
            if grads:
                orc.gradients_ausm(0)
            orc.vis(0)
            exchange(1)
            for s in range(1, d.n_sub):
                if grads:
                    orc.gradients_ausm(s)
            for s in range(1, d.n_sub):
                orc.vis(s)
            orc.rk_stage(0, scheme, rk, dt, None)
            exchange(0)
            for s in range(1, d.n_sub):
                orc.rk_stage(s, scheme, rk, dt, None)
