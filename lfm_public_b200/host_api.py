"""Python view of the C++ host-side case setup (liblfmhost.so, include/lfmhost.h).

Mirrors what the reference does between `main` and `Mesh::solve` (reference: info/lfm_solve.cpp:98-128):
read the case, split submeshes, build geometry/ghosts/halo lists, set initial conditions -- and ends with
the flattened `lfmgpu_desc` that the CUDA library uploads.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._ctypes_defs import Desc, HostOpts, MeshIn

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblfmhost.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `make host` (or __graft_entry__.build())")
        L = C.CDLL(path)
        L.lfmhost_last_error.restype = C.c_char_p
        L.lfmhost_open_case.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.lfmhost_case_from_arrays.argtypes = [C.POINTER(MeshIn), C.POINTER(HostOpts), C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.lfmhost_default_opts.argtypes = [C.POINTER(HostOpts)]
        L.lfmhost_get_opts.argtypes = [C.c_void_p, C.POINTER(HostOpts)]
        L.lfmhost_nbr_count.argtypes = [C.c_void_p]
        L.lfmhost_nbr_rank.argtypes = [C.c_void_p, C.c_int]
        L.lfmhost_export.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.lfmhost_import.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
        L.lfmhost_finish.argtypes = [C.c_void_p]
        L.lfmhost_desc.argtypes = [C.c_void_p]
        L.lfmhost_desc.restype = C.POINTER(Desc)
        L.lfmhost_geometry.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lfmhost_mesh_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 4
        L.lfmhost_hpath_order.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lfmhost_write_field.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.lfmhost_close.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().lfmhost_last_error().decode())


def default_opts(**kw) -> HostOpts:
    o = HostOpts()
    lib().lfmhost_default_opts(C.byref(o))
    for k, v in kw.items():
        if k == "U_inf":
            for i in range(3):
                o.U_inf[i] = float(v[i])
        else:
            setattr(o, k, v)
    return o


def _np_view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    addr = ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class Case:
    """One rank's case: mesh + options + fields + (after finish()) the flattened descriptor."""

    def __init__(self, handle, keepalive=None):
        self._h = handle
        self._keep = keepalive
        self.finished = False

    @classmethod
    def open(cls, case_dir, rank=-1, n_ranks=1):
        h = C.c_void_p()
        _check(lib().lfmhost_open_case(str(case_dir).encode(), rank, n_ranks, C.byref(h)))
        return cls(h)

    @classmethod
    def from_mesh(cls, m, opts: HostOpts, fields, rank=0, n_ranks=1):
        """m: mesh dict of tools.meshgen; fields: dict(p, T, U[, alpha]) in polyMesh cell order."""
        pts = np.ascontiguousarray(m["points"], dtype=np.float64)
        faces = np.ascontiguousarray(m["faces"], dtype=np.int32)
        owner = np.ascontiguousarray(m["owner"], dtype=np.int32)
        nei = np.ascontiguousarray(m["neighbour"], dtype=np.int32)
        P = m["patches"]
        n = len(P)
        names = (C.c_char_p * n)(*[p["name"].encode() for p in P])
        types = (C.c_char_p * n)(*[p["type"].encode() for p in P])
        nbrn = (C.c_char_p * n)(*[p.get("neighbourPatch", "").encode() for p in P])
        refn = (C.c_char_p * n)(*[p.get("referPatch", "").encode() for p in P])
        nf = np.asarray([p["nFaces"] for p in P], dtype=np.int32)
        st = np.asarray([p["startFace"] for p in P], dtype=np.int32)
        myp = np.asarray([p.get("myProcNo", -1) for p in P], dtype=np.int32)
        nbp = np.asarray([p.get("neighbProcNo", -1) for p in P], dtype=np.int32)
        fpa = np.ascontiguousarray(m["faceProcAddressing"], dtype=np.int32) if "faceProcAddressing" in m else None
        csm = np.ascontiguousarray(m["cellSubmesh"], dtype=np.int32) if "cellSubmesh" in m else None
        i32p = C.POINTER(C.c_int32)
        mi = MeshIn(len(pts), pts.ctypes.data_as(C.POINTER(C.c_double)), len(faces), faces.ctypes.data_as(i32p),
                    owner.ctypes.data_as(i32p), len(nei), nei.ctypes.data_as(i32p), int(m["nCells"]), n, names, types,
                    nf.ctypes.data_as(i32p), st.ctypes.data_as(i32p), nbrn, myp.ctypes.data_as(i32p), nbp.ctypes.data_as(i32p),
                    fpa.ctypes.data_as(i32p) if fpa is not None else None, csm.ctypes.data_as(i32p) if csm is not None else None, refn)
        p = np.ascontiguousarray(fields["p"], dtype=np.float64)
        T = np.ascontiguousarray(fields["T"], dtype=np.float64)
        U = np.ascontiguousarray(fields["U"], dtype=np.float64)
        a = np.ascontiguousarray(fields["alpha"], dtype=np.float64) if fields.get("alpha") is not None else None
        h = C.c_void_p()
        _check(lib().lfmhost_case_from_arrays(C.byref(mi), C.byref(opts), p.ctypes.data, T.ctypes.data, U.ctypes.data,
                                              a.ctypes.data if a is not None else None, rank, n_ranks, C.byref(h)))
        return cls(h)

    # -- neighbour exchange ---------------------------------------------------------------------------
    @property
    def nbr_ranks(self):
        return [lib().lfmhost_nbr_rank(self._h, i) for i in range(lib().lfmhost_nbr_count(self._h))]

    def export(self, i) -> bytes:
        d = C.c_void_p()
        n = C.c_size_t()
        _check(lib().lfmhost_export(self._h, i, C.byref(d), C.byref(n)))
        return C.string_at(d, n.value)

    def import_(self, i, data: bytes):
        _check(lib().lfmhost_import(self._h, i, data, len(data)))

    def finish(self):
        _check(lib().lfmhost_finish(self._h))
        self.finished = True
        return self

    # -- views ------------------------------------------------------------------------------------------
    @property
    def opts(self) -> HostOpts:
        o = HostOpts()
        lib().lfmhost_get_opts(self._h, C.byref(o))
        return o

    @property
    def desc_ptr(self):
        assert self.finished
        return lib().lfmhost_desc(self._h)

    @property
    def desc(self) -> Desc:
        return self.desc_ptr.contents

    def arrays(self):
        """numpy views of the descriptor arrays (no copies)."""
        d = self.desc
        real = np.float64 if d.precision == 8 else np.float32
        D, F = d.dim, d.max_slots
        nn = d.n_nbr
        send_start = _np_view(d.send_start, nn + 1, np.int32)
        out = dict(
            face_owner=_np_view(d.face_owner, d.n_faces, np.int32), face_neigh=_np_view(d.face_neigh, d.n_faces, np.int32),
            face_S=_np_view(d.face_S, d.n_faces * D, real).reshape(-1, D), face_d=_np_view(d.face_d, d.n_faces * D, real).reshape(-1, D),
            face_w=_np_view(d.face_w, d.n_faces, real), vol_inv=_np_view(d.vol_inv, d.n_cells, real),
            sponge_sigma=_np_view(d.sponge_sigma, d.n_cells, real), q0=_np_view(d.q0, d.n_cells * (D + 2), real).reshape(-1, D + 2),
            cell_gid=_np_view(d.cell_gid, d.n_cells, np.int32),
            cell_slot_face=_np_view(d.cell_slot_face, d.n_cells * F, np.int32).reshape(-1, F),
            bc_cell=_np_view(d.bc_cell, d.n_bc_ghosts, np.int32), bc_kind=_np_view(d.bc_kind, d.n_bc_ghosts, np.int32),
            bc_patch=_np_view(d.bc_patch, d.n_bc_ghosts, np.int32), bc_face=_np_view(d.bc_face, d.n_bc_ghosts, np.int32),
            nbr_rank=_np_view(d.nbr_rank, nn, np.int32), send_start=send_start,
            send_cell=_np_view(d.send_cell, int(send_start[-1]) if nn else 0, np.int32),
            recv_start=_np_view(d.recv_start, nn + 1, np.int32),
        )
        return out

    def geometry(self):
        npt, nf, ni, nc = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        lib().lfmhost_mesh_sizes(self._h, C.byref(npt), C.byref(nf), C.byref(ni), C.byref(nc))
        fa = np.zeros((nf.value, 3)); fc = np.zeros((nf.value, 3)); cc = np.zeros((nc.value, 3)); cv = np.zeros(nc.value)
        lib().lfmhost_geometry(self._h, fa.ctypes.data, fc.ctypes.data, cc.ctypes.data, cv.ctypes.data)
        return dict(face_areas=fa, face_centres=fc, cell_centres=cc, cell_volumes=cv)

    def hpath_order(self):
        """Hamiltonian-path numbering of the case's mesh (hpathRenumber plugin restated): (order[new] = old, stats dict)."""
        nc = C.c_int32()
        lib().lfmhost_mesh_sizes(self._h, None, None, None, C.byref(nc))
        order = np.zeros(nc.value, dtype=np.int32)
        st = np.zeros(4)
        _check(lib().lfmhost_hpath_order(self._h, order.ctypes.data, st.ctypes.data))
        return order, dict(boundary_cells=int(st[0]), boundary_walk_ok=bool(st[1]), interior_walk_ok=bool(st[2]), interior_path_fraction=float(st[3]))

    def to_mesh_order(self, values):
        """[n_cells, ...] in traversal order -> polyMesh cell order."""
        gid = self.arrays()["cell_gid"]
        out = np.empty_like(values)
        out[gid] = values
        return out

    def close(self):
        if self._h:
            lib().lfmhost_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def exchange_in_process(cases):
    """Setup-time neighbour exchange when all ranks live in this process (cases[r] is rank r)."""
    for r, c in enumerate(cases):
        for i, nb in enumerate(c.nbr_ranks):
            other = cases[nb]
            j = other.nbr_ranks.index(r)
            other.import_(j, c.export(i))
    for c in cases:
        c.finish()
    return cases


def exchange_distributed(case: Case, rank: int):
    """Setup-time neighbour exchange over torch.distributed (gloo or nccl world already initialised)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    payload = {nb: case.export(i) for i, nb in enumerate(case.nbr_ranks)}
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)
    for i, nb in enumerate(case.nbr_ranks):
        case.import_(i, gathered[nb][rank])
    case.finish()
    return case
