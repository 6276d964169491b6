"""Shared test helpers: the case zoo, running the reference binary (oracle/_ref), reading OpenFOAM fields."""
from __future__ import annotations

import os
import re
import shutil
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from lfm_public_b200 import host_api  # noqa: E402
from lfm_public_b200.tools import casegen, meshgen  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "lfm_solve_ref")
REF_BIN_SP = os.path.join(ROOT, "oracle", "_ref", "lfm_solve_ref_sp")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def have_ref(sp=False):
    return os.path.exists(REF_BIN_SP if sp else REF_BIN)


# ------------------------------------------------------------------------------------------------
# case zoo: name -> builder returning (mesh, blocks or None, two_d, dict options)
# Small on purpose: the oracle and the reference finish each in well under a second.
# ------------------------------------------------------------------------------------------------

def _quad2d():
    return meshgen.hex_box(16, 12, 1, lengths=(4.0, 3.0, 0.1), two_d=True)


def _hex3d(z_cyclic=True):
    return meshgen.hex_box(10, 8, 6, lengths=(2.0, 1.5, 1.0), z_cyclic=z_cyclic)


CASES = {
    # 2D structured, M1, sponge on, serial  (config 1 analogue)
    "quad2d_m1": dict(mesh=_quad2d, two_d=True, blocks=None,
                      opts=dict(solver=0, dimension=2, deltaT=2e-3, Ls=1.0, haveResiduals=True)),
    # same mesh, 4 ranks, M1, split halo payload (commType 2)
    "quad2d_m1_p4": dict(mesh=_quad2d, two_d=True, blocks=(2, 2, 1),
                         opts=dict(solver=0, dimension=2, deltaT=2e-3, Ls=1.0, commType=2)),
    # 2 ranks, packed payload (commType 1), M2  (config 3 analogue)
    "quad2d_m2_p2_packed": dict(mesh=_quad2d, two_d=True, blocks=(2, 1, 1),
                                opts=dict(solver=1, dimension=2, deltaT=2e-3, Ls=1.0, commType=1)),
    # 2D unstructured (triangle prisms, shuffled numbering), M2  (config 2 analogue)
    "tri2d_m2": dict(mesh=lambda: meshgen.tri_prism_box(10, 8, lengths=(2.0, 1.6), shuffle_seed=7), two_d=True, blocks=None,
                     opts=dict(solver=1, dimension=2, deltaT=2e-3, Ls=0.5)),
    # 3D hexes with a cyclic span, M2, serial and 2x2 ranks  (config 4/5 analogue)
    "hex3d_m2": dict(mesh=_hex3d, two_d=False, blocks=None,
                     opts=dict(solver=1, dimension=3, deltaT=2e-3, Ls=0.4, mu=7.17948717948718e-05)),
    "hex3d_m2_p4": dict(mesh=_hex3d, two_d=False, blocks=(2, 2, 1),
                        opts=dict(solver=1, dimension=3, deltaT=2e-3, Ls=0.4, commType=2, mu=7.17948717948718e-05)),
    "hex3d_m1_p8": dict(mesh=lambda: _hex3d(False), two_d=False, blocks=(2, 2, 2),
                        opts=dict(solver=0, dimension=3, deltaT=2e-3, Ls=0.4, commType=2)),
    # O-grid cylinder with wall + inlet/outlet + cyclic span, 3D M2 (3D_Cylinder_Re3900 analogue)
    "ogrid3d_m2": dict(mesh=lambda: meshgen.ogrid_cylinder(8, 16, 4, r_in=0.5, r_out=6.0, span=1.0, stretch=4.0), two_d=False,
                       blocks=None, opts=dict(solver=1, dimension=3, deltaT=1e-3, Ls=2.0, mu=7.17948717948718e-05, haveForces=True)),
    # LES closure: calc_VIS_Smagorinsky instead of calc_VIS (turbulenceProperties simulationType LES)
    "hex3d_m2_les_p4": dict(mesh=_hex3d, two_d=False, blocks=(2, 2, 1),
                            opts=dict(solver=1, dimension=3, deltaT=2e-3, Ls=0.4, commType=2, mu=7.17948717948718e-05, simulationType="LES")),
    "ogrid3d_m1_les": dict(mesh=lambda: meshgen.ogrid_cylinder(8, 16, 4, r_in=0.5, r_out=6.0, span=1.0, stretch=4.0), two_d=False,
                           blocks=None, opts=dict(solver=0, dimension=3, deltaT=1e-3, Ls=2.0, mu=7.17948717948718e-05, simulationType="LES")),
    # z-periodic box split in z too: the cyclic pair is cut by the decomposition -> processorCyclic patches (faceId = index
    # inside the patch, hashed tag; polyMeshReaderOF.cpp:251-261, 460-475), next to plain processor patches of the same rank pair
    "hex3d_m2_pc8": dict(mesh=_hex3d, two_d=False, blocks=(2, 2, 2),
                         opts=dict(solver=1, dimension=3, deltaT=2e-3, Ls=0.4, commType=2, mu=7.17948717948718e-05)),
    "hex3d_m1_pc2": dict(mesh=_hex3d, two_d=False, blocks=(1, 1, 2),
                         opts=dict(solver=0, dimension=3, deltaT=2e-3, Ls=0.4, commType=1, mu=7.17948717948718e-05)),
    # solver 2: M2 + AUSM+up pressure dissipation with minmod-limited reconstruction (one_rk_step_M2AUSM)
    "hex3d_ausm_p4": dict(mesh=_hex3d, two_d=False, blocks=(2, 2, 1),
                          opts=dict(solver=2, dimension=3, deltaT=2e-3, Ls=0.4, commType=2, mu=7.17948717948718e-05, minmodExists=True)),
    "ogrid2d_ausm": dict(mesh=lambda: meshgen.ogrid_cylinder(8, 16, 1, r_in=0.5, r_out=6.0, two_d=True, stretch=4.0), two_d=True,
                         blocks=None, opts=dict(solver=2, dimension=2, deltaT=1e-3, Ls=2.0, minmodExists=True)),
    "quad2d_ausm_nominmod": dict(mesh=_quad2d, two_d=True, blocks=None, opts=dict(solver=2, dimension=2, deltaT=2e-3, Ls=1.0)),
    "ogrid2d_m1": dict(mesh=lambda: meshgen.ogrid_cylinder(8, 16, 1, r_in=0.5, r_out=6.0, two_d=True, stretch=4.0), two_d=True,
                       blocks=None, opts=dict(solver=0, dimension=2, deltaT=1e-3, Ls=2.0, haveForces=True, haveAverage=True)),
}

N_STEPS = 6


def build_case(name, case_dir, n_steps=N_STEPS, spec=None, **override):
    spec = spec or CASES[name]
    m = spec["mesh"]()
    opts = dict(spec["opts"])
    opts.update(override)
    dt = opts["deltaT"]
    opts.setdefault("endTime", dt * n_steps)
    opts.setdefault("writeInterval", n_steps)
    cell_rank = meshgen.block_assignment(m, spec["blocks"]) if spec["blocks"] else None
    if os.path.exists(case_dir):
        shutil.rmtree(case_dir)
    o = casegen.write_case(case_dir, m, cell_rank=cell_rank, two_d=spec["two_d"], **opts)
    o["n_ranks"] = int(cell_rank.max()) + 1 if cell_rank is not None else 1
    o["parallel"] = cell_rank is not None
    return m, o


def open_ranks(case_dir, o):
    """Host-side setup of every rank of a case, in this process."""
    if not o["parallel"]:
        return [host_api.Case.open(case_dir).finish()]
    cases = [host_api.Case.open(case_dir, r, o["n_ranks"]) for r in range(o["n_ranks"])]
    return host_api.exchange_in_process(cases)


# ------------------------------------------------------------------------------------------------
# reference binary
# ------------------------------------------------------------------------------------------------

def run_reference(case_dir, o, sp=False, dump=True, timeout=600):
    env = dict(os.environ)
    env["LFM_WRITE_PRECISION"] = "17"
    args = [REF_BIN_SP if sp else REF_BIN]
    if o["parallel"]:
        env["LFM_MPI_NP"] = str(o["n_ranks"])
        args.append("-p")
        if dump:
            d = os.path.join(case_dir, "dump")
            os.makedirs(d, exist_ok=True)
            env["LFM_MPI_DUMP_DIR"] = d
    out = subprocess.run(args, cwd=case_dir, env=env, capture_output=True, text=True, timeout=timeout)
    if "Simulation finished successfully" not in out.stdout:
        raise RuntimeError("reference run failed:\n" + out.stdout[-2000:] + out.stderr[-2000:])
    return out.stdout


def read_field(path, ncomp=1):
    raw = open(path, "rb").read()
    head = raw[:raw.index(b"}") + 1]
    if b"binary" in head:                               # OpenFOAM binary stream format: N( raw doubles )
        i = raw.index(b"internalField")
        m = re.search(rb"(\d+)\s*\(", raw[i:])
        n = int(m.group(1))
        a0 = i + m.end()
        a = np.frombuffer(raw[a0:a0 + 8 * n * ncomp], dtype="<f8").copy()
        assert raw[a0 + 8 * n * ncomp:a0 + 8 * n * ncomp + 1] == b")"
        return a.reshape(-1, ncomp) if ncomp > 1 else a
    s = raw.decode()
    i = s.index("internalField")
    j = s.index("boundaryField")
    body = s[s.index("(", i) + 1:s.rindex(")", i, j)]
    a = np.array(body.replace("(", " ").replace(")", " ").split(), dtype=np.float64)
    return a.reshape(-1, ncomp) if ncomp > 1 else a


def time_name(t):
    return "%.12g" % t


def read_reference_q(case_dir, o, t, D):
    """Conservative-equivalent primitives the reference writes (rho, U, E, p) per rank, in polyMesh cell order
    (reference: src/cfd_v0.cpp:3000-3033 updateSolutionPrimitives)."""
    out = []
    dirs = [os.path.join(case_dir, f"processor{r}") for r in range(o["n_ranks"])] if o["parallel"] else [case_dir]
    for d in dirs:
        td = os.path.join(d, time_name(t))
        out.append(dict(rho=read_field(os.path.join(td, "rho")), U=read_field(os.path.join(td, "U"), 3)[:, :D],
                        E=read_field(os.path.join(td, "E")), p=read_field(os.path.join(td, "p"))))
    return out


def primitives_from_q(q, gm1):
    """The same derived fields from conservatives, with the reference's expressions (cfd_v0.cpp:3017-3025)."""
    D = q.shape[1] - 2
    r = q[:, 0]
    uvw = q[:, 1:D + 1] / r[:, None]
    Umag2 = np.zeros_like(r)
    for k in range(D):
        Umag2 = Umag2 + uvw[:, k] * uvw[:, k]
    E = q[:, D + 1] / r
    p = r * gm1 * (E - 0.5 * Umag2)
    return dict(rho=r, U=uvw, E=E, p=p)


def read_dump(case_dir, src, dst, dtype):
    """Messages rank src sent to rank dst, in order: list of (tag, np.array)."""
    path = os.path.join(case_dir, "dump", f"send_r{src}_to{dst}.bin")
    out = []
    with open(path, "rb") as f:
        while True:
            h = f.read(8)
            if len(h) < 8:
                break
            tag, nbytes = struct.unpack("ii", h)
            out.append((tag, np.frombuffer(f.read(nbytes), dtype=dtype)))
    return out


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))
