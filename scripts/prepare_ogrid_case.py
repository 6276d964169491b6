#!/usr/bin/env python
"""BASELINE config 4 analogue (examples/3D_Cylinder_Re3900 ships without its polyMesh): a synthetic O-grid around a unit
cylinder at the S-mesh size of the example (170 x 200 x 30 = 1.02 M hexahedra; wall cylinder, inlet/outlet outer boundary
at 38 D, cyclic span pi D, alpha = distance to the outer boundary), the dictionaries of examples/3D_Cylinder_Re3900/S
(solver 1 = M2, dimension 3, RK5, commType 2, haloCommType 1, dt 2e-3, Ls 13, M 0.2, mu 7.179e-5, Pr 0.75), decomposed into
8 ranks (2 radial x 4 azimuthal blocks), advanced N steps by the UNMODIFIED reference binary on the CPU.
The result (tmp_cases/ogrid8) is what scripts/run_example_dropin.py compares the 8-GPU drop-in run against.

    python scripts/prepare_ogrid_case.py [steps] [nr nth nz]
"""
import os
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lfm_public_b200.tools import casegen, meshgen  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    nr, nth, nz = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (170, 200, 30)
    dst = os.path.join(ROOT, "tmp_cases", "ogrid8")
    if os.path.exists(dst):
        shutil.rmtree(dst)
    t0 = time.time()
    m = meshgen.ogrid_cylinder(nr, nth, nz, r_in=0.5, r_out=38.0, span=np.pi, stretch=6.0)
    cell_rank = meshgen.block_assignment(m, (2, 4, 1))
    dt = 2.0e-3
    casegen.write_case(dst, m, cell_rank=cell_rank, solver=1, dimension=3, rkOrder=5, commType=2, haloCommType=1, deltaT=dt,
                       endTime=dt * steps, writeInterval=steps, Ls=13.0, M=0.2, mu=7.17948717948718e-05, Pr=0.75,
                       haveResiduals=False, printInfoFreq=1)
    print(f"case written in {time.time() - t0:.0f} s: {m['nCells']} cells, 8 ranks", flush=True)
    env = dict(os.environ, LFM_WRITE_PRECISION="17", LFM_MPI_NP="8")
    t0 = time.time()
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "lfm_solve_ref"), "-p"], cwd=dst, env=env, capture_output=True, text=True)
    open(os.path.join(dst, "log.ref"), "w").write(out.stdout + out.stderr)
    print(out.stdout[-600:])
    print(f"reference run (8 ranks on {os.cpu_count()} cores): {time.time() - t0:.0f} s")
    shutil.rmtree(os.path.join(dst, "output"), ignore_errors=True)


if __name__ == "__main__":
    main()
