"""processorCyclic patches (a cyclic pair cut by the decomposition; reference: polyMeshReaderOF.cpp:251-261 getFaceId,
:460-475 getBoundaryTag, mesh_reader.cpp:404-473 faceTagMapping): the decomposed run must be the serial periodic run up
to summation order, and the patches must be what decomposePar would write."""
import os

import numpy as np
import pytest

import common
import oracle_lib
from lfm_public_b200 import host_api
from lfm_public_b200.tools import foamcase


@pytest.mark.parametrize("name", ["hex3d_m2_pc8", "hex3d_m1_pc2"])
def test_decomposed_periodic_run_equals_serial(name, tmp_path):
    d = str(tmp_path / name)
    m, o = common.build_case(name, d)
    ranks = common.open_ranks(d, o)
    oracles = [oracle_lib.Oracle(c) for c in ranks]
    oracle_lib.run(oracles, o["solver"], o["deltaT"], common.N_STEPS)
    serial = host_api.Case.open(d).finish()
    so = oracle_lib.Oracle(serial)
    oracle_lib.run([so], o["solver"], o["deltaT"], common.N_STEPS)
    qs = serial.to_mesh_order(so.download(0))
    n_pc = 0
    for r, (c, orc) in enumerate(zip(ranks, oracles)):
        md = os.path.join(d, f"processor{r}", "constant", "polyMesh")
        n, t = foamcase._body(os.path.join(md, "cellProcAddressing"))
        idx = foamcase._numbers(t, np.int64)
        q = c.to_mesh_order(orc.download(0))
        assert common.rel_max(q, qs[idx]) < 1e-13
        b = open(os.path.join(md, "boundary")).read()
        n_pc += b.count("type            processorCyclic")
        for line in b.splitlines():
            if "through" in line:
                assert line.strip().startswith(f"procBoundary{r}to") and line.strip().split("through")[1] in ("periodic_m", "periodic_p")
    assert n_pc == len(ranks)        # one processorCyclic patch per rank: its half of the cut periodic plane
