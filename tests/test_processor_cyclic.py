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


def test_per_rank_generator_keeps_the_span_periodic(tmp_path):
    """meshgen.hex_block (the weak-scaling bench's per-rank generator) with z_cyclic=True and two blocks in z writes the same
    processorCyclic patches as decompose() of the global periodic box: rank by rank the flattened descriptors are identical,
    and so are the fields after the run."""
    from lfm_public_b200 import defs
    from lfm_public_b200.tools import casegen, meshgen
    n, blocks = (4, 4, 3), (2, 2, 2)
    G = tuple(a * b for a, b in zip(n, blocks))
    size = 1.0 / G[0]
    glob = meshgen.hex_box(*G, lengths=(G[0] * size, G[1] * size, G[2] * size), z_cyclic=True)
    parts = meshgen.decompose(glob, meshgen.block_assignment(glob, blocks))
    opts = host_api.default_opts(solver=1, dimension=3, delta_t=1e-3, Ls=0.1, mu0=7.17948717948718e-05, mach=0.2, comm_type=2)
    fields_g = casegen.synthetic_fields(glob, meshgen.cell_centres_estimate(glob))
    a_cases, b_cases = [], []
    for r in range(8):
        mb = meshgen.hex_block(n, blocks, r, z_cyclic=True)
        ma = parts[r]
        assert [(p["name"], p["type"], p["nFaces"]) for p in ma["patches"]] == [(p["name"], p["type"], p["nFaces"]) for p in mb["patches"]]
        assert sum(p["type"] == "processorCyclic" for p in mb["patches"]) == 1
        f = {k: np.asarray(v)[ma["cellProcAddressing"]] for k, v in fields_g.items()}
        f["alpha"] = np.ones(ma["nCells"])
        a_cases.append(host_api.Case.from_mesh(ma, opts, f, r, 8))
        b_cases.append(host_api.Case.from_mesh(mb, opts, f, r, 8))
    host_api.exchange_in_process(a_cases)
    host_api.exchange_in_process(b_cases)
    for ca, cb in zip(a_cases, b_cases):
        xa, xb = ca.arrays(), cb.arrays()
        for k in xa:
            assert np.array_equal(np.asarray(xa[k]), np.asarray(xb[k])), k
    oa, ob = [oracle_lib.Oracle(c) for c in a_cases], [oracle_lib.Oracle(c) for c in b_cases]
    oracle_lib.run(oa, 1, 1e-3, 3)
    oracle_lib.run(ob, 1, 1e-3, 3)
    for x, y in zip(oa, ob):
        q = x.download(defs.FIELD_Q)
        assert np.isfinite(q).all() and np.array_equal(q, y.download(defs.FIELD_Q))
