"""Degenerate inputs on the CPU side: meshes whose cells all touch the boundary (one submesh only), a single cell, the
smallest periodic box that still has an interior submesh.  The restatement and the host-side flattening must follow the
reference's own executable bit for bit there too (the GPU counterparts of these cases are listed in DESIGN.md 9)."""
import os

import numpy as np
import pytest

import common
import oracle_lib
from lfm_public_b200 import host_api
from lfm_public_b200.tools import casegen, meshgen

N = 4

EDGE = {
    "quad_3x2_all_boundary": (lambda: meshgen.hex_box(3, 2, 1, lengths=(3, 2, 0.1), two_d=True), True, dict(solver=0, dimension=2, deltaT=1e-3, Ls=0.5), 1),
    "quad_single_cell": (lambda: meshgen.hex_box(1, 1, 1, lengths=(1, 1, 0.1), two_d=True), True, dict(solver=1, dimension=2, deltaT=1e-3), 1),
    "hex_2x2x2_all_boundary": (lambda: meshgen.hex_box(2, 2, 2, lengths=(1, 1, 1), z_cyclic=False), False, dict(solver=1, dimension=3, deltaT=1e-3, mu=7e-5), 1),
    "hex_3x3x3_periodic_ausm": (lambda: meshgen.hex_box(3, 3, 3, lengths=(1, 1, 1), z_cyclic=True), False,
                                dict(solver=2, dimension=3, deltaT=1e-3, mu=7e-5, minmodExists=True), 2),
}


@pytest.mark.skipif(not common.have_ref(), reason="reference binary not built")
@pytest.mark.parametrize("name", sorted(EDGE))
def test_edge_case_matches_reference(name, tmp_path):
    mesh, two_d, opts, n_sub = EDGE[name]
    d = str(tmp_path / name)
    o = casegen.write_case(d, mesh(), two_d=two_d, endTime=opts["deltaT"] * N, writeInterval=N, **opts)
    o["n_ranks"], o["parallel"] = 1, False
    common.run_reference(d, o, dump=False)
    c = host_api.Case.open(d).finish()
    assert c.desc.n_sub == n_sub
    orc = oracle_lib.Oracle(c)
    oracle_lib.run([orc], o["solver"], o["deltaT"], N)
    ref = common.read_reference_q(d, o, o["deltaT"] * N, o["dimension"])[0]
    mine = common.primitives_from_q(c.to_mesh_order(orc.download(0)), c.desc.c.gamma_m1)
    assert np.isfinite(ref["rho"]).all() and len(ref["rho"]) == c.desc.n_cells
    for k in ("rho", "U", "E", "p"):
        assert np.array_equal(mine[k], ref[k]), f"{name} {k}"


def test_farfield_patch_is_refused(tmp_path):
    """`type farfield` (cfd_v0.cpp:1136-1215): the reference's ghost state there depends on m_dAoA, which it never assigns, so the
    GPU path's host side refuses the case loudly instead of advancing it with ghosts nobody set."""
    import re
    from lfm_public_b200 import host_api
    from lfm_public_b200.tools import casegen, meshgen
    d = str(tmp_path / "ff")
    casegen.write_case(d, meshgen.hex_box(4, 3, 1, lengths=(4, 3, 0.1), two_d=True), two_d=True, solver=0, dimension=2, deltaT=1e-3, endTime=1e-2)
    bfile = os.path.join(d, "constant", "polyMesh", "boundary")
    s = open(bfile).read()
    s2, n = re.subn(r"(outlet\s*\{\s*type\s+)patch", r"\1farfield", s)
    assert n == 1
    open(bfile, "w").write(s2)
    with pytest.raises(Exception, match="farfield"):
        host_api.Case.open(d).finish()
