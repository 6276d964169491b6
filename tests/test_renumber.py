"""The standalone renumbering tool (the renumberMesh / hpathRenumber pre-processing step): the
renumbered case is a valid polyMesh describing the same problem, and the solve on it gives the same fields (up to the
summation-order round-off) in the permuted order."""
import os

import numpy as np
import pytest

import common
import oracle_lib
from lfm_public_b200 import host_api
from lfm_public_b200.tools import foamcase, meshgen, renumber


@pytest.mark.parametrize("method", ["morton", "hilbert", "rcm"])
def test_renumbered_case_solves_identically(method, tmp_path):
    case_dir = str(tmp_path / "tri")
    m, o = common.build_case("tri2d_m2", case_dir)          # shuffled numbering: the worst case for locality
    out_dir = str(tmp_path / ("tri_" + method))
    m0, m1, new_of_old = renumber.renumber_case(case_dir, out_dir, method)
    assert sorted(new_of_old.tolist()) == list(range(m0["nCells"]))
    # upper-triangular face order of OpenFOAM: internal faces sorted by (owner, neighbour), owner < neighbour
    nif = len(m1["neighbour"])
    ow, ne = m1["owner"][:nif].astype(np.int64), m1["neighbour"].astype(np.int64)
    assert (ow < ne).all() and (np.diff(ow * m1["nCells"] + ne) > 0).all()
    assert renumber.locality(m1)["median"] < renumber.locality(m0)["median"]
    # same geometry, cell by cell
    a = host_api.Case.open(case_dir).finish()
    b = host_api.Case.open(out_dir).finish()
    va, vb = a.geometry()["cell_volumes"], b.geometry()["cell_volumes"]
    assert np.allclose(vb[new_of_old], va, rtol=1e-13, atol=0)
    # same solution after N steps (the two numberings sum the same faces in a different order)
    oa, ob = oracle_lib.Oracle(a), oracle_lib.Oracle(b)
    oracle_lib.run([oa], o["solver"], o["deltaT"], common.N_STEPS)
    oracle_lib.run([ob], o["solver"], o["deltaT"], common.N_STEPS)
    qa, qb = a.to_mesh_order(oa.download(0)), b.to_mesh_order(ob.download(0))
    assert common.rel_max(qb[new_of_old], qa) < 1e-11
    # the reader round-trips what the generators write
    back = foamcase.read_polymesh(os.path.join(out_dir, "constant", "polyMesh"))
    assert np.array_equal(back["owner"], m1["owner"]) and np.array_equal(back["faces"], m1["faces"]) and np.array_equal(back["points"], m1["points"])
