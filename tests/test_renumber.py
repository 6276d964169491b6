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


@pytest.mark.parametrize("method", ["morton", "hilbert", "rcm", "hpath"])
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


@pytest.mark.parametrize("name", ["tri2d_m2", "quad2d_m1", "ogrid2d_m1"])
def test_hpath_numbering_invariants(name, tmp_path):
    """The restated hpathRenumber plugin: a permutation, the boundary submesh first (the reference's own split: cells with a
    point on a non-empty boundary face), and a path -- most consecutive cells of the interior part share a face."""
    case_dir = str(tmp_path / name)
    m, o = common.build_case(name, case_dir)
    case = host_api.Case.open(case_dir)
    order, st = case.hpath_order()
    n = len(order)
    assert sorted(order.tolist()) == list(range(n))
    pm = foamcase.read_polymesh(os.path.join(case_dir, "constant", "polyMesh"))
    # the reference's submesh split, restated with numpy
    bpts = np.zeros(len(pm["points"]), bool)
    nif = len(pm["neighbour"])
    valid = np.ones(len(pm["owner"]), bool)
    for p in pm["patches"]:
        if p["type"] == "empty":
            valid[p["startFace"]:p["startFace"] + p["nFaces"]] = False
    bf = np.nonzero(valid[nif:])[0] + nif
    fp = pm["faces"][bf]
    bpts[fp[fp >= 0]] = True
    sub = np.ones(n, int)
    sub[pm["owner"][bf]] = 0
    touch = (bpts[np.maximum(pm["faces"][:nif], 0)] & (pm["faces"][:nif] >= 0)).any(1)
    sub[pm["owner"][:nif][touch]] = 0
    sub[pm["neighbour"][touch]] = 0
    nb = int((sub == 0).sum())
    assert st["boundary_cells"] == nb and (sub[order[:nb]] == 0).all() and (sub[order[nb:]] == 1).all()
    # adjacency along the numbering
    adj = set()
    for a, b in zip(pm["owner"][:nif].tolist(), pm["neighbour"].tolist()):
        adj.add((a, b)); adj.add((b, a))
    inner = order[nb:]
    if len(inner) > 10:
        assert st["interior_walk_ok"]
        hops = sum((int(a), int(b)) in adj for a, b in zip(inner[:-1], inner[1:]))
        k = int(round(st["interior_path_fraction"] * len(inner)))
        assert k >= 0.6 * len(inner)                        # most of the interior lies on the path itself
        assert hops >= 0.95 * (k - 1)                       # and the path moves from a cell to one of its neighbours
    outer = order[:nb]
    hops = sum((int(a), int(b)) in adj for a, b in zip(outer[:-1], outer[1:]))
    assert hops >= 0.5 * (nb - 1)
    case.close()
