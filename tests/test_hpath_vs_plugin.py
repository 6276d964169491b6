"""The hpath numbering restated in lfm_public_b200/host/hpath.cpp against the reference's OWN renumberMesh plugin:
oracle/_ref/hpath_plugin is /root/reference/hpathRenumber/hpathRenumber.C compiled unchanged against a stand-in for the few
OpenFOAM types it touches (oracle/shim/openfoam_stub/, oracle/Makefile target `hpath`).  Both sides read the same polyMesh
through the OpenFOAM-free reader (the driver also checks that the reader's cells() is what primitiveMesh::calcCells builds) and
see the same cell centres, so equality of the two orders pins the ALGORITHM: submesh split, boundary walk, starting cells,
both next-cell rules, the step-back, dead ends.  Cases: the zoo's 2D/3D meshes, larger shuffled triangle and O-grid meshes
(dead ends, several trials) and -- where /root/reference is present -- the meshes of the reference's example cases
(525 000 quads, 1 959 342 triangle prisms), whose plugin orders are also committed as SHA-256 fixtures
(tests/golden/hpath_examples.json, scripts/make_hpath_fixtures.py)."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import common
from lfm_public_b200 import host_api
from lfm_public_b200.tools import casegen, meshgen

PLUGIN = os.path.join(common.ROOT, "oracle", "_ref", "hpath_plugin")
needs_plugin = pytest.mark.skipif(not os.path.exists(PLUGIN), reason="oracle/_ref/hpath_plugin not built (needs /root/reference)")


def _plugin_order(case_dir, tmp_path):
    out = str(tmp_path / "order.bin")
    r = subprocess.run([PLUGIN, os.path.join(case_dir, "constant", "polyMesh"), out], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.fromfile(out, dtype=np.int32)


@needs_plugin
@pytest.mark.parametrize("name", ["tri2d_m2", "quad2d_m1", "ogrid2d_m1", "hex3d_m2", "ogrid3d_m2"])
def test_zoo_meshes(name, tmp_path):
    case_dir = str(tmp_path / name)
    common.build_case(name, case_dir)
    mine, st = host_api.Case.open(case_dir).hpath_order()
    assert np.array_equal(mine, _plugin_order(case_dir, tmp_path)), f"{name}: {st}"


LARGER = {
    "tri_60x48_shuffled": (lambda: meshgen.tri_prism_box(60, 48, lengths=(2.0, 1.6), shuffle_seed=3), True),
    "tri_150x120_shuffled": (lambda: meshgen.tri_prism_box(150, 120, lengths=(2.0, 1.6), shuffle_seed=11), True),
    "ogrid2d_48x160": (lambda: meshgen.ogrid_cylinder(48, 160, 1, r_in=0.5, r_out=6.0, two_d=True, stretch=4.0), True),
    "hex_24x20x16": (lambda: meshgen.hex_box(24, 20, 16, lengths=(2.0, 1.5, 1.0), z_cyclic=True), False),
    "ogrid3d_16x48x8": (lambda: meshgen.ogrid_cylinder(16, 48, 8, r_in=0.5, r_out=6.0, span=1.0, stretch=4.0), False),
}


@needs_plugin
@pytest.mark.parametrize("name", sorted(LARGER))
def test_larger_meshes(name, tmp_path):
    make, two_d = LARGER[name]
    case_dir = str(tmp_path / name)
    casegen.write_case(case_dir, make(), two_d=two_d, solver=1, dimension=2 if two_d else 3, deltaT=1e-3, endTime=1e-3)
    mine, st = host_api.Case.open(case_dir).hpath_order()
    ref = _plugin_order(case_dir, tmp_path)
    assert sorted(ref.tolist()) == list(range(len(ref)))
    assert np.array_equal(mine, ref), f"{name}: first difference at {int(np.nonzero(mine != ref)[0][0])}; {st}"


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="the reference's example meshes are not here")
@pytest.mark.parametrize("name", ["cylinder_vortex", "cylinder_vortex_unstructured"])
def test_reference_example_meshes_match_the_committed_plugin_orders(name, tmp_path):
    """BASELINE.json config 2 is 'the unstructured cylinder with hpath renumbering': the restatement's order on the reference's
    own meshes equals what the plugin returned for them (fixture made by the plugin, scripts/make_hpath_fixtures.py)."""
    sys.path.insert(0, os.path.join(common.ROOT, "scripts"))
    import make_hpath_fixtures
    fix = json.load(open(os.path.join(common.GOLDEN_DIR, "hpath_examples.json")))[name]
    case_dir = make_hpath_fixtures.extract(name, str(tmp_path))
    mine, st = host_api.Case.open(case_dir).hpath_order()
    assert len(mine) == fix["n_cells"]
    assert hashlib.sha256(np.ascontiguousarray(mine, dtype=np.int32).tobytes()).hexdigest() == fix["sha256"], st
