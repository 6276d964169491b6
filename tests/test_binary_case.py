"""OpenFOAM binary stream format (`writeFormat binary`; what decomposePar leaves for the reference's 3D cases,
examples/3D_Cylinder_Re3900/*/system/controlDict): the OpenFOAM-free host setup reads binary polyMesh lists
(points, faceCompactList faces, owner/neighbour, *ProcAddressing) and binary nonuniform fields, and the flattened
rank it hands to lfmgpu_create is identical to the one built from the ASCII case."""
import shutil

import numpy as np
import pytest

import common
from common import CASES
from lfm_public_b200.tools import foamcase


@pytest.mark.parametrize("label_bytes", [4, 8])
@pytest.mark.parametrize("name", ["tri2d_m2", "hex3d_m2_p4", "quad2d_m1_p4"])
def test_binary_case_flattens_identically(name, label_bytes, tmp_path):
    a_dir, b_dir = str(tmp_path / "ascii"), str(tmp_path / "binary")
    m, o = common.build_case(name, a_dir)
    shutil.copytree(a_dir, b_dir)
    foamcase.to_binary(b_dir, label_bytes=label_bytes)
    assert b"format      binary" in open(f"{b_dir}/{'processor0/' if o['parallel'] else ''}constant/polyMesh/owner", "rb").read(300)
    ca, cb = common.open_ranks(a_dir, o), common.open_ranks(b_dir, o)
    assert len(ca) == len(cb) == o["n_ranks"]
    for x, y in zip(ca, cb):
        ax, ay = x.arrays(), y.arrays()
        assert ax.keys() == ay.keys()
        for k in ax:
            assert np.array_equal(np.asarray(ax[k]), np.asarray(ay[k])), k
        assert x.desc.n_cells == y.desc.n_cells and x.desc.n_faces == y.desc.n_faces


def test_binary_errors(tmp_path):
    a_dir = str(tmp_path / "c")
    common.build_case("tri2d_m2", a_dir)
    foamcase.to_binary(a_dir)
    p = f"{a_dir}/constant/polyMesh/owner"
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:len(raw) // 2])          # truncated raw block
    from lfm_public_b200 import host_api
    with pytest.raises(Exception):
        host_api.Case.open(a_dir).finish()


@pytest.mark.skipif(not common.have_ref(), reason="reference binary not built")
@pytest.mark.parametrize("name", ["quad2d_m1_p4", "tri2d_m2"])
def test_reference_binary_reads_binary_case(name, tmp_path):
    """The reference's own executable (built against the OpenFOAM-free adapters, which use the same reader) advances the
    binary case to the same fields as the ASCII case."""
    a_dir, b_dir = str(tmp_path / "ascii"), str(tmp_path / "binary")
    m, o = common.build_case(name, a_dir)
    shutil.copytree(a_dir, b_dir)
    foamcase.to_binary(b_dir)
    common.run_reference(a_dir, o, dump=False)
    common.run_reference(b_dir, o, dump=False)
    t = o["deltaT"] * common.N_STEPS
    qa, qb = common.read_reference_q(a_dir, o, t, o["dimension"]), common.read_reference_q(b_dir, o, t, o["dimension"])
    for ra, rb in zip(qa, qb):
        assert ra["rho"].std() > 0
        for k in ("rho", "U", "E", "p"):
            assert np.array_equal(ra[k], rb[k]), k


@pytest.mark.skipif(not common.have_ref(), reason="reference binary not built")
def test_write_format_binary(tmp_path):
    """controlDict `writeFormat binary` (the setting of the reference's 3D cases): the time directory is written in the binary
    stream format and holds exactly the doubles of the 17-digit ASCII run."""
    a_dir, b_dir = str(tmp_path / "ascii"), str(tmp_path / "binary")
    m, o = common.build_case("hex3d_m2_p4", a_dir)
    common.build_case("hex3d_m2_p4", b_dir, writeFormat="binary")
    common.run_reference(a_dir, o, dump=False)
    common.run_reference(b_dir, o, dump=False)
    t = o["deltaT"] * common.N_STEPS
    assert b"format      binary" in open(f"{b_dir}/processor0/{common.time_name(t)}/rho", "rb").read(200)
    qa, qb = common.read_reference_q(a_dir, o, t, 3), common.read_reference_q(b_dir, o, t, 3)
    for ra, rb in zip(qa, qb):
        for k in ("rho", "U", "E", "p"):
            assert np.array_equal(ra[k], rb[k]), k
