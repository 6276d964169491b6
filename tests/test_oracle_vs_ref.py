"""Pins the CPU restatement (oracle/lfm_oracle.c) AND the host-side flattening (lfm_public_b200/host/flatten.cpp)
against the reference's own CPU solver: oracle/_ref/lfm_solve_ref is /root/reference/src/*.cpp + info/lfm_solve.cpp
compiled unchanged (oracle/Makefile).  Same case directory in, fields after N steps and every halo message out.

fp64: bit-exact (the restatement follows the reference's loops and expression trees; gcc emits no FMA).
fp32: 1e-5 relative max-norm is the north_star bar; the restatement is also expected to be bit-exact.
"""
import os

import numpy as np
import pytest

import common
import oracle_lib
from common import CASES, N_STEPS
from lfm_public_b200 import host_api

needs_ref = pytest.mark.skipif(not common.have_ref(), reason="oracle/_ref/lfm_solve_ref not built (needs /root/reference)")


def _run_pair(name, tmp_path, sp=False, **override):
    case_dir = str(tmp_path / name)
    m, o = common.build_case(name, case_dir, doublePrecision=not sp, **override)
    common.run_reference(case_dir, o, sp=sp)
    cases = common.open_ranks(case_dir, o)
    oracles = [oracle_lib.Oracle(c) for c in cases]
    return case_dir, o, cases, oracles


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_fields_match_reference_fp64(name, tmp_path):
    case_dir, o, cases, oracles = _run_pair(name, tmp_path)
    res = oracle_lib.run(oracles, o["solver"], o["deltaT"], N_STEPS, want_res=True)
    D = o["dimension"]
    ref = common.read_reference_q(case_dir, o, o["deltaT"] * N_STEPS, D)
    for r, (c, orc) in enumerate(zip(cases, oracles)):
        q = c.to_mesh_order(orc.download(0))
        mine = common.primitives_from_q(q, c.desc.c.gamma_m1)
        assert ref[r]["rho"].std() > 0 and len(ref[r]["rho"]) == c.desc.n_cells      # not a vacuous comparison
        for k in ("rho", "U", "E", "p"):
            assert np.array_equal(mine[k], ref[r][k]), f"{name} rank {r} field {k}: rel max {common.rel_max(mine[k], ref[r][k]):.3e}"
    assert np.isfinite(res).all() and (res > 0).all()


@needs_ref
@pytest.mark.parametrize("name", [n for n in sorted(CASES) if CASES[n]["blocks"]])
def test_halo_messages_bit_exact(name, tmp_path):
    """Every MPI_Isend payload of the reference (pre-loop warm-up + all stages) equals the restatement's pack output."""
    case_dir, o, cases, oracles = _run_pair(name, tmp_path)
    record = {}
    oracle_lib.lockstep_run(oracles, cases, o["solver"], o["deltaT"], N_STEPS, record=record)
    assert record
    n_msgs = 0
    for (src, dst), msgs in record.items():
        ref = common.read_dump(case_dir, src, dst, np.float64)
        assert len(ref) == len(msgs), f"{src}->{dst}: {len(ref)} reference messages vs {len(msgs)}"
        for i, ((tag, a), b) in enumerate(zip(ref, msgs)):
            assert a.shape == b.shape and a.tobytes() == b.tobytes(), f"{name} {src}->{dst} message {i} differs"
            n_msgs += 1
    assert n_msgs == 2 * (1 + N_STEPS * 5) * len(record)
    # and the lockstep (per-function) driver agrees with lfmo_run
    others = [oracle_lib.Oracle(c) for c in cases]
    oracle_lib.run(others, o["solver"], o["deltaT"], N_STEPS)
    for a, b in zip(oracles, others):
        assert np.array_equal(a.download(0), b.download(0))


@needs_ref
@pytest.mark.skipif(not common.have_ref(sp=True), reason="single-precision reference not built")
@pytest.mark.parametrize("name", ["quad2d_m1", "hex3d_m2_p4", "tri2d_m2", "hex3d_ausm_p4", "ogrid2d_ausm", "hex3d_m2_les_p4", "hex3d_m2_pc8"])
def test_fields_match_reference_fp32(name, tmp_path):
    case_dir, o, cases, oracles = _run_pair(name, tmp_path, sp=True)
    oracle_lib.run(oracles, o["solver"], o["deltaT"], N_STEPS)
    D = o["dimension"]
    ref = common.read_reference_q(case_dir, o, o["deltaT"] * N_STEPS, D)
    for r, (c, orc) in enumerate(zip(cases, oracles)):
        q = c.to_mesh_order(orc.download(0)).astype(np.float32)
        assert common.rel_max(q[:, 0], ref[r]["rho"]) <= 1e-5
        assert common.rel_max(q[:, 1:D + 1] / q[:, :1], ref[r]["U"]) <= 1e-5
        assert common.rel_max(q[:, D + 1] / q[:, 0], ref[r]["E"]) <= 1e-5


@needs_ref
def test_cfl_matches_reference_printout(tmp_path):
    """compute_cfl (cfd_v0.cpp:2887) against the value the reference prints each step (10 significant digits)."""
    name = "quad2d_m1"
    case_dir = str(tmp_path / name)
    m, o = common.build_case(name, case_dir)
    out = common.run_reference(case_dir, o)
    cfls = [float(l.split("CFL:")[1].split()[0]) for l in out.splitlines() if "CFL:" in l]
    cases = common.open_ranks(case_dir, o)
    orc = oracle_lib.Oracle(cases[0])
    mine = []
    for s in range(N_STEPS):
        oracle_lib.run([orc], o["solver"], o["deltaT"], 1, first=(s == 0))
        mine.append(orc.cfl(o["deltaT"]))
    assert len(cfls) == N_STEPS
    assert np.allclose(mine, cfls, rtol=2e-10, atol=0)


@pytest.mark.parametrize("name", ["cylinder_vortex", "cylinder_vortex_unstructured"])
def test_oracle_on_the_reference_example_cases(name, tmp_path):
    """The CPU restatement against the reference executable on the reference's own example cases (BASELINE.json configs 1, 2):
    same SHA-256 of rho, U, E, p after three steps as the fixtures scripts/make_example_fixtures.py took from
    oracle/_ref/lfm_solve_ref.  Runs where the case archive exists (tmp_cases/<name>.txz: not committed, see the script)."""
    import hashlib
    import json
    import tarfile
    fpath = os.path.join(common.GOLDEN_DIR, "examples.json")
    arc = os.path.join(common.ROOT, "tmp_cases", name + ".txz")
    if not os.path.exists(fpath) or not os.path.exists(arc):
        pytest.skip("example fixture or case archive absent")
    fx = json.load(open(fpath)).get(name)
    if fx is None:
        pytest.skip("no fixture for this case")
    with tarfile.open(arc, "r:xz") as t:
        t.extractall(tmp_path)
    case = host_api.Case.open(str(tmp_path / name)).finish()
    orc = oracle_lib.Oracle(case)
    oracle_lib.run([orc], fx["solver"], fx["deltaT"], fx["n_steps"], first=True)
    mine = common.primitives_from_q(case.to_mesh_order(orc.download(0)), case.desc.c.gamma_m1)
    for k, want in fx["sha256"].items():
        assert hashlib.sha256(np.ascontiguousarray(mine[k], dtype=np.float64).tobytes()).hexdigest() == want, (name, k)


# Larger than the zoo (whose meshes are a few hundred cells so that every test runs in a blink): the same pinning on meshes where
# every rank has thousands of interior cells, several tiles' worth of halo and all three patch kinds towards a neighbour.
MEDIUM = {
    # 61 440 hexes, 2x2x2 ranks, z periodic: processor + processorCyclic patches, M2 + viscous + sponge (the bench workload's scheme)
    "hex48x40x32_m2_pc8": dict(mesh=lambda: common.meshgen.hex_box(48, 40, 32, lengths=(2.0, 1.5, 1.0), z_cyclic=True), two_d=False, blocks=(2, 2, 2),
                               opts=dict(solver=1, dimension=3, deltaT=1e-3, Ls=0.4, commType=2, mu=7.17948717948718e-05)),
    # 30 720-cell O-grid cylinder in 4 ranks around the cylinder: wall + inlet/outlet + cyclic span, M1, packed payload
    "ogrid3d_32x96x10_m1_p4": dict(mesh=lambda: common.meshgen.ogrid_cylinder(32, 96, 10, r_in=0.5, r_out=6.0, span=1.0, stretch=4.0), two_d=False,
                                   blocks=(2, 2, 1), opts=dict(solver=0, dimension=3, deltaT=5e-4, Ls=2.0, commType=1, mu=7.17948717948718e-05)),
    # 36 000 shuffled triangle prisms, 2D, M2, serial (the unstructured generator has no block decomposition)
    "tri150x120_m2": dict(mesh=lambda: common.meshgen.tri_prism_box(150, 120, lengths=(2.0, 1.6), shuffle_seed=5), two_d=True, blocks=None,
                          opts=dict(solver=1, dimension=2, deltaT=5e-4, Ls=0.5)),
}


@needs_ref
@pytest.mark.parametrize("name", sorted(MEDIUM))
def test_medium_meshes_match_reference_fp64(name, tmp_path):
    case_dir = str(tmp_path / name)
    m, o = common.build_case(name, case_dir, n_steps=3, spec=MEDIUM[name])
    common.run_reference(case_dir, o)
    cases = common.open_ranks(case_dir, o)
    oracles = [oracle_lib.Oracle(c) for c in cases]
    record = {}
    oracle_lib.lockstep_run(oracles, cases, o["solver"], o["deltaT"], 3, record=record)
    D = o["dimension"]
    ref = common.read_reference_q(case_dir, o, o["deltaT"] * 3, D)
    for r, (c, orc) in enumerate(zip(cases, oracles)):
        mine = common.primitives_from_q(c.to_mesh_order(orc.download(0)), c.desc.c.gamma_m1)
        assert len(ref[r]["rho"]) == c.desc.n_cells and ref[r]["rho"].std() > 0
        for k in ("rho", "U", "E", "p"):
            assert np.array_equal(mine[k], ref[r][k]), f"{name} rank {r} field {k}: rel max {common.rel_max(mine[k], ref[r][k]):.3e}"
    for (src, dst), msgs in record.items():
        refm = common.read_dump(case_dir, src, dst, np.float64)
        assert len(refm) == len(msgs)
        for i, ((tag, a), b) in enumerate(zip(refm, msgs)):
            assert a.tobytes() == b.tobytes(), f"{name} {src}->{dst} message {i} differs"
