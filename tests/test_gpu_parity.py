"""Parity tests proper: the CUDA path (through the C ABI, include/lfmgpu.h) against the CPU oracle on the same
flattened rank, and against the committed golden vectors of the reference's own CPU solver.

Bars (BASELINE.json north_star): fp64 fields within 1e-12 relative max-norm (the default --fmad=false build is
expected to be BIT-EXACT, and the tests below assert equality), fp32 within 1e-5, halo packing bit-exact.
"""
import os

import numpy as np
import pytest

import common
import oracle_lib
from common import CASES, N_STEPS
from lfm_public_b200 import defs, gpu_api

pytestmark = pytest.mark.gpu

MODES = [0, 1]          # use_tiles: 0 = face kernel + gather kernels, 1 = fused shared-memory tiles


def _setup(name, tmp_path, use_tiles, sp=False, **override):
    if not use_tiles and CASES[name]["opts"].get("simulationType", "laminar") != "laminar":
        pytest.skip("the Smagorinsky closure is served by the tile kernels only")
    case_dir = str(tmp_path / name)
    m, o = common.build_case(name, case_dir, doublePrecision=not sp, **override)
    cases = common.open_ranks(case_dir, o)
    oracles = [oracle_lib.Oracle(c) for c in cases]
    gpus = [gpu_api.GpuSolver(c, 0) for c in cases]
    for g in gpus:
        g.set_option("use_tiles", use_tiles)
    if len(gpus) > 1:
        gpu_api.init_local(gpus)
    return o, cases, oracles, gpus


def _close(gpus):
    for g in gpus:
        g.close()


@pytest.mark.parametrize("use_tiles", MODES)
@pytest.mark.parametrize("name", sorted(CASES))
def test_fields_bit_exact_fp64(name, use_tiles, tmp_path):
    o, cases, oracles, gpus = _setup(name, tmp_path, use_tiles)
    res_o = oracle_lib.run(oracles, o["solver"], o["deltaT"], N_STEPS, want_res=True)
    res_g = []
    for s in range(N_STEPS):
        gpu_api.step_multi(gpus, o["solver"], o["deltaT"], 1, first=(s == 0), want_res=True)
        res_g.append(sum(g.residual() for g in gpus))
    for r, (orc, g) in enumerate(zip(oracles, gpus)):
        for field in (defs.FIELD_Q, defs.FIELD_DQ, defs.FIELD_DUDX, defs.FIELD_DTDX, defs.FIELD_RES, defs.FIELD_TAUMC,
                      defs.FIELD_SIGMAU, defs.FIELD_QGHOST):
            a, b = g.download(field), orc.download(field)
            assert a.shape == b.shape and a.size > 0 or field == defs.FIELD_QGHOST
            if field == defs.FIELD_RES:
                # RES is only gathered at stage 0 of the last step on both sides
                pass
            assert np.array_equal(a, b), f"{name} rank {r} field {field}: rel max {common.rel_max(a, b):.3e}"
    # residual norms: parallel tree sum vs the CPU's sequential sum
    assert np.allclose(np.array(res_g), res_o, rtol=1e-12, atol=0)
    # golden vectors of the reference binary
    gpath = os.path.join(common.GOLDEN_DIR, name + ".npz")
    if os.path.exists(gpath):
        gold = np.load(gpath)
        for r, (c, g) in enumerate(zip(cases, gpus)):
            mine = common.primitives_from_q(c.to_mesh_order(g.download(defs.FIELD_Q)), c.desc.c.gamma_m1)
            for k in ("rho", "U", "E", "p"):
                assert common.rel_max(mine[k], gold[f"r{r}_{k}"]) <= 1e-12
                assert np.array_equal(mine[k], gold[f"r{r}_{k}"]), f"{name} rank {r} {k} vs golden"
    _close(gpus)


@pytest.mark.parametrize("use_tiles", MODES)
@pytest.mark.parametrize("name", [n for n in sorted(CASES) if CASES[n]["blocks"]])
def test_halo_packing_bit_exact(name, use_tiles, tmp_path):
    """Per-virtual API driven in the Mesh::solve order on all ranks; every packed send buffer equals the oracle's."""
    o, cases, oracles, gpus = _setup(name, tmp_path, use_tiles)
    scheme, dt = o["solver"], o["deltaT"]
    n = len(gpus)
    rk_order = cases[0].desc.c.rk_order
    record_o = {}
    oracle_lib.lockstep_run(oracles, cases, scheme, dt, 2, record=record_o)
    arrs = [c.arrays() for c in cases]
    D = gpus[0].D
    comm_type = cases[0].desc.c.comm_type
    record_g = {}
    les = not cases[0].opts.laminar           # Mesh::solve picks calc_VIS_Smagorinsky when the case is not laminar

    grads = scheme == defs.SCHEME_M2AUSM and bool(cases[0].opts.minmod)   # mesh_solver.cpp:537-548, 582-594

    def vis(g, sub):
        if grads:
            g.calc_gradients_M2AUSM(sub)
        (g.calc_VIS_Smagorinsky if les else g.calc_VIS)(sub)

    def grab(step):
        spc = oracle_lib.scalars_per_cell(D, comm_type, step)
        for r, g in enumerate(gpus):
            buf = g.send_buffer(step)
            a = arrs[r]
            for i, nb in enumerate(a["nbr_rank"]):
                record_g.setdefault((r, int(nb)), []).append(buf[a["send_start"][i] * spc:a["send_start"][i + 1] * spc].copy())

    for g in gpus:
        g.mpi_communication(0)
    grab(0)
    for g in gpus:
        g.set_boundary_conditions()
        g.mpi_wait(0)
    for g in gpus:
        g.mpi_communication(1)
    grab(1)
    for g in gpus:
        g.mpi_wait(1)
        (g.calc_VIS_Smagorinsky if les else g.calc_VIS)(0)   # warm-up: calc_VIS only (mesh_solver.cpp:409-428)
    for _ in range(2):
        for g in gpus:
            g.prepare_for_timestep()
        for rk in range(rk_order):
            for g in gpus:
                g.prepare_for_RKstep(rk)
                g.mpi_wait(0)
                g.set_boundary_conditions()
                vis(g, 0)
            for g in gpus:
                g.mpi_communication(1)
            grab(1)
            for g in gpus:
                for s in range(1, g.n_sub):
                    vis(g, s)
                g.mpi_wait(1)
                g.one_rk_step(0, scheme, rk, dt)
            for g in gpus:
                g.mpi_communication(0)
            grab(0)
            for g in gpus:
                for s in range(1, g.n_sub):
                    g.one_rk_step(s, scheme, rk, dt)
        for g in gpus:
            g.mpi_wait(0)
    assert record_g.keys() == record_o.keys() and record_o
    for key in record_o:
        assert len(record_g[key]) == len(record_o[key])
        for i, (a, b) in enumerate(zip(record_g[key], record_o[key])):
            assert a.tobytes() == b.tobytes(), f"{name} {key} message {i}"
    for orc, g in zip(oracles, gpus):
        assert np.array_equal(g.download(defs.FIELD_Q), orc.download(defs.FIELD_Q))
        assert np.array_equal(g.download(defs.FIELD_QGHOST), orc.download(defs.FIELD_QGHOST))
    _close(gpus)


@pytest.mark.parametrize("use_tiles", MODES)
@pytest.mark.parametrize("name", ["quad2d_m1", "hex3d_m2_p4", "tri2d_m2", "ogrid3d_m2", "hex3d_m1_p8", "hex3d_m2_les_p4", "ogrid3d_m1_les",
                                  "hex3d_ausm_p4", "ogrid2d_ausm"])
def test_fields_fp32(name, use_tiles, tmp_path):
    """fp32: 1e-5 relative max-norm against the float oracle (and the float reference's golden fields)."""
    o, cases, oracles, gpus = _setup(name, tmp_path, use_tiles, sp=True)
    oracle_lib.run(oracles, o["solver"], o["deltaT"], N_STEPS)
    gpu_api.step_multi(gpus, o["solver"], o["deltaT"], N_STEPS, first=True)
    D = gpus[0].D
    for r, (c, orc, g) in enumerate(zip(cases, oracles, gpus)):
        a, b = g.download(defs.FIELD_Q).astype(np.float64), orc.download(defs.FIELD_Q).astype(np.float64)
        for k in range(D + 2):
            assert common.rel_max(a[:, k], b[:, k]) <= 1e-5, f"{name} rank {r} q[{k}]"
    gpath = os.path.join(common.GOLDEN_DIR, name + "_sp.npz")
    if os.path.exists(gpath):
        gold = np.load(gpath)
        for r, (c, g) in enumerate(zip(cases, gpus)):
            q = c.to_mesh_order(g.download(defs.FIELD_Q)).astype(np.float64)
            assert common.rel_max(q[:, 0], gold[f"r{r}_rho"]) <= 1e-5
            assert common.rel_max(q[:, 1:D + 1] / q[:, :1], gold[f"r{r}_U"]) <= 1e-5
            assert common.rel_max(q[:, D + 1] / q[:, 0], gold[f"r{r}_E"]) <= 1e-5
    _close(gpus)


@pytest.mark.parametrize("name", ["quad2d_m1", "ogrid3d_m2", "ogrid2d_m1", "tri2d_m2"])
def test_cfl_dt_average_forces(name, tmp_path):
    o, cases, oracles, gpus = _setup(name, tmp_path, 1)
    orc, g, c = oracles[0], gpus[0], cases[0]
    dt = o["deltaT"]
    patches = sorted(set(int(p) for p, k in zip(c.arrays()["bc_patch"], c.arrays()["bc_kind"]) if k == defs.BC_WALL))
    for s in range(3):
        oracle_lib.run([orc], o["solver"], dt, 1, first=(s == 0))
        gpu_api.step_multi([g], o["solver"], dt, 1, first=(s == 0))
        assert g.compute_cfl(dt) == orc.cfl(dt)
        assert g.compute_dt(0.5) == orc.dt(0.5)
        orc.average(s + 1)
        g.postProcAverage(s + 1)
        for p in patches:
            fo, vo = np.zeros(3), np.zeros(3)
            orc.forces(p, fo.ctypes.data, vo.ctypes.data)
            fg, vg = g.postProcForces(p)
            assert np.array_equal(fg, fo[:g.D]) and np.array_equal(vg, vo[:g.D]) and np.abs(fo).sum() > 0
    assert np.array_equal(g.download(defs.FIELD_PAVG), orc.download(defs.FIELD_PAVG))
    assert np.array_equal(g.download(defs.FIELD_PRMS), orc.download(defs.FIELD_PRMS))
    _close(gpus)


def test_upload_download_roundtrip(tmp_path):
    o, cases, oracles, gpus = _setup("hex3d_m2", tmp_path, 1)
    g = gpus[0]
    q = g.download(defs.FIELD_Q)
    rng = np.random.default_rng(3)
    q2 = q * (1.0 + 1e-3 * rng.standard_normal(q.shape))
    g.upload_q(q2)
    assert np.array_equal(g.download(defs.FIELD_Q), q2)
    oracles[0].upload_q(q2.ctypes.data)
    oracle_lib.run(oracles, o["solver"], o["deltaT"], 2)
    gpu_api.step_multi(gpus, o["solver"], o["deltaT"], 2, first=True)
    assert np.array_equal(g.download(defs.FIELD_Q), oracles[0].download(defs.FIELD_Q))
    assert g.launch_count > 0
    _close(gpus)


def test_medium_mesh_all_paths(tmp_path):
    """A mesh large enough for many tiles / blocks (64x48x40 hexes = 122 880 cells), 3 steps, the three schemes
    (solver 2 with its minmod gradients), both kernel paths, against the oracle bit for bit."""
    from lfm_public_b200.tools import casegen, meshgen
    m = meshgen.hex_box(64, 48, 40, lengths=(4.0, 3.0, 2.5), z_cyclic=True)
    for scheme in (0, 1, 2):
        case_dir = str(tmp_path / f"med{scheme}")
        opts = casegen.write_case(case_dir, m, solver=scheme, dimension=3, deltaT=2e-3, endTime=1.0, Ls=0.8, mu=7.17948717948718e-05,
                                  minmodExists=(scheme == 2))
        from lfm_public_b200 import host_api
        case = host_api.Case.open(case_dir).finish()
        orc = oracle_lib.Oracle(case)
        oracle_lib.run([orc], scheme, 2e-3, 3)
        ref = orc.download(defs.FIELD_Q)
        for use_tiles in MODES:
            g = gpu_api.GpuSolver(case, 0)
            g.set_option("use_tiles", use_tiles)
            gpu_api.step_multi([g], scheme, 2e-3, 3, first=True)
            assert np.array_equal(g.download(defs.FIELD_Q), ref), f"scheme {scheme} tiles {use_tiles}"
            g.close()


@pytest.mark.parametrize("name", ["quad2d_m1", "ogrid3d_m2", "tri2d_m2"])
def test_graph_replay_bit_exact(name, tmp_path, monkeypatch):
    """lfmgpu_step on a rank without neighbours replays pairs of time steps from a CUDA graph (LFMGPU_GRAPH, default on): nine
    steps in one call (one eager + four replayed pairs) must leave exactly the fields of the oracle and of the eager launches,
    and a second call with another dt must re-capture rather than replay the stale graph."""
    n_steps = 9
    fields = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("LFMGPU_GRAPH", mode)
        o, cases, oracles, gpus = _setup(name, tmp_path / mode, 1)
        assert len(gpus) == 1
        g, orc = gpus[0], oracles[0]
        g.warmup()
        l0 = g.launch_count
        g.step(o["solver"], o["deltaT"], n_steps)
        launches = g.launch_count - l0
        g.step(o["solver"], 0.5 * o["deltaT"], 4)
        g.sync()
        if mode == "1":
            oracle_lib.run(oracles, o["solver"], o["deltaT"], n_steps, first=True)
            oracle_lib.run(oracles, o["solver"], 0.5 * o["deltaT"], 4, first=False)
            for field in (defs.FIELD_Q, defs.FIELD_DQ, defs.FIELD_DUDX, defs.FIELD_QGHOST):
                assert np.array_equal(g.download(field), orc.download(field)), f"{name}: graph replay differs from the oracle in field {field}"
        fields[mode] = (g.download(defs.FIELD_Q), launches)
        _close(gpus)
    assert np.array_equal(fields["1"][0], fields["0"][0])
    assert fields["1"][1] == fields["0"][1] > 0          # replayed kernels are counted like launched ones
