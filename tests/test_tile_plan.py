"""The host half of the fused path -- the tile plan lfmgpu_create builds (tile cuts, halo lists, tile-ordered face tables,
local gather lists, shared-memory strides) -- verified on the CPU through lfmgpu_plan_check (include/lfmgpu.h), which runs
the very function the GPU path uses and checks its output against the flattened rank.  No device, no kernel."""
import os
import sys

import pytest

import common
from common import CASES
from lfm_public_b200 import gpu_api, host_api
from lfm_public_b200.tools import casegen, meshgen

sys.path.insert(0, common.ROOT)


@pytest.mark.parametrize("tile_cells", [16, 48, 128])
@pytest.mark.parametrize("name", sorted(CASES))
def test_plan_of_every_case(name, tile_cells, tmp_path):
    case_dir = str(tmp_path / name)
    m, o = common.build_case(name, case_dir)
    for c in common.open_ranks(case_dir, o):
        st = gpu_api.plan_check(c, tile_cells)
        assert st["tileable"] and st["n_tiles"] >= 1
        assert st["tile_cells"] <= max(16, tile_cells)
        assert st["smem_bytes"] <= 75 * 1024
        # every cell is staged at least once, halo cells on top
        assert st["halo_cell_ratio"] >= 0.0 and st["max_staged"] >= 1


def test_plan_fp32_and_small_budget(tmp_path):
    case_dir = str(tmp_path / "c")
    m, o = common.build_case("hex3d_m2_p4", case_dir, doublePrecision=False)
    for c in common.open_ranks(case_dir, o):
        assert gpu_api.plan_check(c)["tileable"]
    d = str(tmp_path / "box")
    casegen.write_case(d, meshgen.hex_box(24, 16, 12, lengths=(3.0, 2.0, 1.5), z_cyclic=True), solver=1, dimension=3, deltaT=1e-3, endTime=1e-2)
    c = host_api.Case.open(d).finish()
    full = gpu_api.plan_check(c, 128)
    st = gpu_api.plan_check(c, 128, 24 * 1024)          # a budget that forces shorter tiles (cuts where the caps are reached)
    assert full["tileable"] and full["tile_cells"] == 128
    assert st["tileable"] and st["smem_bytes"] <= 24 * 1024 and st["n_tiles"] > full["n_tiles"] and st["max_staged"] < full["max_staged"]
    assert not gpu_api.plan_check(c, 128, 1024)["tileable"]          # nothing fits: the unfused kernels serve the rank


def test_plan_of_degenerate_meshes(tmp_path):
    for tag, mesh, two_d, kw in [
        ("single", meshgen.hex_box(1, 1, 1, lengths=(1, 1, 0.1), two_d=True), True, dict(solver=1, dimension=2)),
        ("allbnd", meshgen.hex_box(3, 2, 1, lengths=(3, 2, 0.1), two_d=True), True, dict(solver=0, dimension=2)),
        ("hex222", meshgen.hex_box(2, 2, 2, lengths=(1, 1, 1), z_cyclic=False), False, dict(solver=1, dimension=3)),
    ]:
        d = str(tmp_path / tag)
        casegen.write_case(d, mesh, two_d=two_d, deltaT=1e-3, endTime=1e-2, **kw)
        c = host_api.Case.open(d).finish()
        st = gpu_api.plan_check(c)
        assert st["tileable"] and st["n_tiles"] == 1 and st["halo_face_ratio"] == 0.0


def test_plan_of_the_bench_numbering():
    """Brick-numbered hexahedra (the bench workload, here 128^3 = 2.1 M cells): every invariant holds, the tiles stay inside the
    compile-time strides of the kernels, and the plan is the one the B200 runs reported (tile count, shared memory and
    incoming/own face ratio printed by tools/tune.py in profiles/r2_final_tune128_bricks.log)."""
    import json

    import bench
    case, dt = bench.build_rank_case(128, (1, 1, 1), 0, 1, 8, 1, (8, 4, 4), "morton")
    case.finish()
    st = gpu_api.plan_check(case)
    assert st["tileable"] and st["max_staged"] <= 320 and st["max_faces"] <= 480
    log = open(os.path.join(common.ROOT, "profiles", "r2_final_tune128_bricks.log")).read().splitlines()
    on_gpu = [json.loads(l)["tiles"] for l in log if l.startswith("{")][-1]
    assert st["n_tiles"] == on_gpu["n_tiles"] and st["smem_bytes"] == on_gpu["smem_bytes"] and st["tile_cells"] == on_gpu["tile_cells"]
    assert st["halo_face_ratio"] == on_gpu["halo_face_ratio"]
