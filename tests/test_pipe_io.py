"""Pipelined host I/O (lfmgpu_pipe_*): results equal the plain upload -> step -> download sequence, batch by batch."""
import numpy as np
import pytest

import common
from lfm_public_b200 import gpu_api

pytestmark = pytest.mark.gpu


def test_pipelined_batches_match_sequential(tmp_path):
    m, o = common.build_case("hex3d_m2", str(tmp_path / "c"))
    case = common.open_ranks(str(tmp_path / "c"), o)[0]
    g = gpu_api.GpuSolver(case, 0)
    NQ, n = g.NQ, g.n_cells
    rng = np.random.default_rng(3)
    q0 = np.ascontiguousarray(g.download(0).T)                      # [NQ][n]
    batches = [q0 * (1.0 + 1e-3 * rng.standard_normal(q0.shape)) for _ in range(4)]
    # sequential reference
    want = []
    out = gpu_api.PinnedArray((NQ, n), np.float64)
    inp = gpu_api.PinnedArray((NQ, n), np.float64)
    for b in batches:
        inp.array[...] = b
        g.upload_q_soa_async(inp.ptr, inp.nbytes)
        g.step(o["solver"], o["deltaT"], 1)
        g.download_q_soa_async(out.ptr, out.nbytes)
        g.sync()
        want.append(out.array.copy())
    # pipelined
    ins = [gpu_api.PinnedArray((NQ, n), np.float64) for _ in batches]
    outs = [gpu_api.PinnedArray((NQ, n), np.float64) for _ in batches]
    for a, b in zip(ins, batches):
        a.array[...] = b
    g.pipe_in_start(ins[0].ptr, ins[0].nbytes)
    for k in range(len(batches)):
        g.pipe_in_commit()
        if k + 1 < len(batches):
            g.pipe_in_start(ins[k + 1].ptr, ins[k + 1].nbytes)
        g.step(o["solver"], o["deltaT"], 1)
        g.pipe_out_start()
        g.pipe_out_fetch(outs[k].ptr, outs[k].nbytes)
    g.sync()
    for k in range(len(batches)):
        assert np.array_equal(outs[k].array, want[k]), f"batch {k}"
    assert not np.array_equal(want[0], want[1])
    g.close()
