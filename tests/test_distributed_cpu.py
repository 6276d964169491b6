"""N > 1 host path on CPU: two gloo ranks (torch.distributed.run, 127.0.0.1) do the setup-time neighbour exchange of
bench.py's multi-GPU arm and advance the block-decomposed case with the halo messages travelling over gloo; the result
must equal the single-process run of the same two ranks bit for bit."""
import os
import subprocess
import sys

import numpy as np

import common
import oracle_lib
from lfm_public_b200 import host_api

sys.path.insert(0, common.ROOT)
import bench  # noqa: E402


def test_two_rank_gloo_matches_in_process(tmp_path):
    n, steps = 8, 2
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(common.ROOT, "tests", "_dist_worker.py"), str(tmp_path), str(n), str(steps)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    # the same two ranks in this process
    cases = []
    for r in range(2):
        c, dt = bench.build_rank_case(n, bench.BLOCKS[2], r, 2, 8, 1, (4, 4, 4))
        cases.append(c)
    host_api.exchange_in_process(cases)
    oracles = [oracle_lib.Oracle(c) for c in cases]
    oracle_lib.run(oracles, 1, dt, steps)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        a = cases[r].arrays()
        for k in ("face_owner", "face_neigh", "face_S", "face_d", "face_w", "send_cell", "send_start", "recv_start", "nbr_rank", "q0"):
            assert np.array_equal(got[k], a[k]), f"rank {r}: descriptor array {k} differs between gloo and in-process setup"
        assert len(a["nbr_rank"]) == 1 and a["send_start"][-1] == n * n
        q = oracles[r].download(0)
        assert np.isfinite(q).all() and np.array_equal(got["q"], q), f"rank {r}: fields differ"


def test_eight_rank_bench_layout_gloo_matches_in_process(tmp_path):
    """The layout `bench.py --gpus 8` runs (2x2x2 blocks, morton numbering, z periodic: every rank has a processor AND a
    processorCyclic patch towards the same z neighbour, served as one neighbour with both surfaces) set up rank by rank over gloo and advanced with the halo messages over
    gloo == the eight ranks in one process (what bench.py's parity_nccl compares the NCCL run with)."""
    n, steps, world = 8, 2, 8
    port = 31500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(common.ROOT, "tests", "_dist_worker.py"), str(tmp_path), str(n), str(steps), "bench"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    cases = []
    for r in range(world):
        c, dt = bench.build_rank_case(n, bench.BLOCKS[world], r, world, 8, 1, "morton", "morton", z_cyclic=True)
        cases.append(c)
    host_api.exchange_in_process(cases)
    oracles = [oracle_lib.Oracle(c) for c in cases]
    oracle_lib.run(oracles, 1, dt, steps)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        a = cases[r].arrays()
        for k in ("face_owner", "face_neigh", "send_cell", "send_start", "recv_start", "nbr_rank", "q0"):
            assert np.array_equal(got[k], a[k]), f"rank {r}: descriptor array {k} differs between gloo and in-process setup"
        # x, y and z neighbours; the z neighbour is reached twice (through the cut and through the periodic boundary)
        assert len(a["nbr_rank"]) == 3 and a["send_start"][-1] == 4 * n * n
        q = oracles[r].download(0)
        assert np.isfinite(q).all() and np.array_equal(got["q"], q), f"rank {r}: fields differ"
