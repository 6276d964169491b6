"""Worker of tests/test_distributed_cpu.py: one rank of a 2-rank gloo job; host-side setup exchange over
torch.distributed, then an oracle run with the halo messages exchanged over gloo (the N>1 path of bench.py without a GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import oracle_lib  # noqa: E402
from lfm_public_b200 import host_api  # noqa: E402


def main():
    out_dir, n, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    bench_layout = len(sys.argv) > 4 and sys.argv[4] == "bench"   # the multi-GPU bench's own numbering and z-periodic layout
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    blocks = bench.BLOCKS[world]
    if bench_layout:
        case, dt = bench.build_rank_case(n, blocks, rank, world, 8, 1, "morton", "morton", z_cyclic=True)
    else:
        case, dt = bench.build_rank_case(n, blocks, rank, world, 8, 1, (4, 4, 4))
    host_api.exchange_distributed(case, rank)
    a = case.arrays()
    orc = oracle_lib.Oracle(case)
    real = np.float64
    nbrs = [int(x) for x in a["nbr_rank"]]
    ss, rs = a["send_start"], a["recv_start"]

    def exchange(step):
        spc = oracle_lib.scalars_per_cell(case.desc.dim, case.desc.c.comm_type, step)
        buf = np.ascontiguousarray(orc.pack(step))
        recv = np.zeros(int(rs[-1]) * spc, dtype=real)
        reqs, keep = [], []
        for i, nb in enumerate(nbrs):
            s = torch.from_numpy(buf[int(ss[i]) * spc:int(ss[i + 1]) * spc].copy())
            r = torch.from_numpy(recv[int(rs[i]) * spc:int(rs[i + 1]) * spc])
            keep += [s, r]
            reqs.append(dist.isend(s, nb, tag=step))
            reqs.append(dist.irecv(r, nb, tag=step))
        for q in reqs:
            q.wait()
        orc.unpack(step, recv.ctypes.data)

    oracle_lib.drive_rank(orc, exchange, 1, dt, steps)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), q=orc.download(0), **{k: np.array(v) for k, v in a.items()})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
