#!/usr/bin/env python
"""CPU model behind DESIGN.md 9.1: what fusing the gradient pass into the stage kernel (two-ring halo of q, gradients of
tile + ring 1 recomputed in shared memory) would move and execute per cell-stage, from the bench mesh's own connectivity
(morton numbering, tiles of 128 consecutive cells) and the measured per-kernel figures of profiles/r2b_final_ncu256_summary.txt.

    python scripts/fused_gradient_model.py [n]        # n^3 cells, default 128
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    case, dt = bench.build_rank_case(n, (1, 1, 1), 0, 1, 8, 1, "morton", "morton")
    case.finish()
    a = case.arrays()
    nc = case.desc.n_cells
    own, nei = a["face_owner"].astype(np.int64), a["face_neigh"].astype(np.int64)
    m = nei < nc
    o, nb = own[m], nei[m]
    A = sp.coo_matrix((np.ones(2 * len(o), dtype=np.int8), (np.r_[o, nb], np.r_[nb, o])), shape=(nc, nc)).tocsr()
    TC = 128
    nt = (nc + TC - 1) // TC
    r1s, r2s = [], []
    for t in np.random.default_rng(0).choice(nt, min(nt, 400), replace=False):
        cells = np.arange(t * TC, min((t + 1) * TC, nc))
        s0 = set(cells.tolist())
        n1 = set(A[cells].indices.tolist()) - s0
        n2 = set(A[np.array(sorted(s0 | n1))].indices.tolist()) - s0 - n1
        r1s.append(len(n1))
        r2s.append(len(n2))
    r1, r2 = float(np.median(r1s)), float(np.median(r2s))
    print(f"{n}^3 cells, morton, tiles of {TC}: ring-1 halo median {r1:.0f} (p90 {np.percentile(r1s, 90):.0f}), ring-2 halo median {r2:.0f} (p90 {np.percentile(r2s, 90):.0f})")
    QB, VB, FACE_G, CELL_G = 64, 128, 36, 20          # Q / V record; S[3] + w + idx per face; gather list + volume per cell (bytes)
    faces_per_cell = 6
    # measured today (256^3, per cell-stage): DRAM bytes and warp instructions of the two kernels
    stage_B, grad_B = 11.72e9 / 16777216, 6.02e9 / 16777216
    stage_I, grad_I = 1782e6 / 16777216, 549e6 / 16777216
    saved = VB + VB + 30 + (grad_B - VB)               # V write, V own read, V halo misses, everything the gradient kernel reads
    ring2_q = r2 / TC * QB                             # gathered, neighbours' tiles: L2 hits for the most part
    ring1_geo = r1 / TC * (faces_per_cell * FACE_G + CELL_G)   # face constants + gather lists of the ring-1 cells (other tiles' tables)
    for l2_hit in (0.4, 0.7):
        fused = stage_B + grad_B - saved + (1 - l2_hit) * (ring2_q + ring1_geo)
        print(f"  L2 hit rate {l2_hit:.0%} on the extra reads: {stage_B + grad_B:.0f} B -> {fused:.0f} B per cell-stage ({fused / (stage_B + grad_B) - 1:+.0%}); "
              f"extra L2->SM traffic {ring2_q + ring1_geo:.0f} B")
    grad_evals = (TC + r1) / TC
    fused_I = stage_I + grad_evals * grad_I * 0.8      # 0.8: the gradient kernel's instructions that are its consumers'
    print(f"  gradient evaluations per updated cell {grad_evals:.2f}; warp instructions per cell-stage {stage_I + grad_I:.0f} -> {fused_I:.0f} ({fused_I / (stage_I + grad_I) - 1:+.0%})")
    slot = (TC + r1 + r2) * QB + (TC + r1) * VB
    print(f"  ring slot: {(TC + r1) * (QB + VB) / 1024:.0f} KB today -> {slot / 1024:.0f} KB (+ the ring-1 face constants): 2 slots instead of 3")
    t_today = 2.711 + 1.312
    print(f"  time per stage today {t_today:.2f} ms; fused if the stage kernel stays issue-bound: {2.711 * fused_I / stage_I:.2f} ms; if it became byte-bound at today's rate: "
          f"{t_today * (stage_B + grad_B - saved + 0.6 * (ring2_q + ring1_geo)) / (stage_B + grad_B):.2f} ms")


if __name__ == "__main__":
    main()
