#!/usr/bin/env python
"""Kernel tuning sweep: builds the bench workload once and times the per-kernel device time of a few steps for
every combination of the create-time knobs given on the command line (environment variables read by lfmgpu_create).

    python -m lfm_public_b200.tools.tune --n 256 --set LFMGPU_TILE_CELLS=128,256 --set LFMGPU_STAGE_CFG=0,1,2
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--precision", type=int, default=8)
    ap.add_argument("--scheme", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--tile", default="8,4,4")
    ap.add_argument("--brick-order", default="morton")
    ap.add_argument("--minmod", action="store_true", help="minmodExists: calc_gradients_M2AUSM in the time loop (with --scheme 2)")
    ap.add_argument("--les", action="store_true", help="Smagorinsky closure (calc_VIS_Smagorinsky) instead of laminar")
    ap.add_argument("--set", action="append", default=[], help="NAME=v1,v2,... (environment knob and its values)")
    args = ap.parse_args()
    import bench
    from lfm_public_b200 import gpu_api
    tile = "morton" if args.tile == "morton" else (tuple(int(x) for x in args.tile.split(",")) if args.tile != "none" else None)
    t0 = time.time()
    case, dt = bench.build_rank_case(args.n, (1, 1, 1), 0, 1, args.precision, args.scheme, tile, args.brick_order)
    case.finish()
    print(f"# setup {time.time() - t0:.1f} s", flush=True)
    names, values = [], []
    for s in args.set:
        k, v = s.split("=")
        names.append(k)
        values.append(v.split(","))
    for combo in itertools.product(*values) if names else [()]:
        for k, v in zip(names, combo):
            if v == "":
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        try:
            g = gpu_api.GpuSolver(case, 0)
        except Exception as e:   # a knob combination the library rejects
            print(json.dumps({"knobs": dict(zip(names, combo)), "error": str(e)[:200]}), flush=True)
            continue
        if args.les:
            g.set_option("laminar", 0)
        if args.minmod:
            g.minmod = 1
            g.set_option("minmod", 1)
        g.warmup()
        g.step(args.scheme, dt, 2)
        g.sync()
        g.event_record(0)
        g.step(args.scheme, dt, args.steps)
        g.event_record(1)
        ms = g.event_elapsed_ms(0, 1) / args.steps
        g.enable_kernel_timing(True)
        g.step(args.scheme, dt, args.steps)
        g.sync()
        kt = {}
        for name in ("tile_stage", "tile_grad", "k_flux_face", "k_update_cell", "k_grad_cell", "k_grad_ausm"):
            t, nl = g.kernel_time(name)
            if nl:
                kt[name] = round(t / args.steps, 3)
        g.enable_kernel_timing(False)
        print(json.dumps({"knobs": dict(zip(names, combo)), "ms_per_step": round(ms, 3), "kernel_ms_per_step": kt,
                          "Gcell_stages_per_s": round(g.n_cells * 5 / ms / 1e6, 3), "tiles": g.tile_info()}), flush=True)
        g.close()


if __name__ == "__main__":
    main()
