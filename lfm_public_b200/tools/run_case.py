#!/usr/bin/env python
"""Times the GPU path on a serial LFM/OpenFOAM case directory: python -m lfm_public_b200.tools.run_case <case> [steps]"""
import json
import sys
import time

from lfm_public_b200 import gpu_api, host_api


def main():
    case_dir = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    t0 = time.time()
    case = host_api.Case.open(case_dir).finish()
    o = case.opts
    g = gpu_api.GpuSolver(case, 0)
    g.warmup()
    g.step(o.solver, o.delta_t, 3)
    g.sync()
    g.event_record(0)
    g.step(o.solver, o.delta_t, steps)
    g.event_record(1)
    ms = g.event_elapsed_ms(0, 1) / steps
    g.enable_kernel_timing(True)
    g.step(o.solver, o.delta_t, 2)
    g.sync()
    kt = {}
    for name in ("tile_stage", "tile_grad", "k_flux_face", "k_update_cell", "k_grad_cell"):
        t, nl = g.kernel_time(name)
        if nl:
            kt[name] = round(t / 2, 3)
    print(json.dumps({"case": case_dir, "cells": g.n_cells, "ms_per_step": round(ms, 3), "Gcell_stages_per_s": round(g.n_cells * 5 / ms / 1e6, 3),
                      "kernel_ms_per_step": kt, "tiles": g.tile_info(), "setup_s": round(time.time() - t0, 1)}))
    g.close()


if __name__ == "__main__":
    main()
