#!/bin/bash
# time-boxed checks of the final code: fp32 at scale, graph replay, example case, then the timings
TAG=${1:-c}
mkdir -p gpurun_out
echo "== fp32 128^3"; LFMGPU_PLAN_STATS=1 timeout 120 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --precision 4 --tile morton 2>&1 | grep -v "plan\]" | tail -3 | cut -c1-330
echo "== graph + examples + bench contract"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_examples_gpu.py tests/test_bench_contract.py -m gpu -x -q -k "graph or example or cuda_arm" > gpurun_out/${TAG}_new_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/${TAG}_new_tests.log | cut -c1-300
echo "== fp64 128^3 morton, graph on/off"; timeout 150 python -m lfm_public_b200.tools.tune --n 128 --steps 8 --tile morton --set LFMGPU_GRAPH=1,0 2>&1 | tail -2 | cut -c1-260
echo "== 2D quads 725x725 (C1-sized), graph on/off"; timeout 150 python - <<'PY' 2>&1 | tail -4
import sys, time
sys.path.insert(0, '.')
import os, numpy as np
from lfm_public_b200 import host_api, gpu_api
from lfm_public_b200.tools import meshgen, casegen
import tempfile
d = tempfile.mkdtemp()
casegen.write_case(d, meshgen.hex_box(725, 725, 1, lengths=(7.25, 7.25, 0.01), two_d=True), two_d=True, solver=0, dimension=2, deltaT=1e-4, endTime=1e-3, Ls=1.0)
case = host_api.Case.open(d).finish()
for gmode in ("1", "0"):
    os.environ["LFMGPU_GRAPH"] = gmode
    g = gpu_api.GpuSolver(case, 0)
    g.warmup(); g.step(0, 1e-4, 5); g.sync()
    g.event_record(0); g.step(0, 1e-4, 41); g.event_record(1); g.sync()
    ms = g.event_elapsed_ms(0, 1) / 41
    print(f"quads {g.n_cells} cells M1 fp64 graph={gmode}: {ms:.3f} ms/step = {g.n_cells*5/ms/1e6:.2f} G cell-updates/s")
    g.close()
PY
