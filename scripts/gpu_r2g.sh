#!/bin/bash
# config 1 (examples/cylinder_vortex, 525 000 quads, M1) on 1 B200: CUDA-graph replay of step pairs on and off.  usage: gpu_r2g.sh <tag>
TAG=${1:-r4g}
mkdir -p gpurun_out /tmp/c1 && tar xJf tmp_cases/cylinder_vortex.txz -C /tmp/c1
LFMGPU_GRAPH=1 timeout 25 python -m lfm_public_b200.tools.run_case /tmp/c1/cylinder_vortex 40 > gpurun_out/${TAG}_c1_graph1.log 2>&1; tail -1 gpurun_out/${TAG}_c1_graph1.log | cut -c1-300
LFMGPU_GRAPH=0 timeout 25 python -m lfm_public_b200.tools.run_case /tmp/c1/cylinder_vortex 40 > gpurun_out/${TAG}_c1_graph0.log 2>&1; tail -1 gpurun_out/${TAG}_c1_graph0.log | cut -c1-300
