#!/bin/bash
# full ncu capture (source-level) of one launch of each persistent kernel on the 128^3 bench block
TAG=${1:-n}
N=${2:-128}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_pipe -s 10 -c 1 -o gpurun_out/${TAG}_stage${N} python -m lfm_public_b200.tools.tune --n $N --steps 1 > gpurun_out/${TAG}_ncu_stage.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_stage.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grad_pipe -s 10 -c 1 -o gpurun_out/${TAG}_grad${N} python -m lfm_public_b200.tools.tune --n $N --steps 1 > gpurun_out/${TAG}_ncu_grad.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_grad.log
