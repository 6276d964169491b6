set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_ -s 5 -c 3 -o gpurun_out/r3_prof python -m lfm_public_b200.tools.tune --n 128 --steps 1 > gpurun_out/r3_ncu.log 2>&1; tail -2 gpurun_out/r3_ncu.log
LFMGPU_PLAN_STATS=1 timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --set LFMGPU_TILE_SMEM=75,112 --set LFMGPU_TILE_CELLS=128,192,256 > gpurun_out/r3_tune.log 2>&1; cat gpurun_out/r3_tune.log
