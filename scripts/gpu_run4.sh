set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 4 -c 2 -o gpurun_out/r4_prof python -m lfm_public_b200.tools.tune --n 128 --steps 1 > gpurun_out/r4_ncu.log 2>&1; tail -2 gpurun_out/r4_ncu.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --set LFMGPU_STAGE_CFG=6,10,11 > gpurun_out/r4_tune.log 2>&1; cat gpurun_out/r4_tune.log
