set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r5_pytest_gpu.log 2>&1; tail -5 gpurun_out/r5_pytest_gpu.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --set LFMGPU_STAGE_CFG=6,1,0 > gpurun_out/r5_tune_a.log 2>&1; cat gpurun_out/r5_tune_a.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --tile 4,4,4 --set LFMGPU_TILE_CELLS=64 --set LFMGPU_STAGE_CFG=1,9,2 > gpurun_out/r5_tune_b.log 2>&1; cat gpurun_out/r5_tune_b.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --tile 8,8,4 --set LFMGPU_TILE_CELLS=256 --set LFMGPU_TILE_SMEM=200 --set LFMGPU_STAGE_CFG=12 > gpurun_out/r5_tune_c.log 2>&1; cat gpurun_out/r5_tune_c.log
