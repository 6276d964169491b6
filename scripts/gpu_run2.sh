set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -15 gpurun_out/r2_pytest_gpu.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --set LFMGPU_STAGE_CFG=0,1,2,6,7,8,9 > gpurun_out/r2_tune.log 2>&1; cat gpurun_out/r2_tune.log
timeout 600 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --precision 4 > gpurun_out/r2_tune_fp32.log 2>&1; cat gpurun_out/r2_tune_fp32.log
