set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r15_pytest_gpu.log 2>&1; tail -4 gpurun_out/r15_pytest_gpu.log
timeout 900 python -m lfm_public_b200.tools.tune --n 256 --steps 3 --set LFMGPU_STAGE_CFG=6,20 > gpurun_out/r15_tune_256.log 2>&1; cat gpurun_out/r15_tune_256.log
timeout 900 python -m lfm_public_b200.tools.tune --n 256 --steps 3 --precision 4 --set LFMGPU_STAGE_CFG=0,23 > gpurun_out/r15_tune_256_fp32.log 2>&1; cat gpurun_out/r15_tune_256_fp32.log
