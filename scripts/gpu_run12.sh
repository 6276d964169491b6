set -x
mkdir -p gpurun_out
LFMGPU_STAGE_CFG=20 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r12_pytest_gpu.log 2>&1; tail -5 gpurun_out/r12_pytest_gpu.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --set LFMGPU_STAGE_CFG=6,20,21,22 > gpurun_out/r12_tune_a.log 2>&1; cat gpurun_out/r12_tune_a.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --precision 4 --set LFMGPU_STAGE_CFG=0,20,23 > gpurun_out/r12_tune_fp32.log 2>&1; cat gpurun_out/r12_tune_fp32.log
