set -x
mkdir -p gpurun_out
LFMGPU_STAGE_CFG=30 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r13_pytest_gpu.log 2>&1; tail -5 gpurun_out/r13_pytest_gpu.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --set LFMGPU_STAGE_CFG=6,30,31,32 > gpurun_out/r13_tune_a.log 2>&1; cat gpurun_out/r13_tune_a.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --precision 4 --set LFMGPU_STAGE_CFG=23,30,32 > gpurun_out/r13_tune_fp32.log 2>&1; cat gpurun_out/r13_tune_fp32.log
