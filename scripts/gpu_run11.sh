set -x
mkdir -p gpurun_out
LFMGPU_LIB=$PWD/build/liblfmgpu_fma.so timeout 600 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --set LFMGPU_STAGE_CFG=6,0 > gpurun_out/r11_tune_fma.log 2>&1; cat gpurun_out/r11_tune_fma.log
LFMGPU_LIB=$PWD/build/liblfmgpu_fma.so timeout 600 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --precision 4 > gpurun_out/r11_tune_fma32.log 2>&1; cat gpurun_out/r11_tune_fma32.log
LFMGPU_LIB=$PWD/build/liblfmgpu_fma.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fp32 or medium" > gpurun_out/r11_pytest_fma.log 2>&1; tail -5 gpurun_out/r11_pytest_fma.log
