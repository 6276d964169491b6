set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1a_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1a_pytest_gpu.log
# ncu full with source, 128^3 fp64 M2: grad sub1 + stage sub0 + stage sub1 of the 2nd stage
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_ -s 5 -c 3 -o gpurun_out/r1a_prof python -m lfm_public_b200.tools.tune --n 128 --steps 1 > gpurun_out/r1a_ncu.log 2>&1; tail -2 gpurun_out/r1a_ncu.log
# variants at 128^3
timeout 600 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --set LFMGPU_USE_TILES=1,0 > gpurun_out/r1a_tune_base.log 2>&1; cat gpurun_out/r1a_tune_base.log
LFMGPU_LIB=$PWD/build/liblfmgpu_fma.so timeout 600 python -m lfm_public_b200.tools.tune --n 128 --steps 3 > gpurun_out/r1a_tune_fma.log 2>&1; cat gpurun_out/r1a_tune_fma.log
timeout 600 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --precision 4 > gpurun_out/r1a_tune_fp32.log 2>&1; cat gpurun_out/r1a_tune_fp32.log
