set -x
mkdir -p gpurun_out
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --set LFMGPU_STAGE_CFG=6,13,14,15 > gpurun_out/r6_tune_a.log 2>&1; cat gpurun_out/r6_tune_a.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --precision 4 --set LFMGPU_STAGE_CFG=0,6,1,13 > gpurun_out/r6_tune_fp32.log 2>&1; cat gpurun_out/r6_tune_fp32.log
timeout 900 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --scheme 0 --set LFMGPU_STAGE_CFG=6,0 > gpurun_out/r6_tune_m1.log 2>&1; cat gpurun_out/r6_tune_m1.log
