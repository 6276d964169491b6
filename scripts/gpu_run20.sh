set -x
mkdir -p gpurun_out
timeout 600 python -m lfm_public_b200.tools.tune --n 128 --steps 3 > gpurun_out/r20_base.log 2>&1; cat gpurun_out/r20_base.log
LFMGPU_LIB=$PWD/build/liblfmgpu_exp1.so timeout 600 python -m lfm_public_b200.tools.tune --n 128 --steps 3 > gpurun_out/r20_exp1.log 2>&1; cat gpurun_out/r20_exp1.log
